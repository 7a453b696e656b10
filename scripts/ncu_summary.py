"""Condense an .ncu-rep (ncu --set full) into the handful of counters the roofline discussion uses.
usage: python scripts/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/<name>.txt"""
import csv, subprocess, sys
KEYS = ["gpu__time_duration.sum", "sm__cycles_elapsed.avg", "sm__cycles_active.avg", "sm__cycles_elapsed.avg.per_second",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "sm__warps_active.avg.per_cycle_active", "launch__registers_per_thread",
        "launch__shared_mem_per_block", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "launch__grid_size", "launch__block_size", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__sass_inst_executed_op_shared_ld.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"]
STALLS = "smsp__pcsamp_warps_issue_stalled_"
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
for vals in rows[2:]:
    d = dict(zip(hdr, zip(units, vals)))
    print("kernel:", d.get("Kernel Name", ("", "?"))[1][:100])
    for k in KEYS:
        if k in d:
            print("  %-78s %14s %s" % (k, d[k][1], d[k][0]))
    st = sorted(((float(v[1] or 0), k[len(STALLS):]) for k, v in d.items() if k.startswith(STALLS) and not k.endswith("_not_issued")), reverse=True)
    tot = sum(x for x, _ in st) or 1.
    print("  warp-state samples (pc sampling): " + ", ".join("%s %.1f%%" % (n, 100 * x / tot) for x, n in st[:9]))
