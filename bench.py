#!/usr/bin/env python
"""bench.py -- throughput of the Landau-Poisson hot path on B200 (one JSON line on rank 0).

Metric (BASELINE.json): collision-cell evaluations/s, one evaluation = ComputeQ + conserveMoments
on one spatial cell; the solver does 4 per cell per timestep (LP_ompi.cpp:700-702,
collisionRoutines_1.cpp:919,931,943).  A "step" is one full timestep (SSP-RK3 advection + RK4
collision of every local cell); value = 4 * Nx * steps / time, timesteps/s is reported beside it.

Workload: BASELINE config "two-stream, Nx=256, Nv=32^3 on 8 GPUs" cut to its per-GPU shard: 32
x-cells per GPU, Nv = N = 32 (weak scaling: Nx = 32 * n_gpus).  BASELINE's single-cell homogeneous
config cannot be sharded for the 1->8 sweep; it is measured too (N=1 only) and reported under
"homogeneous_1cell".

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
  torchrun ... bench.py --gpus N ...        (one rank per GPU, NCCL)
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CELLS_PER_GPU = 32
NV = 32
NSPEC = 32
PHYS = dict(Lv=5.25, Lx=4.0, nu=0.05, dt=0.01)     # [TwoStream] section of the reference input deck
A_AMP, K_WAVE = 0.5, 2 * np.pi / 4.
OUT = sys.stdout
METRIC, UNIT = "collision_cell_evals_per_s", "evals/s"


def pairs_per_eval(N):
    return (3 * N * N / 4.) ** 3


def flop_per_eval(N):
    return 10. * pairs_per_eval(N)                  # algorithmic: 6 complex-mul + 4 scale-accumulate per pair


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons while the timed region runs (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag, self.proc = index, [], False, None

    def run(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                self.rows.append([x.strip() for x in line.split(",")])
                if self.stop_flag:
                    break
        except Exception:
            pass

    def finish(self):
        self.stop_flag = True
        if self.proc:
            self.proc.terminate()
        sm, mx, reasons = [], 0., set()
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx = max(mx, float(r[1]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                continue
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons), "samples": len(sm)}


def cpu_eval_rate(reps, warm):
    """ComputeQ + conserveMoments of the reference's own CPU code (oracle/_ref when it was built
    from /root/reference, else the C restatement) on this box's host cores: evals/s on one cell."""
    from oracle.oracle import PortOracle, RefOracle, have_ref
    cfg = dict(Nx=1, Nv=NV, N=NSPEC, homogeneous=True, **PHYS)
    kind, t_init = "port", time.time()
    ora = None
    if have_ref():
        try:
            avail_gb = os.sysconf("SC_AVPHYS_PAGES") * os.sysconf("SC_PAGE_SIZE") / 2 ** 30
            if avail_gb > 8 * NSPEC ** 6 / 2 ** 30 + 4:
                ora = RefOracle(**cfg)              # builds the reference's 8*N^6-byte weight table
                kind = "reference"
        except Exception:
            ora = None
    if ora is None:
        ora = PortOracle(**cfg)
    t_init = time.time() - t_init
    po = PortOracle(**cfg)                            # input data from the oracle side only: nothing of the product on this path
    f = po.setInit_spectral(po.SetInit_4H_Homo())[0]
    for _ in range(warm):
        ora.conserveMoments(ora.ComputeQ(f))
    t = time.time()
    for _ in range(reps):
        ora.conserveMoments(ora.ComputeQ(f))
    dt = (time.time() - t) / reps
    return dict(value=1. / dt, unit=UNIT, cores=int(ora.num_threads), kind=kind,
                sample="%d x (ComputeQ + conserveMoments) on 1 cell, N=%d, OpenMP on %d threads; init %.0f s excluded" % (reps, NSPEC, ora.num_threads, t_init)), dt


def workload_config(world):
    """config of the JSON line, shared by both arms"""
    Nx = CELLS_PER_GPU * world
    return {"workload": "two-stream Landau-Poisson timestep (SSP-RK3 DG advection + RK4 spectral Landau collision), %d x-cells per GPU, Nv=%d, N=%d "
                        "(BASELINE config Nx=256,Nv=32^3 on 8 GPUs, per-GPU shard)" % (CELLS_PER_GPU, NV, NSPEC),
            "Nx": Nx, "Nv": NV, "N": NSPEC, "evals_per_step": 4 * Nx, "parallelism": "x-cells sharded over %d GPU(s)" % world}


def run_reference(args, rank):
    """--impl reference: the reference's CPU implementation of the metric's unit, timed on the host."""
    if rank != 0:
        return
    base, dt = cpu_eval_rate(args.steps, args.warmup)
    line = {"impl": "reference", "metric": METRIC, "value": base["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": dict(workload_config(args.gpus), sample="each step = one ComputeQ + conserveMoments on one cell (the metric's unit) "
                                                              "by the reference's own CPU code on this box's host cores"),
            "cpu_baseline": base, "e2e": {"value": base["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), file=OUT, flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    # stdout carries the one JSON line and nothing else: whatever a library prints there (NCCL's version banner ...)
    # is sent to stderr
    global OUT
    OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    if args.impl == "reference":
        return run_reference(args, rank)
    if args.warmup < 3:
        args.warmup = 3

    import torch
    import __graft_entry__ as graft
    pkg = graft.load_package()
    from lpsolver_b200 import solver
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the hot path has no CPU fallback")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    assert world == args.gpus, "--gpus must equal the number of launched ranks"

    Nx = CELLS_PER_GPU * world
    s = solver.ShardedSolver(Nx, NV, NSPEC, homogeneous=False, rank=rank, world=world, device=local, dist=dist, **PHYS)
    g = s.g
    g.set_stream(torch.cuda.current_stream().cuda_stream)
    U0 = solver.set_init_ld(Nx, NV, PHYS["Lv"], PHYS["Lx"], A_AMP, K_WAVE, True, s.x_begin, s.x_count)
    host = torch.from_numpy(U0).pin_memory()
    host_np = host.numpy()
    back = torch.empty_like(host).pin_memory()
    back_np = back.numpy()

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x):
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    # ---- device-resident throughput ------------------------------------------------------------
    s.upload(host_np)
    s.step(args.warmup)
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    time.sleep(0.3)
    l0 = g.launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    s.step(args.steps)            # one call: the host enqueues step k+1's advection exchange while step k's collisions run
    e1.record()
    barrier()
    t_dev = max_over_ranks(e0.elapsed_time(e1) * 1e-3)
    launches = sum_over_ranks(g.launch_count - l0)
    value = 4. * Nx * args.steps / t_dev
    # instrumented pass for the roofline: the same K steps again with a CUDA-event pair around every launch of the
    # dominant kernel.  It cannot share the timed pass: there a step is a CUDA-graph replay in which the cells run as
    # concurrent chains (api.cu, collide_async), so one kernel's launch has no duration of its own; with the events on,
    # the library launches eagerly, one chain over all cells, and each F2 launch runs alone on the GPU.
    g.profile_computeQ(2)
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e2.record()
    s.step(args.steps)
    e3.record()
    barrier()
    t_instr = max_over_ranks(e2.elapsed_time(e3) * 1e-3)
    cq_ms, cq_n = g.profile_read()
    #     ... and once more with the events around the DG stage kernels of the advection (HBM-bound)
    g.profile_computeQ(3)
    barrier()
    s.step(args.steps)
    barrier()
    dg_ms, dg_n = g.profile_read()
    g.profile_computeQ(False)
    clocks = sampler.finish()

    # ---- end to end: host buffers in, host buffers out, every step --------------------------------
    # (a) serial: one context, upload -> timestep -> download, each call synchronous (what LP_ompi.cpp's loop does
    #     around MPI_Bcast(U)); (b) pipelined, the headline: two contexts on two streams, every step still moves its
    #     whole U from pinned host memory to the device and its whole result back, but the copies of one batch
    #     overlap the kernels of the other (lpgpu_upload_U_async / lpgpu_step_async / lpgpu_download_U_async).
    e2e_steps = max(2, min(args.steps, 5))
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        s.upload(host_np)          # H2D from pinned memory (the reference's MPI_Bcast(U), LP_ompi.cpp:658)
        s.step(1)
        s.download(back_np)        # D2H (the reference's gather for diagnostics/output, LP_ompi.cpp:813-849)
    barrier()
    t_ser = max_over_ranks(time.perf_counter() - t0)
    e2e_serial = {"value": 4. * Nx * e2e_steps / t_ser, "unit": UNIT, "steps": e2e_steps, "timesteps_per_s": e2e_steps / t_ser}

    # (c) as the reference's loop does it: after every timestep the moments, the entropy and the
    #     negativity / KiE-ratio diagnostics (LP_ompi.cpp:817-849) -- GPU reductions, a few doubles back per step
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        s.step(1)
        s.moments()
        s.g.diagnostics_partial()
    barrier()
    t_diag_sync = max_over_ranks(time.perf_counter() - t0)
    #     ... and with the diagnostics of step k running on a side stream over a snapshot while step k+1 runs
    #     (lpgpu_diagnostics_begin / _end): every step's numbers are still produced, one step later
    ar_steps = 2 * e2e_steps
    s.step(1); s.diagnostics_begin(); s.diagnostics_end()      # lazy allocations of the snapshot path
    barrier()
    t0 = time.perf_counter()
    for k in range(ar_steps):
        s.step(1, wait=False)
        if k:
            s.diagnostics_end()
        s.diagnostics_begin()
    diag_last = s.diagnostics_end()
    barrier()
    t_diag = max_over_ranks(time.perf_counter() - t0)
    as_reference = {"value": 4. * Nx * ar_steps / t_diag, "unit": UNIT, "steps": ar_steps, "timesteps_per_s": ar_steps / t_diag,
                    "synchronous": {"value": 4. * Nx * e2e_steps / t_diag_sync, "timesteps_per_s": e2e_steps / t_diag_sync, "steps": e2e_steps},
                    "entropy_last": float(diag_last[1]),
                    "includes": "per-step mass/momentum/energy, entropy, negativity and KiE-ratio diagnostics (host reads ~10 doubles per step); "
                                "the diagnostics of step k run on a side stream over a snapshot while step k+1 runs; 'synchronous' = step, then diagnostics, then the next step"}

    NFLIGHT = 3                                           # copy-in, kernels and copy-out of three batches overlap
    streams = [torch.cuda.Stream() for _ in range(NFLIGHT)]
    # several contexts in flight per rank: their exchanges go through NCCL (one communicator orders them the same way on
    # every rank); flag-synchronised peer writes of independent contexts could wait on each other across hardware queues
    pair = [solver.ShardedSolver(Nx, NV, NSPEC, homogeneous=False, rank=rank, world=world, device=local, dist=dist, stream=st,
                                 exchange="nccl", **PHYS) for st in streams]
    hosts_t = [host] + [host.clone().pin_memory() for _ in range(NFLIGHT - 1)]
    hosts = [h.numpy() for h in hosts_t]
    backs_t = [torch.empty_like(host).pin_memory() for _ in range(NFLIGHT)]
    backs = [b.numpy() for b in backs_t]

    def pipelined(n):
        for k in range(n):
            q = pair[k % NFLIGHT]
            q.synchronize()                               # this context's previous batch has left its host buffers
            q.upload(hosts[k % NFLIGHT], wait=False)
            q.step(1, wait=False)
            q.download(backs[k % NFLIGHT], wait=False)
        for q in pair:
            q.synchronize()

    pipelined(2 * NFLIGHT)                                # warm-up: lazy allocations, function attributes
    pipe_steps = NFLIGHT * max(3, min(args.steps, 10))
    l0p = sum(q.g.launch_count for q in pair)
    barrier()
    t0 = time.perf_counter()
    pipelined(pipe_steps)
    barrier()
    t_e2e = max_over_ranks(time.perf_counter() - t0)
    pipe_launches = sum(q.g.launch_count for q in pair) - l0p
    same = bool(all(np.array_equal(b, back_np) for b in backs))
    for q in pair:
        q.close()
    e2e = {"value": 4. * Nx * pipe_steps / t_e2e, "unit": UNIT, "h2d_bytes_per_step": int(host.numel() * 8 * world),
           "d2h_bytes_per_step": int(back.numel() * 8 * world), "steps": pipe_steps, "timesteps_per_s": pipe_steps / t_e2e,
           "batches_in_flight": NFLIGHT, "gpu_launches": int(pipe_launches), "result_equals_serial": same, "serial": e2e_serial,
           "note": "every step: full U of the batch H2D from pinned memory, one timestep, full U D2H; %d independent batches in flight "
                   "(one context and one stream each) so copies overlap kernels; 'serial' is one context with synchronous calls" % NFLIGHT}

    # ---- roofline of the dominant kernels --------------------------------------------------------
    # (1) the step's dominant kernel: k_fc3_f2_tmem, the y/x-transform + product + inverse kernel of ComputeQ's
    #     FFT-convolution pipeline (~52 % of a step; its launches were bracketed by CUDA events inside the timed
    #     region, profile mode 2).  Bound: the FP64 pipe (ncu: FP64 pipe 45 % busy, DRAM 13 %), so the roofline is
    #     quoted in FP64 TFLOP/s against the DFMA rate measured live on this GPU.  Algorithmic flops of that kernel
    #     (DESIGN.md section 4.1): per cell and per kz plane 14 x (N + M) forward and (M + N) inverse length-M line
    #     transforms at the textbook 5 M log2 M, plus 7 M^2 complex multiply-adds at 8 flop; M = 3N/2 planes.
    # (2) the north star's direct O(N^6) sum (k_computeQ_tiled, computeq_variant = 3), timed here on the same
    #     cells outside the step: FP64-pipe bound, 10 flop per (xi, omega) pair.
    roof = roof_direct = None
    fp64_peak = pkg.lpgpu.fp64_peak_tflops(local)
    peaks = {}
    ppath = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(ppath):
        peaks = json.load(open(ppath))
    traffic = {}
    tpath = os.path.join(ROOT, "profiles", "computeq_traffic.json")
    if os.path.exists(tpath):
        traffic = json.load(open(tpath))
    fp64_src = ("DFMA micro-benchmark run live on this GPU (lpgpu_fp64_peak); MEASURED_PEAKS.json holds only HBM and bf16 peaks; "
                "nominal 148 SM x 64 DFMA/clk x 1.965 GHz = 37.2 TFLOP/s")
    if cq_n > 0:
        Mpad = 3 * NSPEC // 2                      # cyclic transform size per dimension (fc3.cuh)
        line_flop = 5. * Mpad * np.log2(Mpad)
        f2_flop_per_cell = Mpad * ((14 + 1) * (NSPEC + Mpad) * line_flop + 7 * Mpad * Mpad * 8.)
        avg_s = cq_ms * 1e-3 / cq_n
        achieved = f2_flop_per_cell * s.x_count / avg_s / 1e12
        chain_bytes_per_cell = 16 * (NSPEC ** 3 + 2 * 10 * NSPEC * NSPEC * Mpad + 2 * NSPEC * NSPEC * Mpad + NSPEC ** 3)
        roof = {"bound": "fp64", "kernel": "k_fc3_f2_tmem (ComputeQ as FFT convolutions: y/x line transforms + products + inverse x/y of one kz plane per CTA, accumulators in TMEM)",
                "achieved": achieved, "peak": fp64_peak, "unit": "TFLOP/s", "frac": achieved / fp64_peak,
                "traffic": traffic.get("fc3_f2_dram_bytes_per_launch"),
                "avg_launch_ms": avg_s * 1e3, "launches": cq_n, "share_of_step": cq_ms * 1e-3 / t_instr, "measured_in": "instrumented pass: the same %d steps, eager launches, one chain over all cells, CUDA events around every F2 launch (%.3f ms/step); the timed pass replays a CUDA graph with the cells in concurrent chains" % (args.steps, t_instr / args.steps * 1e3),
                "peak_source": fp64_src,
                "algorithmic_flop_per_launch": f2_flop_per_cell * s.x_count,
                "hbm_view": {"algorithmic_bytes_per_launch_whole_chain": chain_bytes_per_cell * s.x_count, "hbm_peak_gbs": peaks.get("hbm_gbs"),
                             "note": "F1+F2+F3 move 10+1 arrays of N^2 M complex once each way; at the measured HBM peak that is ~0.08 ms per launch chain, well under the FP64 time"},
                "direct_form_equivalent_tflops": flop_per_eval(NSPEC) * s.x_count / avg_s / 1e12,
                "note": "bound is the FP64 vector pipe (not hbm/tensor): the contraction is not dense, see DESIGN.md 4.1; direct_form_equivalent_tflops counts the 10 flop/pair of the O(N^6) sum this kernel replaces and exceeds the FP64 peak because the FFT form executes ~50x fewer flops for the same result (parity-tested)"}
    roof_adv = None
    if dg_n > 0:
        # 384 B per DG cell per timestep = 96 (stage 1: 6 read + 6 written doubles) + 144 + 144 (12 read + 6
        # written); the x- and v1-upwind neighbours a DG cell also reads are other cells' own coefficients, served by L1/L2
        adv_bytes_per_launch = 128. * s.x_count * NV ** 3
        avg_dg = dg_ms * 1e-3 / dg_n
        hbm_peak = peaks.get("hbm_gbs") or 6500.
        roof_adv = {"bound": "hbm", "kernel": "k_dg_stage<0..2> (SSP-RK3 stages of the DG upwind advection: I1, I2, I3, I5, H and the stage combination per DG cell)",
                    "achieved": adv_bytes_per_launch / avg_dg / 1e9, "peak": hbm_peak, "unit": "GB/s",
                    "frac": adv_bytes_per_launch / avg_dg / 1e9 / hbm_peak, "traffic": None,
                    "avg_launch_ms": avg_dg * 1e3, "launches": dg_n, "share_of_step": dg_ms * 1e-3 / t_instr,
                    "algorithmic_bytes_per_launch": adv_bytes_per_launch,
                    "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks.get("hbm_gbs") else "fallback 6500 GB/s (B200_PROFILING.md)"}
    try:
        d = pkg.LPGpu(Nx, NV, NSPEC, homogeneous=False, x_begin=s.x_begin, x_count=s.x_count, device=local, computeq_variant=3, **PHYS)
        d.set_stream(torch.cuda.current_stream().cuda_stream)
        d.upload_U(host_np)
        d.sample_device()
        for _ in range(2):
            d.eval_device(s.x_count)
        d.profile_computeQ(True)
        for _ in range(4):
            d.eval_device(s.x_count)
        dq_ms, dq_n = d.profile_read()
        d.close()
        avg_d = dq_ms * 1e-3 / dq_n
        ach = flop_per_eval(NSPEC) * s.x_count / avg_d / 1e12
        roof_direct = {"bound": "fp64", "kernel": "k_computeQ_tiled (computeq_variant=3, the reference's O(N^6) form)", "achieved": ach,
                       "peak": fp64_peak, "unit": "TFLOP/s", "frac": ach / fp64_peak, "traffic": traffic.get("dram_bytes_per_launch"),
                       "avg_launch_ms": avg_d * 1e3, "launches": dq_n, "evals_per_s": s.x_count / avg_d,
                       "peak_source": fp64_src,
                       "algorithmic_flop_per_launch": flop_per_eval(NSPEC) * s.x_count}
    except Exception as e:
        roof_direct = {"error": repr(e)}

    line = None
    if rank == 0:
        ws_mb = (3 * (s.x_count + 2) * 6 * NV ** 3 * 8 + s.x_count * NSPEC ** 3 * 8 * (3 + 2 * 6) + s.x_count * NSPEC * 4 * NV * NV * 16
                 + s.x_count * 11 * NSPEC * NSPEC * (3 * NSPEC // 2) * 16) / 2 ** 20
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": t_dev / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f64", "data": "synthetic",
                "config": dict(workload_config(world), l2="per-GPU working set %.0f MB > 126 MB L2; no flush between steps" % ws_mb,
                               exchange=("none (one GPU)" if world == 1 else
                                         "peer memory: kernels write halo planes and densities into the neighbours' buffers (CUDA IPC), flag-synchronised; no NCCL call in the timestep" if s.exchange == "peer" else
                                         "NCCL all-gather + send/recv per SSP-RK3 stage")),
                "timesteps_per_s": args.steps / t_dev, "roofline": roof, "roofline_direct": roof_direct, "roofline_advection": roof_adv, "e2e": e2e, "as_reference_loop": as_reference,
                "gpu_launches": int(launches), "clocks": clocks}

    # ---- BASELINE's single-cell homogeneous config, and the CPU baseline (N=1 only) --------------
    if world == 1:
        s.close()
        h = solver.ShardedSolver(1, NV, NSPEC, homogeneous=True, device=local, **PHYS)
        h.g.set_stream(torch.cuda.current_stream().cuda_stream)
        h.upload(solver.set_init_4h_homo(NV, PHYS["Lv"]))
        h.step(4)                  # the library replays a CUDA graph of the timestep from the second step on
        torch.cuda.synchronize()
        e0.record()
        nh = 40
        h.step(nh)
        e1.record()
        torch.cuda.synchronize()
        th = e0.elapsed_time(e1) * 1e-3
        line["homogeneous_1cell"] = {"workload": "space-homogeneous collision-only relaxation, 1 cell, Nv=N=32 (FourHump IC)",
                                     "value": 4 * nh / th, "unit": UNIT, "timesteps_per_s": nh / th, "steps": nh}
        h.close()
        if not args.no_cpu_baseline:
            try:
                line["cpu_baseline"], _ = cpu_eval_rate(2, 1)
            except Exception as e:                          # the oracle is a checker; never let it break the GPU number
                line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": os.cpu_count(), "kind": "unavailable", "sample": repr(e)}
    else:
        s.close()
    if rank == 0:
        print(json.dumps(line), file=OUT, flush=True)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
