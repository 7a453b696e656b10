#!/usr/bin/env python
"""bench.py -- throughput of the Landau-Poisson hot path on B200 (one JSON line on rank 0).

Metric (BASELINE.json): collision-cell evaluations/s, one evaluation = ComputeQ + conserveMoments on one spatial
cell; the solver does 4 per cell per timestep (LP_ompi.cpp:700-702, collisionRoutines_1.cpp:919,931,943).  A "step"
is one full timestep (SSP-RK3 advection + RK4 collision of every cell); value = 4 * Nx * steps / time, timesteps/s
is reported beside it.

Workload (default): BASELINE config 5 -- Nx = 512, Nv = N = 32, the "scaling sweep 1/2/4/8" -- run as such at every
GPU count: the 512 x-cells are sharded over the N GPUs (strong scaling; one GPU holds all 512 cells, 11 GB).  The
reference has no bump-on-tail initial condition; the two-stream deck values are used (A = 0.5, Lv = 5.25,
nu = 0.05, dt = 0.01, and the deck's cell size dx = 0.125, i.e. Lx = Nx / 8: see lx_of): throughput does not depend on the data.
  --scaling weak   : BASELINE config 4 cut to its per-GPU shard, 32 x-cells per GPU (Nx = 256 on 8 GPUs).
Also measured on one GPU and reported in the same line: config 3 (Landau damping, Nx = 64, Nv = 24, N = 16 and 24)
and config 2 (one homogeneous cell, Nv = N = 32).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--scaling strong|weak]
  torchrun ... bench.py --gpus N ...        (one rank per GPU, NCCL for the plumbing)
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

NX_STRONG = 512                                      # BASELINE config 5
CELLS_PER_GPU_WEAK = 32                              # BASELINE config 4: Nx = 256 on 8 GPUs
NV = 32
NSPEC = 32
PHYS = dict(Lv=5.25, nu=0.05, dt=0.01)              # [TwoStream] section of the reference input deck; Lx: see lx_of
LX_DECK = 4.0
A_AMP, K_WAVE = 0.5, 2 * np.pi / 4.
OUT = sys.stdout
METRIC, UNIT = "collision_cell_evals_per_s", "evals/s"


def lx_of(Nx):
    """Domain length for Nx cells.  The deck's Lx = 4 with dt = 0.01 is stable for its own mesh (dx = 0.125: CFL = Lv dt / dx
    = 0.42) but not for Nx = 512 (CFL 6.7: the explicit SSP-RK3 DG advection blows up to NaN within 20 steps); the cell
    size of the deck is kept instead and the periodic box grows with Nx (Lx = Nx / 8, a whole number of the k = 2 pi / 4
    perturbation's wavelengths), dt = 0.01 as in every deck of the reference."""
    return max(LX_DECK, Nx / 8.)


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
                self.rows.append((time.perf_counter(), [x.strip() for x in line.split(",")]))
                if self.stop_flag:
                    break
        except Exception:
            pass

    def finish(self, window=None):
        """window = (t0, t1) in perf_counter time: the timed region (samples outside it are idle clocks)."""
        self.stop_flag = True
        if self.proc:
            self.proc.terminate()
        sm, mx, reasons, load = [], 0., set(), []
        for t, r in self.rows:
            try:
                v = float(r[0]); mx = max(mx, float(r[1]))
                sm.append(v)
                if window and window[0] <= t <= window[1] + 0.1:
                    load.append(v)
                for name, flag in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                    if flag.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                continue
        use = load if load else sm
        return {"sm_mhz": float(np.median(use)) if use else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons), "samples": len(sm),
                "samples_in_timed_region": len(load)}


# ---------------------------------------------------------------------------------------------------------------
# the reference's own CPU implementation of the path, timed on this box's host cores (oracle/ is the checker; this
# and the --impl reference arm are the only places bench.py executes it)
def cpu_reference(reps, warm, slices):
    """ComputeQ + conserveMoments of the reference's own CPU code (oracle/_ref when it was built from /root/reference,
    else the C restatement): evals/s on one cell at N = Nv = 32, with every host core.  slices: also time the other
    phases of the reference's timestep on bounded samples -- one whole collision step of one cell (RK4: 3 more ComputeQ,
    3 FS, the IntModes projection, collisionRoutines_1.cpp:903-985) and RK3 on a 4-cell mesh -- so that a
    step-versus-step figure exists next to the isolated-operator one (LP_ompi.cpp:883-886 prints the whole loop)."""
    from oracle import oracle as orc
    threads = orc.set_num_threads()                   # torchrun exports OMP_NUM_THREADS=1: set it explicitly
    cfg = dict(Nx=1, Nv=NV, N=NSPEC, homogeneous=True, Lx=LX_DECK, **PHYS)
    kind, t_init = "port", time.time()
    ora = None
    if orc.have_ref():
        try:
            avail_gb = os.sysconf("SC_AVPHYS_PAGES") * os.sysconf("SC_PAGE_SIZE") / 2 ** 30
            if avail_gb > 8 * NSPEC ** 6 / 2 ** 30 + 4:
                ora = orc.RefOracle(**cfg)            # builds the reference's 8*N^6-byte weight table
                kind = "reference"
        except Exception:
            ora = None
    if ora is None:
        ora = orc.PortOracle(**cfg)
    t_init = time.time() - t_init
    po = orc.PortOracle(**cfg)                         # input data from the oracle side only: nothing of the product on this path
    Uh = po.SetInit_4H_Homo()
    f = po.setInit_spectral(Uh)[0]
    for _ in range(warm):
        ora.conserveMoments(ora.ComputeQ(f))
    times = []
    for _ in range(reps):
        t = time.time()
        ora.conserveMoments(ora.ComputeQ(f))
        times.append(time.time() - t)
    dt = float(np.median(times))
    out = dict(value=1. / dt, unit=UNIT, cores=int(ora.num_threads), kind=kind,
               sample="median of %d x (ComputeQ + conserveMoments) on 1 cell, N=%d, OpenMP on %d threads (set explicitly; %d cores visible); "
                      "min %.1f / max %.1f ms; table init %.0f s excluded" % (reps, NSPEC, ora.num_threads, threads, min(times) * 1e3, max(times) * 1e3, t_init))
    if slices:
        t = time.time()
        ora.collide_step(Uh)
        t_coll = time.time() - t
        nx = 4
        adv = (orc.RefOracle if kind == "reference" else orc.PortOracle)(**dict(cfg, Nx=nx, homogeneous=False), **({"build_weights": False} if kind == "reference" else {}))
        U4 = po_ld(orc, nx)
        adv.RK3(U4)
        t = time.time()
        adv.RK3(U4)
        t_rk3 = (time.time() - t) / nx
        out["timestep_slices"] = {
            "collision_step_s_per_cell": t_coll, "rk3_s_per_cell": t_rk3, "evals_per_collision_step": 4,
            "whole_timestep_evals_per_s": 4. / (t_coll + t_rk3),
            "note": "one whole collision step of the reference on one cell (4 ComputeQ + conserveMoments, 3 FS, the O(Nv^3 N^3) IntModes projection) and "
                    "RK3 on a %d-cell mesh per cell (the reference's field solve is O(Nx^2): per-cell cost grows with Nx, this is its cheapest); cells are "
                    "independent, so a timestep of Nx cells costs Nx times this on these cores" % nx}
    return out, dt


def po_ld(orc, nx):
    p = orc.PortOracle(Nx=nx, Nv=NV, N=NSPEC, Lx=LX_DECK, **PHYS)
    return p.SetInit_LD(A_AMP, K_WAVE, True)


def shape(world, scaling):
    if scaling == "weak":
        return CELLS_PER_GPU_WEAK * world, CELLS_PER_GPU_WEAK
    if NX_STRONG % world:
        raise SystemExit("bench.py: %d GPUs do not divide Nx = %d" % (world, NX_STRONG))
    return NX_STRONG, NX_STRONG // world


def workload_config(world, scaling):
    """config of the JSON line, shared by both arms"""
    Nx, per = shape(world, scaling)
    if scaling == "weak":
        name = ("BASELINE config 4 cut to its per-GPU shard: two-stream Landau-Poisson timestep (SSP-RK3 DG advection + RK4 spectral Landau "
                "collision), %d x-cells per GPU, Nv=%d, N=%d (Nx=256 on 8 GPUs)" % (per, NV, NSPEC))
    else:
        name = ("BASELINE config 5: Landau-Poisson timestep (SSP-RK3 DG advection + RK4 spectral Landau collision), Nx=%d, Nv=%d^3, N=%d, "
                "sharded over %d GPU(s); two-stream deck values (the reference has no bump-on-tail IC)" % (Nx, NV, NSPEC, world))
    return {"workload": name, "Nx": Nx, "Nv": NV, "N": NSPEC, "Lx": lx_of(Nx), "dt": PHYS["dt"], "nu": PHYS["nu"], "cells_per_gpu": per, "evals_per_step": 4 * Nx,
            "parallelism": "x-cells sharded over %d GPU(s)" % world}


def run_reference(args, rank):
    """--impl reference: the reference's CPU implementation of the metric's unit, timed on the host (rank 0 only)."""
    if rank != 0:
        return
    base, dt = cpu_reference(max(args.steps, 5), max(args.warmup, 1), slices=(args.gpus == 1))
    line = {"impl": "reference", "metric": METRIC, "value": base["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": max(args.steps, 5),
            "warmup": max(args.warmup, 1), "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": dict(workload_config(args.gpus, args.scaling),
                           sample="each step = one ComputeQ + conserveMoments on one cell (the metric's unit) by the reference's own CPU code on this "
                                  "box's host cores; the GPU arm's step is a whole timestep of every cell (4 such evaluations per cell plus the "
                                  "transforms, the projection and the advection), so value/value understates the step-versus-step ratio: see "
                                  "cpu_baseline.timestep_slices for the reference's whole timestep per cell"),
            "cpu_baseline": base, "e2e": {"value": base["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), file=OUT, flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=120)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--scaling", default="strong", choices=["strong", "weak"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the config 2 / config 3 / pipelined / direct-kernel side measurements")
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

    Nx, per_gpu = shape(world, args.scaling)
    s = solver.ShardedSolver(Nx, NV, NSPEC, homogeneous=False, rank=rank, world=world, device=local, dist=dist, Lx=lx_of(Nx), **PHYS)
    g = s.g
    g.set_stream(torch.cuda.current_stream().cuda_stream)
    U0 = solver.set_init_ld(Nx, NV, PHYS["Lv"], lx_of(Nx), A_AMP, K_WAVE, True, s.x_begin, s.x_count)
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
    w0 = time.perf_counter()
    e0.record()
    s.step(args.steps)            # one call: the library replays one CUDA graph per timestep
    e1.record()
    barrier()
    w1 = time.perf_counter()
    t_dev = max_over_ranks(e0.elapsed_time(e1) * 1e-3)
    launches = sum_over_ranks(g.launch_count - l0)
    value = 4. * Nx * args.steps / t_dev
    m_after = s.moments()
    finite = bool(np.all(np.isfinite(m_after)))
    # instrumented passes for the rooflines: a few of the same steps again with a CUDA-event pair around every launch of
    # the dominant kernel.  They cannot share the timed pass: there a step is a CUDA-graph replay in which the cells run as
    # concurrent chains (api.cu, collide_async), so one kernel's launch has no duration of its own; with the events on,
    # the library launches eagerly, one chain over all local cells, and each bracketed launch runs alone on the GPU.
    isteps = max(2, min(args.steps, 8))
    g.profile_computeQ(2)
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e2.record()
    s.step(isteps)
    e3.record()
    barrier()
    t_instr = max_over_ranks(e2.elapsed_time(e3) * 1e-3)
    cq_ms, cq_n = g.profile_read()
    #     ... and once more with the events around the DG stage kernels of the advection (HBM-bound)
    g.profile_computeQ(3)
    barrier()
    s.step(isteps)
    barrier()
    dg_ms, dg_n = g.profile_read()
    #     ... and around the moment reduction of the per-step diagnostics (computeMass / Momentum / KiE: HBM-bound)
    g.profile_computeQ(4)
    for _ in range(10):
        g.moments_partial()
    mo_ms, mo_n = g.profile_read()
    g.profile_computeQ(False)
    clocks = sampler.finish((w0, w1))

    # ---- end to end: host buffers in, host buffers out, every step --------------------------------
    # The headline e2e: ONE problem instance through ONE synchronous C-ABI call per timestep, lpgpu_step_host(U_in, U_out):
    # the state lives in (pinned) host memory between steps, as U does in LP_ompi.cpp's loop (:658 MPI_Bcast(U) ... :813);
    # every step moves the whole U to the GPU and back.  Inside the call the chunks of cells are uploaded, advected,
    # collided and downloaded as a pipeline (the download of chunk k overlaps the collisions of chunk k+1).
    e2e_steps = max(2, min(args.steps, 5))
    s.step_host(host_np, back_np)                                # warm
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        s.step_host(host_np, back_np)
    barrier()
    t_host = max_over_ranks(time.perf_counter() - t0)
    # the same work as three whole-shard calls (round 1's e2e): lpgpu_upload_U, lpgpu_step, lpgpu_download_U
    s.upload(host_np); s.step(1); s.download(back_np)           # warm
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        s.upload(host_np)          # H2D from pinned memory (the reference's MPI_Bcast(U), LP_ompi.cpp:658)
        s.step(1)
        s.download(back_np)        # D2H (the reference's gather for diagnostics/output, LP_ompi.cpp:813-849)
    barrier()
    t_ser = max_over_ranks(time.perf_counter() - t0)
    e2e = {"value": 4. * Nx * e2e_steps / t_host, "unit": UNIT, "h2d_bytes_per_step": int(host.numel() * 8 * world),
           "d2h_bytes_per_step": int(back.numel() * 8 * world), "steps": e2e_steps, "timesteps_per_s": e2e_steps / t_host,
           "ms_per_step": t_host / e2e_steps * 1e3,
           "note": "one problem instance, one synchronous call per timestep (lpgpu_step_host): every step uploads the whole U of every rank's shard "
                   "from pinned host memory, runs one timestep and downloads the whole U; chunks of cells are pipelined inside the call",
           "three_calls": {"value": 4. * Nx * e2e_steps / t_ser, "unit": UNIT, "ms_per_step": t_ser / e2e_steps * 1e3,
                           "note": "lpgpu_upload_U, lpgpu_step, lpgpu_download_U: three whole-shard phases, nothing overlapped"}}

    # as the reference's loop does it: after every timestep the moments, the entropy and the
    # negativity / KiE-ratio diagnostics (LP_ompi.cpp:817-849) -- GPU reductions, a few doubles back per step; the
    # diagnostics of step k run on a side stream over a snapshot while step k+1 runs (lpgpu_diagnostics_begin / _end)
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
                    "entropy_last": float(diag_last[1]),
                    "includes": "per-step mass/momentum/energy, entropy, negativity and KiE-ratio diagnostics (host reads ~10 doubles per step); "
                                "the diagnostics of step k run on a side stream over a snapshot while step k+1 runs"}

    # ---- rooflines of the dominant kernels -------------------------------------------------------
    # (1) the step's dominant kernel: the y/x-transform + product + inverse kernel of ComputeQ's FFT-convolution
    #     pipeline (its launches were bracketed by CUDA events, profile mode 2).  Bound: the FP64 pipe, so the roofline
    #     is quoted in FP64 TFLOP/s against the DFMA rate measured live on this GPU.  Algorithmic flops of that kernel
    #     (DESIGN.md section 4.1): per cell and per kz plane 14 x (N + M) forward and (M + N) inverse length-M line
    #     transforms at the textbook 5 M log2 M, plus 7 M^2 complex multiply-adds at 8 flop; M = 3N/2 planes.
    roof = roof_direct = None
    fp64_peak = pkg.lpgpu.fp64_peak_tflops(local)
    peaks = {}
    ppath = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(ppath):
        peaks = json.load(open(ppath))
    traffic = {}
    tpath = os.path.join(ROOT, "profiles", "kernel_traffic.json")
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
        tr_cell = traffic.get("f2_dram_bytes_per_cell")
        roof = {"bound": "fp64", "kernel": traffic.get("f2_kernel", "k_fc3_f2* (ComputeQ as FFT convolutions: y/x line transforms + products + inverse x/y of the kz planes)"),
                "achieved": achieved, "peak": fp64_peak, "unit": "TFLOP/s", "frac": achieved / fp64_peak,
                "traffic": tr_cell * s.x_count if tr_cell else None,
                "traffic_source": traffic.get("f2_source"),
                "avg_launch_ms": avg_s * 1e3, "launches": cq_n, "share_of_step": cq_ms * 1e-3 / t_instr,
                "measured_in": "instrumented pass: %d of the same steps, eager launches, one chain over all %d local cells, CUDA events around every launch of "
                               "this kernel (%.3f ms/step); the timed pass replays a CUDA graph with the cells in concurrent chains" % (isteps, s.x_count, t_instr / isteps * 1e3),
                "peak_source": fp64_src,
                "algorithmic_flop_per_launch": f2_flop_per_cell * s.x_count,
                "hbm_view": {"algorithmic_bytes_per_launch_whole_chain": chain_bytes_per_cell * s.x_count, "hbm_peak_gbs": peaks.get("hbm_gbs"),
                             "note": "F1+F2+F3 move 10+1 arrays of N^2 M complex once each way; at the measured HBM peak that is well under the FP64 time"},
                "direct_form_equivalent_tflops": flop_per_eval(NSPEC) * s.x_count / avg_s / 1e12,
                "note": "bound is the FP64 vector pipe (not hbm/tensor): the contraction is not dense, see DESIGN.md 4.1; direct_form_equivalent_tflops counts the "
                        "10 flop/pair of the O(N^6) sum this kernel replaces and exceeds the FP64 peak because the FFT form executes ~50x fewer flops for the "
                        "same result (parity-tested)"}
    roof_adv = None
    if dg_n > 0:
        # 384 B per DG cell per timestep = 96 (stage 1: 6 read + 6 written doubles) + 144 + 144 (12 read + 6
        # written); the x- and v1-upwind neighbours a DG cell also reads are other cells' own coefficients, served by L1/L2
        adv_bytes_per_launch = 128. * s.x_count * NV ** 3
        avg_dg = dg_ms * 1e-3 / dg_n
        hbm_peak = peaks.get("hbm_gbs") or 6500.
        tr_dg = traffic.get("dg_dram_bytes_per_cell")
        roof_adv = {"bound": "hbm", "kernel": "k_dg_stage<0..2> (SSP-RK3 stages of the DG upwind advection: I1, I2, I3, I5, H and the stage combination per DG cell)",
                    "achieved": adv_bytes_per_launch / avg_dg / 1e9, "peak": hbm_peak, "unit": "GB/s",
                    "frac": adv_bytes_per_launch / avg_dg / 1e9 / hbm_peak, "traffic": tr_dg * s.x_count if tr_dg else None,
                    "traffic_source": traffic.get("dg_source"),
                    "avg_launch_ms": avg_dg * 1e3, "launches": dg_n, "share_of_step": dg_ms * 1e-3 / t_instr,
                    "algorithmic_bytes_per_launch": adv_bytes_per_launch,
                    "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks.get("hbm_gbs") else "fallback 6500 GB/s (B200_PROFILING.md)"}

    roof_mom = None
    if mo_n > 0:
        mom_bytes = 40. * s.x_count * NV ** 3          # U0, U2, U3, U4, U5 of every DG cell, read once
        avg_mo = mo_ms * 1e-3 / mo_n
        hbm_peak = peaks.get("hbm_gbs") or 6500.
        roof_mom = {"bound": "hbm", "kernel": "k_moments_cell (computeMass, computeMomentum, computeKiE: per-cell partial sums of five DG coefficients)",
                    "achieved": mom_bytes / avg_mo / 1e9, "peak": hbm_peak, "unit": "GB/s", "frac": mom_bytes / avg_mo / 1e9 / hbm_peak,
                    "avg_launch_ms": avg_mo * 1e3, "launches": mo_n, "algorithmic_bytes_per_launch": mom_bytes}
    line = None
    if rank == 0:
        ws_mb = (3 * (s.x_count + 2) * 6 * NV ** 3 * 8 + s.x_count * NSPEC ** 3 * 8 * (3 + 2 * 6) + s.x_count * NSPEC * 4 * NV * NV * 16
                 + s.x_count * 11 * NSPEC * NSPEC * (3 * NSPEC // 2) * 16) / 2 ** 20
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": t_dev / args.steps * 1e3, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
                "dtype": "f64", "data": "synthetic",
                "config": dict(workload_config(world, args.scaling), l2="per-GPU working set %.0f MB > 126 MB L2; no flush between steps" % ws_mb,
                               exchange=("none (one GPU)" if world == 1 else
                                         "peer memory: kernels write halo planes and densities into the neighbours' buffers (CUDA IPC), flag-synchronised; no NCCL call in the timestep" if s.exchange == "peer" else
                                         "NCCL all-gather + send/recv per SSP-RK3 stage")),
                "timesteps_per_s": args.steps / t_dev, "timed_region_s": t_dev, "state_finite_after_timed_region": finite,
                "roofline": roof, "roofline_advection": roof_adv, "roofline_moments": roof_mom, "e2e": e2e, "as_reference_loop": as_reference,
                "gpu_launches": int(launches), "clocks": clocks}
    s.close()

    # ---- side measurements on one GPU: the other BASELINE configs, the direct kernel, the pipelined loop, the CPU baseline ----
    if world == 1 and not args.no_extras:
        def timed_steps(sv, n):
            sv.step(3)
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); sv.step(n); b.record()
            torch.cuda.synchronize()
            return a.elapsed_time(b) * 1e-3
        # config 2: one homogeneous cell
        h = solver.ShardedSolver(1, NV, NSPEC, homogeneous=True, device=local, Lx=LX_DECK, **PHYS)
        h.g.set_stream(torch.cuda.current_stream().cuda_stream)
        h.upload(solver.set_init_4h_homo(NV, PHYS["Lv"]))
        nh = 200
        th = timed_steps(h, nh)
        line["homogeneous_1cell"] = {"workload": "BASELINE config 2: space-homogeneous collision-only relaxation, 1 cell, Nv=N=32 (FourHump IC)",
                                     "value": 4 * nh / th, "unit": UNIT, "timesteps_per_s": nh / th, "steps": nh}
        h.close()
        # config 3: Landau damping, Nx = 64, Nv = 24, with the reference's own pairing N = 16 and with N = Nv = 24
        c3 = {}
        for n_spec in (16, 24):
            c = solver.ShardedSolver(64, 24, n_spec, homogeneous=False, device=local, Lv=5.25, Lx=4 * np.pi, nu=0.05, dt=0.01)
            c.g.set_stream(torch.cuda.current_stream().cuda_stream)
            c.upload(solver.set_init_ld(64, 24, 5.25, 4 * np.pi, 0.2, 0.5, False))
            n3 = 200
            t3 = timed_steps(c, n3)
            c3["N%d" % n_spec] = {"value": 4 * 64 * n3 / t3, "unit": UNIT, "timesteps_per_s": n3 / t3, "ms_per_step": t3 / n3 * 1e3, "steps": n3}
            c.close()
        line["landau_damping_nx64_nv24"] = dict(c3, workload="BASELINE config 3: 1D-3V collisional Landau damping, Nx=64, Nv=24^3 on one GPU; N=16 is the reference's own "
                                                              "pairing for Nv=24 (LP_ompi.cpp:637), N=24 the survey's N=Nv reading (the reference itself is unstable there, DESIGN.md 4.7)")
        # the north star's direct O(N^6) sum (k_computeQ_tiled, computeq_variant = 3) on 32 cells: FP64-pipe bound, 10 flop per pair
        try:
            nd = 32
            d = pkg.LPGpu(nd, NV, NSPEC, homogeneous=False, device=local, computeq_variant=3, Lx=LX_DECK, **PHYS)
            d.set_stream(torch.cuda.current_stream().cuda_stream)
            d.upload_U(solver.set_init_ld(nd, NV, PHYS["Lv"], LX_DECK, A_AMP, K_WAVE, True))
            d.sample_device()
            for _ in range(2):
                d.eval_device(nd)
            d.profile_computeQ(True)
            for _ in range(4):
                d.eval_device(nd)
            dq_ms, dq_n = d.profile_read()
            d.close()
            avg_d = dq_ms * 1e-3 / dq_n
            ach = flop_per_eval(NSPEC) * nd / avg_d / 1e12
            roof_direct = {"bound": "fp64", "kernel": "k_computeQ_tiled (computeq_variant=3, the reference's O(N^6) form), 32 cells per launch", "achieved": ach,
                           "peak": fp64_peak, "unit": "TFLOP/s", "frac": ach / fp64_peak, "traffic": traffic.get("direct_dram_bytes_per_launch"),
                           "avg_launch_ms": avg_d * 1e3, "launches": dq_n, "evals_per_s": nd / avg_d,
                           "peak_source": fp64_src, "algorithmic_flop_per_launch": flop_per_eval(NSPEC) * nd}
        except Exception as e:
            roof_direct = {"error": repr(e)}
        line["roofline_direct"] = roof_direct
        # pipelined end to end (an extra, not the headline): three independent 64-cell problem instances in flight, one
        # context and one stream each, so the copies of one batch overlap the kernels of another
        try:
            NFLIGHT, nxp = 3, 64
            streams = [torch.cuda.Stream() for _ in range(NFLIGHT)]
            ctxs = [solver.ShardedSolver(nxp, NV, NSPEC, homogeneous=False, device=local, stream=st, Lx=lx_of(nxp), **PHYS) for st in streams]
            Up = solver.set_init_ld(nxp, NV, PHYS["Lv"], lx_of(nxp), A_AMP, K_WAVE, True)
            hosts = [torch.from_numpy(Up.copy()).pin_memory() for _ in range(NFLIGHT)]
            backs = [torch.empty_like(hosts[0]).pin_memory() for _ in range(NFLIGHT)]

            def pipelined(n):
                for k in range(n):
                    q = ctxs[k % NFLIGHT]
                    q.synchronize()                               # this context's previous batch has left its host buffers
                    q.upload(hosts[k % NFLIGHT].numpy(), wait=False)
                    q.step(1, wait=False)
                    q.download(backs[k % NFLIGHT].numpy(), wait=False)
                for q in ctxs:
                    q.synchronize()

            pipelined(2 * NFLIGHT)
            pipe_steps = NFLIGHT * 6
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            pipelined(pipe_steps)
            torch.cuda.synchronize()
            t_pipe = time.perf_counter() - t0
            same = bool(all(np.array_equal(b.numpy(), backs[0].numpy()) for b in backs))
            for q in ctxs:
                q.close()
            line["e2e"]["pipelined"] = {"value": 4. * nxp * pipe_steps / t_pipe, "unit": UNIT, "cells_per_batch": nxp, "batches_in_flight": NFLIGHT, "steps": pipe_steps,
                                        "batches_agree": same,
                                        "note": "three independent 64-cell instances, enqueue-only calls (lpgpu_upload_U_async / lpgpu_step_async / "
                                                "lpgpu_download_U_async): every step still moves its whole U both ways"}
        except Exception as e:
            line["e2e"]["pipelined"] = {"error": repr(e)}
        if not args.no_cpu_baseline:
            try:
                line["cpu_baseline"], _ = cpu_reference(20, 3, slices=False)
            except Exception as e:                          # the oracle is a checker; never let it break the GPU number
                line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": os.cpu_count(), "kind": "unavailable", "sample": repr(e)}
    if rank == 0:
        print(json.dumps(line), file=OUT, flush=True)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
