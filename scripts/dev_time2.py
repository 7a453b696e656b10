"""Developer timing helper: CUDA-event time of repeated eval_device batches, min/median over rounds.
usage: python scripts/dev_time2.py N B [rounds] [reps]"""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
import __graft_entry__ as graft
pkg = graft.load_package()
N = int(sys.argv[1]); B = int(sys.argv[2]); rounds = int(sys.argv[3]) if len(sys.argv) > 3 else 7; reps = int(sys.argv[4]) if len(sys.argv) > 4 else 20
cfg = dict(Nx=B, Nv=N, N=N, Lv=5.25, Lx=4 * np.pi, nu=0.05, dt=0.01)
g = pkg.LPGpu(**cfg)
g.set_stream(torch.cuda.current_stream().cuda_stream)
rng = np.random.default_rng(0)
g.upload_U(0.01 * rng.standard_normal(B * N ** 3 * 6)); g.sample_device()
for _ in range(5): g.eval_device(B)
ts = []
for _ in range(rounds):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(reps): g.eval_device(B)
    e1.record(); torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1) / reps)
print("N=%d B=%d: eval batch min %.4f ms, median %.4f ms -> %.0f evals/s" % (N, B, min(ts), float(np.median(ts)), B / (min(ts) * 1e-3)))
