"""Developer timing helper (not the bench): times lpgpu_eval_device for a few (N, B) with CUDA
events around the context's stream.  usage: python scripts/dev_time.py N B [variant] [reps]"""
import sys, time, ctypes as C
import numpy as np
sys.path.insert(0, ".")
import __graft_entry__ as graft
pkg = graft.load_package()
N = int(sys.argv[1]); B = int(sys.argv[2]); variant = int(sys.argv[3]) if len(sys.argv) > 3 else 0
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 5
cfg = dict(Nx=B, Nv=N, N=N, Lv=5.25, Lx=4 * np.pi, nu=0.05, dt=0.01)
g = pkg.LPGpu(computeq_variant=variant, **cfg)
rng = np.random.default_rng(0)
U = 0.01 * rng.standard_normal(B * N ** 3 * 6)
g.upload_U(U); g.sample_device(); g.eval_device(B); g.synchronize()
t = time.time()
for _ in range(reps): g.eval_device(B)
g.synchronize()
dt = (time.time() - t) / reps
pairs = (3 * N * N / 4) ** 3
print("N=%d B=%d variant=%d: %.3f ms per batch, %.1f evals/s, %.2f TFLOP/s algorithmic" % (N, B, variant, dt * 1e3, B / dt, 10 * pairs * B / dt / 1e12))
