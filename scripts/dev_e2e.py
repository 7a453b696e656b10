"""Developer timing of lpgpu_step_host at the headline size (Nx = 512, Nv = N = 32): ms per step, pinned buffers."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import __graft_entry__ as graft
pkg = graft.load_package()
from lpsolver_b200 import solver
Nx = int(sys.argv[1]) if len(sys.argv) > 1 else 512
s = solver.ShardedSolver(Nx, 32, 32, Lv=5.25, Lx=Nx / 8., nu=0.05, dt=0.01)
host = torch.from_numpy(solver.set_init_ld(Nx, 32, 5.25, Nx / 8., 0.5, 2 * np.pi / (Nx / 8.), True)).pin_memory()
back = torch.empty_like(host).pin_memory()
s.step_host(host.numpy(), back.numpy())
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(5): s.step_host(host.numpy(), back.numpy())
torch.cuda.synchronize(); t = (time.perf_counter() - t0) / 5
print("LPGPU_HOST_CHUNK=%s  step_host %.2f ms/step  %.0f evals/s" % (os.environ.get("LPGPU_HOST_CHUNK", "-"), t * 1e3, 4 * Nx / t))
s.upload(host.numpy()); s.step(1); torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(5): s.step(1)
torch.cuda.synchronize(); print("device-resident step %.2f ms" % ((time.perf_counter() - t0) / 5 * 1e3))
t0 = time.perf_counter()
for _ in range(3): s.upload(host.numpy())
torch.cuda.synchronize(); print("upload %.2f ms" % ((time.perf_counter() - t0) / 3 * 1e3))
t0 = time.perf_counter()
for _ in range(3): s.download(back.numpy())
torch.cuda.synchronize(); print("download %.2f ms" % ((time.perf_counter() - t0) / 3 * 1e3))
