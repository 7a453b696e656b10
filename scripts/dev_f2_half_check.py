"""Developer check of the experimental two-thread-third F2 kernel: run with LPGPU_F2_HALF=1.
Parity of ComputeQ (2 cells, N = 32) against the tiled direct sum, then the kernel's time on 32 cells."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import __graft_entry__ as graft
pkg = graft.load_package()
from lpsolver_b200 import solver
cfg = dict(Nx=2, Nv=32, N=32, Lv=5.25, Lx=4.0, nu=0.05, dt=0.01)
g, d = pkg.LPGpu(**cfg), pkg.LPGpu(computeq_variant=3, **cfg)
U = solver.set_init_ld(2, 32, 5.25, 4.0, 0.5, np.pi / 2, True)
g.upload_U(U); f = g.setInit_spectral()
f = f * (1 + 0.1 * np.sin(np.arange(f.shape[1])))
a, b = g.ComputeQ(f), d.ComputeQ(f)
print("F2_HALF=%s: ComputeQ rel err vs direct sum %.3e" % (os.environ.get("LPGPU_F2_HALF"), np.abs(a - b).max() / np.abs(b).max()), flush=True)
g.close(); d.close()
g = pkg.LPGpu(32, 32, 32, Lv=5.25, Lx=4.0, nu=0.05, dt=0.01)
g.upload_U(solver.set_init_ld(32, 32, 5.25, 4.0, 0.5, np.pi / 2, True)); g.sample_device()
for _ in range(2): g.eval_device(32)
g.profile_computeQ(2)
for _ in range(10): g.eval_device(32)
ms, n = g.profile_read()
print("F2 kernel: %.2f us per launch over %d launches" % (ms / n * 1e3, n))
