"""Developer timing of the per-step diagnostics (not a test): where the as-reference loop spends its time."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import __graft_entry__ as graft
pkg = graft.load_package()
from lpsolver_b200 import solver
PHYS = dict(Lv=5.25, Lx=4.0, nu=0.05, dt=0.01)
s = solver.ShardedSolver(32, 32, 32, **PHYS)
s.upload(solver.set_init_ld(32, 32, 5.25, 4.0, 0.5, np.pi / 2, True))
s.step(2)
def t(fn, n=10):
    fn(); torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n * 1e3
print("step           %.3f ms" % t(lambda: s.step(1)))
print("moments()      %.3f ms" % t(s.moments))
print("moments_partial%.3f ms" % t(s.g.moments_partial))
print("diagnostics    %.3f ms" % t(s.g.diagnostics_partial))
print("marginal_sums  %.3f ms" % t(s.g.marginal_sums))
