"""Developer helper: a few homogeneous single-cell steps (for an ncu launch list)."""
import sys
import numpy as np
sys.path.insert(0, ".")
import __graft_entry__ as graft
pkg = graft.load_package()
from lpsolver_b200 import solver
N = 32
h = solver.ShardedSolver(1, N, N, homogeneous=True, Lv=5.25, Lx=4.0, nu=0.05, dt=0.01)
h.upload(solver.set_init_4h_homo(N, 5.25))
for _ in range(4):
    h.g.step(1)
h.close()
