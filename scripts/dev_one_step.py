"""Developer script for ncu: three warm timesteps, then one more (eager launches, 32 cells, Nv = N = 32)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import __graft_entry__ as graft
pkg = graft.load_package()
from lpsolver_b200 import solver
s = solver.ShardedSolver(32, 32, 32, Lv=5.25, Lx=4.0, nu=0.05, dt=0.01)
s.upload(solver.set_init_ld(32, 32, 5.25, 4.0, 0.5, np.pi / 2, True))
l0 = s.g.launch_count
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 4):
    s.step(1)
print("launches per step:", (s.g.launch_count - l0) / (int(sys.argv[1]) if len(sys.argv) > 1 else 4), "before:", l0)
