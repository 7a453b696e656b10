"""Three warm collision steps of 32 cells at N = Nv = 32 (eager, one chain) for profiler captures."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import __graft_entry__ as graft
pkg = graft.load_package()
from lpsolver_b200 import solver
ncell = int(sys.argv[1]) if len(sys.argv) > 1 else 32
g = pkg.LPGpu(ncell, 32, 32, Lv=5.25, Lx=ncell / 8., nu=0.05, dt=0.01)
g.upload_U(solver.set_init_ld(ncell, 32, 5.25, ncell / 8., 0.5, np.pi / 2, True))
g.profile_computeQ(1)          # eager launches, one chain over all cells
for _ in range(3): g.step(1)
g.synchronize()
g.close()
