"""Developer helper: a few timesteps of an Nx-cell problem (Nv = N = 32) for an ncu launch list.  usage: dev_cells.py Nx"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import __graft_entry__ as graft
pkg = graft.load_package()
from lpsolver_b200 import solver
Nx = int(sys.argv[1])
s = solver.ShardedSolver(Nx, 32, 32, Lv=5.25, Lx=max(4.0, Nx / 8.), nu=0.05, dt=0.01)
s.upload(solver.set_init_ld(Nx, 32, 5.25, max(4.0, Nx / 8.), 0.5, np.pi / 2, True))
s.step(1); s.step(1); s.step(2)
s.close()
