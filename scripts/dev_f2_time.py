"""Developer timing of the dominant ComputeQ kernel alone (k_fc3_f2_tmem, CUDA events inside the library around every
launch; N = Nv = 32) and of F1 + F2 + F3 together.  usage: python scripts/dev_f2_time.py [cells ...]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import __graft_entry__ as graft
pkg = graft.load_package()
from lpsolver_b200 import solver
for ncell in [int(a) for a in sys.argv[1:]] or [32, 128]:
    g = pkg.LPGpu(ncell, 32, 32, Lv=5.25, Lx=ncell / 8., nu=0.05, dt=0.01)
    g.upload_U(solver.set_init_ld(ncell, 32, 5.25, ncell / 8., 0.5, np.pi / 2, True)); g.sample_device()
    for _ in range(3): g.eval_device(ncell)
    for mode in (2, 1):
        g.profile_computeQ(mode)
        for _ in range(10): g.eval_device(ncell)
        ms, n = g.profile_read()
        print("%d cells: %s %.2f us per launch (%.3f us per cell)" % (ncell, "F2 kernel" if mode == 2 else "F1+F2+F3  ", ms / n * 1e3, ms / n * 1e3 / ncell), flush=True)
    g.close()
