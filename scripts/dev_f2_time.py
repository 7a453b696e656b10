"""Developer timing of the dominant ComputeQ kernel alone (k_fc3_f2_tmem, CUDA events inside the library around every
launch; 32 cells, N = Nv = 32).  usage: python scripts/dev_f2_time.py [reps]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import __graft_entry__ as graft
pkg = graft.load_package()
from lpsolver_b200 import solver
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
g = pkg.LPGpu(32, 32, 32, Lv=5.25, Lx=4.0, nu=0.05, dt=0.01)
g.upload_U(solver.set_init_ld(32, 32, 5.25, 4.0, 0.5, np.pi / 2, True))
g.sample_device()
for _ in range(3):
    g.eval_device(32)
g.profile_computeQ(2)
for _ in range(reps):
    g.eval_device(32)
ms, n = g.profile_read()
print("%s k_fc3_f2_tmem: %.2f us per launch over %d launches" % (os.environ.get("LPGPU_F2_SKEW_NS", "-"), ms / n * 1e3, n))
