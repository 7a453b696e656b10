"""Developer check: round-off of the two GPU ComputeQ kernels against the CPU oracle at full size.
usage: python scripts/dev_precision.py N"""
import sys, time
import numpy as np
sys.path.insert(0, ".")
import __graft_entry__ as graft
from oracle.oracle import PortOracle
pkg = graft.load_package()
N = int(sys.argv[1])
cfg = dict(Nx=1, Nv=N, N=N, Lv=5.25, Lx=4 * np.pi, nu=0.05, dt=0.01)
ora = PortOracle(homogeneous=True, **cfg)
f = ora.setInit_spectral(ora.SetInit_4H_Homo())[0]
rel = lambda a, b: np.max(np.abs(a - b)) / np.max(np.abs(b))
for name, ff in (("symmetric", f), ("asymmetric", f * (1 + 0.1 * np.sin(np.arange(f.size))))):
    t = time.time(); qo = ora.ComputeQ(ff); t = time.time() - t
    g0 = pkg.LPGpu(homogeneous=True, **cfg); q0 = g0.ComputeQ(ff)[0]; c0 = g0.conserveMoments(q0)[0]; g0.close()
    g1 = pkg.LPGpu(homogeneous=True, computeq_variant=1, **cfg); q1 = g1.ComputeQ(ff)[0]; c1 = g1.conserveMoments(q1)[0]; g1.close()
    co = ora.conserveMoments(qo)
    print("%s N=%d oracle %.1fs (%d thr): raw max|q|=%.3e  tiled-vs-oracle %.2e simple-vs-oracle %.2e tiled-vs-simple %.2e | conserved max=%.3e: %.2e %.2e %.2e"
          % (name, N, t, ora.num_threads, np.abs(qo).max(), rel(q0, qo), rel(q1, qo), rel(q0, q1), np.abs(co).max(), rel(c0, co), rel(c1, co), rel(c0, c1)))
