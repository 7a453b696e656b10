"""Developer check: the N = Nv = 24 timestep (the reference's blow-up case) against the oracle, folded and unfolded projection."""
import importlib, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("landau-poisson-solver_b200")
from oracle.oracle import PortOracle
N = int(sys.argv[1]) if len(sys.argv) > 1 else 24
cfg = dict(Nx=4, Nv=24, N=N, Lv=5.25, Lx=4 * np.pi, nu=0.05, dt=0.01)
ora = PortOracle(**cfg)
U0 = ora.SetInit_LD(0.2, 0.5)
want = ora.step(U0)
g = pkg.LPGpu(**cfg)
g.upload_U(U0)
g.step(1)
got = g.download_U()
rel = lambda a, b: np.max(np.abs(a - b)) / np.max(np.abs(b))
print("N", N, "relerr U", rel(got, want), "relerr dU", rel(got - U0, want - U0), "max|U0|", np.abs(U0).max(), "max|dU|", np.abs(want - U0).max())
