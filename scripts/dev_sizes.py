"""Developer check: homogeneous collide step and FS against the CPU oracle for mixed (N, Nv)."""
import sys
import numpy as np
sys.path.insert(0, ".")
import __graft_entry__ as graft
pkg = graft.load_package()
from oracle.oracle import PortOracle
for N, Nv in ((24, 8), (8, 24), (24, 16), (16, 24), (8, 8), (32, 8), (8, 32)):
    cfg = dict(Nx=1, Nv=Nv, N=N, Lv=5.25, Lx=4.0, nu=0.05, dt=0.01)
    ora = PortOracle(homogeneous=True, **cfg)
    g = pkg.LPGpu(homogeneous=True, **cfg)
    U = ora.SetInit_4H_Homo()
    rng = np.random.default_rng(0)
    x = rng.standard_normal((N ** 3, 2))
    efs = np.abs(g.FS(x)[0][:, 0] - ora.FS(x)[:, 0]).max() / np.abs(ora.FS(x)[:, 0]).max()
    eff = np.abs(g.fft3D(x)[0] - ora.fft3D(x)).max() / np.abs(ora.fft3D(x)).max()
    f = ora.setInit_spectral(U)[0]
    esm = np.abs(g.upload_U(U) or g.setInit_spectral()[0] - f).max()
    g.upload_U(U)
    g.collide_step()
    got = g.download_U()
    want = ora.step(U)
    g.close()
    print("N=%d Nv=%d: FS err %.2e fft3D err %.2e sample err %.2e collide-step rel err of update %.3e (max|dU| %.3e)" % (
        N, Nv, efs, eff, esm, np.abs(got - want).max() / np.abs(want - U).max(), np.abs(want - U).max()), flush=True)
