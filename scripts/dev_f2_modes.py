"""Developer check of the ComputeQ y/x-stage kernels (LPGPU_F2_MODE: 0 = the library's choice by cell count, 1 = k_fc3_f2_tmem,
3 = the warp-specialised persistent k_fc3_f2s): parity of ComputeQ against the tiled direct sum at 2 and 9
cells (one plane per CTA / several planes per CTA, ragged), then the kernel's time at 32 and 128 cells.
usage: LPGPU_F2_MODE=k python scripts/dev_f2_modes.py"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import __graft_entry__ as graft
pkg = graft.load_package()
from lpsolver_b200 import solver
mode = os.environ.get("LPGPU_F2_MODE", "0")
for ncell in (2, 9):
    cfg = dict(Nx=ncell, Nv=32, N=32, Lv=5.25, Lx=4.0, nu=0.05, dt=0.01)
    g, d = pkg.LPGpu(**cfg), pkg.LPGpu(computeq_variant=3, **cfg)
    U = solver.set_init_ld(ncell, 32, 5.25, 4.0, 0.5, np.pi / 2, True)
    g.upload_U(U); f = g.setInit_spectral()
    f = f * (1 + 0.1 * np.sin(np.arange(f.shape[1]) + np.arange(ncell)[:, None]))
    a, b = g.ComputeQ(f), d.ComputeQ(f)
    a2 = g.ComputeQ(f)
    print("mode %s, %d cells: ComputeQ rel err vs direct sum %.3e, run-to-run identical %s" % (mode, ncell, np.abs(a - b).max() / np.abs(b).max(), np.array_equal(a, a2)), flush=True)
    g.close(); d.close()
for ncell in (32, 128):
    g = pkg.LPGpu(ncell, 32, 32, Lv=5.25, Lx=ncell / 8., nu=0.05, dt=0.01)
    g.upload_U(solver.set_init_ld(ncell, 32, 5.25, ncell / 8., 0.5, np.pi / 2, True)); g.sample_device()
    for _ in range(2): g.eval_device(ncell)
    g.profile_computeQ(2)
    for _ in range(10): g.eval_device(ncell)
    ms, n = g.profile_read()
    print("mode %s, %d cells: F2 kernel %.2f us per launch (%.3f us per cell) over %d launches" % (mode, ncell, ms / n * 1e3, ms / n * 1e3 / ncell, n), flush=True)
    g.close()
