"""Regenerates the committed fixtures in tests/golden/ from the UNMODIFIED reference.

Run in the build container (needs /root/reference and oracle/_ref/libref.so, i.e. after
`make -C oracle ref`):   python tests/golden/make_golden.py

* reference_moments.json  -- the numbers of /root/reference/tests/Moments_Test{0..4}.dc (the
  reference's own golden files: 6 rows of %11.8g moments per test case).
* ref_vectors.npz         -- element-wise outputs of the reference's functions (ComputeQ,
  conserveMoments, fft3D, FS, setInit_spectral, RK4 via the collision branch, RK3, field
  integrals, moments) on a tiny deterministic input, computed by oracle/_ref/libref.so.
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle.oracle import RefOracle  # noqa: E402

REF_TESTS = "/root/reference/tests"


def moments_files():
    out = {}
    for k in range(5):
        name = "Moments_Test%d.dc" % k
        rows = [[float(x) for x in line.split()] for line in open(os.path.join(REF_TESTS, name)) if line.strip()]
        out[name] = rows
    return out


def deterministic_U(n):
    """Smooth, sign-changing perturbation (no RNG): makes every DG coefficient non-trivial."""
    i = np.arange(n, dtype=np.float64)
    return 1 + 0.05 * np.sin(0.37 * i + 0.1) , 1e-4 * np.cos(0.11 * i)


def vectors():
    cfg = dict(Nx=3, Nv=6, N=8, Lv=5.25, Lx=4 * np.pi, nu=0.05, dt=0.01)
    R = RefOracle(**cfg)
    U0 = R.SetInit_LD(0.2, 0.5)
    a, b = deterministic_U(U0.size)
    U0 = U0 * a + b
    f = R.setInit_spectral(U0)
    x = np.stack([f[1], 0.3 * f[1][::-1]], 1)
    q = R.ComputeQ(f[1] * (1 + 0.1 * np.sin(np.arange(f[1].size))))
    out = dict(cfg=json.dumps(cfg), U0=U0, f=f, x=x, fft3D=R.fft3D(x), FS=R.FS(x), qHat=q,
               qHat_conserved=R.conserveMoments(q), U_collide=R.collide_step(U0), stage_spectra=R.stage_spectra(),
               field=R.field(U0), U_RK3=R.RK3(U0), moments=R.moments(U0))
    out["U_step"] = R.step(U0)
    # homogeneous (RK4_Homo) on the same velocity grid
    Rh = RefOracle(homogeneous=True, **cfg)
    Uh = Rh.SetInit_4H_Homo()
    out["Uh0"] = Uh
    out["Uh_collide"] = Rh.collide_step(Uh)
    out["moments_h"] = Rh.moments(out["Uh_collide"])
    return out


if __name__ == "__main__":
    json.dump(moments_files(), open(os.path.join(HERE, "reference_moments.json"), "w"), indent=1)
    np.savez_compressed(os.path.join(HERE, "ref_vectors.npz"), **vectors())
    print("wrote", os.listdir(HERE))
