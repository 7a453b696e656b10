"""Golden vectors for the other two collision kernels the reference accepts (InputParsing.cpp:202-238):
gamma = 0 (Maxwell molecules) and gamma = 1 (hard spheres), from the UNMODIFIED reference (oracle/_ref/libref.so).
No reference test deck uses them, so these are outputs of the reference run here.

Run in the build container after `make -C oracle ref`:   python tests/golden/make_gamma_golden.py
Writes tests/golden/ref_gamma.npz: for each gamma, gHat3 at a few (xi, omega) pairs, ComputeQ of one spectral sample,
its conserved spectrum, the state after one collision step (RK4_Inhomo) and after a homogeneous one (RK4_Homo).
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle.oracle import RefOracle  # noqa: E402
sys.path.insert(0, HERE)
from make_golden import deterministic_U  # noqa: E402

PAIRS = np.array([[0.3, -0.2, 0.5, 0.0, 0.0, 0.0], [0.3, -0.2, 0.5, 0.7, 0.0, -0.4], [-1.1, 0.6, 0.2, 0.9, -0.8, 0.35],
                  [0.0, 0.0, 0.0, 0.25, 0.5, -0.75], [2.0, 1.0, -3.0, -2.5, 0.5, 1.5]])


def main():
    cfg = dict(Nx=2, Nv=6, N=8, Lv=5.25, Lx=4 * np.pi, nu=0.05, dt=0.01)
    out = dict(cfg=json.dumps(cfg), pairs=PAIRS)
    for gamma in (0, 1):
        R = RefOracle(gamma=gamma, **cfg)
        U0 = R.SetInit_LD(0.2, 0.5)
        a, b = deterministic_U(U0.size)
        U0 = U0 * a + b
        f = R.setInit_spectral(U0)
        fin = f[1] * (1 + 0.1 * np.sin(np.arange(f[1].size)))
        q = R.ComputeQ(fin)
        tag = "g%d_" % gamma
        out.update({tag + "U0": U0, tag + "f": fin, tag + "qHat": q, tag + "qHat_conserved": R.conserveMoments(q),
                    tag + "U_collide": R.collide_step(U0), tag + "gHat3": np.array([R.gHat3(p[:3], p[3:]) for p in PAIRS])})
        Rh = RefOracle(homogeneous=True, gamma=gamma, **cfg)
        Uh = Rh.SetInit_4H_Homo()
        out[tag + "Uh0"] = Uh
        out[tag + "Uh_collide"] = Rh.collide_step(Uh)
    np.savez_compressed(os.path.join(HERE, "ref_gamma.npz"), **out)
    print("wrote ref_gamma.npz:", sorted(out))


if __name__ == "__main__":
    main()
