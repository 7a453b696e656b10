"""Headline-size fixtures (N = Nv = 32) from the UNMODIFIED reference: tests/golden/ref_n32.npz.

Run in the build container (needs /root/reference, oracle/_ref/libref.so, ~10 GB of memory for the
reference's N^6 weight table and ~10 minutes on 8 threads):

    python tests/golden/make_n32_golden.py

Contents (everything computed by oracle/_ref/libref.so, i.e. the reference's own ComputeQ /
conserveMoments / RK4_Homo / RK4_Inhomo / RK3, collisionRoutines_1.cpp:691-774, :903-985, :1087-1167):

* homogeneous FourHump cell (BASELINE config 2): f, qHat = ComputeQ(f * (1 + 0.1 sin)), its conserved
  form, the state after one collision step (full), the three later stage spectra (strided sample);
* two-cell two-stream shard (BASELINE configs 4/5, Lx = 4, A = 0.5): the state after the reference's
  collision step and after a whole timestep (RK3 + collisions), as strided samples plus sums, and the
  six moments before/after.
Large arrays are stored as strided samples (stride coprime to every array extent) with full sums, to
keep the fixture at a few MB.
"""
import json
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle.oracle import RefOracle  # noqa: E402

STRIDE = 7


def main():
    out = {}
    t0 = time.time()
    cfg = dict(Nx=1, Nv=32, N=32, Lv=5.25, Lx=4.0, nu=0.05, dt=0.01)
    R = RefOracle(homogeneous=True, **cfg)
    print("weights built %.0f s" % (time.time() - t0), flush=True)
    Uh = R.SetInit_4H_Homo()
    f = R.setInit_spectral(Uh)[0]
    fa = f * (1 + 0.1 * np.sin(np.arange(f.size)))
    q = R.ComputeQ(fa)
    out.update(cfg_h=json.dumps(cfg), stride=STRIDE, Uh0=Uh, f_h=f, qHat=q, qHat_conserved=R.conserveMoments(q))
    print("ComputeQ done %.0f s" % (time.time() - t0), flush=True)
    Uc = R.collide_step(Uh)
    out.update(Uh_collide=Uc, stage_spectra_h=R.stage_spectra().reshape(3, -1)[:, ::STRIDE], moments_h0=R.moments(Uh), moments_h1=R.moments(Uc))
    print("homogeneous collide_step done %.0f s" % (time.time() - t0), flush=True)

    cfg2 = dict(Nx=2, Nv=32, N=32, Lv=5.25, Lx=4.0, nu=0.05, dt=0.01)
    R = RefOracle(**cfg2)
    print("weights rebuilt %.0f s" % (time.time() - t0), flush=True)
    U0 = R.SetInit_LD(0.5, 2 * np.pi / 4., True)
    Uc = R.collide_step(U0)
    print("2-cell collide_step done %.0f s" % (time.time() - t0), flush=True)
    Us = R.step(U0)
    print("2-cell step done %.0f s" % (time.time() - t0), flush=True)
    out.update(cfg_2=json.dumps(cfg2), U0_2_sample=U0[::STRIDE], U0_2_sum=U0.sum(),
               Uc_2_sample=Uc[::STRIDE], Uc_2_sum=Uc.sum(), Uc_2_abs=np.abs(Uc).sum(),
               Us_2_sample=Us[::STRIDE], Us_2_sum=Us.sum(), Us_2_abs=np.abs(Us).sum(),
               moments_2_0=R.moments(U0), moments_2_c=R.moments(Uc), moments_2_s=R.moments(Us), field_2=R.field(U0))
    np.savez_compressed(os.path.join(HERE, "ref_n32.npz"), **out)
    print("wrote ref_n32.npz (%.1f MB) in %.0f s" % (os.path.getsize(os.path.join(HERE, "ref_n32.npz")) / 2 ** 20, time.time() - t0))


if __name__ == "__main__":
    main()
