"""Generate tests/golden/ref_outputs.npz: the Marginals / PhiVals / FieldVals / EntropyVals files the UNMODIFIED
reference writes for its own test0, test4, test1 and test2 decks, plus a `Second = True`
restart of test0 (3 steps, then 2 more from the U_*.dc checkpoint, LP_ompi.cpp:529-571) (oracle/_ref/solver, built from /root/reference by
oracle/Makefile).  Run in the build container; the GPU box only reads the committed .npz.

    python tests/golden/make_output_goldens.py
"""
import glob
import os
import shutil
import subprocess
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
SOLVER = os.path.join(ROOT, "oracle", "_ref", "solver")


def rows(path):
    return np.array([[float(x) for x in line.split()] for line in open(path) if line.strip()])


out = {}
for case in ("test0", "test4", "test1", "test2"):
    with tempfile.TemporaryDirectory() as tmp:
        shutil.copy(os.path.join(HERE, "LPsolver-input-%s.txt" % case), os.path.join(tmp, "LPsolver-input.txt"))
        subprocess.run([SOLVER], cwd=tmp, check=True, capture_output=True)
        for kind in ("Marginals", "PhiVals", "FieldVals", "EntropyVals"):
            f = glob.glob(os.path.join(tmp, "Data", kind + "_*"))
            assert len(f) == 1, (case, kind, f)
            r = rows(f[0]) if os.path.getsize(f[0]) else np.zeros((0, 0))
            out["%s_%s" % (case, kind)] = r
            print(case, kind, r.shape)

# Second = True: the reference run for 3 steps, then restarted from its own checkpoint for 2 more
deck = open(os.path.join(HERE, "LPsolver-input-test0.txt")).read()
with tempfile.TemporaryDirectory() as tmp:
    a = deck.replace("flag     = Test0", "flag     = RestartA").replace("nT       = 5 ", "nT       = 3 ")
    assert a != deck
    open(os.path.join(tmp, "LPsolver-input.txt"), "w").write(a)
    subprocess.run([SOLVER], cwd=tmp, check=True, capture_output=True)
    ua = glob.glob(os.path.join(tmp, "Data", "U_*RestartA.dc"))
    assert len(ua) == 1
    b = deck.replace("flag     = Test0", "flag     = RestartB").replace("nT       = 5 ", "nT       = 2 ")
    b = b.replace("First            = True", "First            = False").replace("Second           = False", "Second           = True")
    b += "\n[Second]\nName = %s\n" % os.path.basename(ua[0])
    open(os.path.join(tmp, "LPsolver-input.txt"), "w").write(b)
    subprocess.run([SOLVER], cwd=tmp, check=True, capture_output=True)
    ub = glob.glob(os.path.join(tmp, "Data", "U_*RestartB.dc"))
    assert len(ub) == 1
    U = np.fromfile(ub[0])
    out["restart_name_a"] = os.path.basename(ua[0])
    out["restart_U_sample"] = U[::5]
    out["restart_U_sum"] = U.sum()
    out["restart_U_abs"] = np.abs(U).sum()
    out["restart_Moments"] = rows(glob.glob(os.path.join(tmp, "Data", "Moments_*RestartB.dc"))[0])
    print("restart", U.shape, out["restart_Moments"].shape)
np.savez_compressed(os.path.join(HERE, "ref_outputs.npz"), **out)
