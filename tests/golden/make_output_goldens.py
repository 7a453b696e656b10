"""Generate tests/golden/ref_outputs.npz: the Marginals / PhiVals / FieldVals / EntropyVals files the UNMODIFIED
reference writes for its own test0, test4 and test1 decks (oracle/_ref/solver, built from /root/reference by
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
for case in ("test0", "test4", "test1"):
    with tempfile.TemporaryDirectory() as tmp:
        shutil.copy(os.path.join(HERE, "LPsolver-input-%s.txt" % case), os.path.join(tmp, "LPsolver-input.txt"))
        subprocess.run([SOLVER], cwd=tmp, check=True, capture_output=True)
        for kind in ("Marginals", "PhiVals", "FieldVals", "EntropyVals"):
            f = glob.glob(os.path.join(tmp, "Data", kind + "_*"))
            assert len(f) == 1, (case, kind, f)
            r = rows(f[0]) if os.path.getsize(f[0]) else np.zeros((0, 0))
            out["%s_%s" % (case, kind)] = r
            print(case, kind, r.shape)
np.savez_compressed(os.path.join(HERE, "ref_outputs.npz"), **out)
