"""Generate tests/golden/ref_test1.npz: the state the UNMODIFIED reference (oracle/_ref/solver, built from
/root/reference by oracle/Makefile) reaches after the five steps of its own test 1 deck -- Doping = True,
LinearLandau = True, MassConsOnly = True.  The reference dumps U as raw doubles (LP_ompi.cpp:896); a strided
sample of it and a few norms are committed (the full array is 3 MB).  Run in the build container:

    python tests/golden/make_test1_golden.py
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
STRIDE = 37

with tempfile.TemporaryDirectory() as tmp:
    shutil.copy(os.path.join(HERE, "LPsolver-input-test1.txt"), os.path.join(tmp, "LPsolver-input.txt"))
    subprocess.run([SOLVER], cwd=tmp, check=True, capture_output=True)
    f = glob.glob(os.path.join(tmp, "Data", "U_*"))
    assert len(f) == 1, f
    U = np.fromfile(f[0])
    mom = np.array([[float(x) for x in line.split()] for line in open(glob.glob(os.path.join(tmp, "Data", "Moments_*"))[0]) if line.strip()])
assert U.size == 16 * 16 ** 3 * 6
np.savez_compressed(os.path.join(HERE, "ref_test1.npz"), stride=STRIDE, U_sample=U[::STRIDE], U_sum=U.sum(), U_abs_sum=np.abs(U).sum(),
                    U_sq_sum=(U * U).sum(), coeff_sums=U.reshape(-1, 6).sum(0), moments=mom)
print("wrote ref_test1.npz:", U[::STRIDE].size, "samples")
