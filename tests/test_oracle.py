"""CPU tests (no GPU): the oracle is pinned to the reference.

1. the C restatement (oracle/lp_oracle.c) against the committed golden vectors, which are outputs
   of the unmodified reference (tests/golden/make_golden.py), and against the reference's own
   golden moment files tests/Moments_Test0.dc / Moments_Test4.dc with moment_differ.sh's gates;
2. where oracle/_ref/libref.so exists (the reference compiled behind shims), element-wise against
   the live reference on further inputs."""
import json
import os

import numpy as np
import pytest

from conftest import relerr
from oracle.oracle import PortOracle, RefOracle, have_ref

GOLD = os.path.join(os.path.dirname(__file__), "golden")
TEST0 = dict(Nx=16, Nv=16, N=8, Lv=5.25, Lx=4 * np.pi, nu=0.05, dt=0.01)


@pytest.fixture(scope="module")
def vec():
    z = np.load(os.path.join(GOLD, "ref_vectors.npz"))
    return z, json.loads(str(z["cfg"]))


def test_port_matches_reference_vectors(vec):
    z, cfg = vec
    P = PortOracle(**cfg)
    U0 = z["U0"]
    assert relerr(P.setInit_spectral(U0), z["f"]) < 1e-15
    assert relerr(P.fft3D(z["x"]), z["fft3D"]) < 1e-14
    assert relerr(P.FS(z["x"]), z["FS"]) < 1e-14
    f1 = z["f"][1] * (1 + 0.1 * np.sin(np.arange(z["f"][1].size)))
    assert relerr(P.ComputeQ(f1), z["qHat"]) < 1e-14
    assert relerr(P.conserveMoments(z["qHat"]), z["qHat_conserved"]) < 1e-13
    assert np.max(np.abs(P.field(U0) - z["field"])) < 1e-12 * cfg["Lx"]
    assert relerr(P.moments(U0), z["moments"]) < 1e-13
    for name, got in (("U_collide", P.collide_step(U0)), ("U_RK3", P.RK3(U0)), ("U_step", P.step(U0))):
        assert relerr(got, z[name]) < 1e-13, name
        assert relerr(got - U0, z[name] - U0) < 1e-10, name
    Ph = PortOracle(homogeneous=True, **cfg)
    assert relerr(Ph.SetInit_4H_Homo(), z["Uh0"]) < 1e-15
    got = Ph.collide_step(z["Uh0"])
    assert relerr(got - z["Uh0"], z["Uh_collide"] - z["Uh0"]) < 1e-10


def test_separable_projection_equals_literal_IntModes(vec):
    z, cfg = vec
    P = PortOracle(**cfg)
    a = P.collide_step(z["U0"])
    P.set_direct_intmodes(1)
    b = P.collide_step(z["U0"])
    assert relerr(a - z["U0"], b - z["U0"]) < 1e-11


def _gate(row, gold):
    """tests/moment_differ.sh:9-13, row 6: mass 2e-6, momenta 1e-10 (absolute), total energy."""
    assert abs(row[0] - gold[0]) <= 2e-6
    for d in (1, 2, 3):
        assert abs(row[d] - gold[d]) <= 1e-10
    tot_col = len(gold) - 1
    assert row[tot_col] - gold[tot_col] <= 3e-5 and gold[tot_col] - row[tot_col] <= 5e-8


def test_port_reproduces_Moments_Test0():
    gold = json.load(open(os.path.join(GOLD, "reference_moments.json")))["Moments_Test0.dc"]
    P = PortOracle(**TEST0)
    U = P.SetInit_LD(0.2, 0.5)
    for step in range(6):
        m = P.moments(U)
        row = [m[0], m[1], m[2], m[3], m[4], m[5], np.sqrt(m[5]), np.log(np.sqrt(m[5])), m[4] + m[5]]
        for col in (0, 4, 5, 6, 7, 8):                      # every printed digit of every row
            assert abs(row[col] - gold[step][col]) <= 6e-8 * max(1.0, abs(gold[step][col])), (step, col)
        if step == 5:
            _gate(row, gold[5])
        U = P.step(U)


def test_port_reproduces_Moments_Test3_full_and_linear():
    """FullandLinear = True (ComputeQ_FandL, conserveAllMoments_FandL, RK4_FandL): every printed digit of the
    non-noise columns of all six rows of tests/Moments_Test3.dc -- the reference's own golden pins this variant."""
    gold = json.load(open(os.path.join(GOLD, "reference_moments.json")))["Moments_Test3.dc"]
    P = PortOracle(**TEST0)
    P.set_fandl(True)
    U = P.SetInit_LD(0.2, 0.5)
    for step in range(6):
        m = P.moments(U)
        row = [m[0], m[1], m[2], m[3], m[4], m[5], np.sqrt(m[5]), np.log(np.sqrt(m[5])), m[4] + m[5]]
        for col in (0, 4, 5, 6, 7, 8):
            assert abs(row[col] - gold[step][col]) <= 6e-8 * max(1.0, abs(gold[step][col])), (step, col)
        if step == 5:
            _gate(row, gold[5])
        U = P.step(U)
    # and it differs from the plain operator (Test0) where the golden files differ
    gold0 = json.load(open(os.path.join(GOLD, "reference_moments.json")))["Moments_Test0.dc"]
    assert abs(gold[5][4] - gold0[5][4]) > 5e-6


TEST1_DOPING = dict(NL=0.001, NH=1., eps=0.1, T_L=0.4, T_R=0.4)      # [Doping] section of LPsolver-input-test1.txt


def test_port_reproduces_Moments_Test1_doping_linear_massonly():
    """Reference test 1: Doping (SetInit_ND, *_Doping field integrals, Dirichlet walls in I3), LinearLandau
    (ComputeQLinear / RK4Linear against the transform of the initial state) and MassConsOnly.  Every printed digit of
    all six rows of tests/Moments_Test1.dc, and the final state the unmodified reference dumps (ref_test1.npz,
    tests/golden/make_test1_golden.py) element by element."""
    gold = json.load(open(os.path.join(GOLD, "reference_moments.json")))["Moments_Test1.dc"]
    z = np.load(os.path.join(GOLD, "ref_test1.npz"))
    P = PortOracle(**TEST0)
    P.set_doping(**TEST1_DOPING)
    U0 = P.SetInit_ND()
    P.set_linear_landau(U0)
    P.set_mass_cons_only(True)
    U = U0
    for step in range(6):
        m = P.moments(U)
        row = [m[0], m[1], m[2], m[3], m[4], m[5], np.sqrt(m[5]), np.log(np.sqrt(m[5])), m[4] + m[5]]
        for col in (0, 1, 4, 5, 6, 7, 8):                   # P1 is a real signal here (walls), not round-off
            assert abs(row[col] - gold[step][col]) <= 6e-8 * max(1.0, abs(gold[step][col])), (step, col)
        if step == 5:
            _gate([float("%.8g" % x) for x in row], gold[5])   # moment_differ.sh compares the printed files (%11.8g)
        else:
            U = P.step(U)
    st = int(z["stride"])
    assert relerr(U[::st], z["U_sample"]) < 1e-12
    assert relerr((U - U0)[::st], z["U_sample"] - U0[::st]) < 1e-10
    assert abs(U.sum() - float(z["U_sum"])) < 1e-10 * float(z["U_abs_sum"])


def test_port_reproduces_Moments_Test4():
    gold = json.load(open(os.path.join(GOLD, "reference_moments.json")))["Moments_Test4.dc"]
    P = PortOracle(homogeneous=True, **TEST0)
    U = P.SetInit_4H_Homo()
    for _ in range(5):
        U = P.step(U)
    m = P.moments(U)
    _gate([m[0], m[1], m[2], m[3], 0., 0., 0., m[4]], gold[5])
    assert abs(m[4] - gold[5][7]) <= 5e-8


@pytest.mark.ref
@pytest.mark.skipif(not have_ref(), reason="oracle/_ref/libref.so not built (needs /root/reference)")
def test_port_matches_live_reference():
    cfg = dict(Nx=4, Nv=8, N=8, Lv=5.25, Lx=4.0, nu=0.05, dt=0.01)
    P, R = PortOracle(**cfg), RefOracle(**cfg)
    for a, b in zip(P.grids(), R.grids()):
        assert np.array_equal(a, b)
    Cp, CCp = P.conservation()
    Cr, CCr = R.conservation()
    assert relerr(Cp, Cr) < 1e-15 and relerr(CCp, CCr) < 1e-13
    for xi in (0, 77, 292, 511):
        assert relerr(P.weight_row(xi), R.weight_row(xi)) < 1e-15
    for idx in [(0, 1, 2, 3, 4, 5), (4, 4, 4, 0, 0, 0), (7, 3, 4, 7, 7, 1)]:
        assert relerr(P.IntModes(*idx), R.IntModes(*idx)) < 1e-15
    assert relerr(P.SetInit_LD(0.5, np.pi / 2, True), R.SetInit_LD(0.5, np.pi / 2, True)) < 1e-15
    assert relerr(P.SetInit_4H(), R.SetInit_4H()) < 1e-15
    U = R.SetInit_LD(0.5, np.pi / 2, True)
    rng = np.random.default_rng(7)
    U = U * (1 + 0.1 * rng.standard_normal(U.shape)) + 1e-3 * rng.standard_normal(U.shape)
    assert np.max(np.abs(P.field(U) - R.field(U))) < 1e-12 * max(1.0, np.max(np.abs(R.field(U))))
    assert relerr(P.moments(U), R.moments(U)) < 1e-13
    a, b = P.step(U), R.step(U)
    assert relerr(a, b) < 1e-13 and relerr(a - U, b - U) < 1e-10


@pytest.mark.parametrize("gamma", [0, 1])
def test_port_matches_reference_for_other_kernels(gamma):
    """gamma = 0 (Maxwell molecules) and 1 (hard spheres), collisionRoutines_1.cpp:38-96, 116-157: no reference deck uses
    them, so the port is pinned by outputs of the unmodified reference (tests/golden/make_gamma_golden.py)."""
    z = np.load(os.path.join(GOLD, "ref_gamma.npz"))
    cfg = json.loads(str(z["cfg"]))
    t = "g%d_" % gamma
    P = PortOracle(gamma=gamma, **cfg)
    got = np.array([P.gHat3(p[:3], p[3:]) for p in z["pairs"]])
    assert relerr(got, z[t + "gHat3"]) < 1e-15
    assert relerr(P.ComputeQ(z[t + "f"]), z[t + "qHat"]) < 1e-14
    assert relerr(P.conserveMoments(z[t + "qHat"]), z[t + "qHat_conserved"]) < 1e-13
    U0 = z[t + "U0"]
    assert relerr(P.collide_step(U0) - U0, z[t + "U_collide"] - U0) < 1e-10
    Ph = PortOracle(homogeneous=True, gamma=gamma, **cfg)
    Uh = z[t + "Uh0"]
    assert relerr(Ph.collide_step(Uh) - Uh, z[t + "Uh_collide"] - Uh) < 1e-10
    if have_ref():
        R = RefOracle(gamma=gamma, **cfg)
        f = z[t + "f"] * (1 + 0.05 * np.cos(0.3 * np.arange(z[t + "f"].size)))
        assert relerr(P.ComputeQ(f), R.ComputeQ(f)) < 1e-14


def test_port_matches_reference_at_headline_size():
    """N = Nv = 32 (BASELINE configs 2, 4, 5): the C restatement against tests/golden/ref_n32.npz, outputs of the
    unmodified reference at that size (make_n32_golden.py: its 8.6 GB weight table, ~9 minutes on 8 threads)."""
    z = np.load(os.path.join(GOLD, "ref_n32.npz"))
    cfg = json.loads(str(z["cfg_h"]))
    P = PortOracle(homogeneous=True, **cfg)
    Uh = P.SetInit_4H_Homo()
    assert relerr(Uh, z["Uh0"]) < 1e-14
    f = P.setInit_spectral(z["Uh0"])[0]
    assert relerr(f, z["f_h"]) < 1e-15
    fa = z["f_h"] * (1 + 0.1 * np.sin(np.arange(f.size)))
    q = P.ComputeQ(fa)
    assert relerr(q, z["qHat"]) < 1e-14
    assert relerr(P.conserveMoments(z["qHat"]), z["qHat_conserved"]) < 1e-13
    got = P.collide_step(z["Uh0"])
    assert relerr(got, z["Uh_collide"]) < 1e-13
    assert relerr(got - z["Uh0"], z["Uh_collide"] - z["Uh0"]) < 1e-10
    m = P.moments(got)
    assert np.allclose(m[[0, 4]], z["moments_h1"][[0, 4]], rtol=1e-12) and np.all(np.abs(m[1:4] - z["moments_h1"][1:4]) < 1e-12)
