"""CPU check of the FFT-convolution ComputeQ pipeline (csrc/fc3.cuh).  The per-thread phase functions of
the CUDA kernels are __host__ __device__; tests/emul/fc3_emul.cpp runs them under a thread-loop emulator
(every barrier-separated phase for all thread ids, CTA by CTA) and this test compares the result with the
plain O(N^6) sum of collisionRoutines_1.cpp:706-773 -- the index algebra and the transform identities are
thereby covered in the container without a GPU.  (The GPU parity tests compare the real kernels.)"""
import ctypes
import os
import subprocess

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "emul", "fc3_emul.cpp")
LIB = os.path.join(HERE, "emul", "libfc3_emul.so")
HDR = os.path.join(HERE, "..", "landau-poisson-solver_b200", "csrc", "fc3.cuh")
P = ctypes.POINTER(ctypes.c_double)


@pytest.fixture(scope="module")
def emul():
    stale = not os.path.exists(LIB) or any(os.path.getmtime(LIB) < os.path.getmtime(p) for p in (SRC, HDR))
    if stale:
        subprocess.check_call(["g++", "-O2", "-fopenmp", "-std=c++17", "-shared", "-fPIC", "-o", LIB, SRC])
    return ctypes.CDLL(LIB)


@pytest.mark.parametrize("N,B", [(8, 3), (16, 2), (24, 1), (32, 1)])
def test_pipeline_matches_direct_sum(emul, N, B):
    rng = np.random.default_rng(N)
    fh = rng.standard_normal((B, N ** 3, 2))
    G = rng.standard_normal((N ** 3, 7))
    E = (np.arange(N) - N / 2) * 0.37
    got, want = np.zeros_like(fh), np.zeros_like(fh)
    assert emul.fc3_emulate(N, B, fh.ctypes.data_as(P), G.ctypes.data_as(P), E.ctypes.data_as(P), got.ctypes.data_as(P)) == 0
    assert emul.fc3_direct(N, B, fh.ctypes.data_as(P), G.ctypes.data_as(P), E.ctypes.data_as(P), want.ctypes.data_as(P)) == 0
    assert np.abs(got - want).max() < 1e-13 * np.abs(want).max()


@pytest.mark.parametrize("N,B", [(8, 2), (16, 1), (24, 1), (32, 1)])
def test_quarter_variant_matches_direct_sum(emul, N, B):
    """F2Q (k_fc3_f2q): the y/x stage on three-warp CTAs -- one array at a time, the x stage as two halves per line
    (M/2-point transforms), the inverse x combined from the halves on the way into the inverse y."""
    rng = np.random.default_rng(50 + N)
    fh = rng.standard_normal((B, N ** 3, 2))
    G = rng.standard_normal((N ** 3, 7))
    E = (np.arange(N) - N / 2) * 0.37
    got, want = np.zeros_like(fh), np.zeros_like(fh)
    assert emul.fc3_emulate_quarter(N, B, fh.ctypes.data_as(P), G.ctypes.data_as(P), E.ctypes.data_as(P), got.ctypes.data_as(P)) == 0
    assert emul.fc3_direct(N, B, fh.ctypes.data_as(P), G.ctypes.data_as(P), E.ctypes.data_as(P), want.ctypes.data_as(P)) == 0
    assert np.abs(got - want).max() < 1e-13 * np.abs(want).max()


def test_product_split_over_three_ctas(emul):
    """F2 may split its seven products over three CTAs per (cell, kz) when few cells are in flight; F3 sums the parts."""
    N, B = 16, 1
    rng = np.random.default_rng(3)
    fh = rng.standard_normal((B, N ** 3, 2))
    G = rng.standard_normal((N ** 3, 7))
    E = (np.arange(N) - N / 2) * 0.37
    got, want = np.zeros_like(fh), np.zeros_like(fh)
    assert emul.fc3_emulate_split(N, B, fh.ctypes.data_as(P), G.ctypes.data_as(P), E.ctypes.data_as(P), got.ctypes.data_as(P), 3) == 0
    assert emul.fc3_direct(N, B, fh.ctypes.data_as(P), G.ctypes.data_as(P), E.ctypes.data_as(P), want.ctypes.data_as(P)) == 0
    assert np.abs(got - want).max() < 1e-13 * np.abs(want).max()


@pytest.mark.parametrize("N,B", [(8, 2), (16, 1)])
def test_linear_operator_pipeline_matches_direct_sum(emul, N, B):
    """LinearLandau: Q(f, M) -- F1 builds the seven u arrays from the stored Maxwellian transform, the v arrays from
    fhat (ComputeQLinear, collisionRoutines_1.cpp:1185-1269)."""
    rng = np.random.default_rng(100 + N)
    fh = rng.standard_normal((B, N ** 3, 2))
    mh = rng.standard_normal((B, N ** 3, 2))
    G = rng.standard_normal((N ** 3, 7))
    E = (np.arange(N) - N / 2) * 0.37
    got, want = np.zeros_like(fh), np.zeros_like(fh)
    assert emul.fc3_emulate_linear(N, B, fh.ctypes.data_as(P), mh.ctypes.data_as(P), G.ctypes.data_as(P), E.ctypes.data_as(P), got.ctypes.data_as(P)) == 0
    assert emul.fc3_direct_linear(N, B, fh.ctypes.data_as(P), mh.ctypes.data_as(P), G.ctypes.data_as(P), E.ctypes.data_as(P), want.ctypes.data_as(P)) == 0
    assert np.abs(got - want).max() < 1e-13 * np.abs(want).max()
    assert emul.fc3_direct(N, B, fh.ctypes.data_as(P), G.ctypes.data_as(P), E.ctypes.data_as(P), got.ctypes.data_as(P)) == 0
    assert np.abs(got - want).max() > 1e-3 * np.abs(want).max()        # and it is a different operator


def test_unsupported_size_is_refused(emul):
    z = np.zeros(8)
    assert emul.fc3_emulate(10, 1, z.ctypes.data_as(P), z.ctypes.data_as(P), z.ctypes.data_as(P), z.ctypes.data_as(P)) == 1


@pytest.mark.parametrize("N", [8, 16, 24, 32])
@pytest.mark.parametrize("sign", [-1, 1])
def test_register_fft_matches_definition(emul, N, sign):
    rng = np.random.default_rng(N + sign)
    x = rng.standard_normal(N) + 1j * rng.standard_normal(N)
    xin = np.ascontiguousarray(np.stack([x.real, x.imag], axis=1))
    out = np.zeros_like(xin)
    assert emul.fc3_fftN(N, sign, xin.ctypes.data_as(P), out.ctypes.data_as(P)) == 0
    k = np.arange(N)
    want = np.exp(sign * 2j * np.pi * np.outer(k, k) / N) @ x
    assert np.abs(out[:, 0] + 1j * out[:, 1] - want).max() < 1e-14 * np.abs(want).max() * N


def test_line_halves_match_definition(emul):
    """fwd_half / inv_half (the x stage of k_fc3_f2q): X[2q + h] of a 48-point line whose last 16 inputs are zero, and the
    full inverse rebuilt from the two 24-point inverses."""
    rng = np.random.default_rng(7)
    x = rng.standard_normal(32) + 1j * rng.standard_normal(32)
    xin = np.ascontiguousarray(np.stack([x.real, x.imag], axis=1))
    out = np.zeros((48, 2))
    assert emul.fc3_half_lines(0, xin.ctypes.data_as(P), out.ctypes.data_as(P)) == 0
    X = np.exp(-2j * np.pi * np.outer(np.arange(48), np.arange(32)) / 48) @ x
    got = out[:, 0] + 1j * out[:, 1]
    for h in range(2):
        assert np.abs(got[h * 24:(h + 1) * 24] - X[h::2]).max() < 1e-13 * np.abs(X).max()
    Z = rng.standard_normal(48) + 1j * rng.standard_normal(48)
    zin = np.zeros((48, 2))
    for h in range(2):
        zin[h * 24:(h + 1) * 24, 0], zin[h * 24:(h + 1) * 24, 1] = Z[h::2].real, Z[h::2].imag
    assert emul.fc3_half_lines(1, zin.ctypes.data_as(P), out.ctypes.data_as(P)) == 0
    want = np.exp(2j * np.pi * np.outer(np.arange(48), np.arange(48)) / 48) @ Z
    assert np.abs(out[:, 0] + 1j * out[:, 1] - want).max() < 1e-13 * np.abs(want).max()
