"""CPU tests of the host side: the C ABI library loads and exports every declared symbol, the
input-file reader and output formatting mirror the reference driver, the initial conditions match
the oracle, and the sharding/exchange logic is right for world_size 2 (gloo)."""
import os
import re
import sys

import numpy as np
import pytest

from conftest import ROOT, relerr
from oracle.oracle import PortOracle

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def test_library_exports_every_declared_symbol(pkg):
    L = pkg.lpgpu.load_library()
    header = open(os.path.join(ROOT, "include", "lpgpu.h")).read()
    declared = sorted(set(re.findall(r"\b(lpgpu_[A-Za-z0-9_]+)\s*\(", header)))
    assert declared, "no declarations found in include/lpgpu.h"
    for name in declared:
        assert hasattr(L, name), "liblpgpu.so does not export " + name
    assert sorted(pkg.lpgpu.EXPORTS) == declared


def test_no_cpu_fallback(pkg):
    L = pkg.lpgpu.load_library()
    if L.lpgpu_device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(pkg.LPGpuError, match="no CUDA device"):
        pkg.LPGpu(4, 8, 8, 5.25, 12.5, 0.05, 0.01)


def test_init_refuses_bad_parameters(pkg):
    """Error behaviour of the boundary: the reference prints and exit(1)s on bad input (InputParsing.cpp); the ABI returns
    LPGPU_EINVAL with a message, before any device call, so this runs without a GPU."""
    base = dict(Nx=4, Nv=8, N=8, Lv=5.25, Lx=4.0, nu=0.05, dt=0.01)
    bad = [dict(base, N=7), dict(base, N=34), dict(base, Nv=7), dict(base, dt=0.), dict(base, Lv=-1.), dict(base, gamma=2),
           dict(base, gamma=0, full_and_linear=True), dict(base, linear_landau=True, full_and_linear=True),
           dict(base, x_begin=2, x_count=3), dict(base, doping=dict(NL=0.001, NH=1., eps=0., T_L=0.4, T_R=0.4))]
    for kw in bad:
        with pytest.raises(pkg.lpgpu.LPGpuError) as e:
            pkg.LPGpu(**kw)
        assert "lpgpu_init" in str(e.value), (kw, str(e.value))
    import ctypes as C
    L = pkg.lpgpu.load_library()
    assert L.lpgpu_finalize(None) == 0                       # finalize(NULL) is a no-op, like free(NULL)
    assert L.lpgpu_step(None, 1) != 0 and b"null context" in L.lpgpu_last_error()
    assert L.lpgpu_peer_export(None, None) != 0


def test_product_never_imports_oracle():
    pkgdir = os.path.join(ROOT, "landau-poisson-solver_b200")
    for dirpath, _, files in os.walk(pkgdir):
        for f in files:
            if f.endswith((".py", ".cu", ".cpp", ".h", ".c")) or f == "Makefile":
                txt = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in txt.lower().replace("# oracle", ""), os.path.join(dirpath, f)


def test_input_file_and_output_names(pkg):
    from lpsolver_b200 import solver
    c0 = solver.RunConfig.from_file(os.path.join(GOLD, "LPsolver-input-test0.txt"))
    assert (c0.nT, c0.Nx, c0.Nv, c0.N, c0.nu, c0.dt, c0.gamma) == (5, 16, 16, 8, 0.05, 0.01, -3)
    assert c0.ic == "Damping" and not c0.homogeneous and c0.A_amp == 0.2 and c0.k_wave == 0.5 and c0.Lv == 5.25
    assert abs(c0.Lx - 4 * np.pi) < 1e-15                       # default Lx = 2 pi / k_wave
    # the file name the reference's bats test expects (tests/LPsolver_tests:13)
    assert c0.moments_filename() == "Data/Moments_nu0.05A0.2k0.5Nx16Lx12.5664Nv16Lv5.25SpectralN8dt0.01nT5_Test0.dc"
    c4 = solver.RunConfig.from_file(os.path.join(GOLD, "LPsolver-input-test4.txt"))
    assert c4.homogeneous and c4.ic == "FourHump"
    assert c4.moments_filename() == "Data/Moments_nu0.05A0k0.5Nv16Lv5.25SpectralN8dt0.01nT5_Test4.dc"
    row = solver.format_moments_row([12.566371, 3.9e-13, -1.4e-17, 3.4e-16, 7.541212, 0.50124002], False)
    assert len(row.split()) == 9 and row.split()[0] == "12.566371"


def test_unsupported_options_fail_loudly(pkg, tmp_path):
    from lpsolver_b200 import solver
    txt = open(os.path.join(GOLD, "LPsolver-input-test0.txt")).read().replace("Damping          = True", "Damping          = False").replace(
        "TwoHump          = False", "TwoHump          = True")
    p = tmp_path / "in.txt"
    p.write_text(txt)
    with pytest.raises(NotImplementedError):
        solver.RunConfig.from_file(str(p))
    p.write_text("gamma = 2\n" + open(os.path.join(GOLD, "LPsolver-input-test0.txt")).read())
    with pytest.raises(ValueError):             # ReadGamma accepts -3, 0 and 1 only (InputParsing.cpp:202-238)
        solver.RunConfig.from_file(str(p))
    p.write_text("gamma = 1\n" + open(os.path.join(GOLD, "LPsolver-input-test0.txt")).read())
    assert solver.RunConfig.from_file(str(p)).gamma == 1


def test_reference_test1_deck_is_parsed(pkg):
    """Doping + LinearLandau + MassConsOnly (tests/LPsolver-input-test1.txt): flags, [Doping] section, file name,
    and SetInit_ND against the oracle."""
    from lpsolver_b200 import solver
    c1 = solver.RunConfig.from_file(os.path.join(GOLD, "LPsolver-input-test1.txt"))
    assert c1.ic == "Doping" and c1.linear_landau and c1.mass_cons_only and not c1.full_and_linear
    assert c1.doping == dict(NL=0.001, NH=1.0, eps=0.1, T_L=0.4, T_R=0.4)
    assert c1.moments_filename() == "Data/Moments_nu0.05A0k0.5Nx16Lx12.5664Nv16Lv5.25SpectralN8dt0.01nT5_Test1.dc"
    P = PortOracle(Nx=c1.Nx, Nv=c1.Nv, N=c1.N, Lv=c1.Lv, Lx=c1.Lx, nu=c1.nu, dt=c1.dt)
    P.set_doping(**c1.doping)
    assert relerr(c1.initial_condition(), P.SetInit_ND()) < 1e-13
    assert relerr(c1.initial_condition(4, 8), P.SetInit_ND().reshape(c1.Nx, -1)[4:12].reshape(-1)) < 1e-13
    assert list(solver.doping_profile(16, 0.001, 1.)) == [1.] * 5 + [0.001] * 5 + [1.] * 6


def test_initial_conditions_match_oracle(pkg):
    from lpsolver_b200 import solver
    cfg = dict(Nx=6, Nv=8, N=8, Lv=5.25, Lx=4 * np.pi, nu=0.05, dt=0.01)
    P = PortOracle(**cfg)
    assert relerr(solver.set_init_ld(6, 8, 5.25, 4 * np.pi, 0.2, 0.5), P.SetInit_LD(0.2, 0.5)) < 1e-14
    assert relerr(solver.set_init_4h(6, 8, 5.25, 4 * np.pi), P.SetInit_4H()) < 1e-13
    P2 = PortOracle(**dict(cfg, Lx=4.0))
    assert relerr(solver.set_init_ld(6, 8, 5.25, 4.0, 0.5, np.pi / 2, True), P2.SetInit_LD(0.5, np.pi / 2, True)) < 1e-14
    Ph = PortOracle(homogeneous=True, **cfg)
    assert relerr(solver.set_init_4h_homo(8, 5.25), Ph.SetInit_4H_Homo()) < 1e-13
    # a shard of the IC equals the slice of the full IC
    full = solver.set_init_ld(6, 8, 5.25, 4 * np.pi, 0.2, 0.5)
    part = solver.set_init_ld(6, 8, 5.25, 4 * np.pi, 0.2, 0.5, False, 2, 3)
    assert np.array_equal(part, full[2 * 512 * 6:5 * 512 * 6])


def test_shard_range(pkg):
    from lpsolver_b200 import solver
    assert [solver.shard_range(256, 8, r) for r in (0, 3, 7)] == [(0, 32), (96, 32), (224, 32)]
    with pytest.raises(ValueError):
        solver.shard_range(10, 4, 0)


def _gloo_worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    import __graft_entry__ as graft
    solver = __import__("importlib").import_module("lpsolver_b200.solver") if graft.load_package() else None
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    Nx, plane = 8, 12
    x0, n = solver.shard_range(Nx, world, rank)
    # fake state: plane p of global cell i holds the value i; per-cell sums (m_i, s_i) = (i, -i)
    send_left = torch.full((plane,), float(x0), dtype=torch.float64)
    send_right = torch.full((plane,), float(x0 + n - 1), dtype=torch.float64)
    recv_left, recv_right = torch.zeros(plane, dtype=torch.float64), torch.zeros(plane, dtype=torch.float64)
    ms_local = torch.tensor([v for i in range(x0, x0 + n) for v in (float(i), -float(i))], dtype=torch.float64)
    ms_all = torch.zeros(2 * Nx, dtype=torch.float64)
    solver.exchange_stage(dist, rank, world, ms_local, ms_all, send_left, send_right, recv_left, recv_right)
    ok = (recv_left == float((x0 - 1) % Nx)).all() and (recv_right == float((x0 + n) % Nx)).all()
    ok = ok and ms_all.tolist() == [v for i in range(Nx) for v in (float(i), -float(i))]
    # the order the sharded driver uses: halo planes on their own communicator first, then the density all-gather
    grp = dist.new_group(backend="gloo")
    recv_left.zero_(); recv_right.zero_(); ms_all.zero_()
    solver.exchange_halo(dist, rank, world, send_left, send_right, recv_left, recv_right, group=grp)
    solver.exchange_density(dist, world, ms_local, ms_all)
    ok = ok and (recv_left == float((x0 - 1) % Nx)).all() and (recv_right == float((x0 + n) % Nx)).all()
    ok = ok and ms_all.tolist() == [v for i in range(Nx) for v in (float(i), -float(i))]
    q.put((rank, bool(ok)))
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4])
def test_exchange_stage_gloo(pkg, world):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + world + (os.getpid() % 500)
    procs = [ctx.Process(target=_gloo_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(60)
    assert sorted(res) == [(r, True) for r in range(world)]
