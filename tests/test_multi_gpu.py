"""Sharded solver: both exchanges on >= 2 GPUs (skipped otherwise), and the peer-memory exchange between two processes that
share one GPU (runs on the single-GPU box too)."""
import os
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("world", [2])
def test_sharded_solver_matches_single_gpu(world):
    import torch
    if torch.cuda.device_count() < world:
        pytest.skip("needs %d GPUs" % world)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
           "--master-addr", "127.0.0.1", "--master-port", "29611", os.path.join(ROOT, "scripts", "mgpu_check.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=240, cwd=ROOT)
    assert "MGPU_OK" in out.stdout, out.stdout[-2000:] + out.stderr[-2000:]


def _peer_worker(rank, world, port, q):
    import numpy as np
    import torch
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    import __graft_entry__ as graft
    graft.load_package()
    from lpsolver_b200 import solver
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.cuda.set_device(0)
    cfg = dict(Nx=8, Nv=8, N=8, Lv=5.25, Lx=4.0, nu=0.05, dt=0.01)
    U0 = solver.set_init_ld(cfg["Nx"], cfg["Nv"], cfg["Lv"], cfg["Lx"], 0.5, np.pi / 2, True)
    n = U0.size // world
    s = solver.ShardedSolver(rank=rank, world=world, device=0, dist=dist, exchange="peer", **cfg)
    s.upload(U0[rank * n:(rank + 1) * n])
    s.step(2)                                   # eager, captured
    s.step(2)                                   # replays
    mine = s.step_host(s.download())            # a fifth step on the host-resident shard (lpgpu_step_host: peers inside)
    s.close()                                   # raises if a wait for the peer ever timed out
    q.put((rank, mine))
    dist.barrier()
    dist.destroy_process_group()


def test_peer_exchange_between_two_processes_on_one_gpu():
    """lpgpu_peer_export / _import with CUDA IPC: two processes, one shard each, both on cuda:0 (time-sliced).  The halo
    planes and densities cross through kernel writes into the other process's buffers; the result must equal the unsharded
    context bit for bit (the collision kernels see 4 cells per launch instead of 8, which changes no sum)."""
    import numpy as np
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 1:
        pytest.skip("needs a GPU")
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29700 + (os.getpid() % 200)
    procs = [ctx.Process(target=_peer_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=180) for _ in procs)
    for p in procs:
        p.join(60)
    import __graft_entry__ as graft
    graft.load_package()
    from lpsolver_b200 import solver
    cfg = dict(Nx=8, Nv=8, N=8, Lv=5.25, Lx=4.0, nu=0.05, dt=0.01)
    U0 = solver.set_init_ld(cfg["Nx"], cfg["Nv"], cfg["Lv"], cfg["Lx"], 0.5, np.pi / 2, True)
    one = solver.ShardedSolver(device=0, **cfg)
    one.upload(U0)
    one.step(5)
    want = one.download()
    one.close()
    got = np.concatenate([res[0], res[1]])
    err = np.max(np.abs((got - U0) - (want - U0))) / np.max(np.abs(want - U0))
    assert err < 1e-12, err



def _timeout_worker(rank, world, port, q):
    import time
    import numpy as np
    import torch
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    import __graft_entry__ as graft
    pkg = graft.load_package()
    from lpsolver_b200 import solver
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.cuda.set_device(0)
    cfg = dict(Nx=8, Nv=8, N=8, Lv=5.25, Lx=4.0, nu=0.05, dt=0.01)
    U0 = solver.set_init_ld(cfg["Nx"], cfg["Nv"], cfg["Lv"], cfg["Lx"], 0.5, np.pi / 2, True)
    n = U0.size // world
    s = solver.ShardedSolver(rank=rank, world=world, device=0, dist=dist, exchange="peer", **cfg)
    s.g.peer_set_timeout(1.0)
    s.upload(U0[rank * n:(rank + 1) * n])
    stepped = closed = "ok"
    if rank == 0:                               # rank 1 never steps: rank 0's first wait for its densities must time out
        try:
            s.step(1)
        except pkg.lpgpu.LPGpuError as e:
            stepped = "error: " + str(e)
    else:
        time.sleep(4.0)
    try:
        s.close()                               # all ranks learn about the failure and raise together
    except pkg.lpgpu.LPGpuError as e:
        closed = "error: " + str(e)
    q.put((rank, stepped, closed))
    dist.barrier()
    dist.destroy_process_group()


def test_peer_wait_timeout_fails_the_step_on_every_rank():
    """A rank that dies or falls behind by more than the bound must fail the run, not let it compute with stale halo
    planes: lpgpu_step returns an error on the waiting rank and ShardedSolver.close() raises on every rank."""
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 1:
        pytest.skip("needs a GPU")
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29900 + (os.getpid() % 90)
    procs = [ctx.Process(target=_timeout_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = {r: (a, b) for r, a, b in (q.get(timeout=180) for _ in procs)}
    for p in procs:
        p.join(60)
    assert res[0][0].startswith("error") and "timed out" in res[0][0]
    assert res[0][1].startswith("error") and res[1][1].startswith("error")


def test_cpp_driver_with_two_ranks_matches_one_rank(tmp_path):
    """`lpsolver --ranks 2`: two forked processes, one shard each (on a one-GPU box both on cuda:0), CUDA IPC handles exchanged
    over socket pairs, diagnostics reduced on rank 0, which writes the files.  The final state must equal the one-rank run bit
    for bit (the sharded advection reproduces the unsharded one exactly and collisions are per cell), and the Moments rows the
    golden of the reference's test 0."""
    import json, shutil
    import numpy as np
    exe = os.path.join(ROOT, "landau-poisson-solver_b200", "host", "lpsolver")
    if not os.path.exists(exe):
        pytest.skip("host driver not built")
    deck = os.path.join(ROOT, "tests", "golden", "LPsolver-input-test0.txt")
    out = {}
    for ranks in (1, 2):
        d = tmp_path / ("r%d" % ranks)
        d.mkdir()
        shutil.copy(deck, d / "LPsolver-input.txt")
        r = subprocess.run([exe, "--quiet", "--ranks", str(ranks)], cwd=d, capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stdout[-1500:] + r.stderr[-1500:]
        names = sorted(os.listdir(d / "Data"))
        out[ranks] = {n: open(d / "Data" / n, "rb").read() for n in names}
    assert sorted(out[1]) == sorted(out[2])
    uname = [n for n in out[1] if n.startswith("U_")][0]
    assert np.array_equal(np.frombuffer(out[1][uname], dtype=np.float64), np.frombuffer(out[2][uname], dtype=np.float64))
    mname = [n for n in out[1] if n.startswith("Moments_")][0]
    rows = [[float(x) for x in line.split()] for line in out[2][mname].decode().splitlines() if line.strip()]
    gold = json.load(open(os.path.join(ROOT, "tests", "golden", "reference_moments.json")))["Moments_Test0.dc"]
    assert len(rows) == 6
    for row, grow in zip(rows, gold):
        for col in (0, 4, 5, 6, 7, 8):
            assert abs(row[col] - grow[col]) <= 1.5e-7 * max(1.0, abs(grow[col])), (row, grow, col)
    for kind in ("Marginals_", "PhiVals_", "EntropyVals_"):
        n = [x for x in out[1] if x.startswith(kind)][0]
        a = np.array([[float(x) for x in line.split()] for line in out[1][n].decode().splitlines() if line.strip()])
        b = np.array([[float(x) for x in line.split()] for line in out[2][n].decode().splitlines() if line.strip()])
        assert a.shape == b.shape and np.allclose(a, b, rtol=1e-7, atol=1e-12), kind
