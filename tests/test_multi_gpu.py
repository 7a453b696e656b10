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
    mine = s.download()
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
    one.step(4)
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
