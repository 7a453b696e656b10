"""Run under torchrun: the x-sharded solver must reproduce the single-GPU solver (to round-off: 1e-12 of the update),
with both exchanges -- "peer" (CUDA IPC: kernels write halo planes and densities into the peers' memory, the timestep is
one graph) and "nccl" (all-gather + send/recv per stage through torch.distributed).  Prints MGPU_OK on rank 0."""
import os, sys
import numpy as np
import torch
import torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as graft
graft.load_package()
from lpsolver_b200 import solver

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
cfg = dict(Nx=8 * world, Nv=8, N=8, Lv=5.25, Lx=4.0, nu=0.05, dt=0.01)
U0 = solver.set_init_ld(cfg["Nx"], cfg["Nv"], cfg["Lv"], cfg["Lx"], 0.5, np.pi / 2, True)
n = U0.size // world
NSTEPS = 7                     # eager, captured, five replays
got, mom = {}, {}
for mode in ("peer", "nccl"):
    s = solver.ShardedSolver(rank=rank, world=world, device=local, dist=dist, exchange=mode, **cfg)
    s.upload(U0[rank * n:(rank + 1) * n])
    s.step(3)
    s.step(NSTEPS - 3)
    mine = torch.from_numpy(s.download()).cuda()
    allU = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(allU, mine)
    got[mode] = torch.cat(allU).cpu().numpy()
    mom[mode] = s.moments()
    s.close()
if rank == 0:
    one = solver.ShardedSolver(device=local, **cfg)
    one.upload(U0)
    one.step(NSTEPS)
    want = one.download()
    mom1 = one.moments()
    one.close()
    ok = True
    for mode in ("peer", "nccl"):
        # not bit-identical to one GPU on purpose: the ComputeQ launch shape depends on the cells per launch
        err = np.max(np.abs((got[mode] - U0) - (want - U0))) / np.max(np.abs(want - U0))
        print("%s: rel err of the %d-step update = %.3e, moments diff = %.3e" % (mode, NSTEPS, err, np.max(np.abs(mom[mode] - mom1))))
        ok = ok and err < 1e-12 and np.allclose(mom[mode], mom1, rtol=1e-13, atol=1e-15)
    print("peer == nccl bit for bit:", bool(np.array_equal(got["peer"], got["nccl"])))
    print("MGPU_OK" if ok else "MGPU_FAIL")
dist.destroy_process_group()
