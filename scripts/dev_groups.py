"""Developer timing (not the bench): ms per timestep of the 32-cell shard (Nv = N = 32) for LPGPU_GROUPS = 1, 2, 3, 4
(concurrent collision chains, api.cu), each in its own process because the knob is read once.
usage: python scripts/dev_groups.py [ncell] [steps]"""
import os, subprocess, sys
CHILD = r'''
import sys, os
sys.path.insert(0, os.getcwd())
import numpy as np, torch
import __graft_entry__ as graft
pkg = graft.load_package()
from lpsolver_b200 import solver
nc, steps = int(sys.argv[1]), int(sys.argv[2])
s = solver.ShardedSolver(nc, 32, 32, Lv=5.25, Lx=max(4.0, nc / 8.), nu=0.05, dt=0.01)
U0 = solver.set_init_ld(nc, 32, 5.25, max(4.0, nc / 8.), 0.5, np.pi / 2, True)
s.upload(U0); s.step(4)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize(); e0.record(); s.step(steps); e1.record(); torch.cuda.synchronize()
U = s.download()
print("groups=%s: %.4f ms/step  checksum %.17g" % (os.environ.get("LPGPU_GROUPS", "default"), e0.elapsed_time(e1) / steps, float(np.abs(U).sum())))
'''
nc = sys.argv[1] if len(sys.argv) > 1 else "32"
steps = sys.argv[2] if len(sys.argv) > 2 else "20"
for g in (sys.argv[3].split(",") if len(sys.argv) > 3 else ("1", "2", "3", "4")):
    env = dict(os.environ, LPGPU_GROUPS=g)
    r = subprocess.run([sys.executable, "-c", CHILD, nc, steps], env=env, capture_output=True, text=True)
    print(r.stdout.strip() or r.stderr[-800:])
