"""Run under torchrun on >= 2 GPUs: the distributed trace solve must reproduce the single-GPU solve of the same
global mesh (every rank also solves the whole problem on its own GPU and compares its part)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hybridsbp_b200 as hs                                   # noqa: E402
from hybridsbp_b200 import dist_trace                          # noqa: E402

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
ctx = hs.Context(local)
p, N, nbx, nby = 4, int(sys.argv[1]) if len(sys.argv) > 1 else 31, 2, 3
dt, g, gd, info = dist_trace.build_strip_problem(ctx, rank, world, nbx, nby, N, p, dist=dist if world > 1 else None)
lam, u, st = dt.solve(g, gd, tol=1e-12, maxit=2000)
# reference: the whole mesh on this GPU (world = 1 problem with nbx * world columns)
dt1, g1, gd1, info1 = dist_trace.build_strip_problem(ctx, 0, 1, nbx * world, nby, N, p, dist=None)
lam1, u1, st1 = dt1.solve(g1, gd1, tol=1e-12, maxit=2000)
lm, lm1 = info["lm"], info1["lm"]
s1 = info1["tr"].FTolambdastarts
rows = np.concatenate([np.arange(s1[f] - 1, s1[f + 1] - 1) for f in lm.faces])
npb = (N + 1) ** 2
cols = np.concatenate([np.arange(e * npb, (e + 1) * npb) for e in lm.blocks])
el = float(torch.linalg.norm(lam - lam1[rows]) / torch.linalg.norm(lam1))
eu = float(torch.linalg.norm(u - u1[cols]) / torch.linalg.norm(u1))
print("rank %d/%d: cut faces %d, outer iterations %d (single GPU %d), |dlam| %.2e |du| %.2e" %
      (rank, world, info["cut_faces"], st["outer_iterations"], st1["outer_iterations"], el, eu), flush=True)
assert st["converged"] == 1 and el < 1e-9 and eu < 1e-9
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
