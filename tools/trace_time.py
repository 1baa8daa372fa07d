"""Time the (distributed) trace solve on a synthetic strip: python tools/trace_time.py nbx nby N [p] [tol]"""
import os, sys, time
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hybridsbp_b200 as hs
from hybridsbp_b200 import dist_trace
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dist = None
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
ctx = hs.Context(local)
nbx, nby, N = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
p = int(sys.argv[4]) if len(sys.argv) > 4 else 4
tol = float(sys.argv[5]) if len(sys.argv) > 5 else 1e-10
t0 = time.time()
dt, g, gd, info = dist_trace.build_strip_problem(ctx, rank, world, nbx, nby, N, p, dist=dist)
torch.cuda.synchronize(); t1 = time.time()
lam, u, st = dt.solve(g, gd, tol=tol, maxit=5000)
torch.cuda.synchronize(); t2 = time.time()
print("rank %d: blocks %d N %d p %d local_mode %d lambda %d cut %d | setup %.2f s solve %.3f s outer %d conv %d res %.1e" %
      (rank, info["blocks"], N, p, info["local_mode"], info["lambda_points"], info["cut_faces"], t1 - t0, t2 - t1,
       st["outer_iterations"], st["converged"], st["rel_residual"]), flush=True)
if dist is not None:
    dist.barrier(); dist.destroy_process_group()
