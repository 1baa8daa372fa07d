"""Hybrid trace-CG solve on the synthetic warped mesh with blocks of 256 x 256 points (BASELINE config 4: 32 x 32 blocks).
usage: python tools/trace_c4.py [nbx] [nby] [tol] [condense 0/1]   -- one GPU; prints one JSON line."""
import json, sys, time
import numpy as np
sys.path.insert(0, ".")
import torch
import hybridsbp_b200 as hs
from hybridsbp_b200 import dist_trace
nbx = int(sys.argv[1]) if len(sys.argv) > 1 else 32
nby = int(sys.argv[2]) if len(sys.argv) > 2 else nbx
tol = float(sys.argv[3]) if len(sys.argv) > 3 else 1e-10
condense = (sys.argv[4] != "0") if len(sys.argv) > 4 else True
N, p = 255, 4
ctx = hs.Context(0)
torch.cuda.set_device(0)
t0 = time.perf_counter()
dt, g, gd, info = dist_trace.build_strip_problem(ctx, 0, 1, nbx, nby, N, p, condense=condense)
torch.cuda.synchronize()
t_setup = time.perf_counter() - t0
t0 = time.perf_counter()
lam, u, st = dt.solve(g, gd, tol=tol, maxit=100000)
torch.cuda.synchronize()
t_solve = time.perf_counter() - t0
# residuals of the coupled system [M Fbar; Fbar^T D][u; lam] = [g; gd] with the matrix-free operators (independent of S_e)
blk, tr = info["blk"], info["tr"]
from hybridsbp_b200.parallel import _Ptr
ctx.sync()
Mu = torch.empty_like(u); Fl = torch.zeros_like(u); FTu = torch.empty_like(lam)
torch.cuda.synchronize()
blk.apply(_Ptr(u), _Ptr(Mu)); tr.Fbar_add(_Ptr(lam), 1.0, _Ptr(Fl)); tr.FbarT(_Ptr(u), _Ptr(FTu)); ctx.sync()
D = torch.as_tensor(tr.D(), device=u.device)
res_vol = float(torch.linalg.norm(g - Mu - Fl) / torch.linalg.norm(g))
res_lam = float(torch.linalg.norm(gd - FTu - D * lam) / torch.linalg.norm(gd))
print(json.dumps({"blocks": nbx * nby, "points_per_block": (N + 1) ** 2, "p": p, "lambda_points": info["lambda_points"],
                  "volume_points": info["volume_points"], "setup_seconds": t_setup, "solve_seconds": t_solve,
                  "outer_iterations": st["outer_iterations"], "converged": st["converged"],
                  "rel_residual": st["rel_residual"], "tol": tol, "local_solver": info["local_mode"], "condensed": condense,
                  "check_rel_residual_volume_equations": res_vol, "check_rel_residual_trace_equations": res_lam}))
