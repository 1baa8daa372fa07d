"""Hybrid trace-CG solve on the synthetic warped mesh with blocks of 256 x 256 points (BASELINE config 4: 32 x 32 blocks) through
hsbp_trace_solve.  usage: python tools/trace_c4.py [nbx] [nby] [tol] [condense 0/1]   -- one GPU; prints one JSON line with the
independent residuals of the coupled system [M Fbar; Fbar^T D][u; lam] = [g; gd] (matrix-free operators, not the condensed S_e)."""
import json, sys, time
import numpy as np
sys.path.insert(0, ".")
import hybridsbp_b200 as hs
from hybridsbp_b200 import dist_trace
nbx = int(sys.argv[1]) if len(sys.argv) > 1 else 32
nby = int(sys.argv[2]) if len(sys.argv) > 2 else nbx
tol = float(sys.argv[3]) if len(sys.argv) > 3 else 1e-10
condense = (sys.argv[4] != "0") if len(sys.argv) > 4 else True
N, p = 255, 4
ctx = hs.Context(0)
tm = {}
t0 = time.perf_counter()
pr = dist_trace.StripProblem(ctx, 0, 1, nbx, nby, N, p, condense=condense, timings=tm)
ctx.sync()
t_setup = time.perf_counter() - t0
pr.solve(tol=1e-2, maxit=4)
t0 = time.perf_counter()
st = pr.solve(tol=tol, maxit=100000)
ctx.sync()
t_solve = time.perf_counter() - t0
blk, tr = pr.blk, pr.tr
Mu, Fl, FTu = ctx.empty(blk.VNp), ctx.array(np.zeros(blk.VNp)), ctx.empty(tr.lNp)
blk.apply(pr.u, Mu); tr.Fbar_add(pr.lam, 1.0, Fl); tr.FbarT(pr.u, FTu); ctx.sync()
g, gd, lam = pr.g.get(), pr.gd.get(), pr.lam.get()
res_vol = float(np.linalg.norm(g - Mu.get() - Fl.get()) / np.linalg.norm(g))
res_lam = float(np.linalg.norm(gd - FTu.get() - tr.D() * lam) / np.linalg.norm(gd))
print(json.dumps({"blocks": nbx * nby, "points_per_block": (N + 1) ** 2, "p": p, "lambda_points": pr.info["lambda_points"],
                  "volume_points": pr.info["volume_points"], "setup_seconds": t_setup,
                  "setup_breakdown_seconds": {k: round(v, 3) for k, v in tm.items()}, "solve_seconds": t_solve,
                  "outer_iterations": st["outer_iterations"], "converged": st["converged"], "cg_loop_ms": st["cg_loop_ms"],
                  "rel_residual": st["rel_residual"], "true_rel_residual": st["true_rel_residual"], "tol": tol,
                  "local_solver": pr.info["local_mode"], "condensed": condense,
                  "check_rel_residual_volume_equations": res_vol, "check_rel_residual_trace_equations": res_lam}))
