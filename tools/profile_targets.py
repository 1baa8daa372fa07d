"""Small workloads for ncu captures of the kernels besides k_sweep:
  chol   K2a: dense Cholesky setup + solves on config 1 level 2 (56 blocks x 35x35 points, ld = 1248)
  pcg    K2b: a batched Jacobi-PCG local solve on 64 blocks x 64x64 points
  trace  K3/K4: a short trace solve (face gather / scatter, lambda kernels)"""
import sys
import numpy as np
sys.path.insert(0, ".")
import hybridsbp_b200 as hs
from hybridsbp_b200 import square_circle as sc, dist_trace

what = sys.argv[1] if len(sys.argv) > 1 else "chol"
ctx = hs.Context(0)
if what == "chol":
    mesh = sc.load_mesh(sc.default_mesh_path())
    r = sc.solve_level(ctx, mesh, 4, 34, local_mode=hs.LOCAL_CHOLESKY, tol=1e-6, maxit=20)
    print("chol: outer iterations", r["stats"]["outer_iterations"])
else:
    pr = dist_trace.StripProblem(ctx, 0, 1, 8, 8, 63, 4, local_mode=hs.LOCAL_PCG, condense=(what == "trace"),
                                 coarse_modes=2 if what == "trace" else 0)
    st = pr.solve(tol=1e-3, maxit=3)
    print("pcg/trace:", st)
