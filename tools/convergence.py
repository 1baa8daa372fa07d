"""Refinement sweeps of configs 1 and 2 on the GPU (square_circle.jl:204-428 and its flower_v2 counterpart): errors and
observed rates per level, wall time per level.  usage: python tools/convergence.py [levels] [cg tolerance]"""
import sys, time
import numpy as np
sys.path.insert(0, ".")
import hybridsbp_b200 as hs
from hybridsbp_b200 import square_circle as sc, flower
levels = int(sys.argv[1]) if len(sys.argv) > 1 else 4
tol = float(sys.argv[2]) if len(sys.argv) > 2 else 1e-12
ctx = hs.Context(0)
for name, mod, mesh in (("square_circle (config 1)", sc, sc.load_mesh(sc.default_mesh_path())),
                        ("flower_v2 (config 2)", flower, flower.load_mesh())):
    for p in (4, 6):
        eps, teps = [], []
        for lvl in range(levels):
            N = 17 * 2 ** lvl
            t0 = time.time()
            r = mod.solve_level(ctx, mesh, p, N, tol=tol, maxit=20000)
            eps.append(r["eps"]); teps.append(r["tau_eps"])
            print("%s p=%d N=%3d: eps %.4e tau_eps %.4e  CG iterations %d converged %d  wall %.1f s" %
                  (name, p, N, r["eps"], r["tau_eps"], r["stats"]["outer_iterations"], r["stats"]["converged"], time.time() - t0),
                  flush=True)
        e, t = np.array(eps), np.array(teps)
        print("   rates eps    :", np.round(np.log2(e[:-1] / e[1:]), 2))
        print("   rates tau_eps:", np.round(np.log2(t[:-1] / t[1:]), 2), flush=True)
