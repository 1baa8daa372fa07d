#!/usr/bin/env python
"""Run the reference's own Julia sources (read from /root/reference) through tests/refexec/minijulia.py -- Julia itself is not
installed in the build image.  TEST INFRASTRUCTURE, CPU only.

    python tools/run_reference.py square_circle [--p 4] [--levels 2] [--n0 17]     # square_circle.jl: errors and rates per level
    python tools/run_reference.py bp1 [--n 40]                                     # BP1.jl setup + one odefun call at t = 0
    python tools/run_reference.py script check_residual.jl                         # a script as it is; prints what its @show lines show
    python tools/run_reference.py flower [--p 4]                                   # the reference's functions on meshes/flower_v2.inp
"""
import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("what", choices=["square_circle", "bp1", "script", "flower"])
    ap.add_argument("name", nargs="?", default=None)
    ap.add_argument("--p", type=int, default=4)
    ap.add_argument("--levels", type=int, default=2)
    ap.add_argument("--n0", type=int, default=17)
    ap.add_argument("--n", type=int, default=40)
    a = ap.parse_args()
    from refexec import drivers
    from refexec.minijulia import Interp
    if not os.path.isdir(drivers.REF):
        sys.exit("the reference tree is not mounted at %s" % drivers.REF)
    t0 = time.time()
    if a.what == "square_circle":
        cap, _, _ = drivers.run_square_circle(p=a.p, levels=a.levels, N0=a.n0, keep=("ϵ", "τϵ", "lvl", "λ", "u"))
        eps, teps = np.array(cap[-1]["ϵ"]), np.array(cap[-1]["τϵ"])
        for k in range(a.levels):
            print("level %d  N = %3d  eps = %.9e  tau_eps = %.9e  (%d lambda points, %d volume points)"
                  % (k + 1, a.n0 * 2 ** k, eps[k], teps[k], cap[k]["λ"].size, cap[k]["u"].size))
        if a.levels > 1:
            print("rates", (np.log(eps[:-1]) - np.log(eps[1:])) / np.log(2), (np.log(teps[:-1]) - np.log(teps[1:])) / np.log(2))
    elif a.what == "bp1":
        it, sol, yf = drivers.run_bp1_setup(a.n)
        prob = sol.get("prob")
        y0 = np.array(prob.get("u0")); d = np.zeros_like(y0)
        prob.get("f")(d, y0.copy(), prob.get("p"), 0.0)
        print("N = %d: psi0 in [%.6f, %.6f], V(t = 0) in [%.3e, %.3e], solver options %s"
              % (a.n, y0[:a.n + 1].min(), y0[:a.n + 1].max(), d[a.n + 1:].min(), d[a.n + 1:].max(), sorted(sol.get("options"))))
    elif a.what == "flower":
        c = drivers.run_flower(a.p, a.n0)
        print("flower_v2: %d blocks, %d reversed faces, |lambda| = %.12e, |u| = %.12e" % (np.asarray(c["EToV"]).shape[1],
              int((~np.asarray(c["EToO"]).astype(bool)).sum()), np.linalg.norm(c["λ"]), np.linalg.norm(c["u"])))
    else:
        if a.name is None: sys.exit("script name missing")
        path = os.path.join(drivers.REF, a.name)
        it = Interp(os.path.dirname(path))
        it.include(os.path.basename(path))
        for v in it.log: print(v)
    print("[%.1f s]" % (time.time() - t0))


if __name__ == "__main__":
    main()
