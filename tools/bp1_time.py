"""Time the BP1 ODE right-hand side at the reference's size (N = 200, p = 2) with both local solvers and integrate
a few years with the banded Cholesky (the reference's `cholesky(M-tilde)` kept for the whole run, BP1.jl:78)."""
import sys, time
import numpy as np
sys.path.insert(0, ".")
import hybridsbp_b200 as hs
from hybridsbp_b200 import bp1
ctx = hs.Context(0)
su = bp1.setup(N=int(sys.argv[2]) if len(sys.argv) > 2 else 200)
for name, mode in (("banded Cholesky", hs.LOCAL_BAND), ("Jacobi-PCG", hs.LOCAL_PCG)):
    t0 = time.time()
    f = bp1.Fault(ctx, su, local_tol=1e-12, local_mode=mode)
    tsetup = time.time() - t0
    y = su.psi_delta0.copy()
    f.rhs(0.0, y)
    t0 = time.time(); nrep = 10
    for _ in range(nrep):
        d, rej = f.rhs(3.0e7, y)
    dt = (time.time() - t0) / nrep
    print("BP1 rhs at N=%d, %s: setup %.2f s, %.2f ms per call, local iterations %d" %
          (su.N, name, tsetup, dt * 1e3, f.last_stats["local_iterations"]), flush=True)
    if mode == hs.LOCAL_BAND:
        years = float(sys.argv[1]) if len(sys.argv) > 1 else 5.0
        t0 = time.time()
        ts, ys, nrej = bp1.integrate(f.rhs, su.psi_delta0, 0.0, years * bp1.YEAR_SECONDS, bp1.YEAR_SECONDS)
        print("integrated %.1f years: %d steps, %d rejected, %.1f s wall; max slip %.4e, max V %.3e" %
              (years, len(ts) - 1, nrej, time.time() - t0, ys[-1][su.N + 1:].max(),
               f.rhs(ts[-1], ys[-1])[0][su.N + 1:].max()), flush=True)
    f.close()
