"""SEAS BP1 (seas/BP1/BP1.jl) time integration at the reference's resolution: `gpu` runs the device-resident ODE stage
(hsbp_bp1_rhs, banded Cholesky local solver), `oracle` the CPU restatement of odefun.jl (test infrastructure; run by hand
to produce the series the GPU run is compared with).  Controls as in the reference: Tsit5, dt0 = 1 year, infinity norm, steps rejected
when the root-find fails (BP1.jl:149-161); abstol 1e-6 / reltol 1e-3 are the package defaults BP1.jl:160 falls back to (its
atol / rtol keywords are not the integrator's names).
usage: python tools/bp1_run.py gpu|oracle N years out.npz [max_steps]"""
import sys, time
import numpy as np
sys.path.insert(0, ".")
from hybridsbp_b200 import bp1

which, N, years, out = sys.argv[1], int(sys.argv[2]), float(sys.argv[3]), sys.argv[4]
max_steps = int(sys.argv[5]) if len(sys.argv) > 5 else 10 ** 9
su = bp1.setup(N=N)
n = su.N + 1
if which == "gpu":
    import hybridsbp_b200 as hs
    ctx = hs.Context(0)
    f = bp1.Fault(ctx, su, local_tol=1e-12)
    rhs = f.rhs
else:
    from oracle.bp1 import OdeFun
    rhs = OdeFun(su.p, su.N, su.metrics, su.LFtoB, su.RSa, su.params)
t0 = time.time()
ts, ys, nrej = bp1.integrate(rhs, su.psi_delta0, 0.0, years * bp1.YEAR_SECONDS, bp1.YEAR_SECONDS, abstol=1e-6, reltol=1e-3,
                             stop_on_underflow=True, max_steps=max_steps)
wall = time.time() - t0
ts, ys = np.asarray(ts), np.asarray(ys)
V = np.array([rhs(t, y)[0][n:] for t, y in zip(ts, ys)])
np.savez_compressed(out, t=ts, y=ys, V=V, wall=wall, nrej=nrej)
Vmax = V.max(axis=1)
events = int(np.sum((Vmax[1:] > 1e-2) & (Vmax[:-1] <= 1e-2)))
if ts[-1] < years * bp1.YEAR_SECONDS * (1 - 1e-12):
    print("step size underflow at t = %.6e s (%.2f years): every trial step is rejected by the root-find / state update"
          % (ts[-1], ts[-1] / bp1.YEAR_SECONDS))
print("%s: N=%d, %.0f years: %d steps, %d rejected, %.1f s wall (%.1f ms per accepted step); %d events (max V > 1e-2 m/s), "
      "max slip %.3f m, max V %.3e m/s" % (which, N, years, len(ts) - 1, nrej, wall, 1e3 * wall / max(1, len(ts) - 1), events,
                                         ys[-1][n:].max(), Vmax.max()), flush=True)
