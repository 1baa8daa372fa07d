"""Host side of the SEAS BP1 antiplane earthquake-cycle benchmark over libhsbp (mirror of the reference's
seas/BP1/BP1.jl `main` and odefun.jl).

  setup(...)        BP1.jl:5-146   physical parameters, stretched single block, D/D/N/N operator,
                                   rate-and-state fields a(depth), tau_z0, theta, psi0
  Fault.rhs(t, y)   odefun.jl:8-121  on the GPU (hsbp_bp1_rhs): boundary scatter, local solve, traction,
                                   per-node bracketed Newton, state evolution
  integrate(...)    BP1.jl:148-161 adaptive explicit Runge-Kutta with infinity-norm error control and
                                   step rejection through `isoutofdomain`

Integrator note.  The reference calls OrdinaryDiffEq's Tsit5, a dependency that is neither vendored nor
pinned.  `integrate` restates it from the published tableau (Tsitouras 2011) and the
package's documented step control (PI controller, infinity norm, dt0 = one year, rejection through
isoutofdomain; abstol 1e-6 / reltol 1e-3, which is what `solve` falls back to because BP1.jl:160 passes the
unknown keywords atol / rtol).  Parity for this path is defined against the oracle's odefun driven by the
oracle's own copy of the same integrator (oracle/bp1.py).
"""
import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import host
from ._lib import Bp1Params, Bp1Stats, _f64, lib
from .blocks import Blocks, LOCAL_PCG, LOCAL_BAND

YEAR_SECONDS = 31556926          # odefun.jl:1


@dataclass
class Bp1Setup:
    p: int
    N: int
    metrics: object
    LFtoB: tuple
    RSa: np.ndarray
    params: dict
    psi_delta0: np.ndarray
    yf: np.ndarray


def setup(N=200, SBPp=2, Lx=80.0, Ly=80.0):
    """BP1.jl:8-146 (host arithmetic only)."""
    Vp = 1e-9; rho = 2.670; cs = 3.464; sigma_n = 50.0
    RSamin, RSamax, RSb, RSDc, RSf0, RSV0, RSVinit, RSH1, RSH2 = 0.01, 0.025, 0.015, 0.016, 0.6, 1e-6, 1e-9, 15.0, 18.0
    mu = cs ** 2 * rho
    eta = mu / (2 * cs)
    el_x = el_y = 10e12                      # BP1.jl:63-64: effectively uniform spacing

    def xt(r, s):
        q = np.arctan(Lx / el_x)
        return el_x * np.tan(q * (0.5 * r + 0.5)), el_x / np.cos(q * (0.5 * r + 0.5)) ** 2 * q * 0.5, np.zeros_like(s)

    def yt(r, s):
        q = np.arctan(Ly / el_y)
        return el_y * np.tan(q * (0.5 * s + 0.5)), np.zeros_like(r), el_y / np.cos(q * (0.5 * s + 0.5)) ** 2 * q * 0.5

    metrics = host.create_metrics(SBPp, N, N, xt, yt)
    LFtoB = (host.BC_DIRICHLET, host.BC_DIRICHLET, host.BC_NEUMANN, host.BC_NEUMANN)      # BP1.jl:73
    yf = metrics.facecoord[1][0]
    RSa = RSamin - (RSamin - RSamax) * np.minimum(1.0, np.maximum(0.0, (RSH1 - yf) / (RSH1 - RSH2)))   # BP1.jl:99-102
    tau_z0 = sigma_n * RSamax * np.arcsinh(RSVinit / (2 * RSV0) * np.exp((RSf0 + RSb * np.log(RSV0 / RSVinit)) / RSamax)) \
        + eta * RSVinit                                                                    # BP1.jl:104-106
    theta = (RSDc / RSV0) * np.exp((RSa / RSb) * np.log((2 * RSV0 / RSVinit) *
                                                        np.sinh((tau_z0 - eta * RSVinit) / (RSa * sigma_n))) - RSf0 / RSb)
    psi0 = RSf0 + RSb * np.log(RSV0 * theta / RSDc)                                        # BP1.jl:113
    y0 = np.zeros(2 * (N + 1))
    y0[:N + 1] = psi0
    params = dict(Vp=Vp, mu_shear=mu, sigma_n=sigma_n, eta=eta, V0=RSV0, tau_z0=float(tau_z0), Dc=RSDc, f0=RSf0, b=RSb,
                  ftol=1e-9, atolx=1e-9, rtolx=1e-9, maxiter=500)
    return Bp1Setup(SBPp, N, metrics, LFtoB, RSa, params, y0, yf)


class Fault:
    """Device-resident BP1 right-hand side (hsbp_bp1_*)."""

    def __init__(self, ctx, su: Bp1Setup, local_tol=1e-13, local_maxit=200000, local_mode=LOCAL_BAND, condense=True):
        self.su = su
        m = su.metrics
        self.blk = Blocks(ctx, su.p, [su.N], [su.N])
        fl = lambda a: np.asarray(a).reshape(-1, order="F")
        self.blk.set_metrics(fl(m.crr), fl(m.css), fl(m.crs))
        self.blk.set_bc(np.asarray(su.LFtoB, dtype=np.int64))
        self.blk.compute_tau(2.0)
        # the reference keeps cholesky(M-tilde) for the whole run (BP1.jl:78) and back-solves in every odefun call
        # (odefun.jl:43): the banded factorisation is its direct counterpart; LOCAL_PCG is the matrix-free variant
        self.blk.local_setup(local_mode, tol=local_tol, maxit=local_maxit)
        prm = Bp1Params(**su.params)
        a, pa = _f64(su.RSa)
        sj, psj = _f64(m.sJ[0])
        h = C.c_void_p()
        ctx._check(lib().hsbp_bp1_create(self.blk.h, 1, 1, 2, pa, psj, C.byref(prm), C.byref(h)))
        self.h = h
        self.ctx = ctx
        self.blk._children.add(self)
        self.n = su.N + 1
        self.last_stats = None
        # odefun needs u = M-tilde^-1 ge only through the traction on the fault: condense that map once (N + 2 local
        # solves), every later right-hand side is one small kernel; displacement() solves on demand
        if condense:
            ctx._check(lib().hsbp_bp1_condense(self.h, 1))
        self.condensed = bool(condense)

    def rhs(self, t, y):
        """(dy, rejected) = odefun(y, t)."""
        y, py = _f64(y)
        out = np.empty(2 * self.n)
        st = Bp1Stats()
        self.ctx._check(lib().hsbp_bp1_rhs(self.h, float(t), py, C.c_void_p(out.ctypes.data), C.byref(st)))
        self.last_stats = st.as_dict()
        return out, bool(st.rejected)

    def displacement(self):
        u = np.empty(self.blk.VNp)
        self.ctx._check(lib().hsbp_bp1_get_u(self.h, C.c_void_p(u.ctypes.data)))
        return u

    def close(self):
        if self.h is not None:
            h, self.h = self.h, None
            if self.blk.h is not None and self.ctx.h is not None:      # the C object points into its blocks / context
                lib().hsbp_bp1_destroy(h)
            self.blk.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


# ---- Tsit5: the integrator BP1.jl:159-161 asks OrdinaryDiffEq for --------------------------------------------------
# Tableau of Ch. Tsitouras, "Runge-Kutta pairs of order 5(4) satisfying only the first column simplifying assumption",
# Computers & Mathematics with Applications 62 (2011) 770-775 (the pair OrdinaryDiffEq implements as Tsit5; the package
# itself is not vendored by the reference).  7 stages, first-same-as-last; BT = b - bhat weighs the
# stages in the error estimate.  tests/test_oracle_bp1.py checks the order conditions and the observed order.
TSIT5_C = np.array([0.0, 0.161, 0.327, 0.9, 0.9800255409045097, 1.0, 1.0])
TSIT5_A = [[],
           [0.161],
           [-0.008480655492356989, 0.335480655492357],
           [2.8971530571054935, -6.359448489975075, 4.3622954328695815],
           [5.325864828439257, -11.748883564062828, 7.4955393428898365, -0.09249506636175525],
           [5.86145544294642, -12.92096931784711, 8.159367898576159, -0.071584973281401, -0.028269050394068383],
           [0.09646076681806523, 0.01, 0.4798896504144996, 1.379008574103742, -3.290069515436081, 2.324710524099774]]
TSIT5_BT = np.array([-0.00178001105222577714, -0.0008164344596567469, 0.007880878010261995, -0.1447110071732629,
                     0.5823571654525552, -0.45808210592918697, 0.015151515151515152])


def integrate(rhs, y0, t0, t1, dt0, abstol=1e-6, reltol=1e-3, max_steps=10 ** 9, qmin=0.2, qmax=10.0, gamma=0.9,
              beta1=7.0 / 50.0, beta2=2.0 / 25.0, qoldinit=1e-4, stop_on_underflow=False, tstops=None):
    """Integrate y' = rhs(t, y) -> (dy, rejected) with Tsit5 and the step control `solve(prob, Tsit5(); dt, isoutofdomain,
    internalnorm = (x, _) -> norm(x, Inf))` runs with (BP1.jl:159-161):
      * error estimate dt * sum(BT_i k_i) scaled by abstol + reltol * max(|y_old|, |y_new|), infinity norm;
        abstol = 1e-6, reltol = 1e-3 are the package defaults (BP1.jl:160 passes the unknown keywords atol / rtol);
      * PI step-size controller with the defaults of a 5th-order pair (beta1 = 7/50, beta2 = 2/25, gamma = 9/10, qmin = 1/5,
        qmax = 10, steady-state band [1, 1], qoldinit = 1e-4);
      * a step whose stages set `rejected` (the reference's reject_step flag, read by `stepcheck` = isoutofdomain,
        BP1.jl:149-159) is redone with dt * qmin;
      * tstops: times the integration must hit exactly (output times of the parity tests).
    Returns (ts, ys, nrejected)."""
    t, y = float(t0), np.array(y0, dtype=float)
    ts, ys = [t], [y.copy()]
    dt = float(dt0)
    nrej = 0
    qold = qoldinit
    k1, bad = rhs(t, y)
    if bad:
        raise RuntimeError("right-hand side rejected the initial state")
    stops = sorted(float(x) for x in tstops) if tstops is not None else []
    si = 0
    steps = 0
    while t < t1 and steps < max_steps:
        while si < len(stops) and stops[si] <= t:
            si += 1
        tend = min(t1, stops[si]) if si < len(stops) else t1
        clipped = dt >= tend - t
        h = tend - t if clipped else dt
        K = [k1]
        ok = True
        for s in range(1, 7):
            ys_ = y + h * sum(a * k for a, k in zip(TSIT5_A[s], K))
            ks, bad = rhs(tend if (clipped and s >= 5) else t + TSIT5_C[s] * h, ys_)
            if bad:
                ok = False
                break
            K.append(ks)
        if not ok:                                         # isoutofdomain: dt * qmin
            dt = h * qmin
            nrej += 1
        else:
            ynew = ys_                                     # the last stage point is the new solution (a_7j = b_j)
            est = h * sum(b * k for b, k in zip(TSIT5_BT, K))
            EEst = np.max(np.abs(est) / (abstol + reltol * np.maximum(np.abs(y), np.abs(ynew))))   # internalnorm = Inf norm
            if EEst == 0.0:
                q11, q = 0.0, 1.0 / qmax
            else:
                q11 = EEst ** beta1
                q = max(1.0 / qmax, min(1.0 / qmin, (q11 / qold ** beta2) / gamma))
            if EEst <= 1.0:
                t = tend if clipped else t + h
                y = ynew
                k1 = K[6]                                  # first-same-as-last
                ts.append(t); ys.append(y.copy())
                steps += 1
                qold = max(EEst, qoldinit)
                dtn = h / q
                dt = max(dtn, dt) if clipped else dtn      # a step shortened to hit a stop does not shrink the next one
                continue
            dt = h / min(1.0 / qmin, q11 / gamma)
            nrej += 1
        if dt <= 4.0 * np.spacing(max(1.0, abs(t))):     # dtmin of the reference's integrator: the resolution of t (coseismic
                                                          # steps are milliseconds at t ~ 1e10 s)
            if stop_on_underflow:            # return the series up to here (the caller reports the stop time)
                break
            raise RuntimeError("step size underflow at t = %g" % t)
    return np.array(ts), np.array(ys), nrej
