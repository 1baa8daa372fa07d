"""Host side of the SEAS BP1 antiplane earthquake-cycle benchmark over libhsbp (mirror of the reference's
seas/BP1/BP1.jl `main` and odefun.jl).

  setup(...)        BP1.jl:5-146   physical parameters, stretched single block, D/D/N/N operator,
                                   rate-and-state fields a(depth), tau_z0, theta, psi0
  Fault.rhs(t, y)   odefun.jl:8-121  on the GPU (hsbp_bp1_rhs): boundary scatter, local solve, traction,
                                   per-node bracketed Newton, state evolution
  integrate(...)    BP1.jl:148-161 adaptive explicit Runge-Kutta with infinity-norm error control and
                                   step rejection through `isoutofdomain`

Integrator note.  The reference calls OrdinaryDiffEq's Tsit5, a dependency that is neither vendored nor
pinned; its tableau is not part of the reference.  `integrate` is a Dormand-Prince 5(4)
pair with the same controls (dt0 = one year, infinity norm, rejection callback; abstol 1e-6 / reltol 1e-3,
which is what `solve` falls back to because BP1.jl:160 passes the unknown keywords atol / rtol).
Parity for this path is defined against the oracle's odefun driven by this same integrator.
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

    def __init__(self, ctx, su: Bp1Setup, local_tol=1e-13, local_maxit=200000, local_mode=LOCAL_BAND):
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


# ---- adaptive Runge-Kutta (Dormand-Prince 5(4)) with rejection callback ------------------------------
_C = np.array([0, 1 / 5, 3 / 10, 4 / 5, 8 / 9, 1, 1])
_A = [[],
      [1 / 5],
      [3 / 40, 9 / 40],
      [44 / 45, -56 / 15, 32 / 9],
      [19372 / 6561, -25360 / 2187, 64448 / 6561, -212 / 729],
      [9017 / 3168, -355 / 33, 46732 / 5247, 49 / 176, -5103 / 18656],
      [35 / 384, 0, 500 / 1113, 125 / 192, -2187 / 6784, 11 / 84]]
_B5 = np.array([35 / 384, 0, 500 / 1113, 125 / 192, -2187 / 6784, 11 / 84, 0])
_B4 = np.array([5179 / 57600, 0, 7571 / 16695, 393 / 640, -92097 / 339200, 187 / 2100, 1 / 40])


def integrate(rhs, y0, t0, t1, dt0, abstol=1e-6, reltol=1e-3, max_steps=10 ** 9, qmin=0.2, qmax=10.0, safety=0.9,
              stop_on_underflow=False):
    """Integrate y' = rhs(t, y) -> (dy, rejected).  A step is rejected when the error test fails or when any
    stage reports `rejected` (the reference's isoutofdomain / reject_step mechanism, BP1.jl:149-159).
    Returns (ts, ys, nrejected)."""
    t, y = float(t0), np.array(y0, dtype=float)
    ts, ys = [t], [y.copy()]
    dt = float(dt0)
    nrej = 0
    k1, bad = rhs(t, y)
    if bad:
        raise RuntimeError("right-hand side rejected the initial state")
    steps = 0
    while t < t1 and steps < max_steps:
        dt = min(dt, t1 - t)
        K = [k1]
        ok = True
        for s in range(1, 7):
            ys_ = y + dt * sum(a * k for a, k in zip(_A[s], K))
            ks, bad = rhs(t + _C[s] * dt, ys_)
            if bad:
                ok = False
                break
            K.append(ks)
        if ok:
            y5 = y + dt * sum(b * k for b, k in zip(_B5, K))
            y4 = y + dt * sum(b * k for b, k in zip(_B4, K))
            err = np.max(np.abs(y5 - y4) / (abstol + reltol * np.maximum(np.abs(y), np.abs(y5))))   # infinity norm
            if err <= 1.0:
                t += dt
                y = y5
                k1 = K[6]                      # first-same-as-last
                ts.append(t); ys.append(y.copy())
                steps += 1
                dt *= min(qmax, max(qmin, safety * err ** -0.2)) if err > 0 else qmax
                continue
            dt *= max(qmin, safety * err ** -0.2)
        else:
            dt *= 0.5
        nrej += 1
        if dt <= 4.0 * np.spacing(max(1.0, abs(t))):     # dtmin of the reference's integrator: the resolution of t (coseismic
                                                          # steps are milliseconds at t ~ 1e10 s)
            if stop_on_underflow:            # return the series up to here (the caller reports the stop time)
                break
            raise RuntimeError("step size underflow at t = %g" % t)
    return np.array(ts), np.array(ys), nrej
