"""ORACLE -- test infrastructure only.  CPU restatement of the reference's SEAS BP1 right-hand side
(seas/BP1/odefun.jl:8-121) on top of oracle/hybrid.py: sparse locoperator, direct local solve (SuperLU in place
of CHOLMOD), computetraction_mod, rateandstate + newtbndv per fault node, sequentially and with the reference's
early returns on failure.

Parity status: the right-hand side is pinned by executing the reference's own statements -- BP1.jl's setup and odefun.jl
interpreted by tests/refexec/minijulia.py give the same d(psi, delta)/dt as OdeFun to 1e-13 at three states, and the same
rejection on a NaN state (tests/test_reference_executed.py; golden copy tests/golden/refexec/bp1_odefun_N40.npz).  The
integrator is NOT pinned that way: OrdinaryDiffEq is an unvendored, unpinned dependency, Tsit5 below is restated from the
published tableau (order conditions in tests/test_oracle_bp1.py).  Also pinned by construction: at t = 0 with psi = psi0,
delta = 0 the slip rate equals the initial rate 1e-9 that tau_z0 and theta are built from (BP1.jl:104-113).
"""
import numpy as np

from . import hybrid as orc


class OdeFun:
    def __init__(self, p, N, metrics, LFtoB, RSa, params):
        self.lop = orc.locoperator(p, N, N, metrics, LFtoB)                     # BP1.jl:75
        self.F = orc.default_factorization(self.lop.Mt)                          # BP1.jl:78-79
        self.LFtoB = LFtoB
        self.RSa = np.asarray(RSa, float)
        self.prm = dict(params)
        self.n = N + 1
        self.u = np.zeros((N + 1) ** 2)

    def __call__(self, t, y):
        """-> (dpsiV, rejected) exactly as odefun fills dψV (odefun.jl:28-108)."""
        P = self.prm
        n = self.n
        psi, delta = y[:n], y[n:].copy()
        bc_D = lambda lf, x, yy: (2 - lf) * (delta / 2) + (lf - 1) * np.full(x.shape, t * P["Vp"] / 2)   # :36
        bc_N = lambda lf, x, yy, nx, ny: np.zeros(x.shape)
        ge = np.zeros(n * n)
        orc.locbcarray_mod(ge, self.lop, self.LFtoB, bc_D, bc_N)                 # :42
        self.u = self.F.solve(ge)                                                # :43
        out = np.zeros(2 * n)
        dtau = -P["mu_shear"] * orc.computetraction_mod(self.lop, 1, self.u, delta)               # :59
        for k in range(n):                                                       # :69-108
            taun = dtau[k] + P["tau_z0"]
            if np.isnan(taun):
                return out, True
            VR = abs(taun / P["eta"]); VL = -VR
            f = lambda V: orc.rateandstate(V, psi[k], P["sigma_n"], taun, P["eta"], self.RSa[k], P["V0"])
            Vn, _, it = orc.newtbndv(f, VL, VR, 0.0, ftol=P["ftol"], atolx=P["atolx"], rtolx=P["rtolx"],
                                     maxiter=P["maxiter"])
            if np.isnan(Vn) or it < 0:
                return out, True
            out[n + k] = Vn
            d = (P["b"] * P["V0"] / P["Dc"]) * (np.exp((P["f0"] - psi[k]) / P["b"]) - abs(Vn) / P["V0"])   # :101
            if not np.isfinite(d):
                return out, True
            out[k] = d
        return out, False


# ---- the integrator of BP1.jl:159-161, restated for the oracle runs -------------------------------------------------
# Tsit5 (Tsitouras 2011, the tableau OrdinaryDiffEq uses) with the package's default PI step-size control; see the
# notes in hybridsbp_b200/bp1.py.  This copy is deliberately written on its own (explicit stage formulas, no shared
# tables) so that the GPU-side driver is checked against a second statement of the same method.
def tsit5(f, y0, t0, t1, dt0, abstol=1e-6, reltol=1e-3, tstops=(), max_steps=10 ** 9, stop_on_underflow=False):
    """f(t, y) -> (dy, rejected).  Returns (ts, ys, nrejected)."""
    c2, c3, c4, c5 = 0.161, 0.327, 0.9, 0.9800255409045097
    a21 = 0.161
    a31, a32 = -0.008480655492356989, 0.335480655492357
    a41, a42, a43 = 2.8971530571054935, -6.359448489975075, 4.3622954328695815
    a51, a52, a53, a54 = 5.325864828439257, -11.748883564062828, 7.4955393428898365, -0.09249506636175525
    a61, a62, a63, a64, a65 = 5.86145544294642, -12.92096931784711, 8.159367898576159, -0.071584973281401, -0.028269050394068383
    a71, a72, a73, a74, a75, a76 = (0.09646076681806523, 0.01, 0.4798896504144996, 1.379008574103742, -3.290069515436081,
                                    2.324710524099774)
    bt = (-0.00178001105222577714, -0.0008164344596567469, 0.007880878010261995, -0.1447110071732629, 0.5823571654525552,
          -0.45808210592918697, 0.015151515151515152)
    beta1, beta2, gamma, qmin, qmax, qoldinit = 7 / 50, 2 / 25, 9 / 10, 1 / 5, 10.0, 1e-4
    t, y = float(t0), np.array(y0, dtype=float)
    ts, ys, nrej, qold, dt = [t], [y.copy()], 0, qoldinit, float(dt0)
    k1, bad = f(t, y)
    if bad:
        raise RuntimeError("right-hand side rejected the initial state")
    stops = sorted(float(x) for x in tstops)
    steps = 0
    while t < t1 and steps < max_steps:
        ahead = [x for x in stops if x > t]
        tend = min([t1] + ahead[:1])
        clipped = dt >= tend - t
        h = tend - t if clipped else dt
        tn = tend if clipped else t + h
        out_of_domain = False
        k = [k1]
        for stage in range(2, 8):
            if stage == 2:
                yy, tt = y + h * (a21 * k[0]), t + c2 * h
            elif stage == 3:
                yy, tt = y + h * (a31 * k[0] + a32 * k[1]), t + c3 * h
            elif stage == 4:
                yy, tt = y + h * (a41 * k[0] + a42 * k[1] + a43 * k[2]), t + c4 * h
            elif stage == 5:
                yy, tt = y + h * (a51 * k[0] + a52 * k[1] + a53 * k[2] + a54 * k[3]), t + c5 * h
            elif stage == 6:
                yy, tt = y + h * (a61 * k[0] + a62 * k[1] + a63 * k[2] + a64 * k[3] + a65 * k[4]), tn
            else:
                yy, tt = y + h * (a71 * k[0] + a72 * k[1] + a73 * k[2] + a74 * k[3] + a75 * k[4] + a76 * k[5]), tn
            ks, bad = f(tt, yy)
            if bad:
                out_of_domain = True
                break
            k.append(ks)
        if out_of_domain:
            dt = h * qmin
            nrej += 1
        else:
            ynew = yy
            est = h * (bt[0] * k[0] + bt[1] * k[1] + bt[2] * k[2] + bt[3] * k[3] + bt[4] * k[4] + bt[5] * k[5] + bt[6] * k[6])
            EEst = float(np.max(np.abs(est) / (abstol + reltol * np.maximum(np.abs(y), np.abs(ynew)))))
            if EEst == 0.0:
                q11, q = 0.0, 1.0 / qmax
            else:
                q11 = EEst ** beta1
                q = max(1.0 / qmax, min(1.0 / qmin, (q11 / qold ** beta2) / gamma))
            if EEst <= 1.0:
                t, y, k1 = tn, ynew, k[6]
                ts.append(t); ys.append(y.copy())
                steps += 1
                qold = max(EEst, qoldinit)
                dt = max(h / q, dt) if clipped else h / q
                continue
            dt = h / min(1.0 / qmin, q11 / gamma)
            nrej += 1
        if dt <= 4.0 * np.spacing(max(1.0, abs(t))):
            if stop_on_underflow:
                break
            raise RuntimeError("step size underflow at t = %g" % t)
    return np.array(ts), np.array(ys), nrej
