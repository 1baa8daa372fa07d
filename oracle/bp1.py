"""ORACLE -- test infrastructure only.  CPU restatement of the reference's SEAS BP1 right-hand side
(seas/BP1/odefun.jl:8-121) on top of oracle/hybrid.py: sparse locoperator, direct local solve (SuperLU in place
of CHOLMOD), computetraction_mod, rateandstate + newtbndv per fault node, sequentially and with the reference's
early returns on failure.

Parity status: unpinned by golden data (the reference stores no BP1 output and its integrator dependency is
unpinned); pinned by construction identities (tests/test_oracle_bp1.py): at t = 0 with psi = psi0, delta = 0 the
slip rate equals the initial rate 1e-9 that tau_z0 and theta are built from (BP1.jl:104-113).
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
