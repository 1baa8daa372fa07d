"""ORACLE -- test infrastructure only.  CPU restatement of SEAS BP1 on the multiblock mesh seas/BP1/meshes/BP1_v1.inp in the
style of the reference's drivers (the reference ships the mesh without a driver): the reference's own functions do all
the work --

  read_inp_2d, connectivityarrays, transfinite_blend (corner form), create_metrics, locoperator   (global_curved.jl:19-506, 802-956)
  LocalGlobalOperators, assembleλmatrix, cholesky(B) -> here a sparse direct solve                (:706-797; square_circle.jl:297-314)
  locbcarray! with the jump branch, in_jump as square_circle.jl:335-350, LocalToGLobalRHS!          (:596-623, 730-740)
  lambda = BF \\ b, u = M \\ (g - Fbar' lambda)                                                    (square_circle.jl:376-388)
  computetraction on the minus side of every fault face                                           (:638-644, square_circle.jl:405-416)
  rateandstate + newtbndv per fault node, state evolution                                         (:1031-1075; odefun.jl:69-108)

Boundary data: side set 1 Dirichlet u = sign(x) Vp t / 2, side set 2 traction free, side set 7 slip from the state, side
set 8 slip Vp t.  Parity status: unpinned by golden data (no reference driver exists for this mesh); pinned by the
single-block benchmark it reduces to (tests/test_bp1_multiblock_gpu.py: symmetric problem, same friction law)."""
import numpy as np
import scipy.sparse.linalg as spla

from . import hybrid as orc


class MultiblockOdeFun:
    def __init__(self, filename, p, N, params, RSa, fault_faces, steady_faces, sign):
        verts, EToV, EToF, FToB, _ = orc.read_inp_2d(filename)
        FToE, FToLF, EToO, EToS = orc.connectivityarrays(EToV, EToF)
        ne = EToV.shape[1]
        lop = []
        for e in range(ne):
            vx, vy = verts[0, EToV[:, e] - 1], verts[1, EToV[:, e] - 1]
            xf = lambda r, s, v=vx: orc.transfinite_blend_corners(v[0], v[1], v[2], v[3], r, s)
            yf = lambda r, s, v=vy: orc.transfinite_blend_corners(v[0], v[1], v[2], v[3], r, s)
            lop.append(orc.locoperator(p, N, N, orc.create_metrics(p, N, N, xf, yf), FToB[EToF[:, e] - 1]))
        Ns = [N] * ne
        M, FbarT, D, vstarts, FTol = orc.LocalGlobalOperators(lop, Ns, Ns, FToB, FToE, FToLF, EToO, EToS)
        B = orc.assemblelambdamatrix(FTol, vstarts, EToF, FToB, M.F, D, FbarT)
        self.BF = spla.splu(B.tocsc())
        self.lop, self.M, self.FbarT, self.vstarts, self.FTol = lop, M, FbarT, np.asarray(vstarts), np.asarray(FTol)
        self.EToF, self.FToB, self.conn = EToF, FToB, (FToE, FToLF, EToO, EToS)
        self.prm, self.RSa = dict(params), np.asarray(RSa, float)
        self.fault_faces, self.steady_faces, self.sign = [int(f) for f in fault_faces], [int(f) for f in steady_faces], sign
        self.nl = N + 1
        self.n = self.nl * len(self.fault_faces)
        self.u = None
        self.lam = None

    def stress_change(self, delta, t):
        P, nl = self.prm, self.nl
        FToE, FToLF, EToO, EToS = self.conn
        EToF, FToB = self.EToF, self.FToB
        jump = {f: self.sign[f] * delta[i * nl:(i + 1) * nl] for i, f in enumerate(self.fault_faces)}
        for f in self.steady_faces:
            jump[f] = self.sign[f] * np.full(nl, P["Vp"] * t)
        g = np.zeros(self.vstarts[-1] - 1)
        gd = np.zeros(self.FTol[-1] - 1)
        bc_D = lambda lf, x, y: np.sign(x) * (P["Vp"] * t / 2)
        bc_N = lambda lf, x, y, nx, ny: np.zeros(x.shape)
        for e in range(len(self.lop)):
            def in_jump(lf, x, y, e=e):                       # square_circle.jl:335-350
                f = EToF[lf - 1, e] - 1
                d = jump[int(f)]
                if EToS[lf - 1, e] == 1:
                    assert EToO[lf - 1, e]
                    return -d
                return d if EToO[lf - 1, e] else d[::-1]
            views = []
            for lf in range(4):
                f = EToF[lf, e] - 1
                sl = gd[self.FTol[f] - 1:self.FTol[f + 1] - 1]
                views.append(sl if EToO[lf, e] else sl[::-1])
            orc.locbcarray(g[self.vstarts[e] - 1:self.vstarts[e + 1] - 1], views, self.lop[e], FToB[EToF[:, e] - 1], bc_D, bc_N, in_jump)
        bl = np.zeros(len(gd)); u = np.zeros(len(g))
        orc.LocalToGLobalRHS(bl, g, gd, u, self.M.F, self.FbarT, self.vstarts)
        lam = self.BF.solve(bl)
        rhs = g - self.FbarT.T @ lam
        for e in range(len(self.lop)):
            sl = slice(self.vstarts[e] - 1, self.vstarts[e + 1] - 1)
            u[sl] = self.M.F[e].solve(rhs[sl])
        self.u, self.lam = u, lam
        out = np.empty(self.n)
        for i, f in enumerate(self.fault_faces):
            e1, lf1 = FToE[0, f] - 1, FToLF[0, f]
            lo = self.lop[e1]
            T = orc.computetraction(lo, lf1, u[self.vstarts[e1] - 1:self.vstarts[e1 + 1] - 1], lam[self.FTol[f] - 1:self.FTol[f + 1] - 1],
                                    jump[f])
            out[i * nl:(i + 1) * nl] = P["mu_shear"] * lo.nx[lf1 - 1] * T
        return out

    def __call__(self, t, y):
        """-> (dpsiV, rejected): the node loop of odefun.jl:69-108 on the stress change of this state"""
        P, n = self.prm, self.n
        psi, delta = y[:n], y[n:]
        dtau = self.stress_change(delta, t)
        out = np.zeros(2 * n)
        for k in range(n):
            taun = dtau[k] + P["tau_z0"]
            if np.isnan(taun):
                return out, True
            VR = abs(taun / P["eta"]); VL = -VR
            f = lambda V: orc.rateandstate(V, psi[k], P["sigma_n"], taun, P["eta"], self.RSa[k], P["V0"])
            Vn, _, it = orc.newtbndv(f, VL, VR, 0.0, ftol=P["ftol"], atolx=P["atolx"], rtolx=P["rtolx"], maxiter=P["maxiter"])
            if np.isnan(Vn) or it < 0:
                return out, True
            out[n + k] = Vn
            d = (P["b"] * P["V0"] / P["Dc"]) * (np.exp((P["f0"] - psi[k]) / P["b"]) - abs(Vn) / P["V0"])
            if not np.isfinite(d):
                return out, True
            out[k] = d
        return out, False
