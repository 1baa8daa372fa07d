"""ORACLE -- test infrastructure only (never imported by the product path).

CPU restatement (numpy / scipy.sparse) of the reference's block-local operator
construction and hybridized (trace) global assembly, function by function:

  transfinite_blend (+ _sbp, _corners: the three methods)   global_curved.jl:19-78
  connectivityarrays      global_curved.jl:82-132
  create_metrics          global_curved.jl:136-209
  locoperator             global_curved.jl:211-506
  glolambdaoperator       global_curved.jl:510-565   (reference name: gloλoperator)
  locbcarray_mod / locbcarray / computetraction(_mod) / locsourcearray
                          global_curved.jl:569-654
  SBPLocalOperator1 / LocalGlobalOperators / bcstarts / LocalToGLobalRHS /
  assemblelambdamatrix    global_curved.jl:659-797
  read_inp_2d             global_curved.jl:802-956
  rateandstate / newtbndv global_curved.jl:1031-1075

The sparse matrices (M-tilde, F_k, Fbar^T, B) are *assembled* exactly as the
reference does; the product never assembles them (it is matrix-free), which is
what makes this a meaningful checker.

Third-party arithmetic the reference delegates to (SuiteSparse CHOLMOD via the
`factorization` callback, global_curved.jl:698) is replaced by scipy's SuperLU
(`splu`) on the same SPD matrices; any correct direct solver agrees to
O(cond * eps).

Parity status: the reference ships no golden vectors and Julia is not installed
here, so the reference's own source text is EXECUTED instead: an interpreter
for the Julia subset these files use (tests/refexec/minijulia.py) runs
create_metrics, locoperator, read_inp_2d, connectivityarrays, the trace
operators, assembleλmatrix and the whole square_circle.jl driver statement by
statement, and every function of this module has to reproduce those outputs
(tests/test_reference_executed.py: operators to 1e-14, lambda and u of the
driver to 1e-11).  The outputs are committed as golden vectors
(tests/golden/refexec/, generator tools/gen_refexec_golden.py) for the tests
that run without the reference tree.  Also pinned by the reference's
identities, see tests/test_oracle_*.py.  What stays unpinned: CHOLMOD's
arithmetic (replaced as described above) and anything the interpreter's own
numpy / scipy runtime would get wrong in the same way as this module.

Conventions kept from the reference: entity ids stored in arrays (EToV, EToF,
FToE, FToLF, EToS) are 1-based; offsets vstarts / FTolambdastarts are 1-based
starts; fields are (Nr+1) x (Ns+1) arrays with r as the FIRST index, flattened
r-fastest (order="F").
"""
import re
from types import SimpleNamespace

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

from .sbp import diagonal_sbp_D1, variable_diagonal_sbp_D2

BC_DIRICHLET = 1
BC_NEUMANN = 2
BC_LOCKED_INTERFACE = 0
BC_JUMP_INTERFACE = 7


# --------------------------------------------------------------------------
# transfinite blend (global_curved.jl:19-78)
# --------------------------------------------------------------------------
def transfinite_blend(a1, a2, a3, a4, a1s, a2s, a3r, a4r, r, s):
    """(x, xr, xs) for edge curves a1..a4 and their derivatives, global_curved.jl:19-51."""
    c = np.array([a1(-1.0), a2(-1.0), a1(1.0), a2(1.0)], dtype=float)
    d = np.array([a3(-1.0), a3(1.0), a4(-1.0), a4(1.0)], dtype=float)
    assert np.allclose(c, d), "edge curves do not meet at the corners"   # :25
    x = ((1 + r) * a2(s) / 2 + (1 - r) * a1(s) / 2 +
         (1 + s) * a4(r) / 2 + (1 - s) * a3(r) / 2 -
         ((1 + r) * (1 + s) * a2(1.0) + (1 - r) * (1 + s) * a1(1.0) +
          (1 + r) * (1 - s) * a2(-1.0) + (1 - r) * (1 - s) * a1(-1.0)) / 4)
    xr = (a2(s) / 2 - a1(s) / 2 + (1 + s) * a4r(r) / 2 + (1 - s) * a3r(r) / 2 -
          (+(1 + s) * a2(1.0) - (1 + s) * a1(1.0) + (1 - s) * a2(-1.0) - (1 - s) * a1(-1.0)) / 4)
    xs = ((1 + r) * a2s(s) / 2 + (1 - r) * a1s(s) / 2 + a4(r) / 2 - a3(r) / 2 -
          (+(1 + r) * a2(1.0) + (1 - r) * a1(1.0) - (1 + r) * a2(-1.0) - (1 - r) * a1(-1.0)) / 4)
    return x, xr, xs


def transfinite_blend_sbp(a1, a2, a3, a4, r, s, p):
    """Edge derivatives by the order-p SBP first derivative, global_curved.jl:53-64 (as written: the closures for the r-derivatives
    ignore their argument and differentiate a3(r), a4(r))."""
    Nrp, Nsp = r.shape
    Dr = diagonal_sbp_D1(p, Nrp - 1)[0]
    Ds = diagonal_sbp_D1(p, Nsp - 1)[0]
    a2s = lambda t: (Ds @ a2(t).T).T                 # a2(s) * Ds'
    a1s = lambda t: (Ds @ a1(t).T).T
    a4r = lambda t: Dr @ a4(r)
    a3r = lambda t: Dr @ a3(r)
    return transfinite_blend(a1, a2, a3, a4, a1s, a2s, a3r, a4r, r, s)


def transfinite_blend_corners(v1, v2, v3, v4, r, s):
    """Straight-sided block from its four corner values, global_curved.jl:66-78."""
    e1 = lambda a: v1 * (1 - a) / 2 + v3 * (1 + a) / 2
    e2 = lambda a: v2 * (1 - a) / 2 + v4 * (1 + a) / 2
    e3 = lambda a: v1 * (1 - a) / 2 + v2 * (1 + a) / 2
    e4 = lambda a: v3 * (1 - a) / 2 + v4 * (1 + a) / 2
    e1a = lambda a: -v1 / 2 + v3 / 2
    e2a = lambda a: -v2 / 2 + v4 / 2
    e3a = lambda a: -v1 / 2 + v2 / 2
    e4a = lambda a: -v3 / 2 + v4 / 2
    return transfinite_blend(e1, e2, e3, e4, e1a, e2a, e3a, e4a, r, s)


# --------------------------------------------------------------------------
# connectivity (global_curved.jl:82-132)
# --------------------------------------------------------------------------
_LFTOLV = ((0, 2), (1, 3), (0, 1), (2, 3))          # :99, 0-based local vertices


def connectivityarrays(EToV, EToF):
    EToV = np.asarray(EToV)
    EToF = np.asarray(EToF)
    nelems = EToV.shape[1]
    nfaces = int(EToF.max())
    FToE = np.zeros((2, nfaces), dtype=np.int64)
    FToLF = np.zeros((2, nfaces), dtype=np.int64)
    EToO = np.zeros((4, nelems), dtype=bool)
    EToS = np.zeros((4, nelems), dtype=np.int64)
    for e in range(nelems):
        for lf in range(4):
            gf = EToF[lf, e] - 1
            if FToE[0, gf] == 0:
                assert FToLF[0, gf] == 0
                FToE[0, gf] = e + 1
                FToLF[0, gf] = lf + 1
                EToO[lf, e] = True
                EToS[lf, e] = 1
            else:
                assert FToE[1, gf] == 0 and FToLF[1, gf] == 0
                FToE[1, gf] = e + 1
                FToLF[1, gf] = lf + 1
                EToS[lf, e] = 2
                ne = FToE[0, gf] - 1
                nf = FToLF[0, gf] - 1
                nv = [EToV[v, ne] for v in _LFTOLV[nf]]
                lv = [EToV[v, e] for v in _LFTOLV[lf]]
                if nv == lv:
                    EToO[lf, e] = True
                elif nv[::-1] == lv:
                    EToO[lf, e] = False
                else:
                    raise RuntimeError("problem with connectivity")
    return FToE, FToLF, EToO, EToS


# --------------------------------------------------------------------------
# metrics (global_curved.jl:136-209)
# --------------------------------------------------------------------------
def create_metrics(pm, Nr, Ns, xf=None, yf=None):
    if xf is None:
        xf = lambda r, s: (r, np.ones_like(r), np.zeros_like(r))
    if yf is None:
        yf = lambda r, s: (s, np.zeros_like(s), np.ones_like(s))
    Nrp, Nsp = Nr + 1, Ns + 1
    assert pm <= 8
    r1 = np.linspace(-1, 1, Nrp)
    s1 = np.linspace(-1, 1, Nsp)
    r = np.repeat(r1[:, None], Nsp, axis=1)         # r[i, j] = r_i   (:151)
    s = np.repeat(s1[None, :], Nrp, axis=0)         # s[i, j] = s_j   (:152)
    x, xr, xs = xf(r, s)
    y, yr, ys = yf(r, s)
    x, xr, xs, y, yr, ys = [np.broadcast_to(np.asarray(a, float), (Nrp, Nsp)).copy()
                            for a in (x, xr, xs, y, yr, ys)]
    J = xr * ys - xs * yr
    assert J.min() > 0                               # :157
    rx = ys / J
    sx = -yr / J
    ry = -xs / J
    sy = xr / J
    crr = J * (rx * rx + ry * ry)
    crs = J * (sx * rx + sy * ry)
    css = J * (sx * sx + sy * sy)

    def unit(nx, ny):
        sJ = np.hypot(nx, ny)
        return nx / sJ, ny / sJ, sJ
    nx1, ny1, sJ1 = unit(-ys[0, :], xs[0, :])       # :170-175
    nx2, ny2, sJ2 = unit(ys[-1, :], -xs[-1, :])     # :177-182
    nx3, ny3, sJ3 = unit(yr[:, 0], -xr[:, 0])       # :184-189
    nx4, ny4, sJ4 = unit(-yr[:, -1], xr[:, -1])     # :191-196
    return SimpleNamespace(
        coord=(x, y),
        facecoord=((x[0, :].copy(), x[-1, :].copy(), x[:, 0].copy(), x[:, -1].copy()),
                   (y[0, :].copy(), y[-1, :].copy(), y[:, 0].copy(), y[:, -1].copy())),
        crr=crr, css=css, crs=crs, J=J,
        sJ=(sJ1, sJ2, sJ3, sJ4), nx=(nx1, nx2, nx3, nx4), ny=(ny1, ny2, ny3, ny4),
        rx=rx, ry=ry, sx=sx, sy=sy)


# --------------------------------------------------------------------------
# locoperator (global_curved.jl:211-506)
# --------------------------------------------------------------------------
PENALTY = {2: (2, 0.363636363, 1 / 2),              # (l, beta, alpha)  :402-413
           4: (4, 0.2505765857, 17 / 48),
           6: (7, 0.1878687080, 13649 / 43200)}


def _unit(n, k, shape_col=True):
    e = sp.csc_matrix(([1.0], ([k], [0])), shape=(n, 1))
    return e


def locoperator(p, Nr, Ns, metrics=None, LFToB=(BC_DIRICHLET,) * 4, tauscale=2.0,
                crr=None, css=None, crs=None):
    """Assemble M-tilde, F_k, ... exactly as global_curved.jl:211-506 does."""
    if metrics is None:
        metrics = create_metrics(p, Nr, Ns)
    crr = metrics.crr if crr is None else crr
    css = metrics.css if css is None else css
    crs = metrics.crs if crs is None else crs
    csr = crs
    J = metrics.J
    Nrp, Nsp = Nr + 1, Ns + 1
    Np = Nrp * Nsp
    if p not in PENALTY:
        raise ValueError("unknown order")

    Dr, HrI, Hr, _ = diagonal_sbp_D1(p, Nr)         # :226
    Qr = (Hr @ Dr).tocsc()
    QrT = Qr.T.tocsc()
    Ds, HsI, Hs, _ = diagonal_sbp_D1(p, Ns)         # :230
    Qs = (Hs @ Ds).tocsc()
    QsT = Qs.T.tocsc()
    Ir = sp.identity(Nrp, format="csc")
    Is = sp.identity(Nsp, format="csc")
    hs_d = Hs.diagonal()
    hr_d = Hr.diagonal()

    # rr part: one 1-D variable-coefficient operator per s-line (:261-285)
    def lines_r():
        A_rows, A_cols, A_vals = [], [], []
        S0t, SNt = ([], [], []), ([], [], [])
        for j in range(Nsp):
            b = crr[:, j]
            _, S0e, SNe, _, _, Ae, _ = variable_diagonal_sbp_D2(p, Nr, b)
            Ae = Ae.tocoo(); S0e = S0e.tocoo(); SNe = SNe.tocoo()
            A_rows.append(Ae.row + j * Nrp); A_cols.append(Ae.col + j * Nrp); A_vals.append(hs_d[j] * Ae.data)
            for T, S in ((S0t, S0e), (SNt, SNe)):
                T[0].append(S.row + j * Nrp); T[1].append(S.col + j * Nrp); T[2].append(hs_d[j] * S.data)
        cat = lambda L: np.concatenate(L)
        A = sp.csc_matrix((cat(A_vals), (cat(A_rows), cat(A_cols))), shape=(Np, Np))
        S0 = sp.csc_matrix((cat(S0t[2]), (cat(S0t[0]), cat(S0t[1]))), shape=(Np, Np))
        SN = sp.csc_matrix((cat(SNt[2]), (cat(SNt[0]), cat(SNt[1]))), shape=(Np, Np))
        return A, S0, SN

    # ss part: one per r-line (:313-339)
    def lines_s():
        A_rows, A_cols, A_vals = [], [], []
        S0t, SNt = ([], [], []), ([], [], [])
        for i in range(Nrp):
            b = css[i, :]
            _, S0e, SNe, _, _, Ae, _ = variable_diagonal_sbp_D2(p, Ns, b)
            Ae = Ae.tocoo(); S0e = S0e.tocoo(); SNe = SNe.tocoo()
            A_rows.append(i + Nrp * Ae.row); A_cols.append(i + Nrp * Ae.col); A_vals.append(hr_d[i] * Ae.data)
            for T, S in ((S0t, S0e), (SNt, SNe)):
                T[0].append(i + Nrp * S.row); T[1].append(i + Nrp * S.col); T[2].append(hr_d[i] * S.data)
        cat = lambda L: np.concatenate(L)
        A = sp.csc_matrix((cat(A_vals), (cat(A_rows), cat(A_cols))), shape=(Np, Np))
        S0 = sp.csc_matrix((cat(S0t[2]), (cat(S0t[0]), cat(S0t[1]))), shape=(Np, Np))
        SN = sp.csc_matrix((cat(SNt[2]), (cat(SNt[0]), cat(SNt[1]))), shape=(Np, Np))
        return A, S0, SN

    Arr, Sr0, SrN = lines_r()
    Ass, Ss0, SsN = lines_s()
    Sr0T, SrNT, Ss0T, SsNT = Sr0.T.tocsc(), SrN.T.tocsc(), Ss0.T.tocsc(), SsN.T.tocsc()

    crs_flat = crs.reshape(-1, order="F")
    Csr = sp.diags(crs_flat, format="csc")
    Asr = sp.kron(QsT, Ir, format="csc") @ Csr @ sp.kron(Is, Qr, format="csc")      # :352
    Ars = sp.kron(Is, QrT, format="csc") @ Csr @ sp.kron(Qs, Ir, format="csc")      # :353
    A = Arr + Ass + Ars + Asr                                                        # :356

    Er0 = sp.csc_matrix(([1.0], ([0], [0])), shape=(Nrp, Nrp))
    ErN = sp.csc_matrix(([1.0], ([Nr], [Nr])), shape=(Nrp, Nrp))
    Es0 = sp.csc_matrix(([1.0], ([0], [0])), shape=(Nsp, Nsp))
    EsN = sp.csc_matrix(([1.0], ([Ns], [Ns])), shape=(Nsp, Nsp))
    er0 = sp.csc_matrix(([1.0], ([0], [0])), shape=(Nrp, 1))
    erN = sp.csc_matrix(([1.0], ([Nr], [0])), shape=(Nrp, 1))
    es0 = sp.csc_matrix(([1.0], ([0], [0])), shape=(Nsp, 1))
    esN = sp.csc_matrix(([1.0], ([Ns], [0])), shape=(Nsp, 1))
    er0T, erNT, es0T, esNT = er0.T.tocsc(), erN.T.tocsc(), es0.T.tocsc(), esN.T.tocsc()

    crs0 = sp.diags(crs[:, 0], format="csc")        # :383  (s = first line, runs over r)
    crsN = sp.diags(crs[:, Ns], format="csc")       # :384
    csr0 = sp.diags(csr[0, :], format="csc")        # :385  (r = first point, runs over s)
    csrN = sp.diags(csr[Nr, :], format="csc")       # :386

    H1, H1I = Hs, HsI
    H2, H2I = Hs, HsI
    H3, H3I = Hr, HrI
    H4, H4I = Hr, HrI

    l, beta, alpha = PENALTY[p]
    psimin = (crr + css - np.sqrt((crr - css) ** 2 + 4 * crs ** 2)) / 2              # :418
    assert psimin.min() > 0
    hr = 2.0 / Nr
    hs = 2.0 / Ns
    psi1 = psimin[0, :].copy(); psi2 = psimin[Nr, :].copy()
    psi3 = psimin[:, 0].copy(); psi4 = psimin[:, Ns].copy()
    for k in range(1, l):                                                             # :428-433
        psi1 = np.minimum(psi1, psimin[k, :])
        psi2 = np.minimum(psi2, psimin[Nr - k, :])
        psi3 = np.minimum(psi3, psimin[:, k])
        psi4 = np.minimum(psi4, psimin[:, Ns - k])
    t1 = (2 * tauscale / hr) * (crr[0, :] ** 2 / beta + crs[0, :] ** 2 / alpha) / psi1
    t2 = (2 * tauscale / hr) * (crr[Nr, :] ** 2 / beta + crs[Nr, :] ** 2 / alpha) / psi2
    t3 = (2 * tauscale / hs) * (css[:, 0] ** 2 / beta + crs[:, 0] ** 2 / alpha) / psi3
    t4 = (2 * tauscale / hs) * (css[:, Ns] ** 2 / beta + crs[:, Ns] ** 2 / alpha) / psi4
    tau1, tau2 = sp.diags(t1, format="csc"), sp.diags(t2, format="csc")
    tau3, tau4 = sp.diags(t3, format="csc"), sp.diags(t4, format="csc")

    K = lambda a, b: sp.kron(a, b, format="csc")
    C1 = (Sr0 + Sr0T) + K(csr0 @ Qs + QsT @ csr0, Er0) + K(tau1 @ H1, Er0)          # :444
    C2 = -(SrN + SrNT) - K(csrN @ Qs + QsT @ csrN, ErN) + K(tau2 @ H2, ErN)
    C3 = (Ss0 + Ss0T) + K(Es0, crs0 @ Qr + QrT @ crs0) + K(Es0, tau3 @ H3)
    C4 = -(SsN + SsNT) - K(EsN, crsN @ Qr + QrT @ crsN) + K(EsN, tau4 @ H4)

    G1 = -K(Is, er0T) @ Sr0 - K(csr0 @ Qs, er0T)                                     # :450
    G2 = K(Is, erNT) @ SrN + K(csrN @ Qs, erNT)
    G3 = -K(es0T, Ir) @ Ss0 - K(es0T, crs0 @ Qr)
    G4 = K(esNT, Ir) @ SsN + K(esNT, crsN @ Qr)

    F1 = G1.T - K(tau1 @ H1, er0)                                                    # :455
    F2 = G2.T - K(tau2 @ H2, erN)
    F3 = G3.T - K(es0, tau3 @ H3)
    F4 = G4.T - K(esN, tau4 @ H4)

    HfI_F1T = H1I @ G1 - K(tau1, er0T)                                               # :460
    HfI_F2T = H2I @ G2 - K(tau2, erNT)
    HfI_F3T = H3I @ G3 - K(es0T, tau3)
    HfI_F4T = H4I @ G4 - K(esNT, tau4)
    HfI_G = (H1I @ G1, H2I @ G2, H3I @ G3, H4I @ G4)

    Mt = A + C1 + C2 + C3 + C4                                                       # :470
    F = tuple(f.tocsc() for f in (F1, F2, F3, F4))
    tau = (tau1, tau2, tau3, tau4)
    HfI = (H1I, H2I, H3I, H4I)
    for lf in range(4):                                                               # :477-486
        b = LFToB[lf]
        if b == BC_NEUMANN:
            Mt = Mt - F[lf] @ (sp.diags(1.0 / tau[lf].diagonal()) @ HfI[lf]) @ F[lf].T
        elif not (b == BC_DIRICHLET or b == BC_LOCKED_INTERFACE or b >= BC_JUMP_INTERFACE):
            raise ValueError("invalid bc")
    JH = sp.diags(J.reshape(-1, order="F"), format="csc") @ K(Hs, Hr)                # :491
    return SimpleNamespace(
        Mt=Mt.tocsc(), F=F,
        HfI_FT=tuple(m.tocsc() for m in (HfI_F1T, HfI_F2T, HfI_F3T, HfI_F4T)),
        HfI_G=tuple(m.tocsc() for m in HfI_G),
        G=tuple(m.tocsc() for m in (G1, G2, G3, G4)),
        A=A.tocsc(),
        coord=metrics.coord, facecoord=metrics.facecoord, JH=JH,
        sJ=metrics.sJ, nx=metrics.nx, ny=metrics.ny,
        Hf=(H1, H2, H3, H4), HfI=HfI, tau=tau,
        bctype=tuple(int(b) for b in LFToB), Nr=Nr, Ns=Ns, p=p)


# --------------------------------------------------------------------------
# trace operators (global_curved.jl:510-565)
# --------------------------------------------------------------------------
def glolambdaoperator(lop, vstarts, FToB, FToE, FToLF, EToO, EToS, Nr, Ns):
    nfaces = len(FToB)
    FTolstarts = np.zeros(nfaces + 1, dtype=np.int64)
    FTolstarts[0] = 1
    IT, JT, VT, VD = [], [], [], []
    for f in range(nfaces):
        if FToB[f] == BC_DIRICHLET or FToB[f] == BC_NEUMANN:
            FTolstarts[f + 1] = FTolstarts[f]
            continue
        em, ep = FToE[0, f] - 1, FToE[1, f] - 1
        fm, fp = FToLF[0, f] - 1, FToLF[1, f] - 1
        nl = (Ns[em] if fm <= 1 else Nr[em]) + 1
        assert nl == (Ns[ep] if fp <= 1 else Nr[ep]) + 1                             # :528
        FTolstarts[f + 1] = FTolstarts[f] + nl
        assert EToO[fm, em] and EToS[fm, em] == 1                                    # :531
        Fm = lop[em].F[fm].tocoo()
        IT.append(Fm.col + (FTolstarts[f] - 1))
        JT.append(Fm.row + (vstarts[em] - 1))
        VT.append(Fm.data)
        assert EToS[fp, ep] == 2
        Fp = lop[ep].F[fp].tocoo()
        tm = lop[em].tau[fm].diagonal()
        tp = lop[ep].tau[fp].diagonal()
        if EToO[fp, ep]:
            IT.append(Fp.col + (FTolstarts[f] - 1))
        else:
            IT.append((FTolstarts[f + 1] - 1) - 1 - Fp.col)                           # :549, flipped
            tp = tp[::-1]
        JT.append(Fp.row + (vstarts[ep] - 1))
        VT.append(Fp.data)
        VD.append(lop[em].Hf[fm].diagonal() * (tm + tp))
    lNp = FTolstarts[nfaces] - 1
    VNp = vstarts[len(lop)] - 1
    cat = lambda L, dt: np.concatenate(L) if L else np.zeros(0, dtype=dt)
    FbarT = sp.csc_matrix((cat(VT, float), (cat(IT, np.int64), cat(JT, np.int64))), shape=(lNp, VNp))
    return FTolstarts, FbarT, cat(VD, float)


# --------------------------------------------------------------------------
# right-hand sides (global_curved.jl:569-654)
# --------------------------------------------------------------------------
def locbcarray_mod(ge, lop, LFToB, bc_Dirichlet, bc_Neumann, bcargs=()):
    xf, yf = lop.facecoord
    ge[:] = 0
    for lf in range(4):
        if LFToB[lf] == BC_DIRICHLET:
            vf = bc_Dirichlet(lf + 1, xf[lf], yf[lf], *bcargs)
        elif LFToB[lf] == BC_NEUMANN:
            gN = bc_Neumann(lf + 1, xf[lf], yf[lf], lop.nx[lf], lop.ny[lf], *bcargs)
            vf = lop.sJ[lf] * gN / lop.tau[lf].diagonal()
        elif LFToB[lf] == BC_LOCKED_INTERFACE:
            continue
        else:
            raise ValueError("invalid bc")
        ge -= lop.F[lf] @ np.asarray(vf, float)


def locbcarray(ge, gde, lop, LFToB, bc_Dirichlet, bc_Neumann, in_jump, bcargs=()):
    """gde: 4 writable views into the global g_delta (already orientation-mapped)."""
    xf, yf = lop.facecoord
    ge[:] = 0
    for lf in range(4):
        if LFToB[lf] == BC_DIRICHLET:
            vf = bc_Dirichlet(lf + 1, xf[lf], yf[lf], *bcargs)
        elif LFToB[lf] == BC_NEUMANN:
            gN = bc_Neumann(lf + 1, xf[lf], yf[lf], lop.nx[lf], lop.ny[lf], *bcargs)
            vf = lop.sJ[lf] * gN / lop.tau[lf].diagonal()
        elif LFToB[lf] == BC_LOCKED_INTERFACE:
            continue
        elif LFToB[lf] >= BC_JUMP_INTERFACE:
            vf = in_jump(lf + 1, xf[lf], yf[lf], *bcargs) / 2
            gde[lf][:] -= lop.Hf[lf].diagonal() * lop.tau[lf].diagonal() * vf         # :616
        else:
            raise ValueError("invalid bc")
        ge -= lop.F[lf] @ np.asarray(vf, float)


def computetraction_mod(lop, lf, u, delta):
    """global_curved.jl:627-634 (lf is 1-based)."""
    k = lf - 1
    return (lop.HfI_FT[k] @ u + lop.tau[k].diagonal() * (delta - delta / 2)) / lop.sJ[k]


def computetraction(lop, lf, u, lam, delta):
    """global_curved.jl:638-644 (lf is 1-based)."""
    k = lf - 1
    return (lop.HfI_FT[k] @ u + lop.tau[k].diagonal() * (lam - delta / 2)) / lop.sJ[k]


def locsourcearray(ge, source, lop, volargs=()):
    x, y = lop.coord
    ge += lop.JH @ source(x.reshape(-1, order="F"), y.reshape(-1, order="F"), *volargs)


# --------------------------------------------------------------------------
# global containers (global_curved.jl:659-741)
# --------------------------------------------------------------------------
class _LU:
    """Stand-in for the object the reference's `factorization` callback returns."""

    def __init__(self, A):
        self.n = A.shape[0]
        self._lu = spla.splu(sp.csc_matrix(A))

    def solve(self, b):
        return self._lu.solve(np.asarray(b, float))


def default_factorization(A):
    return _LU(A)


def SBPLocalOperator1(lop, Nr, Ns, factorization=default_factorization):
    nelems = len(lop)
    vstarts = np.zeros(nelems + 1, dtype=np.int64)
    vstarts[0] = 1
    VH, X, Y, E, factors = [], [], [], [], []
    for e in range(nelems):
        Npe = (Nr[e] + 1) * (Ns[e] + 1)
        vstarts[e + 1] = vstarts[e] + Npe
        VH.append(lop[e].JH.diagonal())
        x, y = lop[e].coord
        X.append(x.reshape(-1, order="F")); Y.append(y.reshape(-1, order="F"))
        E.append(np.full(Npe, e + 1, dtype=np.int64))
        factors.append(factorization(lop[e].Mt))
    return SimpleNamespace(offset=vstarts, H=np.concatenate(VH), X=np.concatenate(X),
                           Y=np.concatenate(Y), E=np.concatenate(E), F=factors)


def LocalGlobalOperators(lop, Nr, Ns, FToB, FToE, FToLF, EToO, EToS,
                         factorization=default_factorization):
    M = SBPLocalOperator1(lop, Nr, Ns, factorization)
    FTolstarts, FbarT, D = glolambdaoperator(lop, M.offset, FToB, FToE, FToLF, EToO, EToS, Nr, Ns)
    return M, FbarT, D, M.offset, FTolstarts


def bcstarts(FToB, FToE, FToLF, bctype, Nr, Ns):
    if np.isscalar(bctype):
        bctype = (bctype,)
    nfaces = len(FToB)
    out = np.zeros(nfaces + 1, dtype=np.int64)
    out[0] = 1
    for f in range(nfaces):
        if FToB[f] in bctype:
            e = FToE[0, f] - 1
            lf = FToLF[0, f]
            out[f + 1] = out[f] + (Ns[e] if lf in (1, 2) else Nr[e]) + 1
        else:
            out[f + 1] = out[f]
    return out


def LocalToGLobalRHS(b, g, gd, u, factors, FbarT, vstarts):
    """b = gd - Fbar^T Mtilde^{-1} g   (global_curved.jl:730-740)."""
    u[:] = 0
    for e in range(len(factors)):
        sl = slice(vstarts[e] - 1, vstarts[e + 1] - 1)
        if np.max(np.abs(g[sl])) > 0:
            u[sl] = factors[e].solve(g[sl])
    b[:] = gd - FbarT @ u


def assemblelambdamatrix(FTolstarts, vstarts, EToF, FToB, factors, D, FbarT):
    """Explicit Schur complement B = D - Fbar^T Mtilde^{-1} Fbar (global_curved.jl:743-797)."""
    nfaces = len(FTolstarts) - 1
    nelems = len(vstarts) - 1
    lNp = FTolstarts[nfaces] - 1
    Fbar = FbarT.T.tocsc()
    FbarT_r = FbarT.tocsr()
    I = [np.arange(lNp)]; Jc = [np.arange(lNp)]; V = [np.asarray(D, float)]
    has_l = lambda f: FToB[f] == BC_LOCKED_INTERFACE or FToB[f] >= BC_JUMP_INTERFACE
    for e in range(nelems):
        v0, v1 = vstarts[e] - 1, vstarts[e + 1] - 1
        for lf in range(4):
            f = EToF[lf, e] - 1
            if not has_l(f):
                continue
            l0, l1 = FTolstarts[f] - 1, FTolstarts[f + 1] - 1
            rhs = Fbar[v0:v1, l0:l1].toarray()
            X = np.column_stack([factors[e].solve(rhs[:, c]) for c in range(rhs.shape[1])])
            for lf2 in range(4):
                f2 = EToF[lf2, e] - 1
                if not has_l(f2):
                    continue
                m0, m1 = FTolstarts[f2] - 1, FTolstarts[f2 + 1] - 1
                C = FbarT_r[m0:m1, v0:v1] @ X                   # (l2 x l1) = Fbar2^T M^-1 Fbar1
                rr, cc = np.meshgrid(np.arange(l0, l1), np.arange(m0, m1), indexing="ij")
                I.append(rr.ravel()); Jc.append(cc.ravel()); V.append(-(C.T).ravel())
    B = sp.csc_matrix((np.concatenate(V), (np.concatenate(I), np.concatenate(Jc))), shape=(lNp, lNp))
    assert abs(B - B.T).max() <= 1e-8 * abs(B).max()             # :794 (B ≈ B')
    return B


# --------------------------------------------------------------------------
# Abaqus .inp reader (global_curved.jl:802-956)
# --------------------------------------------------------------------------
def _seek(lines, pattern, first=0):
    rx = re.compile(pattern)
    for l in range(first, len(lines)):
        if rx.search(lines[l]):
            return l
    return -1


def read_inp_2d(filename, bc_map=None):
    """-> (verts 2 x nv, EToV 4 x ne, EToF 4 x ne, FToB nf, EToBlock ne); ids 1-based."""
    if bc_map is None:
        bc_map = list(range(1, 10001))
    try:
        with open(filename) as f:
            lines = f.read().split("\n")
    except OSError:
        raise RuntimeError('InpRead cannot open "%s" ' % filename)
    num_line = re.compile(r"^\s*[0-9]*\s*,.*")
    ln = _seek(lines, "NSET=ALLNODES")
    if ln < 0:
        raise RuntimeError("did not find: NSET=ALLNODES")
    nn = 0
    for l in range(ln + 1, len(lines)):
        if num_line.match(lines[l]):
            nn += 1
        else:
            break
    Vx = np.full(nn, np.nan); Vy = np.full(nn, np.nan)
    for l in range(ln + 1, ln + 1 + nn):
        d = [t for t in re.split(r"\s|,", lines[l]) if t]
        Vx[int(d[0]) - 1] = float(d[1]); Vy[int(d[0]) - 1] = float(d[2])
    # elements
    ne = 0
    ln = _seek(lines, "ELEMENT")
    while ln >= 0:
        for l in range(ln + 1, len(lines)):
            if num_line.match(lines[l]):
                ne += 1
            else:
                break
        ln = _seek(lines, "ELEMENT", ln + 1)
    if ne == 0:
        raise RuntimeError("did not find any element")
    EToV = np.zeros((4, ne), dtype=np.int64)
    EToBlock = np.zeros(ne, dtype=np.int64)
    ln = _seek(lines, "ELEMENT")
    while ln >= 0:
        blk = int([t for t in re.split(r"[^0-9]", lines[ln]) if t][-1])
        for l in range(ln + 1, min(ln + 1 + ne, len(lines))):
            d = [t for t in re.split(r"\s|,", lines[l]) if t]
            try:
                en, v1, v2, v4, v3 = (int(d[0]), int(d[1]), int(d[2]), int(d[3]), int(d[4]))
            except (ValueError, IndexError):
                break
            EToV[:, en - 1] = (v1, v2, v3, v4)      # z-order (:862-871)
            EToBlock[en - 1] = blk
        ln = _seek(lines, "ELEMENT", ln + 1)
    # faces
    EToF = np.zeros((4, ne), dtype=np.int64)
    seen = {}
    for e in range(ne):
        for lf, (a, b) in enumerate(_LFTOLV):
            vs = (EToV[a, e], EToV[b, e])
            if vs[0] > vs[1]:
                vs = (vs[1], vs[0])
            if vs not in seen:
                seen[vs] = len(seen) + 1
            EToF[lf, e] = seen[vs]
    nf = len(seen)
    FToB = np.full(nf, BC_LOCKED_INTERFACE, dtype=np.int64)
    inp_to_zorder = (3, 2, 4, 1)                     # :911
    ln = _seek(lines, r"\*ELSET")
    lead_int = re.compile(r"^\s*[0-9]+")
    while ln >= 0:
        foo = [t for t in re.split(r"[^0-9]", lines[ln]) if t]
        bc = bc_map[int(foo[0]) - 1]
        face = inp_to_zorder[int(foo[1]) - 1]
        for l in range(ln + 1, len(lines)):
            if not lead_int.match(lines[l]):
                break
            for tok in [t for t in re.split(r"\s|,", lines[l]) if t]:
                elm = int(tok)
                if bc == 3:
                    bc = BC_LOCKED_INTERFACE
                FToB[EToF[face - 1, elm - 1] - 1] = bc
                assert bc in (BC_DIRICHLET, BC_NEUMANN, BC_LOCKED_INTERFACE) or bc >= BC_JUMP_INTERFACE
        ln = _seek(lines, r"\*ELSET", ln + 1)
    return np.vstack([Vx, Vy]), EToV, EToF, FToB, EToBlock


# --------------------------------------------------------------------------
# rate-and-state friction (global_curved.jl:1031-1075)
# --------------------------------------------------------------------------
def rateandstate(V, psi, sigma_n, phi, eta, a, V0):
    Y = (1.0 / (2.0 * V0)) * np.exp(psi / a)
    f = a * np.arcsinh(V * Y)
    dfdV = a * (1.0 / np.sqrt(1 + (V * Y) ** 2)) * Y
    g = sigma_n * f + eta * V - phi
    dgdV = sigma_n * dfdV + eta
    return g, dgdV


def newtbndv(func, xL, xR, x, ftol=1e-6, maxiter=500, minchange=0.0, atolx=1e-4, rtolx=1e-4):
    fL, _ = func(xL)
    fR, _ = func(xR)
    if fL * fR > 0:
        return float("nan"), float("nan"), -maxiter
    f, df = func(x)
    dxlr = xR - xL
    for it in range(1, maxiter + 1):
        dx = -f / df
        x = x + dx
        if x < xL or x > xR or abs(dx) / dxlr < minchange:
            x = (xR + xL) / 2
            dx = (xR - xL) / 2
        f, df = func(x)
        if f * fL > 0:
            fL, xL = f, x
        else:
            fR, xR = f, x
        dxlr = xR - xL
        if abs(f) < ftol and abs(dx) < atolx + rtolx * (abs(dx) + abs(x)):
            return x, f, it
    return x, f, -maxiter
