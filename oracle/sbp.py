"""ORACLE -- test infrastructure only (never imported by the product path).

CPU restatement (numpy / scipy.sparse) of the reference's 1-D diagonal-norm
summation-by-parts operators:

  diagonal_sbp_D1            reference diagonal_sbp.jl:67-161
  variable_diagonal_sbp_D2   reference diagonal_sbp.jl:474-764

Coefficient data come from oracle/sbp_tables.json (made by
oracle/gen_sbp_tables.py from the reference's published tables).

Parity status: the reference ships no golden vectors and its host language is
not available in the build container.  Pinned (1) by executing the reference's
own statements: tests/refexec/minijulia.py interprets diagonal_sbp.jl as it
lies under /root/reference, and diagonal_sbp_D1 / variable_diagonal_sbp_D2 here
return the same matrices to 2e-15 (tests/test_reference_executed.py; a second,
regex-based execution of the coefficient statements is in
tests/test_reference_text_extraction.py); (2) by the reference's own
identities (tests/test_oracle_sbp.py): SBP property
Q + Q^T = diag(-1, 0, ..., 0, 1), accuracy conditions on polynomials, symmetry
and zero row sums of M, mirror symmetry of the two closures, the PSD remainder
of check_residual.jl:8-17 and the constant-coefficient limit.

Indexing: everything here is 0-based; "row i" of the reference is row i-1.
"""
import json
import os

import numpy as np
import scipy.sparse as sp

_HERE = os.path.dirname(os.path.abspath(__file__))
with open(os.path.join(_HERE, "sbp_tables.json")) as _f:
    TABLES = json.load(_f)


def _coef(c):
    """float of a stored coefficient: ["num","den"] -> float(num)/float(den)."""
    if isinstance(c, list):
        return float(int(c[0])) / float(int(c[1]))
    return float(c)


def d1_tables(p):
    t = TABLES["D1"][str(p)]
    return (np.array(t["d"]), np.array(t["bd"]), np.array(t["bhinv"]))


def diagonal_sbp_D1(p, N, xc=(-1.0, 1.0)):
    """(D, HI, H, r) as in diagonal_sbp.jl:67-161 (p in 2, 4, 6)."""
    if str(p) not in TABLES["D1"]:
        raise ValueError("Operators for order %d are not implemented" % p)
    d, bd, bhinv = d1_tables(p)
    bm, bn = bd.shape
    Np = N + 1
    if Np < 2 * bm or Np < bn:                      # :129-131
        raise ValueError("Grid not big enough to support the operator")
    h = (xc[1] - xc[0]) / N
    assert h > 0
    hv = np.ones(Np)
    hv[:bm] = 1.0 / bhinv                           # :137
    hv[Np - bm:] = 1.0 / bhinv[::-1]                # :138
    H = sp.diags(h * hv, format="csc")
    HI = sp.diags(1.0 / (h * hv), format="csc")
    half = p // 2
    rows, cols, vals = [], [], []
    for i in range(bm, Np - bm):                    # interior rows, :142-145
        for k in range(-half, half + 1):
            rows.append(i); cols.append(i + k); vals.append(d[k + half] / h)
    for i in range(bm):                             # closures, :147-152
        for j in range(bn):
            rows.append(i); cols.append(j); vals.append(bd[i, j] / h)
            rows.append(Np - 1 - i); cols.append(Np - 1 - j); vals.append(-bd[i, j] / h)
    D = sp.csc_matrix((vals, (rows, cols)), shape=(Np, Np))
    D.eliminate_zeros()
    r = np.linspace(xc[0], xc[1], Np)
    return D, HI, H, r


# interior rows of the variable-coefficient stiffness matrix M (before /h):
# INTERIOR[p][o] = list of (shift, c): M[i, i+o] = sum c * b[i + shift]
INTERIOR = {
    # diagonal_sbp.jl:495-503
    2: {-1: [(-1, -0.5), (0, -0.5)],
        0: [(-1, 0.5), (0, 1.0), (1, 0.5)],
        1: [(0, -0.5), (1, -0.5)]},
    # diagonal_sbp.jl:567-582
    4: {-2: [(0, 1 / 8), (-1, -1 / 6), (-2, 1 / 8)],
        -1: [(1, -1 / 6), (0, -1 / 2), (-1, -1 / 2), (-2, -1 / 6)],
        0: [(2, 1 / 24), (1, 5 / 6), (0, 3 / 4), (-1, 5 / 6), (-2, 1 / 24)],
        1: [(2, -1 / 6), (1, -1 / 2), (0, -1 / 2), (-1, -1 / 6)],
        2: [(2, 1 / 8), (1, -1 / 6), (0, 1 / 8)]},
    # diagonal_sbp.jl:719-727 (as written there, including the line the
    # reference flags with "Bug here?")
    6: {-3: [(-3, -11 / 360), (-2, 1 / 40), (-1, 1 / 40), (0, -11 / 360)],
        -2: [(-3, 1 / 20), (-2, 7 / 40), (-1, -3 / 10), (0, 7 / 40), (1, 1 / 20)],
        -1: [(-3, -1 / 40), (-2, -3 / 10), (-1, -17 / 40), (0, -17 / 40), (1, -3 / 10), (2, -1 / 40)],
        0: [(-3, 1 / 180), (-2, 1 / 8), (-1, 19 / 20), (0, 101 / 180), (1, 19 / 20), (2, 1 / 8), (3, 1 / 180)],
        1: [(-2, -1 / 40), (-1, -3 / 10), (0, -17 / 40), (1, -17 / 40), (2, -3 / 10), (3, -1 / 40)],
        2: [(-1, 1 / 20), (0, 7 / 40), (1, -3 / 10), (2, 7 / 40), (3, 1 / 20)],
        3: [(0, -11 / 360), (1, 1 / 40), (2, 1 / 40), (3, -11 / 360)]},
}
# the p = 2 stencil above is written relative to row i for every offset:
#   M[i,i-1] = -(b[i-1]+b[i])/2, M[i,i] = (b[i-1]+2b[i]+b[i+1])/2, M[i,i+1] = -(b[i]+b[i+1])/2

D2VAR_BS = {2: np.array([3 / 2, -2.0, 1 / 2]),
            4: np.array(TABLES["D2var"]["4"]["BS"]),
            6: np.array(TABLES["D2var"]["6"]["BS"])}
D2VAR_BHINV = {2: np.array([2.0]),
               4: np.array(TABLES["D2var"]["4"]["bhinv"]),
               6: np.array(TABLES["D2var"]["6"]["bhinv"])}


def closure_terms(p):
    """{(i, j): [(k, c)]} 0-based closure block, M0[i,j] = sum c*b[k]."""
    if p == 2:                                       # diagonal_sbp.jl:488-490
        return {(0, 0): [(0, 0.5), (1, 0.5)]}, 1
    t = TABLES["D2var"][str(p)]
    out = {}
    for i, j, terms in t["closure"]:
        out[(i - 1, j - 1)] = [(k - 1, _coef(c)) for k, c in terms]
    return out, t["size"]


def stiffness_dense(p, N, b):
    """Dense (N+1)x(N+1) M of diagonal_sbp.jl:485-733 *before* the division by h."""
    b = np.asarray(b, dtype=float)
    Np = N + 1
    assert b.shape == (Np,)
    clo, m = closure_terms(p)
    half = p // 2
    if p == 2:
        nclo = 1
        minNp = 2
    else:
        nclo = m
        minNp = 2 * m                                # enough room for both blocks
    assert Np >= minNp
    M = np.zeros((Np, Np))
    for i in range(Np):
        for o in range(-half, half + 1):
            j = i + o
            if j < 0 or j >= Np:
                continue
            in_first = i < nclo and j < nclo
            in_last = i >= Np - nclo and j >= Np - nclo
            if in_first or in_last:
                continue
            if p == 2:
                # the reference writes the tridiagonal part for all rows (:492-503)
                # and *adds* the two corner entries (:488-494): sparse() sums duplicates
                pass
            acc = 0.0
            first = True
            for sh, c in INTERIOR[p][o]:
                term = c * b[i + sh]
                acc = term if first else acc + term
                first = False
            M[i, j] = acc
    if p == 2:
        # rows 2..N (1-based) carry the diagonal formula, rows 1 and N+1 only the
        # closure value; the off-diagonals cover all neighbours (:492-503)
        M[0, 0] = (b[0] + b[1]) / 2
        M[N, N] = (b[N - 1] + b[N]) / 2
        return M
    for (i, j), terms in clo.items():
        acc0 = 0.0
        accN = 0.0
        first = True
        for k, c in terms:
            t0 = c * b[k]
            tN = c * b[N - k]
            acc0 = t0 if first else acc0 + t0
            accN = tN if first else accN + tN
            first = False
        M[i, j] = acc0
        M[N - i, N - j] = accN
    return M


def variable_diagonal_sbp_D2(p, N, B, xc=(-1.0, 1.0)):
    """(D, S0, SN, HI, H, M, r) as in diagonal_sbp.jl:482-764; M is already /h (:746)."""
    if p not in (2, 4, 6):
        raise ValueError("Operators for order %d are not implemented" % p)
    B = np.asarray(B, dtype=float)
    Np = N + 1
    assert B.shape == (Np,)                          # :483
    bhinv = D2VAR_BHINV[p]
    BS = D2VAR_BS[p]
    bm = len(bhinv)
    if Np < 2 * bm:
        raise ValueError("Grid not big enough to support the operator")
    h = (xc[1] - xc[0]) / N
    assert h > 0
    M = sp.csc_matrix(stiffness_dense(p, N, B) / h)  # :746
    hv = np.ones(Np)
    hv[:bm] = bhinv
    hv[Np - bm:] = bhinv[::-1]
    HI = sp.diags(hv / h, format="csc")              # :751
    H = sp.diags(h / hv, format="csc")               # :752
    nb = len(BS)
    S0 = sp.csc_matrix((-B[0] * BS / h, (np.zeros(nb, int), np.arange(nb))), shape=(Np, Np))       # :755
    SN = sp.csc_matrix((B[N] * BS / h, (np.full(nb, N), N - np.arange(nb))), shape=(Np, Np))       # :756-757
    D = HI @ (-M + SN - S0)                          # :758
    r = np.linspace(xc[0], xc[1], Np)
    return D, S0, SN, HI, H, M, r
