"""Numpy emulation of the line-marching volume kernel (hybridsbp_b200/csrc/k_sweep.cuh).

This is host-side test infrastructure: it restates, step for step, the algorithm the CUDA kernel
uses -- pair form of the stiffness terms, s-direction register windows with a lag of H lines, the
"march away from the block end" prologue that evaluates closure rows directly, the read-modify-write
treatment of the dense closure block of Q^T, and the chunking / direction logic -- so that the
algebra can be checked against the oracle's assembled sparse operator on the CPU
(tests/test_sweep_algorithm.py) before any GPU time is spent.  It is never on the product path.

Volume operator (reference locoperator, global_curved.jl:261-353):
    A u = Arr u + Ass u + Qs^T[crs o (Qr u)] + Qr^T[crs o (Qs u)]
"""
import numpy as np

from sbp_coeffs import Coeffs


# ---- 1-D operators along r (whole lines; the kernel does these with lanes + closure rows) -------
def m_apply_line(cf, b, u):
    """M(b) u on one line in pair form (both closures)."""
    Np = len(u)
    N = Np - 1
    H, MC = cf.H, cf.MC
    out = np.zeros(Np)
    # interior-formula pairs (a, a+o) not inside a closure block
    for a in range(Np):
        for o in range(1, H + 1):
            j = a + o
            if j > N:
                continue
            if j < MC or a > N - MC:
                continue                     # both ends inside a closure block
            if a < MC or j > N - MC:
                continue                     # row a (or row j) is a closure row: handled below, but the
                                             # interior row of the pair still needs the coupling
            c = cf.pair_interior(a, o, b)
            f = c * (u[j] - u[a])
            out[a] += f
            out[j] -= f
    # couplings between a closure row and an interior row: interior row side
    for a in range(max(0, MC - H), MC):
        for o in range(1, H + 1):
            j = a + o
            if j >= MC:
                c = cf.pair_interior(a, o, b)
                out[j] -= c * (u[j] - u[a])
    br, ur = b[::-1], u[::-1]
    for a in range(max(0, MC - H), MC):
        for o in range(1, H + 1):
            j = a + o
            if j >= MC:
                c = cf.pair_interior(a, o, br)
                out[N - j] -= c * (ur[j] - ur[a])
    # closure rows, evaluated directly
    for i in range(MC):
        out[i] = cf.closure_row(i, b, u)
        out[N - i] = cf.closure_row(i, br, ur)
    return out


def q_apply_line(cf, u):
    Np = len(u)
    N = Np - 1
    H, BM, BN = cf.H, cf.BM, cf.BN
    out = np.zeros(Np)
    for i in range(BM, Np - BM):
        out[i] = sum(cf.d[o + H] * u[i + o] for o in range(-H, H + 1) if o != 0)
    for i in range(BM):
        out[i] = sum(cf.Qc[i, j] * u[j] for j in range(BN))
        out[N - i] = -sum(cf.Qc[i, j] * u[N - j] for j in range(BN))
    return out


def qt_apply_line(cf, w):
    """Q^T w: rows i >= BM are minus the interior stencil, rows i < BM use the QTc table."""
    Np = len(w)
    N = Np - 1
    H, BM, BN = cf.H, cf.BM, cf.BN
    out = np.zeros(Np)
    for i in range(BM, Np - BM):
        out[i] = -sum(cf.d[o + H] * w[i + o] for o in range(-H, H + 1) if o != 0)
    for i in range(BM):
        out[i] = sum(cf.QTc[i, k] * w[k] for k in range(BN))
        out[N - i] = -sum(cf.QTc[i, k] * w[N - k] for k in range(BN))
    return out


# ---- the marching algorithm ---------------------------------------------------------------------
def plan_chunks(cf, Nsp, nchunks_per_side):
    """Lower half marches up from line 0, upper half marches down from line Ns; each half is cut
    into chunks.  Returns [(direction, o0, o1)] in marching coordinates of that direction."""
    K = Nsp // 2
    out = []
    for direction, n in ((+1, K), (-1, Nsp - K)):
        per = -(-n // nchunks_per_side)
        o0 = 0
        while o0 < n:
            out.append((direction, o0, min(n, o0 + per)))
            o0 += per
    return out


def sweep_chunk(cf, y, u, crr, css, crs, direction, o0, o1):
    """Process marching lines [o0, o1) of one block.  Arrays are [Nrp, Nsp]; y is written in place."""
    H, MC, BM, BN, NK, LB = cf.H, cf.MC, cf.BM, cf.BN, cf.NK, cf.LB
    Nrp, Nsp = u.shape
    Nr, Ns = Nrp - 1, Nsp - 1
    hr, hs = 2.0 / Nr, 2.0 / Ns
    sig = float(direction)
    L = (lambda m: m) if direction > 0 else (lambda m: Ns - m)
    prologue = (o0 == 0)
    sc_ss = np.array([hr * cf.hweight(i, Nr) / hs for i in range(Nrp)])
    W = 2 * H + 1
    uw = [np.zeros(Nrp) for _ in range(W)]        # u at lines j-2H .. j
    bw = [np.zeros(Nrp) for _ in range(LB)]       # scaled css at lines j-2H+1 .. j
    cw = [np.zeros(Nrp) for _ in range(H + 1)]    # sigma*crs at lines j-H .. j
    acc = [np.zeros(Nrp) for _ in range(W)]       # accumulators of lines j-H .. j+H
    jstart = 0 if prologue else o0 - H
    jend = o1 - 1 + H
    assert jstart >= 0 and jend <= Ns - MC - H, "chunk must stay clear of the far closure"
    if prologue:
        for l in range(BM):
            y[:, L(l)] = 0.0                      # lines that receive read-modify-write contributions
    for j in range(jstart, jend + 1):
        jl = L(j)
        U = u[:, jl]
        uw = uw[1:] + [U.copy()]
        bw = bw[1:] + [css[:, jl] * sc_ss]
        cw = cw[1:] + [sig * crs[:, jl]]
        acc = acc[1:] + [np.zeros(Nrp)]
        # -- arrival: r-direction work on line j
        rr = m_apply_line(cf, crr[:, jl], U)
        acc[H] = acc[H] + (hs * cf.hweight(j, Ns) / hr) * rr
        t = cw[H] * q_apply_line(cf, U)
        if j >= BM:
            for o in range(-H, H + 1):
                if o != 0:
                    acc[H + o] = acc[H + o] + cf.d[o + H] * t          # (Q^T t)(j+o) += Q[j][j+o] t(j)
        else:
            for l in range(BN):
                c = cf.Qc[j, l]
                if c == 0.0:
                    continue
                if l < BM:
                    y[:, L(l)] += c * t                                  # dense block: global RMW
                else:
                    o = l - j
                    assert 1 <= o <= H
                    acc[H + o] = acc[H + o] + c * t
        # -- s-direction stiffness: pairs (a, a+o), a = j-H  (coefficient needs b up to line a+H = j)
        a = j - H
        if a >= 0:
            bline = lambda x: bw[LB - 1 - (j - x)]        # scaled css of line x, x in [j-2H+1, j]
            uline = lambda x: uw[W - 1 - (j - x)]         # u of line x, x in [j-2H, j]
            for o in range(1, H + 1):
                if prologue and a + o < MC:
                    continue                              # pair inside the closure block: direct rows below
                c = sum(cc * bline(a + sh) for sh, cc in cf.interior_pair[o])
                f = c * (uline(a + o) - uline(a))
                if not (prologue and a < MC):
                    acc[0] = acc[0] + f                   # row a
                acc[o] = acc[o] - f                       # row a+o
        # -- output of line jo = j-H
        jo = j - H
        if o0 <= jo < o1:
            if jo >= BM:
                qs = sum(cf.d[o + H] * uw[W - 1 - (j - (jo + o))] for o in range(-H, H + 1) if o != 0)
            else:
                qs = sum(cf.Qc[jo, l] * u[:, L(l)] for l in range(BN))   # direct loads
            w = cw[0] * qs
            val = acc[0] + qt_apply_line(cf, w)
            if prologue and jo < MC:
                bcol = [css[:, L(k)] for k in range(NK)]
                ucol = [u[:, L(k)] for k in range(NK)]
                val = val + sc_ss * cf.closure_row(jo, bcol, ucol)
            if prologue and jo < BM:
                y[:, L(jo)] += val
            else:
                y[:, L(jo)] = val


def sweep_block(p, u, crr, css, crs, nchunks_per_side=2):
    cf = Coeffs(p)
    y = np.full_like(u, np.nan)
    for direction, o0, o1 in plan_chunks(cf, u.shape[1], nchunks_per_side):
        sweep_chunk(cf, y, u, crr, css, crs, direction, o0, o1)
    return y
