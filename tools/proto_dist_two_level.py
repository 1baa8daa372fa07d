"""Prototype (CPU, oracle only; ranks emulated in one process) of the library-resident distributed trace CG:

  * PCG on B = D - Fbar^T M^-1 Fbar over a mesh partitioned by blocks, lambda replicated on cut faces;
  * first level: exact face blocks B_ff (explicit inverses);
  * second level: q Legendre modes per face, coarse matrix A_c = Z^T B Z eliminated rank by rank:
        I_r = coarse dofs of rank r's uncut faces (A_II is block diagonal over ranks),  G = dofs of the cut faces,
        S_G = A_GG - sum_r A_GI_r A_II_r^-1 A_IG_r   (small, replicated),
    so that one iteration needs TWO reductions:
        (1) p.q                                  (sum of rank-local parts, available before the cut-face exchange)
        (2) [r.z1, r.r, b_I.t, y]                (y = b_G - E^T b_I, t = A_II^-1 b_I, E = A_II^-1 A_IG)
    after which  c_G = S_G^-1 y,  c_I = t - E c_G,  r.z = r.z1 + b_I.t + y.c_G  on every rank.

usage: python tools/proto_dist_two_level.py [world] [nbx_per_rank] [nby] [N] [p] [q]"""
import sys
import numpy as np
import scipy.sparse as sp
sys.path.insert(0, ".")
from hybridsbp_b200 import synthetic, parallel
from hybridsbp_b200.host import connectivityarrays
from oracle import hybrid as orc
from tests.util import warped_metrics

world = int(sys.argv[1]) if len(sys.argv) > 1 else 3
nbx = int(sys.argv[2]) if len(sys.argv) > 2 else 2
nby = int(sys.argv[3]) if len(sys.argv) > 3 else 3
N = int(sys.argv[4]) if len(sys.argv) > 4 else 11
p = int(sys.argv[5]) if len(sys.argv) > 5 else 4
q = int(sys.argv[6]) if len(sys.argv) > 6 else 2

gnbx = nbx * world
_, EToV, EToF, FToB = synthetic.block_grid_connectivity(gnbx, nby)
FToE, FToLF, EToO, EToS = connectivityarrays(EToV, EToF)
ne = gnbx * nby
owner = (np.arange(ne) % gnbx) // nbx
lops = [orc.locoperator(p, N, N, warped_metrics(p, N, N, e % gnbx, e // gnbx, gnbx, nby), FToB[EToF[:, e] - 1]) for e in range(ne)]
M, FbarT, D, vstarts, Fl = orc.LocalGlobalOperators(lops, [N] * ne, [N] * ne, FToB, FToE, FToLF, EToO, EToS)
B = orc.assemblelambdamatrix(Fl, vstarts, EToF, FToB, M.F, D, FbarT).toarray()
B = 0.5 * (B + B.T)
n = B.shape[0]
starts = np.asarray(Fl) - 1
rng = np.random.default_rng(0)
rhs = rng.uniform(-1, 1, n)
nl = N + 1
s = np.linspace(-1, 1, nl)
Lq = np.stack([np.polynomial.legendre.Legendre.basis(k)(s) for k in range(q)], axis=1)

# ---- reference: global two-level PCG ---------------------------------------------------------------------
lam_faces = [f for f in range(len(FToB)) if starts[f + 1] > starts[f]]
Binv = {f: np.linalg.inv(B[starts[f]:starts[f + 1], starts[f]:starts[f + 1]]) for f in lam_faces}
Z = np.zeros((n, q * len(lam_faces)))
for i, f in enumerate(lam_faces):
    Z[starts[f]:starts[f + 1], q * i:q * i + q] = Lq
Acinv = np.linalg.inv(Z.T @ B @ Z)


def prec_ref(r):
    z = np.zeros(n)
    for f in lam_faces:
        z[starts[f]:starts[f + 1]] = Binv[f] @ r[starts[f]:starts[f + 1]]
    return z + Z @ (Acinv @ (Z.T @ r))


def pcg_ref(tol=1e-10, maxit=2000):
    x = np.zeros(n); r = rhs.copy(); z = prec_ref(r); pp = z.copy(); rz = r @ z; b2 = rhs @ rhs
    for it in range(1, maxit + 1):
        Bp = B @ pp
        al = rz / (pp @ Bp)
        x += al * pp; r -= al * Bp
        z = prec_ref(r); rz2 = r @ z
        if np.sqrt(r @ r / b2) <= tol:
            return x, it
        pp = z + (rz2 / rz) * pp; rz = rz2
    return x, maxit


x_ref, it_ref = pcg_ref()

# ---- rank-local data ----------------------------------------------------------------------------------------
# global numbering of the cut faces (Gamma): increasing global face id
is_cut = np.array([starts[f + 1] > starts[f] and owner[FToE[0, f] - 1] != owner[FToE[1, f] - 1] for f in range(len(FToB))])
gamma_of = -np.ones(len(FToB), dtype=np.int64)
gamma_of[is_cut] = np.arange(is_cut.sum())
nG = int(is_cut.sum())


class Rank:
    def __init__(self, rank):
        self.rank = rank
        lm = parallel.localize(rank, owner, EToF, FToB, FToE, FToLF, EToO, EToS)
        self.lm = lm
        self.faces = [int(f) for f in lm.faces if starts[f + 1] > starts[f]]       # global ids of local lambda faces
        self.rows = np.concatenate([np.arange(starts[f], starts[f + 1]) for f in self.faces])
        self.off = {f: i * nl for i, f in enumerate(self.faces)}
        cols = np.concatenate([np.arange(vstarts[e] - 1, vstarts[e + 1] - 1) for e in lm.blocks])
        FT = FbarT.tocsr()[self.rows][:, cols].toarray()
        Minv = np.linalg.inv(sp.block_diag([lops[e].Mt for e in lm.blocks]).toarray())
        self.S = FT @ Minv @ FT.T                                  # this rank's side of Fbar^T M^-1 Fbar (local lambda layout)
        self.D = D[self.rows]                                      # completed D
        self.cut = [f for f in self.faces if is_cut[f]]
        self.unc = [f for f in self.faces if not is_cut[f]]
        self.owned = {f: (not is_cut[f]) or owner[FToE[0, f] - 1] == rank for f in self.faces}
        self.w = np.concatenate([np.full(nl, 1.0 if self.owned[f] else 0.0) for f in self.faces])
        self.nloc = len(self.rows)

    def sl(self, f):
        return slice(self.off[f], self.off[f] + nl)


ranks = [Rank(r) for r in range(world)]


def exchange_sum(vals):
    """vals[r][f] = this rank's part on cut face f  ->  total[f] (what send/recv + add gives both ranks)"""
    tot = {}
    for r in ranks:
        for f in r.cut:
            tot[f] = tot.get(f, 0) + vals[r.rank][f]
    return tot


# first level: face-block inverses (cut faces: D_f - (own + partner))
own = {r.rank: {f: r.S[r.sl(f), r.sl(f)] for f in r.cut} for r in ranks}
tot = exchange_sum(own)
for r in ranks:
    r.Binv = {}
    for f in r.faces:
        Sff = tot[f] if is_cut[f] else r.S[r.sl(f), r.sl(f)]
        r.Binv[f] = np.linalg.inv(np.diag(r.D[r.sl(f)]) - Sff)

# second level
for r in ranks:
    nI, nGl = q * len(r.unc), q * len(r.cut)
    Zl = np.zeros((r.nloc, q * len(r.faces)))
    order = r.unc + r.cut                                  # coarse columns: uncut faces first, then the cut ones
    for i, f in enumerate(order):
        Zl[r.sl(f), q * i:q * i + q] = Lq
    # rank-local part of Z^T B Z: -Z^T S Z for all local faces + Z^T D Z for owned faces
    A = -Zl.T @ r.S @ Zl + Zl.T @ ((r.w * r.D)[:, None] * Zl)
    r.Zl, r.nI, r.nGl = Zl, nI, nGl
    r.AIIinv = np.linalg.inv(A[:nI, :nI])
    r.E = r.AIIinv @ A[:nI, nI:]                           # nI x nGl
    r.gidx = np.concatenate([q * gamma_of[f] + np.arange(q) for f in r.cut]) if r.cut else np.zeros(0, dtype=np.int64)
    r.SG_part = np.zeros((q * nG, q * nG))
    r.SG_part[np.ix_(r.gidx, r.gidx)] = A[nI:, nI:] - A[nI:, :nI] @ r.E
SGinv = np.linalg.inv(sum(r.SG_part for r in ranks))       # all-reduce at setup, replicated


def dist_pcg(tol=1e-10, maxit=2000):
    st = [dict(lam=np.zeros(r.nloc), r=rhs[r.rows].copy(), p=None) for r in ranks]
    b2 = sum((s_["r"] * r.w) @ s_["r"] for r, s_ in zip(ranks, st))

    def precond(first=False):
        # local: z1, partials, coarse restriction, t = AIIinv b_I, y part -> ONE all-reduce -> finish
        red = np.zeros(3 + q * nG)
        loc = []
        for r, s_ in zip(ranks, st):
            z1 = np.zeros(r.nloc)
            for f in r.faces:
                z1[r.sl(f)] = r.Binv[f] @ s_["r"][r.sl(f)]
            bc = r.Zl.T @ (r.w * s_["r"])                # owner-counted restriction (cut faces counted once)
            bI, bG = bc[:r.nI], bc[r.nI:]
            t = r.AIIinv @ bI
            y = np.zeros(q * nG)
            y[r.gidx] += bG
            y[r.gidx] -= r.E.T @ bI
            red += np.concatenate([[(r.w * s_["r"]) @ z1, (r.w * s_["r"]) @ s_["r"], bI @ t], y])
            loc.append((z1, t))
        rz1, rr, s1, y = red[0], red[1], red[2], red[3:]
        cG = SGinv @ y
        rz = rz1 + s1 + y @ cG
        zs = []
        for r, (z1, t) in zip(ranks, loc):
            c = np.concatenate([t - r.E @ cG[r.gidx], cG[r.gidx]])
            zs.append(z1 + r.Zl @ c)
        return zs, rz, rr

    zs, rz, rr = precond()
    for s_, z in zip(st, zs):
        s_["p"] = z.copy()
    for it in range(1, maxit + 1):
        # matvec: local part, then reduction (1) concurrently with the cut-face exchange
        pq = 0.0
        qs, contrib = [], {}
        for r, s_ in zip(ranks, st):
            c = r.S @ s_["p"]                              # this rank's side of Fbar^T M^-1 Fbar p
            pq += (r.w * r.D * s_["p"]) @ s_["p"] - s_["p"] @ c
            contrib[r.rank] = {f: c[r.sl(f)] for f in r.cut}
            qs.append((r.D * s_["p"], c))
        tot_c = exchange_sum(contrib)
        al = rz / pq
        for r, s_, (base, c) in zip(ranks, st, qs):
            qv = base - c
            for f in r.cut:
                qv[r.sl(f)] = base[r.sl(f)] - tot_c[f]
            s_["lam"] += al * s_["p"]
            s_["r"] -= al * qv
        zs, rz2, rr = precond()
        if np.sqrt(rr / b2) <= tol:
            break
        for s_, z in zip(st, zs):
            s_["p"] = z + (rz2 / rz) * s_["p"]
        rz = rz2
    x = np.zeros(n)
    for r, s_ in zip(ranks, st):
        x[r.rows] = s_["lam"]
    # replicated entries identical?
    for f in np.where(is_cut)[0]:
        vals = [s_["lam"][r.sl(f)] for r, s_ in zip(ranks, st) if f in r.off]
        assert len(vals) == 2 and np.array_equal(vals[0], vals[1])
    return x, it


x_d, it_d = dist_pcg()
print("mesh %d x %d blocks (%d ranks), %d lambda points, %d cut faces" % (gnbx, nby, world, n, nG))
print("global two-level PCG : %d iterations" % it_ref)
print("distributed          : %d iterations, |x - x_ref| / |x_ref| = %.2e" % (it_d, np.linalg.norm(x_d - x_ref) / np.linalg.norm(x_ref)))
x_dir = np.linalg.solve(B, rhs)
print("vs direct solve      : %.2e" % (np.linalg.norm(x_d - x_dir) / np.linalg.norm(x_dir)))
