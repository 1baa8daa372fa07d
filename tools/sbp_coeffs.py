"""Coefficient data of the line-marching kernel (hybridsbp_b200/csrc/k_sweep.cuh), derived from
oracle/sbp_tables.json.  Single source of truth for

  * tools/gen_sweep_tables.py  (writes hybridsbp_b200/csrc/sweep_tables_gen.h)
  * tools/proto_sweep.py       (numpy emulation of the kernel's algorithm, CPU-tested against the oracle)

Everything is in "pair form": the variable-coefficient stiffness matrix M(b) of
diagonal_sbp.jl:474-746 is symmetric with zero row sums, so
    (M u)_i = sum_{j != i} M_ij (u_j - u_i)
and only the off-diagonal couplings M_ij = sum_k c_k b_k are needed:
  interior couplings  M[i][i+o], o = 1..H, from the interior stencil (diagonal_sbp.jl:495-503, 567-582, 719-727)
  closure couplings   M[i][j], i < j < MC, from the closure block (diagonal_sbp.jl:515-537, 595-640)
"""
import json
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
with open(os.path.join(os.path.dirname(_HERE), "oracle", "sbp_tables.json")) as _f:
    TABLES = json.load(_f)


def _coef(c):
    if isinstance(c, list):
        return float(int(c[0])) / float(int(c[1]))
    return float(c)


# interior couplings M[i][i+o] = sum_sh c * b[i+sh]   (o > 0 only; the matrix is symmetric)
INTERIOR_PAIR = {
    2: {1: [(0, -0.5), (1, -0.5)]},
    4: {1: [(2, -1 / 6), (1, -1 / 2), (0, -1 / 2), (-1, -1 / 6)],
        2: [(2, 1 / 8), (1, -1 / 6), (0, 1 / 8)]},
    6: {1: [(-2, -1 / 40), (-1, -3 / 10), (0, -17 / 40), (1, -17 / 40), (2, -3 / 10), (3, -1 / 40)],
        2: [(-1, 1 / 20), (0, 7 / 40), (1, -3 / 10), (2, 7 / 40), (3, 1 / 20)],
        3: [(0, -11 / 360), (1, 1 / 40), (2, 1 / 40), (3, -11 / 360)]},
}


class Coeffs:
    """All constants of order p in {2, 4, 6}."""

    def __init__(self, p):
        self.p = p
        self.H = p // 2
        d1 = TABLES["D1"][str(p)]
        self.d = np.array(d1["d"], dtype=float)                 # interior first-derivative stencil, offsets -H..H
        bd = np.array(d1["bd"], dtype=float)
        self.BM, self.BN = bd.shape
        self.hw = 1.0 / np.array(d1["bhinv"], dtype=float)      # norm weights of the first BM points (units of h)
        # closure rows of Q = H D (pure numbers): Qc[k][j] = hw[k] * bd[k][j]
        self.Qc = self.hw[:, None] * bd
        # first BM rows of Q^T over columns 0..BN-1: QTc[i][k] = Q[k][i]; rows k >= BM are interior rows
        Qbig = self.Q_dense(4 * self.BN)
        self.QTc = Qbig.T[: self.BM, : self.BN].copy()
        assert np.all(Qbig.T[: self.BM, self.BN:] == 0)
        # second derivative
        self.interior_pair = INTERIOR_PAIR[p]
        if p == 2:
            self.MC, self.NK = 1, 2
            self.closure_pair = {}
            self.BS = np.array([1.5, -2.0, 0.5])
        else:
            t = TABLES["D2var"][str(p)]
            self.MC = t["size"]
            self.NK = {4: 8, 6: 12}[p]
            self.BS = np.array(t["BS"], dtype=float)
            self.closure_pair = {}
            for i, j, terms in t["closure"]:
                if i < j:
                    self.closure_pair[(i - 1, j - 1)] = [(k - 1, _coef(c)) for k, c in terms]
        # window lengths of the marching kernel
        self.LB = 2 * self.H          # b window: M[a][a+o] at row a = j-H needs b(j-2H+1 .. j)
        sh_min = min(sh for o in self.interior_pair.values() for sh, _ in o)
        sh_max = max(sh for o in self.interior_pair.values() for sh, _ in o)
        assert sh_max == self.H and sh_min >= -(self.H - 1), (sh_min, sh_max)

    def Q_dense(self, Np):
        """Q = H*D on Np points (pure numbers)."""
        H, BM, BN = self.H, self.BM, self.BN
        Q = np.zeros((Np, Np))
        for i in range(BM, Np - BM):
            for o in range(-H, H + 1):
                Q[i, i + o] = self.d[o + H]
        for i in range(BM):
            for j in range(BN):
                Q[i, j] = self.hw[i] * TABLES["D1"][str(self.p)]["bd"][i][j]
                Q[Np - 1 - i, Np - 1 - j] = -Q[i, j]
        return Q

    def hweight(self, i, N):
        if i < self.BM:
            return self.hw[i]
        if i > N - self.BM:
            return self.hw[N - i]
        return 1.0

    # ---- pair couplings ------------------------------------------------------------------
    def pair_interior(self, a, o, b):
        """M[a][a+o] from the interior stencil; b is indexable by absolute index."""
        acc = 0.0
        for sh, c in self.interior_pair[o]:
            acc = acc + c * b[a + sh]
        return acc

    def pair_closure(self, i, j, b):
        """M[i][j], i < j < MC, near-end closure; b indexable 0..NK-1."""
        acc = 0.0
        for k, c in self.closure_pair[(i, j)]:
            acc = acc + c * b[k]
        return acc

    def closure_row(self, i, b, u):
        """(M u)_i for a near-end closure row i < MC: all couplings of the row, b and u indexable 0..NK-1."""
        acc = 0.0
        for j in range(self.MC):
            if j == i:
                continue
            key = (min(i, j), max(i, j))
            acc = acc + self.pair_closure(key[0], key[1], b) * (u[j] - u[i])
        for o in range(1, self.H + 1):          # interior-type couplings that stick out of the block
            j = i + o
            if j >= self.MC:
                acc = acc + self.pair_interior(i, o, b) * (u[j] - u[i])
        return acc
