"""Numpy model of the library's distributed trace solve (hybridsbp_b200/csrc/api_cg.cuh, k_cg.cuh), test infrastructure:
it consumes exactly the tables the host layer hands to hsbp_trace_set_partition (parallel.LocalMesh.partition_arrays)
and runs the same algorithm over torch.distributed (gloo) on CPU, with the rank-local operator backed by the oracle:

  * lambda replicated on cut faces; the rank-local part of p.q is reduced before the cut-face exchange completes q;
  * first level: explicit inverses of B_ff = D_f - (own + partner) S_e[f, f];
  * second level: Legendre modes per face, Z^T B Z eliminated rank by rank (I = uncut faces of the rank, G = all cut
    faces): t = A_II^-1 b_I, y = b_G - E^T b_I, ONE all-reduce of [r.z1, r.r, b_I.t, y], c_G = S_G^-1 y, c_I = t - E c_G.
"""
import numpy as np


class RankModel:
    def __init__(self, S, D_own, lstarts, lm, rank, dist=None):
        """S: this rank's side of Fbar^T M^-1 Fbar in the local lambda layout (dense); D_own: Hf * tau of the local sides;
        lstarts: 1-based local FTolambdastarts; lm: parallel.LocalMesh"""
        import torch
        self.torch, self.dist, self.rank = torch, dist, rank
        self.S, self.st = np.asarray(S), np.asarray(lstarts) - 1
        self.n = int(self.st[-1])
        self.lam_faces = [i for i in range(len(self.st) - 1) if self.st[i + 1] > self.st[i]]
        faces, partner, gamma, self.n_gamma = lm.partition_arrays()
        # message order: partners by rank, faces of one partner by gamma (hsbp_trace_set_partition)
        order = sorted(range(len(faces)), key=lambda c: (partner[c], gamma[c]))
        self.cut = [int(faces[c]) - 1 for c in order]
        self.cut_partner = [int(partner[c]) for c in order]
        self.cut_gamma = [int(gamma[c]) for c in order]
        self.owned = {i: bool(lm.owned[i]) for i in self.lam_faces}
        self.unc = [i for i in self.lam_faces if i not in set(self.cut)]
        self.w = np.ones(self.n)
        for i in self.lam_faces:
            if not self.owned[i]:
                self.w[self.sl(i)] = 0.0
        self.D = self._exchange_faces({i: D_own[self.sl(i)] for i in self.cut}, base=D_own.copy(), sign=+1, replace=True)
        self.Binv, self.modes = None, 0

    def sl(self, i):
        return slice(int(self.st[i]), int(self.st[i + 1]))

    # -- communication -------------------------------------------------------------------------------------------
    def _exchange(self, send):
        """send: list of arrays per cut face (message order) -> the partner's arrays in the same order"""
        if self.dist is None or not self.cut:
            return [np.zeros_like(s) for s in send]
        torch, dist = self.torch, self.dist
        ops, bufs = [], []
        peers = sorted(set(self.cut_partner))
        for q in peers:
            idx = [c for c in range(len(self.cut)) if self.cut_partner[c] == q]
            s = torch.from_numpy(np.concatenate([send[c].reshape(-1) for c in idx]))
            r = torch.empty_like(s)
            ops += [dist.P2POp(dist.isend, s, q), dist.P2POp(dist.irecv, r, q)]
            bufs.append((idx, s, r))
        for h in dist.batch_isend_irecv(ops):
            h.wait()
        out = [None] * len(send)
        for idx, s, r in bufs:
            o = 0
            for c in idx:
                out[c] = r.numpy()[o:o + send[c].size].reshape(send[c].shape).copy(); o += send[c].size
        return out

    def _exchange_faces(self, own, base, sign, replace=False):
        send = [np.asarray(own[i]) for i in self.cut]
        recv = self._exchange(send)
        out = base
        for c, i in enumerate(self.cut):
            both = send[c] + recv[c]
            out[self.sl(i)] = both if replace else base[self.sl(i)] + sign * both
        return out

    def _allreduce(self, v):
        if self.dist is None:
            return v
        t = self.torch.from_numpy(np.ascontiguousarray(v, dtype=np.float64))
        self.dist.all_reduce(t)
        return t.numpy()

    # -- setup ------------------------------------------------------------------------------------------------------
    def setup_face_blocks(self):
        own = [self.S[self.sl(i), self.sl(i)] for i in self.cut]
        rec = self._exchange(own)
        part = {i: own[c] + rec[c] for c, i in enumerate(self.cut)}
        self.Binv = {}
        for i in self.lam_faces:
            Sff = part[i] if i in part else self.S[self.sl(i), self.sl(i)]
            self.Binv[i] = np.linalg.inv(np.diag(self.D[self.sl(i)]) - Sff)

    def setup_coarse(self, modes):
        self.modes = q = modes
        order = self.unc + self.cut
        Z = np.zeros((self.n, q * len(order)))
        for j, i in enumerate(order):
            s = np.linspace(-1, 1, self.st[i + 1] - self.st[i])
            for m in range(q):
                Z[self.sl(i), q * j + m] = np.polynomial.legendre.Legendre.basis(m)(s)
        A = -Z.T @ self.S @ Z + Z.T @ ((self.w * self.D)[:, None] * Z)
        nI = q * len(self.unc)
        self.Z, self.nI = Z, nI
        self.AIIinv = np.linalg.inv(A[:nI, :nI]) if nI else np.zeros((0, 0))
        self.E = self.AIIinv @ A[:nI, nI:]
        self.gidx = (np.concatenate([q * g + np.arange(q) for g in self.cut_gamma]) if self.cut else np.zeros(0)).astype(np.int64)
        nG = q * self.n_gamma
        SG = np.zeros((nG, nG))
        SG[np.ix_(self.gidx, self.gidx)] = A[nI:, nI:] - A[nI:, :nI] @ self.E
        SG = self._allreduce(SG)
        self.SGinv = np.linalg.inv(SG) if nG else np.zeros((0, 0))

    # -- operators --------------------------------------------------------------------------------------------------
    def schur_apply(self, x):
        c = self.S @ x
        return self._exchange_faces({i: c[self.sl(i)] for i in self.cut}, base=self.D * x - c * self._uncut_mask(), sign=-1)

    def _uncut_mask(self):
        m = np.ones(self.n)
        for i in self.cut:
            m[self.sl(i)] = 0.0
        return m

    def precond(self, r):
        """-> (z, r.z, r.r) with one all-reduce"""
        z1 = np.zeros(self.n)
        for i in self.lam_faces:
            z1[self.sl(i)] = self.Binv[i] @ r[self.sl(i)] if self.Binv is not None else r[self.sl(i)] / self.D[self.sl(i)]
        wr = self.w * r
        if not self.modes:
            red = self._allreduce(np.array([wr @ z1, wr @ r]))
            return z1, red[0], red[1]
        q = self.modes
        bc = self.Z.T @ wr
        bI, bG = bc[:self.nI], bc[self.nI:]
        t = self.AIIinv @ bI
        y = np.zeros(q * self.n_gamma)
        y[self.gidx] = bG - self.E.T @ bI
        red = self._allreduce(np.concatenate([[wr @ z1, wr @ r, bI @ t], y]))
        y = red[3:]
        cG = self.SGinv @ y
        c = np.concatenate([t - self.E @ cG[self.gidx], cG[self.gidx]])
        return z1 + self.Z @ c, red[0] + red[2] + y @ cG, red[1]

    def solve(self, b, tol=1e-10, maxit=5000):
        lam, r = np.zeros(self.n), b.copy()
        z, rz, b2 = self.precond(r)
        p = z.copy()
        rr, it = b2, 0
        while it < maxit and rr > tol * tol * b2:
            c = self.S @ p
            pq = self._allreduce(np.array([(self.w * self.D * p) @ p - p @ c]))[0]        # before the exchange
            qv = self._exchange_faces({i: c[self.sl(i)] for i in self.cut}, base=self.D * p - c * self._uncut_mask(), sign=-1)
            al = rz / pq
            lam += al * p
            r -= al * qv
            z, rz2, rr = self.precond(r)
            it += 1
            p = z + (rz2 / rz) * p
            rz = rz2
        return lam, it, float(np.sqrt(rr / b2)) if b2 > 0 else 0.0
