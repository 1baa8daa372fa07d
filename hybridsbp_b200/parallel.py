"""Multi-GPU trace solve: the mesh's blocks are partitioned across the GPUs of one node (one process per GPU,
torch.distributed / NCCL for the plumbing), SURVEY.md section 8e.

Blocks are independent in M-tilde u and in the local solves; coupling is only through faces, each shared by two
blocks (global_curved.jl:525-554).  A face whose two blocks live on different ranks (a *cut face*) carries lambda on
both ranks, kept identical: every rank computes its own side's contribution to Fbar^T z (the C-ABI does that when
the remote side's FToE entry is 0) and the partner's contribution -- Nf doubles per cut face, all faces of one
partner packed into one message -- is exchanged with point-to-point send / recv and added.  Inner products count a
cut face once (on the rank of its minus side) and are completed by an all-reduce.  Volume vectors never move.

The numerics are exactly those of the single-device solve (hsbp_trace_solve): Jacobi-preconditioned CG on
B = D - Fbar^T M^-1 Fbar, then u = M^-1 (g - Fbar lambda) (square_circle.jl:376-388).

Everything in this file is host logic on torch tensors; the local operator is an object with
    lNp, get_D() / set_D(D), rhs(g, gd) -> b, schur_apply(lam) -> B lam, back_substitute(g, lam) -> u
(GpuLocalOperator below wraps the C-ABI; the CPU tests plug in an oracle-backed one and run over gloo).
"""
import ctypes as C
from dataclasses import dataclass
from typing import Dict, List

import numpy as np

from . import host


# ---- partitioning ----------------------------------------------------------------------------------
def partition_contiguous(nblocks, world):
    """owner[e] for contiguous, equally sized ranges of blocks."""
    per = -(-nblocks // world)
    return np.minimum(np.arange(nblocks) // per, world - 1).astype(np.int64)


@dataclass
class LocalMesh:
    blocks: np.ndarray          # global ids (0-based) of the local blocks, increasing
    faces: np.ndarray           # global ids (0-based) of the faces touched by local blocks, increasing
    EToF: np.ndarray            # 4 x nlocal, local face ids, 1-based
    FToB: np.ndarray
    FToE: np.ndarray            # 2 x nfaces_local, local block ids 1-based, 0 = on another rank
    FToLF: np.ndarray
    EToO: np.ndarray
    EToS: np.ndarray
    cut: Dict[int, List[int]]   # partner rank -> local face ids (0-based) of the cut faces, by increasing global id
    owned: np.ndarray           # per local face: this rank counts it in inner products


def localize(rank, owner, EToF, FToB, FToE, FToLF, EToO, EToS):
    """Connectivity of one rank's blocks in local numbering (reference conventions, 1-based ids in arrays)."""
    owner = np.asarray(owner)
    blocks = np.where(owner == rank)[0]
    g2l = -np.ones(len(owner), dtype=np.int64)
    g2l[blocks] = np.arange(len(blocks))
    faces = np.unique(EToF[:, blocks]) - 1
    f2l = {int(f): i for i, f in enumerate(faces)}
    lEToF = np.vectorize(lambda f: f2l[int(f) - 1] + 1)(EToF[:, blocks]).astype(np.int64)
    lFToE = np.zeros((2, len(faces)), dtype=np.int64)
    cut: Dict[int, List[int]] = {}
    owned = np.ones(len(faces), dtype=bool)
    has_lambda = lambda f: FToB[f] == host.BC_LOCKED_INTERFACE or FToB[f] >= host.BC_JUMP_INTERFACE
    for i, f in enumerate(faces):
        for side in range(2):
            e = FToE[side, f] - 1
            if e >= 0 and owner[e] == rank:
                lFToE[side, i] = g2l[e] + 1
        if has_lambda(f):
            em, ep = FToE[0, f] - 1, FToE[1, f] - 1
            if owner[em] != owner[ep]:
                partner = int(owner[ep] if owner[em] == rank else owner[em])
                cut.setdefault(partner, []).append(i)
                owned[i] = owner[em] == rank                 # the minus side's rank counts the face
    return LocalMesh(blocks, faces, lEToF, FToB[faces].copy(), lFToE, FToLF[:, faces].copy(),
                     EToO[:, blocks].copy(), EToS[:, blocks].copy(), cut, owned)


# ---- the distributed solve --------------------------------------------------------------------------
class DistributedTrace:
    """CG on the trace system of a partitioned mesh.  `op` is this rank's local operator, `starts` its 1-based
    FTolambdastarts over the local faces, `lm` the LocalMesh.  `dist` is torch.distributed (already initialised)
    or None for a single process."""

    def __init__(self, op, starts, lm: LocalMesh, dist=None, device="cpu"):
        import torch
        self.torch = torch
        self.op, self.lm, self.dist, self.device = op, lm, dist, device
        self.n = int(starts[-1] - 1)
        rng = lambda i: np.arange(starts[i] - 1, starts[i + 1] - 1)
        self.cut_idx = {q: torch.as_tensor(np.concatenate([rng(i) for i in fl]) if fl else np.zeros(0, np.int64),
                                           device=device) for q, fl in sorted(lm.cut.items())}
        w = np.ones(self.n)
        for i in range(len(lm.faces)):
            if not lm.owned[i]:
                w[rng(i)] = 0.0
        self.w = torch.as_tensor(w, device=device)
        # D = Hf (tau_minus + tau_plus): complete the cut faces with the partner's half (global_curved.jl:556-557)
        D = torch.as_tensor(op.get_D(), device=device)
        D = self._exchange_add(D, None)
        op.set_D(D.cpu().numpy())
        self.D = D
        self.messages = sum(len(v) for v in self.cut_idx.values())
        self.starts = np.asarray(starts)
        self.coarse = None

    # ---- optional second level: a few polynomial modes per face (tools/proto_coarse_space.py) ----------------
    def setup_coarse_space(self, modes=2):
        """Additive coarse correction  Z (Z^T B Z)^-1 Z^T  on top of the first-level preconditioner: Z holds the Legendre
        modes 0 .. modes-1 of every face that carries lambda.  With two modes the CG iteration count no longer grows with
        the number of blocks across the mesh (prototype: 12 x 12 blocks 171 -> 37).  The coarse matrix is global and
        replicated; it is probed with coloured coarse vectors (distributed matvec + all-reduce).  Faces must have one
        size.  Call after the first-level preconditioner is set up."""
        torch, dist = self.torch, self.dist
        st = self.starts
        lam_faces = [i for i in range(len(self.lm.faces)) if st[i + 1] > st[i]]
        nls = {int(st[i + 1] - st[i]) for i in lam_faces}
        if len(nls) != 1:
            raise ValueError("the coarse space needs faces of one size")
        nl = nls.pop()
        # global numbering of the faces that carry lambda (compressed global face ids)
        gid = np.asarray(self.lm.faces)[lam_faces]
        nfg = torch.tensor([int(np.max(self.lm.faces)) + 1], device=self.device)
        if dist is not None:
            dist.all_reduce(nfg, op=dist.ReduceOp.MAX)
        mark = torch.zeros(int(nfg.item()), dtype=torch.float64, device=self.device)
        mark[torch.as_tensor(gid, device=self.device)] = 1.0
        if dist is not None:
            dist.all_reduce(mark, op=dist.ReduceOp.MAX)
        number = (torch.cumsum(mark, 0) - 1).long()
        nfc = int(mark.sum().item())
        cidx = number[torch.as_tensor(gid, device=self.device)]                     # coarse face index of every local face
        s = np.linspace(-1.0, 1.0, nl)
        Lq = np.stack([np.polynomial.legendre.Legendre.basis(k)(s) for k in range(modes)], axis=1)     # nl x modes
        Lq = torch.as_tensor(Lq, device=self.device)
        rows = torch.as_tensor(np.concatenate([np.arange(st[i] - 1, st[i + 1] - 1) for i in lam_faces]), device=self.device)
        nc = nfc * modes
        state = dict(nl=nl, modes=modes, Lq=Lq, rows=rows, cidx=cidx, nc=nc, nfl=len(lam_faces))
        self.coarse = None
        # A_c = Z^T B Z.  B couples a face only to the faces that share a block with it, so many columns come out of one
        # matvec: faces of one colour (pairwise without a common neighbour face) are probed together and their responses
        # are separated by support.  The face graph of the whole mesh is gathered on every rank (4 ids per block), the
        # greedy colouring is deterministic, and the restricted responses are all-reduced: every rank fills the same A_c.
        A = torch.zeros(nc, nc, dtype=torch.float64, device=self.device)
        gl = np.asarray(self.lm.faces)[np.asarray(self.lm.EToF) - 1]                    # global face ids, 4 x local blocks
        if dist is not None:
            parts = [None] * dist.get_world_size()
            dist.all_gather_object(parts, gl)
            gl = np.concatenate(parts, axis=1)
        number_np = number.cpu().numpy()
        mark_np = mark.cpu().numpy() > 0
        adj = [set() for _ in range(nfc)]
        for e in range(gl.shape[1]):
            fs = [int(number_np[f]) for f in gl[:, e] if mark_np[f]]
            for a_ in fs:
                adj[a_].update(fs)
        reach = [set().union(*[adj[g] for g in adj[f]]) if adj[f] else {f} for f in range(nfc)]
        colour = -np.ones(nfc, dtype=np.int64)
        for f in range(nfc):
            used = {colour[g] for g in reach[f] if colour[g] >= 0}
            c = 0
            while c in used:
                c += 1
            colour[f] = c
        for c in range(int(colour.max()) + 1):
            members = np.where(colour == c)[0]
            resp_rows, resp_cols = [], []                                               # coarse face g answers to probe face f
            for f in members:
                for g in adj[f]:
                    resp_rows.append(g); resp_cols.append(f)
            rr = torch.as_tensor(np.asarray(resp_rows, dtype=np.int64), device=self.device)
            rc = torch.as_tensor(np.asarray(resp_cols, dtype=np.int64), device=self.device)
            mem = torch.as_tensor(members, device=self.device)
            for k in range(modes):
                cv = torch.zeros(nfc, modes, dtype=torch.float64, device=self.device)
                cv[mem, k] = 1.0
                out = self._restrict(state, self.schur_apply(self._prolong(state, cv.reshape(-1)))).view(nfc, modes)
                for m in range(modes):
                    A[rr * modes + m, rc * modes + k] = out[rr, m]
        state["matvecs"] = (int(colour.max()) + 1) * modes
        A = 0.5 * (A + A.T)
        state["chol"] = torch.linalg.cholesky(A)
        self.coarse = state

    def _prolong(self, cs, c):
        """lam = Z c (every rank fills its own faces, cut faces on both ranks)"""
        torch = self.torch
        cf = c.view(-1, cs["modes"])[cs["cidx"]]                                      # local faces x modes
        lam = torch.zeros(self.n, dtype=torch.float64, device=self.device)
        lam[cs["rows"]] = (cf @ cs["Lq"].T).reshape(-1)
        return lam

    def _restrict(self, cs, r):
        """c = Z^T r over the whole mesh (a cut face is counted by the rank that owns it)"""
        torch = self.torch
        rf = (r * self.w)[cs["rows"]].view(cs["nfl"], cs["nl"])
        c = torch.zeros(cs["nc"] // cs["modes"], cs["modes"], dtype=torch.float64, device=self.device)
        c.index_add_(0, cs["cidx"], rf @ cs["Lq"])
        c = c.reshape(-1)
        if self.dist is not None:
            self.dist.all_reduce(c)
        return c

    def _exchange_add(self, x, base):
        """On every cut face x_f = base_f + (c_mine + c_partner) with c = x_f - base_f, the part only one rank can
        compute (base: what both ranks already hold, e.g. D lam).  The sum of the two parts is commutative, so the
        replicated entries stay bitwise identical on both ranks."""
        if self.dist is None or not self.cut_idx:
            return x
        torch, dist = self.torch, self.dist
        send, recv, ops = {}, {}, []
        for q, idx in self.cut_idx.items():
            s = x[idx] if base is None else x[idx] - base[idx]
            send[q] = s.contiguous()
            recv[q] = torch.empty_like(send[q])
            ops.append(dist.P2POp(dist.isend, send[q], q))
            ops.append(dist.P2POp(dist.irecv, recv[q], q))
        for r in dist.batch_isend_irecv(ops):
            r.wait()
        for q, idx in self.cut_idx.items():
            both = send[q] + recv[q]
            x[idx] = both if base is None else base[idx] + both
        return x

    def dots(self, pairs):
        """weighted inner products, one all-reduce for all of them"""
        torch = self.torch
        v = torch.stack([(a * b * self.w).sum() for a, b in pairs])
        if self.dist is not None:
            self.dist.all_reduce(v)
        return [float(t) for t in v.cpu()]

    def schur_apply(self, lam):
        q = self.op.schur_apply(lam)                      # D lam - (local side of Fbar^T M^-1 Fbar lam)
        return self._exchange_add(q, self.D * lam)

    def rhs(self, g, gd):
        b = self.op.rhs(g, gd)                            # gd - (local side of Fbar^T M^-1 g); gd is replicated
        return self._exchange_add(b, gd)

    def precond(self, r):
        """z = P^-1 r: the local operator's preconditioner (face blocks; cut faces use the completed D on every rank that
        holds them, so the copies of lambda stay identical) or Jacobi with the completed D."""
        z = self.op.precond(r) if getattr(self.op, "has_precond", False) else r / self.D
        if self.coarse is not None:
            cs = self.coarse
            c = self.torch.cholesky_solve(self._restrict(cs, r).unsqueeze(1), cs["chol"]).squeeze(1)
            z = z + self._prolong(cs, c)
        return z

    def solve(self, g, gd, tol=1e-10, maxit=10000):
        """-> (lambda, u, stats); same iteration as hsbp_trace_solve."""
        torch = self.torch
        r = self.rhs(g, gd)
        lam = torch.zeros_like(r)
        p = self.precond(r)
        rz, b2 = self.dots([(r, p), (r, r)])
        it, rr, converged = 0, b2, b2 == 0.0
        while not converged and it < maxit:
            q = self.schur_apply(p)
            (pq,) = self.dots([(p, q)])
            alpha = rz / pq
            lam += alpha * p
            r -= alpha * q
            z = self.precond(r)
            rz_new, rr = self.dots([(r, z), (r, r)])
            it += 1
            if np.sqrt(rr / b2) <= tol:
                converged = True
                break
            p = z + (rz_new / rz) * p
            rz = rz_new
        u = self.op.back_substitute(g, lam)
        return lam, u, dict(outer_iterations=it, converged=int(converged),
                            rel_residual=float(np.sqrt(rr / b2)) if b2 > 0 else 0.0)


def setup_face_block_preconditioner(tr, lm, starts, dist, device):
    """Face-block preconditioner on a partitioned mesh: the diagonal block B_ff of a cut face needs the S_e[f, f] of both
    ranks.  Every rank packs its own blocks per partner (cut faces in increasing global id, the order both ranks share),
    exchanges them point to point and builds D_f - (own + partner): the sum is commutative, so both ranks factorise the
    same matrix and their copies of lambda stay identical."""
    import torch
    if dist is None or not lm.cut:
        tr.precond_setup(1)
        return
    starts = np.asarray(starts)
    sync = (lambda: torch.cuda.synchronize(device)) if torch.device(device).type == "cuda" else (lambda: None)
    ops, parts = [], []
    for q, faces in sorted(lm.cut.items()):
        ids = np.asarray(faces, dtype=np.int64) + 1
        n = int(sum((starts[f + 1] - starts[f]) ** 2 for f in faces))
        own = torch.empty(n, dtype=torch.float64, device=device)
        tr.precond_cut_own(ids, _Ptr(own))
        rec = torch.empty_like(own)
        ops += [dist.P2POp(dist.isend, own, q), dist.P2POp(dist.irecv, rec, q)]
        parts.append((ids, own, rec))
    sync()
    for r in dist.batch_isend_irecv(ops):
        r.wait()
    sync()
    ids = np.concatenate([p[0] for p in parts])
    partner = torch.cat([p[2] for p in parts])
    sync()
    tr.precond_setup_cut(ids, _Ptr(partner))


# ---- local operator over the C-ABI (device pointers of torch tensors) ---------------------------------
class _Ptr:
    def __init__(self, t):
        self.t = t
        self.ptr = C.c_void_p(t.data_ptr())


class GpuLocalOperator:
    """This rank's blocks on its GPU: hybridsbp_b200.Blocks + Trace driven through torch-owned device vectors."""

    def __init__(self, blocks, trace):
        import torch
        self.torch, self.blk, self.tr = torch, blocks, trace
        self.lNp = trace.lNp
        self.dev = torch.device("cuda", blocks.ctx.device)
        self._w = torch.empty(blocks.VNp, dtype=torch.float64, device=self.dev)

    def _sync(self):
        self.torch.cuda.current_stream(self.dev).synchronize()      # torch's stream -> library's stream ordering
        self.blk.ctx.sync()

    def get_D(self):
        return self.tr.D()

    def set_D(self, D):
        self.tr.set_D(D)

    def schur_apply(self, lam):
        out = self.torch.empty_like(lam)
        self._sync()
        self.tr.schur_apply(_Ptr(lam), _Ptr(out))
        self.blk.ctx.sync()
        return out

    has_precond = False

    def precond(self, r):
        out = self.torch.empty_like(r)
        self._sync()
        self.tr.precond_apply(_Ptr(r), _Ptr(out))
        self.blk.ctx.sync()
        return out

    def rhs(self, g, gd):
        b = self.torch.empty_like(gd)
        self._sync()
        self.tr.rhs(_Ptr(g), _Ptr(gd), _Ptr(b))
        self.blk.ctx.sync()
        return b

    def back_substitute(self, g, lam):
        self._w.copy_(g)
        u = self.torch.empty_like(g)
        self._sync()
        self.tr.Fbar_add(_Ptr(lam), -1.0, _Ptr(self._w))
        self.blk.local_solve(_Ptr(self._w), _Ptr(u))
        self.blk.ctx.sync()
        return u
