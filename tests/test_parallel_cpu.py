"""World-size-2 test (gloo, CPU) of the multi-GPU host logic in hybridsbp_b200/parallel.py: block partition,
cut-face bookkeeping, point-to-point exchange of the partial Fbar^T contributions, masked inner products.
The local operator is backed by the oracle's sparse matrices here (the GPU one wraps the C-ABI); the distributed
solution must equal the single-process oracle solve of the whole mesh."""
import os
import socket

import numpy as np
import pytest
import scipy.sparse as sp
import scipy.sparse.linalg as spla

from hybridsbp_b200 import parallel
from oracle import hybrid as orc
from tests.util import random_spd_metrics


def mesh_four_blocks_with_flips():
    """2 x 2 blocks, two of them rotated so that interface faces meet with opposite orientation"""
    v = lambda ix, iy: 1 + ix + 3 * iy
    blocks = [(v(0, 0), v(1, 0), v(0, 1), v(1, 1)), (v(2, 1), v(2, 0), v(1, 1), v(1, 0)),
              (v(1, 2), v(0, 2), v(1, 1), v(0, 1)), (v(1, 1), v(2, 1), v(1, 2), v(2, 2))]
    EToV = np.array(blocks).T
    fv = np.array([[0, 2], [1, 3], [0, 1], [2, 3]])
    EToF = np.zeros((4, 4), dtype=np.int64)
    known = {}
    for e in range(4):
        for lf in range(4):
            a, b = EToV[fv[lf], e]
            EToF[lf, e] = known.setdefault((min(a, b), max(a, b)), len(known) + 1)
    count = np.bincount(EToF.reshape(-1) - 1, minlength=len(known))
    FToB = np.where(count == 2, orc.BC_LOCKED_INTERFACE, orc.BC_DIRICHLET).astype(np.int64)
    FToB[np.where(count == 1)[0][::2]] = orc.BC_NEUMANN
    FToB[np.where(count == 2)[0][0]] = orc.BC_JUMP_INTERFACE
    return EToV, EToF, FToB


class OracleLocalOperator:
    """rank-local pieces of the oracle's global sparse operators (restriction of Fbar^T to the local blocks'
    columns = this rank's side of every face)"""

    def __init__(self, lm, lops_all, FbarT_all, vstarts_all, starts_all, Dpart):
        import torch
        self.torch = torch
        cols = np.concatenate([np.arange(vstarts_all[e] - 1, vstarts_all[e + 1] - 1) for e in lm.blocks])
        rows = np.concatenate([np.arange(starts_all[f] - 1, starts_all[f + 1] - 1) for f in lm.faces])
        self.FT = FbarT_all.tocsr()[rows][:, cols].tocsc()
        self.M = spla.splu(sp.block_diag([lops_all[e].Mt for e in lm.blocks]).tocsc())
        self.rows, self.cols = rows, cols
        self.lNp = len(rows)
        self._D = Dpart[rows].copy()

    def get_D(self):
        return self._D.copy()

    def set_D(self, D):
        self._D = np.asarray(D).copy()

    def schur_apply(self, lam):
        l = lam.numpy()
        return self.torch.from_numpy(self._D * l - self.FT @ self.M.solve(self.FT.T @ l))

    def rhs(self, g, gd):
        return self.torch.from_numpy(gd.numpy() - self.FT @ self.M.solve(g.numpy()))

    def back_substitute(self, g, lam):
        return self.torch.from_numpy(self.M.solve(g.numpy() - self.FT.T @ lam.numpy()))


def build_global(p=4, seed=3):
    EToV, EToF, FToB = mesh_four_blocks_with_flips()
    conn = orc.connectivityarrays(EToV, EToF)
    FToE, FToLF, EToO, EToS = conn
    assert (~EToO).any()
    rng = np.random.default_rng(seed)
    N = 3 * p - 1
    ne = EToV.shape[1]
    lops = [orc.locoperator(p, N, N, random_spd_metrics(p, N, N, rng, scale2=0.2), FToB[EToF[:, e] - 1]) for e in range(ne)]
    Ns = [N] * ne
    M, FbarT, D, vstarts, starts = orc.LocalGlobalOperators(lops, Ns, Ns, FToB, FToE, FToLF, EToO, EToS)
    g = rng.uniform(-1, 1, vstarts[-1] - 1)
    gd = rng.uniform(-1, 1, starts[-1] - 1)
    return dict(EToF=EToF, FToB=FToB, conn=conn, lops=lops, M=M, FbarT=FbarT, D=D, vstarts=vstarts, starts=starts,
                g=g, gd=gd, N=N, ne=ne)


def partial_D(G, owner, rank):
    """Hf * tau of the sides that live on `rank`, in the global lambda layout and minus-side orientation"""
    FToE, FToLF, EToO, EToS = G["conn"]
    Dp = np.zeros(G["starts"][-1] - 1)
    for f in range(len(G["FToB"])):
        if G["starts"][f + 1] == G["starts"][f]:
            continue
        sl = slice(G["starts"][f] - 1, G["starts"][f + 1] - 1)
        for side in range(2):
            e, k = FToE[side, f] - 1, FToLF[side, f] - 1
            if owner[e] != rank:
                continue
            t = G["lops"][e].Hf[k].diagonal() * G["lops"][e].tau[k].diagonal()
            Dp[sl] += t if EToO[k, e] else t[::-1]
    return Dp


class OracleFaceBlocks:
    """CPU stand-in for the C-ABI's face-block preconditioner (hsbp_trace_precond_cut_own / _setup_cut / _apply) on top of
    OracleLocalOperator: exercises parallel.setup_face_block_preconditioner over gloo."""

    def __init__(self, op, lstarts):
        self.op, self.st = op, np.asarray(lstarts) - 1
        X = op.M.solve(op.FT.T.toarray())                       # M^-1 Fbar (local columns)
        self.S = op.FT @ X                                      # this rank's side of Fbar^T M^-1 Fbar
        self.blocks = None

    def _own(self, f):
        a, b = self.st[f], self.st[f + 1]
        return self.S[a:b, a:b]

    def precond_cut_own(self, ids, out):
        out.t[:] = out.t.new_tensor(np.concatenate([self._own(f - 1).reshape(-1, order="F") for f in ids]))

    def precond_setup(self, kind):
        self.precond_setup_cut([], None)

    def precond_setup_cut(self, ids, partner):
        part, o = {}, 0
        for f in ids:
            nl = self.st[f] - self.st[f - 1]
            part[f - 1] = partner.t.numpy()[o:o + nl * nl].reshape(nl, nl, order="F"); o += nl * nl
        D = self.op.get_D()
        self.blocks = []
        for f in range(len(self.st) - 1):
            a, b = self.st[f], self.st[f + 1]
            if b > a:
                Bff = np.diag(D[a:b]) - (self._own(f) + part[f] if f in part else self._own(f))
                self.blocks.append((a, b, Bff))

    def apply(self, r):
        z = np.zeros(len(r))
        for a, b, Bff in self.blocks:
            z[a:b] = np.linalg.solve(Bff, r.numpy()[a:b])
        return r.new_tensor(z)


def _worker(rank, world, port, owner, out, face_blocks=False, coarse=0):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        G = build_global()
        FToE, FToLF, EToO, EToS = G["conn"]
        lm = parallel.localize(rank, owner, G["EToF"], G["FToB"], FToE, FToLF, EToO, EToS)
        op = OracleLocalOperator(lm, G["lops"], G["FbarT"], G["vstarts"], G["starts"], partial_D(G, owner, rank))
        lstarts = np.concatenate([[1], 1 + np.cumsum([G["starts"][f + 1] - G["starts"][f] for f in lm.faces])])
        dt = parallel.DistributedTrace(op, lstarts, lm, dist=dist)
        if face_blocks:
            fb = OracleFaceBlocks(op, lstarts)
            parallel.setup_face_block_preconditioner(fb, lm, lstarts, dist, "cpu")
            op.has_precond, op.precond = True, fb.apply
        if coarse:
            dt.setup_coarse_space(coarse)
            cs = dt.coarse                                     # coloured probes == column-by-column Z^T B Z, on every rank
            Ac = cs["chol"] @ cs["chol"].T
            for j in (0, cs["nc"] - 1):
                c = torch.zeros(cs["nc"], dtype=torch.float64); c[j] = 1.0
                col = dt._restrict(cs, dt.schur_apply(dt._prolong(cs, c)))
                assert torch.allclose(Ac[:, j], col, rtol=1e-9, atol=1e-11 * float(Ac.abs().max()))
        lam, u, st = dt.solve(torch.from_numpy(G["g"][op.cols]), torch.from_numpy(G["gd"][op.rows]), tol=1e-13, maxit=500)
        np.savez(out % rank, lam=lam.numpy(), u=u.numpy(), rows=op.rows, cols=op.cols, D=dt.D.numpy(),
                 it=st["outer_iterations"], conv=st["converged"], ncut=sum(len(v) for v in lm.cut.values()))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("owner,face_blocks,coarse", [([0, 0, 1, 1], False, 0), ([0, 1, 1, 0], False, 0), ([0, 1, 1, 0], True, 0),
                                                      ([0, 0, 1, 1], True, 2), ([0, 1, 1, 0], False, 1)])
def test_two_rank_trace_solve_equals_single_process(tmp_path, owner, face_blocks, coarse):
    import torch.multiprocessing as mp
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    out = str(tmp_path / "rank%d.npz")
    mp.spawn(_worker, args=(2, port, np.array(owner), out, face_blocks, coarse), nprocs=2, join=True)
    G = build_global()
    B = orc.assemblelambdamatrix(G["starts"], G["vstarts"], G["EToF"], G["FToB"], G["M"].F, G["D"], G["FbarT"])
    bl = np.zeros(G["starts"][-1] - 1); uu = np.zeros(G["vstarts"][-1] - 1)
    orc.LocalToGLobalRHS(bl, G["g"], G["gd"], uu, G["M"].F, G["FbarT"], G["vstarts"])
    lam_ref = np.linalg.solve(B.toarray(), bl)
    rhs = G["g"] - G["FbarT"].T @ lam_ref
    u_ref = np.concatenate([G["M"].F[e].solve(rhs[G["vstarts"][e] - 1:G["vstarts"][e + 1] - 1]) for e in range(G["ne"])])
    cuts, its = 0, []
    for rank in range(2):
        r = np.load(out % rank)
        assert r["conv"] == 1
        its.append(int(r["it"]))
        assert np.allclose(r["D"], G["D"][r["rows"]], rtol=1e-13)             # completed with the partner's half
        assert np.linalg.norm(r["lam"] - lam_ref[r["rows"]]) <= 1e-10 * np.linalg.norm(lam_ref)
        assert np.linalg.norm(r["u"] - u_ref[r["cols"]]) <= 1e-10 * np.linalg.norm(u_ref)
        cuts += int(r["ncut"])
    assert cuts > 0 and cuts % 2 == 0          # both ranks see the same cut faces
    assert its[0] == its[1]
    if face_blocks:                            # exact diagonal blocks, completed across the cut: far fewer iterations
        assert its[0] < 60, its


def test_coarse_space_makes_the_iteration_count_mesh_independent():
    """Single process, oracle-backed operator on warped n x n block meshes: face blocks alone need more iterations as the
    mesh grows, face blocks + two Legendre modes per face do not (tools/proto_coarse_space.py)."""
    import torch
    from hybridsbp_b200 import synthetic
    from hybridsbp_b200.host import connectivityarrays
    from tests.util import warped_metrics
    p, N = 2, 9
    its = {}
    for nb in (3, 6):
        _, EToV, EToF, FToB = synthetic.block_grid_connectivity(nb, nb)
        FToE, FToLF, EToO, EToS = connectivityarrays(EToV, EToF)
        ne = nb * nb
        lops = [orc.locoperator(p, N, N, warped_metrics(p, N, N, e % nb, e // nb, nb, nb), FToB[EToF[:, e] - 1]) for e in range(ne)]
        M, FbarT, D, vstarts, starts = orc.LocalGlobalOperators(lops, [N] * ne, [N] * ne, FToB, FToE, FToLF, EToO, EToS)
        G = dict(lops=lops, conn=(FToE, FToLF, EToO, EToS), starts=starts, FToB=FToB, EToF=EToF, D=D)
        owner = np.zeros(ne, dtype=np.int64)
        lm = parallel.localize(0, owner, EToF, FToB, FToE, FToLF, EToO, EToS)
        rng = np.random.default_rng(nb)
        g, gd = rng.uniform(-1, 1, vstarts[-1] - 1), rng.uniform(-1, 1, starts[-1] - 1)
        for coarse in (0, 2):
            op = OracleLocalOperator(lm, lops, FbarT, vstarts, starts, partial_D(G, owner, 0))
            lstarts = np.concatenate([[1], 1 + np.cumsum([starts[f + 1] - starts[f] for f in lm.faces])])
            dt = parallel.DistributedTrace(op, lstarts, lm, dist=None)
            fb = OracleFaceBlocks(op, lstarts)
            parallel.setup_face_block_preconditioner(fb, lm, lstarts, None, "cpu")
            if fb.blocks is None:                                 # single process: no cut faces, plain setup
                fb.precond_setup_cut([], None)
            op.has_precond, op.precond = True, fb.apply
            if coarse:
                dt.setup_coarse_space(coarse)
                # the coloured probes reproduce Z^T B Z column by column
                cs = dt.coarse
                assert cs["matvecs"] < cs["nc"]
                Ac = cs["chol"] @ cs["chol"].T
                for j in (0, 1, cs["nc"] // 2, cs["nc"] - 1):
                    c = torch.zeros(cs["nc"], dtype=torch.float64); c[j] = 1.0
                    col = dt._restrict(cs, dt.schur_apply(dt._prolong(cs, c)))
                    assert torch.allclose(Ac[:, j], col, rtol=1e-9, atol=1e-11 * float(Ac.abs().max()))
            lam, u, st = dt.solve(torch.from_numpy(g[op.cols]), torch.from_numpy(gd[op.rows]), tol=1e-10, maxit=2000)
            assert st["converged"] == 1
            its[(nb, coarse)] = st["outer_iterations"]
    assert its[(6, 0)] > 1.4 * its[(3, 0)], its              # first level alone: grows with the mesh
    assert its[(6, 2)] < 1.6 * its[(3, 2)], its              # with the coarse space: nearly flat (27 -> 80 vs 23 -> 31)
    assert its[(6, 2)] < 0.5 * its[(6, 0)], its
