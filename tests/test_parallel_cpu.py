"""World-size-2 tests (gloo, CPU) of the multi-GPU path: the host logic in hybridsbp_b200/parallel.py (block partition,
cut-face tables for hsbp_trace_set_partition) drives tests/dist_model.py, a numpy model of the algorithm the library
runs over NCCL (point-to-point exchange of the partial Fbar^T contributions, owner-counted inner products, face-block
and coarse levels eliminated rank by rank).  The rank-local operator is backed by the oracle's sparse matrices; the
distributed solution must equal the single-process oracle solve of the whole mesh."""
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


def local_pieces(G, lm, owner, rank):
    """this rank's side of Fbar^T M^-1 Fbar (dense, local lambda layout), partial D, local starts, row / column maps"""
    cols = np.concatenate([np.arange(G["vstarts"][e] - 1, G["vstarts"][e + 1] - 1) for e in lm.blocks])
    rows = np.concatenate([np.arange(G["starts"][f] - 1, G["starts"][f + 1] - 1) for f in lm.faces]).astype(np.int64)
    FT = G["FbarT"].tocsr()[rows][:, cols].tocsc()
    Mloc = spla.splu(sp.block_diag([G["lops"][e].Mt for e in lm.blocks]).tocsc())
    S = FT @ Mloc.solve(FT.T.toarray())
    lstarts = np.concatenate([[1], 1 + np.cumsum([G["starts"][f + 1] - G["starts"][f] for f in lm.faces])])
    return dict(S=S, D=partial_D(G, owner, rank)[rows], lstarts=lstarts, rows=rows, cols=cols, FT=FT, M=Mloc)


def _worker(rank, world, port, owner, out, face_blocks, coarse):
    import torch.distributed as dist
    from tests.dist_model import RankModel
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        G = build_global()
        FToE, FToLF, EToO, EToS = G["conn"]
        lm = parallel.localize(rank, owner, G["EToF"], G["FToB"], FToE, FToLF, EToO, EToS)
        L = local_pieces(G, lm, owner, rank)
        m = RankModel(L["S"], L["D"], L["lstarts"], lm, rank, dist=dist)
        if face_blocks:
            m.setup_face_blocks()
        if coarse:
            m.setup_coarse(coarse)
        # b = gd - (own + partner) Fbar^T M^-1 g
        own = L["FT"] @ L["M"].solve(G["g"][L["cols"]])
        gd = G["gd"][L["rows"]]
        b = m._exchange_faces({i: own[m.sl(i)] for i in m.cut}, base=gd - own * m._uncut_mask(), sign=-1)
        lam, it, res = m.solve(b, tol=1e-13, maxit=500)
        u = L["M"].solve(G["g"][L["cols"]] - L["FT"].T @ lam)
        np.savez(out % rank, lam=lam, u=u, rows=L["rows"], cols=L["cols"], D=m.D, it=it, res=res, ncut=len(m.cut),
                 ngamma=lm.n_gamma)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("owner,face_blocks,coarse", [([0, 0, 1, 1], False, 0), ([0, 1, 1, 0], False, 0), ([0, 1, 1, 0], True, 0),
                                                      ([0, 0, 1, 1], True, 2), ([0, 1, 1, 0], False, 1), ([0, 1, 0, 1], True, 2)])
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
    cuts, its, lam_by_row = 0, [], {}
    for rank in range(2):
        r = np.load(out % rank)
        assert r["res"] <= 1e-13
        its.append(int(r["it"]))
        assert np.allclose(r["D"], G["D"][r["rows"]], rtol=1e-13)             # completed with the partner's half
        assert np.linalg.norm(r["lam"] - lam_ref[r["rows"]]) <= 1e-10 * np.linalg.norm(lam_ref)
        assert np.linalg.norm(r["u"] - u_ref[r["cols"]]) <= 1e-10 * np.linalg.norm(u_ref)
        cuts += int(r["ncut"])
        assert 2 * int(r["ngamma"]) == 2 * int(r["ncut"])                     # two ranks: every cut face is seen by both
        for row, v in zip(r["rows"], r["lam"]):
            if row in lam_by_row:
                assert lam_by_row[row] == v                                    # replicated copies are bitwise identical
            lam_by_row[row] = v
    assert cuts > 0 and cuts % 2 == 0
    assert its[0] == its[1]
    if face_blocks:                            # exact diagonal blocks, completed across the cut: far fewer iterations
        assert its[0] < 60, its


def test_partition_tables_are_consistent_across_ranks():
    """what hsbp_trace_set_partition receives: both sides of a cut face name each other and share the global index"""
    from hybridsbp_b200 import synthetic
    from hybridsbp_b200.host import connectivityarrays
    world, nbx, nby = 4, 2, 3
    _, EToV, EToF, FToB = synthetic.block_grid_connectivity(nbx * world, nby)
    FToE, FToLF, EToO, EToS = connectivityarrays(EToV, EToF)
    owner = (np.arange(nbx * world * nby) % (nbx * world)) // nbx
    seen = {}
    for rank in range(world):
        lm = parallel.localize(rank, owner, EToF, FToB, FToE, FToLF, EToO, EToS)
        faces, partner, gamma, ng = lm.partition_arrays()
        assert ng == (world - 1) * nby
        assert len(set(gamma)) == len(gamma) and gamma.min() >= 0 and gamma.max() < ng
        for f, q, g in zip(faces, partner, gamma):
            gf = int(lm.faces[f - 1])
            assert (lm.FToE[:, f - 1] > 0).sum() == 1                              # exactly one side is local
            seen.setdefault(gf, []).append((rank, int(q), int(g)))
    assert len(seen) == (world - 1) * nby
    for gf, lst in seen.items():
        (r0, q0, g0), (r1, q1, g1) = lst
        assert (r0, q0) == (q1, r1) and g0 == g1


def test_coarse_space_makes_the_iteration_count_mesh_independent():
    """Single process, oracle-backed model on warped n x n block meshes: face blocks alone need more iterations as the
    mesh grows, face blocks + two Legendre modes per face do not (tools/proto_coarse_space.py)."""
    from hybridsbp_b200 import synthetic
    from hybridsbp_b200.host import connectivityarrays
    from tests.dist_model import RankModel
    from tests.util import warped_metrics
    p, N = 2, 9
    its = {}
    for nb in (3, 6):
        _, EToV, EToF, FToB = synthetic.block_grid_connectivity(nb, nb)
        FToE, FToLF, EToO, EToS = connectivityarrays(EToV, EToF)
        ne = nb * nb
        lops = [orc.locoperator(p, N, N, warped_metrics(p, N, N, e % nb, e // nb, nb, nb), FToB[EToF[:, e] - 1]) for e in range(ne)]
        M, FbarT, D, vstarts, starts = orc.LocalGlobalOperators(lops, [N] * ne, [N] * ne, FToB, FToE, FToLF, EToO, EToS)
        G = dict(lops=lops, conn=(FToE, FToLF, EToO, EToS), starts=starts, vstarts=vstarts, FToB=FToB, EToF=EToF, D=D, FbarT=FbarT)
        owner = np.zeros(ne, dtype=np.int64)
        lm = parallel.localize(0, owner, EToF, FToB, FToE, FToLF, EToO, EToS)
        L = local_pieces(G, lm, owner, 0)
        b = np.random.default_rng(nb).uniform(-1, 1, len(L["rows"]))
        for coarse in (0, 2):
            m = RankModel(L["S"], L["D"], L["lstarts"], lm, 0)
            m.setup_face_blocks()
            if coarse:
                m.setup_coarse(coarse)
            lam, it, res = m.solve(b, tol=1e-10, maxit=2000)
            assert res <= 1e-10
            its[(nb, coarse)] = it
    assert its[(6, 0)] > 1.4 * its[(3, 0)], its              # first level alone: grows with the mesh
    assert its[(6, 2)] < 1.6 * its[(3, 2)], its              # with the coarse space: nearly flat
    assert its[(6, 2)] < 0.5 * its[(6, 0)], its
