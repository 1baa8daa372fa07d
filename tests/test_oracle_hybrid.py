"""Pins the oracle's restatement of the block operator and trace assembly (oracle/hybrid.py) with
the reference's own identities and fixtures: the Schur-complement identity of
global_op_eigenvalues.jl:84, B ~ B' (global_curved.jl:794), positive definiteness of M-tilde under
random SPD coefficient tensors (local_op_eigenvalues.jl:32-50), exactness on constants
(seas/BP1/single_block.jl) and the mesh reader on the reference's .inp files."""
import os

import numpy as np
import pytest
import scipy.sparse as sp

from oracle import hybrid as orc
from tests.util import random_spd_metrics

HERE = os.path.dirname(os.path.abspath(__file__))
MESH = os.path.join(os.path.dirname(HERE), "meshes")


@pytest.mark.parametrize("p", [2, 4, 6])
@pytest.mark.parametrize("bcs", [(1, 1, 1, 1), (1, 2, 2, 2)])
def test_local_operator_is_spd_under_random_tensors(p, bcs):
    """local_op_eigenvalues.jl:41-50 (Dirichlet and D/N/N/N, tauscale = 1 there; 2 is the default)."""
    N = 3 * p - 1 if p > 2 else 5
    rng = np.random.default_rng(777 + p)
    for _ in range(6):
        m = random_spd_metrics(p, N, N, rng)
        lop = orc.locoperator(p, N, N, m, bcs, tauscale=2.0)
        Mt = lop.Mt.toarray()
        assert np.abs(Mt - Mt.T).max() < 1e-11 * np.abs(Mt).max()
        assert np.linalg.eigvalsh((Mt + Mt.T) / 2).min() > 0


def two_block_mesh():
    """hand-built connectivity of global_op_eigenvalues.jl:12-19: two blocks side by side."""
    EToV = np.array([[1, 2], [2, 3], [4, 5], [5, 6]], dtype=np.int64)
    EToF = np.array([[1, 2], [2, 3], [4, 5], [6, 7]], dtype=np.int64)
    FToB = np.array([1, 0, 1, 1, 1, 1, 1], dtype=np.int64)
    return EToV, EToF, FToB


@pytest.mark.parametrize("p", [2, 4, 6])
def test_schur_complement_identity(p):
    """B = D - Fbar^T A11^-1 Fbar (global_op_eigenvalues.jl:84) and B ~ B' (global_curved.jl:794)."""
    EToV, EToF, FToB = two_block_mesh()
    FToE, FToLF, EToO, EToS = orc.connectivityarrays(EToV, EToF)
    N = 3 * p - 1 if p > 2 else 5
    rng = np.random.default_rng(p)
    Nr = [N, N]; Ns = [N, N]
    lop = []
    for e in range(2):
        m = random_spd_metrics(p, N, N, rng)
        lop.append(orc.locoperator(p, N, N, m, FToB[EToF[:, e] - 1], tauscale=1.0))
    M, FbarT, D, vstarts, FTol = orc.LocalGlobalOperators(lop, Nr, Ns, FToB, FToE, FToLF, EToO, EToS)
    B = orc.assemblelambdamatrix(FTol, vstarts, EToF, FToB, M.F, D, FbarT).toarray()
    A11 = sp.block_diag([l.Mt for l in lop]).toarray()
    Bref = np.diag(D) - FbarT.toarray() @ np.linalg.solve(A11, FbarT.toarray().T)
    assert np.abs(B - Bref).max() < 1e-9 * np.abs(Bref).max()
    assert np.abs(B - B.T).max() < 1e-9 * np.abs(B).max()
    assert np.linalg.eigvalsh((B + B.T) / 2).min() > 0


def test_connectivity_orientation_and_sides():
    EToV, EToF, FToB = two_block_mesh()
    FToE, FToLF, EToO, EToS = orc.connectivityarrays(EToV, EToF)
    assert FToE[:, 1].tolist() == [1, 2] and FToLF[:, 1].tolist() == [2, 1]
    assert EToS[1, 0] == 1 and EToS[0, 1] == 2 and EToO[0, 1]
    # reversing the second block's face vertices flips the orientation flag
    EToV2 = EToV.copy(); EToV2[:, 1] = [5, 6, 2, 3]
    _, _, EToO2, _ = orc.connectivityarrays(EToV2, EToF)
    assert not EToO2[0, 1]


def test_single_block_reproduces_constants():
    """seas/BP1/single_block.jl: exact solution 1 with D/D/N/N boundary data; M-tilde \\ g is all ones."""
    p, N = 2, 24
    el = 1e13
    Lx = Ly = 80.0
    xt = lambda r, s: (Lx / 2 * (1 + np.tan(np.arctan(1 / el) * 0 + r) * 0 + r), np.full_like(r, Lx / 2), np.zeros_like(r))
    yt = lambda r, s: (Ly / 2 * (1 + s), np.zeros_like(s), np.full_like(s, Ly / 2))
    m = orc.create_metrics(p, N, N, xt, yt)
    bcs = (1, 1, 2, 2)
    lop = orc.locoperator(p, N, N, m, bcs)
    g = np.zeros((N + 1) ** 2)
    orc.locbcarray_mod(g, lop, bcs, lambda lf, x, y: np.ones_like(x), lambda lf, x, y, nx, ny: np.zeros_like(x))
    u = orc.default_factorization(lop.Mt).solve(g)
    assert np.abs(u - 1.0).max() < 1e-9


MESHES = {
    # name: (bc_map or None, nverts, nelems, nfaces, {bc: count})
    "square_circle.inp": ([1, 1, 2, 2, 7], 73, 56, 128, {0: 88, 7: 8, 1: 16, 2: 16}),   # bc_map of square_circle.jl:11-12
    "flower_v2.inp": (None, 85, 67, 151, {0: 99, 7: 18, 1: 12, 2: 22}),
    "1_1_block.inp": (None, 4, 1, 4, None),
    "BP1_v1.inp": (None, 215, 194, 408, None),
}


@pytest.mark.parametrize("name", sorted(MESHES))
def test_read_inp_2d_counts(name):
    """read_inp_2d (global_curved.jl:802-956) on copies of the reference's mesh files (meshes/)."""
    bc_map, nv, ne, nf, counts = MESHES[name]
    verts, EToV, EToF, FToB, EToBlock = orc.read_inp_2d(os.path.join(MESH, name), bc_map)
    assert verts.shape == (2, nv) and EToV.shape == (4, ne) and EToF.shape == (4, ne) and FToB.shape == (nf,)
    assert EToF.min() == 1 and EToF.max() == nf
    if counts is not None:
        got = {int(k): int((FToB == k).sum()) for k in np.unique(FToB)}
        assert got == counts
    # every face belongs to one or two elements and the connectivity arrays build
    FToE, FToLF, EToO, EToS = orc.connectivityarrays(EToV, EToF)
    assert (FToE[0] > 0).all()
    interior = FToE[1] > 0
    assert np.all((FToB[~interior] != 0) | (FToB[~interior] >= 0))


def test_rateandstate_and_bracketed_newton():
    """global_curved.jl:1031-1075: the root satisfies g(V) = 0 inside the bracket; no bracket -> NaN, iter < 0."""
    a, V0, sn, eta = 0.015, 1e-6, 50.0, 32.04 / (2 * 3.464)
    psi, tau = 0.6, 31.0
    f = lambda V: orc.rateandstate(V, psi, sn, tau, eta, a, V0)
    V, fv, it = orc.newtbndv(f, -tau / eta, tau / eta, 0.0, ftol=1e-9, atolx=1e-9, rtolx=1e-9)
    assert it > 0 and abs(f(V)[0]) < 1e-9 and -tau / eta <= V <= tau / eta
    h = 1e-7 * abs(V)
    assert abs((f(V + h)[0] - f(V - h)[0]) / (2 * h) - f(V)[1]) < 1e-5 * f(V)[1]
    Vn, fn, itn = orc.newtbndv(f, 2 * tau / eta, 3 * tau / eta, 0.0)
    assert np.isnan(Vn) and itn < 0
