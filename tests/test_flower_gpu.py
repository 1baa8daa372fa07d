"""BASELINE config 2 on the GPU: hybridized trace solve on meshes/flower_v2.inp (67 blocks, reversed-orientation
interfaces, jump / Dirichlet / Neumann faces), p = 4 and 6, against the oracle's assembled sparse path on identical
inputs (lambda, u within 1e-10), plus the refinement sweep's convergence rate."""
import numpy as np
import pytest
import scipy.sparse.linalg as spla

from hybridsbp_b200 import flower, host
from oracle import hybrid as orc

pytestmark = pytest.mark.gpu


def oracle_level(mesh, p, N, r):
    verts, EToV, EToF, FToB, dom = mesh
    ne = EToV.shape[1]
    FToE, FToLF, EToO, EToS = r["conn"]
    E = flower.Smooth
    lops = [orc.locoperator(p, N, N, orc.create_metrics(p, N, N, *flower.block_maps(verts, EToV, EToF, FToB, e)),
                            FToB[EToF[:, e] - 1]) for e in range(ne)]
    Ns = [N] * ne
    M, FbarT, D, vstarts, FTol = orc.LocalGlobalOperators(lops, Ns, Ns, FToB, FToE, FToLF, EToO, EToS)
    assert np.array_equal(FTol, r["FTols"])
    g = np.zeros(vstarts[-1] - 1); gd = np.zeros(FTol[-1] - 1)
    for e in range(ne):
        bcD = lambda lf, x, y: E.v(x, y, 0)
        bcN = lambda lf, x, y, nx, ny: nx * E.vx(x, y, 0) + ny * E.vy(x, y, 0)
        in_jump = lambda lf, x, y: np.zeros_like(x)            # continuous exact solution: no slip on the fault
        views = []
        for lf in range(4):
            f = EToF[lf, e] - 1
            sl = gd[FTol[f] - 1:FTol[f + 1] - 1]
            views.append(sl if EToO[lf, e] else sl[::-1])
        ge = g[vstarts[e] - 1:vstarts[e + 1] - 1]
        orc.locbcarray(ge, views, lops[e], FToB[EToF[:, e] - 1], bcD, bcN, in_jump)
        orc.locsourcearray(ge, lambda x, y: -E.laplace(x, y, 0), lops[e])
    B = orc.assemblelambdamatrix(FTol, vstarts, EToF, FToB, M.F, D, FbarT)
    bl = np.zeros(FTol[-1] - 1); u = np.zeros(vstarts[-1] - 1)
    orc.LocalToGLobalRHS(bl, g, gd, u, M.F, FbarT, vstarts)
    lam = spla.spsolve(B.tocsc(), bl)
    rhs = g - FbarT.T @ lam
    for e in range(ne):
        sl = slice(vstarts[e] - 1, vstarts[e + 1] - 1)
        u[sl] = M.F[e].solve(rhs[sl])
    return dict(g=g, gd=gd, lam=lam, u=u)


@pytest.mark.parametrize("p", [4, 6])
def test_flower_level1_matches_oracle(ctx, p):
    mesh = flower.load_mesh()
    _, EToV, EToF, FToB, _ = mesh
    _, _, EToO, _ = host.connectivityarrays(EToV, EToF)
    assert (~EToO).any(), "flower_v2 has interfaces that meet with opposite orientation"
    N = 17
    r = flower.solve_level(ctx, mesh, p, N, tol=1e-13)
    assert r["stats"]["converged"] == 1, r["stats"]
    o = oracle_level(mesh, p, N, r)
    assert np.linalg.norm(r["g_full"] - o["g"]) <= 1e-12 * np.linalg.norm(o["g"])
    assert np.linalg.norm(r["lam"] - o["lam"]) <= 1e-10 * np.linalg.norm(o["lam"]), r["stats"]
    assert np.linalg.norm(r["u"] - o["u"]) <= 1e-10 * np.linalg.norm(o["u"]), r["stats"]


@pytest.mark.parametrize("p,lo", [(4, 3.3), (6, 4.0)])
def test_flower_convergence(ctx, p, lo):
    mesh = flower.load_mesh()
    eps = [flower.solve_level(ctx, mesh, p, N, tol=1e-13)["eps"] for N in (17, 34)]
    rate = np.log2(eps[0] / eps[1])
    assert rate > lo, (eps, rate)
