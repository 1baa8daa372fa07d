"""BASELINE config 1 on the GPU: the hybridized trace solve of square_circle.jl (56 blocks, curved faces on the
circle, jump interface, Dirichlet + Neumann data) against the oracle's assembled sparse path on identical inputs
(lambda and u within 1e-10 relative, north star), plus the convergence rates of the refinement sweep."""
import numpy as np
import pytest
import scipy.sparse.linalg as spla

from hybridsbp_b200 import host, square_circle as sc
from oracle import hybrid as orc

pytestmark = pytest.mark.gpu


def oracle_level(mesh, p, N, r):
    """the reference's trace method (square_circle.jl:297-388) with the oracle's sparse operators, fed with the
    driver's geometry and data"""
    verts, EToV, EToF, FToB, dom = mesh
    ne = EToV.shape[1]
    FToE, FToLF, EToO, EToS = r["conn"]
    lops = []
    for e in range(ne):
        hm = r["mets"][e]
        om = orc.create_metrics(p, N, N, *sc.block_maps(verts, EToV, EToF, FToB, e))
        assert np.allclose(om.crr, hm.crr, rtol=1e-13) and np.allclose(om.crs, hm.crs, rtol=1e-12, atol=1e-14)
        lops.append(orc.locoperator(p, N, N, om, FToB[EToF[:, e] - 1]))
    Ns = [N] * ne
    M, FbarT, D, vstarts, FTol = orc.LocalGlobalOperators(lops, Ns, Ns, FToB, FToE, FToLF, EToO, EToS)
    assert np.array_equal(FTol, r["FTols"])
    FTod = orc.bcstarts(FToB, FToE, FToLF, orc.BC_JUMP_INTERFACE, Ns, Ns)
    assert np.array_equal(FTod, r["FTods"])
    delta = r["delta"]
    g = np.zeros(vstarts[-1] - 1); gd = np.zeros(FTol[-1] - 1)
    E = sc.ExactSolution
    for e in range(ne):
        bcD = lambda lf, x, y: E.v(x, y, dom[e])
        bcN = lambda lf, x, y, nx, ny: nx * E.vx(x, y, dom[e]) + ny * E.vy(x, y, dom[e])

        def in_jump(lf, x, y):
            f = EToF[lf - 1, e] - 1
            d = delta[FTod[f] - 1:FTod[f + 1] - 1]
            if EToS[lf - 1, e] == 1:
                assert EToO[lf - 1, e]
                return -d
            return d if EToO[lf - 1, e] else d[::-1]
        views = []
        for lf in range(4):
            f = EToF[lf, e] - 1
            sl = gd[FTol[f] - 1:FTol[f + 1] - 1]
            views.append(sl if EToO[lf, e] else sl[::-1])
        ge = g[vstarts[e] - 1:vstarts[e + 1] - 1]
        orc.locbcarray(ge, views, lops[e], FToB[EToF[:, e] - 1], bcD, bcN, in_jump)
        orc.locsourcearray(ge, lambda x, y: -E.laplace(x, y, dom[e]), lops[e])
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
def test_square_circle_level1_matches_oracle(ctx, p):
    mesh = sc.load_mesh(sc.default_mesh_path())
    N = 17
    r = sc.solve_level(ctx, mesh, p, N, tol=1e-13)
    assert r["stats"]["converged"] == 1, r["stats"]
    o = oracle_level(mesh, p, N, r)
    assert np.linalg.norm(r["gd"] - o["gd"]) <= 1e-12 * np.linalg.norm(o["gd"])
    assert np.linalg.norm(r["g_full"] - o["g"]) <= 1e-12 * np.linalg.norm(o["g"])
    assert np.linalg.norm(r["lam"] - o["lam"]) <= 1e-10 * np.linalg.norm(o["lam"]), r["stats"]
    assert np.linalg.norm(r["u"] - o["u"]) <= 1e-10 * np.linalg.norm(o["u"]), r["stats"]


def test_square_circle_convergence_rates(ctx):
    """refinement sweep (square_circle.jl:204-428) at p = 4: the L2 error converges at about order 4"""
    mesh = sc.load_mesh(sc.default_mesh_path())
    eps, teps = [], []
    for N in (17, 34):
        r = sc.solve_level(ctx, mesh, 4, N, tol=1e-12)
        assert r["stats"]["converged"] == 1
        eps.append(r["eps"]); teps.append(r["tau_eps"])
    rate = np.log2(eps[0] / eps[1])
    trate = np.log2(teps[0] / teps[1])
    assert 3.3 < rate < 5.5, (eps, rate)
    assert trate > 2.0, (teps, trate)
