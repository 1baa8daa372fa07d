"""The trace method of square_circle.jl (:204-428) assembled from the ORACLE's functions, with the product's host-side mesh
loader, block maps and exact solution: what the executed reference driver (drivers.run_square_circle) and the golden vectors
generated from it are compared with on the CPU.  TEST INFRASTRUCTURE."""
import numpy as np
import scipy.sparse.linalg as spla

from hybridsbp_b200 import square_circle as sc
from oracle import hybrid as orc


def oracle_square_circle_level(p, N, mesh=None, maps=None, exact=None, slip=None):
    """maps / exact: block maps and manufactured solution (default: square_circle's); slip(x, y): given slip on the jump faces instead
    of the jump of the exact solution"""
    mesh = mesh or sc.load_mesh(sc.default_mesh_path())
    verts, EToV, EToF, FToB, dom = mesh
    maps = maps or sc.block_maps
    ne = EToV.shape[1]
    conn = orc.connectivityarrays(EToV, EToF)
    FToE, FToLF, EToO, EToS = conn
    lops = []
    for e in range(ne):
        om = orc.create_metrics(p, N, N, *maps(verts, EToV, EToF, FToB, e))
        lops.append(orc.locoperator(p, N, N, om, FToB[EToF[:, e] - 1]))
    Ns = [N] * ne
    M, FbarT, D, vstarts, FTol = orc.LocalGlobalOperators(lops, Ns, Ns, FToB, FToE, FToLF, EToO, EToS)
    jump_codes = tuple(sorted(set(int(b) for b in FToB if b >= orc.BC_JUMP_INTERFACE))) or (orc.BC_JUMP_INTERFACE,)
    FTod = orc.bcstarts(FToB, FToE, FToLF, jump_codes, Ns, Ns)
    E = exact or sc.ExactSolution
    delta = np.zeros(FTod[-1] - 1)
    for f in range(len(FToB)):                                                    # square_circle.jl:321-330
        if FToB[f] >= orc.BC_JUMP_INTERFACE:
            e1, e2 = FToE[:, f] - 1
            lf1 = FToLF[0, f] - 1
            xf, yf = lops[e1].facecoord[0][lf1], lops[e1].facecoord[1][lf1]
            delta[FTod[f] - 1:FTod[f + 1] - 1] = slip(xf, yf) if slip else E.v(xf, yf, dom[e2]) - E.v(xf, yf, dom[e1])
    g = np.zeros(vstarts[-1] - 1); gd = np.zeros(FTol[-1] - 1)
    for e in range(ne):                                                           # :332-364
        bcD = lambda lf, x, y: E.v(x, y, dom[e])
        bcN = lambda lf, x, y, nx, ny: nx * E.vx(x, y, dom[e]) + ny * E.vy(x, y, dom[e])

        def in_jump(lf, x, y):
            f = EToF[lf - 1, e] - 1
            d = delta[FTod[f] - 1:FTod[f + 1] - 1]
            if EToS[lf - 1, e] == 1:
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
    orc.LocalToGLobalRHS(bl, g, gd, u, M.F, FbarT, vstarts)                       # :372
    lam = spla.spsolve(B.tocsc(), bl)                                             # :373
    rhs = g - FbarT.T @ lam                                                       # :375-376
    for e in range(ne):
        sl = slice(vstarts[e] - 1, vstarts[e + 1] - 1)
        u[sl] = M.F[e].solve(rhs[sl])
    eps = 0.0                                                                     # :393-400
    for e in range(ne):
        x, y = lops[e].coord
        sl = slice(vstarts[e] - 1, vstarts[e + 1] - 1)
        dlt = u[sl] - E.v(x.reshape(-1, order="F"), y.reshape(-1, order="F"), dom[e])
        eps += dlt @ (lops[e].JH @ dlt)
    teps = 0.0                                                                    # :402-420
    tauf = np.zeros(FTod[-1] - 1)
    for f in range(len(FToB)):
        if FToB[f] >= orc.BC_JUMP_INTERFACE:
            e1 = FToE[0, f] - 1; lf1 = FToLF[0, f] - 1
            xf, yf = lops[e1].facecoord[0][lf1], lops[e1].facecoord[1][lf1]
            nx, ny = lops[e1].nx[lf1], lops[e1].ny[lf1]
            tex = nx * E.vx(xf, yf, dom[e1]) + ny * E.vy(xf, yf, dom[e1])
            tr = orc.computetraction(lops[e1], lf1 + 1, u[vstarts[e1] - 1:vstarts[e1 + 1] - 1],
                                     lam[FTol[f] - 1:FTol[f + 1] - 1], delta[FTod[f] - 1:FTod[f + 1] - 1])
            tauf[FTod[f] - 1:FTod[f + 1] - 1] = tr
            dt = tr - tex
            teps += dt @ (lops[e1].Hf[lf1].diagonal() * lops[e1].sJ[lf1] * dt)
    return dict(mesh=mesh, conn=conn, lops=lops, FbarT=FbarT, D=D, vstarts=vstarts, FTol=FTol, FTod=FTod, delta=delta, g=g, gd=gd,
                B=B, bl=bl, lam=lam, u=u, tauf=tauf, eps=float(np.sqrt(eps)), teps=float(np.sqrt(teps)))
