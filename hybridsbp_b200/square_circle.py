"""Host-side mirror of the reference driver square_circle.jl (BASELINE config 1): hybridized Poisson solve on
the 56-block square-with-circle mesh, curved block faces on the circle, manufactured solution with a jump
across the circle, refinement sweep with L2 / traction error.

What runs where
  host (this file)   mesh reading, vertex snapping (square_circle.jl:27-33), edge maps and transfinite blending
                     (:221-285), exact solution and data (:98-201), the small per-face bookkeeping of the right-hand
                     sides (:321-366), error norms (:396-428)
  device (libhsbp)   everything the reference spends its time on: the block operators, g = -sum F_k v_k, the trace
                     solve lambda = B^-1 (g_delta - Fbar^T M^-1 g), u = M^-1 (g - Fbar lambda), traction operators

`geometry()` and `ExactSolution` are also what the parity test feeds to the oracle, so both sides see identical inputs.
"""
import os

import numpy as np

from . import host
from .blocks import Blocks, Trace, LOCAL_BAND, LOCAL_CHOLESKY, LOCAL_PCG  # noqa: F401

BC_MAP = [host.BC_DIRICHLET, host.BC_DIRICHLET, host.BC_NEUMANN, host.BC_NEUMANN, host.BC_JUMP_INTERFACE]   # :11-12


def load_mesh(filename):
    verts, EToV, EToF, FToB, EToDomain = host.read_inp_2d(filename, BC_MAP)
    for v in range(verts.shape[1]):                      # :27-33  pull near-circle vertices onto r = 1
        x, y = verts[:, v]
        if abs(np.hypot(x, y) - 1) < 1e-5:
            q = np.arctan2(y, x)
            verts[:, v] = np.cos(q), np.sin(q)
    return verts, EToV, EToF, FToB, EToDomain


def block_maps(verts, EToV, EToF, FToB, e):
    """(xt, yt) callbacks of block e (0-based) for create_metrics, square_circle.jl:221-283."""
    x1, x2, x3, x4 = verts[0, EToV[:, e] - 1]
    y1, y2, y3, y4 = verts[1, EToV[:, e] - 1]
    lin = lambda a, b: (lambda al: a * (1 - al) / 2 + b * (1 + al) / 2)
    dlin = lambda a, b: (lambda al: -a / 2 + b / 2 + 0 * al)
    ex = [lin(x1, x3), lin(x2, x4), lin(x1, x2), lin(x3, x4)]
    exa = [dlin(x1, x3), dlin(x2, x4), dlin(x1, x2), dlin(x3, x4)]
    ey = [lin(y1, y3), lin(y2, y4), lin(y1, y2), lin(y3, y4)]
    eya = [dlin(y1, y3), dlin(y2, y4), dlin(y1, y2), dlin(y3, y4)]
    fb = FToB[EToF[:, e] - 1]
    if fb[0] == host.BC_JUMP_INTERFACE or fb[1] == host.BC_JUMP_INTERFACE:
        raise NotImplementedError("curved face 1 / 2 not implemented yet")         # :251-256

    def arc(Qa, Qb):
        beta = (Qb - Qa) / 2
        ang = lambda al: Qa * (1 - al) / 2 + Qb * (1 + al) / 2
        return (lambda al: np.cos(ang(al)), lambda al: np.sin(ang(al)),
                lambda al: -beta * np.sin(ang(al)), lambda al: beta * np.cos(ang(al)))
    if fb[2] == host.BC_JUMP_INTERFACE:                                            # :257-268
        Q1, Q2 = np.arctan2(y1, x1), np.arctan2(y2, x2)
        if not (-np.pi / 2 < Q1 - Q2 < np.pi / 2):
            Q2 -= np.sign(Q2) * 2 * np.pi
        ex[2], ey[2], exa[2], eya[2] = arc(Q1, Q2)
    if fb[3] == host.BC_JUMP_INTERFACE:                                            # :269-280
        Q3, Q4 = np.arctan2(y3, x3), np.arctan2(y4, x4)
        if not (-np.pi / 2 < Q3 - Q4 < np.pi / 2):
            raise NotImplementedError("curved face 4 angle correction not implemented yet")
        ex[3], ey[3], exa[3], eya[3] = arc(Q3, Q4)
    xt = lambda r, s: host.transfinite_blend(ex[0], ex[1], ex[2], ex[3], exa[0], exa[1], exa[2], exa[3], r, s)
    yt = lambda r, s: host.transfinite_blend(ey[0], ey[1], ey[2], ey[3], eya[0], eya[1], eya[2], eya[3], r, s)
    return xt, yt


class ExactSolution:
    """square_circle.jl:98-201; `dom` = EToDomain of the block (1 inside the circle, 2 outside)."""
    c = np.e / (1 + np.e)

    @classmethod
    def v(cls, x, y, dom):
        r, th = np.hypot(x, y), np.arctan2(y, x)
        if dom == 1:
            return cls.c * (1 - np.exp(-r ** 2)) * r * np.sin(th)
        return (r - 1) ** 2 * np.cos(th) + (r - 1) * np.sin(th)

    @classmethod
    def _polar(cls, x, y, dom):
        r, th = np.hypot(x, y), np.arctan2(y, x)
        if dom == 1:
            dv_dr = cls.c * (2 * r ** 2 * np.exp(-r ** 2) + 1 - np.exp(-r ** 2)) * np.sin(th)
            dv_dth = cls.c * (1 - np.exp(-r ** 2)) * r * np.cos(th)
        else:
            dv_dr = 2 * (r - 1) * np.cos(th) + np.sin(th)
            dv_dth = -(r - 1) ** 2 * np.sin(th) + (r - 1) * np.cos(th)
        return r, th, dv_dr, dv_dth

    @classmethod
    def vx(cls, x, y, dom):
        r, th, dv_dr, dv_dth = cls._polar(x, y, dom)
        return dv_dr * np.cos(th) + dv_dth * (-np.sin(th) / r)

    @classmethod
    def vy(cls, x, y, dom):
        r, th, dv_dr, dv_dth = cls._polar(x, y, dom)
        return dv_dr * np.sin(th) + dv_dth * (np.cos(th) / r)

    @classmethod
    def laplace(cls, x, y, dom):
        r, th = np.hypot(x, y), np.arctan2(y, x)
        if dom == 1:                                      # :158-162
            u_r = cls.c * (2 * r ** 2 * np.exp(-r ** 2) + 1 - np.exp(-r ** 2)) * np.sin(th)
            u_rr = cls.c * np.exp(-r ** 2) * (6 * r - 4 * r ** 3) * np.sin(th)
            return u_rr + (1 / r) * u_r - (cls.c / r ** 2) * (1 - np.exp(-r ** 2)) * r * np.sin(th)
        return (2 * np.cos(th) + (1 / r) * (2 * (r - 1) * np.cos(th) + np.sin(th)) +
                (1 / r ** 2) * (-(r - 1) ** 2 * np.cos(th) - (r - 1) * np.sin(th)))


def geometry(mesh, p, N, maps=None):
    """metrics of every block at level size N (:285)."""
    verts, EToV, EToF, FToB, _ = mesh
    maps = maps or block_maps
    out = []
    for e in range(EToV.shape[1]):
        xt, yt = maps(verts, EToV, EToF, FToB, e)
        out.append(host.create_metrics(p, N, N, xt, yt))
    return out


def jump_data(mesh, conn, mets, FTods, N, exact=None, slip=None):
    """delta on the jump faces: exact value on the plus side minus the minus side (:321-330), or a given slip(x, y) evaluated at
    the minus side's face points."""
    verts, EToV, EToF, FToB, dom = mesh
    FToE, FToLF, EToO, EToS = conn
    exact = exact or ExactSolution
    delta = np.zeros(FTods[-1] - 1)
    for f in range(len(FToB)):
        if FToB[f] >= host.BC_JUMP_INTERFACE:
            e1, e2 = FToE[:, f] - 1
            lf1 = FToLF[0, f] - 1
            xf, yf = mets[e1].facecoord[0][lf1], mets[e1].facecoord[1][lf1]
            delta[FTods[f] - 1:FTods[f + 1] - 1] = slip(xf, yf) if slip else exact.v(xf, yf, dom[e2]) - exact.v(xf, yf, dom[e1])
    return delta


def face_data(mesh, conn, mets, taus, FTols, FTods, delta, N, p, exact=None):
    """Boundary / jump data of every block face and g_delta, exactly what locbcarray! feeds to F_k (:335-363,
    global_curved.jl:596-623).  taus[e][lf]: penalty vectors, returns (v[e][lf] or None, g_delta)."""
    verts, EToV, EToF, FToB, dom = mesh
    FToE, FToLF, EToO, EToS = conn
    ne = EToV.shape[1]
    exact = exact or ExactSolution
    gd = np.zeros(FTols[-1] - 1)
    v = [[None] * 4 for _ in range(ne)]
    for e in range(ne):
        m = mets[e]
        for lf in range(4):
            f = EToF[lf, e] - 1
            bc = FToB[f]
            xf, yf = m.facecoord[0][lf], m.facecoord[1][lf]
            if bc == host.BC_DIRICHLET:
                v[e][lf] = exact.v(xf, yf, dom[e])
            elif bc == host.BC_NEUMANN:
                gN = m.nx[lf] * exact.vx(xf, yf, dom[e]) + m.ny[lf] * exact.vy(xf, yf, dom[e])
                v[e][lf] = m.sJ[lf] * gN / taus[e][lf]
            elif bc >= host.BC_JUMP_INTERFACE:
                d = delta[FTods[f] - 1:FTods[f + 1] - 1]
                if EToS[lf, e] == 1:
                    assert EToO[lf, e]
                    dj = -d
                else:
                    dj = d if EToO[lf, e] else d[::-1]
                vf = dj / 2
                v[e][lf] = vf
                Hf = host.norm_weights(p, N)
                contrib = Hf * taus[e][lf] * vf
                sl = slice(FTols[f] - 1, FTols[f + 1] - 1)
                if EToO[lf, e]:
                    gd[sl] -= contrib
                else:
                    gd[sl] -= contrib[::-1]
    return v, gd


def solve_level(ctx, mesh, p, N, local_mode=None, tol=1e-12, maxit=5000, maps=None, exact=None, jump_code=None,
                condense=True, coarse_modes=2, slip=None):
    """One refinement level on the GPU.  Returns dict(eps, tau_eps, lam, u, stats, ...)."""
    verts, EToV, EToF, FToB, dom = mesh
    ne, nf = EToV.shape[1], len(FToB)
    conn = host.connectivityarrays(EToV, EToF)
    FToE, FToLF, EToO, EToS = conn
    exact = exact or ExactSolution
    jump_code = host.BC_JUMP_INTERFACE if jump_code is None else jump_code
    mets = geometry(mesh, p, N, maps)
    fl = lambda a: np.asarray(a).reshape(-1, order="F")
    blk = Blocks(ctx, p, [N] * ne, [N] * ne)
    blk.set_metrics(np.concatenate([fl(m.crr) for m in mets]), np.concatenate([fl(m.css) for m in mets]),
                    np.concatenate([fl(m.crs) for m in mets]))
    bcs = np.array([[FToB[f - 1] for f in EToF[:, e]] for e in range(ne)], dtype=np.int64)
    blk.set_bc(bcs.reshape(-1))
    blk.compute_tau(2.0)
    if local_mode is None:
        # the reference factorises every block (cholesky(M-tilde), square_circle.jl:299): dense factors for small
        # blocks, banded factors beyond
        local_mode = LOCAL_CHOLESKY if (N + 1) ** 2 <= 2500 else LOCAL_BAND
    blk.local_setup(local_mode, tol=1e-14, maxit=200000)
    tr = Trace(blk, FToB, FToE, FToLF, EToO, EToS)
    if condense:
        # ... and forms B from per-block products (assembleλmatrix, :313): here the dense S_e stay block-wise and B is
        # applied inside a CG preconditioned with its exact diagonal blocks
        tr.condense()
        tr.precond_setup(1)
        if coarse_modes > 0:               # second level: Legendre modes per face (iteration counts independent of the mesh)
            tr.coarse_setup(coarse_modes)
    FTols = tr.FTolambdastarts
    jump_codes = tuple(sorted(set(int(b) for b in FToB if b >= host.BC_JUMP_INTERFACE))) or (jump_code,)
    FTods = host.bcstarts(FToB, FToE, FToLF, jump_codes, [N] * ne, [N] * ne)
    tau = blk.get_tau()
    taus = [[tau[blk.face_slice(e, lf + 1)] for lf in range(4)] for e in range(ne)]
    delta = jump_data(mesh, conn, mets, FTods, N, exact, slip)
    v, gd = face_data(mesh, conn, mets, taus, FTols, FTods, delta, N, p, exact)
    Hw = host.norm_weights(p, N)
    JH = [(m.J * Hw[:, None] * Hw[None, :]).reshape(-1, order="F") for m in mets]      # global_curved.jl:491
    # g = - sum_k F_k v_k  (device)  +  JH * source  (host, elementwise)
    vface = np.zeros(blk.FNp)
    for e in range(ne):
        for lf in range(4):
            if v[e][lf] is not None:
                vface[blk.face_slice(e, lf + 1)] = v[e][lf]
    g = np.zeros(blk.VNp)
    for e in range(ne):
        x, y = mets[e].coord
        g[blk.vol_slice(e)] = JH[e] * (-exact.laplace(fl(x), fl(y), dom[e]))      # :364-365
    dg = ctx.array(g)
    dv = ctx.array(vface)
    blk.face_F_add(dv, -1.0, dg)
    dgd, dlam, du = ctx.array(gd), ctx.empty(tr.lNp), ctx.empty(blk.VNp)
    stats = tr.solve(dg, dgd, dlam, du, tol=tol, maxit=maxit)
    u, lam = du.get(), dlam.get()
    # errors (:396-422)
    eps2 = 0.0
    for e in range(ne):
        x, y = mets[e].coord
        d = u[blk.vol_slice(e)] - exact.v(fl(x), fl(y), dom[e])
        eps2 += d @ (JH[e] * d)
    dtr = ctx.empty(blk.FNp)
    blk.face_traction(du, dtr)
    trv = dtr.get()
    teps2 = 0.0
    for f in range(nf):
        if FToB[f] >= host.BC_JUMP_INTERFACE:
            e1, lf1 = FToE[0, f] - 1, FToLF[0, f] - 1
            m = mets[e1]
            xf, yf = m.facecoord[0][lf1], m.facecoord[1][lf1]
            tex = m.nx[lf1] * exact.vx(xf, yf, dom[e1]) + m.ny[lf1] * exact.vy(xf, yf, dom[e1])
            lamf = lam[FTols[f] - 1:FTols[f + 1] - 1]
            df = delta[FTods[f] - 1:FTods[f + 1] - 1]
            t = (trv[blk.face_slice(e1, lf1 + 1)] + taus[e1][lf1] * (lamf - df / 2)) / m.sJ[lf1]     # computetraction
            dt = t - tex
            teps2 += dt @ (Hw * m.sJ[lf1] * dt)
    out = dict(eps=np.sqrt(eps2), tau_eps=np.sqrt(teps2), lam=lam, u=u, stats=stats, g=g, gd=gd, delta=delta,
               FTols=FTols, FTods=FTods, mets=mets, conn=conn, vstarts=blk.vstarts, g_full=dg.get())
    tr.close(); blk.close()
    return out


def default_mesh_path():
    here = os.path.dirname(os.path.abspath(__file__))
    return os.path.join(os.path.dirname(here), "meshes", "square_circle.inp")


def main(ctx=None, p=4, N0=17, levels=3):
    """The refinement sweep of square_circle.jl:204-428 (SBPp and the number of levels are arguments here)."""
    import hybridsbp_b200 as hs
    ctx = ctx or hs.Context(0)
    mesh = load_mesh(default_mesh_path())
    eps, teps = [], []
    for lvl in range(levels):
        r = solve_level(ctx, mesh, p, N0 * 2 ** lvl)
        eps.append(r["eps"]); teps.append(r["tau_eps"])
        print("(lvl, eps, tau_eps) =", (lvl + 1, r["eps"], r["tau_eps"]), r["stats"])
    eps, teps = np.array(eps), np.array(teps)
    print((np.log(eps[:-1]) - np.log(eps[1:])) / np.log(2))
    print((np.log(teps[:-1]) - np.log(teps[1:])) / np.log(2))
    return eps, teps


if __name__ == "__main__":
    main()
