"""Synthetic warped multiblock meshes (BASELINE.json configs 4/5).

nbx x nby blocks tile [0, nbx] x [0, nby]; the block-grid coordinates (xi, eta) are warped by

    x = xi  + A sin(2 pi xi / L) sin(2 pi eta / L)
    y = eta - A sin(2 pi xi / L) sin(2 pi eta / L),     L = max(nbx, nby) (or given), A = L / 40

so that every block is genuinely curvilinear (crs != 0).  Connectivity comes out in exactly the
format read_inp_2d produces (built like the hand-made two-block mesh of
global_op_eigenvalues.jl:12-19); left/right outer faces are Dirichlet, bottom/top Neumann, all
interior faces locked interfaces.
"""
import numpy as np

from .host import BC_DIRICHLET, BC_LOCKED_INTERFACE, BC_NEUMANN, _FACE_VERTS


def block_grid_connectivity(nbx, nby, x0=0):
    """(verts, EToV, EToF, FToB) of an nbx x nby grid of blocks; block e = bx + nbx * by."""
    vid = lambda ix, iy: 1 + ix + (nbx + 1) * iy
    ne = nbx * nby
    verts = np.zeros((2, (nbx + 1) * (nby + 1)))
    for iy in range(nby + 1):
        for ix in range(nbx + 1):
            verts[:, vid(ix, iy) - 1] = (x0 + ix, iy)
    EToV = np.zeros((4, ne), dtype=np.int64)
    for by in range(nby):
        for bx in range(nbx):
            EToV[:, bx + nbx * by] = (vid(bx, by), vid(bx + 1, by), vid(bx, by + 1), vid(bx + 1, by + 1))
    EToF = np.zeros((4, ne), dtype=np.int64)
    known = {}
    for e in range(ne):
        for lf in range(4):
            a, b = EToV[_FACE_VERTS[lf], e]
            EToF[lf, e] = known.setdefault((min(a, b), max(a, b)), len(known) + 1)
    FToB = np.full(len(known), BC_LOCKED_INTERFACE, dtype=np.int64)
    for by in range(nby):
        for bx in range(nbx):
            e = bx + nbx * by
            if bx == 0:
                FToB[EToF[0, e] - 1] = BC_DIRICHLET
            if bx == nbx - 1:
                FToB[EToF[1, e] - 1] = BC_DIRICHLET
            if by == 0:
                FToB[EToF[2, e] - 1] = BC_NEUMANN
            if by == nby - 1:
                FToB[EToF[3, e] - 1] = BC_NEUMANN
    return verts, EToV, EToF, FToB


def warp_maps(bx, by, L, A):
    """xf, yf callbacks ((r, s) -> (x, x_r, x_s)) of block (bx, by) for create_metrics."""
    k = 2 * np.pi / L

    def parts(r, s):
        xi = bx + (r + 1) / 2
        et = by + (s + 1) / 2
        sx, cx, se, ce = np.sin(k * xi), np.cos(k * xi), np.sin(k * et), np.cos(k * et)
        return xi, et, A * sx * se, A * k * cx * se / 2, A * k * sx * ce / 2

    def xf(r, s):
        xi, et, w, wr, ws = parts(r, s)
        return xi + w, 0.5 + wr, ws

    def yf(r, s):
        xi, et, w, wr, ws = parts(r, s)
        return et - w, -wr, 0.5 - ws

    return xf, yf


def warped_coefficients(nbx, nby, N, L=None, A=None, bx0=0, dtype=np.float64):
    """crr, css, crs of all blocks of the warped mesh, concatenated block by block (r fastest),
    evaluated for a whole block row at a time.  Block columns are bx0 .. bx0+nbx-1 of a grid whose
    warp period is L (weak-scaling strips of a wider mesh share one L)."""
    L = float(max(nbx, nby)) if L is None else float(L)
    A = L / 40.0 if A is None else float(A)
    k = 2 * np.pi / L
    t = np.linspace(-1.0, 1.0, N + 1)
    npb = (N + 1) ** 2
    crr = np.empty(nbx * nby * npb, dtype=dtype)
    css = np.empty_like(crr)
    crs = np.empty_like(crr)
    bxs = bx0 + np.arange(nbx)
    xi = bxs[:, None] + (t[None, :] + 1) / 2                      # [bx, i]
    sx, cx = np.sin(k * xi), np.cos(k * xi)
    for by in range(nby):
        et = by + (t + 1) / 2                                       # [j]
        se, ce = np.sin(k * et), np.cos(k * et)
        wr = (A * k / 2) * cx[:, None, :] * se[None, :, None]       # [bx, j, i]
        ws = (A * k / 2) * sx[:, None, :] * ce[None, :, None]
        xr, xs_, yr, ys = 0.5 + wr, ws, -wr, 0.5 - ws
        J = xr * ys - xs_ * yr
        if not J.min() > 0:
            raise AssertionError("non-positive Jacobian")
        rx, sx_, ry, sy = ys / J, -yr / J, -xs_ / J, xr / J
        sl = slice(by * nbx * npb, (by + 1) * nbx * npb)
        crr[sl] = (J * (rx * rx + ry * ry)).reshape(-1)
        crs[sl] = (J * (sx_ * rx + sy * ry)).reshape(-1)
        css[sl] = (J * (sx_ * sx_ + sy * sy)).reshape(-1)
    return crr, css, crs


def block_bcs(EToF, FToB):
    """LFToB of every block (4 x ne -> flat 4*ne, block-major) = FToB[EToF[:, e]]."""
    return FToB[np.asarray(EToF) - 1].T.reshape(-1).copy()
