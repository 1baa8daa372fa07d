"""BASELINE config 2: refinement sweep of the hybridized trace solve on the curved multiblock mesh
meshes/flower_v2.inp (67 straight-sided quadrilateral blocks, 151 faces: 99 locked + 18 jump (side set 7, the
fault) + 12 Dirichlet (side set 1) + 22 Neumann (side set 2) with read_inp_2d's default bc_map).

The reference ships this mesh but no driver for it; this one is written in the style of
square_circle.jl and reuses its machinery (hybridsbp_b200/square_circle.py): corner-blended block maps
(global_curved.jl:66-78), a smooth manufactured solution (continuous, so the jump data on the fault is zero),
Dirichlet / Neumann data and source from it, L2 and fault-traction errors per level.
"""
import os

import numpy as np

from . import host, square_circle as sc


def load_mesh(filename=None):
    if filename is None:
        here = os.path.dirname(os.path.abspath(__file__))
        filename = os.path.join(os.path.dirname(here), "meshes", "flower_v2.inp")
    return host.read_inp_2d(filename)


def block_maps(verts, EToV, EToF, FToB, e):
    """straight-sided block from its corners (transfinite_blend's corner form, global_curved.jl:66-78)"""
    x1, x2, x3, x4 = verts[0, EToV[:, e] - 1]
    y1, y2, y3, y4 = verts[1, EToV[:, e] - 1]
    return (lambda r, s: host.transfinite_blend(x1, x2, x3, x4, r, s),
            lambda r, s: host.transfinite_blend(y1, y2, y3, y4, r, s))


class Smooth:
    """u = sin(a x + 0.3) cos(b y - 0.2) + 0.1 x y  (any block)"""
    a, b = 0.9, 0.7

    @classmethod
    def v(cls, x, y, dom):
        return np.sin(cls.a * x + 0.3) * np.cos(cls.b * y - 0.2) + 0.1 * x * y

    @classmethod
    def vx(cls, x, y, dom):
        return cls.a * np.cos(cls.a * x + 0.3) * np.cos(cls.b * y - 0.2) + 0.1 * y

    @classmethod
    def vy(cls, x, y, dom):
        return -cls.b * np.sin(cls.a * x + 0.3) * np.sin(cls.b * y - 0.2) + 0.1 * x

    @classmethod
    def laplace(cls, x, y, dom):
        return -(cls.a ** 2 + cls.b ** 2) * np.sin(cls.a * x + 0.3) * np.cos(cls.b * y - 0.2)


def solve_level(ctx, mesh, p, N, **kw):
    return sc.solve_level(ctx, mesh, p, N, maps=block_maps, exact=Smooth, **kw)


def main(ctx=None, p=4, N0=17, levels=3):
    import hybridsbp_b200 as hs
    ctx = ctx or hs.Context(0)
    mesh = load_mesh()
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
