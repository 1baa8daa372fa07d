"""The product's host-side mirror of the reference's setup functions (hybridsbp_b200/host.py: read_inp_2d,
connectivityarrays, transfinite_blend, create_metrics, bcstarts) against the oracle's line-faithful restatement
(oracle/hybrid.py; global_curved.jl:19-209, 714-728, 802-956) on copies of all four meshes the reference ships.
The two were written independently (different parsers, different blend algebra): integer outputs must be identical,
floating-point outputs agree to rounding."""
import os

import numpy as np
import pytest

from hybridsbp_b200 import host
from oracle import hybrid as orc

MESH = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "meshes")
MESHES = {"square_circle.inp": [1, 1, 2, 2, 7], "flower_v2.inp": None, "1_1_block.inp": None, "BP1_v1.inp": None}


@pytest.mark.parametrize("name", sorted(MESHES))
def test_reader_and_connectivity_are_identical(name):
    bc_map = MESHES[name]
    h = host.read_inp_2d(os.path.join(MESH, name), bc_map)
    o = orc.read_inp_2d(os.path.join(MESH, name), bc_map)
    for a, b, what in zip(h, o, ("verts", "EToV", "EToF", "FToB", "EToBlock")):
        a, b = np.asarray(a), np.asarray(b)
        assert a.shape == b.shape, what
        if what == "verts":
            assert np.array_equal(a, b), what                 # the same decimal strings parsed by the same float()
        else:
            assert np.array_equal(a.astype(np.int64), b.astype(np.int64)), what
    hc = host.connectivityarrays(h[1], h[2])
    oc = orc.connectivityarrays(o[1], o[2])
    for a, b, what in zip(hc, oc, ("FToE", "FToLF", "EToO", "EToS")):
        assert np.array_equal(np.asarray(a), np.asarray(b)), what
    # offsets of the faces of one kind (bcstarts, global_curved.jl:714-728) for every code that occurs
    ne = h[1].shape[1]
    Nr = Ns = np.full(ne, 17)
    for code in np.unique(h[3]):
        assert np.array_equal(host.bcstarts(h[3], hc[0], hc[1], int(code), Nr, Ns), orc.bcstarts(o[3], oc[0], oc[1], int(code), Nr, Ns))


@pytest.mark.parametrize("name", sorted(MESHES))
def test_blend_and_metrics_agree_on_every_block(name):
    """straight-sided blocks from the mesh's corner vertices (the corner form, global_curved.jl:66-78) and, on the same
    blocks, curved edges through the general form (:19-51): an arc-like bulge on every edge with analytic derivatives"""
    bc_map = MESHES[name]
    verts, EToV, EToF, FToB, _ = orc.read_inp_2d(os.path.join(MESH, name), bc_map)
    p, N = 4, 13
    r, s = host.reference_grid(N, N)
    ro, so = orc.create_metrics(p, N, N).coord                       # the oracle's reference grid
    assert np.array_equal(r, ro) and np.array_equal(s, so)
    worst = 0.0
    for e in range(EToV.shape[1]):
        vx, vy = verts[0, EToV[:, e] - 1], verts[1, EToV[:, e] - 1]
        for v in (vx, vy):
            a = host.transfinite_blend(v[0], v[1], v[2], v[3], r, s)
            b = orc.transfinite_blend_corners(v[0], v[1], v[2], v[3], r, s)
            for x, y in zip(a, b):
                worst = max(worst, np.max(np.abs(np.asarray(x) - np.asarray(y))) / max(1.0, np.max(np.abs(y))))
        # curved edges: edge k of coordinate v bulges by amp * (1 - t^2) (keeps the corners)
        amp = 0.02 * (1 + e % 3)
        def edges(v):
            lin = lambda a, b: (lambda t: a * (1 - t) / 2 + b * (1 + t) / 2 + amp * (1 - t * t))
            der = lambda a, b: (lambda t: (b - a) / 2 - 2 * amp * t + 0 * t)
            return (lin(v[0], v[2]), lin(v[1], v[3]), lin(v[0], v[1]), lin(v[2], v[3]),
                    der(v[0], v[2]), der(v[1], v[3]), der(v[0], v[1]), der(v[2], v[3]))
        for v in (vx, vy):
            a = host.transfinite_blend(*edges(v), r, s)
            b = orc.transfinite_blend(*edges(v), r, s)
            for x, y in zip(a, b):
                worst = max(worst, np.max(np.abs(x - y)) / max(1.0, np.max(np.abs(y))))
        xf = lambda rr, ss: host.transfinite_blend(vx[0], vx[1], vx[2], vx[3], rr, ss)
        yf = lambda rr, ss: host.transfinite_blend(vy[0], vy[1], vy[2], vy[3], rr, ss)
        xo = lambda rr, ss: orc.transfinite_blend_corners(vx[0], vx[1], vx[2], vx[3], rr, ss)
        yo = lambda rr, ss: orc.transfinite_blend_corners(vy[0], vy[1], vy[2], vy[3], rr, ss)
        mh, mo = host.create_metrics(p, N, N, xf, yf), orc.create_metrics(p, N, N, xo, yo)
        for fld in ("crr", "css", "crs", "J"):
            a, b = getattr(mh, fld), getattr(mo, fld)
            assert np.max(np.abs(a - b)) <= 1e-13 * np.max(np.abs(b)), (e, fld)
        for k in range(4):
            for fld in ("sJ", "nx", "ny"):
                a, b = getattr(mh, fld)[k], getattr(mo, fld)[k]
                assert np.max(np.abs(a - b)) <= 1e-13 * max(1.0, np.max(np.abs(b))), (e, fld, k)
            for c in range(2):
                assert np.max(np.abs(mh.facecoord[c][k] - mo.facecoord[c][k])) <= 1e-13 * max(1.0, np.max(np.abs(mo.facecoord[c][k])))
    assert worst <= 2e-13, worst          # different order of the same additions (BP1_v1 spans 400 km)


@pytest.mark.parametrize("p", [2, 4, 6])
def test_host_d1_matrix_and_generated_table(p):
    """the host mirror's first-derivative operator (table generated by tools/gen_host_d1.py) against the oracle's"""
    import importlib.util
    import os
    from oracle import sbp as osbp
    for N in (3 * p + 1, 40):
        assert np.max(np.abs(host.d1_matrix(p, N) - osbp.diagonal_sbp_D1(p, N)[0].toarray())) < 1e-14
    if p > 2:
        with pytest.raises(ValueError):
            host.d1_matrix(p, p)                                           # grid too small for the closure (diagonal_sbp.jl:129-131)
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("gen_host_d1", os.path.join(root, "tools", "gen_host_d1.py"))
    mod = importlib.util.module_from_spec(spec); spec.loader.exec_module(mod)
    assert mod.render() == open(os.path.join(root, "hybridsbp_b200", "_sbp_d1.py")).read()
