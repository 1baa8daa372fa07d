"""CUDA path against outputs of the reference itself (tests/golden/refexec/: the reference's Julia statements executed by
tests/refexec/minijulia.py, see tools/gen_refexec_golden.py) -- no oracle in between.  Through the C-ABI: operator apply, penalty
parameters and face operators on a curved block with every boundary-condition type; the first level of square_circle.jl (lambda and u
within 1e-10, the north-star tolerance); the BP1 right-hand side."""
import os
from types import SimpleNamespace

import numpy as np
import pytest

from tests.util import upload_blocks

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "refexec")


@pytest.mark.parametrize("generic", [False, True])
@pytest.mark.parametrize("p", [2, 4, 6])
def test_apply_tau_and_face_operators_vs_reference_output(ctx, p, generic):
    import hybridsbp_b200 as hs
    g = np.load(os.path.join(GOLD, "locoperator_p%d.npz" % p))
    nb = len(g["bc"])
    m = SimpleNamespace(crr=g["crr"], css=g["css"], crs=g["crs"])
    blk = upload_blocks(hs, ctx, p, [m] * nb, g["bc"])                       # tauscale = 2: the reference's default
    if generic:
        blk.force_generic(True)
    u = g["u"].reshape(-1)
    y = blk.apply_host(u)
    tau = blk.get_tau()
    du = ctx.array(u)
    dft, dtr = ctx.empty(blk.FNp), ctx.empty(blk.FNp)
    blk.face_FT(du, dft); blk.face_traction(du, dtr)
    ft, tr = dft.get(), dtr.get()
    for e in range(nb):
        sl = blk.vol_slice(e)
        assert np.max(np.abs(y[sl] - g["y"][e])) <= 1e-12 * g["scale"][e], (p, e)
        for lf in range(1, 5):
            fs = blk.face_slice(e, lf)
            assert np.max(np.abs(tau[fs] - g["tau"][e][lf - 1])) <= 1e-13 * np.max(g["tau"][e][lf - 1])
            assert np.max(np.abs(ft[fs] - g["FTu"][e][lf - 1])) <= 1e-11 * np.max(np.abs(g["FTu"][e][lf - 1]))
            assert np.max(np.abs(tr[fs] - g["traction_op_u"][e][lf - 1])) <= 1e-11 * np.max(np.abs(g["traction_op_u"][e][lf - 1]))
    blk.close()


@pytest.mark.parametrize("p", [4, 6])
def test_square_circle_level1_vs_reference_output(ctx, p):
    from hybridsbp_b200 import square_circle as sc
    g = np.load(os.path.join(GOLD, "square_circle_p%d.npz" % p))
    mesh = sc.load_mesh(sc.default_mesh_path())
    verts, EToV, EToF, FToB, dom = mesh
    assert np.array_equal(EToV, g["EToV"]) and np.array_equal(EToF, g["EToF"]) and np.array_equal(FToB, g["FToB"])
    assert np.array_equal(verts, g["verts"]) and np.array_equal(dom, g["EToDomain"])
    r = sc.solve_level(ctx, mesh, p, int(g["N"]), tol=1e-13)
    assert r["stats"]["converged"] == 1, r["stats"]
    rel = lambda a, b: np.linalg.norm(a - b) / np.linalg.norm(b)
    assert np.array_equal(r["FTols"], g["FTolstarts"]) and np.array_equal(r["FTods"], g["FTodstarts"])
    assert rel(r["delta"], g["delta"]) <= 1e-13
    assert rel(r["gd"], g["gdelta"]) <= 1e-12
    assert rel(r["g_full"][::37], g["g_sample"]) <= 1e-12
    assert rel(r["lam"], g["lam"]) <= 1e-10, r["stats"]
    assert rel(r["u"], g["u"]) <= 1e-10, r["stats"]


def test_square_circle_level3_vs_reference_output(ctx):
    """third level of square_circle.jl (N = 68: 69 points per line, banded block factors, condensed trace system with the two-level
    preconditioner) against the executed reference's direct sparse solves"""
    from hybridsbp_b200 import square_circle as sc
    g = np.load(os.path.join(GOLD, "square_circle_p4_N68.npz"))
    r = sc.solve_level(ctx, sc.load_mesh(sc.default_mesh_path()), 4, int(g["N"]), tol=1e-13)
    assert r["stats"]["converged"] == 1, r["stats"]
    rel = lambda a, b: np.linalg.norm(a - b) / np.linalg.norm(b)
    assert np.array_equal(r["FTols"], g["FTolstarts"]) and np.array_equal(r["FTods"], g["FTodstarts"])
    assert rel(r["delta"], g["delta"]) <= 1e-13 and rel(r["gd"], g["gdelta"]) <= 1e-12
    assert rel(r["lam"], g["lam"]) <= 1e-10, (rel(r["lam"], g["lam"]), r["stats"])
    assert rel(r["u"][::31], g["u_sample"]) <= 1e-10 and abs(np.linalg.norm(r["u"]) - float(g["u_norm"])) <= 1e-10 * float(g["u_norm"])
    assert abs(r["eps"] - g["eps"][2]) <= 1e-2 * g["eps"][2] and abs(r["tau_eps"] - g["teps"][2]) <= 1e-3 * g["teps"][2]


def test_flower_reversed_faces_vs_reference_output(ctx):
    """67 blocks, 27 faces seen in reversed orientation from their plus side, given slip on the 18 jump faces"""
    from hybridsbp_b200 import flower
    g = np.load(os.path.join(GOLD, "flower_p4.npz"))
    r = flower.solve_level(ctx, flower.load_mesh(), 4, int(g["N"]), tol=1e-13, slip=lambda x, y: 0.3 * np.sin(x) * np.cos(2 * y))
    assert r["stats"]["converged"] == 1, r["stats"]
    rel = lambda a, b: np.linalg.norm(a - b) / np.linalg.norm(b)
    assert np.array_equal(r["FTols"], g["FTolstarts"]) and np.array_equal(r["FTods"], g["FTodstarts"])
    assert rel(r["delta"], g["delta"]) <= 1e-13 and rel(r["gd"], g["gdelta"]) <= 1e-12 and rel(r["g_full"][::37], g["g_sample"]) <= 1e-12
    assert rel(r["lam"], g["lam"]) <= 1e-10, r["stats"]
    assert rel(r["u"], g["u"]) <= 1e-10, r["stats"]


def test_bp1_rhs_vs_reference_output(ctx):
    from hybridsbp_b200 import bp1, LOCAL_BAND
    g = np.load(os.path.join(GOLD, "bp1_odefun_N40.npz"))
    N = int(g["N"])
    n = N + 1
    su = bp1.setup(N=N)
    assert np.allclose(su.psi_delta0, g["y0"], rtol=1e-14, atol=0)
    for condense in (True, False):
        gpu = bp1.Fault(ctx, su, local_mode=LOCAL_BAND, condense=condense)
        for t, y, d in zip(g["t"], g["y"], g["dydt"]):
            dg, rejected = gpu.rhs(float(t), y)
            assert not rejected, gpu.last_stats
            assert np.max(np.abs(dg[n:] - d[n:])) <= 1e-8 * np.max(np.abs(d[n:])), gpu.last_stats
            assert np.max(np.abs(dg[:n] - d[:n])) <= 1e-8 * np.max(np.abs(d[:n]))
        gpu.close()
