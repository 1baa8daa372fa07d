"""Oracle against the golden vectors of tests/golden/refexec/: outputs of the reference's own statements, executed by
tests/refexec/minijulia.py where the reference tree is mounted (tools/gen_refexec_golden.py; tests/test_reference_executed.py checks
that the files are reproducible).  These tests need only the committed files, so they also run on the GPU box."""
import os

import numpy as np
import pytest

from oracle import hybrid as orc

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "refexec")


def curved_xf(r, s):
    return r + 0.1 * np.sin(2 * r) * np.cos(s) + 0.2 * s, 1 + 0.2 * np.cos(2 * r) * np.cos(s), -0.1 * np.sin(2 * r) * np.sin(s) + 0.2


def curved_yf(r, s):
    return s + 0.15 * np.sin(r + s), 0.15 * np.cos(r + s), 1 + 0.15 * np.cos(r + s)


@pytest.mark.parametrize("p", [2, 4, 6])
def test_oracle_locoperator_vs_reference_output(p):
    g = np.load(os.path.join(GOLD, "locoperator_p%d.npz" % p))
    N = int(g["N"])
    om = orc.create_metrics(p, N, N, curved_xf, curved_yf)
    for k in ("crr", "css", "crs", "J"):
        assert np.max(np.abs(getattr(om, k) - g[k])) <= 5e-15 * np.max(np.abs(g[k]))
    for k, bc in enumerate(g["bc"]):
        lop = orc.locoperator(p, N, N, om, tuple(int(b) for b in bc))
        u = g["u"][k]
        assert np.max(np.abs(lop.Mt @ u - g["y"][k])) <= 1e-14 * g["scale"][k]
        for lf in range(4):
            assert np.max(np.abs(lop.tau[lf].diagonal() - g["tau"][k][lf])) <= 1e-14 * np.max(g["tau"][k][lf])
            assert np.max(np.abs(lop.F[lf].T @ u - g["FTu"][k][lf])) <= 1e-13 * np.max(np.abs(g["FTu"][k][lf]))
            assert np.max(np.abs(lop.HfI_FT[lf] @ u - g["traction_op_u"][k][lf])) <= 1e-13 * np.max(np.abs(g["traction_op_u"][k][lf]))


@pytest.mark.parametrize("p", [4, 6])
def test_oracle_square_circle_vs_reference_output(p):
    from tests.refexec.oracle_driver import oracle_square_circle_level
    g = np.load(os.path.join(GOLD, "square_circle_p%d.npz" % p))
    o = oracle_square_circle_level(p, int(g["N"]))
    verts, EToV, EToF, FToB, dom = o["mesh"]
    assert np.array_equal(EToV, g["EToV"]) and np.array_equal(EToF, g["EToF"]) and np.array_equal(FToB, g["FToB"])
    assert np.array_equal(dom, g["EToDomain"]) and np.array_equal(verts, g["verts"])
    assert np.array_equal(o["vstarts"], g["vstarts"]) and np.array_equal(o["FTol"], g["FTolstarts"]) and np.array_equal(o["FTod"], g["FTodstarts"])
    rel = lambda a, b: np.linalg.norm(a - b) / np.linalg.norm(b)
    assert rel(o["delta"], g["delta"]) < 1e-14 and rel(o["gd"], g["gdelta"]) < 1e-13 and rel(o["g"][::37], g["g_sample"]) < 1e-13
    assert rel(o["bl"], g["blambda"]) < 1e-12
    assert rel(o["lam"], g["lam"]) < 1e-11 and rel(o["u"], g["u"]) < 1e-11
    assert abs(o["eps"] - g["eps"]) < 1e-6 * g["eps"] and abs(o["teps"] - g["teps"]) < 1e-6 * g["teps"]


@pytest.mark.skipif(not os.environ.get("HSBP_SLOW_TESTS"), reason="one minute of sparse factorisations; set HSBP_SLOW_TESTS=1 "
                    "(measured: lambda 1.4e-13, u 1.2e-13, profiles/r02c_reference_executed_convergence.txt)")
def test_oracle_square_circle_level3_vs_reference_output():
    from tests.refexec.oracle_driver import oracle_square_circle_level
    g = np.load(os.path.join(GOLD, "square_circle_p4_N68.npz"))
    o = oracle_square_circle_level(4, int(g["N"]))
    rel = lambda a, b: np.linalg.norm(a - b) / np.linalg.norm(b)
    assert np.array_equal(o["FTol"], g["FTolstarts"]) and np.array_equal(o["FTod"], g["FTodstarts"])
    assert rel(o["delta"], g["delta"]) < 1e-14 and rel(o["gd"], g["gdelta"]) < 1e-13
    assert rel(o["lam"], g["lam"]) < 1e-11 and rel(o["u"][::31], g["u_sample"]) < 1e-11
    assert abs(o["eps"] - g["eps"][2]) < 1e-4 * g["eps"][2] and abs(o["teps"] - g["teps"][2]) < 1e-6 * g["teps"][2]


def test_oracle_flower_vs_reference_output():
    from tests.refexec.oracle_driver import oracle_square_circle_level
    from hybridsbp_b200 import flower
    g = np.load(os.path.join(GOLD, "flower_p4.npz"))
    o = oracle_square_circle_level(4, int(g["N"]), mesh=flower.load_mesh(), maps=flower.block_maps, exact=flower.Smooth,
                                   slip=lambda x, y: 0.3 * np.sin(x) * np.cos(2 * y))
    assert np.array_equal(np.asarray(o["conn"][2]).astype(int), g["EToO"]) and np.array_equal(o["conn"][3], g["EToS"])
    assert np.array_equal(o["FTol"], g["FTolstarts"]) and np.array_equal(o["FTod"], g["FTodstarts"])
    rel = lambda a, b: np.linalg.norm(a - b) / np.linalg.norm(b)
    assert rel(o["delta"], g["delta"]) < 1e-14 and rel(o["gd"], g["gdelta"]) < 1e-13 and rel(o["bl"], g["blambda"]) < 1e-12
    assert rel(o["lam"], g["lam"]) < 1e-11 and rel(o["u"], g["u"]) < 1e-11 and rel(o["tauf"], g["traction"]) < 1e-11


def test_oracle_bp1_odefun_vs_reference_output():
    from hybridsbp_b200 import bp1
    from oracle.bp1 import OdeFun
    g = np.load(os.path.join(GOLD, "bp1_odefun_N40.npz"))
    N = int(g["N"])
    su = bp1.setup(N=N)
    assert np.allclose(su.psi_delta0, g["y0"], rtol=1e-14, atol=0) and np.allclose(su.RSa, g["RSa"], rtol=1e-15)
    assert abs(su.params["tau_z0"] - float(g["tau_z0"])) <= 1e-15 * float(g["tau_z0"])
    o = OdeFun(su.p, su.N, su.metrics, su.LFtoB, su.RSa, su.params)
    for t, y, d in zip(g["t"], g["y"], g["dydt"]):
        do, rejected = o(float(t), y)
        assert not rejected
        assert np.max(np.abs(do - d)) <= 1e-13 * np.max(np.abs(d))
