"""GPU parity of the BP1 ODE stage (hsbp_bp1_rhs: boundary scatter, local solve, traction, per-node bracketed
Newton, state evolution) against the oracle's odefun restatement (seas/BP1/odefun.jl:8-121), and of a short
earthquake-cycle integration (north star: slip / slip-rate series within 1e-6 relative)."""
import numpy as np
import pytest

from hybridsbp_b200 import bp1
from oracle.bp1 import OdeFun

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", params=["band", "pcg"])
def case(ctx, request):
    from hybridsbp_b200 import LOCAL_BAND, LOCAL_PCG
    su = bp1.setup(N=40)
    gpu = bp1.Fault(ctx, su, local_mode=LOCAL_BAND if request.param == "band" else LOCAL_PCG)
    ref = OdeFun(su.p, su.N, su.metrics, su.LFtoB, su.RSa, su.params)
    yield su, gpu, ref
    gpu.close()


def test_rhs_matches_oracle(case):
    su, gpu, ref = case
    n = su.N + 1
    rng = np.random.default_rng(5)
    for t, dscale in ((0.0, 0.0), (3.0e7, 1e-2), (9.0e8, 0.5)):
        y = su.psi_delta0.copy()
        y[:n] += 0.02 * rng.uniform(-1, 1, n)
        y[n:] = dscale * rng.uniform(0, 1, n)
        d_gpu, rej_gpu = gpu.rhs(t, y)
        d_ref, rej_ref = ref(t, y)
        assert rej_gpu == rej_ref == False, gpu.last_stats
        V_ref, V_gpu = d_ref[n:], d_gpu[n:]
        assert np.max(np.abs(V_gpu - V_ref)) <= 1e-8 * np.max(np.abs(V_ref)), gpu.last_stats
        assert np.max(np.abs(d_gpu[:n] - d_ref[:n])) <= 1e-8 * np.max(np.abs(d_ref[:n]))
        u = gpu.displacement()
        assert np.linalg.norm(u - ref.u) <= 1e-10 * max(np.linalg.norm(ref.u), 1e-300)


def test_rejection_is_reported_not_raised(case):
    su, gpu, ref = case
    y = su.psi_delta0.copy()
    y[3] = np.nan
    d, rejected = gpu.rhs(0.0, y)
    assert rejected and gpu.last_stats["failed_nodes"] >= 1
    assert ref(0.0, y)[1]


def test_short_cycle_integration(case):
    su, gpu, ref = case
    n = su.N + 1
    t1 = 3 * bp1.YEAR_SECONDS
    ts_g, ys_g, rej_g = bp1.integrate(gpu.rhs, su.psi_delta0, 0.0, t1, bp1.YEAR_SECONDS)
    ts_r, ys_r, rej_r = bp1.integrate(ref, su.psi_delta0, 0.0, t1, bp1.YEAR_SECONDS)
    assert len(ts_g) == len(ts_r) and np.allclose(ts_g, ts_r, rtol=1e-9)
    slip_g, slip_r = ys_g[:, n:], ys_r[:, n:]
    assert np.max(np.abs(slip_g - slip_r)) <= 1e-6 * np.max(np.abs(slip_r))
    V_g = np.array([gpu.rhs(t, y)[0][n:] for t, y in zip(ts_g, ys_g)])
    V_r = np.array([ref(t, y)[0][n:] for t, y in zip(ts_r, ys_r)])
    assert np.max(np.abs(V_g - V_r) / np.abs(V_r)) <= 1e-6


@pytest.mark.parametrize("mode", ["band", "pcg"])
def test_rhs_at_reference_resolution_matches_stored_oracle_output(ctx, mode):
    """N = 200 (the reference's BP1 resolution, BP1.jl:8): odefun at five states of the first earthquake cycle, including
    the coseismic phase, against the oracle's stored output (tests/golden/bp1)."""
    import os
    from hybridsbp_b200 import LOCAL_BAND, LOCAL_PCG
    gold = np.load(os.path.join(os.path.dirname(__file__), "golden", "bp1", "states_N200.npz"))
    su = bp1.setup(N=200)
    n = su.N + 1
    f = bp1.Fault(ctx, su, local_tol=1e-13, local_mode=LOCAL_BAND if mode == "band" else LOCAL_PCG)
    for t, y, d_ref in zip(gold["t"], gold["y"], gold["dpsiV"]):
        d, rej = f.rhs(float(t), y)
        assert not rej
        tol = 1e-9 if mode == "band" else 1e-7
        assert np.max(np.abs(d[n:] - d_ref[n:])) <= tol * np.max(np.abs(d_ref[n:]))
        assert np.max(np.abs(d[:n] - d_ref[:n])) <= tol * max(np.max(np.abs(d_ref[:n])), 1e-300)
    f.close()
