"""SEAS BP1 on the multiblock mesh meshes/BP1_v1.inp: trace solve with slip on the jump interfaces
(K4) + rate-and-state stage on the fault nodes (K5), GPU against the oracle's restatement with the reference's own functions
(oracle/bp1_multiblock.py: locbcarray! with the jump branch, assembleλmatrix + direct solve, computetraction, newtbndv)."""
import numpy as np
import pytest

from hybridsbp_b200 import bp1, bp1_multiblock as mb
from oracle.bp1_multiblock import MultiblockOdeFun

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", params=[(2, 5), (4, 11)])
def case(ctx, request):
    p, N = request.param
    su = mb.setup(N=N, SBPp=p)
    ref = MultiblockOdeFun(mb.default_mesh_path(), su.p, su.N, su.params, su.RSa, su.fault_faces, su.steady_faces, su.sign)
    gpu = mb.FaultOperator(ctx, su, mode="solve")
    yield su, gpu, ref
    gpu.close()


def test_mesh_roles(case):
    su, gpu, ref = case
    assert len(su.fault_faces) == 13 and len(su.steady_faces) == 9          # side sets 7 and 8 of BP1_v1.inp
    assert su.depth.min() >= -1e-12 and su.depth.max() <= 40 + 1e-9           # the frictional fault reaches 40 km
    assert gpu.n == ref.n == 13 * (su.N + 1)


def test_stress_change_and_rhs_match_the_oracle(case):
    su, gpu, ref = case
    n = gpu.n
    rng = np.random.default_rng(3)
    for t, dscale in ((0.0, 0.0), (3.0e8, 1e-2), (5.0e9, 2.0)):
        y = su.psi_delta0.copy()
        y[:n] += 0.02 * rng.uniform(-1, 1, n)
        y[n:] = dscale * rng.uniform(0, 1, n)
        s_gpu, s_ref = gpu.stress_change(y[n:], t), ref.stress_change(y[n:], t)
        assert gpu.trace_stats["converged"] == 1 and gpu.trace_stats["true_rel_residual"] <= 1e-11
        assert np.max(np.abs(s_gpu - s_ref)) <= 1e-9 * max(np.max(np.abs(s_ref)), 1e-12), (t, gpu.trace_stats)
        lam_gpu = gpu.dlam.get()
        assert np.linalg.norm(lam_gpu - ref.lam) <= 1e-10 * max(np.linalg.norm(ref.lam), 1e-300)
        assert np.linalg.norm(gpu.du.get() - ref.u) <= 1e-10 * max(np.linalg.norm(ref.u), 1e-300)
        d_gpu, rej_gpu = gpu.rhs(t, y)
        d_ref, rej_ref = ref(t, y)
        assert rej_gpu == rej_ref == False, gpu.last_stats
        assert np.max(np.abs(d_gpu[n:] - d_ref[n:])) <= 1e-8 * np.max(np.abs(d_ref[n:]))
        assert np.max(np.abs(d_gpu[:n] - d_ref[:n])) <= 1e-8 * np.max(np.abs(d_ref[:n]))
    # at t = 0 with the initial state the fault slips at the construction rate (BP1.jl:104-113)
    d0, _ = gpu.rhs(0.0, su.psi_delta0)
    assert np.abs(d0[n:] - 1e-9).max() < 1e-16


def test_condensed_fault_operator_and_short_integration(ctx, case):
    su, gpu, ref = case
    n = gpu.n
    cond = mb.FaultOperator(ctx, su, mode="condensed")
    rng = np.random.default_rng(4)
    y = su.psi_delta0.copy()
    y[:n] += 0.02 * rng.uniform(-1, 1, n)
    y[n:] = 0.3 * rng.uniform(0, 1, n)
    a, _ = cond.rhs(2.0e9, y)
    b, _ = gpu.rhs(2.0e9, y)
    assert np.max(np.abs(a[n:] - b[n:])) <= 1e-9 * np.max(np.abs(b[n:]))
    assert np.max(np.abs(a[:n] - b[:n])) <= 1e-9 * np.max(np.abs(b[:n]))
    # three years of loading, GPU (condensed) against the oracle, same integrator, at tolerances where the integration is well
    # conditioned (at the package defaults it runs at its stability limit and amplifies round-off, tests/test_bp1_gpu.py)
    t1 = 3 * bp1.YEAR_SECONDS
    from oracle.bp1 import tsit5
    ts_g, ys_g, _ = bp1.integrate(cond.rhs, su.psi_delta0, 0.0, t1, bp1.YEAR_SECONDS, abstol=1e-10, reltol=1e-7)
    ts_r, ys_r, _ = tsit5(ref, su.psi_delta0, 0.0, t1, bp1.YEAR_SECONDS, abstol=1e-10, reltol=1e-7)
    assert len(ts_g) == len(ts_r), (len(ts_g), len(ts_r))
    assert np.allclose(ts_g, ts_r, rtol=1e-7), np.max(np.abs(ts_g - ts_r) / ts_r[-1])
    e_slip = np.max(np.abs(ys_g[:, n:] - ys_r[:, n:])) / np.max(np.abs(ys_r[:, n:]))
    V_g = np.array([cond.rhs(t, y)[0][n:] for t, y in zip(ts_g, ys_g)])
    V_r = np.array([ref(t, y)[0][n:] for t, y in zip(ts_r, ys_r)])
    e_V = np.max(np.abs(V_g - V_r) / np.abs(V_r))
    assert e_slip <= 1e-6 and e_V <= 1e-6, (e_slip, e_V)
    cond.close()
