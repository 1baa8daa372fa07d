"""GPU parity of the BP1 ODE stage (hsbp_bp1_rhs: boundary scatter, local solve, traction, per-node bracketed
Newton, state evolution) against the oracle's odefun restatement (seas/BP1/odefun.jl:8-121), and of a short
earthquake-cycle integration (north star: slip / slip-rate series within 1e-6 relative)."""
import numpy as np
import pytest

from hybridsbp_b200 import bp1
from oracle.bp1 import OdeFun

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", params=["band", "pcg", "band_solve_every_call"])
def case(ctx, request):
    """band / pcg: the local solve condensed onto the fault (hsbp_bp1_condense, one small kernel per odefun);
    band_solve_every_call: one banded back-solve per odefun call, as the reference does (odefun.jl:43)"""
    from hybridsbp_b200 import LOCAL_BAND, LOCAL_PCG
    su = bp1.setup(N=40)
    gpu = bp1.Fault(ctx, su, local_mode=LOCAL_PCG if request.param == "pcg" else LOCAL_BAND,
                    condense=request.param != "band_solve_every_call")
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


@pytest.mark.parametrize("mode", ["band", "pcg", "band_solve_every_call"])
def test_rhs_at_reference_resolution_matches_stored_oracle_output(ctx, mode):
    """N = 200 (the reference's BP1 resolution, BP1.jl:8): odefun at five states of the first earthquake cycle, including
    the coseismic phase, against the oracle's stored output (tests/golden/bp1)."""
    import os
    from hybridsbp_b200 import LOCAL_BAND, LOCAL_PCG
    gold = np.load(os.path.join(os.path.dirname(__file__), "golden", "bp1", "states_N200.npz"))
    su = bp1.setup(N=200)
    n = su.N + 1
    f = bp1.Fault(ctx, su, local_tol=1e-13, local_mode=LOCAL_PCG if mode == "pcg" else LOCAL_BAND,
                  condense=mode != "band_solve_every_call")
    for t, y, d_ref in zip(gold["t"], gold["y"], gold["dpsiV"]):
        d, rej = f.rhs(float(t), y)
        assert not rej
        tol = 1e-7 if mode == "pcg" else 1e-9
        assert np.max(np.abs(d[n:] - d_ref[n:])) <= tol * np.max(np.abs(d_ref[n:]))
        assert np.max(np.abs(d[:n] - d_ref[:n])) <= tol * max(np.max(np.abs(d_ref[:n])), 1e-300)
    f.close()


def test_series_through_the_first_earthquake_matches_the_oracle(ctx):
    """North star: slip / slip-rate series within 1e-6.  The GPU odefun (condensed fault operator, banded factor) drives the
    host's Tsit5 through 235 years of loading and the first earthquake at tolerances where the integration is well
    conditioned (reltol 1e-8, abstol 1e-11), with output forced at the fixed times of the oracle's stored series
    (tests/golden/bp1/event_N16.npz, tools/gen_bp1_event_golden.py).
      * before the event (up to one minute before onset) the series agree at fixed times to 1e-6;
      * inside the event a fixed-time comparison is ill-posed: the onset time of a frictional instability is
        exponentially sensitive (perturbing the ORACLE's own odefun output by 1e-13 relative moves V at fixed coseismic
        times by 3e-5).  There the two trajectories are compared up to a shift dt_k along the
        trajectory, y_gpu(t_k) = y_ref(t_k) + dt_k y'_ref(t_k) + residual: slip, state and slip-rate residuals within 1e-6
        and the shift itself below a millisecond after 7.4e9 s.  Measured on B200: shift 3.7e-5 s, residuals 1e-14 (slip),
        2e-13 (state), 3e-11 (slip rate); without the alignment V differs by 7e-6 at fixed coseismic times, 2e-8 before."""
    import os
    from hybridsbp_b200 import LOCAL_BAND
    gold = np.load(os.path.join(os.path.dirname(__file__), "golden", "bp1", "event_N16.npz"))
    N = int(gold["N"])
    n = N + 1
    su = bp1.setup(N=N)
    f = bp1.Fault(ctx, su, local_mode=LOCAL_BAND)
    stops = [float(x) for x in gold["t"]]
    ts, ys, nrej = bp1.integrate(f.rhs, su.psi_delta0, 0.0, stops[-1], bp1.YEAR_SECONDS, abstol=float(gold["abstol"]),
                                 reltol=float(gold["reltol"]), tstops=stops)
    assert abs((len(ts) - 1) - int(gold["steps"])) <= 2, (len(ts) - 1, int(gold["steps"]))      # the same step sequence
    idx = [int(np.where(ts == s_)[0][0]) for s_ in stops]
    yg, yr, Fr, Vdot = ys[idx], gold["y"], gold["F"], gold["Vdot"]
    Vg = np.array([f.rhs(t, y)[0][n:] for t, y in zip(stops, yg)])
    Vr = Fr[:, n:]
    t_on = float(gold["t_onset"])
    worst = dict(pre_slip=0.0, pre_V=0.0, shift=0.0, slip=0.0, psi=0.0, V=0.0, fixed_time_V=0.0)
    for k, t in enumerate(stops):
        d = yg[k] - yr[k]
        slip_scale = max(np.abs(yr[k][n:]).max(), 1e-6)
        if t <= t_on - 60.0:
            worst["pre_slip"] = max(worst["pre_slip"], np.abs(d[n:]).max() / slip_scale)
            worst["pre_V"] = max(worst["pre_V"], (np.abs(Vg[k] - Vr[k]) / np.abs(Vr[k])).max())
        worst["fixed_time_V"] = max(worst["fixed_time_V"], (np.abs(Vg[k] - Vr[k]) / np.abs(Vr[k])).max())
        dt_k = float((d * Fr[k]).sum() / (Fr[k] * Fr[k]).sum())
        r = d - dt_k * Fr[k]
        worst["shift"] = max(worst["shift"], abs(dt_k))
        worst["slip"] = max(worst["slip"], np.abs(r[n:]).max() / slip_scale)
        worst["psi"] = max(worst["psi"], np.abs(r[:n]).max() / np.abs(yr[k][:n]).max())
        worst["V"] = max(worst["V"], (np.abs(Vg[k] - Vr[k] - dt_k * Vdot[k]) / np.abs(Vr[k])).max())
    print("BP1 through the first earthquake, GPU vs oracle:", {k: float("%.3g" % v) for k, v in worst.items()})
    assert Vr.max() > 0.5                                  # the stored series does contain the earthquake (peak 1.18 m/s between samples)
    assert worst["pre_slip"] <= 1e-6 and worst["pre_V"] <= 1e-6, worst
    assert worst["slip"] <= 1e-6 and worst["psi"] <= 1e-6, worst
    assert worst["V"] <= 1e-6 and worst["shift"] <= 1e-3, worst
    assert worst["fixed_time_V"] <= 1e-4, worst
    f.close()
