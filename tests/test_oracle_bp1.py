"""Pins the oracle's BP1 right-hand side (oracle/bp1.py, seas/BP1/odefun.jl) by construction identities: the
reference stores no BP1 output.  tau_z0 and theta are built so that the initial slip rate is RSVinit = 1e-9
everywhere on the fault (BP1.jl:104-113)."""
import numpy as np

from hybridsbp_b200 import bp1
from oracle.bp1 import OdeFun


def test_initial_slip_rate_is_the_construction_rate():
    su = bp1.setup(N=24)
    f = OdeFun(su.p, su.N, su.metrics, su.LFtoB, su.RSa, su.params)
    dy, rejected = f(0.0, su.psi_delta0)
    n = su.N + 1
    assert not rejected
    assert np.abs(dy[n:] - 1e-9).max() < 1e-17
    # a(depth) ramps from 0.010 to 0.025 between 15 and 18 km (BP1.jl:96-102)
    assert su.RSa.min() == 0.01 and su.RSa.max() == 0.025
    assert np.all(np.diff(su.RSa) >= 0)


def test_rejection_flag_on_unbracketed_root():
    su = bp1.setup(N=24)
    f = OdeFun(su.p, su.N, su.metrics, su.LFtoB, su.RSa, su.params)
    y = su.psi_delta0.copy()
    y[0] = np.nan                      # psi NaN -> residual NaN -> no convergence -> reject (odefun.jl:91-96)
    dy, rejected = f(0.0, y)
    assert rejected


def test_integrator_reproduces_a_known_solution():
    """Dormand-Prince pair + controller on y' = -y with the BP1 controls (infinity norm, rejection hook)."""
    calls = {"n": 0}

    def rhs(t, y):
        calls["n"] += 1
        return -y, False
    ts, ys, nrej = bp1.integrate(rhs, np.array([1.0, 2.0]), 0.0, 3.0, 0.5, abstol=1e-10, reltol=1e-10)
    assert abs(ts[-1] - 3.0) < 1e-14
    assert np.abs(ys[-1] - np.array([1.0, 2.0]) * np.exp(-3.0)).max() < 1e-9


def test_integrator_takes_millisecond_steps_at_late_times():
    """The smallest step of the restated integrator is the resolution of t (OrdinaryDiffEq's dtmin), not a fraction of t:
    coseismic steps are milliseconds at t ~ 1e10 s (the first BP1 earthquake at N = 200 starts at 9.7e9 s)."""
    import numpy as np
    from hybridsbp_b200 import bp1
    t0 = 1.0e10
    rhs = lambda t, y: (-1000.0 * y, False)                       # needs dt < 3e-3 for stability
    ts, ys, nrej = bp1.integrate(rhs, np.array([1.0]), t0, t0 + 0.05, 1.0e7, abstol=1e-8, reltol=1e-6)
    assert abs(ts[-1] - (t0 + 0.05)) <= 4 * np.spacing(t0)
    assert np.min(np.diff(ts)) < 3e-3
    assert abs(ys[-1][0] - np.exp(-1000.0 * (ts[-1] - t0))) <= 1e-4
    # a right-hand side that always rejects ends in a reported underflow, not an endless loop
    with __import__("pytest").raises(RuntimeError):
        bp1.integrate(lambda t, y: (y, t > t0), np.array([1.0]), t0, t0 + 1.0, 0.5)
    ts2, _, _ = bp1.integrate(lambda t, y: (y, t > t0), np.array([1.0]), t0, t0 + 1.0, 0.5, stop_on_underflow=True)
    assert len(ts2) == 1
