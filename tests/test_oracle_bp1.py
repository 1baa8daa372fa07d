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
    """Tsit5 + controller on y' = -y with the BP1 controls (infinity norm, rejection hook)."""
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


def test_tsit5_tableau_satisfies_the_order_conditions():
    """The tableau is restated from the publication (Tsitouras 2011), not from the reference tree (OrdinaryDiffEq is an
    unpinned dependency): pin it by the Runge-Kutta order conditions -- all 17 up to order 5 for the propagating weights
    b = a_7j, all 8 up to order 4 for the embedded weights b - BT."""
    A = np.zeros((7, 7))
    for i, row in enumerate(bp1.TSIT5_A):
        A[i, :len(row)] = row
    c = bp1.TSIT5_C
    assert np.allclose(A.sum(axis=1), c, atol=1e-15)
    e = np.ones(7)
    C = np.diag(c)
    b = A[6].copy()
    conds = {          # elementary weight, 1 / gamma(tree)
        1: [(lambda w: w @ e, 1.0)],
        2: [(lambda w: w @ c, 1 / 2)],
        3: [(lambda w: w @ c ** 2, 1 / 3), (lambda w: w @ A @ c, 1 / 6)],
        4: [(lambda w: w @ c ** 3, 1 / 4), (lambda w: w @ C @ A @ c, 1 / 8), (lambda w: w @ A @ c ** 2, 1 / 12), (lambda w: w @ A @ A @ c, 1 / 24)],
        5: [(lambda w: w @ c ** 4, 1 / 5), (lambda w: w @ C @ C @ A @ c, 1 / 10), (lambda w: w @ C @ A @ c ** 2, 1 / 15),
            (lambda w: w @ C @ A @ A @ c, 1 / 30), (lambda w: w @ (A @ c) ** 2, 1 / 20), (lambda w: w @ A @ c ** 3, 1 / 20),
            (lambda w: w @ A @ C @ A @ c, 1 / 40), (lambda w: w @ A @ A @ c ** 2, 1 / 60), (lambda w: w @ A @ A @ A @ c, 1 / 120)],
    }
    for order in range(1, 6):
        for fn, val in conds[order]:
            assert abs(fn(b) - val) < 5e-15, (order, fn(b), val)
    bhat = b - bp1.TSIT5_BT
    for order in range(1, 5):
        for fn, val in conds[order]:
            assert abs(fn(bhat) - val) < 5e-15, (order, fn(bhat), val)
    assert max(abs(fn(bhat) - val) for fn, val in conds[5]) > 1e-4          # the embedded solution is only 4th order


def test_tsit5_observed_order_and_the_two_restatements_agree():
    from oracle.bp1 import tsit5
    rhs = lambda t, y: (np.array([y[1], -y[0]]) * (1 + 0.5 * np.sin(t)), False)
    exact = lambda t: np.array([np.sin(t - 0.5 * np.cos(t) + 0.5), np.cos(t - 0.5 * np.cos(t) + 0.5)])
    errs = []
    for nsteps in (40, 80, 160):
        h = 2.0 / nsteps
        # fixed steps: a huge tolerance makes the controller accept everything, tstops force the grid
        ts, ys, _ = bp1.integrate(rhs, exact(0.0), 0.0, 2.0, h, abstol=1e30, reltol=1e30, tstops=np.arange(1, nsteps + 1) * h)
        assert len(ts) == nsteps + 1
        errs.append(np.abs(ys[-1] - exact(2.0)).max())
    rates = np.log2(np.array(errs[:-1]) / np.array(errs[1:]))
    assert np.all(rates > 4.7) and np.all(rates < 5.6), rates
    # adaptive runs of the host driver and of the oracle's own copy take the same steps
    a = bp1.integrate(rhs, exact(0.0), 0.0, 5.0, 0.3, abstol=1e-9, reltol=1e-7, tstops=[1.0, 2.5])
    o = tsit5(rhs, exact(0.0), 0.0, 5.0, 0.3, abstol=1e-9, reltol=1e-7, tstops=[1.0, 2.5])
    assert len(a[0]) == len(o[0]) and a[2] == o[2]
    assert np.array_equal(a[0], o[0])
    assert np.abs(a[1] - o[1]).max() < 1e-13
    assert 1.0 in a[0] and 2.5 in a[0]
    assert np.abs(a[1][-1] - exact(5.0)).max() < 1e-6
    # out-of-domain stages shrink the step by qmin = 1/5 (loopheader! of the reference's integrator)
    calls = []
    def picky(t, y):
        calls.append(t)
        return -y, (t > 0.55 and len(calls) < 12)
    ts, ys, nrej = bp1.integrate(picky, np.array([1.0]), 0.0, 1.0, 1.0, abstol=1e-6, reltol=1e-3)
    assert nrej >= 1 and abs(ts[-1] - 1.0) < 1e-14
