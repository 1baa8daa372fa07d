#!/usr/bin/env python3
"""Golden series for the BP1 parity test THROUGH THE FIRST EARTHQUAKE (tests/test_bp1_gpu.py).

TEST INFRASTRUCTURE (oracle side).  Runs the oracle's odefun (oracle/bp1.py, seas/BP1/odefun.jl:8-121) with the oracle's
own Tsit5 (BP1.jl:159-161) at a resolution the CPU finishes in seconds (N = 16) and at tolerances where the integration
is well conditioned (reltol 1e-8, abstol 1e-11), forcing output at fixed times: every 20 years of the interseismic
phase, the approach to the first event (30 days ... 1 minute before), every second of the 18 s of coseismic slip, and
the start of the post-seismic phase.  Stored per output time: the state [psi; delta], odefun's output [dpsi; V] and
dV/dt along the trajectory (a directional finite difference of odefun), which the test needs to separate a shift in
the timing of the event from a change of the trajectory itself.

  python tools/gen_bp1_event_golden.py        -> tests/golden/bp1/event_N16.npz"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from hybridsbp_b200 import bp1          # setup() only: host arithmetic of BP1.jl:8-146
from oracle.bp1 import OdeFun, tsit5

N, RELTOL, ABSTOL, T_END_YEARS = 16, 1e-8, 1e-11, 250
Y = bp1.YEAR_SECONDS


def main():
    su = bp1.setup(N=N)
    f = OdeFun(su.p, su.N, su.metrics, su.LFtoB, su.RSa, su.params)
    n = N + 1
    ts, ys, _ = tsit5(f, su.psi_delta0, 0.0, (T_END_YEARS + 10) * Y, Y, abstol=ABSTOL, reltol=RELTOL)
    V = np.array([np.abs(f(t, y)[0][n:]).max() for t, y in zip(ts, ys)])
    fast = np.where(V > 1e-3)[0]
    t_on = float(np.floor(ts[fast[0]]))
    print("first event: slip rate above 1 mm/s from t = %.3f s (%.4f years) for %.1f s, peak %.3f m/s" %
          (ts[fast[0]], ts[fast[0]] / Y, ts[fast[-1]] - ts[fast[0]], V.max()))
    stops = sorted(list(np.arange(20, 240, 20) * float(Y)) + [t_on - 86400.0 * k for k in (30, 5, 1)] + [t_on - 600.0, t_on - 60.0] +
                   list(t_on + np.arange(0.0, 30.0, 1.0)) + [t_on + 3600.0, float(T_END_YEARS * Y)])
    ts, ys, nrej = tsit5(f, su.psi_delta0, 0.0, T_END_YEARS * Y, Y, abstol=ABSTOL, reltol=RELTOL, tstops=stops)
    idx = [int(np.where(ts == s)[0][0]) for s in stops]
    y = ys[idx]
    F = np.array([f(s, yy)[0] for s, yy in zip(stops, y)])
    Vdot = []
    for s, yy, ff in zip(stops, y, F):
        h = 1e-7 / max(1e-30, np.abs(ff[n:]).max())                  # a slip increment of 0.1 micrometre along the trajectory
        Vdot.append((f(s + h, yy + h * ff)[0][n:] - ff[n:]) / h)
    out = os.path.join(ROOT, "tests", "golden", "bp1", "event_N16.npz")
    np.savez_compressed(out, N=N, reltol=RELTOL, abstol=ABSTOL, t=np.array(stops), y=y, F=F, Vdot=np.array(Vdot), t_onset=t_on,
                        steps=len(ts) - 1, rejected=nrej)
    print("wrote", out, "(%d output times, %d accepted + %d rejected steps)" % (len(stops), len(ts) - 1, nrej))


if __name__ == "__main__":
    main()
