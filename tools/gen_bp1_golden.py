"""Regenerate tests/golden/bp1/states_N200.npz (CPU oracle only; about 4 minutes):
    python tools/bp1_run.py oracle 200 315 /tmp/oracle_N200_315y.npz
    python tools/gen_bp1_golden.py /tmp/oracle_N200_315y.npz
States: steps 0, 5, 12, the middle and the end of the stored series; dpsiV = the oracle's odefun at those states."""
import sys
import numpy as np
sys.path.insert(0, ".")
from hybridsbp_b200 import bp1
from oracle.bp1 import OdeFun

o = np.load(sys.argv[1])
idx = [0, 5, 12, len(o["t"]) // 2, len(o["t"]) - 1]
su = bp1.setup(N=200)
ref = OdeFun(su.p, su.N, su.metrics, su.LFtoB, su.RSa, su.params)
t, y = o["t"][idx], o["y"][idx]
d = np.array([ref(tt, yy)[0] for tt, yy in zip(t, y)])
np.savez_compressed("tests/golden/bp1/states_N200.npz", t=t, y=y, dpsiV=d)
