"""CPU check of the line-marching kernel's algorithm (tools/proto_sweep.py emulates k_sweep.cuh step by
step: pair form, lagged windows, prologue closures, read-modify-write of the dense Q^T closure block,
both marching directions, chunk seams) against the oracle's assembled volume operator A-tilde
(reference locoperator, global_curved.jl:261-356), and of the generated closure tables."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))

from oracle import hybrid as orc  # noqa: E402
from oracle import sbp  # noqa: E402
from tests.util import random_spd_metrics  # noqa: E402
import proto_sweep as ps  # noqa: E402
from sbp_coeffs import Coeffs  # noqa: E402


@pytest.mark.parametrize("p,N", [(2, 9), (4, 15), (4, 40), (6, 23), (6, 50)])
def test_pair_form_of_the_1d_operators(p, N):
    rng = np.random.default_rng(p * 7 + N)
    cf = Coeffs(p)
    b = rng.random(N + 1) + 0.05
    u = rng.uniform(-1, 1, N + 1)
    M = sbp.stiffness_dense(p, N, b)
    assert np.abs(M.sum(axis=1)).max() < 1e-14 * np.abs(M).max() * (N + 1)      # zero row sums: pair form is exact
    assert np.abs(M - M.T).max() == 0.0
    assert np.abs(ps.m_apply_line(cf, b, u) - M @ u).max() < 1e-14 * np.abs(M).max() * 4
    D, _, H, _ = sbp.diagonal_sbp_D1(p, N)
    Q = (H @ D).toarray()
    assert np.abs(cf.Q_dense(N + 1) - Q).max() < 1e-15
    assert np.abs(ps.q_apply_line(cf, u) - Q @ u).max() < 1e-14
    assert np.abs(ps.qt_apply_line(cf, u) - Q.T @ u).max() < 1e-14


@pytest.mark.parametrize("p,Nr,Ns,nch", [(2, 15, 31, 1), (2, 20, 40, 2), (4, 31, 31, 1), (4, 25, 63, 2), (4, 20, 100, 3),
                                         (6, 35, 35, 1), (6, 30, 70, 2)])
def test_marching_algorithm_matches_assembled_volume_operator(p, Nr, Ns, nch):
    rng = np.random.default_rng(100 * p + Nr + Ns + nch)
    m = random_spd_metrics(p, Nr, Ns, rng, scale2=0.05)
    lop = orc.locoperator(p, Nr, Ns, m, (1, 1, 1, 1))
    u = rng.uniform(-1, 1, (Nr + 1, Ns + 1))
    y = ps.sweep_block(p, u, m.crr, m.css, m.crs, nch)
    uf = u.reshape(-1, order="F")
    yref = (lop.A @ uf).reshape(Nr + 1, Ns + 1, order="F")
    scale = np.max(abs(lop.A) @ np.abs(uf))
    assert not np.isnan(y).any()
    assert np.max(np.abs(y - yref)) / scale < 1e-14


def test_generated_header_is_current():
    """hybridsbp_b200/csrc/sweep_tables_gen.h must be what tools/gen_sweep_tables.py writes from the tables"""
    import gen_sweep_tables as g
    out = []
    for p in (2, 4, 6):
        g.emit(p, out)
    txt = open(os.path.join(ROOT, "hybridsbp_b200", "csrc", "sweep_tables_gen.h")).read()
    for line in out:
        assert line in txt, line[:80]
