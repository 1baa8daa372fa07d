"""Pins the oracle's 1-D SBP restatement (oracle/sbp.py) with the identities the reference itself
relies on or checks: the reference ships no golden vectors."""
import numpy as np
import pytest

from oracle import sbp


@pytest.mark.parametrize("p", [2, 4, 6])
@pytest.mark.parametrize("N", [23, 40])
def test_d1_sbp_property_and_accuracy(p, N):
    D, HI, H, r = sbp.diagonal_sbp_D1(p, N)
    Q = (H @ D).toarray()
    B = np.zeros((N + 1, N + 1)); B[0, 0] = -1.0; B[N, N] = 1.0
    assert np.abs(Q + Q.T - B).max() < 1e-14                      # summation by parts
    assert np.abs((H @ HI).toarray() - np.eye(N + 1)).max() < 1e-14
    assert abs(H.diagonal().sum() - 2.0) < 1e-13                  # quadrature of the unit function on [-1, 1]
    # boundary accuracy p/2, interior accuracy p (diagonal_sbp.jl:67-161)
    for k in range(p // 2 + 1):
        exact = k * r ** (k - 1) if k > 0 else np.zeros_like(r)
        assert np.abs(D @ r ** k - exact).max() < 1e-11, k
    bm = {2: 1, 4: 4, 6: 6}[p]
    for k in range(p + 1):
        exact = k * r ** (k - 1) if k > 0 else np.zeros_like(r)
        assert np.abs((D @ r ** k - exact)[bm + 3:N - bm - 2]).max() < 1e-10, k


@pytest.mark.parametrize("p", [2, 4, 6])
def test_variable_d2_structure(p):
    N = 30
    rng = np.random.default_rng(p)
    b = rng.random(N + 1) + 0.1
    D, S0, SN, HI, H, M, r = sbp.variable_diagonal_sbp_D2(p, N, b)
    M = M.toarray()
    assert np.abs(M - M.T).max() == 0.0                           # symmetric stiffness matrix
    assert np.abs(M.sum(axis=1)).max() < 1e-12                    # constants are in the null space
    assert np.linalg.eigvalsh(M).min() > -1e-12                   # positive semi-definite
    # mirror symmetry: reversing b reverses the operator (diagonal_sbp.jl:539-562, 642-687)
    _, _, _, _, _, Mr, _ = sbp.variable_diagonal_sbp_D2(p, N, b[::-1])
    assert np.abs(Mr.toarray()[::-1, ::-1] - M).max() < 1e-13
    # D = HI (-M + SN - S0) differentiates (b u')' for low-degree u with constant and linear b
    for bb, ub, exact in ((np.ones(N + 1), r ** 2, 2 * np.ones(N + 1)),
                          (1 + 0.3 * r, r ** 2, 2 * (1 + 0.3 * r) + 0.6 * r)):
        Dv = sbp.variable_diagonal_sbp_D2(p, N, bb)[0]
        err = np.abs(Dv @ ub - exact)
        assert err[3:-3].max() < 1e-10
        if p > 2 or bb[0] == bb[-1]:      # the p = 2 closure is first-order accurate: exact for constant b only
            assert err.max() < 1e-10
    # boundary derivative rows: S0 = -b_0 BS / h on row 0, SN = +b_N BS / h reversed on row N (:755-757)
    h = 2.0 / N
    BS = sbp.D2VAR_BS[p]
    assert np.allclose(S0.toarray()[0, :len(BS)], -b[0] * BS / h)
    assert np.allclose(SN.toarray()[N, ::-1][:len(BS)], b[N] * BS / h)
    assert np.abs((-BS / h) @ r[:len(BS)] - 1.0) < 1e-11 and abs(BS.sum()) < 1e-14   # BS is a first derivative (sign: outward at 0)


@pytest.mark.parametrize("p", [2, 4, 6])
def test_remainder_is_positive_semidefinite(p):
    """check_residual.jl:8-17: R = A - D1^T H diag(b) D1 has non-negative real eigenvalues."""
    N = 3 * p + 12
    rng = np.random.default_rng(10 + p)
    for _ in range(5):
        b = rng.random(N + 1)
        D1, _, H, _ = sbp.diagonal_sbp_D1(p, N)
        A = sbp.variable_diagonal_sbp_D2(p, N, b)[5]
        import scipy.sparse as sp
        R = (A - D1.T @ H @ sp.diags(b) @ D1).toarray()
        ev = np.linalg.eigvals(R)
        assert np.abs(ev.imag).max() < 1e-10
        assert ev.real.min() > -1e-11 * max(1.0, np.abs(ev).max())


def test_grid_too_small_is_rejected():
    with pytest.raises(ValueError):
        sbp.diagonal_sbp_D1(4, 5)
    with pytest.raises(ValueError):
        sbp.diagonal_sbp_D1(8, 40)
