"""CPU emulation of the fast-diagonalisation preconditioner of hybridsbp_b200/csrc/api_fdm.cuh on the oracle's assembled
M-tilde (locoperator, global_curved.jl:211-506): the collapsed 1-D operators, the two-round eigenvalue recipe and the
resulting condition numbers.  Pins the design decisions the CUDA path relies on:
  * on a block with constant coefficients the separable operator equals M-tilde (kappa = 1);
  * collapsing against the constant alone makes lr_i + ls_j - c a difference of large numbers: the resulting operator is
    indefinite on curved blocks, which is why the eigenvalues are Rayleigh quotients against the lowest mode instead;
  * on the smoothly warped blocks of the synthetic mesh kappa(P^-1 M-tilde) stays O(1 - 10) for all boundary-condition mixes."""
import numpy as np
import pytest

from oracle import hybrid as orc
from oracle import sbp
from tests.util import warped_metrics


def _geig(A, H):
    s = 1.0 / np.sqrt(H)
    w, Q = np.linalg.eigh(0.5 * (A + A.T) * s[:, None] * s[None, :])
    return w, Q * s[:, None]


def _fdm(M, Nr, Ns, p):
    """-> (naive d_ij, two-round d_ij, P^-1 as a dense matrix from the two-round recipe)"""
    Nrp, Nsp = Nr + 1, Ns + 1
    Hr = np.array(sbp.diagonal_sbp_D1(p, Nr)[2].diagonal())
    Hs = np.array(sbp.diagonal_sbp_D1(p, Ns)[2].diagonal())
    M4 = M.reshape(Nsp, Nrp, Nsp, Nrp)                       # index (j, i, j', i'): r fastest
    Ar, As = M4.sum(axis=(0, 2)) / 2.0, M4.sum(axis=(1, 3)) / 2.0      # round 1: collapse against the constant
    lr, Vr = _geig(Ar, Hr)
    ls, Vs = _geig(As, Hs)
    d_naive = lr[:, None] + ls[None, :] - M.sum() / 4.0
    w0, v0 = Vs[:, 0], Vr[:, 0]                              # round 2: collapse against the other direction's lowest mode
    Ar2 = np.einsum("j,jikl,k->il", w0, M4, w0)
    As2 = np.einsum("i,jikl,l->jk", v0, M4, v0)
    mr = np.einsum("ia,ij,ja->a", Vr, Ar2, Vr)
    ms = np.einsum("ia,ij,ja->a", Vs, As2, Vs)
    d = mr[:, None] + ms[None, :] - 0.5 * (mr[0] + ms[0])
    d = np.maximum(d, 1e-10 * abs(mr[-1] + ms[-1]))
    VV = np.kron(Vs, Vr)
    Pinv = VV @ np.diag((1.0 / d).T.reshape(-1)) @ VV.T
    return d_naive, d, Pinv


BCS = [(1, 0, 2, 0), (0, 0, 0, 0), (2, 2, 1, 1), (2, 2, 2, 0)]


@pytest.mark.parametrize("p", [2, 4])
def test_exact_on_constant_coefficients(p):
    Nr, Ns = 23, 19
    m = warped_metrics(p, Nr, Ns, 0, 0, 2, 2, amp=0.0)
    for bc in BCS:
        M = orc.locoperator(p, Nr, Ns, m, bc).Mt.toarray()
        d_naive, d, Pinv = _fdm(M, Nr, Ns, p)
        ev = np.linalg.eigvals(Pinv @ M).real
        assert d.min() > 0 and abs(ev.min() - 1) < 1e-8 and abs(ev.max() - 1) < 1e-8, (bc, ev.min(), ev.max())


def test_two_round_eigenvalues_are_needed_and_sufficient_on_warped_blocks():
    p, Nr, Ns = 4, 31, 27
    worst = 0.0
    naive_indefinite = 0
    for bx, by, nb in ((0, 0, 2), (5, 7, 32)):               # a strongly warped block and a block of BASELINE config 4's mesh
        m = warped_metrics(p, Nr, Ns, bx, by, nb, nb)
        for bc in BCS:
            M = orc.locoperator(p, Nr, Ns, m, bc).Mt.toarray()
            d_naive, d, Pinv = _fdm(M, Nr, Ns, p)
            naive_indefinite += int(d_naive.min() <= 0)
            assert d.min() > 0
            ev = np.linalg.eigvals(Pinv @ M).real
            assert ev.min() > 0
            worst = max(worst, ev.max() / ev.min())
    assert naive_indefinite > 0          # the one-round recipe is not usable
    assert worst < 40.0, worst           # kappa(M-tilde) itself is O(N^2) ~ 1e3 - 1e4 here
