"""The reference's `factorization` plugin at its own seam (hsbp_factor_*): the ASSEMBLED sparse M-tilde of a block, exactly
as the oracle's locoperator builds it (global_curved.jl:470-486), is handed over as CSC arrays -- what
`factorization(lop[e].M̃)` receives (global_curved.jl:698) -- factorised on the device and used as `F \\ g` (:734) and
`F' \\ S` (:774).  Checked against the oracle's sparse direct solve, 1e-10 relative."""
import numpy as np
import pytest
import scipy.sparse as sp
import scipy.sparse.linalg as spla

from oracle import hybrid as orc
from tests.util import random_spd_metrics, warped_metrics

pytestmark = pytest.mark.gpu


def test_one_by_one_probe(ctx):
    """SBPLocalOperator1 first calls the plugin on sparse([1], [1], [1.0]) to learn the factor type (global_curved.jl:681)"""
    import hybridsbp_b200 as hs
    F = hs.SpdFactor(ctx, sp.csc_matrix(np.array([[1.0]])))
    assert F.n == 1 and abs(F.solve(np.array([3.0]))[0] - 3.0) < 1e-15
    F2 = hs.SpdFactor(ctx, sp.csc_matrix(np.array([[4.0]])))
    assert abs(F2.solve(np.array([2.0]))[0] - 0.5) < 1e-15
    F.close(); F2.close()


@pytest.mark.parametrize("p", [2, 4, 6])
def test_factor_of_the_assembled_block_matrix(ctx, p):
    import hybridsbp_b200 as hs
    rng = np.random.default_rng(p)
    N = {2: 14, 4: 15, 6: 19}[p]
    m = random_spd_metrics(p, N, N + 2, rng, scale2=0.3)
    lop = orc.locoperator(p, N, N + 2, m, (1, 2, 0, 2))                      # Dirichlet, Neumann, interface, Neumann
    F = hs.SpdFactor(ctx, lop.Mt)
    ref = spla.splu(sp.csc_matrix(lop.Mt))
    g = rng.uniform(-1, 1, lop.Mt.shape[0])
    x, xr = F.solve(g), ref.solve(g)
    assert np.linalg.norm(x - xr) <= 1e-10 * np.linalg.norm(xr)
    # F' \ S for a block of right-hand sides: the face operator of assembleλmatrix (global_curved.jl:774)
    S = lop.F[2].toarray()
    X, Xr = F.solve(S), ref.solve(S)
    assert X.shape == S.shape and np.linalg.norm(X - Xr) <= 1e-10 * np.linalg.norm(Xr)
    F.close()


def test_factor_at_the_bp1_size_and_error_on_an_indefinite_matrix(ctx):
    """the 201 x 201-point block of seas/BP1 (BP1.jl:75-79): 40 401 unknowns, half-bandwidth 404"""
    import hybridsbp_b200 as hs
    from hybridsbp_b200 import bp1
    from hybridsbp_b200._lib import HsbpError
    su = bp1.setup(N=200)
    lop = orc.locoperator(su.p, su.N, su.N, su.metrics, su.LFtoB)
    F = hs.SpdFactor(ctx, lop.Mt)
    ref = spla.splu(sp.csc_matrix(lop.Mt))
    g = np.random.default_rng(9).uniform(-1, 1, lop.Mt.shape[0])
    x, xr = F.solve(g), ref.solve(g)
    assert np.linalg.norm(x - xr) <= 1e-10 * np.linalg.norm(xr)
    F.close()
    A = sp.diags([1.0, -1.0, 2.0]).tocsc()
    with pytest.raises(HsbpError) as e:
        hs.SpdFactor(ctx, A)
    assert "positive definite" in str(e.value)
