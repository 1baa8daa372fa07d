"""Size-independent properties at BASELINE.json's full sizes (config 4: 1024 blocks x 256 x 256 points, p = 4),
where the oracle cannot assemble anything: M-tilde is symmetric positive definite and linear, the two independent
CUDA paths agree, and a local solve inverts the operator (apply -> solve round trip)."""
import numpy as np
import pytest

import hybridsbp_b200 as hs
from hybridsbp_b200 import synthetic

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def c4(ctx):
    nbx = nby = 32
    N, p = 255, 4
    crr, css, crs = synthetic.warped_coefficients(nbx, nby, N)
    _, EToV, EToF, FToB = synthetic.block_grid_connectivity(nbx, nby)
    blk = hs.Blocks(ctx, p, [N] * (nbx * nby), [N] * (nbx * nby))
    blk.set_metrics(crr, css, crs)
    blk.set_bc(synthetic.block_bcs(EToF, FToB))
    blk.compute_tau(2.0)
    yield blk
    blk.close()


def test_symmetry_linearity_definiteness_at_full_size(ctx, c4):
    blk = c4
    rng = np.random.default_rng(11)
    u = rng.uniform(-1, 1, blk.VNp)
    v = rng.uniform(-1, 1, blk.VNp)
    du, dv, dy = ctx.array(u), ctx.array(v), ctx.empty(blk.VNp)
    blk.apply(du, dy); Mu = dy.get()
    assert blk.apply_variant() == 1
    blk.apply(dv, dy); Mv = dy.get()
    vMu, uMv = float(v @ Mu), float(u @ Mv)
    scale = np.linalg.norm(u) * np.linalg.norm(Mv)
    assert abs(vMu - uMv) <= 1e-12 * scale                       # M-tilde = M-tilde^T
    assert float(u @ Mu) > 0 and float(v @ Mv) > 0               # positive definite
    a, b = 0.37, -1.9
    dw = ctx.array(a * u + b * v)
    blk.apply(dw, dy)
    lin = dy.get()
    assert np.max(np.abs(lin - (a * Mu + b * Mv))) <= 1e-12 * np.max(np.abs(Mu))     # linear
    # constants: the volume operator annihilates them, only the Dirichlet-type face penalties remain
    done = ctx.array(np.ones(blk.VNp))
    blk.apply(done, dy)
    M1 = dy.get().reshape(1024, 256, 256)
    assert np.max(np.abs(M1[:, 8:-8, 8:-8])) <= 1e-9 * np.max(np.abs(Mu))


def test_apply_then_local_solve_round_trip_at_full_block_size(ctx):
    """4 blocks of 256 x 256 points: x = M-tilde^-1 (M-tilde x0) through the batched PCG (K2b) on top of k_sweep"""
    nbx, nby, N, p = 2, 2, 255, 4
    crr, css, crs = synthetic.warped_coefficients(nbx, nby, N)
    _, EToV, EToF, FToB = synthetic.block_grid_connectivity(nbx, nby)
    blk = hs.Blocks(ctx, p, [N] * 4, [N] * 4)
    blk.set_metrics(crr, css, crs)
    blk.set_bc(synthetic.block_bcs(EToF, FToB))
    blk.compute_tau(2.0)
    blk.local_setup(hs.LOCAL_PCG, tol=1e-12, maxit=100000)
    x0 = np.random.default_rng(5).uniform(-1, 1, blk.VNp)
    dx0, dg, dx = ctx.array(x0), ctx.empty(blk.VNp), ctx.empty(blk.VNp)
    blk.apply(dx0, dg)
    st = blk.local_solve(dg, dx)
    assert st["failed_blocks"] == 0, st
    # residual-based check (the error itself is amplified by cond(M-tilde) = O(N^2))
    dr = ctx.empty(blk.VNp)
    blk.apply(dx, dr)
    g, r = dg.get(), dr.get()
    assert np.linalg.norm(r - g) <= 1e-10 * np.linalg.norm(g), st
    assert np.linalg.norm(dx.get() - x0) <= 1e-6 * np.linalg.norm(x0), st
    blk.close()


def test_condensed_trace_solve_on_large_blocks(ctx):
    """2 x 2 blocks of 128 x 128 points, FDM-PCG local solver (TF32 preconditioner GEMMs), static condensation and the
    face-block preconditioner: the solution must satisfy the coupled system [M Fbar; Fbar^T D][u; lam] = [g; gd]
    evaluated with the matrix-free operators, and agree with the matrix-free Jacobi-preconditioned solve."""
    from hybridsbp_b200.host import connectivityarrays
    nbx, nby, N, p = 2, 2, 127, 4
    crr, css, crs = synthetic.warped_coefficients(nbx, nby, N)
    _, EToV, EToF, FToB = synthetic.block_grid_connectivity(nbx, nby)
    FToE, FToLF, EToO, EToS = connectivityarrays(EToV, EToF)
    blk = hs.Blocks(ctx, p, [N] * 4, [N] * 4)
    blk.set_metrics(crr, css, crs)
    blk.set_bc(synthetic.block_bcs(EToF, FToB))
    blk.compute_tau(2.0)
    blk.local_setup(hs.LOCAL_FDM, tol=1e-13, maxit=5000)
    tr = hs.Trace(blk, FToB, FToE, FToLF, EToO, EToS)
    rng = np.random.default_rng(11)
    g, gd = rng.uniform(-1, 1, blk.VNp), rng.uniform(-1, 1, tr.lNp)
    dg, dgd = ctx.array(g), ctx.array(gd)
    lam0, u0 = ctx.empty(tr.lNp), ctx.empty(blk.VNp)
    st0 = tr.solve(dg, dgd, lam0, u0, tol=1e-11, maxit=5000)           # matrix-free, Jacobi
    assert st0["converged"] == 1, st0
    tr.condense()
    tr.precond_setup(1)
    lam1, u1 = ctx.empty(tr.lNp), ctx.empty(blk.VNp)
    st1 = tr.solve(dg, dgd, lam1, u1, tol=1e-11, maxit=5000)
    assert st1["converged"] == 1 and st1["local_solves"] == 2 and st1["outer_iterations"] < st0["outer_iterations"], (st0, st1)
    l0, l1, v0, v1 = lam0.get(), lam1.get(), u0.get(), u1.get()
    assert np.linalg.norm(l1 - l0) <= 1e-8 * np.linalg.norm(l0), (st0, st1)
    assert np.linalg.norm(v1 - v0) <= 1e-8 * np.linalg.norm(v0), (st0, st1)
    Mu, Fl, FTu = ctx.empty(blk.VNp), ctx.array(np.zeros(blk.VNp)), ctx.empty(tr.lNp)
    blk.apply(u1, Mu); tr.Fbar_add(lam1, 1.0, Fl); tr.FbarT(u1, FTu)
    assert np.linalg.norm(g - Mu.get() - Fl.get()) <= 1e-9 * np.linalg.norm(g)
    assert np.linalg.norm(gd - FTu.get() - tr.D() * l1) <= 1e-7 * np.linalg.norm(gd)
    tr.close(); blk.close()
