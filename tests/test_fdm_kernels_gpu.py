"""The hand-written GEMMs of the fast-diagonalisation preconditioner (K2d): tcgen05.mma kind::tf32 with TMEM accumulators
(k_tc_gemm) and mma.sync.m8n8k4.f64 (k_dgemm_batched).  z = Vr [(Vr^T R Vs) o Dinv] Vs^T is compared between the fp64 kernels,
the TF32 tensor-core kernels and -- as an independent library result -- strided-batched cuBLAS TF32 GEMMs, on blocks of the
shapes the tensor-core path serves (128 / 256 points per direction); then the PCG solves built on them are checked by an
apply -> solve round trip at config-4 block size."""
import numpy as np
import pytest

import hybridsbp_b200 as hs
from hybridsbp_b200 import synthetic

pytestmark = pytest.mark.gpu


def make_blocks(ctx, nbx, nby, Nr, Ns, p=4):
    from tests.util import warped_metrics, flat
    mets = [warped_metrics(p, Nr, Ns, bx, by, nbx, nby) for by in range(nby) for bx in range(nbx)]
    _, EToV, EToF, FToB = synthetic.block_grid_connectivity(nbx, nby)
    blk = hs.Blocks(ctx, p, [Nr] * (nbx * nby), [Ns] * (nbx * nby))
    blk.set_metrics(np.concatenate([flat(m.crr) for m in mets]), np.concatenate([flat(m.css) for m in mets]),
                    np.concatenate([flat(m.crs) for m in mets]))
    blk.set_bc(synthetic.block_bcs(EToF, FToB))
    blk.compute_tau(2.0)
    return blk


@pytest.mark.parametrize("Nr,Ns", [(255, 255), (127, 255), (255, 127)])
def test_preconditioner_kernels_agree(ctx, Nr, Ns):
    blk = make_blocks(ctx, 2, 2, Nr, Ns)
    rng = np.random.default_rng(Nr + Ns)
    r = rng.uniform(-1, 1, blk.VNp)
    dr, dz = ctx.array(r), ctx.empty(blk.VNp)
    z = {}
    variants = {0: {"fdm_gemm": 0},                                             # fp64, mma.sync f64
                3: {"fdm_gemm": 3, "fdm_tc_variant": 0},                        # default: two fused GEMM pairs (k_fdm_pair)
                "single": {"fdm_gemm": 3, "fdm_tc_variant": 1},                # four single-GEMM launches (k_tc_gemm)
                -1: {"fdm_gemm": -1}}                                           # cuBLAS, comparison only
    for key, opts in variants.items():
        for k, v in opts.items():
            blk.set_option(k, v)
        blk.local_setup(hs.LOCAL_FDM, tol=1e-13, maxit=500)
        blk.local_precondition(dr, dz)
        z[key] = dz.get()
    blk.set_option("fdm_tc_variant", 0)
    n0 = np.linalg.norm(z[0])
    assert n0 > 0
    for key in (3, "single", -1):
        # TF32 (10-bit mantissa) against fp64: four chained 256-term contractions -> a few 1e-4 relative
        assert np.all(np.isfinite(z[key])), key
        assert np.linalg.norm(z[key] - z[0]) <= 3e-3 * n0, (key, np.linalg.norm(z[key] - z[0]) / n0)
    assert np.linalg.norm(z[3] - z[-1]) <= 4e-3 * n0, np.linalg.norm(z[3] - z[-1]) / n0
    assert np.linalg.norm(z[3] - z["single"]) <= 1e-3 * n0, np.linalg.norm(z[3] - z["single"]) / n0
    print('preconditioner, relative to fp64: fused pairs %.2e, single GEMMs %.2e, cuBLAS TF32 %.2e' %
          tuple(np.linalg.norm(z[k] - z[0]) / n0 for k in (3, "single", -1)))
    # the preconditioner is symmetric positive definite: r.z > 0, and (with fp64 kernels) u.P^-1 v = v.P^-1 u
    assert r @ z[3] > 0 and r @ z[0] > 0
    v = rng.uniform(-1, 1, blk.VNp)
    blk.set_option("fdm_gemm", 0)
    blk.local_setup(hs.LOCAL_FDM, tol=1e-13, maxit=500)
    dv = ctx.array(v)
    blk.local_precondition(dv, dz)
    Pv = dz.get()
    assert abs(r @ Pv - v @ z[0]) <= 1e-11 * np.linalg.norm(r) * np.linalg.norm(Pv)
    blk.close()


@pytest.mark.parametrize("gemm", [3, 0])
def test_pcg_on_the_hand_written_kernels_round_trip(ctx, gemm):
    """4 blocks of 256 x 256 points: x = M-tilde^-1 (M-tilde x0) through FDM-PCG with the tensor-core preconditioner"""
    blk = make_blocks(ctx, 2, 2, 255, 255)
    blk.set_option("fdm_gemm", gemm)
    blk.local_setup(hs.LOCAL_FDM, tol=1e-12, maxit=2000)
    x0 = np.random.default_rng(5).uniform(-1, 1, blk.VNp)
    dx0, dg, dx, dr = ctx.array(x0), ctx.empty(blk.VNp), ctx.empty(blk.VNp), ctx.empty(blk.VNp)
    blk.apply(dx0, dg)
    st = blk.local_solve(dg, dx)
    assert st["failed_blocks"] == 0 and st["iterations_max"] < 120, st
    blk.apply(dx, dr)
    g, r = dg.get(), dr.get()
    assert np.linalg.norm(r - g) <= 1e-10 * np.linalg.norm(g), st
    assert np.linalg.norm(dx.get() - x0) <= 1e-6 * np.linalg.norm(x0), st
    if gemm == 3:                      # the library GEMMs give the same iteration count (+- rounding mode of TF32)
        blk.set_option("fdm_gemm", -1)
        blk.local_setup(hs.LOCAL_FDM, tol=1e-12, maxit=2000)
        st_lib = blk.local_solve(dg, dx)
        assert abs(st_lib["iterations_max"] - st["iterations_max"]) <= 6, (st, st_lib)
    blk.close()


def test_converged_blocks_drop_out_without_changing_the_result(ctx):
    """PCG kernels skip blocks that have converged (PcgState.active): the solution is bitwise the one of the lockstep loop.
    Block 1 has a zero right-hand side (x = 0, global_curved.jl:733), the others converge after different iteration counts."""
    blk = make_blocks(ctx, 2, 2, 255, 255)
    blk.set_option("fdm_gemm", 3)
    blk.local_setup(hs.LOCAL_FDM, tol=1e-11, maxit=2000)
    rng = np.random.default_rng(11)
    g = rng.uniform(-1, 1, blk.VNp)
    g[blk.vol_slice(1)] = 0.0
    sl = blk.vol_slice(2)
    g[sl] = 0.0
    g[sl.start + 300] = 1.0                          # a point source: a different iteration count
    dg, dx = ctx.array(g), ctx.empty(blk.VNp)
    xs, its = [], []
    for no_skip in (1, 0):
        blk.set_option("fdm_no_skip", no_skip)
        st = blk.local_solve(dg, dx)
        assert st["failed_blocks"] == 0, st
        xs.append(dx.get()); its.append((st["iterations_max"], st["iterations_sum"]))
    assert its[0] == its[1], its
    assert its[0][1] < 4 * its[0][0]                  # the blocks did stop at different iterations
    assert np.array_equal(xs[0], xs[1])
    assert np.all(xs[0][blk.vol_slice(1)] == 0.0)
    y = blk.apply_host(np.ones(blk.VNp))              # the skip flags do not leak into later applies
    assert np.all(np.isfinite(y)) and np.abs(y[blk.vol_slice(1)]).max() > 0
    blk.close()


@pytest.mark.parametrize("Nr,Ns", [(255, 127), (511, 63)])
def test_batched_jacobi_eigensolver_gives_the_library_preconditioner(ctx, Nr, Ns):
    """setup with the hand-written batched Jacobi eigensolver (k_eig.cuh) against cuSOLVER syevd (comparison knob): the fp64
    preconditioner built on either set of eigenvectors is the same operator (512 points per line: column blocks of 8)"""
    blk = make_blocks(ctx, 2, 2, Nr, Ns)
    rng = np.random.default_rng(77)
    r = rng.uniform(-1, 1, blk.VNp)
    dr, dz = ctx.array(r), ctx.empty(blk.VNp)
    z = []
    blk.set_option("fdm_gemm", 0)
    for lib in (0, 1):
        blk.set_option("fdm_eig_lib", lib)
        blk.local_setup(hs.LOCAL_FDM, tol=1e-13, maxit=500)
        blk.local_precondition(dr, dz)
        z.append(dz.get())
    blk.set_option("fdm_eig_lib", 0)
    assert np.all(np.isfinite(z[0]))
    assert np.linalg.norm(z[0] - z[1]) <= 1e-8 * np.linalg.norm(z[1]), np.linalg.norm(z[0] - z[1]) / np.linalg.norm(z[1])
    blk.close()


@pytest.mark.parametrize("N", [255, 127])
def test_p_dot_Ap_from_the_sweep_kernel(ctx, N):
    """FDM-PCG takes p . M-tilde p from k_sweep (chunk sums + the closure lines summed by the update kernel) instead of a second
    pass over the vectors (the sum itself is checked in test_apply_gpu.py::test_apply_energy...): same solution"""
    blk = make_blocks(ctx, 2, 2, N, 255)
    blk.set_option("fdm_gemm", 3)
    blk.local_setup(hs.LOCAL_FDM, tol=1e-12, maxit=2000)
    g = np.random.default_rng(3).uniform(-1, 1, blk.VNp)
    dg, dx = ctx.array(g), ctx.empty(blk.VNp)
    xs, its = [], []
    for off in (1, 0):
        blk.set_option("fdm_no_fused_dot", off)
        st = blk.local_solve(dg, dx)
        assert st["failed_blocks"] == 0, st
        xs.append(dx.get()); its.append(st["iterations_max"])
    # (the flexible PCG with a TF32 preconditioner is sensitive to the rounding of alpha: the counts differ by a few per cent)
    assert abs(its[0] - its[1]) <= 0.1 * max(its), its
    assert np.linalg.norm(xs[0] - xs[1]) <= 1e-8 * np.linalg.norm(xs[0])
    blk.close()
