"""GPU parity: matrix-free M-tilde u (CUDA, through the C-ABI) against the oracle's assembled
sparse M-tilde (reference locoperator, global_curved.jl:211-506).  Tolerance from the north
star: 1e-12 relative, normwise (||dy||_inf / || |M| |u| ||_inf)."""
import numpy as np
import pytest

from oracle import hybrid as orc
from tests.util import flat, random_spd_metrics, rel_err_apply, upload_blocks, warped_metrics

pytestmark = pytest.mark.gpu
TOL = 1e-12

BCS = [(1, 1, 1, 1), (0, 0, 0, 0), (1, 2, 2, 2), (2, 1, 0, 7), (2, 2, 2, 1), (7, 0, 2, 1)]


@pytest.mark.parametrize("p", [2, 4, 6])
@pytest.mark.parametrize("generic", [True, False])
def test_apply_random_spd_small(ctx, p, generic):
    import hybridsbp_b200 as hs
    rng = np.random.default_rng(777 + p)
    N = 3 * p - 1 if p > 2 else 5
    shapes = [(N, N), (N + 3, N + 1), (N + 8, N + 13), (2 * N + 1, N + 2)] * 2
    mets, bcs, lops = [], [], []
    for i, (Nr, Ns) in enumerate(shapes):
        m = random_spd_metrics(p, Nr, Ns, rng)
        bc = BCS[i % len(BCS)]
        mets.append(m); bcs.append(bc)
        lops.append(orc.locoperator(p, Nr, Ns, m, bc, tauscale=1.0))
    blk = upload_blocks(hs, ctx, p, mets, bcs, tauscale=1.0)
    blk.force_generic(generic)
    u = rng.uniform(-1, 1, blk.VNp)
    du, dy = ctx.array(u), ctx.empty(blk.VNp)
    blk.apply(du, dy)
    y = dy.get()
    # tau parity first (global_curved.jl:418-437)
    tau = blk.get_tau()
    for e, lop in enumerate(lops):
        for lf in range(1, 5):
            tref = lop.tau[lf - 1].diagonal()
            assert np.allclose(tau[blk.face_slice(e, lf)], tref, rtol=1e-13, atol=0)
    for e, lop in enumerate(lops):
        sl = blk.vol_slice(e)
        yref = lop.Mt @ u[sl]
        err = rel_err_apply(y[sl], yref, lop.Mt, u[sl])
        assert err < TOL, (p, e, shapes[e], bcs[e], err)


@pytest.mark.parametrize("p,N", [(2, 40), (4, 31), (4, 63), (6, 47)])
def test_apply_warped_mesh(ctx, p, N):
    import hybridsbp_b200 as hs
    rng = np.random.default_rng(p * 100 + N)
    nbx = nby = 2
    mets, bcs, lops = [], [], []
    for by in range(nby):
        for bx in range(nbx):
            m = warped_metrics(p, N, N, bx, by, nbx, nby, amp=0.15)
            bc = (1 if bx == 0 else 0, 1 if bx == nbx - 1 else 0, 2 if by == 0 else 0, 2 if by == nby - 1 else 0)
            mets.append(m); bcs.append(bc)
            lops.append(orc.locoperator(p, N, N, m, bc))
    blk = upload_blocks(hs, ctx, p, mets, bcs)
    u = rng.uniform(-1, 1, blk.VNp)
    y = blk.apply_host(u)
    for e, lop in enumerate(lops):
        sl = blk.vol_slice(e)
        err = rel_err_apply(y[sl], lop.Mt @ u[sl], lop.Mt, u[sl])
        assert err < TOL, (p, e, err)


@pytest.mark.parametrize("p", [2, 4, 6])
def test_apply_against_extended_precision(ctx, p):
    """the CUDA result and the oracle's fp64 SpMV both against an 80-bit (numpy.longdouble) evaluation of
    M~ u on small blocks (marching-kernel and generic-kernel sizes), in the normwise measure of the parity bar"""
    import hybridsbp_b200 as hs
    assert np.finfo(np.longdouble).nmant >= 63
    rng = np.random.default_rng(4242 + p)
    for N, generic in ((3 * p + 2, True), (34, False)):
        m = random_spd_metrics(p, N, N, rng)
        bc = (1, 2, 0, 7)
        lop = orc.locoperator(p, N, N, m, bc)
        blk = upload_blocks(hs, ctx, p, [m], [bc])
        blk.force_generic(generic)
        u = rng.uniform(-1, 1, blk.VNp)
        y = blk.apply_host(u)
        yld = lop.Mt.toarray().astype(np.longdouble) @ u.astype(np.longdouble)
        scale = np.max(abs(lop.Mt) @ np.abs(u))
        assert float(np.max(np.abs(y - yld))) / scale < TOL
        assert float(np.max(np.abs(lop.Mt @ u - yld))) / scale < 1e-14
        blk.close()


@pytest.mark.parametrize("p", [2, 4, 6])
def test_face_operators(ctx, p):
    """F_k^T u, F_k v and the traction operator against the oracle's sparse F_k / HfI_FT_k."""
    import hybridsbp_b200 as hs
    rng = np.random.default_rng(5 + p)
    N = 3 * p + 2
    shapes = [(N, N + 4), (N + 5, N)]
    mets = [random_spd_metrics(p, a, b, rng) for a, b in shapes]
    bcs = [(0, 7, 1, 2), (2, 0, 0, 1)]
    lops = [orc.locoperator(p, a, b, m, bc, tauscale=1.5) for (a, b), m, bc in zip(shapes, mets, bcs)]
    blk = upload_blocks(hs, ctx, p, mets, bcs, tauscale=1.5)
    u = rng.uniform(-1, 1, blk.VNp)
    v = rng.uniform(-1, 1, blk.FNp)
    du, dv = ctx.array(u), ctx.array(v)
    dft, dtr, dy = ctx.empty(blk.FNp), ctx.empty(blk.FNp), ctx.array(np.zeros(blk.VNp))
    blk.face_FT(du, dft); blk.face_traction(du, dtr); blk.face_F_add(dv, -0.5, dy)
    ft, tr, y = dft.get(), dtr.get(), dy.get()
    for e, lop in enumerate(lops):
        sl = blk.vol_slice(e)
        yref = np.zeros(sl.stop - sl.start)
        for lf in range(1, 5):
            fs = blk.face_slice(e, lf)
            F = lop.F[lf - 1]
            ref = F.T @ u[sl]
            assert np.max(np.abs(ft[fs] - ref)) <= 1e-12 * np.max(abs(F.T) @ np.abs(u[sl]))
            T = lop.HfI_FT[lf - 1]
            assert np.max(np.abs(tr[fs] - T @ u[sl])) <= 1e-12 * np.max(abs(T) @ np.abs(u[sl]))
            yref += -0.5 * (F @ v[fs])
        assert np.max(np.abs(y[sl] - yref)) <= 1e-12 * np.max(np.abs(yref))


SWEEP_CASES = [
    # p, Nr, Ns, points per thread, chunks per side
    (2, 31, 31, 0, 0), (2, 63, 40, 2, 2), (2, 127, 95, 4, 0),
    (4, 31, 40, 0, 1), (4, 63, 33, 2, 0), (4, 127, 64, 4, 2), (4, 127, 64, 2, 1), (4, 255, 37, 0, 0),
    (4, 259, 50, 4, 0), (4, 255, 255, 0, 0),
    (6, 47, 47, 0, 0), (6, 63, 80, 2, 2), (6, 127, 40, 4, 1), (6, 131, 64, 4, 2),
    # odd numbers of points per line (N = 34, 68, 136 of square_circle.jl / flower, N = 200 of BP1): pitched copies, phantom point
    (2, 34, 40, 0, 0), (2, 200, 33, 0, 2), (4, 34, 34, 0, 0), (4, 68, 45, 0, 2), (4, 136, 60, 0, 0), (4, 200, 40, 0, 1),
    (6, 34, 34, 0, 0), (6, 68, 50, 0, 0), (6, 200, 47, 0, 2),
]


@pytest.mark.parametrize("p,Nr,Ns,R,ncs", SWEEP_CASES)
def test_apply_marching_kernel_vs_oracle(ctx, p, Nr, Ns, R, ncs):
    """uniform blocks with >= 32 points per direction take the TMA line-marching kernel (k_sweep): every closure, both
    marching directions, chunk seams, 2 and 4 points per thread, even and odd line lengths"""
    import hybridsbp_b200 as hs
    rng = np.random.default_rng(1000 * p + Nr + Ns)
    nb = 3 if Nr * Ns < 40000 else 2
    mets = [random_spd_metrics(p, Nr, Ns, rng, scale2=0.05) for _ in range(nb)]
    bcs = [BCS[(i + p) % len(BCS)] for i in range(nb)]
    blk = upload_blocks(hs, ctx, p, mets, bcs)
    blk.set_option("sweep_points_per_thread", R)
    blk.set_option("sweep_chunks_per_side", ncs)
    u = rng.uniform(-1, 1, blk.VNp)
    du, dy = ctx.array(u), ctx.empty(blk.VNp)
    dy.set(np.full(blk.VNp, np.nan))
    blk.apply(du, dy)
    assert blk.apply_variant() == 1
    y = dy.get()
    for e in range(nb):
        lop = orc.locoperator(p, Nr, Ns, mets[e], bcs[e])
        sl = blk.vol_slice(e)
        err = rel_err_apply(y[sl], lop.Mt @ u[sl], lop.Mt, u[sl])
        assert err < TOL, (p, e, err)


@pytest.mark.parametrize("p", [2, 4, 6])
def test_apply_marching_equals_generic_at_full_block_size(ctx, p):
    """256 x 256-point blocks (BASELINE config 4 shape): the two independent CUDA paths must agree
    to rounding; the generic one is checked against the oracle at sizes the oracle can assemble."""
    import hybridsbp_b200 as hs
    from hybridsbp_b200 import synthetic
    nbx, nby, N = 4, 2, 255
    crr, css, crs = synthetic.warped_coefficients(nbx, nby, N)
    _, EToV, EToF, FToB = synthetic.block_grid_connectivity(nbx, nby)
    blk = hs.Blocks(ctx, p, [N] * (nbx * nby), [N] * (nbx * nby))
    blk.set_metrics(crr, css, crs)
    blk.set_bc(synthetic.block_bcs(EToF, FToB))
    blk.compute_tau(2.0)
    u = np.random.default_rng(3).uniform(-1, 1, blk.VNp)
    du, dy = ctx.array(u), ctx.empty(blk.VNp)
    blk.force_generic(True)
    blk.apply(du, dy)
    assert blk.apply_variant() == 0
    y0 = dy.get()
    blk.force_generic(False)
    for R, ncs, deep in ((0, 0, 1), (2, 3, 1), (4, 1, 1), (4, 5, 1), (4, 2, 0), (2, 2, 0)):
        blk.set_option("sweep_points_per_thread", R)
        blk.set_option("sweep_chunks_per_side", ncs)
        blk.set_option("sweep_deep", deep)           # css / crs windows in shared-memory rings or in registers
        dy.set(np.full(blk.VNp, np.nan))
        blk.apply(du, dy)
        assert blk.apply_variant() == 1
        y1 = dy.get()
        assert np.max(np.abs(y1 - y0)) <= 1e-12 * np.max(np.abs(y0)), (R, ncs, deep)


def test_compute_tau_rejects_an_indefinite_tensor_anywhere_in_the_block(ctx):
    """global_curved.jl:418-419 asserts psi_min > 0 over the whole block, not only in the layers next to the faces"""
    import hybridsbp_b200 as hs
    p, N = 4, 20
    rng = np.random.default_rng(9)
    m = random_spd_metrics(p, N, N, rng)
    crr, css, crs = flat(m.crr).copy(), flat(m.css).copy(), flat(m.crs).copy()
    blk = hs.Blocks(ctx, p, [N], [N])
    blk.set_metrics(crr, css, crs)
    blk.set_bc([1, 1, 1, 1])
    blk.compute_tau(1.0)                                   # fine
    mid = (N + 1) * (N // 2) + N // 2                      # a point far from every face
    crs[mid] = 10.0 * max(crr[mid], css[mid])              # indefinite there
    blk.set_metrics(crr, css, crs)
    with pytest.raises(hs.HsbpError):
        blk.compute_tau(1.0)
    blk.close()


@pytest.mark.parametrize("p,Nr,Ns,ncs", [(4, 63, 70, 0), (4, 255, 96, 3), (2, 40, 33, 1), (6, 95, 64, 2), (4, 200, 50, 0), (6, 68, 40, 0)])
def test_apply_energy_from_the_sweep_kernel(ctx, p, Nr, Ns, ncs):
    """hsbp_apply_energy: u_e . (M-tilde u)_e per block accumulated inside k_sweep (chunk sums, closure lines added by a small
    kernel) equals the dot product of the vectors; y is the plain apply's y bit for bit"""
    import hybridsbp_b200 as hs
    rng = np.random.default_rng(31 * p + Nr)
    nb = 3
    mets = [random_spd_metrics(p, Nr, Ns, rng, scale2=0.05) for _ in range(nb)]
    bcs = [BCS[(i + 2 * p) % len(BCS)] for i in range(nb)]
    blk = upload_blocks(hs, ctx, p, mets, bcs)
    blk.set_option("sweep_chunks_per_side", ncs)
    u = rng.uniform(-1, 1, blk.VNp)
    du, dy, dy2 = ctx.array(u), ctx.empty(blk.VNp), ctx.empty(blk.VNp)
    blk.apply(du, dy)
    assert blk.apply_variant() == 1
    en = blk.apply_energy(du, dy2)
    y, y2 = dy.get(), dy2.get()
    assert np.array_equal(y, y2)
    for e in range(nb):
        sl = blk.vol_slice(e)
        ref = u[sl] @ y[sl]
        assert abs(en[e] - ref) <= 1e-12 * np.abs(u[sl] * y[sl]).sum(), (e, en[e], ref)
