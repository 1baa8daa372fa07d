"""GPU parity of the local solves, the trace operators and the Schur-complement solve against the
oracle's assembled sparse path (global_curved.jl:510-565, 730-797; square_circle.jl:376-388).
Tolerance from the north star: lambda and u within 1e-10 relative (2-norm) at equal CG tolerance."""
import numpy as np
import pytest
import scipy.sparse as sp

from oracle import hybrid as orc
from tests.util import flat, random_spd_metrics, upload_blocks, warped_metrics

pytestmark = pytest.mark.gpu


def two_block_mesh():
    """the hand-made connectivity of global_op_eigenvalues.jl:12-19"""
    EToV = np.array([[1, 2, 4, 5], [2, 3, 5, 6]]).T
    EToF = np.array([[1, 2, 3, 4], [2, 5, 6, 7]]).T
    FToB = np.full(7, orc.BC_DIRICHLET)
    FToB[1] = orc.BC_LOCKED_INTERFACE
    return EToV, EToF, FToB


def flipped_four_block_mesh():
    """2 x 2 blocks around a centre vertex; blocks 2 and 3 are rotated so that some interface
    faces meet with opposite orientation (EToO == false, global_curved.jl:544-552)."""
    # vertices 1..9 on a 3x3 grid, row-major from bottom-left
    v = lambda ix, iy: 1 + ix + 3 * iy
    blocks = [
        (v(0, 0), v(1, 0), v(0, 1), v(1, 1)),            # standard
        (v(2, 1), v(2, 0), v(1, 1), v(1, 0)),            # rotated: r runs downwards
        (v(1, 2), v(0, 2), v(1, 1), v(0, 1)),            # mirrored twice
        (v(1, 1), v(2, 1), v(1, 2), v(2, 2)),            # standard
    ]
    EToV = np.array(blocks).T
    from hybridsbp_b200.host import _FACE_VERTS
    EToF = np.zeros((4, 4), dtype=np.int64)
    known = {}
    for e in range(4):
        for lf in range(4):
            a, b = EToV[_FACE_VERTS[lf], e]
            EToF[lf, e] = known.setdefault((min(a, b), max(a, b)), len(known) + 1)
    FToB = np.zeros(len(known), dtype=np.int64)
    count = np.zeros(len(known), dtype=int)
    for e in range(4):
        for lf in range(4):
            count[EToF[lf, e] - 1] += 1
    k = 0
    for f in range(len(known)):
        if count[f] == 1:
            FToB[f] = orc.BC_DIRICHLET if k % 2 == 0 else orc.BC_NEUMANN
            k += 1
    FToB[np.where(count == 2)[0][0]] = orc.BC_JUMP_INTERFACE
    return EToV, EToF, FToB


def build_case(hs, ctx, p, N, mesh, rng, tauscale=1.0):
    EToV, EToF, FToB = mesh
    ne = EToV.shape[1]
    FToE, FToLF, EToO, EToS = orc.connectivityarrays(EToV, EToF)
    mets = [random_spd_metrics(p, N, N, rng, scale2=0.2) for _ in range(ne)]
    bcs = [[FToB[f - 1] for f in EToF[:, e]] for e in range(ne)]
    lops = [orc.locoperator(p, N, N, m, bc, tauscale=tauscale) for m, bc in zip(mets, bcs)]
    Nr = [N] * ne
    M, FbarT, D, vstarts, Fl = orc.LocalGlobalOperators(lops, Nr, Nr, FToB, FToE, FToLF, EToO, EToS)
    blk = upload_blocks(hs, ctx, p, mets, bcs, tauscale=tauscale)
    tr = hs.Trace(blk, FToB, FToE, FToLF, EToO, EToS)
    return dict(lops=lops, M=M, FbarT=FbarT, D=D, vstarts=vstarts, Fl=Fl, blk=blk, tr=tr, EToF=EToF, FToB=FToB)


@pytest.mark.parametrize("p", [2, 4, 6])
@pytest.mark.parametrize("meshname", ["two", "flipped"])
def test_trace_operators_and_solve(ctx, p, meshname):
    import hybridsbp_b200 as hs
    rng = np.random.default_rng(31 * p + len(meshname))
    N = 3 * p - 1 if p > 2 else 6
    mesh = two_block_mesh() if meshname == "two" else flipped_four_block_mesh()
    c = build_case(hs, ctx, p, N, mesh, rng)
    blk, tr, FbarT = c["blk"], c["tr"], c["FbarT"]
    if meshname == "flipped":
        _, _, EToO, _ = orc.connectivityarrays(mesh[0], mesh[1])
        assert (~EToO).any(), "this mesh must exercise reversed faces"
    assert np.array_equal(tr.FTolambdastarts, c["Fl"])
    assert np.allclose(tr.D(), c["D"], rtol=1e-13)
    u = rng.uniform(-1, 1, blk.VNp)
    lam = rng.uniform(-1, 1, tr.lNp)
    du, dl = ctx.array(u), ctx.array(lam)
    dlo, dy = ctx.empty(tr.lNp), ctx.array(np.zeros(blk.VNp))
    tr.FbarT(du, dlo)
    ref = FbarT @ u
    assert np.max(np.abs(dlo.get() - ref)) <= 1e-12 * np.max(abs(FbarT) @ np.abs(u))
    tr.Fbar_add(dl, 1.0, dy)
    ref = FbarT.T @ lam
    assert np.max(np.abs(dy.get() - ref)) <= 1e-12 * np.max(abs(FbarT.T) @ np.abs(lam))

    # local solves
    blk.local_setup(hs.LOCAL_PCG, tol=1e-14, maxit=20000)
    g = rng.uniform(-1, 1, blk.VNp)
    dg, dx = ctx.array(g), ctx.empty(blk.VNp)
    st = blk.local_solve(dg, dx)
    assert st["failed_blocks"] == 0, st
    x = dx.get()
    for e, lop in enumerate(c["lops"]):
        sl = blk.vol_slice(e)
        xr = c["M"].F[e].solve(g[sl])
        assert np.linalg.norm(x[sl] - xr) <= 1e-10 * np.linalg.norm(xr), (e, st)

    # Schur complement apply against the oracle's explicit B
    B = orc.assemblelambdamatrix(c["Fl"], c["vstarts"], c["EToF"], c["FToB"], c["M"].F, c["D"], FbarT)
    tr.schur_apply(dl, dlo)
    ref = B @ lam
    assert np.linalg.norm(dlo.get() - ref) <= 1e-10 * np.linalg.norm(ref)

    # full trace solve: lambda = B^-1 (gd - Fbar^T M^-1 g), u = M^-1 (g - Fbar lambda)
    gd = rng.uniform(-1, 1, tr.lNp)
    bl = np.zeros(tr.lNp); uu = np.zeros(blk.VNp)
    orc.LocalToGLobalRHS(bl, g, gd, uu, c["M"].F, FbarT, c["vstarts"])
    lam_ref = np.linalg.solve(B.toarray(), bl)
    rhs = g - FbarT.T @ lam_ref
    u_ref = np.concatenate([c["M"].F[e].solve(rhs[blk.vol_slice(e)]) for e in range(blk.nblocks)])
    dgd, dlam, dsol = ctx.array(gd), ctx.empty(tr.lNp), ctx.empty(blk.VNp)
    st = tr.solve(dg, dgd, dlam, dsol, tol=1e-13, maxit=2000)
    assert st["converged"] == 1, st
    assert np.linalg.norm(dlam.get() - lam_ref) <= 1e-10 * np.linalg.norm(lam_ref), st
    assert np.linalg.norm(dsol.get() - u_ref) <= 1e-10 * np.linalg.norm(u_ref), st


@pytest.mark.parametrize("p", [2, 4, 6])
def test_dense_cholesky_local_solver(ctx, p):
    """K2a: the `factorization` plugin as a batched dense Cholesky (global_curved.jl:698, 734) on small,
    differently shaped blocks; and the trace solve on top of it."""
    import hybridsbp_b200 as hs
    rng = np.random.default_rng(400 + p)
    N = 3 * p - 1 if p > 2 else 6
    shapes = [(N, N), (N + 2, N + 5), (2 * N, N + 1), (N + 7, N)]
    mets = [random_spd_metrics(p, a, b, rng, scale2=0.2) for a, b in shapes]
    bcs = [(1, 1, 1, 1), (1, 2, 2, 2), (0, 1, 2, 7), (2, 0, 1, 1)]
    lops = [orc.locoperator(p, a, b, m, bc) for (a, b), m, bc in zip(shapes, mets, bcs)]
    blk = upload_blocks(hs, ctx, p, mets, bcs)
    blk.local_setup(hs.LOCAL_CHOLESKY)
    g = rng.uniform(-1, 1, blk.VNp)
    dg, dx = ctx.array(g), ctx.empty(blk.VNp)
    st = blk.local_solve(dg, dx)
    assert st["failed_blocks"] == 0 and st["iterations_max"] == 0
    x = dx.get()
    for e, lop in enumerate(lops):
        sl = blk.vol_slice(e)
        xr = orc.default_factorization(lop.Mt).solve(g[sl])
        assert np.linalg.norm(x[sl] - xr) <= 1e-10 * np.linalg.norm(xr), e


@pytest.mark.parametrize("p,stream", [(2, 1), (4, 1), (6, 1), (4, 0)])
def test_banded_cholesky_local_solver(ctx, p, stream):
    """K2c: the `factorization` plugin as a batched banded Cholesky (global_curved.jl:698, 734): blocks of different
    shapes, all boundary-condition types (Neumann faces have the widest coupling), sizes beyond the dense solver."""
    import hybridsbp_b200 as hs
    rng = np.random.default_rng(500 + p)
    N = 3 * p - 1 if p > 2 else 6
    shapes = [(N, N), (N + 2, N + 5), (40, N + 1), (N + 7, 33), (47, 52)]
    mets = [random_spd_metrics(p, a, b, rng, scale2=0.2) for a, b in shapes]
    bcs = [(1, 1, 1, 1), (1, 2, 2, 2), (0, 1, 2, 7), (2, 0, 1, 1), (2, 2, 1, 2)]
    lops = [orc.locoperator(p, a, b, m, bc) for (a, b), m, bc in zip(shapes, mets, bcs)]
    blk = upload_blocks(hs, ctx, p, mets, bcs)
    blk.set_option("band_no_stream", 0 if stream else 1)      # streamed (TMA ring) or plain solve kernel
    blk.local_setup(hs.LOCAL_BAND)
    g = rng.uniform(-1, 1, blk.VNp)
    dg, dx = ctx.array(g), ctx.empty(blk.VNp)
    st = blk.local_solve(dg, dx)
    assert st["failed_blocks"] == 0 and st["iterations_max"] == 0
    x = dx.get()
    for e, lop in enumerate(lops):
        sl = blk.vol_slice(e)
        xr = orc.default_factorization(lop.Mt).solve(g[sl])
        assert np.linalg.norm(x[sl] - xr) <= 1e-10 * np.linalg.norm(xr), (e, np.linalg.norm(x[sl] - xr) / np.linalg.norm(xr))


@pytest.mark.parametrize("p,Nr,Ns", [(6, 110, 20), (6, 215, 19), (4, 180, 17)])
def test_banded_cholesky_wide_bands_stream_with_narrow_panels(ctx, p, Nr, Ns):
    """wide bands (long lines, p = 6: half-bandwidth 8 (Nr+1) + 8): sixteen columns per shared-memory stage would leave room for one
    stage only, the streamed solve then runs on panels of 8 or 4 columns (config 1 / 2 at N = 136, p = 6 took the plain-load kernel
    before: 22 s of condensation instead of 2.7 s)"""
    import hybridsbp_b200 as hs
    rng = np.random.default_rng(900 + Nr)
    mets = [random_spd_metrics(p, Nr, Ns, rng, scale2=0.2) for _ in range(2)]
    bcs = [(1, 2, 0, 7), (2, 1, 1, 0)]
    lops = [orc.locoperator(p, Nr, Ns, m, bc) for m, bc in zip(mets, bcs)]
    blk = upload_blocks(hs, ctx, p, mets, bcs)
    g = rng.uniform(-1, 1, blk.VNp)
    dg, dx = ctx.array(g), ctx.empty(blk.VNp)
    xs = []
    for no_stream in (0, 1):
        blk.set_option("band_no_stream", no_stream)
        blk.local_setup(hs.LOCAL_BAND)
        st = blk.local_solve(dg, dx)
        assert st["failed_blocks"] == 0
        xs.append(dx.get())
    for e, lop in enumerate(lops):
        sl = blk.vol_slice(e)
        xr = orc.default_factorization(lop.Mt).solve(g[sl])
        for x in xs:
            assert np.linalg.norm(x[sl] - xr) <= 1e-10 * np.linalg.norm(xr), (e, np.linalg.norm(x[sl] - xr) / np.linalg.norm(xr))


@pytest.mark.parametrize("p,kind,gemm", [(2, "warped", 0), (4, "warped", 0), (6, "warped", 0), (4, "random", 0),
                                         (4, "warped", 1), (4, "warped", 2), (4, "warped", 3)])
def test_fast_diagonalisation_pcg_local_solver(ctx, p, kind, gemm):
    """K2d: PCG on M-tilde_e preconditioned by the inverse of its separable part (api_fdm.cuh) against the oracle's
    direct solve (global_curved.jl:698, 734); on smoothly warped blocks it must need far fewer iterations than
    Jacobi-PCG, on the random SPD tensors of local_op_eigenvalues.jl:32-38 it only has to stay correct."""
    import hybridsbp_b200 as hs
    from tests.util import warped_metrics
    rng = np.random.default_rng(600 + p)
    Nr, Ns = 39, 33
    if kind == "warped":
        mets = [warped_metrics(p, Nr, Ns, bx, by, 2, 2) for by in range(2) for bx in range(2)]
    else:
        mets = [random_spd_metrics(p, Nr, Ns, rng, scale2=0.2) for _ in range(4)]
    bcs = [(1, 0, 2, 0), (0, 1, 2, 0), (1, 0, 0, 2), (0, 7, 0, 1)]
    lops = [orc.locoperator(p, Nr, Ns, m, bc) for m, bc in zip(mets, bcs)]
    blk = upload_blocks(hs, ctx, p, mets, bcs)
    g = rng.uniform(-1, 1, blk.VNp)
    dg, dx = ctx.array(g), ctx.empty(blk.VNp)
    blk.set_option("fdm_gemm", gemm)      # preconditioner GEMMs: 0 fp64, 1 fp32, 2 fp32 emulated on BF16 tensor cores, 3 TF32
    blk.local_setup(hs.LOCAL_FDM, tol=1e-13, maxit=5000)
    st = blk.local_solve(dg, dx)
    assert st["failed_blocks"] == 0, st
    x = dx.get()
    for e, lop in enumerate(lops):
        sl = blk.vol_slice(e)
        xr = orc.default_factorization(lop.Mt).solve(g[sl])
        assert np.linalg.norm(x[sl] - xr) <= 1e-10 * np.linalg.norm(xr), (e, st)
    if kind == "warped":
        blk.local_setup(hs.LOCAL_PCG, tol=1e-13, maxit=20000)
        st_j = blk.local_solve(dg, dx)
        assert st["iterations_max"] * 2 < st_j["iterations_max"], (st, st_j)
        print("FDM-PCG %d iterations, Jacobi-PCG %d" % (st["iterations_max"], st_j["iterations_max"]))


def test_trace_solve_with_cholesky_local_solver(ctx):
    import hybridsbp_b200 as hs
    p = 4
    rng = np.random.default_rng(77)
    c = build_case(hs, ctx, p, 3 * p - 1, flipped_four_block_mesh(), rng)
    blk, tr, FbarT = c["blk"], c["tr"], c["FbarT"]
    blk.local_setup(hs.LOCAL_CHOLESKY)
    g = rng.uniform(-1, 1, blk.VNp); gd = rng.uniform(-1, 1, tr.lNp)
    B = orc.assemblelambdamatrix(c["Fl"], c["vstarts"], c["EToF"], c["FToB"], c["M"].F, c["D"], FbarT)
    bl = np.zeros(tr.lNp); uu = np.zeros(blk.VNp)
    orc.LocalToGLobalRHS(bl, g, gd, uu, c["M"].F, FbarT, c["vstarts"])
    lam_ref = np.linalg.solve(B.toarray(), bl)
    rhs = g - FbarT.T @ lam_ref
    u_ref = np.concatenate([c["M"].F[e].solve(rhs[blk.vol_slice(e)]) for e in range(blk.nblocks)])
    dg, dgd, dlam, dsol = ctx.array(g), ctx.array(gd), ctx.empty(tr.lNp), ctx.empty(blk.VNp)
    st = tr.solve(dg, dgd, dlam, dsol, tol=1e-13, maxit=2000)
    assert st["converged"] == 1 and st["inner_iterations_sum"] == 0, st
    assert np.linalg.norm(dlam.get() - lam_ref) <= 1e-10 * np.linalg.norm(lam_ref), st
    assert np.linalg.norm(dsol.get() - u_ref) <= 1e-10 * np.linalg.norm(u_ref), st


def test_condensed_trace_solve_matches_matrix_free_and_oracle(ctx):
    """Static condensation (hsbp_trace_condense): the dense S_e = F_e^T M̃_e^-1 F_e reproduce the oracle's assembled
    B = D - Fbar^T M̃^-1 Fbar (assembleλmatrix, global_curved.jl:743-797) and the solve agrees with the direct one."""
    import hybridsbp_b200 as hs
    p = 4
    rng = np.random.default_rng(78)
    c = build_case(hs, ctx, p, 3 * p - 1, flipped_four_block_mesh(), rng)
    blk, tr, FbarT = c["blk"], c["tr"], c["FbarT"]
    blk.local_setup(hs.LOCAL_CHOLESKY)
    B = orc.assemblelambdamatrix(c["Fl"], c["vstarts"], c["EToF"], c["FToB"], c["M"].F, c["D"], FbarT).toarray()
    lam = rng.uniform(-1, 1, tr.lNp)
    dl, dq0, dq1 = ctx.array(lam), ctx.empty(tr.lNp), ctx.empty(tr.lNp)
    tr.schur_apply(dl, dq0)
    tr.condense()
    tr.schur_apply(dl, dq1)
    ref = B @ lam
    assert np.linalg.norm(dq0.get() - ref) <= 1e-11 * np.linalg.norm(ref)
    assert np.linalg.norm(dq1.get() - ref) <= 1e-11 * np.linalg.norm(ref)
    g = rng.uniform(-1, 1, blk.VNp); gd = rng.uniform(-1, 1, tr.lNp)
    bl = np.zeros(tr.lNp); uu = np.zeros(blk.VNp)
    orc.LocalToGLobalRHS(bl, g, gd, uu, c["M"].F, FbarT, c["vstarts"])
    lam_ref = np.linalg.solve(B, bl)
    rhs = g - FbarT.T @ lam_ref
    u_ref = np.concatenate([c["M"].F[e].solve(rhs[blk.vol_slice(e)]) for e in range(blk.nblocks)])
    dg, dgd, dlam, dsol = ctx.array(g), ctx.array(gd), ctx.empty(tr.lNp), ctx.empty(blk.VNp)
    st = tr.solve(dg, dgd, dlam, dsol, tol=1e-13, maxit=2000)
    assert st["converged"] == 1 and st["local_solves"] == 2, st          # rhs + back-substitution only
    assert np.linalg.norm(dlam.get() - lam_ref) <= 1e-10 * np.linalg.norm(lam_ref), st
    assert np.linalg.norm(dsol.get() - u_ref) <= 1e-10 * np.linalg.norm(u_ref), st
    # face-block preconditioner from the diagonal blocks of B: same solution, fewer iterations
    it_jacobi = st["outer_iterations"]
    tr.precond_setup(1)
    Bd = np.zeros_like(B)
    starts = np.asarray(tr.FTolambdastarts) - 1
    for a, b_ in zip(starts[:-1], starts[1:]):
        Bd[a:b_, a:b_] = B[a:b_, a:b_]
    r = rng.uniform(-1, 1, tr.lNp)
    dr, dz = ctx.array(r), ctx.empty(tr.lNp)
    tr.precond_apply(dr, dz)
    z_ref = np.linalg.solve(Bd, r)
    assert np.linalg.norm(dz.get() - z_ref) <= 1e-9 * np.linalg.norm(z_ref)
    st = tr.solve(dg, dgd, dlam, dsol, tol=1e-13, maxit=2000)
    assert st["converged"] == 1 and st["outer_iterations"] < it_jacobi, (st, it_jacobi)
    assert np.linalg.norm(dlam.get() - lam_ref) <= 1e-10 * np.linalg.norm(lam_ref), st
    assert np.linalg.norm(dsol.get() - u_ref) <= 1e-10 * np.linalg.norm(u_ref), st
    tr.precond_setup(0)
    tr.condense(False)
    tr.schur_apply(dl, dq1)
    assert np.linalg.norm(dq1.get() - dq0.get()) == 0.0


def test_error_paths_report_codes_and_messages(ctx):
    """The C-ABI never throws or exits: misuse comes back as a status code with a message."""
    import hybridsbp_b200 as hs
    from hybridsbp_b200._lib import HsbpError
    p = 4
    rng = np.random.default_rng(79)
    c = build_case(hs, ctx, p, 3 * p - 1, flipped_four_block_mesh(), rng)
    blk, tr = c["blk"], c["tr"]
    with pytest.raises(HsbpError) as e:                      # no local solver yet
        tr.condense()
    assert e.value.code < 0 and "hsbp_local_setup" in str(e.value)
    blk.local_setup(hs.LOCAL_CHOLESKY)
    with pytest.raises(HsbpError) as e:                      # face blocks need the condensed matrices
        tr.precond_setup(1)
    assert e.value.code < 0 and "hsbp_trace_condense" in str(e.value)
    with pytest.raises(HsbpError):
        tr.precond_setup(7)
    with pytest.raises(HsbpError):
        blk.set_option("no_such_option", 1)
    with pytest.raises(HsbpError):
        blk.local_setup(99)
    # FDM needs blocks of one size
    mets = [random_spd_metrics(p, 12, 12, rng, scale2=0.2), random_spd_metrics(p, 14, 12, rng, scale2=0.2)]
    b2 = upload_blocks(hs, ctx, p, mets, [(1, 1, 1, 1), (1, 1, 1, 1)])
    with pytest.raises(HsbpError) as e:
        b2.local_setup(hs.LOCAL_FDM)
    assert "one size" in str(e.value)
    b2.close()
    # after the failures the objects still work
    tr.condense(); tr.precond_setup(1)
    lam = rng.uniform(-1, 1, tr.lNp)
    dl, dq = ctx.array(lam), ctx.empty(tr.lNp)
    tr.schur_apply(dl, dq)
    assert np.all(np.isfinite(dq.get()))
