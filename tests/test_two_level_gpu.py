"""GPU parity of the library-resident two-level preconditioned trace CG (hsbp_trace_precond_setup / _coarse_setup / _solve)
against the oracle's assembled Schur complement and direct solve (assembleλmatrix, global_curved.jl:743-797;
square_circle.jl:376-388).  Tolerance from the north star: lambda and u within 1e-10 relative (2-norm)."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

from oracle import hybrid as orc
from tests.util import flat, random_spd_metrics, upload_blocks, warped_metrics
from tests.test_trace_gpu import flipped_four_block_mesh

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def grid_case(hs, ctx, p, N, nbx, nby):
    from hybridsbp_b200 import synthetic
    _, EToV, EToF, FToB = synthetic.block_grid_connectivity(nbx, nby)
    FToE, FToLF, EToO, EToS = orc.connectivityarrays(EToV, EToF)
    ne = nbx * nby
    mets = [warped_metrics(p, N, N, e % nbx, e // nbx, nbx, nby) for e in range(ne)]
    bcs = [[FToB[f - 1] for f in EToF[:, e]] for e in range(ne)]
    return mets, bcs, (EToV, EToF, FToB, FToE, FToLF, EToO, EToS)


def oracle_system(p, N, mets, bcs, conn, tauscale=2.0):
    EToV, EToF, FToB, FToE, FToLF, EToO, EToS = conn
    ne = EToV.shape[1]
    lops = [orc.locoperator(p, N, N, m, bc, tauscale=tauscale) for m, bc in zip(mets, bcs)]
    M, FbarT, D, vstarts, Fl = orc.LocalGlobalOperators(lops, [N] * ne, [N] * ne, FToB, FToE, FToLF, EToO, EToS)
    B = orc.assemblelambdamatrix(Fl, vstarts, EToF, FToB, M.F, D, FbarT).toarray()
    return dict(M=M, FbarT=FbarT, D=D, vstarts=vstarts, Fl=np.asarray(Fl), B=B, ne=ne)


def reference_solution(O, g, gd):
    bl = np.zeros(O["Fl"][-1] - 1); uu = np.zeros(O["vstarts"][-1] - 1)
    orc.LocalToGLobalRHS(bl, g, gd, uu, O["M"].F, O["FbarT"], O["vstarts"])
    lam = np.linalg.solve(O["B"], bl)
    rhs = g - O["FbarT"].T @ lam
    u = np.concatenate([O["M"].F[e].solve(rhs[O["vstarts"][e] - 1:O["vstarts"][e + 1] - 1]) for e in range(O["ne"])])
    return bl, lam, u


def two_level_reference(O, modes):
    """numpy version of the preconditioner: exact face blocks + Legendre coarse space (tools/proto_coarse_space.py)"""
    B, st = O["B"], O["Fl"] - 1
    n = B.shape[0]
    faces = [(a, b) for a, b in zip(st[:-1], st[1:]) if b > a]
    Binv = [np.linalg.inv(B[a:b, a:b]) for a, b in faces]
    Z = np.zeros((n, modes * len(faces)))
    for i, (a, b) in enumerate(faces):
        s = np.linspace(-1, 1, b - a)
        for m in range(modes):
            Z[a:b, modes * i + m] = np.polynomial.legendre.Legendre.basis(m)(s)
    Ac = Z.T @ B @ Z

    def apply(r):
        z = np.zeros(n)
        for (a, b), Bi in zip(faces, Binv):
            z[a:b] = Bi @ r[a:b]
        if modes:
            z += Z @ np.linalg.solve(Ac, Z.T @ r)
        return z
    return apply


@pytest.mark.parametrize("p,modes", [(4, 2), (2, 1), (6, 3), (4, 0)])
def test_two_level_preconditioner_and_solve_match_the_oracle(ctx, p, modes):
    import hybridsbp_b200 as hs
    N = {2: 9, 4: 11, 6: 17}[p]
    nbx, nby = 4, 3
    mets, bcs, conn = grid_case(hs, ctx, p, N, nbx, nby)
    O = oracle_system(p, N, mets, bcs, conn)
    blk = upload_blocks(hs, ctx, p, mets, bcs)
    tr = hs.Trace(blk, conn[2], conn[3], conn[4], conn[5], conn[6])
    blk.local_setup(hs.LOCAL_CHOLESKY)
    tr.condense()
    tr.precond_setup(1)
    tr.coarse_setup(modes)
    assert tr.coarse_size() == modes * int(np.sum(np.diff(O["Fl"]) > 0))
    rng = np.random.default_rng(5 + p)
    # the preconditioner itself
    r = rng.uniform(-1, 1, tr.lNp)
    dr, dz = ctx.array(r), ctx.empty(tr.lNp)
    tr.precond_apply(dr, dz)
    z_ref = two_level_reference(O, modes)(r)
    assert np.linalg.norm(dz.get() - z_ref) <= 1e-9 * np.linalg.norm(z_ref)
    # B itself through the fused scatter + product + gather kernels
    lam = rng.uniform(-1, 1, tr.lNp)
    dl, dq = ctx.array(lam), ctx.empty(tr.lNp)
    tr.schur_apply(dl, dq)
    assert np.linalg.norm(dq.get() - O["B"] @ lam) <= 1e-11 * np.linalg.norm(O["B"] @ lam)
    # the solve
    g = rng.uniform(-1, 1, blk.VNp); gd = rng.uniform(-1, 1, tr.lNp)
    bl, lam_ref, u_ref = reference_solution(O, g, gd)
    dg, dgd, dlam, dsol = ctx.array(g), ctx.array(gd), ctx.empty(tr.lNp), ctx.empty(blk.VNp)
    st = tr.solve(dg, dgd, dlam, dsol, tol=1e-13, maxit=500)
    assert st["converged"] == 1 and st["failed_local_blocks"] == 0, st
    assert st["true_rel_residual"] <= 1e-11, st
    assert abs(st["b_norm"] - np.linalg.norm(bl)) <= 1e-10 * np.linalg.norm(bl)
    assert st["issued_iterations"] >= st["outer_iterations"]
    assert np.linalg.norm(dlam.get() - lam_ref) <= 1e-10 * np.linalg.norm(lam_ref), st
    assert np.linalg.norm(dsol.get() - u_ref) <= 1e-10 * np.linalg.norm(u_ref), st
    if modes >= 2:                                   # fewer iterations than with the face blocks alone
        it2 = st["outer_iterations"]
        tr.coarse_setup(0)
        st1 = tr.solve(dg, dgd, dlam, dsol, tol=1e-13, maxit=500)
        assert st1["converged"] == 1 and st1["coarse_dofs"] == 0
        assert it2 < st1["outer_iterations"], (it2, st1)
        assert np.linalg.norm(dlam.get() - lam_ref) <= 1e-10 * np.linalg.norm(lam_ref), st1
    # the chunked, run-ahead host loop gives the same iteration as a fully synchronous one
    tr.coarse_setup(modes)
    tr.set_option("cg_chunk", 1); tr.set_option("cg_lookahead", 0)
    st_sync = tr.solve(dg, dgd, dlam, dsol, tol=1e-13, maxit=500)
    assert st_sync["outer_iterations"] == st["outer_iterations"] and st_sync["issued_iterations"] == st["outer_iterations"]
    lam_sync = dlam.get()
    tr.set_option("cg_chunk", 7); tr.set_option("cg_lookahead", 2)
    st_ahead = tr.solve(dg, dgd, dlam, dsol, tol=1e-13, maxit=500)
    assert st_ahead["outer_iterations"] == st["outer_iterations"]
    assert np.array_equal(dlam.get(), lam_sync)
    tr.close(); blk.close()


def test_two_level_on_reversed_faces_matrix_free_and_jacobi_first_level(ctx):
    """orientation flips (EToO == false) in the coarse kernels; Z^T B Z from batched local solves instead of condensed blocks;
    first level = D (no face blocks); random SPD tensors"""
    import hybridsbp_b200 as hs
    p, N, modes = 4, 11, 2
    rng = np.random.default_rng(41)
    EToV, EToF, FToB = flipped_four_block_mesh()
    FToE, FToLF, EToO, EToS = orc.connectivityarrays(EToV, EToF)
    assert (~EToO).any()
    mets = [random_spd_metrics(p, N, N, rng, scale2=0.2) for _ in range(4)]
    bcs = [[FToB[f - 1] for f in EToF[:, e]] for e in range(4)]
    conn = (EToV, EToF, FToB, FToE, FToLF, EToO, EToS)
    O = oracle_system(p, N, mets, bcs, conn, tauscale=1.0)
    blk = upload_blocks(hs, ctx, p, mets, bcs, tauscale=1.0)
    tr = hs.Trace(blk, FToB, FToE, FToLF, EToO, EToS)
    blk.local_setup(hs.LOCAL_CHOLESKY)
    tr.coarse_setup(modes)                       # matrix-free: 4 * modes batched local solves
    B, st = O["B"], O["Fl"] - 1
    faces = [(a, b) for a, b in zip(st[:-1], st[1:]) if b > a]
    Z = np.zeros((B.shape[0], modes * len(faces)))
    for i, (a, b) in enumerate(faces):
        s = np.linspace(-1, 1, b - a)
        for m in range(modes):
            Z[a:b, modes * i + m] = np.polynomial.legendre.Legendre.basis(m)(s)
    r = rng.uniform(-1, 1, tr.lNp)
    z_ref = r / O["D"] + Z @ np.linalg.solve(Z.T @ B @ Z, Z.T @ r)
    dr, dz = ctx.array(r), ctx.empty(tr.lNp)
    tr.precond_apply(dr, dz)
    assert np.linalg.norm(dz.get() - z_ref) <= 1e-9 * np.linalg.norm(z_ref)
    g = rng.uniform(-1, 1, blk.VNp); gd = rng.uniform(-1, 1, tr.lNp)
    _, lam_ref, u_ref = reference_solution(O, g, gd)
    dg, dgd, dlam, dsol = ctx.array(g), ctx.array(gd), ctx.empty(tr.lNp), ctx.empty(blk.VNp)
    s1 = tr.solve(dg, dgd, dlam, dsol, tol=1e-13, maxit=2000)
    assert s1["converged"] == 1, s1
    assert np.linalg.norm(dlam.get() - lam_ref) <= 1e-10 * np.linalg.norm(lam_ref), s1
    assert np.linalg.norm(dsol.get() - u_ref) <= 1e-10 * np.linalg.norm(u_ref), s1
    # iterative local solver inside the matvec (host checks every iteration there)
    blk.local_setup(hs.LOCAL_PCG, tol=1e-14, maxit=20000)
    s2 = tr.solve(dg, dgd, dlam, dsol, tol=1e-12, maxit=2000)
    assert s2["converged"] == 1 and s2["issued_iterations"] == s2["outer_iterations"], s2
    assert np.linalg.norm(dlam.get() - lam_ref) <= 1e-10 * np.linalg.norm(lam_ref), s2
    tr.close(); blk.close()


def test_stale_operator_is_refused_and_failed_local_solves_are_reported(ctx):
    import hybridsbp_b200 as hs
    from hybridsbp_b200._lib import HsbpError
    p, N = 4, 11
    mets, bcs, conn = grid_case(hs, ctx, p, N, 2, 2)
    blk = upload_blocks(hs, ctx, p, mets, bcs)
    tr = hs.Trace(blk, conn[2], conn[3], conn[4], conn[5], conn[6])
    # a local solver that cannot converge: condensation must refuse to keep the blocks
    blk.local_setup(hs.LOCAL_PCG, tol=1e-14, maxit=2)
    with pytest.raises(HsbpError) as e:
        tr.condense()
    assert "did not reach its tolerance" in str(e.value)
    assert tr.last_local_stats()["failed_blocks"] > 0
    rng = np.random.default_rng(3)
    g = rng.uniform(-1, 1, blk.VNp); gd = rng.uniform(-1, 1, tr.lNp)
    dg, dgd, dlam, dsol = ctx.array(g), ctx.array(gd), ctx.empty(tr.lNp), ctx.empty(blk.VNp)
    st = tr.solve(dg, dgd, dlam, dsol, tol=1e-10, maxit=3)
    assert st["converged"] == 0 and st["failed_local_blocks"] > 0, st
    # changing the operator invalidates local factors and the trace
    blk.local_setup(hs.LOCAL_CHOLESKY)
    tr.condense()
    blk.compute_tau(1.0)
    with pytest.raises(HsbpError) as e:
        blk.local_solve(dg, dsol)
    assert "hsbp_local_setup" in str(e.value)
    with pytest.raises(HsbpError) as e:
        tr.solve(dg, dgd, dlam, dsol)
    assert "operator changed" in str(e.value)
    tr.close(); blk.close()


def _gpu_count():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.mark.parametrize("args", [(3, 3, 17, 4, 2, 1), (2, 2, 11, 4, 2, 0)])
def test_two_rank_solve_equals_the_single_device_solve(args):
    """NCCL path inside the library (hsbp_comm_init, hsbp_trace_set_partition, send / recv + all-reduce in hsbp_trace_solve):
    2 ranks, each on its own GPU; needs a box with two GPUs (gpurun --gpus 2)."""
    if _gpu_count() < 2:
        pytest.skip("needs two GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29533", os.path.join(ROOT, "tools", "dist_trace_check.py")] + [str(a) for a in args]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    out = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][-1])
    assert out["ok"], out
    assert out["ranks"][0]["iterations"] == out["ranks"][1]["iterations"]
    # the loop's exchanges went through peer memory (2) or, where the partner's memory cannot be mapped, NCCL (1); the tool
    # repeats the solve over NCCL and requires the same solution
    assert all(r["comm_path"] in (1, 2) and r["iterations_nccl"] == r["iterations"] for r in out["ranks"]), out
