"""Weak-scaling trace solve on the synthetic warped multiblock mesh (BASELINE config 5): the global mesh is
(nbx * world) x nby blocks, rank r owns the strip of block columns [r nbx, (r+1) nbx); the 2 nby cut faces per
interior strip boundary are exchanged point to point, CG scalars are all-reduced (hybridsbp_b200/parallel.py)."""
import numpy as np

from . import parallel, synthetic
from .blocks import Blocks, Trace, LOCAL_CHOLESKY, LOCAL_FDM, LOCAL_PCG  # noqa: F401
from .host import connectivityarrays


def build_strip_problem(ctx, rank, world, nbx, nby, N, p, dist=None, local_mode=None, local_tol=1e-13, seed=1234,
                        condense=False, fdm_gemm=3, face_blocks=None, coarse_modes=0):
    """-> (DistributedTrace, g, gd, info).  g, gd are torch tensors on the rank's GPU; the right-hand sides are
    seeded per global block / face so that every world size solves the same global problem on the same mesh."""
    import torch
    gnbx = nbx * world
    _, EToV, EToF, FToB = synthetic.block_grid_connectivity(gnbx, nby)
    FToE, FToLF, EToO, EToS = connectivityarrays(EToV, EToF)
    ne = gnbx * nby
    owner = (np.arange(ne) % gnbx) // nbx
    lm = parallel.localize(rank, owner, EToF, FToB, FToE, FToLF, EToO, EToS)
    L = float(max(gnbx, nby))
    crr, css, crs = synthetic.warped_coefficients(nbx, nby, N, L=L, A=L / 40.0, bx0=rank * nbx)
    nloc = len(lm.blocks)
    blk = Blocks(ctx, p, [N] * nloc, [N] * nloc)
    blk.set_metrics(crr, css, crs)
    blk.set_bc(lm.FToB[lm.EToF - 1].T.reshape(-1))
    blk.compute_tau(2.0)
    if local_mode is None:
        # small blocks: dense Cholesky factors; large, smoothly varying blocks: PCG with the separable preconditioner
        local_mode = LOCAL_CHOLESKY if (N + 1) ** 2 <= 1600 else LOCAL_FDM
    blk.set_option("fdm_gemm", fdm_gemm)
    blk.local_setup(local_mode, tol=local_tol, maxit=200000)
    tr = Trace(blk, lm.FToB, lm.FToE, lm.FToLF, lm.EToO, lm.EToS)
    if condense:
        tr.condense()
    op = parallel.GpuLocalOperator(blk, tr)
    dev = torch.device("cuda", ctx.device)
    dt = parallel.DistributedTrace(op, tr.FTolambdastarts, lm, dist=dist, device=dev)
    if face_blocks is None:
        face_blocks = condense
    if face_blocks:                      # after DistributedTrace has completed D on the cut faces
        parallel.setup_face_block_preconditioner(tr, lm, tr.FTolambdastarts, dist, dev)
        op.has_precond = True
    if coarse_modes > 0:                 # optional second level (parallel.DistributedTrace.setup_coarse_space)
        dt.setup_coarse_space(coarse_modes)
    npb = (N + 1) ** 2
    g = np.concatenate([np.random.default_rng(seed + int(e)).uniform(-1, 1, npb) for e in lm.blocks])
    gd = np.concatenate([np.random.default_rng(seed + 10 ** 6 + int(f)).uniform(-1, 1, tr.FTolambdastarts[i + 1] - tr.FTolambdastarts[i])
                         for i, f in enumerate(lm.faces)] + [np.zeros(0)])
    info = dict(blocks=nloc, lambda_points=int(tr.lNp), cut_faces=sum(len(v) for v in lm.cut.values()),
                volume_points=int(blk.VNp), local_mode=int(local_mode), lm=lm, blk=blk, tr=tr)
    return dt, torch.as_tensor(g, device=dev), torch.as_tensor(gd, device=dev), info
