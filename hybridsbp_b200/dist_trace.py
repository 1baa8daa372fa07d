"""Weak-scaling trace solve on the synthetic warped multiblock mesh (BASELINE config 4 / 5): the global mesh is
(nbx * world) x nby blocks, rank r owns the strip of block columns [r nbx, (r+1) nbx); the 2 nby cut faces per
interior strip boundary are exchanged inside libhsbp (NCCL send / recv), CG scalars and coarse-level data by
all-reduce (hsbp_trace_set_partition, hsbp_trace_solve).  This file only builds the problem and calls the C-ABI."""
import time

import numpy as np

from . import parallel, synthetic
from .blocks import Blocks, Trace, LOCAL_CHOLESKY, LOCAL_FDM, LOCAL_PCG  # noqa: F401
from .host import connectivityarrays


class StripProblem:
    """One rank's part of the strip mesh on its GPU: blocks, trace, right-hand sides (device arrays of the library)."""

    def __init__(self, ctx, rank, world, nbx, nby, N, p, local_mode=None, local_tol=1e-13, seed=1234, condense=True,
                 fdm_gemm=3, face_blocks=None, coarse_modes=2, timings=None, device_geometry=True):
        tm = {} if timings is None else timings
        t0 = time.perf_counter()
        gnbx = nbx * world
        _, EToV, EToF, FToB = synthetic.block_grid_connectivity(gnbx, nby)
        FToE, FToLF, EToO, EToS = connectivityarrays(EToV, EToF)
        ne = gnbx * nby
        owner = (np.arange(ne) % gnbx) // nbx
        lm = parallel.localize(rank, owner, EToF, FToB, FToE, FToLF, EToO, EToS)
        L = float(max(gnbx, nby))
        nloc = len(lm.blocks)
        blk = Blocks(ctx, p, [N] * nloc, [N] * nloc)
        if device_geometry:                   # metrics of the analytic warp generated on the device: nothing crosses PCIe
            blk.set_synthetic_warp(nbx, rank * nbx, L, L / 40.0)
        else:
            crr, css, crs = synthetic.warped_coefficients(nbx, nby, N, L=L, A=L / 40.0, bx0=rank * nbx)
            blk.set_metrics(crr, css, crs)
            del crr, css, crs
        blk.set_bc(lm.FToB[lm.EToF - 1].T.reshape(-1))
        blk.compute_tau(2.0)
        ctx.sync()
        tm["mesh_and_metrics"] = time.perf_counter() - t0
        t0 = time.perf_counter()
        if local_mode is None:
            # small blocks: dense Cholesky factors; large, smoothly varying blocks: PCG with the separable preconditioner
            local_mode = LOCAL_CHOLESKY if (N + 1) ** 2 <= 1600 else LOCAL_FDM
        blk.set_option("fdm_gemm", fdm_gemm)
        blk.local_setup(local_mode, tol=local_tol, maxit=200000)
        ctx.sync()
        tm["local_setup"] = time.perf_counter() - t0
        tr = Trace(blk, lm.FToB, lm.FToE, lm.FToLF, lm.EToO, lm.EToS)
        if world > 1:
            tr.set_partition(*lm.partition_arrays())
        t0 = time.perf_counter()
        if condense:
            tr.condense()
        ctx.sync()
        tm["condense"] = time.perf_counter() - t0
        t0 = time.perf_counter()
        if face_blocks is None:
            face_blocks = condense
        if face_blocks:
            tr.precond_setup(1)
        ctx.sync()
        tm["face_blocks"] = time.perf_counter() - t0
        t0 = time.perf_counter()
        if coarse_modes > 0:
            tr.coarse_setup(coarse_modes)
        ctx.sync()
        tm["coarse"] = time.perf_counter() - t0
        npb = (N + 1) ** 2
        # right-hand sides seeded per global block / face: every world size solves the same global problem
        g = np.concatenate([np.random.default_rng(seed + int(e)).uniform(-1, 1, npb) for e in lm.blocks])
        st = tr.FTolambdastarts
        gd = np.concatenate([np.random.default_rng(seed + 10 ** 6 + int(f)).uniform(-1, 1, st[i + 1] - st[i])
                             for i, f in enumerate(lm.faces)] + [np.zeros(0)])
        self.ctx, self.blk, self.tr, self.lm = ctx, blk, tr, lm
        self.g, self.gd = ctx.array(g), ctx.array(gd)
        self.lam, self.u = ctx.empty(tr.lNp), ctx.empty(blk.VNp)
        self.timings = tm
        self.info = dict(blocks=nloc, lambda_points=int(tr.lNp), cut_faces=sum(len(v) for v in lm.cut.values()),
                         volume_points=int(blk.VNp), local_mode=int(local_mode), coarse_dofs=int(tr.coarse_size()),
                         condensed=bool(condense), face_blocks=bool(face_blocks), coarse_modes=int(coarse_modes))

    def solve(self, tol=1e-10, maxit=10000):
        """-> statistics of hsbp_trace_solve; lambda and u stay on the device (self.lam, self.u)"""
        return self.tr.solve(self.g, self.gd, self.lam, self.u, tol=tol, maxit=maxit)

    def close(self):
        self.tr.close()
        self.blk.close()
