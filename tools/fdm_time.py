"""Local solves on blocks of 256 x 256 points (BASELINE config 4 shape): fast-diagonalisation PCG vs Jacobi-PCG."""
import sys, time
import numpy as np
sys.path.insert(0, ".")
import hybridsbp_b200 as hs
from hybridsbp_b200 import synthetic
ctx = hs.Context(0)
nbx = int(sys.argv[1]) if len(sys.argv) > 1 else 8
nby = int(sys.argv[2]) if len(sys.argv) > 2 else nbx
jac = len(sys.argv) > 3 and sys.argv[3] == "jacobi"
N, p = 255, 4
crr, css, crs = synthetic.warped_coefficients(nbx, nby, N)
_, EToV, EToF, FToB = synthetic.block_grid_connectivity(nbx, nby)
ne = nbx * nby
blk = hs.Blocks(ctx, p, [N] * ne, [N] * ne)
blk.set_metrics(crr, css, crs)
blk.set_bc(synthetic.block_bcs(EToF, FToB))
blk.compute_tau(2.0)
x0 = np.random.default_rng(5).uniform(-1, 1, blk.VNp)
dx0, dg, dx, dr = ctx.array(x0), ctx.empty(blk.VNp), ctx.empty(blk.VNp), ctx.empty(blk.VNp)
blk.apply(dx0, dg)
dz = ctx.empty(blk.VNp)
for name, mode, gemm, opts in (
        ("FDM-PCG, fp64 GEMMs (mma.sync f64, own kernel)", hs.LOCAL_FDM, 0, {}),
        ("FDM-PCG, TF32 (tcgen05: two fused GEMM pairs, TMA, TMEM operand)", hs.LOCAL_FDM, 3, {"fdm_tc_variant": 0}),
        ("FDM-PCG, TF32 (same, converged blocks NOT skipped)", hs.LOCAL_FDM, 3, {"fdm_tc_variant": 0, "fdm_no_skip": 1}),
        ("FDM-PCG, TF32 (tcgen05: four single-GEMM launches)", hs.LOCAL_FDM, 3, {"fdm_tc_variant": 1, "fdm_no_skip": 0}),
        ("FDM-PCG, TF32 GEMMs (cuBLAS, comparison only)", hs.LOCAL_FDM, -1, {})) + \
        ((("Jacobi-PCG", hs.LOCAL_PCG, 0, {}),) if jac else ()):
    for k, v in opts.items():
        blk.set_option(k, v)
    t0 = time.time()
    blk.set_option("fdm_gemm", gemm)
    blk.local_setup(mode, tol=1e-13, maxit=2000)
    ts = time.time() - t0
    for rep in range(2):
        t0 = time.time()
        st = blk.local_solve(dg, dx)
        dt = time.time() - t0
    blk.apply(dx, dr)
    g, r = dg.get(), dr.get()
    pre_ms = float("nan")
    if mode == hs.LOCAL_FDM:                       # the preconditioner application alone (4 batched GEMMs), CUDA events
        blk.local_precondition(dg, dz)
        ctx.timer_start()
        for _ in range(10):
            blk.local_precondition(dg, dz)
        pre_ms = ctx.timer_stop() / 10
    print("%d blocks of 256x256, %s: setup %.2f s, solve %.3f s, preconditioner %.3f ms, iterations max %d mean %.1f, failed %d, "
          "true rel residual %.2e, error %.2e" %
          (ne, name, ts, dt, pre_ms, st["iterations_max"], st["iterations_sum"] / ne, st["failed_blocks"],
           np.linalg.norm(r - g) / np.linalg.norm(g), np.linalg.norm(dx.get() - x0) / np.linalg.norm(x0)), flush=True)
