"""One batched FDM-PCG local solve on 1024 blocks of 256 x 256 points (TF32 preconditioner GEMMs): target of the ncu launch list."""
import sys
import numpy as np
sys.path.insert(0, ".")
import hybridsbp_b200 as hs
from hybridsbp_b200 import synthetic
ctx = hs.Context(0)
nbx = nby = 32
N, p = 255, 4
crr, css, crs = synthetic.warped_coefficients(nbx, nby, N)
_, EToV, EToF, FToB = synthetic.block_grid_connectivity(nbx, nby)
ne = nbx * nby
blk = hs.Blocks(ctx, p, [N] * ne, [N] * ne)
blk.set_metrics(crr, css, crs)
blk.set_bc(synthetic.block_bcs(EToF, FToB))
blk.compute_tau(2.0)
blk.set_option("fdm_gemm", int(sys.argv[1]) if len(sys.argv) > 1 else 3)
blk.local_setup(hs.LOCAL_FDM, tol=1e-13, maxit=1000)
g = ctx.array(np.random.default_rng(5).uniform(-1, 1, blk.VNp))
x = ctx.empty(blk.VNp)
print(blk.local_solve(g, x))
