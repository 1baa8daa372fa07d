"""One batched FDM-PCG local solve on 1024 blocks of 256 x 256 points (TF32 preconditioner GEMMs): target for ncu.
The setup runs 2048 cuSOLVER syevd calls (about 100 kernels each): profile with a kernel filter and a launch cap, e.g.
  ncu -k regex:"k_sweep|k_fpcg|gemm|k_edge" --launch-skip 200 -c 400 --metrics gpu__time_duration.sum ...
(an unfiltered launch list of this script does not finish within a gpurun call)."""
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
