"""Three applications of the fast-diagonalisation preconditioner (4 tensor-core GEMMs each) on nbx x nby blocks of 256 x 256
points: target for ncu (-k regex:k_tc_gemm).  usage: python tools/precond_profile_target.py [nbx] [nby] [fdm_gemm]"""
import sys
import numpy as np
sys.path.insert(0, ".")
import hybridsbp_b200 as hs
from hybridsbp_b200 import synthetic
nbx = int(sys.argv[1]) if len(sys.argv) > 1 else 16
nby = int(sys.argv[2]) if len(sys.argv) > 2 else 16
ctx = hs.Context(0)
N, p = 255, 4
ne = nbx * nby
_, EToV, EToF, FToB = synthetic.block_grid_connectivity(nbx, nby)
blk = hs.Blocks(ctx, p, [N] * ne, [N] * ne)
blk.set_synthetic_warp(nbx, 0, float(max(nbx, nby)), max(nbx, nby) / 40.0)
blk.set_bc(synthetic.block_bcs(EToF, FToB))
blk.compute_tau(2.0)
blk.set_option("fdm_gemm", int(sys.argv[3]) if len(sys.argv) > 3 else 3)
blk.local_setup(hs.LOCAL_FDM, tol=1e-13, maxit=1000)
r = ctx.array(np.random.default_rng(5).uniform(-1, 1, blk.VNp))
z = ctx.empty(blk.VNp)
for _ in range(3):
    blk.local_precondition(r, z)
ctx.sync()
print("ok", float(np.abs(z.get()).sum()))
