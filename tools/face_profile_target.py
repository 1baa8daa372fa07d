"""The face kernels (K3) at config-4 size -- 1024 blocks of 256 x 256 points -- for an ncu launch list:
F_k^T u (hsbp_face_FT), y += a F_k v (hsbp_face_F_add), traction (hsbp_face_traction).
  ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum -k regex:"k_face" python tools/face_profile_target.py"""
import sys
import numpy as np
sys.path.insert(0, ".")
import hybridsbp_b200 as hs
from hybridsbp_b200 import synthetic
ctx = hs.Context(0)
nbx = nby = 32; N, p = 255, 4
ne = nbx * nby
_, EToV, EToF, FToB = synthetic.block_grid_connectivity(nbx, nby)
blk = hs.Blocks(ctx, p, [N] * ne, [N] * ne)
blk.set_synthetic_warp(nbx, 0, 32.0, 0.8)
blk.set_bc(synthetic.block_bcs(EToF, FToB)); blk.compute_tau(2.0)
rng = np.random.default_rng(0)
u = ctx.array(rng.uniform(-1, 1, blk.VNp)); v = ctx.array(rng.uniform(-1, 1, blk.FNp))
ft, tr, y = ctx.empty(blk.FNp), ctx.empty(blk.FNp), ctx.array(np.zeros(blk.VNp))
for _ in range(3):
    blk.face_FT(u, ft); blk.face_traction(u, tr); blk.face_F_add(v, -0.5, y)
ctx.sync()
print("ok", float(np.abs(ft.get()).sum()))
