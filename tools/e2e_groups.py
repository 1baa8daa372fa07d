import sys, time
import numpy as np
sys.path.insert(0, ".")
import hybridsbp_b200 as hs
from hybridsbp_b200 import synthetic
ctx = hs.Context(0)
nb, N, p = 1024, 255, 4
_, EToV, EToF, FToB = synthetic.block_grid_connectivity(32, 32)
blk = hs.Blocks(ctx, p, [N] * nb, [N] * nb)
blk.set_synthetic_warp(32, 0, 32.0, 0.8)
blk.set_bc(synthetic.block_bcs(EToF, FToB)); blk.compute_tau(2.0)
u = np.random.default_rng(1).uniform(-1, 1, blk.VNp); y = np.empty(blk.VNp)
ctx.host_register(u); ctx.host_register(y)
for g in (16, 8, 32, 64, 128, 16):
    blk.set_option("host_groups", g)
    blk.apply_host_pinned(u, y)
    t0 = time.perf_counter()
    for _ in range(5): blk.apply_host_pinned(u, y)
    dt = (time.perf_counter() - t0) / 5
    print("groups %3d: %.2f ms per apply through host buffers, %.2f GDOF/s, %.1f GB/s each way" % (g, dt * 1e3, blk.VNp / dt / 1e9, 8 * blk.VNp / dt / 1e9), flush=True)
