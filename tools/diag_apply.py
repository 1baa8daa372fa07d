"""Diagnostic (GPU box): where does the marching kernel disagree with the generic kernels?"""
import sys
import numpy as np
sys.path.insert(0, ".")
import hybridsbp_b200 as hs
from tests.util import random_spd_metrics, upload_blocks

p, Nr, Ns, R, ncs = [int(a) for a in sys.argv[1:6]]
fold = int(sys.argv[6]) if len(sys.argv) > 6 else 1
ctx = hs.Context(0)
rng = np.random.default_rng(1)
nb = 2
mets = [random_spd_metrics(p, Nr, Ns, rng, scale2=0.05) for _ in range(nb)]
bcs = [(1, 2, 0, 7), (2, 1, 1, 0)]
blk = upload_blocks(hs, ctx, p, mets, bcs)
u = rng.uniform(-1, 1, blk.VNp)
du, dy = ctx.array(u), ctx.empty(blk.VNp)
blk.force_generic(True)
blk.apply(du, dy)
y0 = dy.get()
blk.force_generic(False)
blk.set_option("sweep_points_per_thread", R)
blk.set_option("sweep_chunks_per_side", ncs)
blk.set_option("sweep_fold_faces", fold)
dy.set(np.full(blk.VNp, np.nan))
blk.apply(du, dy)
y1 = dy.get()
print("variant", blk.apply_variant())
for e in range(nb):
    sl = blk.vol_slice(e)
    d = np.abs(y1[sl] - y0[sl]).reshape(Nr + 1, Ns + 1, order="F")
    sc = np.abs(y0[sl]).max()
    bad = np.argwhere(~(d <= 1e-11 * sc))
    print("block", e, "max rel diff", np.nanmax(d) / sc, "nan", np.isnan(d).sum(), "bad points", len(bad))
    if len(bad):
        print("  i range", bad[:, 0].min(), bad[:, 0].max(), " j range", bad[:, 1].min(), bad[:, 1].max())
        print("  first", bad[:12].tolist())
