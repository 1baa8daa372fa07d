import sys, time
import numpy as np
sys.path.insert(0, ".")
import hybridsbp_b200 as hs
from hybridsbp_b200 import square_circle as sc, host
from hybridsbp_b200.blocks import Blocks, Trace, LOCAL_BAND
ctx = hs.Context(0)
mesh = sc.load_mesh(sc.default_mesh_path())
p, N = int(sys.argv[1]), int(sys.argv[2])
verts, EToV, EToF, FToB, dom = mesh
ne = EToV.shape[1]
FToE, FToLF, EToO, EToS = host.connectivityarrays(EToV, EToF)
t0 = time.time(); mets = sc.geometry(mesh, p, N, None); print("geometry (host) %.2f s" % (time.time() - t0))
fl = lambda a: np.asarray(a).reshape(-1, order="F")
blk = Blocks(ctx, p, [N] * ne, [N] * ne)
blk.set_metrics(np.concatenate([fl(m.crr) for m in mets]), np.concatenate([fl(m.css) for m in mets]), np.concatenate([fl(m.crs) for m in mets]))
bcs = np.array([[FToB[f - 1] for f in EToF[:, e]] for e in range(ne)], dtype=np.int64)
blk.set_bc(bcs.reshape(-1)); blk.compute_tau(2.0)
def T(name, f):
    ctx.sync(); t0 = time.time(); r = f(); ctx.sync(); print("%-28s %.2f s" % (name, time.time() - t0), flush=True); return r
T("local_setup (band)", lambda: blk.local_setup(LOCAL_BAND, tol=1e-14, maxit=200000))
tr = Trace(blk, FToB, FToE, FToLF, EToO, EToS)
T("condense", lambda: tr.condense())
T("face blocks", lambda: tr.precond_setup(1))
T("coarse", lambda: tr.coarse_setup(2))
g = ctx.array(np.random.default_rng(0).uniform(-1, 1, blk.VNp)); gd = ctx.array(np.random.default_rng(1).uniform(-1, 1, tr.lNp))
lam, u = ctx.empty(tr.lNp), ctx.empty(blk.VNp)
st = T("solve", lambda: tr.solve(g, gd, lam, u, tol=1e-12, maxit=5000))
print(st["outer_iterations"], st["cg_loop_ms"])
