"""Multi-GPU check of the library-resident trace solve (run under torchrun, one rank per GPU):

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
      tools/dist_trace_check.py [nbx_per_rank] [nby] [N] [p] [modes] [condense]

Every rank solves its strip of the (nbx * world) x nby warped mesh through hsbp_trace_solve on the context's NCCL
communicator, then solves the WHOLE mesh alone on a second, communicator-free context of the same GPU and compares
lambda on its faces and u on its blocks.  Exit code 0 only if every rank agrees to 1e-10 and the copies of lambda on
the cut faces are bitwise identical on both ranks."""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

import hybridsbp_b200 as hs
from hybridsbp_b200 import dist_trace

arg = lambda i, d: int(sys.argv[i]) if len(sys.argv) > i else d
nbx, nby, N, p, modes, condense = arg(1, 3), arg(2, 3), arg(3, 17), arg(4, 4), arg(5, 2), arg(6, 1)
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("gloo")             # only to pass the unique id and gather the verdict: the data path is the library's NCCL
ctx = hs.Context(local)
ctx.comm_init_torch(dist)
assert ctx.world == world and ctx.rank == rank

pr = dist_trace.StripProblem(ctx, rank, world, nbx, nby, N, p, condense=bool(condense), coarse_modes=modes, local_tol=1e-14)
st = pr.solve(tol=1e-12, maxit=5000)             # default: exchanges of the loop through peer memory where it can be mapped
path = pr.tr.comm_path()
lam, u = pr.lam.get(), pr.u.get()
pr.tr.set_option("cg_p2p", 0)                    # the same solve over NCCL: same iteration count, same solution
st_nccl = pr.solve(tol=1e-12, maxit=5000)
path_nccl = pr.tr.comm_path()
lam_nccl = pr.lam.get()
pr.tr.set_option("cg_p2p", 1)
st = pr.solve(tol=1e-12, maxit=5000)             # (timed with the mailboxes and the graph in place)

ctx1 = hs.Context(local)                    # the whole mesh on one device, no communicator
ref = dist_trace.StripProblem(ctx1, 0, 1, nbx * world, nby, N, p, condense=bool(condense), coarse_modes=modes, local_tol=1e-14)
st1 = ref.solve(tol=1e-12, maxit=5000)
lam1, u1 = ref.lam.get(), ref.u.get()
gst = ref.tr.FTolambdastarts
lst = pr.tr.FTolambdastarts
rows = np.concatenate([np.arange(gst[f] - 1, gst[f + 1] - 1) for f in pr.lm.faces] + [np.zeros(0, dtype=np.int64)]).astype(np.int64)
npb = (N + 1) ** 2
cols = np.concatenate([np.arange(e * npb, (e + 1) * npb) for e in pr.lm.blocks])
elam = float(np.linalg.norm(lam - lam1[rows]) / np.linalg.norm(lam1))
eu = float(np.linalg.norm(u - u1[cols]) / np.linalg.norm(u1))

# copies of lambda on the cut faces: bitwise identical on both ranks
cut_vals = {}
for q, fl in pr.lm.cut.items():
    for i in fl:
        cut_vals[int(pr.lm.faces[i])] = lam[lst[i] - 1:lst[i + 1] - 1].tobytes()
allcut = [None] * world
dist.all_gather_object(allcut, cut_vals)
mismatch = 0
for f, v in cut_vals.items():
    others = [c[f] for r, c in enumerate(allcut) if r != rank and f in c]
    assert len(others) == 1
    mismatch += int(others[0] != v)

out = dict(rank=rank, world=world, err_lambda=elam, err_u=eu, iterations=st["outer_iterations"], iterations_single=st1["outer_iterations"],
           converged=st["converged"], true_rel_residual=st["true_rel_residual"], issued=st["issued_iterations"],
           coarse_dofs=st["coarse_dofs"], cut_faces=pr.info["cut_faces"], cut_copies_differ=mismatch, timings=pr.timings,
           comm_path=path, comm_path_second_solve=path_nccl, iterations_nccl=st_nccl["outer_iterations"],
           cg_loop_ms=st["cg_loop_ms"], cg_loop_ms_nccl=st_nccl["cg_loop_ms"],
           nccl_vs_p2p_lambda=float(np.linalg.norm(lam - lam_nccl) / np.linalg.norm(lam)))
res = [None] * world
dist.all_gather_object(res, out)
ok = all(r["err_lambda"] <= 1e-10 and r["err_u"] <= 1e-10 and r["converged"] == 1 and r["cut_copies_differ"] == 0 and
         r["nccl_vs_p2p_lambda"] <= 1e-10 and r["comm_path_second_solve"] == 1 for r in res)
if rank == 0:
    print(json.dumps(dict(ok=ok, ranks=res)))
pr.close(); ref.close()
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
