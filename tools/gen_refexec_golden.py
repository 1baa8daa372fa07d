#!/usr/bin/env python
"""Golden vectors from the REFERENCE ITSELF, run here: its Julia sources are parsed where they lie under /root/reference and
executed statement by statement by tests/refexec/minijulia.py (Julia is not installed; see that module).  The outputs are
written to tests/golden/refexec/ and travel to the GPU box, where the CUDA path is compared with them directly
(tests/test_refexec_golden_gpu.py) and the oracle once more (tests/test_refexec_golden_cpu.py).

    python tools/gen_refexec_golden.py            # rewrites all files (about two minutes)

Files
  locoperator_p{2,4,6}.npz   create_metrics + locoperator (global_curved.jl:136-506) on a curved 35 x 35 block (N = 34, the size of
                             square_circle.jl's second level): coefficients, three boundary-condition sets, tau, y = M~ u, F_k' u,
                             traction operator HfI_FT_k u for a stored u
  square_circle_p{4,6}.npz   square_circle.jl:1-431 at its first level (56 blocks, N = 17): delta, g-delta, b-lambda, lambda, u, errors
  square_circle_p4_N68.npz   the same driver run for three levels; third level (N = 68): lambda, g-delta, delta, every 31st entry of u
  flower_p4.npz              the reference's functions on meshes/flower_v2.inp (27 reversed faces, given slip on the 18 jump faces)
                             through tests/refexec/trace_driver.jl: delta, g-delta, b-lambda, lambda, u, fault traction
  bp1_odefun_N40.npz         seas/BP1/BP1.jl:1-158 (setup) + odefun.jl:8-121 at three states: y, t -> d(psi, delta)/dt
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from refexec.minijulia import Interp                                  # noqa: E402
from refexec.drivers import run_square_circle, run_bp1_setup, run_flower, REF     # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden", "refexec")
CURVED_MAP = """
xfun(r,s) = (r .+ 0.1 .* sin.(2 .* r) .* cos.(s) .+ 0.2 .* s, 1 .+ 0.2 .* cos.(2 .* r) .* cos.(s), -0.1 .* sin.(2 .* r) .* sin.(s) .+ 0.2)
yfun(r,s) = (s .+ 0.15 .* sin.(r .+ s), 0.15 .* cos.(r .+ s), 1 .+ 0.15 .* cos.(r .+ s))
"""
BC_SETS = ((1, 1, 1, 1), (0, 2, 7, 1), (2, 2, 0, 0))
N_LOC = 34


def dense(a):
    return a.toarray() if hasattr(a, "toarray") else np.asarray(a)


def gen_locoperator(p):
    it = Interp(REF)
    it.include("global_curved.jl")
    it.run(CURVED_MAP)
    N = N_LOC
    m = it.call("create_metrics", p, N, N, it.globals.lookup("xfun"), it.globals.lookup("yfun"))
    rng = np.random.default_rng(100 + p)
    u = rng.uniform(-1, 1, (len(BC_SETS), (N + 1) ** 2))
    out = dict(p=p, N=N, bc=np.array(BC_SETS), u=u, crr=np.array(m.get("crr")), css=np.array(m.get("css")), crs=np.array(m.get("crs")),
               J=np.array(m.get("J")), sJ=np.array([np.array(a) for a in m.get("sJ")]))
    ys, scales, taus, fts, trs = [], [], [], [], []
    for k, bc in enumerate(BC_SETS):
        lop = it.call("locoperator", p, N, N, m, bc)
        Mt = lop.get("M̃")
        ys.append(Mt @ u[k]); scales.append(np.max(abs(Mt) @ np.abs(u[k])))
        taus.append(np.array([t.diagonal() for t in lop.get("τ")]))
        fts.append(np.array([F.T @ u[k] for F in lop.get("F")]))
        trs.append(np.array([T @ u[k] for T in lop.get("HfI_FT")]))
    out.update(y=np.array(ys), scale=np.array(scales), tau=np.array(taus), FTu=np.array(fts), traction_op_u=np.array(trs))
    np.savez_compressed(os.path.join(OUT, "locoperator_p%d.npz" % p), **out)
    print("locoperator p=%d written" % p)


def gen_square_circle(p):
    cap, mesh, _ = run_square_circle(p=p, levels=1, N0=17)
    c = cap[0]
    np.savez_compressed(os.path.join(OUT, "square_circle_p%d.npz" % p), p=p, N=17, delta=c["δ"], gdelta=c["gδ"], blambda=c["bλ"],
                        lam=c["λ"], u=c["u"], g_norm=np.linalg.norm(c["g"]), g_sample=c["g"][::37], eps=c["ϵ"][0], teps=c["τϵ"][0],
                        vstarts=c["vstarts"], FTolstarts=c["FToλstarts"], FTodstarts=c["FToδstarts"],
                        EToV=mesh["EToV"], EToF=mesh["EToF"], FToB=mesh["FToB"], EToDomain=mesh["EToDomain"], verts=mesh["verts"])
    print("square_circle p=%d written: eps = %.6e, traction eps = %.6e" % (p, c["ϵ"][0], c["τϵ"][0]))


def gen_square_circle_level3(p=4):
    """third refinement level (N = 68, 69 points per line: the banded local solver and the odd line lengths of k_sweep on the GPU side);
    u is stored as every 31st entry plus its norm to keep the file small"""
    cap, mesh, _ = run_square_circle(p=p, levels=3, N0=17, keep=("λ", "u", "gδ", "δ", "ϵ", "τϵ", "FToλstarts", "FToδstarts", "vstarts"))
    c = cap[2]
    np.savez_compressed(os.path.join(OUT, "square_circle_p%d_N68.npz" % p), p=p, N=68, delta=c["δ"], gdelta=c["gδ"], lam=c["λ"],
                        u_sample=c["u"][::31], u_norm=np.linalg.norm(c["u"]), eps=cap[2]["ϵ"], teps=cap[2]["τϵ"],
                        FTolstarts=c["FToλstarts"], FTodstarts=c["FToδstarts"])
    print("square_circle p=%d, three levels: eps = %s" % (p, ["%.9e" % v for v in cap[2]["ϵ"]]))


def gen_flower(p=4):
    c = run_flower(p, 17)
    np.savez_compressed(os.path.join(OUT, "flower_p%d.npz" % p), p=p, N=17, delta=c["δ"], gdelta=c["gδ"], blambda=c["bλ"], lam=c["λ"], u=c["u"],
                        g_sample=c["g"][::37], traction=c["τf"], FTolstarts=c["FToλstarts"], FTodstarts=c["FToδstarts"],
                        EToO=np.asarray(c["EToO"]).astype(np.int64), EToS=c["EToS"], FToB=c["FToB"])
    print("flower p=%d written" % p)


def bp1_states(y0, N):
    rng = np.random.default_rng(3)
    out = []
    for t in (0.0, 3.1e7, 2.0e9):
        y = y0.copy()
        if t > 0:
            y[N + 1:] += rng.uniform(0, 1e-3 * (1 + t * 1e-9), N + 1)
            y[:N + 1] += rng.uniform(-1e-3, 1e-3, N + 1)
        out.append((t, y))
    return out


def gen_bp1(N=40):
    it, sol, yf = run_bp1_setup(N)
    prob = sol.get("prob")
    y0 = np.array(prob.get("u0"))
    ts, ys, ds = [], [], []
    for t, y in bp1_states(y0, N):
        d = np.zeros(2 * (N + 1))
        prob.get("f")(d, y.copy(), prob.get("p"), t)
        assert not prob.get("p").get("reject_step")[0]
        ts.append(t); ys.append(y); ds.append(d)
    np.savez_compressed(os.path.join(OUT, "bp1_odefun_N%d.npz" % N), N=N, y0=y0, yf=yf, t=np.array(ts), y=np.array(ys), dydt=np.array(ds),
                        RSa=np.array(prob.get("p").get("RSa")), tau_z0=prob.get("p").get("τz0"))
    print("bp1 odefun N=%d written" % N)


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    for p in (2, 4, 6): gen_locoperator(p)
    gen_bp1(40)
    gen_flower(4)
    for p in (4, 6): gen_square_circle(p)
    gen_square_circle_level3(4)
