"""The reference's drivers run through the interpreter (tests/refexec/minijulia.py): the file is parsed as it lies under
/root/reference, three literals are replaced (SBP order, base grid size, number of refinement levels), the plotting statements are
dropped, and a hook at the end of the level loop hands the driver's local variables to the caller.  TEST INFRASTRUCTURE."""
import os

import numpy as np
import scipy.sparse as sp

from .minijulia import Interp, parse_source, NT, JRange

REF = "/root/reference"


def contains_id(node, names):
    if isinstance(node, tuple):
        if len(node) == 2 and node[0] == "id" and node[1] in names: return True
        return any(contains_id(c, names) for c in node)
    if isinstance(node, list):
        return any(contains_id(c, names) for c in node)
    return False


def rewrite(node, fn):
    """bottom-up copy of an AST with fn applied to every tuple node (a statement fn maps to None is dropped)"""
    if isinstance(node, tuple):
        if node and node[0] == "ln":
            inner = rewrite(node[3], fn)
            return None if inner is None else ("ln", node[1], node[2], inner)
        return fn(tuple(rewrite(c, fn) for c in node))
    if isinstance(node, list):
        out = []
        for c in node:
            r = rewrite(c, fn)
            if r is not None: out.append(r)
        return out
    return node


def to_python(v):
    if isinstance(v, NT): return {k: to_python(x) for k, x in zip(v._names, v._values)}
    if isinstance(v, JRange): return v.arr()
    if isinstance(v, tuple): return tuple(to_python(x) for x in v)
    if isinstance(v, np.ndarray): return np.array(v.view(np.ndarray))
    if sp.issparse(v): return v.copy()
    if isinstance(v, dict): return {k: to_python(x) for k, x in v.items()}
    return v


def run_square_circle(p=6, levels=1, N0=17, keep=("λ", "u", "g", "gδ", "bλ", "δ", "vstarts", "FToλstarts", "FToδstarts", "ϵ", "τϵ", "lvl",
                                                  "Nr", "Ns", "B", "FbarT", "D", "lop")):
    """square_circle.jl:1-431 executed; returns (per-level list of dicts of the driver's variables, the mesh arrays)"""
    it = Interp(REF)
    with open(os.path.join(REF, "square_circle.jl")) as f:
        ast = parse_source(f.read(), "square_circle.jl")
    captured, mesh = [], {}

    def hook(env):
        captured.append({k: to_python(env.lookup(k)) for k in keep})
        for k in ("verts", "EToV", "EToF", "FToB", "EToDomain", "FToE", "FToLF", "EToO", "EToS"):
            mesh[k] = to_python(env.lookup(k))

    def fn(n):
        if n[0] == "assign" and n[1] == ("id", "SBPp"): return ("assign", n[1], ("num", p))
        if n[0] == "assign" and n[1] == ("id", "N0") and n[2][0] == "num": return ("assign", n[1], ("num", N0))
        if n[0] == "assign" and n[1] == ("id", "ϵ"): return ("assign", n[1], ("call", ("id", "zeros"), [("num", levels)], [], False))
        if n[0] in ("assign", "call", "for", "show") and contains_id(n, ("PGFPlots", "pgf_axis")): return None
        if n[0] == "call" and n[1] == ("id", "println"): return None
        if n[0] == "for" and n[1][0][0] == ("id", "lvl"):
            return ("for", n[1], ("block", n[2][1] + [("pyhook", hook)]))
        return n

    it.exec_block(rewrite(ast, fn), it.globals)
    return captured, mesh, it


def run_bp1_setup(N=200):
    """seas/BP1/BP1.jl:1-161 executed up to the construction of the ODE problem (the integrator is a package the reference does not
    vendor): returns (interpreter, problem) with problem.f = the reference's `odefun`, .u0 = ψδ, .p = odeparam, and yf"""
    base = os.path.join(REF, "seas", "BP1")
    it = Interp(base)
    it.globals.vars["ODEProblem"] = lambda f, u0, tspan, p: NT(["f", "u0", "tspan", "p"], [f, u0, tspan, p])
    it.globals.vars["Tsit5"] = lambda: None
    it.globals.vars["solve"] = lambda prob, alg, **kw: NT(["prob", "options"], [prob, kw])
    with open(os.path.join(base, "BP1.jl")) as f:
        ast = parse_source(f.read(), "BP1.jl")

    def fn(n):
        if n[0] == "assign" and n[1] == ("id", "N") and n[2][0] == "num": return ("assign", n[1], ("num", N))
        return n
    stmts = [st for st in rewrite(ast, fn)[1] if not (st[3][0] in ("assign", "call") and contains_id(st, ("main", "plot_slip")))]
    it.exec_block(("block", stmts), it.globals)
    sol, yf = it.call("main")
    return it, sol, to_python(yf)


def run_trace_driver(meshfile, p=4, N0=17, jumpcodes=7, slipshift=0.0):
    """tests/refexec/trace_driver.jl (the test suite's own driver for meshes of straight-sided blocks) on top of the reference's
    global_curved.jl; meshfile relative to the reference's root"""
    it = Interp(REF)
    it.globals.vars.update(order=p, npts=N0, meshfile=meshfile, jumpcodes=jumpcodes, slipshift=slipshift)
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "trace_driver.jl")) as f:
        out = it.run(f.read(), "trace_driver.jl")
    return to_python(out)


def run_flower(p=4, N0=17):
    return run_trace_driver("meshes/flower_v2.inp", p, N0, 7)


PLOTTING = ("Plot", "BrailleCanvas", "annotate!", "lineplot!", "display", "PGFPlots", "pgf_axis", "plt", "plt_max", "plt_min")


def run_check_script(name, num_samp=2, capture=()):
    """one of the reference's own check scripts (local_op_eigenvalues.jl, global_op_eigenvalues.jl, check_residual.jl) as written, with
    the number of random samples reduced and the plotting statements dropped; returns (captured variables per `let` block, @show log)"""
    it = Interp(REF)
    with open(os.path.join(REF, name)) as f:
        ast = parse_source(f.read(), name)
    captured = []

    def fn(n):
        if n[0] == "assign" and n[1] == ("id", "num_samp"): return ("assign", n[1], ("num", num_samp))
        if n[0] in ("assign", "call", "for", "show") and contains_id(n, PLOTTING): return None
        if n[0] == "call" and n[1] == ("id", "range") and any(k[0] == "length" for k in n[3]):      # tau sweep of local_op_eigenvalues.jl
            return ("call", n[1], n[2], [(k, ("num", 4) if k == "length" and v == ("num", 100) else v, a) for k, v, a in n[3]], n[4])
        if n[0] == "let":
            hook = lambda env: captured.append({k: to_python(env.vars[k]) for k in capture if k in env.vars})
            return ("let", ("block", n[1][1] + [("pyhook", hook)]))
        return n
    it.exec_block(rewrite(ast, fn), it.globals)
    return captured, it.log
