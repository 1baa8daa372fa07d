"""A second, independent extraction of the SBP coefficient data from the reference's source text, compared against
BOTH consumers of the first one (tools/gen_sbp_tables.py -> oracle/sbp_tables.json -> oracle/sbp.py, and
-> tools/sbp_coeffs.py -> hybridsbp_b200/csrc/sweep_tables_gen.h, the straight-line closure code of the CUDA kernel).

The first extraction parses coefficient terms with regular expressions into tables.  This one never looks at terms:
it rewrites the reference's assignment statements (diagonal_sbp.jl:67-92 and :485-727) into Python expressions
(implicit multiplication made explicit, 1-based ranges turned into index arrays), EXECUTES them for random
coefficients, and compares matrices:
  * the full (N+1) x (N+1) stiffness matrix M(b) of variable_diagonal_sbp_D2 and the first-derivative operator of
    diagonal_sbp_D1 against the oracle (p = 2, 4, 6);
  * the kernel's generated header, compiled for the host with g++ (the CUDA qualifiers defined away), against the same
    matrices: closure rows of M(b) u, of Q u and Q^T w, and the norm weights.
A table-extraction error common to the oracle and the kernel would show up here.

Reads /root/reference, which exists only in the build container: skipped elsewhere (never part of the -m gpu run)."""
import ctypes
import os
import re
import subprocess

import numpy as np
import pytest

from oracle import sbp as osbp

REF = os.environ.get("HSBP_REFERENCE", "/root/reference")
SRC = os.path.join(REF, "diagonal_sbp.jl")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.skipif(not os.path.exists(SRC), reason="the reference tree is not mounted here")


def _lines():
    with open(SRC) as f:
        return f.read().split("\n")


def _explicit_mul(expr):
    """Julia's juxtaposition: (12/17)b1, 0.79b1, 2B[...], 8(16200x1-953), 16200x1 -> with '*'"""
    expr = re.sub(r"(\)|\d)\s*(?=b\d)", r"\1*", expr)
    expr = re.sub(r"(\)|\d)\s*(?=B\[)", r"\1*", expr)
    expr = re.sub(r"(\d)\s*(?=x1)", r"\1*", expr)
    expr = re.sub(r"(\d)\s*(?=\()", r"\1*", expr)
    return expr


def _branch(lines, start_pat, p):
    """lines of the `p == <p>` branch of the function whose header matches start_pat"""
    i0 = next(i for i, l in enumerate(lines) if re.match(start_pat, l))
    i = next(j for j in range(i0, len(lines)) if re.match(r"\s*(?:if|elseif) p == %d\b" % p, lines[j]))
    j = next(k for k in range(i + 1, len(lines)) if re.match(r"\s*(?:elseif|else)\b", lines[k]))
    return lines[i + 1:j]


def reference_stiffness(p, N, B):
    """M * h of variable_diagonal_sbp_D2 (before the division by h at :746), built by executing the reference's
    statements: closure blocks V_M0 / V_MN entry by entry, interior rows from the I_M / J_M / V_M range expressions,
    duplicates summed as sparse() does."""
    body = _branch(_lines(), r"function variable_diagonal_sbp_D2\(p, N, B::AbstractArray", p)
    M = np.zeros((N + 1, N + 1))
    if p == 2:                                  # :485-503, a handful of range expressions
        M[0, 0] += (B[0] + B[1]) / 2
        M[N, N] += (B[N - 1] + B[N]) / 2
        text = "\n".join(body)
        assert "-(B[1:N  ]+B[2:N+1])/2" in text and "(B[1:N-1]+2B[2:N]+B[3:N+1])/2" in text      # the statements being restated
        i = np.arange(1, N + 1)
        M[i, i - 1] += -(B[0:N] + B[1:N + 1]) / 2
        M[i - 1, i] += -(B[0:N] + B[1:N + 1]) / 2
        j = np.arange(1, N)
        M[j, j] += (B[0:N - 1] + 2 * B[1:N] + B[2:N + 1]) / 2
        return M
    size = 6 if p == 4 else 9
    nb = 8 if p == 4 else 12
    entry = re.compile(r"^\s*(V_M0|V_MN)\[(\d+),\s*(\d+)\]\s*=\s*(?:(?:V_M0|V_MN)\[(\d+),\s*(\d+)\]\s*=)?(.*)$")
    blocks = {"V_M0": np.zeros((size, size)), "V_MN": np.zeros((size, size))}
    env0 = {"b%d" % (k + 1): B[k] for k in range(nb)}                       # (b1, ...) = B[1:nb]
    envN = {"b%d" % (k + 1): B[N - k] for k in range(nb)}                   # (b1, ...) = B[N+1:-1:...]
    count = 0
    for l in body:
        m = entry.match(l)
        if not m:
            continue
        name, i, j = m.group(1), int(m.group(2)), int(m.group(3))
        val = eval(_explicit_mul(m.group(6)), {"__builtins__": {}}, env0 if name == "V_M0" else envN)
        blocks[name][i - 1, j - 1] = val
        if m.group(4):
            blocks[name][int(m.group(4)) - 1, int(m.group(5)) - 1] = val
        count += 1
    assert count == size * (size + 1), count            # every entry of both symmetric blocks was seen
    M[:size, :size] += blocks["V_M0"]
    M[N + 1 - size:, N + 1 - size:] += blocks["V_MN"]
    # interior: I_M = [r1; r2; ...], J_M = [...], V_M = [(expr1); (expr2); ...]
    text = "\n".join(body)
    ranges = {}
    for nm, lo, hi in re.findall(r"^\s*(r\w+)\s*=\s*(\d+):N([+-]\d+)\s*$", text, flags=re.M):
        ranges[nm] = (int(lo), int(hi))

    def rng(tok):
        tok = tok.strip()
        sh = 0
        m = re.match(r"^(.*?)\s*\.([+-])\s*(\d+)$", tok)
        if m:
            tok, sh = m.group(1).strip(), int(m.group(3)) * (1 if m.group(2) == "+" else -1)
        tok = tok.strip("()").strip()
        if tok in ranges:
            lo, hi = ranges[tok]
        else:
            m = re.match(r"^(\d+):N([+-]\d+)$", tok)
            lo, hi = int(m.group(1)), int(m.group(2))
        return np.arange(lo, N + hi + 1) + sh                                # 1-based indices

    def block_list(name):
        m = re.search(r"^\s*%s\s*=\s*\[(.*?)\]\s*$" % name, text, flags=re.M | re.S)
        return [x.strip() for x in m.group(1).split(";") if x.strip()]

    I_M, J_M, V_M = block_list("I_M"), block_list("J_M"), block_list("V_M")
    assert len(I_M) == len(J_M) == len(V_M) == p + 1
    for ri, rj, ex in zip(I_M, J_M, V_M):
        ii, jj = rng(ri), rng(rj)
        ex = _explicit_mul(ex)
        ex = re.sub(r"B\[([^\]]*)\]", lambda mm: "B[IDX(%r)]" % mm.group(1), ex)
        vals = eval(ex, {"__builtins__": {}}, {"B": B, "IDX": lambda tok: rng(tok) - 1})
        assert len(ii) == len(jj) == len(vals)
        np.add.at(M, (ii - 1, jj - 1), vals)
    return M


def reference_d1_tables(p):
    """(d, bd, bhinv) of diagonal_sbp_D1 (:69-92) by evaluating the matrix literals"""
    body = "\n".join(_branch(_lines(), r"function diagonal_sbp_D1\(p, N", p))
    env = {}
    m = re.search(r"x1\s*=\s*([0-9.]+)", body)
    if m:
        env["x1"] = float(m.group(1))

    def literal(name):
        mm = re.search(r"^\s*%s\s*=\s*\[(.*?)\]" % name, body, flags=re.M | re.S)
        rows = [r for r in re.split(r";|\n", mm.group(1)) if r.strip()]
        return np.array([[eval(_explicit_mul(tok), {"__builtins__": {}}, env) for tok in r.replace(",", " ").split()] for r in rows], dtype=float)
    return literal("d").ravel(), literal("bd"), literal("bhinv").ravel()


def reference_bs(p):
    if p == 2:
        return np.array([3 / 2, -2, 1 / 2])
    body = "\n".join(_branch(_lines(), r"function variable_diagonal_sbp_D2\(p, N, B::AbstractArray", p))
    m = re.search(r"BS\s*=\s*\[(.*?)\]", body)
    return np.array([eval(t, {"__builtins__": {}}) for t in m.group(1).split()], dtype=float)


@pytest.mark.parametrize("p", [2, 4, 6])
def test_oracle_operators_equal_the_executed_reference_statements(p):
    N = {2: 7, 4: 19, 6: 29}[p]
    rng = np.random.default_rng(100 + p)
    B = rng.uniform(0.5, 2.0, N + 1)
    Mref = reference_stiffness(p, N, B)
    assert np.abs(Mref - Mref.T).max() <= 1e-15 * np.abs(Mref).max()
    assert np.abs(Mref.sum(axis=1)).max() <= 1e-14 * np.abs(Mref).max()                 # zero row sums: what the pair form relies on
    h = 2.0 / N
    D, S0, SN, HI, H, M, r = osbp.variable_diagonal_sbp_D2(p, N, B)
    assert np.abs(M.toarray() * h - Mref).max() <= 2e-15 * np.abs(Mref).max()
    bs = reference_bs(p)
    S0ref = np.zeros((N + 1, N + 1)); S0ref[0, :len(bs)] = -B[0] * bs / h
    SNref = np.zeros((N + 1, N + 1)); SNref[N, N - np.arange(len(bs))] = B[N] * bs / h
    assert np.abs(S0.toarray() - S0ref).max() <= 1e-15 * np.abs(S0ref).max()
    assert np.abs(SN.toarray() - SNref).max() <= 1e-15 * np.abs(SNref).max()
    d, bd, bhinv = reference_d1_tables(p)
    bm, bn = bd.shape
    Dref = np.zeros((N + 1, N + 1))
    for i in range(bm, N + 1 - bm):
        Dref[i, i - p // 2:i + p // 2 + 1] = d
    Dref[:bm, :bn] = bd
    Dref[N + 1 - bm:, N + 1 - bn:] = -bd[::-1, ::-1]
    Dref /= h
    Hv = np.ones(N + 1); Hv[:bm] = 1 / bhinv; Hv[N + 1 - bm:] = 1 / bhinv[::-1]
    D1, HI1, H1, _ = osbp.diagonal_sbp_D1(p, N)
    assert np.abs(D1.toarray() - Dref).max() <= 4e-15 * np.abs(Dref).max()
    assert np.abs(H1.diagonal() - h * Hv).max() <= 1e-15
    assert np.abs(H.diagonal() - h * Hv).max() <= 1e-15


HARNESS = r"""
#include <cstring>
#define __device__
#define __forceinline__ inline
#define __constant__ static const
#include "%(hdr)s"
using namespace hsbp;
extern "C" {
#define EXPORT(P) \
  void d2_rows_##P(const double *b, const double *u, double *out) { d2_closure_rows<P>(b, u, out); } \
  double d2_row_##P(int row, const double *b, const double *u) { return d2_closure_row<P>(row, b, u); } \
  void q_rows_##P(const double *u, double *out) { q_closure_rows<P>(u, out); } \
  double qt_row_##P(int row, const double *w) { return qt_closure_row<P>(row, w); } \
  void sizes_##P(int *s) { s[0] = SweepTab<P>::H; s[1] = SweepTab<P>::MC; s[2] = SweepTab<P>::NK; s[3] = SweepTab<P>::BM; s[4] = SweepTab<P>::BN; } \
  void hw_##P(double *w) { std::memcpy(w, c_sw_hw##P, sizeof(c_sw_hw##P)); }
EXPORT(2) EXPORT(4) EXPORT(6)
}
"""


@pytest.fixture(scope="module")
def kernel_tables(tmp_path_factory):
    d = tmp_path_factory.mktemp("sweeptab")
    hdr = os.path.join(ROOT, "hybridsbp_b200", "csrc", "sweep_tables_gen.h")
    src = d / "harness.cpp"
    src.write_text(HARNESS % {"hdr": hdr})
    lib = d / "libsweeptab.so"
    subprocess.run(["g++", "-std=c++17", "-O1", "-ffp-contract=off", "-shared", "-fPIC", "-o", str(lib), str(src)], check=True)
    return ctypes.CDLL(str(lib))


@pytest.mark.parametrize("p", [2, 4, 6])
def test_kernel_closure_code_equals_the_executed_reference_statements(kernel_tables, p):
    L = kernel_tables
    P = ctypes.POINTER(ctypes.c_double)
    ptr = lambda a: a.ctypes.data_as(P)
    sz = np.zeros(5, dtype=np.int32)
    getattr(L, "sizes_%d" % p)(sz.ctypes.data_as(ctypes.POINTER(ctypes.c_int)))
    Hh, MC, NK, BM, BN = [int(v) for v in sz]
    N = {2: 7, 4: 19, 6: 29}[p]
    rng = np.random.default_rng(200 + p)
    d, bd, bhinv = reference_d1_tables(p)
    assert (BM, BN) == bd.shape and Hh == p // 2
    hw = np.zeros(BM)
    getattr(L, "hw_%d" % p)(ptr(hw))
    assert np.abs(hw - 1 / bhinv).max() <= 1e-15
    getattr(L, "d2_row_%d" % p).restype = ctypes.c_double
    getattr(L, "qt_row_%d" % p).restype = ctypes.c_double
    for trial in range(3):
        B = rng.uniform(0.5, 2.0, N + 1)
        u = rng.uniform(-1, 1, N + 1)
        Mref = reference_stiffness(p, N, B)
        want = (Mref @ u)[:MC]
        scale = (np.abs(Mref) @ np.abs(u))[:MC].max()
        assert np.abs(Mref[:MC, NK:]).max() == 0.0                       # the closure rows only reach the first NK points
        b8, u8, out = B[:NK].copy(), u[:NK].copy(), np.zeros(MC)
        getattr(L, "d2_rows_%d" % p)(ptr(b8), ptr(u8), ptr(out))
        assert np.abs(out - want).max() <= 1e-14 * scale, (p, out, want)
        one = np.array([getattr(L, "d2_row_%d" % p)(r, ptr(b8), ptr(u8)) for r in range(MC)])
        assert np.abs(one - want).max() <= 1e-14 * scale
        # the far end of the line uses the same code on reversed data: the reference's lower-right block must be the mirror image
        bN, uN = B[::-1][:NK].copy(), u[::-1][:NK].copy()
        getattr(L, "d2_rows_%d" % p)(ptr(bN), ptr(uN), ptr(out))
        wantN = (Mref @ u)[::-1][:MC]
        assert np.abs(out - wantN).max() <= 1e-14 * scale
        # Q = H D: closure rows of Q u and of Q^T w
        Q = np.zeros((N + 1, N + 1))
        for i in range(BM, N + 1 - BM):
            Q[i, i - Hh:i + Hh + 1] = d
        Q[:BM, :BN] = bd / bhinv[:, None]
        Q[N + 1 - BM:, N + 1 - BN:] = -(bd / bhinv[:, None])[::-1, ::-1]
        qu = np.zeros(BM)
        uq = u[:BN].copy()
        getattr(L, "q_rows_%d" % p)(ptr(uq), ptr(qu))
        assert np.abs(qu - (Q @ u)[:BM]).max() <= 1e-14
        w = rng.uniform(-1, 1, N + 1)
        assert np.abs(Q.T[:BM, BN:]).max() == 0.0
        wq = w[:BN].copy()
        qt = np.array([getattr(L, "qt_row_%d" % p)(r, ptr(wq)) for r in range(BM)])
        assert np.abs(qt - (Q.T @ w)[:BM]).max() <= 1e-14


def test_penalty_constants_equal_the_reference_source_text():
    """The constants of the penalty parameter tau (global_curved.jl:402-415: layers l, beta, alpha per order) read from the
    reference's text by executing its assignments, against the oracle's table and against the CUDA source (penalty_consts in
    k_generic.cuh, Sbp<P>::LPSI in sbp1d.cuh)."""
    from oracle import hybrid as orc
    src = open(os.path.join(REF, "global_curved.jl")).read().split("\n")
    i0 = next(i for i, l in enumerate(src) if re.match(r"\s*if p == 2\s*$", l) and "l = 2" in src[i + 1])
    ref = {}
    p = None
    for l in src[i0:i0 + 16]:
        m = re.match(r"\s*(?:if|elseif) p == (\d)", l)
        if m:
            p = int(m.group(1)); ref[p] = {}
            continue
        m = re.match(r"\s*(l|β|α) = (.+)$", l)
        if m and p is not None:
            ref[p][m.group(1)] = eval(m.group(2))                 # plain arithmetic: 17 / 48, 13649 / 43200, literals
    assert set(ref) == {2, 4, 6} and all(set(v) == {"l", "β", "α"} for v in ref.values()), ref
    for p, v in ref.items():
        assert orc.PENALTY[p] == (v["l"], v["β"], v["α"])
    gen = open(os.path.join(ROOT, "hybridsbp_b200", "csrc", "k_generic.cuh")).read()
    body = gen[gen.index("penalty_consts(int p"):gen.index("template <int P>", gen.index("penalty_consts(int p"))]
    found = re.findall(r"beta = ([0-9.]+); alpha = ([0-9. /]+);", body)
    assert len(found) == 3
    for (b, a), p in zip(found, (2, 4, 6)):
        assert float(b) == ref[p]["β"] and abs(eval(a) - ref[p]["α"]) == 0.0, (p, b, a)
    sb = open(os.path.join(ROOT, "hybridsbp_b200", "csrc", "sbp1d.cuh")).read()
    for p in (2, 4, 6):
        m = re.search(r"template <> struct Sbp<%d> \{\s*static constexpr int [^;]*LPSI = (\d+);" % p, sb)
        assert m and int(m.group(1)) == ref[p]["l"], (p, m and m.group(1))


def test_bp1_parameters_equal_the_reference_source_text():
    """The physical parameters of the BP1 driver (seas/BP1/BP1.jl:8-26, 54-55, sim_years, tolerances of the solve call :159-161)
    read from the reference's text, against the literals in the driver (hybridsbp_b200/bp1.py setup, which the oracle's twin shares)."""
    txt = open(os.path.join(REF, "seas", "BP1", "BP1.jl")).read()
    ref = {}
    for name in ("Vp", "ρ", "cs", "σn", "RSamin", "RSamax", "RSb", "RSDc", "RSf0", "RSV0", "RSVinit", "RSH1", "RSH2", "N", "SBPp",
                 "Lx", "Ly", "sim_years"):
        m = re.search(r"^\s*%s\s*=\s*([0-9.e+-]+)" % re.escape(name), txt, re.M)
        assert m, name
        ref[name] = float(m.group(1))
    assert ref["N"] == 200 and ref["SBPp"] == 2 and ref["sim_years"] == 1000.0
    m = re.search(r"atol\s*=\s*([0-9.e+-]+)\s*,\s*rtol\s*=\s*([0-9.e+-]+)", txt)
    assert m and (float(m.group(1)), float(m.group(2))) == (1e-5, 1e-3)
    import inspect
    from hybridsbp_b200 import bp1 as prod            # (the oracle's BP1 twin takes the same setup: one set of literals)
    for mod in (prod,):
        srcs = inspect.getsource(mod)
        lits = {}
        for py, jl in (("Vp", "Vp"), ("rho", "ρ"), ("cs", "cs"), ("sigma_n", "σn")):
            m = re.search(r"\b%s = ([0-9.e+-]+)" % py, srcs)
            assert m, (mod.__name__, py)
            lits[jl] = float(m.group(1))
        m = re.search(r"RSamin, RSamax, RSb, RSDc, RSf0, RSV0, RSVinit, RSH1, RSH2 = ([^\n]+)", srcs)
        assert m, mod.__name__
        for jl, v in zip(("RSamin", "RSamax", "RSb", "RSDc", "RSf0", "RSV0", "RSVinit", "RSH1", "RSH2"), m.group(1).split(",")):
            lits[jl] = float(v)
        for k, v in lits.items():
            assert v == ref[k], (mod.__name__, k, v, ref[k])


def _julia_function_to_python(lines, name):
    """Mechanical rewrite of one small Julia function of the reference (scalar arithmetic, if / for / return) into Python source:
    broadcasting dots dropped, `^` -> `**`, `||` / `&&` -> or / and, `for i = a:b` -> range, blocks closed by `end` -> indentation,
    the trailing tuple expression -> return."""
    i0 = next(i for i, l in enumerate(lines) if re.match(r"function %s\(" % name, l))
    head = lines[i0]
    i = i0
    while ")" not in head.split("(", 1)[1] or head.count("(") > head.count(")"):
        i += 1
        head += " " + lines[i].strip()
    args = head[head.index("(") + 1:head.rindex(")")].replace(";", ",")
    out = ["def %s(%s):" % (name, args)]
    depth = 1
    body = []
    for l in lines[i + 1:]:
        t = l.strip()
        if not t:
            continue
        if t == "end":
            depth -= 1
            if depth == 0:
                break
            continue
        t = t.replace(".*", "*").replace("./", "/").replace(".^", "**").replace("^", "**")
        t = re.sub(r"\b(exp|asinh|sqrt)\.\(", r"\1(", t)
        t = t.replace("typeof(x)(NaN)", "float('nan')").replace("||", " or ").replace("&&", " and ")
        m = re.match(r"for (\w+) = (\w+):(\w+)$", t)
        ind = "    " * depth
        if m:
            body.append(ind + "for %s in range(%s, %s + 1):" % m.groups()); depth += 1
        elif t.startswith("if "):
            body.append(ind + t + ":"); depth += 1
        elif t.startswith("elseif "):
            body.append("    " * (depth - 1) + "elif " + t[7:] + ":")
        elif t == "else":
            body.append("    " * (depth - 1) + "else:")
        else:
            body.append(ind + t)
    if re.match(r"\s*\(.*\)$", body[-1]) and "=" not in body[-1]:
        body[-1] = "    return " + body[-1].strip()
    return "\n".join(out + body)


def test_rate_and_state_and_newton_equal_the_executed_reference_statements():
    """rateandstate and newtbndv (global_curved.jl:1029-1075), rewritten mechanically into Python and EXECUTED, against the
    oracle's restatement: identical floating-point results (same operations in the same order) on random fault states, including
    the unbracketed case."""
    import math
    from oracle import hybrid as orc
    lines = open(os.path.join(REF, "global_curved.jl")).read().split("\n")
    # numpy's elementary functions on both sides (libm's differ from them in the last place; Julia has its own again): the test
    # pins the sequence of operations, not the elementary-function library
    ns = {"exp": np.exp, "asinh": np.arcsinh, "sqrt": np.sqrt, "abs": abs}
    exec(_julia_function_to_python(lines, "rateandstate"), ns)
    exec(_julia_function_to_python(lines, "newtbndv"), ns)
    rng = np.random.default_rng(42)
    for _ in range(200):
        a = rng.uniform(0.01, 0.025); psi = rng.uniform(0.4, 0.9); V0 = 1e-6; sn = 50.0; eta = rng.uniform(2.0, 6.0)
        tau = rng.uniform(10.0, 40.0); V = 10 ** rng.uniform(-12, 0)
        g1 = ns["rateandstate"](V, psi, sn, tau, eta, a, V0)
        g2 = orc.rateandstate(V, psi, sn, tau, eta, a, V0)
        assert (float(g1[0]), float(g1[1])) == (float(g2[0]), float(g2[1]))
        fref = lambda v: ns["rateandstate"](v, psi, sn, tau, eta, a, V0)
        forc = lambda v: orc.rateandstate(v, psi, sn, tau, eta, a, V0)
        xr = abs(tau / eta)
        for x0 in (1e-9, V):
            r1 = ns["newtbndv"](fref, -xr, xr, x0, ftol=1e-12, atolx=1e-12, rtolx=1e-12)
            r2 = orc.newtbndv(forc, -xr, xr, x0, ftol=1e-12, atolx=1e-12, rtolx=1e-12)
            assert (float(r1[0]), float(r1[1]), r1[2]) == (float(r2[0]), float(r2[1]), r2[2])
    r1 = ns["newtbndv"](lambda v: (v * v + 1.0, 2 * v), -1.0, 1.0, 0.3)        # no sign change: the reference's failure tuple
    r2 = orc.newtbndv(lambda v: (v * v + 1.0, 2 * v), -1.0, 1.0, 0.3)
    assert math.isnan(r1[0]) and math.isnan(r2[0]) and r1[2] == r2[2] == -500
