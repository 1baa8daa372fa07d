"""Static check of both bindings of the C-ABI against include/hsbp.h: every prototype of the header is parsed (return type and
parameter types), and compared kind by kind (int / int64 / size_t / double / pointer / C string) with

  * the ctypes table of hybridsbp_b200/_lib.py (the binding the tests run), and
  * every `ccall` in julia/HybridSBPB200.jl (the binding a maintainer of the reference would use; Julia is not installed in the
    build image, so this is the mechanical part of its review): symbol declared, return type, number and kinds of the argument
    types, and as many actual arguments as argument types.

The Julia file is tokenised with the lexer of tests/refexec/minijulia.py."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_prototypes():
    txt = open(os.path.join(ROOT, "include", "hsbp.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    txt = re.sub(r"//[^\n]*", "", txt)
    protos = {}
    for m in re.finditer(r"([A-Za-z_][A-Za-z0-9_ \*]*?)\b(hsbp_[A-Za-z0-9_]+)\s*\(([^;{]*?)\)\s*;", txt, flags=re.S):
        ret, name, params = m.group(1).strip(), m.group(2), m.group(3).strip()
        plist = [] if params in ("", "void") else [p.strip() for p in params.split(",")]
        protos[name] = (c_kind(ret, is_return=True), [c_kind(p) for p in plist])
    return protos


def c_kind(decl, is_return=False):
    d = decl.replace("const", " ").strip()
    if "*" in d:
        return "cstring" if re.search(r"\bchar\b", d) else "pointer"
    base = d.split()[0] if is_return or len(d.split()) == 1 else " ".join(d.split()[:-1])
    base = base.strip()
    return {"int": "int", "int64_t": "int64", "size_t": "size_t", "double": "double", "void": "void"}[base]


def ctypes_kind(t):
    if t is None: return "void"
    if t is C.c_int: return "int"
    if t is C.c_int64: return "int64"
    if t is C.c_size_t: return "size_t"
    if t is C.c_double: return "double"
    if t is C.c_char_p: return "cstring"
    if t is C.c_void_p or hasattr(t, "contents") or issubclass(t, C._Pointer): return "pointer"
    raise AssertionError("unmapped ctypes type %r" % (t,))


def julia_kind(tokens):
    s = "".join(tokens)
    if s.startswith(("Ptr{", "Ref{")): return "pointer"
    return {"Cint": "int", "Int64": "int64", "Clonglong": "int64", "Csize_t": "size_t", "Cdouble": "double", "Float64": "double",
            "Cstring": "cstring", "Cvoid": "void", "Nothing": "void"}[s]


def test_header_parses_completely():
    import hybridsbp_b200 as hs
    protos = header_prototypes()
    assert sorted(protos) == hs.declared_symbols()
    assert protos["hsbp_ctx_create"] == ("int", ["int", "pointer"])
    assert protos["hsbp_last_error"] == ("cstring", ["pointer"])
    assert protos["hsbp_face_F_add"] == ("int", ["pointer", "pointer", "double", "pointer"])
    assert protos["hsbp_blocks_num_volume_points"] == ("int64", ["pointer"])


def test_ctypes_signatures_match_header_types():
    import __graft_entry__ as g
    import hybridsbp_b200 as hs
    g.build()
    protos = header_prototypes()
    sig = hs.lib()._signatures
    for name, (ret, params) in protos.items():
        res, args = sig[name]
        assert ctypes_kind(res) == ret, name
        assert [ctypes_kind(a) for a in args] == params, name


def split_top(tokens):
    """split a token list at top-level commas"""
    out, cur, depth = [], [], 0
    for t in tokens:
        if t.kind == "op" and t.val in "([{": depth += 1
        if t.kind == "op" and t.val in ")]}": depth -= 1
        if t.kind == "op" and t.val == "," and depth == 0:
            out.append(cur); cur = []
        else:
            cur.append(t)
    if cur: out.append(cur)
    return out


def julia_ccalls():
    import sys
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from refexec.minijulia import lex
    toks = [t for t in lex(open(os.path.join(ROOT, "julia", "HybridSBPB200.jl")).read()) if t.kind != "nl"]
    calls = []
    for i, t in enumerate(toks):
        if t.kind == "id" and t.val == "ccall" and toks[i + 1].val == "(":
            depth, j = 0, i + 1
            while True:
                if toks[j].kind == "op" and toks[j].val in "([{": depth += 1
                if toks[j].kind == "op" and toks[j].val in ")]}": depth -= 1
                j += 1
                if depth == 0: break
            parts = split_top(toks[i + 2:j - 1])
            sym = [x.val for x in parts[0] if x.kind in ("id",)][0]
            lib = [x.val for x in parts[0] if x.kind in ("id",)][1]
            ret = [str(x.val) for x in parts[1]]
            argt = split_top(parts[2][1:-1])
            calls.append((t.line, sym, lib, ret, [[str(x.val) for x in a] for a in argt], len(parts) - 3))
    return calls


def test_julia_ccalls_match_header():
    protos = header_prototypes()
    calls = julia_ccalls()
    assert len(calls) >= 45
    seen = set()
    for line, sym, lib, ret, argt, nargs in calls:
        where = "julia/HybridSBPB200.jl:%d %s" % (line, sym)
        assert lib == "libhsbp", where
        assert sym in protos, where + " is not declared in include/hsbp.h"
        cret, cparams = protos[sym]
        assert julia_kind(ret) == cret, where
        assert [julia_kind(a) for a in argt] == cparams, where
        assert nargs == len(argt), where + ": %d arguments for %d argument types" % (nargs, len(argt))
        seen.add(sym)
    # the twin covers the calls of the solve path: context, blocks, apply, local solves, trace solve, the factorization plugin
    for must in ("hsbp_ctx_create", "hsbp_blocks_create", "hsbp_blocks_set_metrics", "hsbp_blocks_compute_tau", "hsbp_apply", "hsbp_local_setup",
                 "hsbp_local_solve", "hsbp_trace_create", "hsbp_trace_condense", "hsbp_trace_solve", "hsbp_factor_create", "hsbp_factor_solve"):
        assert must in seen, must


def header_structs():
    txt = open(os.path.join(ROOT, "include", "hsbp.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    out = {}
    for m in re.finditer(r"typedef\s+struct\s*\{(.*?)\}\s*(hsbp_[a-z0-9_]+)\s*;", txt, flags=re.S):
        fields = []
        for decl in m.group(1).split(";"):
            decl = decl.strip()
            if not decl: continue
            typ, names = decl.split(None, 1)
            for nm in names.split(","):
                fields.append((nm.strip(), {"int64_t": "int64", "double": "double", "int": "int"}[typ]))
        out[m.group(2)] = fields
    return out


def julia_structs():
    src = open(os.path.join(ROOT, "julia", "HybridSBPB200.jl")).read()
    out = {}
    for m in re.finditer(r"^struct (\w+)[^\n]*\n(.*?)^end", src, flags=re.S | re.M):
        fields = []
        for line in m.group(2).split("\n"):
            for fm in re.finditer(r"(\w+)::(\w+)", line.split("#")[0]):
                fields.append((fm.group(1), {"Int64": "int64", "Float64": "double", "Cdouble": "double", "Cint": "int"}.get(fm.group(2), fm.group(2))))
        out[m.group(1)] = fields
    return out


def test_struct_layouts_agree():
    """the four plain structs that cross the boundary: field order and types in the header, in ctypes and in the Julia twin"""
    from hybridsbp_b200 import _lib
    hs_, js = header_structs(), julia_structs()
    pairs = {"hsbp_local_stats": ("LocalStats", _lib.LocalStats), "hsbp_trace_stats": ("TraceStats", _lib.TraceStats),
             "hsbp_bp1_params": ("Bp1Params", _lib.Bp1Params), "hsbp_bp1_stats": ("Bp1Stats", _lib.Bp1Stats)}
    for cname, (jname, ct) in pairs.items():
        assert cname in hs_, cname
        cf = hs_[cname]
        pyf = [(n, ctypes_kind(t)) for n, t in ct._fields_]
        assert pyf == cf, cname
        assert jname in js, jname
        assert js[jname] == cf, jname


def check_julia_blocks(src):
    import sys
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from refexec.minijulia import lex
    src = re.sub(r'"""(.*?)"""', lambda m: '"' + m.group(1).replace('"', "'").replace("\n", " ") + '"', src, flags=re.S)   # doc strings
    src = re.sub(r"(\d)_(\d)", r"\1\2", src)                                                                        # 100_000
    toks = lex(src)
    openers = {"function", "if", "for", "while", "let", "begin", "try", "do", "struct"}
    depth, sq, par, curly, stack = 0, 0, 0, 0, []
    prev = None
    for t in toks:
        if t.kind == "op":
            sq += {"[": 1, "]": -1}.get(t.val, 0); par += {"(": 1, ")": -1}.get(t.val, 0); curly += {"{": 1, "}": -1}.get(t.val, 0)
            assert sq >= 0 and par >= 0 and curly >= 0, "unbalanced bracket at line %d" % t.line
        if t.kind == "id" and t.val == "module":
            depth += 1; stack.append(("module", t.line))
        if t.kind == "kw":
            if t.val in openers and sq == 0:
                if t.val == "struct" and prev is not None and prev.kind == "kw" and prev.val == "mutable": pass
                depth += 1; stack.append((t.val, t.line))
            elif t.val == "end" and sq == 0:
                depth -= 1
                assert depth >= 0, "`end` without an opener at line %d" % t.line
                stack.pop()
        if t.kind != "nl": prev = t
    assert (depth, sq, par, curly) == (0, 0, 0, 0), "unclosed blocks: %r" % (stack[-3:],)


def test_julia_twin_block_structure_balances():
    """lexical sanity of the Julia file: every block opener (module, struct, function, if, for, while, let, begin, try, do) has its
    `end` (an `end` inside square brackets is an index), brackets balance, no token the lexer cannot read"""
    src = open(os.path.join(ROOT, "julia", "HybridSBPB200.jl")).read()
    check_julia_blocks(src)
    k = src.rindex("\nend")
    with pytest.raises(AssertionError):
        check_julia_blocks(src[:k] + src[k + 4:])                       # the checker notices a missing `end` ...
    with pytest.raises(AssertionError):
        check_julia_blocks(src.replace("function close(c::Context)", "function close(c::Context", 1))    # ... and an open parenthesis
