#!/usr/bin/env python3
"""Extract the published SBP coefficient tables into oracle/sbp_tables.json and
hybridsbp_b200/csrc/sbp_tables_gen.h (two independent encodings of the same numbers).

TEST INFRASTRUCTURE (oracle side).  Run once in the build container, where
/root/reference is mounted; the JSON it writes is committed, this script is the
record of how it was made.  Nothing at test/bench time reads /root/reference.

What is extracted (numbers only, re-encoded as data -- no reference source text
is kept):
  * first-derivative tables d / bd / bhinv for p = 2, 4, 6
      (reference diagonal_sbp.jl:69-92)
  * variable-coefficient second-derivative closure blocks, i.e. for every
    closure entry (i, j) the list of (k, c) with  M0[i, j] = sum_k c * b_k,
    in the reference's own summation order
      p = 4: diagonal_sbp.jl:515-537   (6x6, b1..b8, rational c)
      p = 6: diagonal_sbp.jl:595-640   (9x9, b1..b12, decimal c)
  * the lower-right closure block is parsed too and checked to be the exact
    mirror image (diagonal_sbp.jl:539-562, 642-687), so only one copy is stored.

Rational coefficients are stored as ["num", "den"] strings; the oracle evaluates
them as float(num) / float(den), which is what the reference's host language
does with integer literals (including those beyond 64 bits).
"""
import json
import os
import re
import sys

REF = os.environ.get("HSBP_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def _func_body(lines, header_prefix):
    """Lines of the function whose header starts with header_prefix."""
    start = next(i for i, l in enumerate(lines) if l.startswith(header_prefix))
    return start


def parse_closure(lines, name, size):
    """{(i, j): [(k, coef_repr), ...]} for lines like  NAME[i, j] = NAME[j, i] = expr."""
    entry_re = re.compile(r"^\s*%s\[(\d+),\s*(\d+)\]\s*=\s*(?:%s\[(\d+),\s*(\d+)\]\s*=)?(.*)$" % (name, name))
    term_re = re.compile(r"([+-]?)\s*(\(\s*\d+\s*/\s*\d+\s*\)|\d+\.\d+|\d+)\s*b(\d+)")
    out = {}
    for l in lines:
        m = entry_re.match(l)
        if not m:
            continue
        i, j = int(m.group(1)), int(m.group(2))
        expr = m.group(5)
        terms = []
        pos = 0
        for t in term_re.finditer(expr):
            gap = expr[pos:t.start()].strip()
            assert gap == "", "unparsed text %r in %r" % (gap, l)
            pos = t.end()
            sign = -1 if t.group(1) == "-" else 1
            c = t.group(2)
            k = int(t.group(3))
            if c.startswith("("):
                num, den = [s.strip() for s in c.strip("()").split("/")]
                coef = [("-" if sign < 0 else "") + num, den]
            else:
                coef = ("-" if sign < 0 else "") + c
            terms.append([k, coef])
        assert expr[pos:].strip() == "", "unparsed tail %r" % expr[pos:]
        assert terms, l
        out[(i, j)] = terms
        if m.group(3) is not None:
            i2, j2 = int(m.group(3)), int(m.group(4))
            assert (i2, j2) == (j, i)
            out[(j, i)] = terms
    assert len(out) == size * size, (name, len(out))
    return out


def _julia_eval(expr, env):
    """Evaluate a numeric expression written with implicit multiplication."""
    e = expr.strip()
    e = re.sub(r"(\d)\s*\(", r"\1*(", e)        # 8(…)   -> 8*(…)
    e = re.sub(r"(\d)(x1)", r"\1*\2", e)         # 16200x1 -> 16200*x1
    return float(eval(e, {"__builtins__": {}}, env))


def parse_matrix_literal(text, env):
    """rows separated by ';' or newline, entries by whitespace (no spaces inside an entry)."""
    text = text.strip().strip("[]")
    rows = [r for r in re.split(r";|\n", text) if r.strip()]
    return [[_julia_eval(tok, env) for tok in r.split()] for r in rows]


def parse_d1(lines):
    start = _func_body(lines, "function diagonal_sbp_D1")
    end = next(i for i in range(start, len(lines)) if lines[i].startswith("end"))
    body = "\n".join(lines[start:end])
    out = {}
    for p in (2, 4, 6):
        m = re.search(r"p == %d\n(.*?)\n  elseif" % p, body, re.S)
        blk = m.group(1)
        env = {}
        mx = re.search(r"x1\s*=\s*([0-9.]+)", blk)
        if mx:
            env["x1"] = float(mx.group(1))
        tab = {}
        for nm in ("bhinv", "d", "bd"):
            mm = re.search(r"\b%s\s*=\s*\[(.*?)\]" % nm, blk, re.S)
            tab[nm] = parse_matrix_literal(mm.group(1), env)
        tab["bhinv"] = tab["bhinv"][0]
        tab["d"] = tab["d"][0]
        if "x1" in env:
            tab["x1"] = env["x1"]
        out[str(p)] = tab
    return out


def main():
    lines = open(os.path.join(REF, "diagonal_sbp.jl")).read().split("\n")
    tables = {"_about": "coefficient data extracted by oracle/gen_sbp_tables.py; "
                        "see that file for provenance (reference diagonal_sbp.jl line ranges)"}
    tables["D1"] = parse_d1(lines)

    vstart = next(i for i, l in enumerate(lines)
                  if l.startswith("function variable_diagonal_sbp_D2(p, N, B::AbstractArray"))
    vlines = lines[vstart:]
    p4 = next(i for i, l in enumerate(vlines) if "elseif p == 4" in l)
    p6 = next(i for i, l in enumerate(vlines) if "elseif p == 6" in l)
    pend = next(i for i, l in enumerate(vlines) if i > p6 and l.strip() == "else")
    d2 = {}
    for p, (a, b), size in ((4, (p4, p6), 6), (6, (p6, pend), 9)):
        blk = vlines[a:b]
        m0 = parse_closure(blk, "V_M0", size)
        mn = parse_closure(blk, "V_MN", size)
        # the far-end block must be the mirror image with the same b numbering
        for (i, j), terms in m0.items():
            assert mn[(size + 1 - i, size + 1 - j)] == terms, (p, i, j)
        bs = next(l for l in blk if re.match(r"\s*BS\s*=", l))
        bh = next(l for l in blk if re.match(r"\s*bhinv\s*=", l))
        d2[str(p)] = {
            "size": size,
            "BS": parse_matrix_literal(bs.split("=", 1)[1].strip().rstrip(";"), {})[0],
            "bhinv": parse_matrix_literal(bh.split("=", 1)[1].strip().rstrip(";"), {})[0],
            "closure": [[i, j, m0[(i, j)]] for i in range(1, size + 1) for j in range(1, size + 1)],
        }
    tables["D2var"] = d2
    dst = os.path.join(ROOT, "oracle", "sbp_tables.json")
    with open(dst, "w") as f:
        json.dump(tables, f, indent=0, separators=(",", ":"))
    print("wrote", dst, os.path.getsize(dst), "bytes")
    write_cuda_header(tables, os.path.join(ROOT, "hybridsbp_b200", "csrc", "sbp_tables_gen.h"))


def _f(c):
    if isinstance(c, list):
        v = float(int(c[0])) / float(int(c[1]))
    else:
        v = float(c)
    return v


def _lit(v):
    return float(v).hex() if v != 0 else "0.0"


def write_cuda_header(tables, dst):
    """Dense device tables: first-derivative closures and, for the variable-coefficient
    second derivative, T[i][j][k] with M0[i][j] = sum_k T[i][j][k] * b[k]."""
    out = ["// GENERATED by tools/gen_sbp_tables.py -- do not edit.",
           "// Published SBP coefficient data (Strand 1994; Mattsson & Nordstrom 2004; Mattsson 2012)",
           "// as used by reference diagonal_sbp.jl:69-92 (D1) and :507-690 (variable D2 closures).",
           "// Literals are C99 hex floats so that the device sees exactly the doubles the host computed.",
           "#pragma once", ""]
    for p in (2, 4, 6):
        t = tables["D1"][str(p)]
        bm, bn = len(t["bd"]), len(t["bd"][0])
        out.append("// ---- p = %d first derivative: interior d[%d], closure bd[%d][%d], norm weights hw = 1/bhinv" % (p, p + 1, bm, bn))
        out.append("#define HSBP_D1_BM_%d %d" % (p, bm))
        out.append("#define HSBP_D1_BN_%d %d" % (p, bn))
        out.append("#define HSBP_D1_D_%d {%s}" % (p, ", ".join(_lit(v) for v in t["d"])))
        out.append("#define HSBP_D1_BD_%d {%s}" % (p, ", ".join(_lit(v) for row in t["bd"] for v in row)))
        out.append("#define HSBP_D1_HW_%d {%s}" % (p, ", ".join(_lit(1.0 / v) for v in t["bhinv"])))
        out.append("")
    for p, nk in ((4, 8), (6, 12)):
        t = tables["D2var"][str(p)]
        m = t["size"]
        T = [[[0.0] * nk for _ in range(m)] for _ in range(m)]
        for i, j, terms in t["closure"]:
            for k, c in terms:
                assert T[i - 1][j - 1][k - 1] == 0.0
                T[i - 1][j - 1][k - 1] = _f(c)
        out.append("// ---- p = %d variable-coefficient second derivative closure tensor [%d][%d][%d]" % (p, m, m, nk))
        out.append("#define HSBP_D2_M_%d %d" % (p, m))
        out.append("#define HSBP_D2_NK_%d %d" % (p, nk))
        out.append("#define HSBP_D2_BS_%d {%s}" % (p, ", ".join(_lit(v) for v in t["BS"])))
        out.append("#define HSBP_D2_T_%d {\\" % p)
        for i in range(m):
            for j in range(m):
                out.append("  %s,\\" % ", ".join(_lit(v) for v in T[i][j]))
        out.append("}")
        out.append("")
    with open(dst, "w") as f:
        f.write("\n".join(out) + "\n")
    print("wrote", dst, os.path.getsize(dst), "bytes")


if __name__ == "__main__":
    sys.exit(main())
