#!/usr/bin/env python3
"""Write hybridsbp_b200/csrc/sweep_tables_gen.h: the closure rows of the 1-D SBP operators as
straight-line device code (zero coefficients never reach the compiler) plus the small tables the
line-marching kernel k_sweep.cuh indexes at run time.

Source of the numbers: tools/sbp_coeffs.py (which reads oracle/sbp_tables.json, extracted from the
published tables the reference uses, diagonal_sbp.jl:69-92 and :507-690).  tools/proto_sweep.py
evaluates the same Coeffs objects on the CPU and tests/test_sweep_algorithm.py checks that
emulation against the oracle's assembled operator, so a wrong entry here shows up without a GPU.
"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from sbp_coeffs import Coeffs  # noqa: E402


def lit(v):
    v = float(v)
    return v.hex() if v != 0.0 else "0.0"


def sum_expr(terms, fmt):
    """terms: [(index, coef)] -> 'c0 * x[i0] + c1 * x[i1] ...' evaluated left to right."""
    parts = []
    for k, c in terms:
        parts.append("%s * %s" % (lit(c), fmt % k))
    return " + ".join(parts) if parts else "0.0"


def emit(p, out):
    cf = Coeffs(p)
    H, MC, NK, BM, BN = cf.H, cf.MC, cf.NK, cf.BM, cf.BN
    out.append("// ============================== p = %d ==============================" % p)
    out.append("template <> struct SweepTab<%d> {" % p)
    out.append("  static constexpr int H = %d, MC = %d, NK = %d, BM = %d, BN = %d;" % (H, MC, NK, BM, BN))
    out.append("  // interior first-derivative stencil, offsets +1..+H (antisymmetric): diagonal_sbp.jl:71,75,83")
    for o in range(1, H + 1):
        out.append("  static constexpr double D%d = %s;" % (o, lit(cf.d[H + o])))
    out.append("};")
    out.append("__constant__ double c_sw_hw%d[%d] = {%s};" % (p, BM, ", ".join(lit(v) for v in cf.hw)))
    out.append("__constant__ double c_sw_Qc%d[%d] = {%s};   // Q = H*D closure rows [BM][BN]" %
               (p, BM * BN, ", ".join(lit(v) for v in cf.Qc.reshape(-1))))
    out.append("")

    # ---- second-derivative closure rows, pair form -------------------------------------------
    def pair_coef_expr(i, j):
        """coupling M[i][j], i < j, for a near-end closure row i < MC (b indexable 0..NK-1)."""
        if j < MC:
            return sum_expr(cf.closure_pair[(i, j)], "b[%d]")
        o = j - i
        return sum_expr([(i + sh, c) for sh, c in cf.interior_pair[o]], "b[%d]")

    pairs = []
    for i in range(MC):
        for j in range(i + 1, MC):
            pairs.append((i, j))
        for o in range(1, H + 1):
            if i + o >= MC:
                pairs.append((i, i + o))
    out.append("// all MC closure rows of M(b) u (pair form, diagonal_sbp.jl closure block + interior couplings that")
    out.append("// stick out of it).  b[0..NK-1], u[0..NK-1] are the first NK points counted from the block end.")
    out.append("template <> __device__ __forceinline__ void d2_closure_rows<%d>(const double *b, const double *u, double *out) {" % p)
    for i in range(MC):
        out.append("  double r%d = 0.0;" % i)
    for (i, j) in pairs:
        out.append("  { const double f = (%s) * (u[%d] - u[%d]); r%d += f;%s }" %
                   (pair_coef_expr(i, j), j, i, i, (" r%d -= f;" % j) if j < MC else ""))
    for i in range(MC):
        out.append("  out[%d] = r%d;" % (i, i))
    out.append("}")
    out.append("// one closure row (row is warp-uniform in the caller)")
    out.append("template <> __device__ __forceinline__ double d2_closure_row<%d>(int row, const double *b, const double *u) {" % p)
    out.append("  double r = 0.0;")
    out.append("  switch (row) {")
    for i in range(MC):
        out.append("    case %d:" % i)
        js = [j for j in range(MC) if j != i] + [i + o for o in range(1, H + 1) if i + o >= MC]
        for j in js:
            a, b_ = (i, j) if i < j else (j, i)
            out.append("      r += (%s) * (u[%d] - u[%d]);" % (pair_coef_expr(a, b_), j, i))
        out.append("      break;")
    out.append("  }")
    out.append("  return r;")
    out.append("}")
    out.append("")

    # ---- first-derivative closure rows ----------------------------------------------------------
    out.append("// (Q u)_k, k < BM, from u[0..BN-1]")
    out.append("template <> __device__ __forceinline__ void q_closure_rows<%d>(const double *u, double *out) {" % p)
    for k in range(BM):
        terms = [(j, cf.Qc[k, j]) for j in range(BN) if cf.Qc[k, j] != 0.0]
        out.append("  out[%d] = %s;" % (k, sum_expr(terms, "u[%d]")))
    out.append("}")
    out.append("// (Q^T w)_i, i < BM, from w[0..BN-1]   (rows of Q^T; interior rows of Q contribute their stencil entries)")
    out.append("template <> __device__ __forceinline__ double qt_closure_row<%d>(int row, const double *w) {" % p)
    out.append("  switch (row) {")
    for i in range(BM):
        terms = [(k, cf.QTc[i, k]) for k in range(BN) if cf.QTc[i, k] != 0.0]
        out.append("    case %d: return %s;" % (i, sum_expr(terms, "w[%d]")))
    out.append("  }")
    out.append("  return 0.0;")
    out.append("}")
    out.append("")


def main():
    out = ["// GENERATED by tools/gen_sweep_tables.py -- do not edit.",
           "// Closure rows of the SBP operators as straight-line code for the line-marching kernel (k_sweep.cuh).",
           "// Numbers: oracle/sbp_tables.json via tools/sbp_coeffs.py (reference diagonal_sbp.jl:69-92, 507-690).",
           "#pragma once",
           "",
           "namespace hsbp {",
           "template <int P> struct SweepTab;",
           "template <int P> __device__ __forceinline__ void d2_closure_rows(const double *b, const double *u, double *out);",
           "template <int P> __device__ __forceinline__ double d2_closure_row(int row, const double *b, const double *u);",
           "template <int P> __device__ __forceinline__ void q_closure_rows(const double *u, double *out);",
           "template <int P> __device__ __forceinline__ double qt_closure_row(int row, const double *w);",
           ""]
    for p in (2, 4, 6):
        emit(p, out)
    out.append("}  // namespace hsbp")
    dst = os.path.join(os.path.dirname(HERE), "hybridsbp_b200", "csrc", "sweep_tables_gen.h")
    with open(dst, "w") as f:
        f.write("\n".join(out) + "\n")
    print("wrote", dst, os.path.getsize(dst), "bytes")


if __name__ == "__main__":
    main()
