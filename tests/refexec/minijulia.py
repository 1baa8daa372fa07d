"""A small interpreter for the subset of Julia the reference's solve path is written in.

TEST INFRASTRUCTURE.  Julia is not installed in this image, so the reference cannot be run as it is.  This module parses the
reference's own source files (diagonal_sbp.jl, global_curved.jl, square_circle.jl, read where they lie under /root/reference)
and EXECUTES their statements one by one on a numpy / scipy runtime with Julia's semantics for the constructs those files use:
1-based column-major arrays, ranges, broadcasting with trailing singleton dimensions, sparse matrices built from (I, J, V)
triplets, adjoints, kron, named tuples, closures, keyword arguments, multiple dispatch by arity and simple type annotations.
Nothing of the reference is copied: the text is read at test time and interpreted.

What it gives the tests: the reference's operators (`locoperator`, `gloλoperator`, `assembleλmatrix`, ...) and its
`square_circle.jl` driver evaluated by the reference's own statements, to compare the oracle with (tests/test_reference_executed.py)
and to generate the golden vectors under tests/golden/refexec/ (tools/gen_refexec_golden.py).

Not a general Julia: no type system beyond what dispatch on Number / AbstractArray / String needs, no modules, no macros other
than the handful the files use (@assert @view @views @inbounds @show @.).
"""
import math
import re
import unicodedata

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla


# ======================================================================================================================
# lexer
# ======================================================================================================================
class Tok:
    __slots__ = ("kind", "val", "sp", "line", "sp_after")

    def __init__(self, kind, val, sp, line):
        self.kind, self.val, self.sp, self.line, self.sp_after = kind, val, sp, line, False

    def __repr__(self):
        return "Tok(%s,%r,l%d)" % (self.kind, self.val, self.line)


OPS = sorted([
    "...", ".==", ".!=", ".<=", ".>=", ".+=", ".-=", ".*=", "./=", ".+", ".-", ".*", "./", ".^", ".\\", ".<", ".>", ".=",
    "==", "!=", "<=", ">=", "&&", "||", "+=", "-=", "*=", "/=", "->", "=>", "::", "<:", "|>",
    "+", "-", "*", "/", "\\", "^", "=", "<", ">", "!", "?", ":", ";", ",", "(", ")", "[", "]", "{", "}", ".", "'", "&", "|", "%",
    "⊗", "≈", "∈", "÷", "≤", "≥", "≠", "$",
], key=len, reverse=True)
KEYWORDS = {"function", "end", "if", "elseif", "else", "for", "while", "return", "break", "continue", "let", "begin", "const",
            "using", "import", "struct", "try", "catch", "do", "in", "where", "global", "local", "true", "false", "mutable"}
NUM_RE = re.compile(r"\d+\.\d+(?:[eE][+-]?\d+)?|\d+\.(?=\s|$|[),;\]])|\d+(?:[eE][+-]?\d+)?")


def _id_start(c):
    return c.isalpha() or c == "_" or (ord(c) > 127 and unicodedata.category(c) in ("Sm", "So", "Lo", "Ll", "Lu") and c not in "⊗≈∈÷≤≥≠")


def _id_cont(c):
    return c.isalnum() or c == "_" or unicodedata.category(c) in ("Mn", "Mc", "No") or c in "′" or (
        ord(c) > 127 and unicodedata.category(c) in ("Lo", "Ll", "Lu", "Lm"))


def lex(src):
    toks, i, n, line = [], 0, len(src), 1
    sp = True
    while i < n:
        c = src[i]
        if c == "\n":
            toks.append(Tok("nl", "\n", sp, line)); line += 1; i += 1; sp = True; continue
        if c in " \t\r":
            i += 1; sp = True; continue
        if src.startswith("#=", i):
            depth, j = 1, i + 2
            while depth and j < n:
                if src.startswith("#=", j): depth += 1; j += 2
                elif src.startswith("=#", j): depth -= 1; j += 2
                else:
                    if src[j] == "\n": line += 1
                    j += 1
            i = j; sp = True; continue
        if c == "#":
            while i < n and src[i] != "\n": i += 1
            continue
        if c == '"':
            j = i + 1; buf = []
            while src[j] != '"':
                if src[j] == "\\":
                    buf.append(src[j:j + 2]); j += 2
                else:
                    buf.append(src[j]); j += 1
            toks.append(Tok("str", "".join(buf), sp, line)); i = j + 1; sp = False; continue
        if c.isdigit():
            m = NUM_RE.match(src, i)
            t = m.group(0)
            toks.append(Tok("num", float(t) if ("." in t or "e" in t or "E" in t) else int(t), sp, line)); i = m.end(); sp = False
            continue
        if _id_start(c):
            j = i + 1
            while j < n and (_id_cont(src[j]) or (src[j] == "!" and not src.startswith("!=", j))): j += 1
            w = src[i:j]
            if j < n and src[j] == '"':                                   # r"..." regex literal, other string macros raw
                k = j + 1
                while src[k] != '"':
                    k += 2 if src[k] == "\\" else 1
                if w == "r": toks.append(Tok("regex", src[j + 1:k], sp, line))
                else: toks.append(Tok("strmacro", (w, src[j + 1:k]), sp, line))
                i = k + 1; sp = False; continue
            toks.append(Tok("kw" if w in KEYWORDS else "id", w, sp, line)); i = j; sp = False; continue
        if c == "@":
            j = i + 1
            while j < n and (_id_cont(src[j]) or src[j] == "."): j += 1
            toks.append(Tok("macro", src[i + 1:j], sp, line)); i = j; sp = False; continue
        for op in OPS:
            if src.startswith(op, i):
                if op == "." and i + 1 < n and src[i + 1].isdigit():
                    m = re.compile(r"\.\d+(?:[eE][+-]?\d+)?").match(src, i)
                    toks.append(Tok("num", float(m.group(0)), sp, line)); i = m.end(); sp = False; break
                toks.append(Tok("op", op, sp, line)); i += len(op); sp = False; break
        else:
            raise SyntaxError("minijulia: cannot lex %r at line %d" % (src[i:i + 20], line))
    toks.append(Tok("nl", "\n", True, line)); toks.append(Tok("eof", None, True, line))
    for a, b in zip(toks, toks[1:]):
        a.sp_after = b.sp or b.kind in ("nl", "eof")
    return toks


# ======================================================================================================================
# parser -> tuples
# ======================================================================================================================
ASSIGN_OPS = {"=", "+=", "-=", "*=", "/=", ".=", ".+=", ".-=", ".*=", "./="}
CMP_OPS = {"==", "!=", "<", "<=", ">", ">=", "≈", "∈", "≤", "≥", "≠", ".==", ".!=", ".<", ".<=", ".>", ".>=", "<:"}
PLUS_OPS = {"+", "-", ".+", ".-", "|"}
TIMES_OPS = {"*", "/", "\\", "⊗", ".*", "./", ".\\", "÷", "%", "&"}


class Parser:
    def __init__(self, toks, fname="?"):
        self.t, self.i, self.fname = toks, 0, fname
        self.ctx = ["block"]             # 'block' | 'paren' | 'bracket'  (how newlines / whitespace are read)
        self.tern = [0]
        self.index_depth = 0

    # ---- token helpers
    def peek(self, skip_nl=None):
        if skip_nl is None: skip_nl = self.ctx[-1] == "paren"
        j = self.i
        if skip_nl:
            while self.t[j].kind == "nl": j += 1
        return self.t[j]

    def next(self, skip_nl=None):
        if skip_nl is None: skip_nl = self.ctx[-1] == "paren"
        if skip_nl:
            while self.t[self.i].kind == "nl": self.i += 1
        tok = self.t[self.i]; self.i += 1
        return tok

    def skip_nl(self):
        while self.t[self.i].kind == "nl" or (self.t[self.i].kind == "op" and self.t[self.i].val == ";"): self.i += 1

    def is_op(self, v, tok=None):
        tok = tok or self.peek()
        return tok.kind == "op" and tok.val == v

    def is_kw(self, v, tok=None):
        tok = tok or self.peek()
        return tok.kind == "kw" and tok.val == v

    def expect_op(self, v):
        tok = self.next()
        if not (tok.kind == "op" and tok.val == v):
            raise SyntaxError("minijulia %s:%d: expected %r, got %r" % (self.fname, tok.line, v, tok.val))
        return tok

    def expect_kw(self, v):
        self.skip_nl()
        tok = self.next()
        if not (tok.kind == "kw" and tok.val == v):
            raise SyntaxError("minijulia %s:%d: expected %r, got %r" % (self.fname, tok.line, v, tok.val))

    # ---- statements
    def parse_program(self):
        out = []
        self.skip_nl()
        while self.peek().kind != "eof":
            line = self.peek().line
            out.append(("ln", line, self.fname, self.parse_statement())); self.skip_nl()
        return ("block", out)

    def parse_block(self, terminators=("end",)):
        out = []
        self.ctx.append("block"); self.tern.append(0)
        saved_index = self.index_depth; self.index_depth = 0
        self.skip_nl()
        while not (self.peek().kind == "kw" and self.peek().val in terminators):
            if self.peek().kind == "eof": raise SyntaxError("minijulia %s: unexpected end of file in block" % self.fname)
            line = self.peek().line
            out.append(("ln", line, self.fname, self.parse_statement())); self.skip_nl()
        self.ctx.pop(); self.tern.pop(); self.index_depth = saved_index
        return ("block", out)

    def parse_statement(self):
        tok = self.peek()
        if tok.kind == "kw":
            v = tok.val
            if v == "function": return self.parse_function()
            if v == "for": return self.parse_for()
            if v == "while":
                self.next(); cond = self.parse_expr(); body = self.parse_block(); self.expect_kw("end"); return ("while", cond, body)
            if v == "return":
                self.next()
                if self.peek(False).kind == "nl" or self.is_kw("end", self.peek(False)): return ("return", None)
                return ("return", self.parse_tuple_expr())
            if v == "break": self.next(); return ("break",)
            if v == "continue": self.next(); return ("continue",)
            if v in ("using", "import"):
                self.next()
                while self.peek(False).kind != "nl": self.next(False)
                return ("nop",)
            if v == "const": self.next(); return self.parse_statement()
            if v == "global": self.next(); return ("global", self.parse_statement())
            if v == "local": self.next(); return self.parse_statement()
            if v in ("struct", "mutable"): return self.parse_struct()
        if tok.kind == "macro":
            return self.parse_macro_statement()
        lhs = self.parse_expr()
        if self.is_op(",", self.peek(False)) and self.ctx[-1] == "block":      # a, b = ...   /   a, b  (bare tuple)
            items = [lhs]
            while self.is_op(",", self.peek(False)):
                self.next(False); items.append(self.parse_expr(no_assign=True))
            lhs = ("tuple", items)
            if self.peek(False).kind == "op" and self.peek(False).val in ASSIGN_OPS:
                op = self.next(False).val
                rhs = self.parse_tuple_expr()
                return self.make_assign(op, lhs, rhs)
        return lhs

    def parse_tuple_expr(self):
        e = self.parse_expr()
        if self.is_op(",", self.peek(False)) and self.ctx[-1] == "block":
            items = [e]
            while self.is_op(",", self.peek(False)):
                self.next(False); items.append(self.parse_expr())
            return ("tuple", items)
        return e

    def parse_macro_statement(self):
        tok = self.next()
        name = tok.val
        if name in ("inbounds", "views", "simd", "inline", "noinline", "fastmath"):
            return self.parse_statement() if name != "views" else ("viewsblock", self.parse_statement())
        if name == ".":
            return ("dotmacro", self.parse_statement())
        if name == "assert":
            e = self.parse_expr()
            return ("assert", e, tok.line)
        if name in ("show", "printf", "info", "warn", "time", "plotting"):
            e = self.parse_tuple_expr()
            return ("show", e)
        if name == "view":
            self.i -= 1
            return self.parse_expr()
        raise SyntaxError("minijulia %s:%d: macro @%s is not supported" % (self.fname, tok.line, name))

    def parse_struct(self):
        tok = self.next()
        if tok.val == "mutable": self.next()
        name = self.next().val
        fields = []
        depth = 0
        if self.is_op("{", self.peek(False)): self.skip_braces()
        while self.peek(False).kind != "nl": self.next(False)
        # field lines are `name::Type`; anything else (inner constructors) is skipped up to the struct's own `end`
        while True:
            self.skip_nl()
            t = self.peek()
            if t.kind == "kw" and t.val == "end" and depth == 0:
                self.next(); break
            if t.kind == "id" and self.t[self.i + 1].kind == "op" and self.t[self.i + 1].val == "::" and depth == 0:
                fields.append(t.val)
            while self.peek(False).kind != "nl":
                t = self.next(False)
                if t.kind == "kw" and t.val in ("function", "if", "for", "while", "begin", "let", "try", "do"): depth += 1
                if t.kind == "kw" and t.val == "end": depth -= 1
        return ("struct", name, fields)

    def skip_braces(self):
        self.expect_op("{"); depth = 1
        while depth:
            t = self.next(True)
            if t.kind == "op" and t.val == "{": depth += 1
            if t.kind == "op" and t.val == "}": depth -= 1

    def parse_params(self):
        """( a, b::T, c = default ; kw = default, kws... )  ->  positional [(name, type, default)], keyword [(name, default)]"""
        self.expect_op("(")
        self.ctx.append("paren"); self.tern.append(0)
        pos, kws, in_kw = [], [], False
        while not self.is_op(")"):
            if self.is_op(";"): self.next(); in_kw = True; continue
            if self.is_op(","): self.next(); continue
            if self.is_op("("):                     # destructuring parameter (a, b)
                raise SyntaxError("minijulia: destructuring parameters are not supported")
            name = self.next().val
            typ = None; default = None; splat = False
            if self.is_op("::"):
                self.next(); typ = self.parse_type()
            if self.is_op("..."):
                self.next(); splat = True
            if self.is_op("="):
                self.next(); default = self.parse_expr(no_assign=True)
            if in_kw: kws.append((name, default, splat))
            else: pos.append((name, typ, default, splat))
        self.expect_op(")")
        self.ctx.pop(); self.tern.pop()
        return pos, kws

    def parse_type(self):
        name = self.next().val
        if self.is_op("{", self.peek(False)) and not self.peek(False).sp: self.skip_braces()
        return name

    def skip_where(self):
        while self.is_kw("where", self.peek(True)):
            self.next(True)
            if self.is_op("{", self.peek(False)): self.skip_braces()
            else:
                self.next(False)
                if self.is_op("<:", self.peek(False)): self.next(False); self.parse_type()

    def parse_function(self):
        self.next()
        tok = self.next()
        name = tok.val
        if self.is_op("{", self.peek(False)): self.skip_braces()
        pos, kws = self.parse_params()
        typevars = self.where_types()
        body = self.parse_block(); self.expect_kw("end")
        return ("func", name, pos, kws, body, typevars)

    def where_types(self):
        """`where T <: Number` / `where {T1 <: Number, ...}`  ->  {T: bound}"""
        tv = {}
        while self.is_kw("where", self.peek(True)):
            self.next(True)
            braces = self.is_op("{", self.peek(False))
            if braces: self.next(False)
            while True:
                nm = self.next(True).val; bound = None
                if self.is_op("<:", self.peek(True)): self.next(True); bound = self.parse_type()
                tv[nm] = bound
                if braces and self.is_op(",", self.peek(True)): self.next(True); continue
                break
            if braces: self.ctx.append("paren"); self.expect_op("}"); self.ctx.pop()
        return tv

    def parse_for(self):
        self.next()
        iters = []
        while True:
            var = self.parse_postfix_only()
            t = self.next()
            if not ((t.kind == "op" and t.val in ("=", "∈")) or (t.kind == "kw" and t.val == "in")):
                raise SyntaxError("minijulia %s:%d: bad for header" % (self.fname, t.line))
            iters.append((var, self.parse_expr(no_assign=True)))
            if self.is_op(",", self.peek(False)): self.next(False); continue
            break
        body = self.parse_block(); self.expect_kw("end")
        return ("for", iters, body)

    def parse_postfix_only(self):
        if self.is_op("("):
            return self.parse_primary()
        return ("id", self.next().val)

    # ---- expressions
    def make_assign(self, op, lhs, rhs):
        if op == "=":
            if lhs[0] == "call" and not lhs[4] and lhs[1][0] == "id":                 # f(x) = expr
                pos = []
                for a in lhs[2]:
                    if a[0] == "id": pos.append((a[1], None, None, False))
                    elif a[0] == "typed": pos.append((a[1][1], a[2], None, False))
                    else: raise SyntaxError("minijulia: unsupported short-form parameter %r" % (a,))
                pos += [(k, None, v, False) for k, v, after in lhs[3] if not after]          # f(a, b = 2a) = ...: positional default
                kws = [(v[1], None, True) if k == "..." else (k, v, False) for k, v, after in lhs[3] if after]
                return ("func", lhs[1][1], pos, kws, ("block", [rhs]), {})
            return ("assign", lhs, rhs)
        if op == ".=": return ("dotassign", lhs, rhs)
        if op.startswith("."): return ("dotassign", lhs, ("bin", "." + op[1:-1], lhs, rhs))
        return ("assign", lhs, ("bin", op[:-1], lhs, rhs))

    def parse_expr(self, no_assign=False):
        lhs = self.parse_ternary()
        tok = self.peek(False) if self.ctx[-1] != "paren" else self.peek()
        if not no_assign and tok.kind == "op" and tok.val in ASSIGN_OPS:
            self.next()
            self.skip_nl_after_operator()
            if self.ctx[-1] == "block":
                rhs = self.parse_tuple_expr_or_assign()
            else:
                rhs = self.parse_expr()
            return self.make_assign(tok.val, lhs, rhs)
        if tok.kind == "op" and tok.val == "=>":
            self.next(); rhs = self.parse_ternary(); return ("call", ("id", "Pair"), [lhs, rhs], [], False)
        return lhs

    def parse_tuple_expr_or_assign(self):
        e = self.parse_expr()                      # chained a = b = c
        if self.is_op(",", self.peek(False)) and self.ctx[-1] == "block":
            items = [e]
            while self.is_op(",", self.peek(False)):
                self.next(False); items.append(self.parse_expr())
            return ("tuple", items)
        return e

    def skip_nl_after_operator(self):
        while self.t[self.i].kind == "nl" and self.ctx[-1] != "bracket": self.i += 1

    def parse_ternary(self):
        cond = self.parse_arrow()
        tok = self.peek(False) if self.ctx[-1] != "paren" else self.peek()
        if tok.kind == "op" and tok.val == "?":
            self.next()
            self.tern[-1] += 1
            a = self.parse_expr()
            self.tern[-1] -= 1
            self.skip_nl_after_operator()
            self.expect_op(":")
            self.skip_nl_after_operator()
            b = self.parse_statement_like()
            return ("ternary", cond, a, b)
        return cond

    def parse_statement_like(self):
        t = self.peek()
        if t.kind == "kw" and t.val in ("break", "continue", "return"):
            return self.parse_statement()
        return self.parse_expr()

    def parse_arrow(self):
        lhs = self.parse_or()
        tok = self.peek(False) if self.ctx[-1] != "paren" else self.peek()
        if tok.kind == "op" and tok.val == "->":
            self.next(); self.skip_nl_after_operator()
            if self.is_kw("begin"):
                body = self.parse_primary()
            else:
                body = self.parse_expr(no_assign=False)
            if lhs[0] == "id": params = [lhs[1]]
            elif lhs[0] == "tuple": params = [p[1] for p in lhs[1]]
            elif lhs[0] == "paren": params = [lhs[1][1]]
            else: raise SyntaxError("minijulia: bad lambda parameters %r" % (lhs,))
            return ("lambda", params, body)
        return lhs

    def binary_level(self, sub, ops, node=None):
        lhs = sub()
        while True:
            tok = self.peek(False) if self.ctx[-1] != "paren" else self.peek()
            if tok.kind != "op" or tok.val not in ops: return lhs
            if self.ctx[-1] == "bracket" and tok.sp and not tok.sp_after and tok.val in ("+", "-"):
                return lhs                                   # `[a -b]`: a new element, not a difference
            self.next(); self.skip_nl_after_operator()
            nxt = self.peek()
            if node and nxt.kind == "kw" and nxt.val in ("break", "continue", "return"):
                rhs = self.parse_statement()                 # cond || continue
            else:
                rhs = sub()
            lhs = (node, lhs, rhs) if node else ("bin", tok.val, lhs, rhs)

    def parse_or(self):
        return self.binary_level(self.parse_and, {"||"}, "or")

    def parse_and(self):
        return self.binary_level(self.parse_cmp, {"&&"}, "and")

    def parse_cmp(self):
        first = self.parse_range()
        operands, ops = [first], []
        while True:
            tok = self.peek(False) if self.ctx[-1] != "paren" else self.peek()
            if tok.kind == "op" and tok.val in CMP_OPS: pass
            elif tok.kind == "kw" and tok.val == "in" and self.ctx[-1] != "bracket": pass
            else: break
            self.next(); self.skip_nl_after_operator()
            ops.append("∈" if tok.val == "in" else tok.val); operands.append(self.parse_range())
        return first if not ops else ("cmp", operands, ops)

    def colon_is_range(self, tok):
        if not (tok.kind == "op" and tok.val == ":"): return False
        if self.tern[-1] > 0 and tok.sp: return False
        nxt = self.t[self.i + 1]
        if nxt.kind == "op" and nxt.val in (",", "]", ")"): return False
        return True

    def parse_range(self):
        a = self.parse_plus()
        tok = self.peek(False) if self.ctx[-1] != "paren" else self.peek()
        if self.colon_is_range(tok):
            self.next(); b = self.parse_plus()
            tok = self.peek(False) if self.ctx[-1] != "paren" else self.peek()
            if self.colon_is_range(tok):
                self.next(); c = self.parse_plus()
                return ("range", a, b, c)
            return ("range", a, None, b)
        return a

    def parse_plus(self):
        return self.binary_level(self.parse_times, PLUS_OPS)

    def parse_times(self):
        return self.binary_level(self.parse_unary, TIMES_OPS)

    def parse_unary(self):
        tok = self.peek()
        if tok.kind == "op" and tok.val in ("-", "+", "!", ".-", ".+"):
            self.next()
            e = self.parse_unary()
            return ("un", tok.val.lstrip(".") if tok.val != "!" else "!", e)
        return self.parse_power()

    def parse_power(self):
        base = self.parse_juxt()
        tok = self.peek(False) if self.ctx[-1] != "paren" else self.peek()
        if tok.kind == "op" and tok.val in ("^", ".^"):
            if self.ctx[-1] == "bracket" and tok.sp and not tok.sp_after: return base
            self.next()
            ex = self.parse_unary_power()
            return ("bin", tok.val, base, ex)
        return base

    def parse_unary_power(self):
        tok = self.peek()
        if tok.kind == "op" and tok.val in ("-", "+"):
            self.next(); return ("un", tok.val, self.parse_unary_power())
        return self.parse_power()

    def parse_juxt(self):
        """numeric literal or parenthesised expression directly followed by an identifier / parenthesis: implicit product"""
        e = self.parse_postfix()
        tok = self.peek(False)
        if not tok.sp and tok.kind in ("id",) and e[0] in ("num", "paren"):
            rhs = self.parse_power()
            return ("bin", "*", e, rhs)
        if not tok.sp and tok.kind == "op" and tok.val == "(" and e[0] == "num":
            rhs = self.parse_power()
            return ("bin", "*", e, rhs)
        return e

    def parse_postfix(self):
        e = self.parse_primary()
        while True:
            tok = self.peek(False)
            if tok.kind != "op" or tok.sp:                    # call / index / field / adjoint need adjacency
                return self.maybe_do(e)
            if tok.val == "(" and e[0] != "num":
                args, kwargs = self.parse_call_args()
                e = ("call", e, args, kwargs, False)
            elif tok.val == "[":
                self.next(False)
                self.ctx.append("paren"); self.tern.append(0); self.index_depth += 1
                idxs = []
                while not self.is_op("]"):
                    idxs.append(self.parse_expr(no_assign=True))
                    if self.is_op(","): self.next()
                self.expect_op("]")
                self.ctx.pop(); self.tern.pop(); self.index_depth -= 1
                e = ("index", e, idxs)
            elif tok.val == ".":
                nxt = self.t[self.i + 1]
                if nxt.kind == "op" and nxt.val == "(":
                    self.next(False)
                    args, kwargs = self.parse_call_args()
                    e = ("call", e, args, kwargs, True)
                elif nxt.kind in ("id", "kw") and not nxt.sp:
                    self.next(False); self.next(False)
                    e = ("field", e, nxt.val)
                elif nxt.kind == "strmacro" and not nxt.sp:
                    self.next(False); self.next(False)
                    e = ("str", nxt.val[1])
                else:
                    return self.maybe_do(e)
            elif tok.val == "'":
                self.next(False); e = ("adj", e)
            elif tok.val == "{" and e[0] in ("id", "field"):
                self.next(False)
                self.ctx.append("paren"); self.tern.append(0)
                params = []
                while not self.is_op("}"):
                    params.append(self.parse_expr(no_assign=True))
                    if self.is_op(","): self.next()
                self.expect_op("}")
                self.ctx.pop(); self.tern.pop()
                e = ("curly", e, params)
            elif tok.val == "::":
                self.next(False); typ = self.parse_type(); e = ("typed", e, typ)
            else:
                return self.maybe_do(e)

    def maybe_do(self, e):
        tok = self.peek(False)
        if tok.kind == "kw" and tok.val == "do" and e[0] == "call":
            self.next(False)
            params = []
            while self.peek(False).kind != "nl":
                t = self.next(False)
                if t.kind == "id": params.append(t.val)
            body = self.parse_block(); self.expect_kw("end")
            return ("call", e[1], [("lambda", params, body)] + e[2], e[3], e[4])
        return e

    def parse_call_args(self):
        self.expect_op("(")
        self.ctx.append("paren"); self.tern.append(0)
        saved_index = self.index_depth; self.index_depth = 0
        args, kwargs, in_kw = [], [], False
        while not self.is_op(")"):
            if self.is_op(";"): self.next(); in_kw = True; continue
            e = self.parse_expr()
            if self.is_op("..."):
                self.next(); e = ("splat", e)
            if self.is_kw("for"):                                   # generator argument
                e = self.parse_generator(e)
            if e[0] == "assign" and e[1][0] == "id": kwargs.append((e[1][1], e[2], in_kw))
            elif in_kw and e[0] == "id": kwargs.append((e[1], e, True))
            elif in_kw and e[0] == "splat": kwargs.append(("...", e[1], True))
            else: args.append(e)
            if self.is_op(","): self.next()
        self.expect_op(")")
        self.ctx.pop(); self.tern.pop(); self.index_depth = saved_index
        return args, kwargs

    def parse_generator(self, e):
        self.next()
        var = self.parse_postfix_only()
        t = self.next()
        it = self.parse_expr(no_assign=True)
        return ("comp", e, var, it)

    def parse_primary(self):
        tok = self.next()
        k, v = tok.kind, tok.val
        if k == "num": return ("num", v)
        if k == "str": return ("str", v)
        if k == "regex": return ("regex", v)
        if k == "strmacro": return ("str", v[1])
        if k == "macro":
            if v == "view":
                e = self.parse_postfix()
                return ("view", e)
            if v == "views":
                return ("viewsblock", self.parse_expr())
            if v == "show":
                return ("show", self.parse_expr())
            raise SyntaxError("minijulia %s:%d: macro @%s in an expression" % (self.fname, tok.line, v))
        if k == "id": return ("id", v)
        if k == "kw":
            if v == "true": return ("num", True)
            if v == "false": return ("num", False)
            if v == "end" and self.index_depth > 0: return ("endidx",)
            if v == "begin":
                body = self.parse_block(); self.expect_kw("end"); return body
            if v == "let":
                binds = []
                while self.peek(False).kind != "nl":                  # let a = 1, b = 2
                    binds.append(self.parse_expr())
                    if self.is_op(",", self.peek(False)): self.next(False)
                body = self.parse_block(); self.expect_kw("end")
                return ("let", ("block", binds + body[1]))
            if v == "if": return self.parse_if()
            if v == "try":
                body = self.parse_block(("catch", "end"))
                cbody, cvar = ("block", []), None
                self.skip_nl()
                if self.is_kw("catch"):
                    self.next()
                    if self.peek(False).kind == "id" and self.t[self.i + 1].kind == "nl":
                        cvar = self.next(False).val
                    cbody = self.parse_block()
                self.expect_kw("end")
                return ("try", body, cvar, cbody)
            if v == "function":
                self.i -= 1; return self.parse_function()
            if v == "for":
                self.i -= 1; return self.parse_for()
        if k == "op":
            if v == "(": return self.parse_paren()
            if v == "[": return self.parse_bracket()
            if v == ":":
                nxt = self.peek(False)
                if nxt.kind in ("id", "kw") and not nxt.sp:
                    self.next(False); return ("sym", nxt.val)
                return ("colon",)
            if v == "⊗" and self.is_op("(", self.peek(False)):
                return ("id", v)
            if v == "$":
                return self.parse_primary()
        raise SyntaxError("minijulia %s:%d: unexpected token %r" % (self.fname, tok.line, v))

    def parse_if(self):
        branches, els = [], None
        cond = self.parse_expr(); body = self.parse_block(("elseif", "else", "end")); branches.append((cond, body))
        while True:
            self.skip_nl()
            t = self.next()
            if t.val == "elseif":
                cond = self.parse_expr(); body = self.parse_block(("elseif", "else", "end")); branches.append((cond, body))
            elif t.val == "else":
                els = self.parse_block(); self.expect_kw("end"); break
            else: break
        return ("if", branches, els)

    def parse_paren(self):
        self.ctx.append("paren"); self.tern.append(0)
        saved_index = self.index_depth; self.index_depth = 0
        try:
            if self.is_op(")"):
                self.next(); return ("tuple", [])
            items, trailing = [], False
            while True:
                e = self.parse_expr()
                if self.is_op("..."): self.next(); e = ("splat", e)
                if self.is_kw("for"): e = self.parse_generator(e)
                items.append(e)
                if self.is_op(","):
                    self.next(); trailing = True
                    if self.is_op(")"): break
                    continue
                if self.is_op(";"):
                    self.next(); continue
                break
            self.expect_op(")")
            if len(items) == 1 and not trailing:
                return ("paren", items[0])
            if all(it[0] == "assign" and it[1][0] == "id" for it in items):
                return ("ntuple", [(it[1][1], it[2]) for it in items])
            return ("tuple", items)
        finally:
            self.ctx.pop(); self.tern.pop(); self.index_depth = saved_index

    def parse_bracket(self):
        """[a, b]  |  [a b; c d] (rows by ';' or newline, columns by blanks)  |  [f(x) for x in xs]"""
        self.ctx.append("bracket"); self.tern.append(0)
        saved_index = self.index_depth; self.index_depth = 0
        try:
            while self.peek(False).kind == "nl": self.next(False)
            if self.is_op("]", self.peek(False)):
                self.next(False); return ("vect", [])
            first = self.parse_expr(no_assign=True)
            if self.is_kw("for", self.peek(True)):
                self.ctx[-1] = "paren"
                self.next()
                var = self.parse_postfix_only(); self.next(); it = self.parse_expr(no_assign=True)
                self.expect_op("]")
                return ("comp", first, var, it)
            if self.is_op(",", self.peek(True)):
                self.ctx[-1] = "paren"
                items = [first]
                while self.is_op(","):
                    self.next()
                    if self.is_op("]"): break
                    items.append(self.parse_expr(no_assign=True))
                self.expect_op("]")
                return ("vect", items)
            rows, row = [], [first]
            while True:
                tok = self.peek(False)
                if tok.kind == "op" and tok.val == "]":
                    self.next(False); rows.append(row); break
                if tok.kind == "nl" or (tok.kind == "op" and tok.val == ";"):
                    self.next(False)
                    while self.peek(False).kind == "nl" or self.is_op(";", self.peek(False)): self.next(False)
                    if row: rows.append(row); row = []
                    continue
                row.append(self.parse_expr(no_assign=True))
            rows = [r for r in rows if r]
            if len(rows) == 1 and len(rows[0]) == 1:
                return ("vect", rows[0])
            return ("matrix", rows)
        finally:
            self.ctx.pop(); self.tern.pop(); self.index_depth = saved_index


def parse_source(src, fname="?"):
    return Parser(lex(src), fname).parse_program()


# ======================================================================================================================
# runtime values
# ======================================================================================================================
class JuliaError(Exception):
    pass


class BreakEx(BaseException):
    pass


class ContinueEx(BaseException):
    pass


class ReturnEx(BaseException):
    def __init__(self, value):
        self.value = value


class JRange:
    """a:b / a:s:b with integer (or float) bounds, 1-based like everything else"""
    def __init__(self, start, step, stop):
        self.start, self.step = start, step
        n = int(math.floor((stop - start) / step)) + 1 if (stop - start) * step >= 0 else 0
        self.n = max(n, 0)
        self.stop = start + (self.n - 1) * step

    def arr(self):
        if isinstance(self.start, (int, np.integer)) and isinstance(self.step, (int, np.integer)):
            return np.arange(self.n, dtype=np.int64) * int(self.step) + int(self.start)
        return self.start + np.arange(self.n) * self.step

    def __len__(self):
        return self.n

    def __iter__(self):
        return iter(self.arr().tolist())

    def __repr__(self):
        return "JRange(%r:%r:%r)" % (self.start, self.step, self.stop)


class RowVec(np.ndarray):
    """adjoint of a vector (1 x n): RowVec * vector is a scalar, RowVec' is the vector again"""
    pass


class NT:
    """named tuple / struct instance"""
    def __init__(self, names, values):
        self._names, self._values = list(names), list(values)

    def get(self, name):
        return self._values[self._names.index(name)]

    def __iter__(self):
        return iter(self._values)

    def __len__(self):
        return len(self._values)


class JType:
    def __init__(self, name, conv, dtype):
        self.name, self.conv, self.dtype = name, conv, dtype

    def __call__(self, x):
        return self.conv(x)

    def __eq__(self, o):
        return isinstance(o, JType) and o.name == self.name

    def __hash__(self):
        return hash(self.name)

    def __repr__(self):
        return self.name


Int64 = JType("Int64", lambda x: int(x), np.int64)
Float64 = JType("Float64", lambda x: float(x), np.float64)
BoolT = JType("Bool", lambda x: bool(x), np.bool_)


class TypeSpec:
    def __init__(self, name, params):
        self.name, self.params = name, tuple(params)

    def __eq__(self, o):
        return isinstance(o, TypeSpec) and (self.name, self.params) == (o.name, o.params)

    def __hash__(self):
        return hash((self.name, self.params))

    def __repr__(self):
        return "%s{%s}" % (self.name, ",".join(map(repr, self.params)))


class Factor:
    """what `cholesky(Symmetric(A))` returns: solves with A (sparse LU of the SPD matrix; the adjoint is itself)"""
    def __init__(self, A):
        self.A = sp.csc_matrix(A, dtype=np.float64)
        self.n = self.A.shape[0]
        self.lu = spla.splu(self.A) if self.n > 1 else None

    def solve(self, b):
        if sp.issparse(b): b = b.toarray()
        b = np.asarray(b, dtype=np.float64)
        if self.n == 1: return b / self.A[0, 0]
        return self.lu.solve(np.ascontiguousarray(b))


class Stub:
    """stands for plotting packages: every attribute and call gives another Stub"""
    def __getattr__(self, k):
        return Stub()

    def __call__(self, *a, **k):
        return Stub()


class JFunction:
    def __init__(self, name, interp=None):
        self.name, self.methods, self.interp = name, [], interp

    def __call__(self, *args, **kwargs):
        return self.interp.apply(self, list(args), kwargs)

    def __repr__(self):
        return "<julia function %s, %d methods>" % (self.name, len(self.methods))


class Closure:
    """x -> body: parameters, body and the environment it was written in"""
    def __init__(self, interp, params, body, env):
        self.interp, self.params, self.body, self.env = interp, params, body, env

    def __call__(self, *args):
        return self.interp.apply(self, list(args))


IDENT = object()
UNDEF = object()
COLON = object()


def is_number(x):
    return isinstance(x, (int, float, bool, np.number, np.bool_)) and not isinstance(x, np.ndarray)


def is_arraylike(x):
    return isinstance(x, (np.ndarray, JRange)) or sp.issparse(x)


def to_arr(x):
    if isinstance(x, JRange): return x.arr()
    if isinstance(x, RowVec): return x
    if isinstance(x, (list, tuple)) and all(is_number(v) for v in x): return np.array(x)
    return x


def plain(x):
    return x.view(np.ndarray) if isinstance(x, RowVec) else x


def csc(x):
    return sp.csc_matrix(x, dtype=np.float64)


def align(a, b):
    """Julia broadcasting: missing dimensions are TRAILING singletons"""
    a, b = to_arr(a), to_arr(b)
    if sp.issparse(a): a = a.toarray()
    if sp.issparse(b): b = b.toarray()
    if isinstance(a, np.ndarray) and isinstance(b, np.ndarray) and a.ndim != b.ndim:
        nd = max(a.ndim, b.ndim)
        if a.ndim < nd: a = plain(a).reshape(a.shape + (1,) * (nd - a.ndim))
        if b.ndim < nd: b = plain(b).reshape(b.shape + (1,) * (nd - b.ndim))
    return a, b


def jl_add(a, b, sign=1):
    a, b = to_arr(a), to_arr(b)
    if sp.issparse(a) or sp.issparse(b):
        if is_number(a) or is_number(b): raise JuliaError("sparse + number")
        r = (csc(a) + csc(b)) if sign > 0 else (csc(a) - csc(b))
        return csc(r)
    if isinstance(a, np.ndarray) and isinstance(b, np.ndarray) and a.shape != b.shape:
        raise JuliaError("dimension mismatch in +/-: %s vs %s" % (a.shape, b.shape))
    return a + b if sign > 0 else a - b


def jl_mul(a, b):
    a, b = to_arr(a), to_arr(b)
    if is_number(a) or is_number(b):
        r = a * b
        return csc(r) if sp.issparse(r) else r
    if isinstance(a, Factor) or isinstance(b, Factor): raise JuliaError("product with a factorization")
    if sp.issparse(a) and sp.issparse(b): return csc(a @ b)
    if sp.issparse(a):
        r = a @ plain(b)
        return np.asarray(r)
    if sp.issparse(b):
        if isinstance(a, RowVec): return np.asarray(plain(a) @ b).reshape(1, -1).view(RowVec)
        if a.ndim == 1: raise JuliaError("vector * sparse matrix")
        return np.asarray(a @ b)
    if isinstance(a, RowVec) and b.ndim == 1:
        return (plain(a).reshape(-1) @ b).item()
    if a.ndim == 1 and b.ndim == 2:
        if b.shape[0] != 1: raise JuliaError("vector * matrix with more than one row")
        return plain(a).reshape(-1, 1) @ plain(b)
    r = plain(a) @ plain(b)
    if isinstance(a, RowVec): r = r.view(RowVec)
    return r


def jl_div(a, b):
    a, b = to_arr(a), to_arr(b)
    if is_number(b):
        if is_number(a): return float(a) / float(b) if not isinstance(a, complex) else a / b
        r = a / float(b)
        return csc(r) if sp.issparse(r) else r
    raise JuliaError("unsupported / between %s and %s" % (type(a), type(b)))


def jl_ldiv(a, b):
    if isinstance(a, Factor): return a.solve(to_arr(b))
    if sp.issparse(a): return Factor(a).solve(to_arr(b))
    if is_number(a): return to_arr(b) / a
    return np.linalg.solve(plain(a), plain(to_arr(b)))


def jl_pow(a, b):
    if is_number(a) and is_number(b):
        if isinstance(a, (int, np.integer)) and isinstance(b, (int, np.integer)) and not isinstance(a, (bool, np.bool_)):
            if b < 0: raise JuliaError("integer to a negative power")
            return int(a) ** int(b)
        return float(a) ** b
    raise JuliaError("matrix power is not supported")


def jl_adj(a):
    a = to_arr(a)
    if isinstance(a, Factor) or is_number(a): return a
    if sp.issparse(a): return csc(a.T)
    if isinstance(a, RowVec): return plain(a).reshape(-1)
    if a.ndim == 1: return a.reshape(1, -1).view(RowVec)
    return a.T


def jl_kron(a, b):
    a, b = to_arr(a), to_arr(b)
    if sp.issparse(a) or sp.issparse(b):
        a2 = csc(plain(a).reshape(-1, 1)) if isinstance(a, np.ndarray) and a.ndim == 1 else csc(a)
        b2 = csc(plain(b).reshape(-1, 1)) if isinstance(b, np.ndarray) and b.ndim == 1 else csc(b)
        return csc(sp.kron(a2, b2, format="csc"))
    a, b = plain(a), plain(b)
    if a.ndim == 1 and b.ndim == 1: return np.kron(a, b)
    a2 = a.reshape(-1, 1) if a.ndim == 1 else a
    b2 = b.reshape(-1, 1) if b.ndim == 1 else b
    return np.asfortranarray(np.kron(a2, b2))


def jl_isapprox(a, b):
    a, b = to_arr(a), to_arr(b)
    if is_number(a) and is_number(b):
        return abs(a - b) <= math.sqrt(np.finfo(float).eps) * max(abs(a), abs(b))
    if sp.issparse(a) or sp.issparse(b):
        a, b = csc(a), csc(b)
        d = spla.norm(a - b); na, nb = spla.norm(a), spla.norm(b)
    else:
        a, b = plain(np.asarray(a, dtype=float)), plain(np.asarray(b, dtype=float))
        if a.shape != b.shape: return False
        d = np.linalg.norm(a - b); na, nb = np.linalg.norm(a), np.linalg.norm(b)
    return bool(d <= math.sqrt(np.finfo(float).eps) * max(na, nb))


def jl_equal(a, b):
    a, b = to_arr(a), to_arr(b)
    if isinstance(a, np.ndarray) or isinstance(b, np.ndarray):
        a, b = np.asarray(a), np.asarray(b)
        return a.shape == b.shape and bool(np.all(a == b))
    if isinstance(a, tuple) and isinstance(b, tuple):
        return len(a) == len(b) and all(jl_equal(x, y) for x, y in zip(a, b))
    return a == b


def flat_f(a):
    """column-major vector of an array (a view when the layout allows)"""
    a = plain(a)
    return a if a.ndim == 1 else a.reshape(-1, order="F")


def idx_positions(ix, n):
    """one index of a getindex / setindex call -> (0-based positions or a slice, shape it contributes: None = dropped dim)"""
    if ix is COLON: return slice(None), (n,)
    if isinstance(ix, JRange):
        if ix.n == 0: return np.zeros(0, dtype=np.int64), (0,)
        if isinstance(ix.start, (int, np.integer)) and isinstance(ix.step, (int, np.integer)):
            lo, st = int(ix.start) - 1, int(ix.step)
            hi = lo + st * ix.n
            if lo < 0 or max(lo, lo + st * (ix.n - 1)) >= n: raise JuliaError("BoundsError: range %r of %d" % (ix, n))
            return slice(lo, hi if hi >= 0 else None, st), (ix.n,)
        ix = ix.arr()
    if is_number(ix):
        if isinstance(ix, (float, np.floating)):
            if ix != int(ix): raise JuliaError("non-integer index %r" % ix)
        k = int(ix) - 1
        if k < 0 or k >= n: raise JuliaError("BoundsError: index %d of %d" % (k + 1, n))
        return k, None
    ix = np.asarray(to_arr(ix))
    if ix.dtype == np.bool_:
        return np.nonzero(flat_f(ix))[0], (int(ix.sum()),)
    pos = flat_f(ix).astype(np.int64) - 1
    if pos.size and (pos.min() < 0 or pos.max() >= n): raise JuliaError("BoundsError: index array out of 1:%d" % n)
    if ix.ndim == 1 and pos.size >= 2:                       # an arithmetic progression (a .+ (1:n)): a slice, so that views alias
        st = int(pos[1] - pos[0])
        if st != 0 and np.all(np.diff(pos) == st):
            hi = int(pos[-1]) + st
            return slice(int(pos[0]), hi if hi >= 0 else None, st), ix.shape
    return pos, ix.shape


def jl_getindex(a, idxs, view=False):
    if isinstance(a, (tuple, list, NT)):
        (ix,) = idxs
        vals = a._values if isinstance(a, NT) else a
        if isinstance(ix, JRange) or isinstance(ix, np.ndarray):
            out = [vals[int(k) - 1] for k in to_arr(ix).tolist()]
            return tuple(out) if isinstance(a, tuple) else out
        if ix is COLON: return a
        if not 1 <= int(ix) <= len(vals): raise JuliaError("BoundsError: index %d of %d" % (ix, len(vals)))
        return vals[int(ix) - 1]
    if isinstance(a, dict):
        (ix,) = idxs
        k = _key(ix)
        if k not in a: raise JuliaError("KeyError: %r" % (k,))
        return a[k]
    if isinstance(a, str):
        (ix,) = idxs
        return a[int(ix) - 1]
    if isinstance(a, JRange):
        (ix,) = idxs
        if is_number(ix):
            if not 1 <= ix <= a.n: raise JuliaError("BoundsError on a range")
            return a.start + (int(ix) - 1) * a.step
        return jl_getindex(a.arr(), idxs)
    if sp.issparse(a):
        if len(idxs) == 1:
            return jl_getindex(a.toarray(), idxs)
        p0, s0 = idx_positions(idxs[0], a.shape[0]); p1, s1 = idx_positions(idxs[1], a.shape[1])
        if s0 is None and s1 is None: return a[p0, p1]
        r = a.tocsr()[p0 if s0 is not None else [p0], :].tocsc()[:, p1 if s1 is not None else [p1]]
        if s0 is None or s1 is None: return np.asarray(r.toarray()).reshape(-1)
        return csc(r)
    if is_number(a) and all(is_number(i) and int(i) == 1 for i in idxs):
        return a                                           # numbers index like one-element collections
    if not isinstance(a, np.ndarray):
        raise JuliaError("cannot index a %s" % type(a).__name__)
    a = plain(a)
    if len(idxs) == 1:
        f = flat_f(a)
        pos, shp = idx_positions(idxs[0], f.shape[0])
        r = f[pos]
        if shp is None: return r.item() if isinstance(r, np.generic) and a.dtype != object else r
        if not view or not isinstance(pos, slice): r = np.array(r)
        return r.reshape(shp, order="F") if len(shp) > 1 else r
    if len(idxs) != a.ndim:
        if len(idxs) > a.ndim and all(is_number(i) and int(i) == 1 for i in idxs[a.ndim:]):
            idxs = idxs[:a.ndim]
        else:
            raise JuliaError("%d indices into a %d-dimensional array" % (len(idxs), a.ndim))
    ps = [idx_positions(ix, a.shape[d]) for d, ix in enumerate(idxs)]
    if all(isinstance(p, (slice, int)) for p, _ in ps):
        r = a[tuple(p for p, _ in ps)]
        if all(s is None for _, s in ps): return r.item() if a.dtype != object else r
        return r if view else np.array(r, order="F")
    grid = np.ix_(*[np.atleast_1d(np.arange(a.shape[d])[p]) for d, (p, _) in enumerate(ps)])
    r = a[grid]
    shape = ()
    for _, s in ps:
        if s is not None: shape += tuple(s)
    return np.asfortranarray(r.reshape(shape, order="F")) if shape else r.item()


def _key(k):
    if isinstance(k, tuple): return tuple(_key(v) for v in k)
    if isinstance(k, (np.integer,)): return int(k)
    return k


def jl_setindex(a, idxs, v):
    v = to_arr(v)
    if isinstance(a, list):
        (ix,) = idxs
        a[int(ix) - 1] = v; return
    if isinstance(a, dict):
        (ix,) = idxs
        a[_key(ix)] = v; return
    if not isinstance(a, np.ndarray):
        raise JuliaError("cannot assign into a %s" % type(a).__name__)
    if sp.issparse(v): v = v.toarray()
    a = plain(a)
    if isinstance(v, np.ndarray): v = plain(v)
    if len(idxs) == 1:
        if a.ndim == 1:
            pos, shp = idx_positions(idxs[0], a.shape[0])
            if isinstance(v, np.ndarray):
                if shp is None:
                    if v.size != 1: raise JuliaError("array assigned to one element")
                    v = v.reshape(()).item()
                else:
                    if v.size != int(np.prod(shp)): raise JuliaError("DimensionMismatch in assignment: %s into %s" % (v.shape, shp))
                    v = flat_f(v)
            a[pos] = v
            return
        pos, shp = idx_positions(idxs[0], a.size)
        if isinstance(pos, slice): pos = np.arange(a.size)[pos]
        sub = np.unravel_index(pos, a.shape, order="F")
        a[sub] = flat_f(v) if isinstance(v, np.ndarray) else v
        return
    ps = [idx_positions(ix, a.shape[d]) for d, ix in enumerate(idxs)]
    if all(isinstance(p, (slice, int)) for p, _ in ps):
        tgt = a[tuple(p for p, _ in ps)]
        if isinstance(v, np.ndarray):
            if isinstance(tgt, np.ndarray):
                if v.size != tgt.size: raise JuliaError("DimensionMismatch in assignment")
                v = v.reshape(tgt.shape, order="F")
            else:
                v = v.reshape(()).item()
        a[tuple(p for p, _ in ps)] = v
        return
    grid = np.ix_(*[np.atleast_1d(np.arange(a.shape[d])[p]) for d, (p, _) in enumerate(ps)])
    a[grid] = v.reshape(a[grid].shape, order="F") if isinstance(v, np.ndarray) else v


def make_range(start, stop, length):
    """range(a, stop = b, length = n): Julia evaluates it in twice-precision arithmetic, i.e. correctly rounded"""
    from fractions import Fraction
    a, b = Fraction(start), Fraction(stop)
    n = int(length)
    if n == 1: return np.array([float(a)])
    return np.array([float(a + (b - a) * Fraction(i, n - 1)) for i in range(n)])


def jl_sparse(*args):
    if len(args) == 1:
        x = to_arr(args[0])
        return csc(x if sp.issparse(x) else np.atleast_2d(x))
    if args[0] is IDENT:
        return csc(sp.identity(int(args[1]), format="csc"))
    I = np.asarray(to_arr(args[0])).astype(np.int64).reshape(-1) - 1
    J = np.asarray(to_arr(args[1])).astype(np.int64).reshape(-1) - 1
    V = to_arr(args[2])
    V = np.full(I.shape, float(V)) if is_number(V) else flat_f(np.asarray(V, dtype=np.float64))
    if not (I.size == J.size == V.size): raise JuliaError("sparse(I, J, V): lengths differ")
    if len(args) >= 5: m, n = int(args[3]), int(args[4])
    else: m, n = (int(I.max()) + 1, int(J.max()) + 1) if I.size else (0, 0)
    if I.size and (I.min() < 0 or J.min() < 0 or I.max() >= m or J.max() >= n): raise JuliaError("sparse: index out of range")
    return sp.coo_matrix((V, (I, J)), shape=(m, n)).tocsc()


def jl_findnz(A):
    A = csc(A); A.sort_indices()
    c = A.tocoo()
    order = np.lexsort((c.row, c.col))
    return (c.row[order].astype(np.int64) + 1, c.col[order].astype(np.int64) + 1, c.data[order].copy())


def _dims(args):
    if len(args) == 1 and isinstance(args[0], tuple): args = args[0]
    return tuple(int(d) for d in args)


def jl_zeros(*args, fill=0):
    dtype = np.float64
    if args and isinstance(args[0], JType):
        dtype, args = args[0].dtype, args[1:]
    return np.full(_dims(args), fill, dtype=dtype, order="F")


def jl_fill(v, *dims):
    dt = np.int64 if isinstance(v, (int, np.integer)) and not isinstance(v, bool) else (np.bool_ if isinstance(v, (bool, np.bool_)) else np.float64)
    return np.full(_dims(dims), v, dtype=dt, order="F")


def jl_reshape(a, *dims):
    a = to_arr(a)
    return np.asfortranarray(plain(a).reshape(_dims(dims), order="F"))


def jl_length(a):
    if isinstance(a, (np.ndarray,)): return int(a.size)
    if isinstance(a, NT): return len(a._values)
    if sp.issparse(a): return int(a.shape[0] * a.shape[1])
    return len(a)


def jl_size(a, d=None):
    a = to_arr(a)
    shp = tuple(int(s) for s in a.shape)
    if d is None: return shp
    return shp[int(d) - 1] if int(d) <= len(shp) else 1


def reduce_all(fn):
    def f(a, *rest):
        if rest: raise JuliaError("reduction with extra arguments is not supported")
        a = to_arr(a)
        if sp.issparse(a): a = a.toarray()
        if isinstance(a, (tuple, list)): a = np.asarray(a)
        r = fn(a)
        return r.item() if isinstance(r, np.generic) else r
    return f


def elementwise(fn, nargs=1):
    def f(*args):
        if len(args) == 2:
            a, b = align(*args)
            return fn(a, b)
        (a,) = args
        a = to_arr(a)
        if sp.issparse(a): a = a.toarray()
        return fn(a)
    f.elementwise = True
    return f


def jl_diag(A):
    if sp.issparse(A): return np.asarray(A.diagonal()).copy()
    return np.array(np.diag(plain(A)))


def jl_diagonal(v):
    v = to_arr(v)
    if sp.issparse(v): return csc(sp.diags(v.diagonal()))
    return csc(sp.diags(np.asarray(flat_f(v), dtype=np.float64)))


def jl_vector(x):
    x = to_arr(x)
    if sp.issparse(x): x = x.toarray()
    return np.array(flat_f(x))


def jl_matrix(x):
    x = to_arr(x)
    if sp.issparse(x): return np.asfortranarray(x.toarray())
    return np.array(plain(x), order="F")


def jl_collect(x):
    if isinstance(x, JRange): return x.arr()
    if isinstance(x, np.ndarray): return np.array(x)
    vals = list(x)
    return np.array(vals) if all(is_number(v) for v in vals) else vals


def jl_flatten(x):
    out = []
    for it in x: out.extend(list(it))
    return out


def jl_string(*a):
    return "".join(str(v) for v in a)


def jl_error(*a):
    raise JuliaError(jl_string(*a))


def jl_split(s, pat=None, keepempty=True, limit=0):
    parts = pat.split(s) if hasattr(pat, "split") else (s.split(pat) if pat is not None else s.split())
    return [p for p in parts if keepempty or p != ""]


def jl_parse(T, s):
    try:
        return T(s.strip()) if T is not Int64 else int(s.strip())
    except ValueError as e:
        raise JuliaError("ArgumentError: cannot parse %r as %s" % (s, T))


def jl_view(a, *idxs):
    return jl_getindex(a, list(idxs), view=True)


def jl_typeof(x):
    if isinstance(x, np.ndarray):
        el = "Int64" if x.dtype.kind == "i" else ("Bool" if x.dtype.kind == "b" else ("Float64" if x.dtype.kind == "f" else "Any"))
        return TypeSpec("Array", (JType(el, None, None), x.ndim))
    if isinstance(x, (bool, np.bool_)): return BoolT
    if isinstance(x, (int, np.integer)): return Int64
    if isinstance(x, (float, np.floating)): return Float64
    return type(x)


def jl_hvcat(rows):
    """[a b; c d]: rows of blank-separated blocks"""
    def hcat(items):
        items = [to_arr(x) for x in items]
        if len(items) == 1: return items[0], True
        if any(sp.issparse(x) for x in items):
            return csc(sp.hstack([csc(x) if sp.issparse(x) else csc(np.atleast_2d(plain(x)).T if x.ndim == 1 else plain(x)) for x in items])), False
        if all(is_number(x) for x in items): return np.array(items).reshape(1, -1), False
        cols = []
        for x in items:
            x = np.asarray(plain(x)) if not is_number(x) else np.array([[x]])
            cols.append(x.reshape(-1, 1) if x.ndim == 1 else x)
        return np.asfortranarray(np.hstack(cols)), False
    hs = [hcat(r) for r in rows]
    if len(hs) == 1: return hs[0][0]
    vals = [h for h, _ in hs]
    if any(sp.issparse(x) for x in vals):
        return csc(sp.vstack([csc(x) for x in vals]))
    if all(is_number(x) or (isinstance(x, np.ndarray) and x.ndim == 1) for x in vals):
        parts = [np.atleast_1d(x) for x in vals]
        if not parts: return np.zeros(0)
        if all(p.dtype == object for p in parts): return np.concatenate(parts)
        return np.concatenate(parts)
    mats = []
    for x in vals:
        x = np.array([[x]]) if is_number(x) else np.asarray(plain(x))
        mats.append(x.reshape(-1, 1) if x.ndim == 1 else x)
    return np.asfortranarray(np.vstack(mats))


def _mul_inplace(b, A, x):
    b[...] = jl_mul(A, x)
    return b


def _fill_inplace(a, v):
    a[...] = v
    return a


def _push(a, *v):
    if isinstance(a, list): a.extend(v)
    return a


def _rot180(A):
    if sp.issparse(A): return csc(A.tocsr()[::-1, :].tocsc()[:, ::-1])
    return np.array(plain(A)[::-1, ::-1])


def _minmax(fn, ufn):
    def f(*args):
        if len(args) == 1:
            return reduce_all(fn)(args[0])
        r = args[0]
        for x in args[1:]:
            if is_number(r) and is_number(x): r = fn((r, x))
            else: r = ufn(*align(r, x))
        return r
    f.elementwise = True
    return f


def _sign(x):
    return np.sign(x)


def _floor(*a):
    if len(a) == 2: return int(math.floor(a[1]))
    return np.floor(a[0])


def _regex(s):
    return re.compile(s)


def _occursin(pat, s):
    if hasattr(pat, "search"): return pat.search(s) is not None
    return pat in s


def _readlines(f):
    return [l.rstrip("\n").rstrip("\r") for l in f.readlines()]


def _ntuple(f, n):
    return tuple(f(i) for i in range(1, int(n) + 1))


def _range(*args, stop=None, length=None, step=None):
    if len(args) == 2 and stop is None: stop = args[1]
    if length is not None: return make_range(args[0], stop, length)
    return JRange(args[0], step if step is not None else 1, stop)


def _copy(x):
    if isinstance(x, np.ndarray): return np.array(x, order="F" if x.ndim > 1 else "C")
    if sp.issparse(x): return x.copy()
    if isinstance(x, list): return list(x)
    if isinstance(x, dict): return dict(x)
    return x


def _blockdiag(*ms):
    return csc(sp.block_diag([csc(m) for m in ms], format="csc"))


def _haskey(d, k):
    return _key(k) in d


def _norm(x, p=2):
    x = to_arr(x)
    if sp.issparse(x): return float(spla.norm(x))
    return float(np.linalg.norm(np.asarray(plain(x)).reshape(-1), p))


def _div(a, b):
    q = abs(int(a)) // abs(int(b))
    return q if (a >= 0) == (b >= 0) else -q


def _sum(a, *r):
    if callable(a) and r: return sum(a(v) for v in r[0])
    return reduce_all(np.sum)(a)


def _transpose(a):
    return jl_adj(a)


def _isempty(a):
    return jl_length(a) == 0


class _Rng:
    gen = np.random.default_rng(777)


def _rand(*d):
    if d and isinstance(d[0], JType): d = d[1:]
    return np.asfortranarray(_Rng.gen.uniform(0.0, 1.0, _dims(d))) if d else float(_Rng.gen.uniform())


def _eigen(A):
    A = np.asarray(jl_matrix(A), dtype=float)
    if np.allclose(A, A.T, rtol=0, atol=1e-14 * max(1.0, np.abs(A).max())):
        w, V = np.linalg.eigh((A + A.T) / 2)
    else:
        w, V = np.linalg.eig(A)
    return NT(["values", "vectors"], [w, V])


def _extrema(a):
    a = np.asarray(to_arr(a))
    return (a.min().item(), a.max().item())


def _enumerate(x):
    return [(i + 1, v) for i, v in enumerate(Interp.iterate(None, x))]


BUILTINS = {
    "size": jl_size, "length": jl_length, "zeros": jl_zeros, "ones": lambda *a: jl_zeros(*a, fill=1), "fill": jl_fill,
    "fill!": _fill_inplace, "sparse": jl_sparse, "spzeros": lambda m, n: csc((int(m), int(n))), "findnz": jl_findnz, "kron": jl_kron,
    "⊗": jl_kron, "diag": jl_diag, "Diagonal": jl_diagonal, "Matrix": jl_matrix, "Vector": jl_vector, "Array": jl_matrix,
    "collect": jl_collect, "reshape": jl_reshape, "minimum": _minmax(np.min, np.minimum), "maximum": _minmax(np.max, np.maximum),
    "min": _minmax(np.min, np.minimum), "max": _minmax(np.max, np.maximum), "sum": _sum,
    "abs": elementwise(np.abs), "sqrt": elementwise(np.sqrt), "exp": elementwise(np.exp), "log": elementwise(np.log),
    "sin": elementwise(np.sin), "cos": elementwise(np.cos), "tan": elementwise(np.tan), "asinh": elementwise(np.arcsinh),
    "sec": elementwise(lambda x: 1.0 / np.cos(x)), "sinh": elementwise(np.sinh), "cosh": elementwise(np.cosh), "tanh": elementwise(np.tanh), "log10": elementwise(np.log10),
    "atan": elementwise(lambda *a: np.arctan2(*a) if len(a) == 2 else np.arctan(*a)), "hypot": elementwise(np.hypot),
    "sign": elementwise(_sign), "isnan": elementwise(np.isnan), "isfinite": elementwise(np.isfinite), "abs2": elementwise(lambda x: x * x),
    "floor": _floor, "div": _div, "range": _range, "transpose": _transpose, "adjoint": jl_adj, "copy": _copy, "push!": _push,
    "haskey": _haskey, "error": jl_error, "string": jl_string, "println": lambda *a: None, "print": lambda *a: None,
    "isapprox": jl_isapprox, "norm": _norm, "rand": _rand, "eigen": _eigen, "eigvals": lambda A: _eigen(A).get("values"), "extrema": _extrema, "enumerate": _enumerate,
    "mod": lambda a, b: a % b, "real": elementwise(np.real), "imag": elementwise(np.imag), "exp10": elementwise(lambda x: 10.0 ** x),
    "ceil": elementwise(np.ceil), "Random": Stub(),
    "view": jl_view, "ntuple": _ntuple, "typeof": jl_typeof, "nnz": lambda A: int(csc(A).nnz), "rot180": _rot180,
    "cholesky": lambda A: Factor(A), "lu": lambda A: Factor(A), "Symmetric": lambda A: A, "mul!": _mul_inplace, "blockdiag": _blockdiag,
    "occursin": _occursin, "split": jl_split, "parse": jl_parse, "readlines": _readlines, "open": lambda fn: open(fn),
    "close": lambda f: f.close(), "Regex": _regex, "isempty": _isempty, "any": reduce_all(np.any), "all": reduce_all(np.all),
    "Int64": Int64, "Int": Int64, "Float64": Float64, "Bool": BoolT, "I": IDENT, "undef": UNDEF, "nothing": None,
    "π": math.pi, "pi": math.pi, "NaN": float("nan"), "Inf": float("inf"), "eps": lambda *a: np.finfo(float).eps,
    "cumsum": lambda a: np.cumsum(to_arr(a)), "sort": lambda a: np.sort(to_arr(a)), "unique": lambda a: np.unique(to_arr(a)),
    "issymmetric": lambda A: bool(abs(csc(A) - csc(A).T).max() == 0), "dot": lambda a, b: float(np.dot(to_arr(a), to_arr(b))),
    "Pair": lambda a, b: (a, b), "first": lambda a: jl_getindex(a, [1]), "last": lambda a: jl_getindex(a, [jl_length(a)]),
    "sparsevec": jl_sparse, "float": lambda x: to_arr(x) * 1.0, "round": lambda *a: np.round(a[-1]), "Tuple": TypeSpec("Tuple", ()),
}


# ======================================================================================================================
# evaluator
# ======================================================================================================================
class Env:
    __slots__ = ("vars", "parent", "is_func")

    def __init__(self, parent=None, is_func=False):
        self.vars, self.parent, self.is_func = {}, parent, is_func

    def lookup(self, name):
        e = self
        while e is not None:
            if name in e.vars: return e.vars[name]
            e = e.parent
        raise JuliaError("UndefVarError: %s" % name)

    def find(self, name):
        e = self
        while e is not None:
            if name in e.vars: return e
            e = e.parent
        return None


DOT_BIN = {
    "+": lambda a, b: a + b, "-": lambda a, b: a - b, "*": lambda a, b: a * b, "/": lambda a, b: np.true_divide(a, b),
    "^": lambda a, b: np.power(np.asarray(a, dtype=float) if isinstance(b, (float, np.floating)) or (isinstance(b, np.ndarray) and b.dtype.kind == "f") or (is_number(b) and b < 0) else a, b),
    "==": lambda a, b: a == b, "!=": lambda a, b: a != b, "<": lambda a, b: a < b, "<=": lambda a, b: a <= b,
    ">": lambda a, b: a > b, ">=": lambda a, b: a >= b,
}
TYPE_TESTS = {
    "Number": is_number, "Real": is_number, "Integer": lambda x: isinstance(x, (int, np.integer)),
    "AbstractArray": is_arraylike, "AbstractVector": is_arraylike, "AbstractMatrix": is_arraylike, "Array": lambda x: isinstance(x, np.ndarray),
    "String": lambda x: isinstance(x, str), "Function": callable, "Any": lambda x: True,
}


class Interp:
    def __init__(self, basedir):
        self.basedir = basedir
        self.globals = Env()
        self.globals.vars.update(BUILTINS)
        self.globals.vars["include"] = self.include
        self.globals.vars["open"] = self.open
        self.globals.vars["Iterators"] = NT(["flatten"], [jl_flatten])
        self.globals.vars["PGFPlots"] = Stub()
        self.structs = {}
        self.include_dirs = []
        self.end_stack = []
        self.views = 0
        self.log = []
        _Rng.gen = np.random.default_rng(777)

    # ---- files
    def include(self, fname):
        import os
        here = self.include_dirs[-1] if self.include_dirs else self.basedir      # relative to the including file
        path = os.path.normpath(fname if os.path.isabs(fname) else os.path.join(here, fname))
        with open(path) as f:
            ast = parse_source(f.read(), os.path.basename(path))
        self.include_dirs.append(os.path.dirname(path))
        try:
            return self.exec_block(ast, self.globals)
        finally:
            self.include_dirs.pop()

    def open(self, fname):
        import os
        d = self.basedir
        while not os.path.isabs(fname) and not os.path.exists(os.path.join(d, fname)) and d != "/":
            d = os.path.dirname(d)                        # the drivers are started from the reference's root or their own folder
        return open(fname if os.path.isabs(fname) else os.path.join(d, fname))

    def run(self, src, fname="<string>", env=None):
        return self.exec_block(parse_source(src, fname), env or self.globals)

    def call(self, name, *args, **kwargs):
        return self.apply(self.globals.lookup(name), list(args), kwargs)

    # ---- statements
    def exec_block(self, node, env):
        val = None
        for st in node[1]:
            val = self.ev(st, env)
        return val

    def assign(self, lhs, val, env):
        k = lhs[0]
        if k == "id":
            e, in_func = env, False
            while e is not None and lhs[1] not in e.vars:
                in_func = in_func or e.is_func
                e = e.parent
            if e is None or (e is self.globals and in_func): e = env       # a new local; globals are not assigned from functions
            e.vars[lhs[1]] = val
        elif k in ("tuple", "paren"):
            items = lhs[1] if k == "tuple" else [lhs[1]]
            vals = list(val.arr()) if isinstance(val, JRange) else (list(val) if not isinstance(val, np.ndarray) else list(flat_f(val)))
            if len(vals) < len(items): raise JuliaError("BoundsError: destructuring %d values into %d names" % (len(vals), len(items)))
            for it, v in zip(items, vals):
                if isinstance(v, np.generic): v = v.item()
                self.assign(it, v, env)
        elif k == "index":
            a = self.ev(lhs[1], env)
            idxs = self.eval_indices(a, lhs[2], env)
            jl_setindex(a, idxs, val)
        elif k == "field":
            raise JuliaError("assignment to a field is not supported")
        elif k == "typed":
            self.assign(lhs[1], val, env)
        elif k == "view":
            self.assign(lhs[1], val, env)
        else:
            raise JuliaError("cannot assign to %s" % k)

    def eval_indices(self, a, idx_nodes, env):
        out = []
        n = len(idx_nodes)
        for d, ix in enumerate(idx_nodes):
            if ix[0] == "colon":
                out.append(COLON); continue
            self.end_stack.append((a, d, n))
            try:
                out.append(self.ev(ix, env))
            finally:
                self.end_stack.pop()
        return out

    def define_function(self, node, env):
        _, name, pos, kws, body, typevars = node
        f = env.vars.get(name)
        if not isinstance(f, JFunction):
            f = JFunction(name, self); env.vars[name] = f
        sig = tuple((p[1], p[3]) for p in pos)
        f.methods = [m for m in f.methods if m["sig"] != sig or m["npos"] != len(pos)]
        f.methods.append({"pos": pos, "kws": kws, "body": body, "env": env, "tv": typevars, "sig": sig, "npos": len(pos)})
        return f

    def match(self, m, args):
        pos = m["pos"]
        nreq = sum(1 for p in pos if p[2] is None and not p[3])
        has_splat = any(p[3] for p in pos)
        if len(args) < nreq or (len(args) > len(pos) and not has_splat): return -1
        score = 0
        for p, a in zip(pos, args):
            t = p[1]
            if t is None: continue
            t = m["tv"].get(t, t) if t in m["tv"] else t
            if t is None: continue
            test = TYPE_TESTS.get(t)
            if test is None: continue
            if not test(a): return -1
            score += 1
        return score

    def apply(self, f, args, kwargs=None):
        kwargs = kwargs or {}
        if isinstance(f, JFunction):
            best, bs = None, -1
            for m in f.methods:
                s = self.match(m, args)
                if s >= bs and s >= 0: best, bs = m, s
            if best is None: raise JuliaError("MethodError: no method of %s for %d arguments" % (f.name, len(args)))
            return self.call_method(f, best, args, kwargs)
        if isinstance(f, Closure):
            env = Env(f.env, True)
            if len(args) != len(f.params): raise JuliaError("MethodError: lambda with %d parameters called with %d" % (len(f.params), len(args)))
            for p, a in zip(f.params, args): env.vars[p] = a
            try:
                return self.ev(f.body, env)
            except ReturnEx as r:
                return r.value
        if isinstance(f, TypeSpec):
            return self.construct(f, args)
        if callable(f):
            return f(*args, **kwargs)
        raise JuliaError("not callable: %r" % (f,))

    def call_method(self, f, m, args, kwargs):
        env = Env(m["env"], True)
        pos = m["pos"]
        for k, p in enumerate(pos):
            name, _, default, splat = p
            if splat: env.vars[name] = tuple(args[k:]); break
            if k < len(args): env.vars[name] = args[k]
            else: env.vars[name] = self.ev(default, env)
        rest = dict(kwargs)
        for name, default, splat in m["kws"]:
            if splat: env.vars[name] = rest; rest = {}; continue
            if name in rest: env.vars[name] = rest.pop(name)
            elif default is not None: env.vars[name] = self.ev(default, env)
            else: raise JuliaError("UndefKeywordError: %s" % name)
        if rest: raise JuliaError("MethodError: %s got unsupported keyword arguments %s" % (f.name, sorted(rest)))
        try:
            return self.exec_block(m["body"], env)
        except ReturnEx as r:
            return r.value

    def construct(self, ts, args):
        if ts.name == "Array":
            T = ts.params[0]
            dims = _dims([a for a in args if a is not UNDEF])
            if isinstance(T, JType) and T.dtype is not None: return np.zeros(dims, dtype=T.dtype, order="F")
            if len(dims) == 1: return [None] * dims[0]
            return np.empty(dims, dtype=object)
        if ts.name == "Dict": return {}
        if ts.name in self.structs:
            return NT(self.structs[ts.name], args)
        raise JuliaError("cannot construct %r" % ts)

    # ---- expressions
    def ev(self, n, env):
        k = n[0]
        return getattr(self, "ev_" + k)(n, env)

    def ev_ln(self, n, env):
        try:
            return self.ev(n[3], env)
        except (BreakEx, ContinueEx, ReturnEx):
            raise
        except Exception as ex:
            tr = getattr(ex, "jl_trace", None)
            if tr is None:
                tr = ex.jl_trace = []
                ex.args = (("%s  [at %s:%d]" % (ex.args[0] if ex.args else "", n[2], n[1])),) + tuple(ex.args[1:])
            tr.append((n[2], n[1]))
            raise

    def ev_nop(self, n, env): return None
    def ev_num(self, n, env): return n[1]
    def ev_regex(self, n, env): return re.compile(n[1])
    def ev_sym(self, n, env): return ":" + n[1]
    def ev_colon(self, n, env): return COLON
    def ev_paren(self, n, env): return self.ev(n[1], env)
    def ev_block(self, n, env): return self.exec_block(n, env)
    def ev_let(self, n, env):
        e = Env(env)
        for st in n[1][1]:                                   # bindings on the `let` line are new locals even if the name exists outside
            inner = st[3] if st[0] == "ln" else st
            if inner[0] == "assign" and inner[1][0] == "id" and st[0] != "ln": e.vars[inner[1][1]] = self.ev(inner[2], env)
        return self.exec_block(("block", [st for st in n[1][1] if st[0] == "ln" or st[0] != "assign"]), e)

    def ev_global(self, n, env):
        st = n[1]
        if st[0] == "assign" and st[1][0] == "id":
            val = self.ev(st[2], env); self.globals.vars[st[1][1]] = val; return val
        return self.ev(st, env)
    def ev_typed(self, n, env): return self.ev(n[1], env)
    def ev_break(self, n, env): raise BreakEx()
    def ev_continue(self, n, env): raise ContinueEx()
    def ev_show(self, n, env):
        try:
            self.log.append(self.ev(n[1], env))
        except JuliaError:
            pass
        return None
    def ev_pyhook(self, n, env): return n[1](env)

    def ev_str(self, n, env):
        s, out, i = n[1], [], 0
        while i < len(s):
            c = s[i]
            if c == "\\":
                nx = s[i + 1]
                out.append({"n": "\n", "t": "\t", "\\": "\\", '"': '"', "$": "$"}.get(nx, "\\" + nx)); i += 2
            elif c == "$" and i + 1 < len(s) and (s[i + 1] == "(" or _id_start(s[i + 1])):
                if s[i + 1] == "(":
                    depth, j = 1, i + 2
                    while depth:
                        depth += {"(": 1, ")": -1}.get(s[j], 0); j += 1
                    out.append(jl_string(self.ev(parse_source(s[i + 2:j - 1])[1][0][3], env))); i = j
                else:
                    j = i + 1
                    while j < len(s) and _id_cont(s[j]): j += 1
                    out.append(jl_string(env.lookup(s[i + 1:j]))); i = j
            else:
                out.append(c); i += 1
        return "".join(out)

    def ev_id(self, n, env):
        return env.lookup(n[1])

    def ev_endidx(self, n, env):
        a, d, nd = self.end_stack[-1]
        return jl_length(a) if nd == 1 else jl_size(a, d + 1)

    def ev_return(self, n, env):
        raise ReturnEx(None if n[1] is None else self.ev(n[1], env))

    def ev_assert(self, n, env):
        if not self.truth(self.ev(n[1], env)): raise JuliaError("AssertionError at line %d" % n[2])

    def ev_struct(self, n, env):
        self.structs[n[1]] = n[2]

    def ev_func(self, n, env):
        return self.define_function(n, env)

    def ev_lambda(self, n, env):
        return Closure(self, n[1], n[2], env)

    def ev_assign(self, n, env):
        val = self.ev(n[2], env)
        self.assign(n[1], val, env)
        return val

    def ev_dotassign(self, n, env):
        lhs = n[1]
        val = self.ev(n[2], env)
        if lhs[0] == "index":
            a = self.ev(lhs[1], env)
            if isinstance(a, (tuple, list, NT, dict)):                      # element of a container: broadcast into that array
                tgt = jl_getindex(a, self.eval_indices(a, lhs[2], env))
                _, v = align(tgt, to_arr(val))
                tgt[...] = v
            else:
                self.assign(lhs, val, env)
        else:
            a = self.ev(lhs, env)
            v = to_arr(val)
            if isinstance(v, np.ndarray):
                _, v = align(a, v)
                v = np.broadcast_to(v, a.shape) if v.shape != a.shape else v
            a[...] = v
        return val

    def ev_dotmacro(self, n, env):
        st = n[1]
        if st[0] == "assign": return self.ev_dotassign(("dotassign", st[1], self.dotted(st[2])), env)
        return self.ev(self.dotted(st), env)

    def dotted(self, n):
        if n[0] == "bin" and n[1] in DOT_BIN: return ("bin", "." + n[1], self.dotted(n[2]), self.dotted(n[3]))
        if n[0] == "paren": return ("paren", self.dotted(n[1]))
        if n[0] == "un": return ("un", n[1], self.dotted(n[2]))
        if n[0] == "call": return ("call", n[1], [self.dotted(a) for a in n[2]], n[3], True)
        return n

    def ev_viewsblock(self, n, env):
        self.views += 1
        try:
            return self.ev(n[1], env)
        finally:
            self.views -= 1

    def ev_view(self, n, env):
        self.views += 1
        try:
            return self.ev(n[1], env)
        finally:
            self.views -= 1

    def ev_if(self, n, env):
        for cond, body in n[1]:
            if self.truth(self.ev(cond, env)): return self.exec_block(body, env)
        if n[2] is not None: return self.exec_block(n[2], env)
        return None

    def truth(self, v):
        if isinstance(v, (bool, np.bool_)): return bool(v)
        raise JuliaError("TypeError: non-boolean (%s) used in boolean context" % type(v).__name__)

    def ev_ternary(self, n, env):
        return self.ev(n[2], env) if self.truth(self.ev(n[1], env)) else self.ev(n[3], env)

    def ev_and(self, n, env):
        a = self.ev(n[1], env)
        return self.ev(n[2], env) if self.truth(a) else False

    def ev_or(self, n, env):
        a = self.ev(n[1], env)
        return True if self.truth(a) else self.ev(n[2], env)

    def ev_while(self, n, env):
        while self.truth(self.ev(n[1], env)):
            try:
                self.exec_block(n[2], Env(env))
            except BreakEx:
                break
            except ContinueEx:
                continue
        return None

    def ev_for(self, n, env):
        self.run_for(n[1], n[2], env)
        return None

    def iterate(self, it):
        if isinstance(it, NT): return list(it._values)
        if isinstance(it, JRange): return it.arr().tolist()
        if isinstance(it, np.ndarray): return [v.item() if isinstance(v, np.generic) else v for v in flat_f(it)]
        if isinstance(it, dict): return list(it.items())
        if is_number(it): return [it]                      # numbers iterate over themselves
        return list(it)

    def run_for(self, iters, body, env):
        var, itn = iters[0]
        for v in self.iterate(self.ev(itn, env)):
            e = Env(env)
            self.assign_local(var, v, e)
            try:
                if len(iters) > 1: self.run_for(iters[1:], body, e)
                else: self.exec_block(body, e)
            except BreakEx:
                break
            except ContinueEx:
                continue

    def assign_local(self, var, v, e):
        if var[0] == "id": e.vars[var[1]] = v
        else:
            for it, x in zip(var[1], list(v)): self.assign_local(it, x, e)

    def ev_try(self, n, env):
        try:
            return self.exec_block(n[1], env)
        except (JuliaError, ValueError, IndexError, KeyError, OSError) as ex:
            e = Env(env)
            if n[2]: e.vars[n[2]] = ex
            return self.exec_block(n[3], e)

    def ev_tuple(self, n, env):
        out = []
        for it in n[1]:
            if it[0] == "splat": out.extend(self.iterate(self.ev(it[1], env)))
            else: out.append(self.ev(it, env))
        return tuple(out)

    def ev_ntuple(self, n, env):
        return NT([nm for nm, _ in n[1]], [self.ev(e, env) for _, e in n[1]])

    def ev_vect(self, n, env):
        vals = [self.ev(it, env) for it in n[1]]
        if vals and all(is_number(v) for v in vals):
            if all(isinstance(v, (bool, np.bool_)) for v in vals): return np.array(vals, dtype=np.bool_)
            if all(isinstance(v, (int, np.integer)) and not isinstance(v, (bool, np.bool_)) for v in vals): return np.array(vals, dtype=np.int64)
            return np.array(vals, dtype=np.float64)
        if not vals: return []                                   # Any[]: grows by push!
        if len(vals) == 1 and isinstance(vals[0], JRange): return [vals[0]]
        return list(vals)

    def ev_matrix(self, n, env):
        return jl_hvcat([[self.ev(it, env) for it in row] for row in n[1]])

    def ev_comp(self, n, env):
        out = []
        for v in self.iterate(self.ev(n[3], env)):
            e = Env(env); self.assign_local(n[2], v, e)
            out.append(self.ev(n[1], e))
        return np.array(out) if out and all(is_number(v) for v in out) else out

    def ev_range(self, n, env):
        a = self.ev(n[1], env); c = self.ev(n[3], env)
        b = 1 if n[2] is None else self.ev(n[2], env)
        return JRange(a, b, c)

    def ev_adj(self, n, env):
        return jl_adj(self.ev(n[1], env))

    def ev_un(self, n, env):
        v = self.ev(n[2], env)
        if n[1] == "!": return not self.truth(v)
        v = to_arr(v)
        if n[1] == "+": return v
        r = -v
        return csc(r) if sp.issparse(r) else r

    def ev_field(self, n, env):
        a = self.ev(n[1], env)
        name = n[2]
        if isinstance(a, NT): return a.get(name)
        if isinstance(a, Stub): return Stub()
        if sp.issparse(a) and name == "nzval": return csc(a).data
        if isinstance(a, Factor) and name in ("L", "U"): raise JuliaError("factor parts are not supported")
        raise JuliaError("type %s has no field %s" % (type(a).__name__, name))

    def ev_curly(self, n, env):
        base = n[1][1] if n[1][0] == "id" else n[1][2]
        params = []
        for p in n[2]:
            try:
                params.append(self.ev(p, env))
            except JuliaError:
                params.append(p[1] if p[0] == "id" else None)
        return TypeSpec(base, params)

    def ev_index(self, n, env):
        a = self.ev(n[1], env)
        idxs = self.eval_indices(a, n[2], env)
        return jl_getindex(a, idxs, view=self.views > 0)

    def ev_cmp(self, n, env):
        vals = [self.ev(n[1][0], env)]
        res = True
        for op, nd in zip(n[2], n[1][1:]):
            if res is False: break
            vals.append(self.ev(nd, env))
            r = self.compare(op, vals[-2], vals[-1])
            if len(n[2]) == 1: return r
            res = res and self.truth(r)
        return res

    def compare(self, op, a, b):
        if op == "==": return jl_equal(a, b)
        if op in ("!=", "≠"): return not jl_equal(a, b)
        if op == "≈": return jl_isapprox(a, b)
        if op == "∈":
            if isinstance(b, JRange): return any(a == v for v in b)
            if is_number(b): return a == b
            return any(jl_equal(a, v) for v in (flat_f(b) if isinstance(b, np.ndarray) else b))
        if op.startswith("."):
            x, y = align(a, b)
            return DOT_BIN[op[1:]](x, y)
        if not (is_number(a) and is_number(b)): raise JuliaError("comparison %s of non-scalars" % op)
        return {"<": a < b, "<=": a <= b, "≤": a <= b, ">": a > b, ">=": a >= b, "≥": a >= b}[op]

    def ev_bin(self, n, env):
        op = n[1]
        a = self.ev(n[2], env); b = self.ev(n[3], env)
        return self.binop(op, a, b)

    def binop(self, op, a, b):
        if op == "+": return jl_add(a, b, 1)
        if op == "-": return jl_add(a, b, -1)
        if op == "*": return jl_mul(a, b)
        if op == "/": return jl_div(a, b)
        if op == "\\": return jl_ldiv(a, b)
        if op == "^": return jl_pow(a, b)
        if op == "⊗": return self.apply(self.globals.lookup("⊗"), [a, b])
        if op == "÷": return _div(a, b)
        if op == "%": return a % b
        if op.startswith("."):
            x, y = align(a, b)
            if is_number(x) and is_number(y):
                return self.binop(op[1:], x, y)
            r = DOT_BIN[op[1:]](x, y)
            return r
        raise JuliaError("operator %s is not supported" % op)

    def ev_call(self, n, env):
        _, fn, argn, kwn, dotted = n
        if fn[0] == "id" and fn[1] in ("plot_blocks", "plot_connectivity"): return None
        f = self.ev(fn, env)
        args = []
        for a in argn:
            if a[0] == "splat": args.extend(self.iterate(self.ev(a[1], env)))
            else: args.append(self.ev(a, env))
        kwargs = {}
        for kname, kv, _ in kwn:
            if kname == "...": kwargs.update(self.ev(kv, env))
            else: kwargs[kname] = self.ev(kv, env)
        if dotted:
            return self.broadcast_call(f, args)
        return self.apply(f, args, kwargs)

    def broadcast_call(self, f, args):
        if getattr(f, "elementwise", False) or isinstance(f, JType):
            if isinstance(f, JType):
                a = to_arr(args[0]); return a.astype(f.dtype) if isinstance(a, np.ndarray) else f(a)
            return f(*args)
        arrs = [to_arr(a) for a in args]
        if all(is_number(a) or not isinstance(a, np.ndarray) for a in arrs): return self.apply(f, args)
        shape = np.broadcast(*[a for a in arrs if isinstance(a, np.ndarray)]).shape
        out = np.empty(shape, dtype=object)
        its = [np.broadcast_to(a, shape) if isinstance(a, np.ndarray) else None for a in arrs]
        for ix in np.ndindex(*shape):
            out[ix] = self.apply(f, [it[ix].item() if it is not None else a for it, a in zip(its, arrs)])
        try:
            return out.astype(np.float64)
        except (TypeError, ValueError):
            return out
