"""The reference executed, not restated: tests/refexec/minijulia.py interprets the reference's own Julia statements (Julia itself is
not installed in this image), and the oracle -- the restatement every GPU parity test is checked against -- has to reproduce what
those statements compute.  Covers the 1-D SBP tables and variable-coefficient operators (diagonal_sbp.jl), create_metrics /
locoperator with curved metrics and every boundary-condition type, the complete square_circle.jl driver (mesh reader, connectivity,
trace operators, assembleλmatrix, LocalToGLobalRHS!, boundary / jump / source data, trace solve, back substitution, error norms)
and BP1's setup + odefun.  Also checks that the committed golden vectors are what tools/gen_refexec_golden.py produces.

Runs where /root/reference is mounted (the CPU tier); the golden vectors carry the result to the GPU box."""
import os
import sys

import numpy as np
import pytest
import scipy.sparse as sp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "tools"))
REF = "/root/reference"
pytestmark = pytest.mark.skipif(not os.path.isdir(REF), reason="the reference tree is not mounted on this machine")

from refexec.minijulia import Interp, JuliaError          # noqa: E402
from oracle import sbp as osbp, hybrid as orc             # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden", "refexec")


def dense(a):
    return a.toarray() if sp.issparse(a) else np.asarray(a, dtype=float)


def relmax(a, b):
    a, b = dense(a), dense(b)
    assert a.shape == b.shape
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)


@pytest.fixture(scope="module")
def ref():
    it = Interp(REF)
    it.include("global_curved.jl")
    return it


# ---- the interpreter itself: known answers for the Julia semantics the reference relies on --------------------------------------
def test_interpreter_semantics():
    it = Interp(REF)
    out = it.run("""
    A = [1 2 3; 4 5 6]                      # rows by ';', columns by blanks
    v = A[:]                                # column-major
    B = reshape(1:6, 2, 3)
    w = [1 -2 3]                            # 'a -b' inside brackets: two elements
    d = [1 - 2 3]
    r = 2:2:7
    x = zeros(4); xv = @view x[2:3]; xv .= 5; xv[1] = 7
    y = zeros(4); yc = y[2:3]; yc[1] = 7    # a copy: y untouched
    col = [10, 20]
    bc = col .* A                           # broadcasting: missing dimensions are trailing
    k = kron([1 2], [1, 1])
    s = sparse([1, 1, 2], [1, 1, 2], [1.0, 2.0, 3.0], 2, 2)    # duplicates are summed
    f(a; scale = 2) = a * scale
    f(a::AbstractArray; scale = 2) = a .* (10scale)
    g = (a, b) -> a + 2b
    t = (first = 1, second = [3.0, 4.0])
    h = 0
    for i = 1:3, j = 1:2
      h += i * j
    end
    q = 7 ÷ 2 + div(9, 2) + 2^3
    (v, B, w, d, collect(r), x, y, bc, k, Matrix(s), f(3), f([1, 2]; scale = 1), g(1, 2), t.second[end], h, q, 4col[2]^2, A', 1 / 2,
     [x[1:2]; 9], A[end, end], A[2, :], (-1 < 0 < 1), !(1 == 2) && true, length(A), size(A, 2))
    """)
    exp = (np.array([1, 4, 2, 5, 3, 6]), np.array([[1, 3, 5], [2, 4, 6]]), np.array([[1, -2, 3]]), np.array([[-1, 3]]), np.array([2, 4, 6]),
           np.array([0, 7.0, 5, 0]), np.zeros(4), np.array([[10, 20, 30], [80, 100, 120]]), np.array([[1, 2], [1, 2]]),
           np.array([[3.0, 0], [0, 3.0]]), 6, np.array([10, 20]), 5, 4.0, 18, 15, 1600, np.array([[1, 4], [2, 5], [3, 6]]), 0.5,
           np.array([0, 7.0, 9]), 6, np.array([4, 5, 6]), True, True, 6, 3)
    assert len(out) == len(exp)
    for k, (a, b) in enumerate(zip(out, exp)):
        assert np.array_equal(np.asarray(a), np.asarray(b)), (k, a, b)
    with pytest.raises(JuliaError):
        it.run("zeros(3) + zeros(4)")
    with pytest.raises(JuliaError):
        it.run("u = zeros(3); u[4]")


def test_interpreter_semantics_scoping_dispatch_and_linear_algebra():
    """more known answers: mesh-grid krons, findnz order, operator precedence, empty / reversed ranges, `end` in indices, linear
    indexing, interpolation, do-blocks, named tuples, short circuits, closures (one binding per loop iteration, captured locals),
    let / global, default arguments, dispatch on Number / AbstractArray, aliasing vs rebinding vs in-place broadcast, comparisons,
    integer and float division, adjoints and solves"""
    it = Interp(REF)
    out = it.run(r'''
    r = range(-1, stop=1, length=5)
    R = ones(1, 3) ⊗ r                 # R[i, j] = r[i]
    S = r[1:3]' ⊗ ones(5)              # S[i, j] = r[j]
    A = sparse([2, 1, 2], [1, 2, 2], [5.0, 6.0, 7.0], 2, 2)
    (I1, J1, V1) = findnz(A)            # column-major order
    (I2, J2, V2) = findnz(sparse(transpose(A)))
    v = [1.0, 2.0, 3.0]
    quad = v' * sparse([1, 2, 3], [1, 2, 3], [2.0, 2.0, 2.0]) * v
    pw = (-2^2, 2.0^-1, -v[2]^2, 2^3^2)
    emp = (length(1:0), length(3:-1:1), collect(5:-2:1))
    idx = 2 .+ (1:3)
    m = reshape(collect(1:12), 3, 4)
    ends = (m[end, 1], m[1, end], m[end], m[end-1, end-1], v[end:-1:1])
    lin = m[[2, 5, 12]]
    sub = m[2:3, [1, 4]]
    name = "blk"
    str = "x$(1 + 2)_$name"
    tup = ntuple(4) do k
      k^2
    end
    nt = (a = 1, b = "two", c = [3.0])
    (a1, b1) = (nt.a, nt.c[1])
    tern = 3 > 2 ? (1:3)[2] : -1
    hits = 0
    sc = (false && (hits += 1; true), true || (hits += 1; true), hits)
    acc = []
    for (i, w) in enumerate([10, 20])
      push!(acc, i * w)
    end
    fs = [() -> k for k = 1:3]
    caps = [f() for f in fs]
    function counter()
      n = 0
      bump() = (n += 1)
      bump(); bump()
      n
    end
    let q = 5
      global from_let = q + 1
    end
    g(a, b = 2a) = a + b
    kind(x::Number) = "number"
    kind(x::AbstractArray) = "array"
    kind(x) = "other"
    kinds = (kind(1.5), kind([1]), kind(1:3), kind(sin), kind("s"))
    w = [1.0, 2.0, 3.0]; alias = w; alias[1] = 9.0
    w2 = w; w2 += [1.0, 1.0, 1.0]
    z = zeros(3); z .= w .* 2
    eqs = ([1, 2] == [1, 2], [1, 2] == [2, 1], [1, 2][end:-1:1] == [2, 1], [1.0, 2.0] ≈ [1.0, 2.0 + 1e-12], 7 ∈ (1, 7), 3 ∈ 1:2)
    dv = (7 / 2, div(7, 2), 7 ÷ 2, 7 % 3, 2 * 3 / 4)
    un = -[1, 2] .+ 1
    cmpc = 1 < 2 <= 2 != 3
    A2 = [1 2; 3 4]
    tr = (A2', A2 * [1, 1], [1, 1]' * A2, A2 .* [10, 20], A2 * A2, A2 \ [5.0, 11.0])
    (R[:, 2], S[4, :], I1, J1, V1, I2, J2, V2, quad, pw, emp, collect(idx), ends, lin, sub, str, tup, (a1, b1), tern, sc, acc, caps, counter(),
     from_let, g(1), g(1, 1), kinds, w, w2, z, eqs, dv, un, cmpc, tr)
    ''')
    r5 = np.array([-1, -0.5, 0, 0.5, 1])
    exp = (r5, r5[:3], [2, 1, 2], [1, 2, 2], [5.0, 6, 7], [2, 1, 2], [1, 2, 2], [6.0, 5, 7], 28.0, (-4, 0.5, -4.0, 512), (0, 3, [5, 3, 1]), [3, 4, 5],
           (3, 10, 12, 8, [3.0, 2, 1]), [2, 5, 12], [[2, 11], [3, 12]], "x3_blk", (1, 4, 9, 16), (1, 3.0), 2, (False, True, 0), [10, 40], [1, 2, 3], 2,
           6, 3, 2, ("number", "array", "array", "other", "other"), [9.0, 2, 3], [10.0, 3, 4], [18.0, 4, 6], (True, False, True, True, True, False),
           (3.5, 3, 3, 1, 1.5), [0, -1], True, ([[1, 3], [2, 4]], [3, 7], [[4, 6]], [[10, 20], [60, 80]], [[7, 10], [15, 22]], [1.0, 2.0]))

    def same(a, b):
        if isinstance(b, tuple):
            return isinstance(a, tuple) and len(a) == len(b) and all(same(x, y) for x, y in zip(a, b))
        if isinstance(b, str): return a == b
        return np.array_equal(np.asarray(a), np.asarray(b))
    assert len(out) == len(exp)
    for k, (a, b) in enumerate(zip(out, exp)):
        assert same(a, b), (k, a, b)


# ---- diagonal_sbp.jl ---------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("p", [2, 4, 6])
def test_sbp_operators_executed(ref, p):
    N = 23
    D, HI, H, r = ref.call("diagonal_sbp_D1", p, N)
    oD, oHI, oH, orr = osbp.diagonal_sbp_D1(p, N)
    assert relmax(D, oD) < 1e-15 and relmax(H, oH) < 1e-15 and relmax(HI, oHI) < 1e-15 and relmax(r, orr) < 1e-15
    B = np.random.default_rng(p).uniform(0.5, 2.0, N + 1)
    out = ref.call("variable_diagonal_sbp_D2", p, N, B)                   # (D, S0, SN, HI, H, M, r)   diagonal_sbp.jl:763
    oo = osbp.variable_diagonal_sbp_D2(p, N, B)
    for a, b in zip(out, oo):
        assert relmax(a, b) < 2e-15
    out = ref.call("variable_diagonal_sbp_D2", p, N, 1.5)                 # the method for a constant coefficient, :478
    oo = osbp.variable_diagonal_sbp_D2(p, N, 1.5 * np.ones(N + 1))
    assert relmax(out[5], oo[5]) < 2e-15


@pytest.mark.parametrize("p", [2, 4, 6])
def test_constant_and_variable_second_derivative_agree_executed(ref, p):
    """the reference's own 'affine mesh test' (global_curved.jl:274-277, commented out there): with b = 1 the variable-coefficient
    stiffness is SN - S0 - H D2 of diagonal_sbp_D2 (diagonal_sbp.jl:203-465), two independently typed sets of tables"""
    N = 25
    D2, S0, SN, HI, H, r = ref.call("diagonal_sbp_D2", p, N)
    Dv, S0v, SNv, HIv, Hv, Mv, rv = ref.call("variable_diagonal_sbp_D2", p, N, np.ones(N + 1))
    assert relmax(Mv, SN - S0 - H @ D2) < 1e-14 and relmax(Dv, D2) < 1e-14 and relmax(S0v, S0) == 0 and relmax(SNv, SN) == 0
    assert relmax(osbp.variable_diagonal_sbp_D2(p, N, np.ones(N + 1))[5], SN - S0 - H @ D2) < 1e-14


# ---- create_metrics / locoperator --------------------------------------------------------------------------------------------------
def curved_maps():
    import gen_refexec_golden as gg
    xf = lambda r, s: (r + 0.1 * np.sin(2 * r) * np.cos(s) + 0.2 * s, 1 + 0.2 * np.cos(2 * r) * np.cos(s), -0.1 * np.sin(2 * r) * np.sin(s) + 0.2)
    yf = lambda r, s: (s + 0.15 * np.sin(r + s), 0.15 * np.cos(r + s), 1 + 0.15 * np.cos(r + s))
    return gg.CURVED_MAP, xf, yf


@pytest.mark.parametrize("p", [2, 4, 6])
def test_locoperator_executed(ref, p):
    src, xf, yf = curved_maps()
    ref.run(src)
    N = 19
    m = ref.call("create_metrics", p, N, N, ref.globals.lookup("xfun"), ref.globals.lookup("yfun"))
    om = orc.create_metrics(p, N, N, xf, yf)
    for k in ("crr", "css", "crs", "J"):
        assert relmax(m.get(k), getattr(om, k)) < 5e-15
    for k in ("sJ", "nx", "ny"):
        for a, b in zip(m.get(k), getattr(om, k)):
            assert relmax(a, b) < 5e-15
    for bc in ((1, 1, 1, 1), (0, 2, 7, 1), (2, 2, 0, 0), (7, 0, 2, 8)):
        lop = ref.call("locoperator", p, N, N, m, bc)
        ol = orc.locoperator(p, N, N, om, bc)
        assert relmax(lop.get("M̃"), ol.Mt) < 5e-15
        assert relmax(lop.get("JH"), ol.JH) < 5e-15
        for name, mine in (("F", ol.F), ("HfI_FT", ol.HfI_FT), ("HfI_G", ol.HfI_G), ("τ", ol.tau), ("Hf", ol.Hf), ("HfI", ol.HfI)):
            for a, b in zip(lop.get(name), mine):
                assert relmax(a, b) < 5e-15, name
        assert tuple(lop.get("bctype")) == tuple(ol.bctype)
    lop = ref.call("locoperator", p, N, N, m, (1, 2, 0, 7), **{"τscale": 1})   # keyword of :214 (local_op_eigenvalues.jl uses 1)
    ol = orc.locoperator(p, N, N, om, (1, 2, 0, 7), tauscale=1.0)
    assert relmax(lop.get("M̃"), ol.Mt) < 5e-15
    for a, b in zip(lop.get("τ"), ol.tau):
        assert relmax(a, b) < 5e-15
    with pytest.raises(JuliaError):
        ref.call("locoperator", p, N, N, m, (1, 3, 1, 1))                 # 'invalid bc', global_curved.jl:484
    # the reference as written needs Nr == Ns: an unused remainder (global_curved.jl:316) multiplies r-direction matrices with
    # s-direction coefficients.  The oracle and the CUDA path have no such restriction (they are compared on Nr != Ns elsewhere).
    with pytest.raises(Exception):
        ref.call("locoperator", p, N, N + 4, ref.call("create_metrics", p, N, N + 4))


@pytest.mark.parametrize("p", [2, 4, 6])
def test_transfinite_blend_three_methods_executed(ref, p):
    """global_curved.jl:19-78: analytic edge derivatives, SBP-differentiated edges (:53-64), corner values (:66-78) -- the reference's
    methods against the oracle's and the product's host-side mirror (whose D1 table is generated, hybridsbp_b200/_sbp_d1.py)"""
    from hybridsbp_b200 import host
    ref.run("""
    b1(t) = -1 .+ 0.1 .* (1 .- t .^ 2)
    b2(t) =  1 .+ 0.05 .* (1 .- t .^ 2)
    b3(t) = t .+ 0.0 .* t
    b4(t) = t .+ 0.0 .* t
    """)
    b1, b2 = (lambda t: -1 + 0.1 * (1 - t ** 2)), (lambda t: 1 + 0.05 * (1 - t ** 2))
    b3 = b4 = lambda t: t + 0.0 * t
    N = 17
    r = np.asfortranarray(np.repeat(np.linspace(-1, 1, N + 1)[:, None], N + 1, axis=1)); s = np.asfortranarray(r.T)
    assert np.max(np.abs(host.d1_matrix(p, N) - dense(ref.call("diagonal_sbp_D1", p, N)[0]))) < 1e-14
    out = ref.call("transfinite_blend", *[ref.globals.lookup(k) for k in ("b1", "b2", "b3", "b4")], r, s, p)
    for a, b, c in zip(out, orc.transfinite_blend_sbp(b1, b2, b3, b4, r, s, p), host.transfinite_blend(b1, b2, b3, b4, r, s, p)):
        assert np.max(np.abs(np.asarray(a) - b)) < 1e-14 and np.max(np.abs(np.asarray(a) - c)) < 1e-14
    out = ref.call("transfinite_blend", 0.1, 1.3, -0.2, 1.1, r, s)
    for a, b, c in zip(out, orc.transfinite_blend_corners(0.1, 1.3, -0.2, 1.1, r, s), host.transfinite_blend(0.1, 1.3, -0.2, 1.1, r, s)):
        assert np.max(np.abs(np.asarray(a) - b)) < 1e-14 and np.max(np.abs(np.asarray(a) - c)) < 1e-14
    with pytest.raises(JuliaError):                                         # edges that do not meet at the corners, :25
        ref.call("transfinite_blend", ref.globals.lookup("b2"), ref.globals.lookup("b2"), ref.globals.lookup("b3"), ref.globals.lookup("b4"), r, s, p)
    with pytest.raises(AssertionError):
        host.transfinite_blend(b2, b2, b3, b4, r, s, p)


def test_penalty_asserts_positive_psi(ref):
    m = ref.call("create_metrics", 2, 12, 12)
    bad = np.array(m.get("crr")); bad[5, 5] = -1.0
    with pytest.raises(JuliaError):
        ref.call("locoperator", 2, 12, 12, m, (1, 1, 1, 1), crr=bad)     # @assert minimum(ψmin) > 0, :419


# ---- mesh reader and connectivity on every mesh the reference ships ------------------------------------------------------------------
@pytest.mark.parametrize("mesh", ["meshes/square_circle.inp", "meshes/flower_v2.inp", "seas/BP1/meshes/1_1_block.inp", "seas/BP1/meshes/BP1_v1.inp"])
def test_read_inp_2d_and_connectivity_executed(ref, mesh):
    """read_inp_2d (global_curved.jl:802-946) and connectivityarrays (:82-132) as the reference runs them, against the product's
    host-side reader (hybridsbp_b200/host.py) and the oracle's, on the repository's copies of the same files"""
    from hybridsbp_b200 import host
    path = os.path.join(REF, mesh)
    mine = os.path.join(ROOT, "meshes", os.path.basename(mesh))
    assert open(path).read() == open(mine).read()
    # square_circle.inp / 1_1_block.inp carry side sets 4 and 5, which the reader's assertion (:937) only accepts after the drivers'
    # bc_map (square_circle.jl:11-12, BP1.jl:36-37); the other two meshes are read with the default map
    bc_map = np.array([1, 1, 2, 2, 7]) if ("square" in mesh or "1_1" in mesh) else None
    if bc_map is None:
        out, hv, ov = ref.call("read_inp_2d", path), host.read_inp_2d(mine), orc.read_inp_2d(mine)
    else:
        out, hv, ov = ref.call("read_inp_2d", path, bc_map=bc_map), host.read_inp_2d(mine, list(bc_map)), orc.read_inp_2d(mine, list(bc_map))
        if "square" in mesh:
            with pytest.raises(JuliaError):
                ref.call("read_inp_2d", path)
    for a, b, c in zip(out, hv, ov):
        assert np.array_equal(np.asarray(a, dtype=float), np.asarray(b, dtype=float))
        assert np.array_equal(np.asarray(a, dtype=float), np.asarray(c, dtype=float))
    conn = ref.call("connectivityarrays", out[1], out[2])
    for a, b, c in zip(conn, host.connectivityarrays(hv[1], hv[2]), orc.connectivityarrays(ov[1], ov[2])):
        assert np.array_equal(np.asarray(a).astype(int), np.asarray(b).astype(int))
        assert np.array_equal(np.asarray(a).astype(int), np.asarray(c).astype(int))
    Nr = np.full(out[1].shape[1], 9); Ns = np.full(out[1].shape[1], 11)
    for codes in (7, (7, 8), 1):
        a = ref.call("bcstarts", out[3], conn[0], conn[1], codes, Nr, Ns)  # :714-728
        b = host.bcstarts(hv[3], conn[0], conn[1], codes if isinstance(codes, tuple) else (codes,), list(Nr), list(Ns))
        assert np.array_equal(np.asarray(a), np.asarray(b))


# ---- the driver of configuration 1 -------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def square_circle_p4():
    from refexec.drivers import run_square_circle
    from refexec.oracle_driver import oracle_square_circle_level
    cap, mesh, _ = run_square_circle(p=4, levels=1, N0=17)
    return cap[0], mesh, oracle_square_circle_level(4, 17)


def test_square_circle_mesh_and_connectivity_executed(square_circle_p4):
    c, mesh, o = square_circle_p4
    verts, EToV, EToF, FToB, dom = o["mesh"]                               # the product's reader + the circle fix-up of :27-33
    for k, a in (("verts", verts), ("EToV", EToV), ("EToF", EToF), ("FToB", FToB), ("EToDomain", dom)):
        assert np.array_equal(np.asarray(mesh[k], dtype=float), np.asarray(a, dtype=float)), k
    for k, a in zip(("FToE", "FToLF", "EToO", "EToS"), o["conn"]):
        assert np.array_equal(np.asarray(mesh[k]).astype(int), np.asarray(a).astype(int)), k
    assert np.array_equal(c["vstarts"], o["vstarts"]) and np.array_equal(c["FToλstarts"], o["FTol"]) and np.array_equal(c["FToδstarts"], o["FTod"])


def test_square_circle_operators_executed(square_circle_p4):
    c, mesh, o = square_circle_p4
    assert relmax(c["FbarT"], o["FbarT"]) < 1e-14
    assert relmax(c["D"], o["D"]) < 1e-14
    assert relmax(c["B"], o["B"]) < 1e-13                                  # assembleλmatrix: 56 x 4 x 18 local solves
    for e in (1, 9, 30, 56):                                               # straight, curved (circle) and boundary blocks
        assert relmax(c["lop"][e]["M̃"], o["lops"][e - 1].Mt) < 1e-14
        for lf in range(4):
            assert relmax(c["lop"][e]["F"][lf], o["lops"][e - 1].F[lf]) < 1e-14


def test_square_circle_solution_executed(square_circle_p4):
    c, mesh, o = square_circle_p4
    rel = lambda a, b: np.linalg.norm(np.asarray(a) - np.asarray(b)) / np.linalg.norm(np.asarray(b))
    assert rel(o["delta"], c["δ"]) < 1e-14 and rel(o["g"], c["g"]) < 1e-13 and rel(o["gd"], c["gδ"]) < 1e-13
    assert rel(o["bl"], c["bλ"]) < 1e-12
    assert rel(o["lam"], c["λ"]) < 1e-11 and rel(o["u"], c["u"]) < 1e-11
    assert abs(o["eps"] - c["ϵ"][0]) < 1e-6 * c["ϵ"][0] and abs(o["teps"] - c["τϵ"][0]) < 1e-6 * c["τϵ"][0]


@pytest.mark.skipif(not os.environ.get("HSBP_SLOW_TESTS"), reason="20 s more of the same driver; set HSBP_SLOW_TESTS=1")
def test_square_circle_order_2_executed():
    """the same driver at SBP order 2 (BP1's order): lambda, u and the error norms of the oracle against the executed reference"""
    from refexec.drivers import run_square_circle
    from refexec.oracle_driver import oracle_square_circle_level
    cap, _, _ = run_square_circle(p=2, levels=1, N0=17, keep=("λ", "u", "gδ", "ϵ", "τϵ"))
    c, o = cap[0], oracle_square_circle_level(2, 17)
    rel = lambda a, b: np.linalg.norm(np.asarray(a) - np.asarray(b)) / np.linalg.norm(np.asarray(b))
    assert rel(o["gd"], c["gδ"]) < 1e-13 and rel(o["lam"], c["λ"]) < 1e-11 and rel(o["u"], c["u"]) < 1e-11
    assert abs(o["eps"] - c["ϵ"][0]) < 1e-7 * c["ϵ"][0] and abs(o["teps"] - c["τϵ"][0]) < 1e-7 * c["τϵ"][0]


# ---- configuration 2: the reference's functions on the flower mesh (reversed faces) --------------------------------------------------
def test_flower_mesh_reversed_faces_executed():
    """gloλoperator's flipped branch (`FToλstarts[f+1] .- Ie`, `rot180(τ)`, global_curved.jl:546-549), in_jump's three branches and
    the reversed g-delta views, with slip data on the jump faces: tests/refexec/trace_driver.jl drives the reference's functions"""
    from refexec.drivers import run_flower
    from refexec.oracle_driver import oracle_square_circle_level
    from hybridsbp_b200 import flower
    c = run_flower(4, 17)
    o = oracle_square_circle_level(4, 17, mesh=flower.load_mesh(), maps=flower.block_maps, exact=flower.Smooth,
                                   slip=lambda x, y: 0.3 * np.sin(x) * np.cos(2 * y))
    assert int((~np.asarray(c["EToO"]).astype(bool)).sum()) == 27 and int((np.asarray(c["FToB"]) == 7).sum()) == 18
    for k, a in zip(("verts", "EToV", "EToF", "FToB"), o["mesh"]):
        assert np.array_equal(np.asarray(c[k], dtype=float), np.asarray(a, dtype=float)), k
    for k, a in zip(("FToE", "FToLF", "EToO", "EToS"), o["conn"]):
        assert np.array_equal(np.asarray(c[k]).astype(int), np.asarray(a).astype(int)), k
    rel = lambda a, b: np.linalg.norm(np.asarray(a) - np.asarray(b)) / np.linalg.norm(np.asarray(b))
    assert relmax(c["FbarT"], o["FbarT"]) < 1e-14 and relmax(c["D"], o["D"]) < 1e-14 and relmax(c["B"], o["B"]) < 1e-13
    assert rel(o["delta"], c["δ"]) < 1e-14 and rel(o["g"], c["g"]) < 1e-13 and rel(o["gd"], c["gδ"]) < 1e-13 and rel(o["bl"], c["bλ"]) < 1e-12
    assert rel(o["lam"], c["λ"]) < 1e-11 and rel(o["u"], c["u"]) < 1e-11 and rel(o["tauf"], c["τf"]) < 1e-11
    g = np.load(os.path.join(GOLD, "flower_p4.npz"))
    assert np.array_equal(g["lam"], c["λ"]) and np.array_equal(g["u"], c["u"]) and np.array_equal(g["traction"], c["τf"])


def test_multiblock_bp1_mesh_two_jump_codes_executed():
    """seas/BP1/meshes/BP1_v1.inp (194 blocks, 104 reversed faces, jump interfaces of two kinds: side sets 7 and 8 -- every
    `>= BC_JUMP_INTERFACE` branch of locoperator, locbcarray!, assembleλmatrix, bcstarts with a tuple of codes): the reference's
    functions driven by tests/refexec/trace_driver.jl against the oracle's, slip on both kinds of interfaces (p = 2, 7 points per line)"""
    from refexec.drivers import run_trace_driver
    from refexec.oracle_driver import oracle_square_circle_level
    from hybridsbp_b200 import flower, host
    c = run_trace_driver("seas/BP1/meshes/BP1_v1.inp", 2, 6, (7, 8), 0.4)
    o = oracle_square_circle_level(2, 6, mesh=host.read_inp_2d(os.path.join(ROOT, "meshes", "BP1_v1.inp")), maps=flower.block_maps,
                                   exact=flower.Smooth, slip=lambda x, y: 0.3 * np.sin(x + 0.4) * np.cos(2 * y))
    FToB = np.asarray(c["FToB"])
    assert {int(k): int((FToB == k).sum()) for k in np.unique(FToB)} == {0: 346, 1: 10, 2: 30, 7: 13, 8: 9}
    assert int((~np.asarray(c["EToO"]).astype(bool)).sum()) == 104
    assert np.array_equal(o["FTol"], c["FToλstarts"]) and np.array_equal(o["FTod"], c["FToδstarts"])
    rel = lambda a, b: np.linalg.norm(np.asarray(a) - np.asarray(b)) / np.linalg.norm(np.asarray(b))
    assert np.linalg.norm(c["δ"]) > 1 and rel(o["delta"], c["δ"]) < 1e-12 and rel(o["g"], c["g"]) < 1e-13 and rel(o["gd"], c["gδ"]) < 1e-12
    assert rel(o["bl"], c["bλ"]) < 1e-12 and rel(o["lam"], c["λ"]) < 1e-11 and rel(o["u"], c["u"]) < 1e-11 and rel(o["tauf"], c["τf"]) < 1e-10


# ---- BP1 ---------------------------------------------------------------------------------------------------------------------------
def test_bp1_setup_and_odefun_executed():
    import gen_refexec_golden as gg
    from refexec.drivers import run_bp1_setup
    from hybridsbp_b200 import bp1
    from oracle.bp1 import OdeFun
    N = 40
    it, sol, yf = run_bp1_setup(N)
    prob, opts = sol.get("prob"), sol.get("options")
    su = bp1.setup(N=N)                                                    # the product's host-side setup
    assert np.allclose(np.array(prob.get("u0")), su.psi_delta0, rtol=1e-14, atol=0)
    assert np.allclose(yf, su.yf, rtol=1e-14, atol=1e-14)
    assert np.allclose(np.array(prob.get("p").get("RSa")), su.RSa, rtol=1e-15)
    assert opts["atol"] == 1e-5 and opts["rtol"] == 1e-3 and opts["dt"] == 31556926 and tuple(prob.get("tspan")) == (0, 1000.0 * 31556926)
    o = OdeFun(su.p, su.N, su.metrics, su.LFtoB, su.RSa, su.params)
    for t, y in gg.bp1_states(np.array(prob.get("u0")), N):
        d = np.zeros(2 * (N + 1))
        prob.get("f")(d, y.copy(), prob.get("p"), t)
        do, rejected = o(t, y)
        assert not rejected and not prob.get("p").get("reject_step")[0]
        assert np.max(np.abs(d - do)) <= 1e-13 * np.max(np.abs(do))
    y = np.array(prob.get("u0")); y[3] = np.nan                            # a failed node: reject_step is set, no exception
    prob.get("f")(np.zeros(2 * (N + 1)), y, prob.get("p"), 0.0)
    assert prob.get("p").get("reject_step")[0] and o(0.0, y)[1]


def test_bp1_odefun_at_the_reference_resolution_executed():
    """N = 200, p = 2 (BP1.jl:28-32 as written): the reference's odefun at the five stored states of tests/golden/bp1/states_N200.npz
    -- initial, interseismic, the middle of the first earthquake (max V = 0.7 m/s), post-seismic -- against the stored outputs the
    GPU tests use (they were generated by the oracle)"""
    from refexec.drivers import run_bp1_setup
    g = np.load(os.path.join(ROOT, "tests", "golden", "bp1", "states_N200.npz"))
    it, sol, yf = run_bp1_setup(200)
    prob = sol.get("prob")
    for k in range(len(g["t"])):
        d = np.zeros(402)
        prob.get("f")(d, g["y"][k].copy(), prob.get("p"), float(g["t"][k]))
        ref = g["dpsiV"][k]
        assert not prob.get("p").get("reject_step")[0]
        assert np.max(np.abs(d[201:] - ref[201:])) <= 1e-12 * np.max(np.abs(ref[201:]))
        assert np.max(np.abs(d[:201] - ref[:201])) <= 1e-12 * np.max(np.abs(ref[:201]))
    assert np.max(np.abs(g["dpsiV"][3][201:])) > 0.5                        # the coseismic state is among them


# ---- the reference's own check scripts, as written -----------------------------------------------------------------------------------
def test_reference_check_scripts_executed():
    """check_residual.jl, global_op_eigenvalues.jl and local_op_eigenvalues.jl run through the interpreter (random samples reduced
    from 1000 to 2, plotting statements dropped): their assertions hold and the quantities they plot have the advertised signs"""
    from refexec.drivers import run_check_script
    _, log = run_check_script("check_residual.jl")
    assert len(log) == 6                                                   # extrema of Re, Im of eig(A - D1' H diag(b) D1) for p = 2, 4, 6
    for re_ext, im_ext in zip(log[0::2], log[1::2]):
        assert re_ext[0] > -1e-12 and re_ext[1] > 1 and im_ext == (0.0, 0.0)
    names = ("min_eig_noSchur", "min_eig_Schur_M", "min_eig_Schur_D")
    cap, _ = run_check_script("global_op_eigenvalues.jl", 2, names)       # contains @assert B ≈ D - F̄ᵀ A11⁻¹ F̄  (:84)
    for k in names:
        assert np.asarray(cap[0][k]).shape == (3, 2) and np.asarray(cap[0][k]).min() > 0
    cap, _ = run_check_script("local_op_eigenvalues.jl", 2, ("min_eig", "max_eig"))
    assert np.asarray(cap[0]["min_eig"]).shape == (3, 2, 2) and np.asarray(cap[0]["min_eig"]).min() > 0      # SPD for Dirichlet and Neumann faces
    sweep = np.asarray(cap[1]["min_eig"])                                  # tau scale 1e-2 ... 1e2: indefinite when the penalty is too small
    assert sweep[:, 0].max() < 0 and sweep[:, -1].min() > 0
    it = Interp(os.path.join(REF, "seas", "BP1"))                          # single_block.jl: u = 1 is reproduced on the stretched block
    it.include("single_block.jl")
    assert it.log[0] == (1, 4) and np.max(np.abs(np.asarray(it.log[-1]) - 1)) < 1e-9


# ---- the committed golden vectors are what the generator writes ----------------------------------------------------------------------
def test_golden_vectors_are_reproducible(tmp_path, square_circle_p4, monkeypatch):
    import gen_refexec_golden as gg
    monkeypatch.setattr(gg, "OUT", str(tmp_path))
    gg.gen_locoperator(4); gg.gen_bp1(40)
    for name in ("locoperator_p4.npz", "bp1_odefun_N40.npz"):
        a, b = np.load(os.path.join(GOLD, name)), np.load(os.path.join(str(tmp_path), name))
        assert sorted(a.files) == sorted(b.files)
        for k in a.files:
            assert np.array_equal(a[k], b[k]), (name, k)
    g = np.load(os.path.join(GOLD, "square_circle_p4.npz"))
    c = square_circle_p4[0]
    assert np.array_equal(g["lam"], c["λ"]) and np.array_equal(g["u"], c["u"]) and np.array_equal(g["gdelta"], c["gδ"])
