"""Host-side mirror of the reference's entry points for the hybridized SBP path.

The north star keeps the host code in the reference's own language (Julia); that toolchain is
not installed in the build container, so the host side above the C-ABI is this Python module
with the SAME entry points, argument meaning and array conventions (julia/HybridSBPB200.jl is
the `ccall` twin):

  read_inp_2d            global_curved.jl:802-956     Abaqus .inp -> verts, EToV, EToF, FToB, EToBlock
  connectivityarrays     global_curved.jl:82-132      -> FToE, FToLF, EToO, EToS
  transfinite_blend      global_curved.jl:19-78       (three call forms)
  create_metrics         global_curved.jl:136-209
  locoperator            global_curved.jl:211-506     -> LocalOperator (host record; nothing assembled)
  bcstarts               global_curved.jl:714-728

Only cheap O(mesh) host work happens here.  Everything the reference spends its time on
(assembling and factorising sparse matrices) is replaced by libhsbp's CUDA kernels.

Array conventions are the reference's: ids stored in arrays are 1-based, EToV/EToF are 4 x ne,
FToE/FToLF are 2 x nf, fields are (Nr+1) x (Ns+1) with r the first index (flattened r-fastest).
"""
import re
from dataclasses import dataclass, field
from typing import Tuple

import numpy as np

BC_DIRICHLET = 1
BC_NEUMANN = 2
BC_LOCKED_INTERFACE = 0
BC_JUMP_INTERFACE = 7

# local face -> its two local vertices (z-order numbering of the block corners)
_FACE_VERTS = np.array([[0, 2], [1, 3], [0, 1], [2, 3]])


# ------------------------------------------------------------------------------------------
# mesh reader
# ------------------------------------------------------------------------------------------
_RE_DATA = re.compile(r"^\s*[0-9]*\s*,.*")
_RE_LEADINT = re.compile(r"^\s*[0-9]+")


def _tokens(line):
    return [t for t in re.split(r"[\s,]", line) if t]


def _find(lines, needle, start=0):
    rx = re.compile(needle)
    for i in range(start, len(lines)):
        if rx.search(lines[i]):
            return i
    return -1


def read_inp_2d(filename, bc_map=None):
    """Parse a Cubit/Abaqus .inp quad mesh exactly like the reference reader does.

    Returns (verts[2, nv], EToV[4, ne], EToF[4, ne], FToB[nf], EToBlock[ne]).  Side-set id i is
    mapped to bc_map[i-1]; a mapped value of 3 becomes a locked interface.
    """
    try:
        with open(filename) as fh:
            lines = fh.read().split("\n")
    except OSError:
        raise RuntimeError('InpRead cannot open "%s" ' % filename)
    bc_of = (lambda sid: sid) if bc_map is None else (lambda sid: bc_map[sid - 1])

    # nodes
    at = _find(lines, "NSET=ALLNODES")
    if at < 0:
        raise RuntimeError("did not find: NSET=ALLNODES")
    rows = []
    for l in lines[at + 1:]:
        if not _RE_DATA.match(l):
            break
        rows.append(_tokens(l))
    verts = np.full((2, len(rows)), np.nan)
    for t in rows:
        verts[:, int(t[0]) - 1] = (float(t[1]), float(t[2]))

    # elements: every *ELEMENT section; the trailing integer of the header is the block id
    sections = []
    at = _find(lines, "ELEMENT")
    while at >= 0:
        blk = int(re.findall(r"[0-9]+", lines[at])[-1])
        body = []
        for l in lines[at + 1:]:
            if not _RE_DATA.match(l):
                break
            body.append([int(x) for x in _tokens(l)[:5]])
        sections.append((blk, body))
        at = _find(lines, "ELEMENT", at + 1)
    ne = sum(len(b) for _, b in sections)
    if ne == 0:
        raise RuntimeError("did not find any element")
    EToV = np.zeros((4, ne), dtype=np.int64)
    EToBlock = np.zeros(ne, dtype=np.int64)
    for blk, body in sections:
        for en, a, b, c, d in body:
            EToV[:, en - 1] = (a, b, d, c)          # counter-clockwise file order -> z-order
            EToBlock[en - 1] = blk

    # faces, numbered in order of first appearance (element-major, local face minor)
    EToF = np.zeros((4, ne), dtype=np.int64)
    known = {}
    for e in range(ne):
        for lf in range(4):
            a, b = EToV[_FACE_VERTS[lf], e]
            key = (a, b) if a <= b else (b, a)
            EToF[lf, e] = known.setdefault(key, len(known) + 1)
    FToB = np.zeros(len(known), dtype=np.int64)       # BC_LOCKED_INTERFACE

    # side sets: "*ELSET, ELSET=SS<set>_E<inpface>"
    inp_face_to_local = (3, 2, 4, 1)
    at = _find(lines, r"\*ELSET")
    while at >= 0:
        nums = re.findall(r"[0-9]+", lines[at])
        bc = bc_of(int(nums[0]))
        lf = inp_face_to_local[int(nums[1]) - 1]
        if bc == 3:
            bc = BC_LOCKED_INTERFACE
        if not (bc in (BC_DIRICHLET, BC_NEUMANN, BC_LOCKED_INTERFACE) or bc >= BC_JUMP_INTERFACE):
            raise ValueError("invalid bc code %r in side set" % (bc,))
        for l in lines[at + 1:]:
            if not _RE_LEADINT.match(l):
                break
            for tok in _tokens(l):
                FToB[EToF[lf - 1, int(tok) - 1] - 1] = bc
        at = _find(lines, r"\*ELSET", at + 1)
    return verts, EToV, EToF, FToB, EToBlock


# ------------------------------------------------------------------------------------------
# connectivity
# ------------------------------------------------------------------------------------------
def connectivityarrays(EToV, EToF):
    """(FToE, FToLF, EToO, EToS): the first block that owns a face is its minus side (EToS=1,
    EToO=True); the second is the plus side, EToO tells whether it runs along the face the same way."""
    EToV = np.asarray(EToV, dtype=np.int64)
    EToF = np.asarray(EToF, dtype=np.int64)
    ne = EToV.shape[1]
    nf = int(EToF.max())
    FToE = np.zeros((2, nf), dtype=np.int64)
    FToLF = np.zeros((2, nf), dtype=np.int64)
    EToO = np.ones((4, ne), dtype=bool)
    EToS = np.zeros((4, ne), dtype=np.int64)
    for e in range(ne):
        for lf in range(4):
            f = EToF[lf, e] - 1
            side = 0 if FToE[0, f] == 0 else 1
            if side == 1 and FToE[1, f] != 0:
                raise RuntimeError("problem with connectivity: face %d has more than two blocks" % (f + 1))
            FToE[side, f] = e + 1
            FToLF[side, f] = lf + 1
            EToS[lf, e] = side + 1
            if side == 1:
                mine = tuple(EToV[_FACE_VERTS[lf], e])
                e0, lf0 = FToE[0, f] - 1, FToLF[0, f] - 1
                theirs = tuple(EToV[_FACE_VERTS[lf0], e0])
                if mine == theirs:
                    EToO[lf, e] = True
                elif mine == theirs[::-1]:
                    EToO[lf, e] = False
                else:
                    raise RuntimeError("problem with connectivity")
    return FToE, FToLF, EToO, EToS


# ------------------------------------------------------------------------------------------
# transfinite blend
# ------------------------------------------------------------------------------------------
def d1_matrix(p, N):
    """Dense first-derivative SBP operator on [-1, 1] with N + 1 points (diagonal_sbp.jl:67-157)."""
    from ._sbp_d1 import D1_TABLES
    if p not in D1_TABLES:
        raise ValueError("Operators for order %d are not implemented" % p)
    d, bd = np.array(D1_TABLES[p]["d"]), np.array(D1_TABLES[p]["bd"])
    bm, bn = bd.shape
    Np = N + 1
    if Np < 2 * bm or Np < bn:
        raise ValueError("Grid not big enough to support the operator")
    D = np.zeros((Np, Np))
    D[:bm, :bn] = bd                                     # :144-146
    D[Np - bm:, Np - bn:] = -bd[::-1, ::-1]              # :147-149
    h = p // 2
    for i in range(bm, Np - bm):                         # :139-142
        D[i, i - h:i + h + 1] = d
    return D / (2.0 / N)


def transfinite_blend(*args):
    """Three call forms, as in the reference:
      transfinite_blend(a1, a2, a3, a4, a1s, a2s, a3r, a4r, r, s)   edge curves + derivatives        (global_curved.jl:19-51)
      transfinite_blend(a1, a2, a3, a4, r, s, p)                    edge derivatives by the SBP operator of order p (:53-64)
      transfinite_blend(v1, v2, v3, v4, r, s)                       straight block from corner values  (:66-78)
    Returns (x, x_r, x_s)."""
    if len(args) == 7:
        a1, a2, a3, a4, r, s, p = args
        Nrp, Nsp = r.shape
        Dr, Ds = d1_matrix(p, Nrp - 1), d1_matrix(p, Nsp - 1)
        return transfinite_blend(a1, a2, a3, a4, lambda t: a1(t) @ Ds.T, lambda t: a2(t) @ Ds.T,
                                 lambda t: Dr @ a3(r), lambda t: Dr @ a4(r), r, s)          # as written: a3r, a4r differentiate a(r)
    if len(args) == 6:
        v1, v2, v3, v4, r, s = args
        lin = lambda a, b: (lambda t: a * (1 - t) / 2 + b * (1 + t) / 2)
        con = lambda a, b: (lambda t: (b - a) / 2)
        return transfinite_blend(lin(v1, v3), lin(v2, v4), lin(v1, v2), lin(v3, v4),
                                 con(v1, v3), con(v2, v4), con(v1, v2), con(v3, v4), r, s)
    a1, a2, a3, a4, a1s, a2s, a3r, a4r, r, s = args
    c11, c21, c12, c22 = a1(-1.0), a2(-1.0), a1(1.0), a2(1.0)      # corners (r,s) = (-,-) (+,-) (-,+) (+,+)
    if not np.allclose([c11, c21, c12, c22], [a3(-1.0), a3(1.0), a4(-1.0), a4(1.0)]):
        raise AssertionError("edge curves do not meet at the corners")
    rp, rm, sp, sm = 1 + r, 1 - r, 1 + s, 1 - s
    x = (rp * a2(s) + rm * a1(s) + sp * a4(r) + sm * a3(r)) / 2 \
        - (rp * sp * c22 + rm * sp * c12 + rp * sm * c21 + rm * sm * c11) / 4
    xr = (a2(s) - a1(s) + sp * a4r(r) + sm * a3r(r)) / 2 - (sp * (c22 - c12) + sm * (c21 - c11)) / 4
    xs = (rp * a2s(s) + rm * a1s(s) + a4(r) - a3(r)) / 2 - (rp * (c22 - c21) + rm * (c12 - c11)) / 4
    return x, xr, xs


# ------------------------------------------------------------------------------------------
# metrics
# ------------------------------------------------------------------------------------------
@dataclass
class Metrics:
    coord: Tuple[np.ndarray, np.ndarray]
    facecoord: Tuple[Tuple[np.ndarray, ...], Tuple[np.ndarray, ...]]
    crr: np.ndarray
    css: np.ndarray
    crs: np.ndarray
    J: np.ndarray
    sJ: Tuple[np.ndarray, ...]
    nx: Tuple[np.ndarray, ...]
    ny: Tuple[np.ndarray, ...]
    rx: np.ndarray
    ry: np.ndarray
    sx: np.ndarray
    sy: np.ndarray


def reference_grid(Nr, Ns):
    r = np.linspace(-1.0, 1.0, Nr + 1)[:, None] * np.ones((1, Ns + 1))
    s = np.ones((Nr + 1, 1)) * np.linspace(-1.0, 1.0, Ns + 1)[None, :]
    return r, s


def create_metrics(pm, Nr, Ns, xf=None, yf=None):
    """Curvilinear metric terms of one block.  xf, yf: (r, s) -> (x, x_r, x_s)."""
    if pm > 8:
        raise AssertionError("pm <= 8")
    r, s = reference_grid(Nr, Ns)
    shp = r.shape
    if xf is None:
        x, xr, xs = r, np.ones(shp), np.zeros(shp)
    else:
        x, xr, xs = (np.broadcast_to(np.asarray(a, float), shp).copy() for a in xf(r, s))
    if yf is None:
        y, yr, ys = s, np.zeros(shp), np.ones(shp)
    else:
        y, yr, ys = (np.broadcast_to(np.asarray(a, float), shp).copy() for a in yf(r, s))
    J = xr * ys - xs * yr
    if not J.min() > 0:
        raise AssertionError("non-positive Jacobian")
    rx, sx, ry, sy = ys / J, -yr / J, -xs / J, xr / J
    crr = J * (rx * rx + ry * ry)
    crs = J * (sx * rx + sy * ry)
    css = J * (sx * sx + sy * sy)
    # outward (unnormalised) normals of faces 1..4 and their lengths
    raw = ((-ys[0, :], xs[0, :]), (ys[-1, :], -xs[-1, :]), (yr[:, 0], -xr[:, 0]), (-yr[:, -1], xr[:, -1]))
    sJ = tuple(np.hypot(a, b) for a, b in raw)
    nx = tuple(a / l for (a, _), l in zip(raw, sJ))
    ny = tuple(b / l for (_, b), l in zip(raw, sJ))
    fx = (x[0, :].copy(), x[-1, :].copy(), x[:, 0].copy(), x[:, -1].copy())
    fy = (y[0, :].copy(), y[-1, :].copy(), y[:, 0].copy(), y[:, -1].copy())
    return Metrics((x, y), (fx, fy), crr, css, crs, J, sJ, nx, ny, rx, ry, sx, sy)


# ------------------------------------------------------------------------------------------
# 1-D norm weights (needed on the host only for JH and the face norms of error measures)
# ------------------------------------------------------------------------------------------
_HW = {2: np.array([1 / 2]),
       4: np.array([17 / 48, 59 / 48, 43 / 48, 49 / 48]),
       6: np.array([13649 / 43200, 12013 / 8640, 2711 / 4320, 5359 / 4320, 7877 / 8640, 43801 / 43200])}


def norm_weights(p, N):
    """Diagonal of the SBP norm H on [-1, 1] with N+1 points (diagonal_sbp.jl:133-139)."""
    w = np.ones(N + 1)
    b = _HW[p]
    w[:b.size] = b
    w[N + 1 - b.size:] = b[::-1]
    return w * (2.0 / N)


@dataclass
class LocalOperator:
    """What the host needs to know about one block.  The reference's locoperator returns assembled
    sparse matrices here; in this build the operators live on the GPU (Blocks) and this record
    only carries geometry and boundary-condition data."""
    p: int
    Nr: int
    Ns: int
    metrics: Metrics
    bctype: Tuple[int, int, int, int]
    tauscale: float
    crr: np.ndarray
    css: np.ndarray
    crs: np.ndarray
    coord: Tuple[np.ndarray, np.ndarray] = field(init=False)
    facecoord: tuple = field(init=False)
    sJ: tuple = field(init=False)
    nx: tuple = field(init=False)
    ny: tuple = field(init=False)

    def __post_init__(self):
        m = self.metrics
        self.coord, self.facecoord, self.sJ, self.nx, self.ny = m.coord, m.facecoord, m.sJ, m.nx, m.ny

    @property
    def JH(self):
        """diag of J * (Hs kron Hr), r fastest (global_curved.jl:491)."""
        hr, hs = norm_weights(self.p, self.Nr), norm_weights(self.p, self.Ns)
        return (self.metrics.J * hr[:, None] * hs[None, :]).reshape(-1, order="F")

    def Hf(self, lf):
        """diag of the face norm of local face lf (1-based)."""
        return norm_weights(self.p, self.Ns if lf <= 2 else self.Nr)


def locoperator(p, Nr, Ns, metrics=None, LFToB=(BC_DIRICHLET,) * 4, tauscale=2.0, crr=None, css=None, crs=None):
    """Same call as the reference's locoperator (tauscale is its keyword τscale)."""
    if p not in (2, 4, 6):
        raise ValueError("unknown order")
    if metrics is None:
        metrics = create_metrics(p, Nr, Ns)
    for b in LFToB:
        if not (b in (BC_DIRICHLET, BC_NEUMANN, BC_LOCKED_INTERFACE) or b >= BC_JUMP_INTERFACE):
            raise ValueError("invalid bc")
    return LocalOperator(p, Nr, Ns, metrics, tuple(int(b) for b in LFToB), float(tauscale),
                         metrics.crr if crr is None else crr,
                         metrics.css if css is None else css,
                         metrics.crs if crs is None else crs)


def bcstarts(FToB, FToE, FToLF, bctype, Nr, Ns):
    """1-based offsets of the faces whose code is in bctype (layout of the jump vector delta)."""
    codes = (bctype,) if np.isscalar(bctype) else tuple(bctype)
    nf = len(FToB)
    npts = np.zeros(nf, dtype=np.int64)
    for f in range(nf):
        if FToB[f] in codes:
            e, lf = FToE[0, f] - 1, FToLF[0, f]
            npts[f] = (Ns[e] if lf <= 2 else Nr[e]) + 1
    return np.concatenate([[1], 1 + np.cumsum(npts)]).astype(np.int64)
