"""Device-resident block-local operators (thin object wrapper over the hsbp_blocks_* C-ABI).

Replaces what the reference's `locoperator` assembles per block -- the sparse
M-tilde, F_k, HfI_FT_k (global_curved.jl:211-506) -- by matrix-free CUDA kernels.
"""
import ctypes as C
import weakref

import numpy as np

from ._lib import Context, DeviceArray, LocalStats, TraceStats, _f64, _i64, lib

LOCAL_PCG = 1
LOCAL_CHOLESKY = 2
LOCAL_BAND = 3
LOCAL_FDM = 4


class Blocks:
    """A set of blocks on one GPU.

    p: SBP interior order (2, 4, 6); Nr, Ns: per-block grid sizes (length nblocks).
    Volume vectors are the blocks' (Nr+1)x(Ns+1) fields, r fastest, concatenated
    (the reference's `vstarts` layout); face vectors hold faces 1..4 of each block.
    """

    def __init__(self, ctx: Context, p, Nr, Ns):
        self.ctx = ctx
        self.p = int(p)
        self.Nr = np.ascontiguousarray(Nr, dtype=np.int64).ravel()
        self.Ns = np.ascontiguousarray(Ns, dtype=np.int64).ravel()
        assert self.Nr.shape == self.Ns.shape
        self.nblocks = self.Nr.size
        h = C.c_void_p()
        _, pNr = _i64(self.Nr)
        _, pNs = _i64(self.Ns)
        ctx._check(lib().hsbp_blocks_create(ctx.h, self.p, self.nblocks, pNr, pNs, C.byref(h)))
        self.h = h
        self._children = weakref.WeakSet()       # Trace / bp1.Fault objects built on these blocks (closed first)
        ctx._children.add(self)
        self.VNp = lib().hsbp_blocks_num_volume_points(h)
        self.FNp = lib().hsbp_blocks_num_face_points(h)
        npts = (self.Nr + 1) * (self.Ns + 1)
        self.vstarts = np.concatenate([[1], 1 + np.cumsum(npts)]).astype(np.int64)     # 1-based, as the reference
        nface = 2 * (self.Nr + 1) + 2 * (self.Ns + 1)
        self.fstarts = np.concatenate([[0], np.cumsum(nface)]).astype(np.int64)         # 0-based block face offsets

    # -- setup -------------------------------------------------------------------------------
    def set_metrics(self, crr, css, crs):
        a, pa = _f64(crr); b, pb = _f64(css); c, pc = _f64(crs)
        assert a.size == b.size == c.size == self.VNp
        self.ctx._check(lib().hsbp_blocks_set_metrics(self.h, pa, pb, pc))

    def set_metrics_dev(self, crr: DeviceArray, css: DeviceArray, crs: DeviceArray):
        self.ctx._check(lib().hsbp_blocks_set_metrics_dev(self.h, crr.ptr, css.ptr, crs.ptr))

    # -- geometry on the device (transfinite_blend / create_metrics, global_curved.jl:19-51, 136-209) ----------
    def blend_dev(self, edges: DeviceArray, x: DeviceArray, xr: DeviceArray, xs: DeviceArray):
        """x, x_r, x_s of every block from its edge curves: `edges` holds per block [a1 | a2 | a3 | a4] sampled at the grid
        points followed by [a1' | a2' | a3' | a4'] (2 * FNp doubles in all)"""
        assert edges.n == 2 * self.FNp
        self.ctx._check(lib().hsbp_blocks_blend_dev(self.h, edges.ptr, x.ptr, xr.ptr, xs.ptr))

    def set_geometry_dev(self, xr, xs, yr, ys, J=None, sJ=None, nx=None, ny=None):
        """create_metrics on the device: coefficient fields from the derivatives of the block maps; optional outputs J
        (volume layout) and sJ, nx, ny (block-face layout)"""
        P = lambda a: a.ptr if a is not None else None
        self.ctx._check(lib().hsbp_blocks_set_geometry_dev(self.h, xr.ptr, xs.ptr, yr.ptr, ys.ptr, P(J), P(sJ), P(nx), P(ny)))

    def set_synthetic_warp(self, nbx, bx0, L, A, x=None, y=None):
        """metrics of the synthetic warped multiblock mesh generated on the device"""
        P = lambda a: a.ptr if a is not None else None
        self.ctx._check(lib().hsbp_blocks_set_synthetic_warp(self.h, int(nbx), int(bx0), float(L), float(A), P(x), P(y)))

    def set_bc(self, bctype):
        a, pa = _i64(np.asarray(bctype).reshape(-1))
        assert a.size == 4 * self.nblocks
        self.ctx._check(lib().hsbp_blocks_set_bc(self.h, pa))

    def compute_tau(self, tauscale=2.0):
        self.ctx._check(lib().hsbp_blocks_compute_tau(self.h, float(tauscale)))

    def set_tau(self, tau):
        a, pa = _f64(tau)
        assert a.size == self.FNp
        self.ctx._check(lib().hsbp_blocks_set_tau(self.h, pa))

    def get_tau(self):
        out = np.empty(self.FNp)
        self.ctx._check(lib().hsbp_blocks_get_tau(self.h, C.c_void_p(out.ctypes.data)))
        return out

    def face_slice(self, e, lf):
        """Slice of block e's local face lf (1-based) inside a face vector."""
        nsp, nrp = self.Ns[e] + 1, self.Nr[e] + 1
        start = self.fstarts[e] + (0, nsp, 2 * nsp, 2 * nsp + nrp)[lf - 1]
        return slice(int(start), int(start + (nsp if lf <= 2 else nrp)))

    def vol_slice(self, e):
        return slice(int(self.vstarts[e] - 1), int(self.vstarts[e + 1] - 1))

    # -- operators ---------------------------------------------------------------------------
    def apply(self, u: DeviceArray, y: DeviceArray):
        """y = M-tilde u   (device vectors)."""
        self.ctx._check(lib().hsbp_apply(self.h, u.ptr, y.ptr))

    def apply_host(self, u, y=None):
        """y = M-tilde u through host buffers (H2D + kernels + D2H inside the call)."""
        u, pu = _f64(u)
        assert u.size == self.VNp
        if y is None:
            y = np.empty(self.VNp)
        self.ctx._check(lib().hsbp_apply_host(self.h, pu, C.c_void_p(y.ctypes.data)))
        return y

    def apply_energy(self, u: DeviceArray, y: DeviceArray):
        """y = M-tilde u; returns u_e . (M-tilde u)_e per block, computed inside the sweep kernel (hsbp_apply_energy)."""
        en = np.zeros(self.nblocks)
        self.ctx._check(lib().hsbp_apply_energy(self.h, u.ptr, y.ptr, C.c_void_p(en.ctypes.data)))
        return en

    def apply_timed(self, u: DeviceArray, y: DeviceArray):
        """apply with per-stage CUDA-event times: (volume ms, face gather ms, face scatter ms)."""
        ms = np.zeros(3)
        self.ctx._check(lib().hsbp_apply_timed(self.h, u.ptr, y.ptr, C.c_void_p(ms.ctypes.data)))
        return ms

    def apply_host_pinned(self, u, y):
        """apply_host on caller buffers that were registered with ctx.host_register."""
        self.ctx._check(lib().hsbp_apply_host(self.h, C.c_void_p(u.ctypes.data), C.c_void_p(y.ctypes.data)))

    def apply_variant(self):
        return lib().hsbp_apply_variant(self.h)

    def force_generic(self, on=True):
        self.ctx._check(lib().hsbp_blocks_force_generic(self.h, 1 if on else 0))

    def set_option(self, name, value):
        """tuning / testing knobs of the kernels, see hsbp_blocks_set_option in include/hsbp.h"""
        self.ctx._check(lib().hsbp_blocks_set_option(self.h, name.encode(), int(value)))

    def face_FT(self, u: DeviceArray, ft: DeviceArray):
        self.ctx._check(lib().hsbp_face_FT(self.h, u.ptr, ft.ptr))

    def face_F_add(self, v: DeviceArray, alpha, y: DeviceArray):
        self.ctx._check(lib().hsbp_face_F_add(self.h, v.ptr, float(alpha), y.ptr))

    def face_traction(self, u: DeviceArray, tr: DeviceArray):
        self.ctx._check(lib().hsbp_face_traction(self.h, u.ptr, tr.ptr))

    # -- local solves (the reference's `factorization` plugin) --------------------------------
    def local_setup(self, mode=LOCAL_PCG, tol=1e-13, maxit=100000):
        self.ctx._check(lib().hsbp_local_setup(self.h, int(mode), float(tol), int(maxit)))

    def local_solve(self, g: DeviceArray, u: DeviceArray):
        """u = M-tilde^-1 g for every block; returns the solver statistics."""
        st = LocalStats()
        self.ctx._check(lib().hsbp_local_solve(self.h, g.ptr, u.ptr, C.byref(st)))
        return st.as_dict()

    def local_precondition(self, r: DeviceArray, z: DeviceArray):
        """z = P^-1 r of the fast-diagonalisation preconditioner alone (testing / profiling hook)"""
        self.ctx._check(lib().hsbp_local_precondition(self.h, r.ptr, z.ptr))

    def close(self):
        if self.h is not None:
            for child in list(self._children):       # traces / BP1 stages point into this object on the C side
                child.close()
            if self.ctx.h is not None:
                lib().hsbp_blocks_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Trace:
    """Trace (lambda) operators of a multiblock mesh on top of a Blocks object.

    Replaces glo-lambda-operator's sparse Fbar^T / D (global_curved.jl:510-565) and the explicit Schur
    complement of assemble-lambda-matrix (:743-797) by matrix-free kernels and a device CG."""

    def __init__(self, blocks: Blocks, FToB, FToE, FToLF, EToO, EToS):
        self.blocks = blocks
        self.ctx = blocks.ctx
        FToB = np.ascontiguousarray(FToB, dtype=np.int64)
        nf = FToB.size
        col = lambda a, dt: np.ascontiguousarray(np.asarray(a).T, dtype=dt).reshape(-1)     # column-major flattening
        fe, fl = col(FToE, np.int64), col(FToLF, np.int64)
        eo, es = col(EToO, np.uint8), col(EToS, np.int64)
        assert fe.size == 2 * nf and eo.size == 4 * blocks.nblocks
        h = C.c_void_p()
        P = lambda a: a.ctypes.data_as(C.POINTER(C.c_int64))
        self.ctx._check(lib().hsbp_trace_create(blocks.h, nf, P(FToB), P(fe), P(fl), C.c_void_p(eo.ctypes.data),
                                                P(es), C.byref(h)))
        self.h = h
        blocks._children.add(self)
        self.nfaces = nf
        self.lNp = lib().hsbp_trace_num_lambda(h)
        self.FTolambdastarts = np.zeros(nf + 1, dtype=np.int64)
        self.ctx._check(lib().hsbp_trace_get_starts(h, P(self.FTolambdastarts)))

    def D(self):
        out = np.empty(self.lNp)
        self.ctx._check(lib().hsbp_trace_get_D(self.h, C.c_void_p(out.ctypes.data)))
        return out

    def set_partition(self, faces, partner, gamma, n_gamma_total):
        """tell the library which local faces are cut by the block partition (parallel.LocalMesh.partition_arrays);
        collective: completes D on the cut faces with the partner's half"""
        f, pf = _i64(faces); q, pq = _i64(partner); g, pg = _i64(gamma)
        assert f.size == q.size == g.size
        self.ctx._check(lib().hsbp_trace_set_partition(self.h, f.size, pf, pq, pg, int(n_gamma_total)))

    def FbarT(self, u: DeviceArray, lam: DeviceArray):
        self.ctx._check(lib().hsbp_trace_FbarT(self.h, u.ptr, lam.ptr))

    def Fbar_add(self, lam: DeviceArray, alpha, y: DeviceArray):
        self.ctx._check(lib().hsbp_trace_Fbar_add(self.h, lam.ptr, float(alpha), y.ptr))

    def condense(self, enable=True):
        """Form S_e = F_e^T M-tilde_e^-1 F_e per block (assembleλmatrix's products, global_curved.jl:759-790); later
        schur_apply / solve calls use them instead of local solves."""
        self.ctx._check(lib().hsbp_trace_condense(self.h, 1 if enable else 0))

    def precond_setup(self, kind=1):
        """0: Jacobi (D); 1: block-Jacobi with the exact diagonal blocks of B as explicit inverses (needs condense());
        collective on a partitioned mesh (the partner's half of a cut face's block is fetched)."""
        self.ctx._check(lib().hsbp_trace_precond_setup(self.h, int(kind)))

    def coarse_setup(self, modes=2):
        """second level: `modes` Legendre polynomials per face (0 = off); collective on a partitioned mesh"""
        self.ctx._check(lib().hsbp_trace_coarse_setup(self.h, int(modes)))

    def coarse_size(self):
        return lib().hsbp_trace_coarse_size(self.h)

    def set_option(self, name, value):
        self.ctx._check(lib().hsbp_trace_set_option(self.h, name.encode(), int(value)))

    def comm_path(self):
        """0: one rank, 1: NCCL, 2: peer memory (how the last solve exchanged data inside its iteration loop)"""
        return lib().hsbp_trace_comm_path(self.h)

    def last_local_stats(self):
        st = LocalStats()
        self.ctx._check(lib().hsbp_trace_last_local_stats(self.h, C.byref(st)))
        return st.as_dict()

    def precond_apply(self, r: DeviceArray, z: DeviceArray):
        self.ctx._check(lib().hsbp_trace_precond_apply(self.h, r.ptr, z.ptr))

    def schur_apply(self, lam: DeviceArray, out: DeviceArray):
        self.ctx._check(lib().hsbp_trace_schur_apply(self.h, lam.ptr, out.ptr))

    def rhs(self, g: DeviceArray, gdelta: DeviceArray, b: DeviceArray):
        self.ctx._check(lib().hsbp_trace_rhs(self.h, g.ptr, gdelta.ptr, b.ptr))

    def solve(self, g: DeviceArray, gdelta: DeviceArray, lam: DeviceArray, u: DeviceArray, tol=1e-10, maxit=10000):
        st = TraceStats()
        self.ctx._check(lib().hsbp_trace_solve(self.h, g.ptr, gdelta.ptr, lam.ptr, u.ptr, float(tol), int(maxit),
                                               C.byref(st)))
        return st.as_dict()

    def close(self):
        if self.h is not None:
            if self.blocks.h is not None and self.ctx.h is not None:     # the C object points into its blocks / context
                lib().hsbp_trace_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class SpdFactor:
    """The reference's `factorization` plugin at its own seam (hsbp_factor_*): a device-resident banded Cholesky factor
    of ONE assembled sparse SPD matrix, e.g. `lop[e].M̃` exactly as locoperator builds it (global_curved.jl:470-486).
    `F.solve(g)` is the reference's `F \\ g` (global_curved.jl:734); a 1 x 1 matrix -- the probe SBPLocalOperator1 makes to
    learn the factor type (:681) -- is fine."""

    def __init__(self, ctx: Context, A):
        import scipy.sparse as sp
        A = sp.csc_matrix(A)
        A.sort_indices()
        assert A.shape[0] == A.shape[1]
        self.ctx, self.n = ctx, A.shape[0]
        cp, pcp = _i64(A.indptr)
        ri, pri = _i64(A.indices)
        nz, pnz = _f64(A.data)
        h = C.c_void_p()
        ctx._check(lib().hsbp_factor_create(ctx.h, self.n, pcp, pri, pnz, 0, C.byref(h)))
        self.h = h
        ctx._children.add(self)

    def solve(self, g):
        g = np.ascontiguousarray(g, dtype=np.float64)
        cols = g.reshape(self.n, -1, order="F") if g.ndim > 1 else g.reshape(self.n, 1)
        gin = np.ascontiguousarray(cols.T)                          # right-hand sides one after the other
        out = np.empty_like(gin)
        self.ctx._check(lib().hsbp_factor_solve(self.h, C.c_void_p(gin.ctypes.data), C.c_void_p(out.ctypes.data), gin.shape[0]))
        return out.T.reshape(g.shape, order="F") if g.ndim > 1 else out[0]

    def close(self):
        if self.h is not None:
            if self.ctx.h is not None:
                lib().hsbp_factor_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
