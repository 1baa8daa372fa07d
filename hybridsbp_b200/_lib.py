"""ctypes binding of libhsbp.so (the C-ABI declared in include/hsbp.h).

This is the Python twin of the `ccall` layer a Julia host would use
(INTEGRATION.md): same symbols, same argument order, no array abstraction in
between.  There is deliberately no fallback: if the shared library is missing
or a call fails, an exception is raised.
"""
import ctypes as C
import weakref
import os
import re

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("HSBP_LIB", os.path.join(_HERE, "libhsbp.so"))   # HSBP_LIB: kernel-variant experiments
HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "hsbp.h")


class HsbpError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("libhsbp error %d: %s" % (code, msg))
        self.code = code


def declared_symbols(header=HEADER_PATH):
    """Names of all functions include/hsbp.h declares."""
    txt = open(header).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(hsbp_[A-Za-z0-9_]+)\s*\(", txt)))


_lib = None


def lib():
    """Load libhsbp.so (once).  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise HsbpError(-100, "libhsbp.so not built (run python -c 'import __graft_entry__ as g; g.build()'); "
                              "there is no CPU fallback")
    L = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
    vp, i64p, dp, cint, i64, dbl = C.c_void_p, C.POINTER(C.c_int64), C.c_void_p, C.c_int, C.c_int64, C.c_double
    sig = {
        "hsbp_version": (cint, []),
        "hsbp_ctx_create": (cint, [cint, C.POINTER(vp)]),
        "hsbp_ctx_destroy": (cint, [vp]),
        "hsbp_last_error": (C.c_char_p, [vp]),
        "hsbp_malloc": (cint, [vp, C.c_size_t, C.POINTER(vp)]),
        "hsbp_free": (cint, [vp, vp]),
        "hsbp_h2d": (cint, [vp, vp, vp, C.c_size_t]),
        "hsbp_d2h": (cint, [vp, vp, vp, C.c_size_t]),
        "hsbp_memset0": (cint, [vp, vp, C.c_size_t]),
        "hsbp_sync": (cint, [vp]),
        "hsbp_host_register": (cint, [vp, vp, C.c_size_t]),
        "hsbp_host_unregister": (cint, [vp, vp]),
        "hsbp_stream": (vp, [vp]),
        "hsbp_timer_start": (cint, [vp]),
        "hsbp_timer_stop": (cint, [vp, C.POINTER(dbl)]),
        "hsbp_blocks_create": (cint, [vp, cint, i64, i64p, i64p, C.POINTER(vp)]),
        "hsbp_blocks_destroy": (cint, [vp]),
        "hsbp_blocks_num_volume_points": (i64, [vp]),
        "hsbp_blocks_num_face_points": (i64, [vp]),
        "hsbp_blocks_set_metrics": (cint, [vp, dp, dp, dp]),
        "hsbp_blocks_set_metrics_dev": (cint, [vp, dp, dp, dp]),
        "hsbp_blocks_blend_dev": (cint, [vp, dp, dp, dp, dp]),
        "hsbp_blocks_set_geometry_dev": (cint, [vp, dp, dp, dp, dp, dp, dp, dp, dp]),
        "hsbp_blocks_set_synthetic_warp": (cint, [vp, i64, i64, dbl, dbl, dp, dp]),
        "hsbp_blocks_set_bc": (cint, [vp, i64p]),
        "hsbp_blocks_compute_tau": (cint, [vp, dbl]),
        "hsbp_blocks_set_tau": (cint, [vp, dp]),
        "hsbp_blocks_get_tau": (cint, [vp, dp]),
        "hsbp_apply": (cint, [vp, dp, dp]),
        "hsbp_apply_host": (cint, [vp, dp, dp]),
        "hsbp_apply_energy": (cint, [vp, dp, dp, dp]),
        "hsbp_apply_timed": (cint, [vp, dp, dp, dp]),
        "hsbp_apply_variant": (cint, [vp]),
        "hsbp_blocks_force_generic": (cint, [vp, cint]),
        "hsbp_blocks_set_option": (cint, [vp, C.c_char_p, i64]),
        "hsbp_face_FT": (cint, [vp, dp, dp]),
        "hsbp_face_F_add": (cint, [vp, dp, dbl, dp]),
        "hsbp_face_traction": (cint, [vp, dp, dp]),
        "hsbp_local_setup": (cint, [vp, cint, dbl, i64]),
        "hsbp_local_solve": (cint, [vp, dp, dp, vp]),
        "hsbp_local_precondition": (cint, [vp, dp, dp]),
        "hsbp_factor_create": (cint, [vp, i64, i64p, i64p, dp, cint, C.POINTER(vp)]),
        "hsbp_factor_destroy": (cint, [vp]),
        "hsbp_factor_size": (i64, [vp]),
        "hsbp_factor_solve": (cint, [vp, dp, dp, i64]),
        "hsbp_factor_solve_dev": (cint, [vp, dp, dp]),
        "hsbp_trace_create": (cint, [vp, i64, i64p, i64p, i64p, vp, i64p, C.POINTER(vp)]),
        "hsbp_trace_destroy": (cint, [vp]),
        "hsbp_trace_num_lambda": (i64, [vp]),
        "hsbp_trace_comm_path": (cint, [vp]),
        "hsbp_trace_get_starts": (cint, [vp, i64p]),
        "hsbp_trace_get_D": (cint, [vp, dp]),
        "hsbp_trace_FbarT": (cint, [vp, dp, dp]),
        "hsbp_trace_Fbar_add": (cint, [vp, dp, dbl, dp]),
        "hsbp_trace_schur_apply": (cint, [vp, dp, dp]),
        "hsbp_trace_condense": (cint, [vp, cint]),
        "hsbp_trace_precond_setup": (cint, [vp, cint]),
        "hsbp_trace_precond_apply": (cint, [vp, dp, dp]),
        "hsbp_trace_coarse_setup": (cint, [vp, cint]),
        "hsbp_trace_coarse_size": (i64, [vp]),
        "hsbp_trace_set_option": (cint, [vp, C.c_char_p, i64]),
        "hsbp_trace_last_local_stats": (cint, [vp, vp]),
        "hsbp_trace_set_partition": (cint, [vp, i64, i64p, i64p, i64p, i64]),
        "hsbp_comm_unique_id": (cint, [vp]),
        "hsbp_comm_init": (cint, [vp, vp, cint, cint]),
        "hsbp_comm_destroy": (cint, [vp]),
        "hsbp_comm_rank": (cint, [vp]),
        "hsbp_comm_world": (cint, [vp]),
        "hsbp_comm_allreduce_sum": (cint, [vp, dp, i64]),
        "hsbp_trace_rhs": (cint, [vp, dp, dp, dp]),
        "hsbp_trace_solve": (cint, [vp, dp, dp, dp, dp, dbl, i64, vp]),
        "hsbp_bp1_create": (cint, [vp, i64, i64, i64, dp, dp, vp, C.POINTER(vp)]),
        "hsbp_bp1_destroy": (cint, [vp]),
        "hsbp_bp1_condense": (cint, [vp, cint]),
        "hsbp_bp1_rhs": (cint, [vp, dbl, dp, dp, vp]),
        "hsbp_bp1_get_u": (cint, [vp, dp]),
        "hsbp_fault_create": (cint, [vp, i64, dp, dp, dp, vp, C.POINTER(vp)]),
        "hsbp_fault_destroy": (cint, [vp]),
        "hsbp_fault_rhs": (cint, [vp, dbl, dp, dp, vp]),
        "hsbp_fault_stage": (cint, [vp, dp, dp, dp, vp]),
        "hsbp_peak_fp64_fma": (cint, [vp, C.POINTER(dbl)]),
        "hsbp_peak_fp64_dmma": (cint, [vp, C.POINTER(dbl)]),
        "hsbp_peak_dgemm": (cint, [vp, i64, C.POINTER(dbl)]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = args
    L._signatures = sig
    _lib = L
    return L


class LocalStats(C.Structure):
    _fields_ = [("iterations_max", C.c_int64), ("iterations_sum", C.c_int64),
                ("failed_blocks", C.c_int64), ("max_rel_residual", C.c_double)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


class TraceStats(C.Structure):
    _fields_ = [("outer_iterations", C.c_int64), ("converged", C.c_int64), ("rel_residual", C.c_double),
                ("inner_iterations_sum", C.c_int64), ("inner_iterations_max", C.c_int64),
                ("local_solves", C.c_int64), ("true_rel_residual", C.c_double), ("failed_local_blocks", C.c_int64),
                ("max_local_rel_residual", C.c_double), ("coarse_dofs", C.c_int64), ("issued_iterations", C.c_int64),
                ("b_norm", C.c_double), ("cg_loop_ms", C.c_double)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


class Bp1Params(C.Structure):
    _fields_ = [("Vp", C.c_double), ("mu_shear", C.c_double), ("sigma_n", C.c_double), ("eta", C.c_double),
                ("V0", C.c_double), ("tau_z0", C.c_double), ("Dc", C.c_double), ("f0", C.c_double), ("b", C.c_double),
                ("ftol", C.c_double), ("atolx", C.c_double), ("rtolx", C.c_double), ("maxiter", C.c_int64)]


class Bp1Stats(C.Structure):
    _fields_ = [("rejected", C.c_int64), ("failure_bits", C.c_int64), ("failed_nodes", C.c_int64),
                ("newton_iterations_max", C.c_int64), ("local_iterations", C.c_int64)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


def _i64(a):
    a = np.ascontiguousarray(a, dtype=np.int64)
    return a, a.ctypes.data_as(C.POINTER(C.c_int64))


def _f64(a):
    a = np.ascontiguousarray(a, dtype=np.float64)
    return a, C.c_void_p(a.ctypes.data)


class DeviceArray:
    """A device buffer of float64 owned through hsbp_malloc / hsbp_free."""

    def __init__(self, ctx, n):
        self.ctx = ctx
        self.n = int(n)
        p = C.c_void_p()
        ctx._check(lib().hsbp_malloc(ctx.h, self.n * 8, C.byref(p)))
        self.ptr = p

    @classmethod
    def from_host(cls, ctx, a):
        a, pa = _f64(a)
        d = cls(ctx, a.size)
        ctx._check(lib().hsbp_h2d(ctx.h, d.ptr, pa, a.size * 8))
        return d

    def set(self, a):
        a, pa = _f64(a)
        assert a.size == self.n
        self.ctx._check(lib().hsbp_h2d(self.ctx.h, self.ptr, pa, a.size * 8))

    def zero(self):
        self.ctx._check(lib().hsbp_memset0(self.ctx.h, self.ptr, self.n * 8))

    def get(self):
        out = np.empty(self.n, dtype=np.float64)
        self.ctx._check(lib().hsbp_d2h(self.ctx.h, C.c_void_p(out.ctypes.data), self.ptr, self.n * 8))
        return out

    def free(self):
        if self.ptr is not None and self.ctx.h is not None:
            lib().hsbp_free(self.ctx.h, self.ptr)
        self.ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class Context:
    def __init__(self, device=0):
        h = C.c_void_p()
        rc = lib().hsbp_ctx_create(int(device), C.byref(h))
        if rc != 0:
            raise HsbpError(rc, "hsbp_ctx_create failed (a B200 / sm_100 GPU is required; no CPU fallback)")
        self.h = h
        self.device = int(device)
        # objects that hold handles created on this context (Blocks); they are destroyed first: the C objects keep raw
        # pointers to their parents (hsbp_trace -> hsbp_blocks -> hsbp_ctx), and Python's finalisation order is arbitrary
        self._children = weakref.WeakSet()

    def _check(self, rc):
        if rc != 0:
            raise HsbpError(rc, lib().hsbp_last_error(self.h).decode())

    def sync(self):
        self._check(lib().hsbp_sync(self.h))

    def array(self, a):
        return DeviceArray.from_host(self, a)

    def empty(self, n):
        return DeviceArray(self, n)

    def host_register(self, a):
        self._check(lib().hsbp_host_register(self.h, C.c_void_p(a.ctypes.data), a.nbytes))

    def host_unregister(self, a):
        self._check(lib().hsbp_host_unregister(self.h, C.c_void_p(a.ctypes.data)))

    def timer_start(self):
        self._check(lib().hsbp_timer_start(self.h))

    def timer_stop(self):
        ms = C.c_double()
        self._check(lib().hsbp_timer_stop(self.h, C.byref(ms)))
        return ms.value

    def stream(self):
        return lib().hsbp_stream(self.h)

    # -- multi-GPU: one context = one NCCL rank (hsbp_comm_*) ------------------------------------------
    @staticmethod
    def comm_unique_id():
        """128 bytes made by one rank; hand them to every other rank (any transport) and call comm_init everywhere"""
        buf = C.create_string_buffer(128)
        rc = lib().hsbp_comm_unique_id(buf)
        if rc != 0:
            raise HsbpError(rc, "hsbp_comm_unique_id failed (NCCL not loadable?)")
        return buf.raw

    def comm_init(self, unique_id, rank, world):
        assert len(unique_id) == 128
        self._check(lib().hsbp_comm_init(self.h, C.c_char_p(bytes(unique_id)), int(rank), int(world)))

    def comm_init_torch(self, dist):
        """communicator over the ranks of an initialised torch.distributed process group (the id travels through it)"""
        rank, world = dist.get_rank(), dist.get_world_size()
        box = [self.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        self.comm_init(box[0], rank, world)

    @property
    def rank(self):
        return lib().hsbp_comm_rank(self.h)

    @property
    def world(self):
        return lib().hsbp_comm_world(self.h)

    def fp64_peaks(self, dgemm_n=8192):
        """measured fp64 denominators of this device in TFLOP/s: CUDA-core FMA, mma.sync f64, cuBLAS DGEMM"""
        out = {}
        for name, fn, args in (("fma", lib().hsbp_peak_fp64_fma, ()), ("dmma", lib().hsbp_peak_fp64_dmma, ()),
                               ("dgemm", lib().hsbp_peak_dgemm, (int(dgemm_n),))):
            v = C.c_double()
            self._check(fn(self.h, *args, C.byref(v)))
            out[name] = v.value
        return out

    def allreduce_sum(self, x: "DeviceArray"):
        self._check(lib().hsbp_comm_allreduce_sum(self.h, x.ptr, x.n))

    def close(self):
        if self.h is not None:
            for child in list(self._children):
                child.close()
            lib().hsbp_ctx_destroy(self.h)
            self.h = None
