"""ORACLE -- build the C pieces of the CPU baseline with gcc (test/bench infrastructure only)."""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_build")
LIB = os.path.join(OUT, "libspmv.so")


def build(force=False):
    src = os.path.join(HERE, "spmv_csr.c")
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= os.path.getmtime(src):
        return LIB
    os.makedirs(OUT, exist_ok=True)
    subprocess.run(["gcc", "-O3", "-march=native", "-fopenmp", "-shared", "-fPIC", "-o", LIB, src], check=True)
    return LIB


class BlockSpMV:
    """CSR product over a list of independent square blocks (scipy matrices) through spmv_csr.c."""

    def __init__(self, mats):
        L = ctypes.CDLL(build())
        self._f = L.hsbp_oracle_spmv_blocks
        self._f.restype = None
        self.max_threads = L.hsbp_oracle_max_threads()
        mats = [m.tocsr() for m in mats]
        for m in mats:
            m.sort_indices()
        self.nb = len(mats)
        self.rows = np.array([m.shape[0] for m in mats], dtype=np.int64)
        self.row_off = np.concatenate([[0], np.cumsum(self.rows)[:-1]]).astype(np.int64)
        nnz = np.array([m.nnz for m in mats], dtype=np.int64)
        self.nnz_off = np.concatenate([[0], np.cumsum(nnz)[:-1]]).astype(np.int64)
        self.rp_off = np.concatenate([[0], np.cumsum(self.rows + 1)[:-1]]).astype(np.int64)
        self.rowptr = np.concatenate([m.indptr.astype(np.int64) for m in mats])
        self.col = np.concatenate([m.indices.astype(np.int32) for m in mats])
        self.val = np.concatenate([m.data.astype(np.float64) for m in mats])
        self.n = int(self.rows.sum())
        self.nnz = int(nnz.sum())

    def __call__(self, x, y=None, nthreads=0):
        x = np.ascontiguousarray(x, dtype=np.float64)
        if y is None:
            y = np.empty(self.n)
        p = lambda a: a.ctypes.data_as(ctypes.c_void_p)
        self._f(ctypes.c_int64(self.nb), p(self.rp_off), p(self.nnz_off), p(self.row_off), p(self.rows),
                p(self.rowptr), p(self.col), p(self.val), p(x), p(y), ctypes.c_int(nthreads))
        return y
