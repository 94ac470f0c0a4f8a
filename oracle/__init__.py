"""CPU oracle (TEST INFRASTRUCTURE): ctypes access to oracle/liboracle.so.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "liboracle.so")
_lib = None
i64 = C.c_int64
pi = C.POINTER(C.c_int64)
pd = C.POINTER(C.c_double)


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB):
            from . import build as _b
            _b.build()
        _lib = C.CDLL(_LIB)
        L = _lib
        L.oracle_chol_analyze.restype = C.c_void_p
        L.oracle_chol_analyze.argtypes = [i64, pi, pi, C.c_char, pi]
        L.oracle_chol_free.argtypes = [C.c_void_p]
        L.oracle_chol_factorize.argtypes = [C.c_void_p, pd]
        L.oracle_chol_solve.argtypes = [C.c_void_p, C.c_int, pd, i64, i64]
        L.oracle_chol_diag.argtypes = [C.c_void_p, pd]
        L.oracle_chol_get_perm.argtypes = [C.c_void_p, pi]
        L.oracle_chol_dense_L.argtypes = [C.c_void_p, pd]
        for f in ("oracle_chol_nnzL", "oracle_chol_nsuper", "oracle_chol_minor"):
            getattr(L, f).restype = i64
            getattr(L, f).argtypes = [C.c_void_p]
        L.oracle_chol_flops.restype = C.c_double
        L.oracle_chol_flops.argtypes = [C.c_void_p]
        L.oracle_blas_threads.argtypes = [C.c_int]
    return _lib


def _p(a, t):
    return a.ctypes.data_as(t)


class CholOracle:
    """Supernodal left-looking LL^T of the `uplo` triangle of a CCS matrix (colptr,rowind,values)."""

    def __init__(self, n, colptr, rowind, uplo="L", perm=None):
        self.n = int(n)
        self.colptr = np.ascontiguousarray(colptr, dtype=np.int64)
        self.rowind = np.ascontiguousarray(rowind, dtype=np.int64)
        p = np.ascontiguousarray(perm, dtype=np.int64) if perm is not None else None
        self.h = lib().oracle_chol_analyze(self.n, _p(self.colptr, pi), _p(self.rowind, pi), uplo.encode(),
                                           _p(p, pi) if p is not None else None)
        if not self.h:
            raise ValueError("oracle: invalid permutation")

    def __del__(self):
        if getattr(self, "h", None):
            lib().oracle_chol_free(self.h)
            self.h = None

    def factorize(self, values):
        v = np.ascontiguousarray(values, dtype=np.float64)
        st = lib().oracle_chol_factorize(self.h, _p(v, pd))
        if st:
            raise ArithmeticError(int(lib().oracle_chol_minor(self.h)))

    def solve(self, B, sys=0):
        X = np.array(B, dtype=np.float64, order="F", copy=True)
        if self.n == 0 or X.size == 0:
            return X
        X2 = X.reshape(self.n, -1, order="F")
        st = lib().oracle_chol_solve(self.h, sys, _p(X2, pd), X2.shape[1], max(self.n, 1))
        if st:
            raise ValueError("oracle solve failed: %d" % st)
        return X

    def diag(self):
        d = np.zeros(self.n)
        lib().oracle_chol_diag(self.h, _p(d, pd))
        return d

    def perm(self):
        p = np.zeros(self.n, dtype=np.int64)
        lib().oracle_chol_get_perm(self.h, _p(p, pi))
        return p

    def dense_L(self):
        L = np.zeros((self.n, self.n), order="F")
        lib().oracle_chol_dense_L(self.h, _p(L, pd))
        return L

    @property
    def nnzL(self):
        return int(lib().oracle_chol_nnzL(self.h))

    @property
    def flops(self):
        return float(lib().oracle_chol_flops(self.h))


def _klu_sigs():
    L = lib()
    if getattr(L, "_klu_ready", False):
        return L
    L.oracle_klu_factor.restype = C.c_void_p
    L.oracle_klu_factor.argtypes = [i64, pi, pi, pd, pi, pi, C.c_double, pi]
    L.oracle_klu_free.argtypes = [C.c_void_p]
    L.oracle_klu_refactor.argtypes = [C.c_void_p, pi, pi, pd]
    L.oracle_klu_solve.argtypes = [C.c_void_p, C.c_int, pd, i64, i64]
    L.oracle_klu_det.restype = C.c_double
    L.oracle_klu_det.argtypes = [C.c_void_p]
    L.oracle_klu_nnz.restype = i64
    L.oracle_klu_nnz.argtypes = [C.c_void_p, C.c_int]
    L.oracle_klu_flops.restype = C.c_double
    L.oracle_klu_flops.argtypes = [C.c_void_p]
    L.oracle_klu_pnum.argtypes = [C.c_void_p, pi]
    L._klu_ready = True
    return L


class KluOracle:
    """Gilbert-Peierls LU with KLU's scaling and pivot rule on a CCS matrix; optional row/column pre-ordering."""

    def __init__(self, n, colptr, rowind, values, P0=None, Q=None, tol=1e-3):
        L = _klu_sigs()
        self.n = int(n)
        self.colptr = np.ascontiguousarray(colptr, dtype=np.int64)
        self.rowind = np.ascontiguousarray(rowind, dtype=np.int64)
        v = np.ascontiguousarray(values, dtype=np.float64)
        p0 = np.ascontiguousarray(P0, dtype=np.int64) if P0 is not None else None
        q = np.ascontiguousarray(Q, dtype=np.int64) if Q is not None else None
        sing = i64(-1)
        self.h = L.oracle_klu_factor(self.n, _p(self.colptr, pi), _p(self.rowind, pi), _p(v, pd),
                                     _p(p0, pi) if p0 is not None else None, _p(q, pi) if q is not None else None,
                                     tol, C.byref(sing))
        if not self.h:
            raise ArithmeticError("singular matrix")

    def __del__(self):
        if getattr(self, "h", None):
            lib().oracle_klu_free(self.h)
            self.h = None

    def refactor(self, values):
        v = np.ascontiguousarray(values, dtype=np.float64)
        st = lib().oracle_klu_refactor(self.h, _p(self.colptr, pi), _p(self.rowind, pi), _p(v, pd))
        if st:
            raise ArithmeticError("singular matrix")

    def solve(self, B, trans="N"):
        X = np.array(B, dtype=np.float64, order="F", copy=True)
        if self.n == 0 or X.size == 0:
            return X
        X2 = X.reshape(self.n, -1, order="F")
        lib().oracle_klu_solve(self.h, 0 if trans == "N" else 1, _p(X2, pd), X2.shape[1], max(self.n, 1))
        return X

    def det(self):
        return float(lib().oracle_klu_det(self.h))

    @property
    def pnum(self):
        """the pivot sequence chosen by the oracle's own threshold pivoting: pnum[k] = original row of pivotal row k"""
        out = np.zeros(max(self.n, 1), dtype=np.int64)
        lib().oracle_klu_pnum(self.h, _p(out, pi))
        return out[:self.n]

    @property
    def nnz_L(self):
        return int(lib().oracle_klu_nnz(self.h, 0))

    @property
    def nnz_U(self):
        return int(lib().oracle_klu_nnz(self.h, 1))

    @property
    def flops(self):
        return float(lib().oracle_klu_flops(self.h))
