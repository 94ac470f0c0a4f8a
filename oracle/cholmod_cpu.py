"""TEST INFRASTRUCTURE: a CPU `kvxopt.cholmod` look-alike backed by the oracle (oracle/chol_oracle.c), used only as the
CPU arm when the reference IPM (probe build oracle/_ref) is timed / compared on BASELINE configs 3 and 5.  It covers
the calls misc.kkt_chol2 makes (symbolic, numeric, solve, spsolve, options).  Never imported by the product."""
import ctypes as C
import sys

import numpy as np

from . import CholOracle

options = {}


class _Factor:
    def __init__(self, oracle, kind):
        self.o = oracle
        self.numeric = False
        self.kind = kind


def _ccs(A):
    cp, ri, vx = A.CCS
    return (np.array(cp, dtype=np.int64).reshape(-1), np.array(ri, dtype=np.int64).reshape(-1),
            np.array(vx, dtype=np.float64).reshape(-1))


def symbolic(A, p=None, uplo="L"):
    """fill-reducing permutation from the engine's HOST analysis (integer work, no GPU), numeric work in the oracle"""
    from kvxopt_b200 import _lib as L
    n = A.size[0]
    cp, ri, _ = _ccs(A)
    perm = None
    if n > 0:
        h = L.vp()
        pp = np.array(p, dtype=np.int64).reshape(-1) if p is not None else None
        st = L.fn["b200s_chol_analyze"](n, L.ptr_i64(cp), L.ptr_i64(ri), uplo.encode(), L.ptr_i64(pp), None, C.byref(h))
        if st != 0:
            raise ValueError("symbolic factorization failed")
        perm = np.zeros(n, dtype=np.int64)
        L.fn["b200s_chol_get_perm"](h, L.ptr_i64(perm))
        L.fn["b200s_chol_free"](h)
    return _Factor(CholOracle(n, cp, ri, uplo, perm), type(A))


def numeric(A, F):
    _, _, vx = _ccs(A)
    F.o.factorize(vx)
    F.numeric = True


def solve(F, B, sys=0, nrhs=-1, ldB=0, offsetB=0):
    n = F.o.n
    if n == 0 or len(B) == 0:
        return
    a = np.asarray(memoryview(B))
    X = F.o.solve(np.array(a, order="F"), sys)
    a[...] = X.reshape(a.shape, order="F")


def spsolve(F, B, sys=0):
    import kvxopt
    n, k = B.size
    if n == 0 or k == 0:
        return kvxopt.spmatrix([], [], [], (n, k), "d")
    D = np.array(kvxopt.matrix(B), order="F")
    X = F.o.solve(D, sys)
    return kvxopt.sparse(kvxopt.matrix(X))


def install(kvxopt_module):
    sys.modules[kvxopt_module.__name__ + ".cholmod"] = sys.modules[__name__]
    setattr(kvxopt_module, "cholmod", sys.modules[__name__])
