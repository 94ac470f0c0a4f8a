"""Device-side reduced KKT solver: the B200 counterpart of the reference's `misc.kkt_chol2`
(reference src/python/misc.py:1352-1567) behind the reference's own KKT plug-in API
(`kktsolver(W) -> f`, `f(x, y, z)` in place; src/python/coneprog.py:323-345, doc/source/coneprog.rst:1288-1332).

    from kvxopt import solvers
    from kvxopt_b200 import kkt
    factor = kkt.chol2(G, {'l': m, 'q': [], 's': []}, A)          # same call as misc.kkt_chol2(G, dims, A)
    sol = solvers.coneqp(P, q, G, h, kktsolver=lambda W: factor(W, P))
    sol = solvers.conelp(c, G, h, kktsolver=factor)                # LP: H is None
    # or in one line:  sol = solvers.qp(P, q, G, h, kktsolver=kkt.qp_kktsolver(P, G))

Where misc.kkt_chol2 assembles S = H + G' W^-1 W^-T G on the host every interior-point iteration (sparse gemm + syrk +
add), uploads it, and goes through five cholmod.solve calls per KKT solve, this solver keeps G, A and the pattern of S
on the device: `factor` uploads the m scaling entries W['di'] (and H's values), `solve` uploads and downloads x, y, z
once.  All floating-point work runs in libb200sparse.so (csrc/kkt_gpu.cu); there is no CPU fallback.

Like the reference function it handles problems with componentwise inequalities only (dims['q'] and dims['s'] empty);
nonlinear constraints (mnl > 0, used by solvers.cp) are not supported.  Error behaviour follows the reference: a
singular S at the FIRST factorization switches to S + A'A for the rest of the solve (misc.py:1427-1447); later
failures raise ArithmeticError, which the interior-point drivers report as a singular KKT matrix.
"""
import ctypes as C

import numpy as np

from . import _lib as L
from .cholmod import _ccs, _dense_view, _is_kvx, _is_dense, _size, _values

fn = L.fn

__all__ = ["chol2", "qp_kktsolver", "lp_kktsolver"]


class _Handle:
    def __init__(self, h):
        self.h = h

    def __del__(self):
        if getattr(self, "h", None):
            fn["b200s_kkt_free"](self.h)
            self.h = None


def _sparse_ccs(M, what):
    """CCS arrays of a dense or sparse kvxopt / numpy / scipy matrix"""
    if _is_dense(M):
        if _is_kvx(M):
            a = np.asarray(memoryview(M)).reshape(tuple(M.size), order="F") if M.size[0] * M.size[1] else np.zeros(tuple(M.size))
        else:
            a = np.asarray(M, dtype=np.float64)
        import scipy.sparse as sp
        M = sp.csc_matrix(a)
    cp, ri, vx = _ccs(M)
    return cp, ri, vx


def _vec(v, n, what):
    flat, nr, nc = _dense_view(v)
    if nr * nc != n:
        raise TypeError("%s must be a dense 'd' matrix with %d entries" % (what, n))
    return flat


def _raise(st):
    if st == L.NOT_POSDEF:
        raise ArithmeticError("KKT matrix is not positive definite")
    if st == L.NO_DEVICE:
        raise RuntimeError("kvxopt_b200.kkt needs a CUDA device: " + L.last_error())
    if st == L.OUT_OF_MEMORY:
        raise MemoryError(L.last_error())
    raise ValueError("KKT solver failed: %s (%s)" % (L.strerror(st), L.last_error()))


def chol2(G, dims, A, mnl=0):
    """Same contract as misc.kkt_chol2(G, dims, A, mnl): returns factor(W, H=None, Df=None), which returns
    solve(x, y, z)."""
    if dims["q"] or dims["s"]:
        raise ValueError("kktsolver option 'kkt_chol2' is implemented only for problems with no second-order or "
                         "semidefinite cone constraints")
    if mnl:
        raise ValueError("kvxopt_b200.kkt.chol2 does not support nonlinear constraints (mnl > 0)")
    p, n = _size(A)
    ml = dims["l"]
    if _size(G) != (ml, n):
        raise TypeError("G must be a %d x %d matrix" % (ml, n))
    Gc = _sparse_ccs(G, "G")
    Ac = _sparse_ccs(A, "A") if p > 0 else (None, None, None)
    state = {"handle": None, "firstcall": True, "singular": False, "Hkey": None}

    def _create(H):
        Hc = None
        if H is not None:
            if _size(H) != (n, n):
                raise TypeError("H must be a %d x %d matrix" % (n, n))
            Hc = _sparse_ccs(H, "H")
        h = C.c_void_p()
        st = fn["b200s_kkt_create"](n, ml, p, L.ptr_i64(Gc[0]), L.ptr_i64(Gc[1]), L.ptr_f64(Gc[2]),
                                    L.ptr_i64(Ac[0]), L.ptr_i64(Ac[1]), L.ptr_f64(Ac[2]),
                                    L.ptr_i64(Hc[0]) if Hc else None, L.ptr_i64(Hc[1]) if Hc else None, C.byref(h))
        if st != L.OK:
            _raise(st)
        state["handle"] = _Handle(h)
        state["Hpattern"] = (Hc[0].copy(), Hc[1].copy()) if Hc else None

    def factor(W, H=None, Df=None):
        if Df is not None:
            raise ValueError("kvxopt_b200.kkt.chol2 does not support nonlinear constraints")
        Hx = None
        if state["handle"] is None:
            _create(H)
        if H is not None:
            if state["Hpattern"] is None:
                raise ValueError("H was None in the first call and cannot appear later")
            # the pattern of H was fixed by the first call (as the reference's S += H on a fixed F['S'] assumes);
            # only the values travel -- no index conversion per interior-point iteration
            Hx = _values(H) if not _is_dense(H) else _sparse_ccs(H, "H")[2]
            if len(Hx) != len(state["Hpattern"][1]):
                raise ValueError("the sparsity pattern of H changed between calls")
        elif state["Hpattern"] is not None:
            raise ValueError("H was given in the first call and is missing now")
        di = _vec(W["di"], ml, "W['di']") if ml else np.zeros(0)
        di = np.ascontiguousarray(di, dtype=np.float64)
        handle = state["handle"]          # kept alive by the returned solve() even when the caller drops factor()
        h = handle.h
        minor = C.c_int64(0)
        st = fn["b200s_kkt_factor"](h, L.ptr_f64(di), L.ptr_f64(Hx) if Hx is not None else None, C.byref(minor))
        if st == L.NOT_POSDEF and state["firstcall"] and not state["singular"]:
            # misc.py:1427-1447: S singular in the first call => S + A'A from now on
            state["singular"] = True
            st = fn["b200s_kkt_set_singular"](h, 1)
            if st != L.OK:
                _raise(st)
            st = fn["b200s_kkt_factor"](h, L.ptr_f64(di), L.ptr_f64(Hx) if Hx is not None else None, C.byref(minor))
        state["firstcall"] = False
        if st != L.OK:
            _raise(st)

        def solve(x, y, z):
            xf = _vec(x, n, "x")
            yf = _vec(y, p, "y") if p else None
            zf = _vec(z, ml, "z") if ml else None
            st2 = fn["b200s_kkt_solve"](handle.h, L.ptr_f64(xf), L.ptr_f64(yf) if p else None, L.ptr_f64(zf) if ml else None)
            if st2 != L.OK:
                _raise(st2)

        return solve

    factor.info = lambda: info(state["handle"].h) if state["handle"] else {}
    return factor


def info(h):
    inf = L.KktInfo()
    fn["b200s_kkt_info"](h, C.byref(inf))
    return {k: getattr(inf, k) for k, _ in inf._fields_ if k not in ("Sp", "Si")}


def _dims_l(G):
    return {"l": _size(G)[0], "q": [], "s": []}


def _empty_A(n, like):
    if _is_kvx(like):
        import importlib
        kv = importlib.import_module(type(like).__module__.split(".")[0])
        return kv.spmatrix([], [], [], (0, n), "d")
    import scipy.sparse as sp
    return sp.csc_matrix((0, n))


def qp_kktsolver(P, G, A=None):
    """kktsolver callable for solvers.qp / solvers.coneqp with componentwise inequalities G x <= h:
    solvers.qp(P, q, G, h, A, b, kktsolver=kkt.qp_kktsolver(P, G, A))"""
    n = _size(P)[0]
    factor = chol2(G, _dims_l(G), A if A is not None else _empty_A(n, G))
    kktsolver = lambda W: factor(W, P)
    kktsolver.info = factor.info
    return kktsolver


def lp_kktsolver(G, A=None):
    """kktsolver callable for solvers.lp / solvers.conelp: solvers.lp(c, G, h, A, b, kktsolver=kkt.lp_kktsolver(G, A))"""
    n = _size(G)[1]
    factor = chol2(G, _dims_l(G), A if A is not None else _empty_A(n, G))
    kktsolver = lambda W: factor(W)
    kktsolver.info = factor.info
    return kktsolver
