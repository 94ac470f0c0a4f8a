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

__all__ = ["chol", "chol2", "ldl", "ldl2", "qp_kktsolver", "lp_kktsolver"]


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


def _checked_values(H, pattern):
    """values of H after checking that it still has the pattern (colptr, rowind) fixed by the first call: the term lists
    and the scatter maps built then are only valid for that pattern (a same-length, different pattern would be scattered
    through the old map and factor silently wrong)"""
    cp, ri, vx = _sparse_ccs(H, "H")
    if cp.shape != pattern[0].shape or ri.shape != pattern[1].shape or not (np.array_equal(cp, pattern[0]) and
                                                                            np.array_equal(ri, pattern[1])):
        raise ValueError("the sparsity pattern of H changed between calls")
    return vx


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
            Hx = _checked_values(H, state["Hpattern"])
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


def _dense_colmajor(M, shape, what):
    """dense column-major float64 copy of a dense or sparse kvxopt / numpy / scipy matrix"""
    if _is_dense(M):
        if _is_kvx(M):
            a = np.asarray(memoryview(M)).reshape(tuple(M.size), order="F") if M.size[0] * M.size[1] else np.zeros(tuple(M.size))
        else:
            a = np.asarray(M, dtype=np.float64)
            if a.ndim == 1:
                a = a.reshape(-1, 1)
    else:
        import scipy.sparse as sp
        cp, ri, vx = _ccs(M)
        a = sp.csc_matrix((vx, ri, cp), shape=_size(M)).toarray()
    if tuple(a.shape) != tuple(shape):
        raise TypeError("%s must be a %d x %d matrix" % ((what,) + tuple(shape)))
    return np.asfortranarray(a, dtype=np.float64)


class _DenseHandle:
    def __init__(self, h):
        self.h = h

    def __del__(self):
        if getattr(self, "h", None):
            fn["b200s_kktd_free"](self.h)
            self.h = None


def chol(G, dims, A, mnl=0):
    """Same contract as misc.kkt_chol(G, dims, A, mnl) (reference src/python/misc.py:1213-1349), the 'chol' kktsolver:
    QR of A' once, then per interior-point iteration the dense K = [Q1 Q2]'(H + G' W^-1 W^-T G)[Q1 Q2] and the dense
    Cholesky factorization of its (2,2) block of order n-p -- here on the device (csrc/kktd_gpu.cu: register-tiled FP64
    GEMM for the SYRK and the compact-WY products, the multifrontal engine's one-supernode case for the Cholesky
    factorization and its solves).  Returns factor(W, H=None, Df=None) -> solve(x, y, z), in place: x, y, z := ux, uy, W uz.
    Componentwise inequalities only (dims['q'], dims['s'] empty, mnl = 0), like kkt.chol2.  No CPU fallback."""
    if dims["q"] or dims["s"]:
        raise ValueError("kvxopt_b200.kkt.chol is implemented only for problems with no second-order or semidefinite "
                         "cone constraints")
    if mnl:
        raise ValueError("kvxopt_b200.kkt.chol does not support nonlinear constraints (mnl > 0)")
    p, n = _size(A)
    ml = dims["l"]
    Gd = _dense_colmajor(G, (ml, n), "G")
    Ad = _dense_colmajor(A, (p, n), "A") if p else np.zeros((0, n), order="F")
    h = C.c_void_p()
    st = fn["b200s_kktd_create"](n, ml, p, L.ptr_f64(Gd.reshape(-1, order="F")), L.ptr_f64(Ad.reshape(-1, order="F")) if p else None,
                                 C.byref(h))
    if st == L.SINGULAR:
        raise ArithmeticError("Rank(A) < p")
    if st != L.OK:
        _raise(st)
    handle = _DenseHandle(h)

    def factor(W, H=None, Df=None):
        if Df is not None:
            raise ValueError("kvxopt_b200.kkt.chol does not support nonlinear constraints")
        di = np.ascontiguousarray(_vec(W["di"], ml, "W['di']") if ml else np.zeros(0), dtype=np.float64)
        Hd = _dense_colmajor(H, (n, n), "H") if H is not None else None
        minor = C.c_int64(0)
        st = fn["b200s_kktd_factor"](handle.h, L.ptr_f64(di), L.ptr_f64(Hd.reshape(-1, order="F")) if Hd is not None else None,
                                     C.byref(minor))
        if st == L.NOT_POSDEF:
            raise ArithmeticError(int(minor.value) + 1)       # lapack.potrf's info (1-based), as the reference raises it
        if st != L.OK:
            _raise(st)

        def solve(x, y, z):
            xf = _vec(x, n, "x")
            yf = _vec(y, p, "y") if p else None
            zf = _vec(z, ml, "z") if ml else None
            st2 = fn["b200s_kktd_solve"](handle.h, L.ptr_f64(xf), L.ptr_f64(yf) if p else None, L.ptr_f64(zf) if ml else None)
            if st2 != L.OK:
                _raise(st2)

        return solve

    def _info():
        inf = L.KktdInfo()
        fn["b200s_kktd_info"](handle.h, C.byref(inf))
        return inf.asdict()

    factor.info = _info
    factor._handle = handle
    return factor


class _SpmvHandle:
    def __init__(self, h):
        self.h = h

    def __del__(self):
        if getattr(self, "h", None):
            fn["b200s_spmv_free"](self.h)
            self.h = None


class _CholHandle:
    def __init__(self, h):
        self.h = h

    def __del__(self):
        if getattr(self, "h", None):
            fn["b200s_chol_free"](self.h)
            self.h = None


def ldl(G, dims, A, mnl=0, kktreg=None):
    """Sparse counterpart of misc.kkt_ldl(G, dims, A, mnl, kktreg) (reference src/python/misc.py:1055-1130): the
    3 x 3 system

        [ H           A'   G' W^-1 ]   [ ux   ]   [ bx      ]
        [ A           0    0       ] * [ uy   ] = [ by      ]
        [ W^-T G      0   -I       ]   [ W uz ]   [ W^-T bz ]

    is factored as ONE sparse symmetric indefinite matrix, P K P' = L D L' without pivoting, by the supernodal engine in
    its signed mode (the C ABI's `supernodal = 0`: what cholmod.options['supernodal'] = 0 selects), where the reference
    forms K densely and calls LAPACK sytrf.  Returns factor(W, H=None, Df=None) -> solve(x, y, z) with the reference's
    in-place contract (x, y, z := ux, uy, W uz).

    Elimination order.  Without `kktreg` the (2,2) block is zero and, for linear programs, so is H: the order is
    z first (pivots -1), then x by approximate minimum degree on the pattern of S = H + G'G (pivots = those of the
    Cholesky factorization of H + G' W^-2 G), then y (pivots of -A S^-1 A'): every pivot is nonzero whenever the
    reduced system of misc.kkt_chol2 is positive definite.  With `kktreg` (coneprog.py:430-434) K is quasi-definite, any
    symmetric permutation has an LDL' factorization, and the fill-reducing ordering is applied to all of K.

    Componentwise inequalities only (dims['q'], dims['s'] empty, mnl = 0), as kkt.chol2.  The pattern of K is analysed
    once; each factor() sends only K's values; solve() is one host-buffer call.  No CPU fallback."""
    import scipy.sparse as sp
    if dims["q"] or dims["s"]:
        raise ValueError("kvxopt_b200.kkt.ldl is implemented only for problems with no second-order or semidefinite "
                         "cone constraints")
    if mnl:
        raise ValueError("kvxopt_b200.kkt.ldl does not support nonlinear constraints (mnl > 0)")
    p, n = _size(A)
    m = dims["l"]
    if _size(G) != (m, n):
        raise TypeError("G must be a %d x %d matrix" % (m, n))
    Gp, Gi, Gx = _sparse_ccs(G, "G")
    if p > 0:
        Ap, Ai, Ax = _sparse_ccs(A, "A")
    else:
        Ap, Ai, Ax = np.zeros(n + 1, dtype=np.int64), np.zeros(0, dtype=np.int64), np.zeros(0)
    N = n + p + m
    reg = float(kktreg) if kktreg else 0.0
    state = {"handle": None, "Hnnz": None}

    def _create(H):
        if H is not None:
            if _size(H) != (n, n):
                raise TypeError("H must be a %d x %d matrix" % (n, n))
            Hp, Hi, Hx = _sparse_ccs(H, "H")
            state["Hpat"] = (Hp.copy(), Hi.copy())
            Hc = np.repeat(np.arange(n, dtype=np.int64), np.diff(Hp))
            keep = np.nonzero(Hi >= Hc)[0]                       # lower triangle, as sytrf reads it
            Hi, Hc = Hi[keep], Hc[keep]
            # every diagonal entry of the x block has a slot, stored or not (kktreg lands there)
            have = np.zeros(n, dtype=bool); have[Hi[Hi == Hc]] = True
        else:
            keep = np.zeros(0, dtype=np.int64); Hi = Hc = keep; have = np.zeros(n, dtype=bool)
        xdiag = np.nonzero(~have)[0]
        Gc = np.repeat(np.arange(n, dtype=np.int64), np.diff(Gp))
        Ac = np.repeat(np.arange(n, dtype=np.int64), np.diff(Ap))
        rows = np.concatenate([Hi, xdiag, n + Ai, n + p + Gi, n + np.arange(p), n + p + np.arange(m)])
        cols = np.concatenate([Hc, xdiag, Ac, Gc, n + np.arange(p), n + p + np.arange(m)])
        K = sp.csc_matrix((np.arange(1, len(rows) + 1, dtype=np.float64), (rows, cols)), shape=(N, N))
        K.sort_indices()
        if K.nnz != len(rows):
            raise ValueError("H, A and G must not hold duplicate entries")
        src = K.data.astype(np.int64) - 1                        # CCS slot -> index into the concatenated value list
        kp = np.ascontiguousarray(K.indptr, dtype=np.int64)
        ki = np.ascontiguousarray(K.indices, dtype=np.int64)
        if reg:
            perm = None
        else:
            # fill-reducing order of the x block from the engine's host analysis of S = H + G'G
            Gm = sp.csc_matrix((np.ones(len(Gi)), Gi, Gp), shape=(m, n))
            Sp = (Gm.T @ Gm + sp.identity(n)).tocsc()
            if len(Hi):
                Hm = sp.csc_matrix((np.ones(len(Hi)), (Hi, Hc)), shape=(n, n))
                Sp = (Sp + Hm + Hm.T).tocsc()
            Sl = sp.tril(Sp).tocsc(); Sl.sort_indices()
            hs = C.c_void_p()
            st = fn["b200s_chol_analyze"](n, L.ptr_i64(np.ascontiguousarray(Sl.indptr, dtype=np.int64)),
                                          L.ptr_i64(np.ascontiguousarray(Sl.indices, dtype=np.int64)), b"L", None, None,
                                          C.byref(hs))
            if st != L.OK:
                _raise(st)
            px = np.zeros(n, dtype=np.int64)
            fn["b200s_chol_get_perm"](hs, L.ptr_i64(px))
            fn["b200s_chol_free"](hs)
            perm = np.ascontiguousarray(np.concatenate([n + p + np.arange(m), px, n + np.arange(p)]), dtype=np.int64)
        o = L.CholOpts()
        fn["b200s_chol_default_opts"](C.byref(o))
        o.supernodal = 0                                         # signed LDL' (no pivoting)
        h = C.c_void_p()
        st = fn["b200s_chol_analyze"](N, L.ptr_i64(kp), L.ptr_i64(ki), b"L", L.ptr_i64(perm) if perm is not None else None,
                                      C.byref(o), C.byref(h))
        if st != L.OK:
            _raise(st)
        state.update(handle=_CholHandle(h), Hnnz=(len(keep) if H is not None else None), keep=keep, src=src, kp=kp, ki=ki,
                     nxd=len(xdiag), Grow=Gi, vals=np.zeros(len(rows)), kv=np.zeros(len(rows)),
                     u=np.zeros(N), hdiag=(np.nonzero(Hi == Hc)[0] if H is not None else None))

    def factor(W, H=None, Df=None):
        if Df is not None:
            raise ValueError("kvxopt_b200.kkt.ldl does not support nonlinear constraints")
        if state["handle"] is None:
            _create(H)
        if (H is None) != (state["Hnnz"] is None):
            raise ValueError("H must be given in every call or in none")
        di = np.ascontiguousarray(_vec(W["di"], m, "W['di']") if m else np.zeros(0), dtype=np.float64)
        v = state["vals"]
        k0 = 0
        if H is not None:
            Hx = _checked_values(H, state["Hpat"])[state["keep"]]
            v[:len(Hx)] = Hx
            if reg:
                v[state["hdiag"]] += reg
            k0 = len(Hx)
        v[k0:k0 + state["nxd"]] = reg; k0 += state["nxd"]
        v[k0:k0 + len(Ax)] = Ax; k0 += len(Ax)
        v[k0:k0 + len(Gx)] = di[state["Grow"]] * Gx; k0 += len(Gx)         # W^-T G = diag(di) G
        v[k0:k0 + p] = -reg; k0 += p
        v[k0:k0 + m] = -1.0 - reg
        kv = state["kv"]
        np.take(v, state["src"], out=kv)
        handle = state["handle"]
        minor = C.c_int64(0)
        st = fn["b200s_chol_factorize"](handle.h, None, None, L.ptr_f64(kv), C.byref(minor))
        state["factors"] = state.get("factors", 0) + 1
        if st == L.NOT_POSDEF:
            raise ArithmeticError("zero pivot in the LDL' factorization of the KKT matrix (column %d)" % minor.value)
        if st != L.OK:
            _raise(st)
        u = state["u"]

        def solve(x, y, z):
            xf = _vec(x, n, "x")
            yf = _vec(y, p, "y") if p else None
            zf = _vec(z, m, "z") if m else None
            u[:n] = xf
            if p:
                u[n:n + p] = yf
            if m:
                u[n + p:] = di * zf
            st2 = fn["b200s_chol_solve"](handle.h, 0, L.ptr_f64(u), 1, N)
            if st2 != L.OK:
                _raise(st2)
            xf[:] = u[:n]
            if p:
                yf[:] = u[n:n + p]
            if m:
                zf[:] = u[n + p:]

        return solve

    def _info():
        if state["handle"] is None:
            return {}
        inf = L.CholInfo()
        fn["b200s_chol_info"](state["handle"].h, C.byref(inf))
        return {"order": N, "nnz_L": inf.nnz_L, "flops": inf.flops, "nsuper": inf.nsuper, "max_front_rows": inf.max_front_rows,
                "ms_analyze": inf.ms_analyze, "ms_factor_last": inf.ms_factor, "ms_solve_last": inf.ms_solve,
                "factorizations": state.get("factors", 0)}

    factor.info = _info
    factor._state = state          # assembled pattern / values, read by the host-side tests
    return factor


def ldl2(G, dims, A, mnl=0):
    """Sparse counterpart of misc.kkt_ldl2(G, dims, A, mnl) (reference src/python/misc.py:1128-1210): the 2 x 2 system

        [ H + G' W^-1 W^-T G   A' ]   [ ux ]   [ bx + G' W^-1 W^-T bz ]
        [ A                    0  ] * [ uy ] = [ by                   ],      W uz = W^-T (G ux - bz)

    factored as ONE sparse symmetric quasi-definite matrix, P K P' = L D L' without pivoting, by the engine's signed mode
    (x columns in approximate-minimum-degree order of the pattern of S = H + G'G, then y: the x pivots are those of the
    Cholesky factorization of S, the y pivots those of -A S^-1 A').  The reference forms K densely (sytrf, or potrf when
    there are no equalities).  Unlike kkt.chol2 there is no dense p x p Schur complement, so problems with many equality
    constraints stay sparse.  Componentwise inequalities only (dims['q'], dims['s'] empty, mnl = 0).  The pattern of K and the
    list of products g_ki g_kj behind every entry of G'D^2 G are built once; factor() sends di^2 and the values of H and A, the
    entries of K are evaluated on the device (b200s_spmv_apply: K's values as a fixed sparse linear map of those numbers) and
    factored where they are; solve() is two sparse products with G on the host around one host-buffer solve.  No CPU fallback."""
    import scipy.sparse as sp
    if dims["q"] or dims["s"]:
        raise ValueError("kvxopt_b200.kkt.ldl2 is implemented only for problems with no second-order or semidefinite "
                         "cone constraints")
    if mnl:
        raise ValueError("kvxopt_b200.kkt.ldl2 does not support nonlinear constraints (mnl > 0)")
    p, n = _size(A)
    m = dims["l"]
    if _size(G) != (m, n):
        raise TypeError("G must be a %d x %d matrix" % (m, n))
    Gp, Gi, Gx = _sparse_ccs(G, "G")
    Gm = sp.csc_matrix((Gx, Gi, Gp), shape=(m, n))
    Gr = Gm.tocsr()
    if p > 0:
        Ap, Ai, Ax = _sparse_ccs(A, "A")
    else:
        Ap, Ai, Ax = np.zeros(n + 1, dtype=np.int64), np.zeros(0, dtype=np.int64), np.zeros(0)
    N = n + p
    state = {"handle": None, "Hnnz": None}

    def _create(H):
        # products behind the lower triangle of G' D^2 G: for every row k of G and every pair i >= j of its columns
        rp, ci, gv = Gr.indptr, Gr.indices, Gr.data
        cnt = np.diff(rp)
        ti, tj, tk, tc = [], [], [], []
        for r in np.unique(cnt):                    # rows grouped by their number of entries: vectorised pair lists
            if r == 0:
                continue
            rows = np.nonzero(cnt == r)[0]
            base = rp[rows][:, None] + np.arange(r)[None, :]
            cols, vals = ci[base], gv[base]
            a, b = np.tril_indices(r)
            ii, jj = cols[:, a], cols[:, b]
            hi, lo = np.maximum(ii, jj), np.minimum(ii, jj)
            ti.append(hi.ravel()); tj.append(lo.ravel())
            tk.append(np.repeat(rows, len(a))); tc.append((vals[:, a] * vals[:, b]).ravel())
        ti = np.concatenate(ti) if ti else np.zeros(0, dtype=np.int64)
        tj = np.concatenate(tj) if tj else np.zeros(0, dtype=np.int64)
        tk = np.concatenate(tk) if tk else np.zeros(0, dtype=np.int64)
        tc = np.concatenate(tc) if tc else np.zeros(0)
        if H is not None:
            if _size(H) != (n, n):
                raise TypeError("H must be a %d x %d matrix" % (n, n))
            Hp, Hi, Hx = _sparse_ccs(H, "H")
            state["Hpat"] = (Hp.copy(), Hi.copy())
            Hc = np.repeat(np.arange(n, dtype=np.int64), np.diff(Hp))
            keep = np.nonzero(Hi >= Hc)[0]
            Hi, Hc = Hi[keep], Hc[keep]
        else:
            keep = np.zeros(0, dtype=np.int64); Hi = Hc = keep
        Ac = np.repeat(np.arange(n, dtype=np.int64), np.diff(Ap))
        # pattern of K (lower): S entries (terms and H, merged), diagonal of x, A rows, diagonal of y
        rows = np.concatenate([ti, Hi, np.arange(n), n + Ai, n + np.arange(p)]).astype(np.int64)
        cols = np.concatenate([tj, Hc, np.arange(n), Ac, n + np.arange(p)]).astype(np.int64)
        key = cols * N + rows
        uniq, slot = np.unique(key, return_inverse=True)           # CCS order: column-major, rows ascending
        kcol, krow = uniq // N, uniq % N
        kp = np.zeros(N + 1, dtype=np.int64)
        np.add.at(kp, kcol + 1, 1)
        kp = np.cumsum(kp)
        ki = np.ascontiguousarray(krow, dtype=np.int64)
        nt, nh = len(ti), len(Hi)
        Sl = sp.csc_matrix((np.ones(len(uniq)), ki, kp), shape=(N, N))[:n, :n].tocsc()
        Sl.sort_indices()
        hs = C.c_void_p()
        st = fn["b200s_chol_analyze"](n, L.ptr_i64(np.ascontiguousarray(Sl.indptr, dtype=np.int64)),
                                      L.ptr_i64(np.ascontiguousarray(Sl.indices, dtype=np.int64)), b"L", None, None, C.byref(hs))
        if st != L.OK:
            _raise(st)
        px = np.zeros(n, dtype=np.int64)
        fn["b200s_chol_get_perm"](hs, L.ptr_i64(px))
        fn["b200s_chol_free"](hs)
        perm = np.ascontiguousarray(np.concatenate([px, n + np.arange(p)]), dtype=np.int64)
        o = L.CholOpts()
        fn["b200s_chol_default_opts"](C.byref(o))
        o.supernodal = 0                                            # signed LDL' (no pivoting)
        h = C.c_void_p()
        st = fn["b200s_chol_analyze"](N, L.ptr_i64(kp), L.ptr_i64(ki), b"L", L.ptr_i64(perm), C.byref(o), C.byref(h))
        if st != L.OK:
            _raise(st)
        # the stored entries of K as a fixed linear map of w = [di^2; values of H (lower); values of A], evaluated on the device
        # (b200s_spmv_*): rows = entries of K in CCS order, terms in the order [G'D^2G products, H, A] within a row
        na = len(Ai)
        trow = np.concatenate([slot[:nt], slot[nt:nt + nh], slot[nt + nh + n:nt + nh + n + na]]).astype(np.int64)
        tcol = np.concatenate([tk, m + np.arange(nh), m + nh + np.arange(na)]).astype(np.int64)
        tval = np.concatenate([tc, np.ones(nh + na)])
        order = np.argsort(trow, kind="stable")
        mp = np.zeros(len(uniq) + 1, dtype=np.int64)
        np.add.at(mp, trow + 1, 1)
        mp = np.cumsum(mp)
        state.update(handle=_CholHandle(h), asm=None, Hnnz=(nh if H is not None else None), keep=keep, kp=kp, ki=ki,
                     asm_csr=(np.ascontiguousarray(mp), np.ascontiguousarray(tcol[order]), np.ascontiguousarray(tval[order])),
                     nk=len(uniq), w=np.zeros(m + nh + na), u=np.zeros(N))

    def factor(W, H=None, Df=None):
        if Df is not None:
            raise ValueError("kvxopt_b200.kkt.ldl2 does not support nonlinear constraints")
        if state["handle"] is None:
            _create(H)
        if (H is None) != (state["Hnnz"] is None):
            raise ValueError("H must be given in every call or in none")
        di = np.ascontiguousarray(_vec(W["di"], m, "W['di']") if m else np.zeros(0), dtype=np.float64)
        d2 = di * di
        w = state["w"]
        w[:m] = d2
        if H is not None:
            w[m:m + state["Hnnz"]] = _checked_values(H, state["Hpat"])[state["keep"]]
        w[len(w) - len(Ax):] = Ax
        if state["asm"] is None:                 # the map goes to the device once per pattern
            mp, mc, mv = state["asm_csr"]
            hm = C.c_void_p()
            st = fn["b200s_spmv_create"](state["nk"], len(w), L.ptr_i64(mp), L.ptr_i64(mc), L.ptr_f64(mv), C.byref(hm))
            if st != L.OK:
                _raise(st)
            state["asm"] = _SpmvHandle(hm)
        kv_dev = C.c_void_p()
        st = fn["b200s_spmv_apply"](state["asm"].h, L.ptr_f64(w), C.byref(kv_dev))        # K's values, assembled in HBM
        if st != L.OK:
            _raise(st)
        handle = state["handle"]
        minor = C.c_int64(0)
        st = fn["b200s_chol_factorize_dev"](handle.h, kv_dev, C.byref(minor))
        state["factors"] = state.get("factors", 0) + 1
        if st == L.NOT_POSDEF:
            raise ArithmeticError("zero pivot in the LDL' factorization of the KKT matrix (column %d)" % minor.value)
        if st != L.OK:
            _raise(st)
        u = state["u"]

        def solve(x, y, z):
            xf = _vec(x, n, "x")
            yf = _vec(y, p, "y") if p else None
            zf = _vec(z, m, "z") if m else np.zeros(0)
            u[:n] = xf + Gm.T @ (d2 * zf) if m else xf
            if p:
                u[n:] = yf
            st2 = fn["b200s_chol_solve"](handle.h, 0, L.ptr_f64(u), 1, N)
            if st2 != L.OK:
                _raise(st2)
            xf[:] = u[:n]
            if p:
                yf[:] = u[n:]
            if m:
                zf[:] = di * (Gm @ u[:n] - zf)

        return solve

    def _info():
        if state["handle"] is None:
            return {}
        inf = L.CholInfo()
        fn["b200s_chol_info"](state["handle"].h, C.byref(inf))
        return {"order": N, "nnz_L": inf.nnz_L, "flops": inf.flops, "nsuper": inf.nsuper, "max_front_rows": inf.max_front_rows,
                "ms_analyze": inf.ms_analyze, "ms_factor_last": inf.ms_factor, "ms_solve_last": inf.ms_solve,
                "factorizations": state.get("factors", 0)}

    factor.info = _info
    factor._state = state
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
