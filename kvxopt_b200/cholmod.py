"""Drop-in mirror of `kvxopt.cholmod` (reference src/C/cholmod.c) on top of libb200sparse.so.

Same function names, arguments, return values, factor capsules and exceptions as the reference
module (cholmod.c:988-1024): `options`, `symbolic`, `numeric`, `solve`, `spsolve`, `linsolve`,
`splinsolve`, `diag`, `getfactor`.  Matrices may be kvxopt `spmatrix`/`matrix` objects (the reference's
own types, consumed through `.CCS` and the buffer protocol) or scipy.sparse / numpy arrays.

All numeric work runs in the CUDA library; without a GPU every numeric call raises RuntimeError.
`install()` registers this module as `kvxopt.cholmod` so that the unmodified reference
`kvxopt.misc.kkt_chol2` (misc.py:21,1431-1558) uses it.
"""
import ctypes as C
import sys

import numpy as np

from . import _lib as L

fn = L.fn

# reference cholmod.c:98-125 -- recognised keys; anything else raises ValueError on every call
options = {}
# engine extensions that are NOT part of the reference's `options` contract (b200s_chol_opts fields): 'max_merge_cols'
engine_options = {}

_NAME_L = b"CHOLMOD SYM D FACTOR L"     # cholmod.c:44-48
_NAME_U = b"CHOLMOD SYM D FACTOR U"
_NAME_ZL = b"CHOLMOD SYM Z FACTOR L"
_NAME_ZU = b"CHOLMOD SYM Z FACTOR U"

_py = C.pythonapi
_py.PyCapsule_New.restype = C.py_object
_py.PyCapsule_New.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p]
_py.PyCapsule_GetPointer.restype = C.c_void_p
_py.PyCapsule_GetPointer.argtypes = [C.py_object, C.c_char_p]
_py.PyCapsule_GetName.restype = C.c_char_p
_py.PyCapsule_GetName.argtypes = [C.py_object]
_py.PyCapsule_IsValid.restype = C.c_int
_py.PyCapsule_IsValid.argtypes = [C.py_object, C.c_char_p]


# Raw-address variants for use inside capsule destructors: the capsule is being deallocated there
# (refcount 0), so it must not be wrapped in a py_object (that would resurrect and re-free it).
# PYFUNCTYPE: these are CPython API functions and must be called with the GIL held (CFUNCTYPE would drop it).
_raw_GetPointer = C.PYFUNCTYPE(C.c_void_p, C.c_void_p, C.c_char_p)(("PyCapsule_GetPointer", _py))
_raw_GetName = C.PYFUNCTYPE(C.c_char_p, C.c_void_p)(("PyCapsule_GetName", _py))


@C.CFUNCTYPE(None, C.c_void_p)
def _capsule_destructor(capsule_addr):        # kvxopt_free_cholmod_factor, cholmod.c:210-214
    try:
        name = _raw_GetName(capsule_addr)
        ptr = _raw_GetPointer(capsule_addr, name)
        if ptr:
            _FAMILY.pop(ptr, None)
            _ZFLAG.pop(ptr, None)
            fn["b200s_chol_free"](ptr)
    except Exception:   # never raise from a destructor
        pass


def _raise_status(st, what):
    if st == L.OUT_OF_MEMORY:
        raise MemoryError()
    if st in (L.NO_DEVICE, L.CUDA_ERROR):
        raise RuntimeError("kvxopt_b200.cholmod: %s (%s)" % (L.strerror(st), L.last_error()))
    raise ValueError(what)


def _set_options():
    """cholmod.c:87-129: defaults, print=0, supernodal=2, then overlay `options`."""
    o = L.CholOpts()
    fn["b200s_chol_default_opts"](C.byref(o))
    o.max_merge_cols = int(engine_options.get("max_merge_cols", 0))
    for key, value in options.items():
        if not isinstance(key, str):
            continue
        if key == "supernodal" and isinstance(value, int) and not isinstance(value, bool):
            o.supernodal = value
        elif key == "print" and isinstance(value, int) and not isinstance(value, bool):
            pass
        elif key == "nmethods" and isinstance(value, int) and not isinstance(value, bool):
            o.nmethods = value
        elif key == "postorder" and isinstance(value, bool):
            o.postorder = int(value)
        elif key == "dbound" and isinstance(value, float):
            o.dbound = value
        else:
            raise ValueError("invalid value for CHOLMOD parameter: %-.20s" % key)
    return o


# ---- adapters for the reference's matrix types ------------------------------------------------------

def _is_kvx(obj):
    return type(obj).__module__.split(".")[0] in ("kvxopt", "cvxopt")


def _is_spmatrix(A):
    if _is_kvx(A):
        return type(A).__name__ == "spmatrix"
    return hasattr(A, "tocsc") and hasattr(A, "shape")


def _is_dense(B):
    if _is_kvx(B):
        return type(B).__name__ == "matrix"
    return isinstance(B, np.ndarray)


def _typecode(A):
    if _is_kvx(A):
        return A.typecode
    k = np.dtype(A.dtype).kind
    return {"f": "d", "c": "z", "i": "i", "u": "i"}.get(k, "?")


def _size(A):
    return tuple(A.size) if _is_kvx(A) else tuple(A.shape)


def _ccs(A):
    """(colptr int64, rowind int64, values float64) of a sparse matrix without changing its pattern; the values of a complex
    ('z') matrix come back as interleaved (re, im) pairs, 2 nnz doubles"""
    vt = np.complex128 if _typecode(A) == "z" else np.float64
    if _is_kvx(A):
        cp, ri, vx = A.CCS        # fresh copies owned by these matrix objects: viewed, not copied again
        return (_flat_view(cp, np.int64), _flat_view(ri, np.int64), _flat_view(vx, vt).view(np.float64))
    import scipy.sparse as sp
    if not sp.isspmatrix_csc(A) and not (hasattr(sp, "csc_array") and isinstance(A, sp.csc_array)):
        A = A.tocsc()
    if not A.has_sorted_indices:
        A = A.copy()
        A.sort_indices()
    return (np.ascontiguousarray(A.indptr, dtype=np.int64), np.ascontiguousarray(A.indices, dtype=np.int64),
            np.ascontiguousarray(A.data, dtype=vt).view(np.float64))


def _flat_view(m, dtype):
    """1-D numpy view of a kvxopt dense matrix (buffer protocol; the array keeps the matrix alive); a copy only when the
    element type differs"""
    if len(m) == 0:
        return np.zeros(0, dtype=dtype)
    return np.ascontiguousarray(np.asarray(memoryview(m)).reshape(-1), dtype=dtype)


def _values(A):
    """float64 value array of a sparse matrix in CCS order (a view when the container already stores it that way)"""
    if _is_kvx(A):
        return np.array(A.V, dtype=np.float64).reshape(-1)
    import scipy.sparse as sp
    if not sp.isspmatrix_csc(A) and not (hasattr(sp, "csc_array") and isinstance(A, sp.csc_array)):
        A = A.tocsc()
    if not A.has_sorted_indices:
        A = A.copy()
        A.sort_indices()
    return np.ascontiguousarray(A.data, dtype=np.float64)


def _dense_view(B, z=False):
    """flat column-major view that shares memory with B, plus (nrows, ncols); float64, or complex128 with z=True"""
    if _is_kvx(B):
        a = np.asarray(memoryview(B))
        nrows, ncols = B.size
    else:
        a = B
        if a.ndim == 1:
            nrows, ncols = a.shape[0], 1
        else:
            nrows, ncols = a.shape
    if a.dtype != (np.complex128 if z else np.float64):
        raise TypeError("B must a dense matrix of the same numerical type as F")
    if a.ndim == 1:
        if not a.flags.c_contiguous:
            raise TypeError("B must be contiguous")
        flat = a
    else:
        if not a.flags.f_contiguous:
            raise TypeError("B must be stored column-major (Fortran order)")
        flat = a.reshape(-1, order="F")
    if a.size and (not np.shares_memory(flat, a) or not flat.flags.writeable):
        raise TypeError("B must be a writable column-major array")
    return flat, nrows, ncols


def _kv_module(like):
    """the kvxopt package `like` belongs to (an object of it, or its name), else None"""
    import importlib
    if isinstance(like, str):
        return importlib.import_module(like)
    if like is not None and _is_kvx(like):
        return importlib.import_module(type(like).__module__.split(".")[0])
    return None


def _make_spmatrix(like, values, rowind, colptr, size, tc="d"):
    kv = _kv_module(like)
    if kv is not None:
        cols = np.repeat(np.arange(size[1], dtype=np.int64), np.diff(colptr))
        return kv.spmatrix(kv.matrix(values, (len(values), 1), tc) if len(values) else [],
                           kv.matrix(rowind, (len(rowind), 1), "i") if len(rowind) else [],
                           kv.matrix(cols, (len(cols), 1), "i") if len(cols) else [], size, tc)
    import scipy.sparse as sp
    return sp.csc_matrix((values, rowind, colptr), shape=size)


def _make_matrix(like, values, size, tc="d"):
    kv = _kv_module(like)
    if kv is not None:
        return kv.matrix(values, size, tc)
    return np.asarray(values, dtype=np.complex128 if tc == "z" else np.float64).reshape(size, order="F")


def _factor_handle(F, want_typecode=None):
    if type(F).__name__ != "PyCapsule":
        raise TypeError("F is not a Capsule")
    name = _py.PyCapsule_GetName(F)
    if name is None:
        raise TypeError("F is not a Capsule")
    if name not in (_NAME_L, _NAME_U, _NAME_ZL, _NAME_ZU):
        raise TypeError("F is not a CHOLMOD factor")
    return _py.PyCapsule_GetPointer(F, name), ("L" if name in (_NAME_L, _NAME_ZL) else "U")


def _is_z(h):
    """complex factor object (made by b200s_chol_analyze_z: Hermitian matrix factored through its real embedding)"""
    return _info(h).zn > 0 or _ZFLAG.get(h if isinstance(h, int) else h.value, False)


def _order(inf):
    """order of the matrix as the caller sees it (complex order for 'z' factors)"""
    return inf.zn if inf.zn > 0 else inf.n


def _info(h):
    inf = L.CholInfo()
    fn["b200s_chol_info"](h, C.byref(inf))
    return inf


def _analyze(A, p, uplo, o):
    if not _is_spmatrix(A) or _size(A)[0] != _size(A)[1]:
        raise TypeError("A is not a square sparse matrix")
    if _typecode(A) not in ("d", "z"):
        raise TypeError("A is not a square sparse matrix")
    n = _size(A)[0]
    perm = None
    if p is not None:
        if _is_kvx(p):
            if type(p).__name__ != "matrix" or p.typecode != "i":
                raise TypeError("p must be a matrix with typecode 'i'")
            perm = np.array(p, dtype=np.int64).reshape(-1)
        else:
            perm = np.asarray(p)
            if perm.dtype.kind not in "iu":
                raise TypeError("p must be a matrix with typecode 'i'")
            perm = np.ascontiguousarray(perm, dtype=np.int64).reshape(-1)
        if perm.size != n:
            raise TypeError("length of p is too small")
        if n and (perm.min() < 0 or perm.max() >= n or np.unique(perm).size != n):
            raise ValueError("p is not a valid permutation")
    if uplo not in ("L", "U"):
        raise ValueError("possible values of uplo are: 'L', 'U'")
    cp, ri, vx = _ccs(A)
    h = C.c_void_p()
    z = _typecode(A) == "z"
    st = fn["b200s_chol_analyze_z" if z else "b200s_chol_analyze"](n, L.ptr_i64(cp), L.ptr_i64(ri), uplo.encode(),
                                                                   L.ptr_i64(perm), C.byref(o), C.byref(h))
    if st != L.OK:
        _raise_status(st, "symbolic factorization failed")
    if z:
        _ZFLAG[h.value] = True      # also marks complex factors of order 0 (zn = 0 there)
    return h, vx


def _factorize(h, vx, cp=None, ri=None):
    minor = C.c_int64(0)
    st = fn["b200s_chol_factorize_z" if _is_z(h) else "b200s_chol_factorize"](h, L.ptr_i64(cp), L.ptr_i64(ri), L.ptr_f64(vx),
                                                                              C.byref(minor))
    if st == L.NOT_POSDEF:
        raise ArithmeticError(int(minor.value))
    if st != L.OK:
        _raise_status(st, "factorization failed")


def symbolic(A, p=None, uplo="L"):
    """F = symbolic(A, p=None, uplo='L')  -- cholmod.c:244-291"""
    o = _set_options()
    h, _ = _analyze(A, p, uplo, o)
    _FAMILY[h.value] = type(A).__module__.split(".")[0] if _is_kvx(A) else None   # family diag()/getfactor() return
    if _typecode(A) == "z":         # cholmod.c:286-290: the capsule name carries the numerical type
        return _py.PyCapsule_New(h, _NAME_ZL if uplo == "L" else _NAME_ZU, C.cast(_capsule_destructor, C.c_void_p))
    return _py.PyCapsule_New(h, _NAME_L if uplo == "L" else _NAME_U, C.cast(_capsule_destructor, C.c_void_p))


def numeric(A, F):
    """numeric(A, F): numeric factorization with the pattern analysed in F -- cholmod.c:322-398.
    Raises ArithmeticError(k) when the matrix is not positive definite (k = failing column)."""
    _set_options()
    if not _is_spmatrix(A) or _size(A)[0] != _size(A)[1]:
        raise TypeError("A is not a sparse matrix")
    h, uplo = _factor_handle(F)
    if _typecode(A) != ("z" if _is_z(h) else "d"):      # cholmod.c:343-357
        raise TypeError("F is not the CHOLMOD factor of a '%s' matrix" % _typecode(A))
    inf = _info(h)
    if _size(A)[0] != _order(inf):
        raise ValueError("factorization failed")
    # A's own pattern travels with the values (cholmod.c:340-358 rebuilds the cholmod_sparse from A on every call): the
    # library compares it with the analysed pattern and re-maps the values when A stores a subset of it
    cp, ri, vx = _ccs(A)
    _factorize(h, vx, cp, ri)


def _check_numeric(h):
    inf = _info(h)
    if inf.minor < inf.n:
        raise ArithmeticError("singular matrix")
    if not inf.is_numeric and inf.n > 0:
        raise ValueError("called with symbolic factor")
    return inf


def _solve_dense(h, n, B, sys, nrhs, ldB, offsetB):
    z = _is_z(h)
    if not _is_dense(B) or _typecode(B) != ("z" if z else "d"):      # cholmod.c:460-464
        raise TypeError("B must a dense matrix of the same numerical type as F")
    flat, nrows, ncols = _dense_view(B, z)
    if nrhs < 0:
        nrhs = ncols
    if n == 0 or nrhs == 0:
        return
    if ldB == 0:
        ldB = max(1, nrows)
    if ldB < max(1, n):
        raise ValueError("ldB must be at least max(1,n)")
    if offsetB < 0:
        raise ValueError("offsetB must be a nonnegative integer")
    if offsetB + (nrhs - 1) * ldB + n > flat.size:
        raise TypeError("length of B is too small")
    if z:       # a complex vector IS its real embedding: (re, im) pairs, 2n real unknowns per column
        sub = flat[offsetB:].view(np.float64)
        st = fn["b200s_chol_solve"](h, sys, L.ptr_f64(sub), nrhs, 2 * ldB)
    else:
        sub = flat[offsetB:]
        st = fn["b200s_chol_solve"](h, sys, L.ptr_f64(sub), nrhs, ldB)
    if st != L.OK:
        _raise_status(st, "solve step failed")


def solve(F, B, sys=0, nrhs=-1, ldB=0, offsetB=0):
    """solve(F, B, sys=0, nrhs, ldB, offsetB): B overwritten by the solution -- cholmod.c:429-499"""
    _set_options()
    h, _ = _factor_handle(F)
    inf = _check_numeric(h)
    if sys < 0 or sys > 8:
        raise ValueError("invalid value for sys")
    _solve_dense(h, _order(inf), B, sys, nrhs, ldB, offsetB)


def _spsolve(h, n, B, sys):
    z = _is_z(h)
    if not _is_spmatrix(B) or _typecode(B) != ("z" if z else "d"):   # cholmod.c:557-561
        raise TypeError("B must a sparse matrix of the same numerical type as F")
    if _size(B)[0] != n:
        raise ValueError("incompatible dimensions for B")
    bp, bi, bx = _ccs(B)
    ncols = _size(B)[1]
    xp, xi, xx = L.p_i64(), L.p_i64(), L.p_f64()
    st = fn["b200s_chol_spsolve_z" if z else "b200s_chol_spsolve"](h, sys, n, ncols, L.ptr_i64(bp), L.ptr_i64(bi), L.ptr_f64(bx),
                                                                   C.byref(xp), C.byref(xi), C.byref(xx))
    if st != L.OK:
        _raise_status(st, "solve step failed")
    colptr = L.take_array(xp, ncols + 1, np.int64)
    nnz = int(colptr[-1])
    rowind = L.take_array(xi, max(nnz, 1), np.int64)[:nnz]
    if z:
        values = L.take_array(xx, 2 * max(nnz, 1), np.float64)[:2 * nnz].view(np.complex128)
    else:
        values = L.take_array(xx, max(nnz, 1), np.float64)[:nnz]
    return _make_spmatrix(B, values, rowind, colptr, (n, ncols), "z" if z else "d")


def spsolve(F, B, sys=0):
    """X = spsolve(F, B, sys=0) with sparse B, returns a new sparse matrix -- cholmod.c:524-587"""
    _set_options()
    h, _ = _factor_handle(F)
    inf = _check_numeric(h)
    if sys < 0 or sys > 8:
        raise ValueError("invalid value for sys")
    return _spsolve(h, _order(inf), B, sys)


def linsolve(A, B, p=None, uplo="L", nrhs=-1, ldB=0, offsetB=0):
    """linsolve(A, B, p=None, uplo='L', nrhs, ldB, offsetB): solves A X = B in place -- cholmod.c:618-753"""
    o = _set_options()
    if not _is_spmatrix(A) or _size(A)[0] != _size(A)[1]:
        raise TypeError("A is not a sparse matrix")
    n = _size(A)[0]
    if not _is_dense(B) or _typecode(B) != _typecode(A):
        raise TypeError("B must be a dense matrix of the same numerical type as A")
    flat, nrows, ncols = _dense_view(B, _typecode(A) == "z")
    nr = ncols if nrhs < 0 else nrhs
    if n == 0 or nr == 0:
        return
    ld = max(1, nrows) if ldB == 0 else ldB
    if ld < max(1, n):
        raise ValueError("ldB must be at least max(1,n)")
    if offsetB < 0:
        raise ValueError("offsetB must be a nonnegative integer")
    if offsetB + (nr - 1) * ld + n > flat.size:
        raise TypeError("length of B is too small")
    h, vx = _analyze(A, p, uplo, o)
    try:
        _factorize(h, vx)
        _solve_dense(h, n, B, 0, nrhs, ldB, offsetB)
    finally:
        _ZFLAG.pop(h.value, None)
        fn["b200s_chol_free"](h)


def splinsolve(A, B, p=None, uplo="L"):
    """X = splinsolve(A, B, p=None, uplo='L') with sparse B -- cholmod.c:774-881"""
    o = _set_options()
    if not _is_spmatrix(A) or _size(A)[0] != _size(A)[1]:
        raise TypeError("A is not a square sparse matrix")
    n = _size(A)[0]
    if not _is_spmatrix(B) or _typecode(A) != _typecode(B):
        raise TypeError("B must be a sparse matrix of the same type as A")
    if _size(B)[0] != n:
        raise ValueError("incompatible dimensions for B")
    h, vx = _analyze(A, p, uplo, o)
    try:
        _factorize(h, vx)
        return _spsolve(h, n, B, 0)
    finally:
        _ZFLAG.pop(h.value, None)
        fn["b200s_chol_free"](h)


def diag(F):
    """d = diag(F): diagonal of the supernodal Cholesky factor L as an n x 1 matrix -- cholmod.c:900-945"""
    _set_options()
    h, _ = _factor_handle(F)
    inf = _info(h)
    if not inf.is_numeric and inf.n > 0:
        raise ValueError("F must be a numeric Cholesky factor")
    z = _is_z(h)
    n = _order(inf)
    d = np.zeros(2 * n if z else n, dtype=np.float64)
    st = fn["b200s_chol_diag_z" if z else "b200s_chol_diag"](h, L.ptr_f64(d))
    if st == L.INVALID:       # LDL' semantics (options['supernodal'] = 0): cholmod.c:919-922
        raise ValueError("F must be a nonsingular supernodal Cholesky factor")
    if st != L.OK:
        _raise_status(st, "diag failed")
    if z:                     # cholmod.c:923-924: a 'z' matrix for complex factors
        return _make_matrix(_FAMILY.get(h), d.view(np.complex128), (n, 1), "z")
    return _make_matrix(_FAMILY.get(h), d, (n, 1))


def getfactor(F):
    """L = getfactor(F): the Cholesky factor as a sparse lower-triangular matrix -- cholmod.c:948-985"""
    _set_options()
    h, _ = _factor_handle(F)
    inf = _info(h)
    if not inf.is_numeric and inf.n > 0:
        raise ValueError("F must be a numeric Cholesky factor")
    lp, li, lx = L.p_i64(), L.p_i64(), L.p_f64()
    z = _is_z(h)
    n = _order(inf)
    st = fn["b200s_chol_get_L_z" if z else "b200s_chol_get_L"](h, C.byref(lp), C.byref(li), C.byref(lx))
    if st != L.OK:
        _raise_status(st, "getfactor failed")
    colptr = L.take_array(lp, n + 1, np.int64)
    nnz = int(colptr[-1]) if n else 0
    rowind = L.take_array(li, max(nnz, 1), np.int64)[:nnz]
    if z:
        values = L.take_array(lx, 2 * max(nnz, 1), np.float64)[:2 * nnz].view(np.complex128)
    else:
        values = L.take_array(lx, max(nnz, 1), np.float64)[:nnz]
    return _make_spmatrix(_FAMILY.get(h), values, rowind, colptr, (n, n), "z" if z else "d")


# diag/getfactor have no matrix argument: they return the container family (kvxopt or scipy/numpy) of the
# matrix that was given to symbolic().  Keyed by factor handle, cleared by the capsule destructor.
_FAMILY = {}
_ZFLAG = {}        # handles of complex factor objects (covers order 0, where the library's zn is 0 as well)


def factor_info(F):
    """extension: symbolic/numeric statistics of a factor (nnz(L), flops, timings) as a dict"""
    h, _ = _factor_handle(F)
    return _info(h).asdict()


def set_solve_sweeps(F, mode=-1):
    """extension: bit 0 / bit 1 of `mode` run the forward / backward sweep of one-right-hand-side solves as one persistent
    kernel per level (the default, -1) instead of two launches per 128-column block step; bit-identical solutions"""
    h, _ = _factor_handle(F)
    st = fn["b200s_chol_set_solve_sweeps"](h, int(mode))
    if st != L.OK:
        _raise_status(st, "set_solve_sweeps failed")


def factor_perm(F):
    """extension: the fill-reducing permutation held by F (L->Perm)"""
    h, _ = _factor_handle(F)
    inf = _info(h)
    p = np.zeros(inf.n, dtype=np.int64)
    fn["b200s_chol_get_perm"](h, L.ptr_i64(p))
    return p[::2] // 2 if inf.zn > 0 else p      # complex factors: the pair ordering (2k, 2k+1) back to complex indices


def install(kvxopt_module=None):
    """Register this module as `<kvxopt>.cholmod` (what reference src/python/misc.py:21 imports)."""
    if kvxopt_module is None:
        import kvxopt as kvxopt_module
    name = kvxopt_module.__name__
    sys.modules[name + ".cholmod"] = sys.modules[__name__]
    setattr(kvxopt_module, "cholmod", sys.modules[__name__])
    # misc.py:21 binds `cholmod` at import time: rebind it where that already happened, or the KKT solvers of an already
    # imported kvxopt.misc would silently keep the module they found first
    for sub in ("misc", "coneprog", "cvxprog", "solvers"):
        m = sys.modules.get(name + "." + sub)
        if m is not None and hasattr(m, "cholmod"):
            setattr(m, "cholmod", sys.modules[__name__])
    return sys.modules[__name__]
