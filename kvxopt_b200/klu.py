"""Drop-in mirror of `kvxopt.klu` (reference src/C/klu.c) on top of libb200sparse.so.

Same function names, arguments, return values, capsule names and exceptions as the reference module
(klu.c:830-858): `linsolve`, `symbolic`, `numeric`, `solve`, `get_numeric`, `get_det` -- plus the
batched extension the B200 engine exists for: `refactor_batch` / `solve_batch` (klu_refactor
semantics: same pattern, same pivot sequence, fresh row scaling; the reference never calls
klu_refactor, klu.c:296-301 documents it only).

The pivot search of `numeric` runs once on the host and its values are the factor of the analysed matrix; batched
refactorizations and all solves run on the device.  Complex ('z') matrices: the complex factor of the host pivot search serves
`get_numeric` / `get_det`, the solves run on the device through the real embedding of order 2n.
No GPU => RuntimeError.
"""
import ctypes as C
import sys

import numpy as np

from . import _lib as L
from .cholmod import (_ccs, _dense_view, _is_dense, _is_kvx, _is_spmatrix, _make_spmatrix, _py, _raw_GetPointer, _size,
                      _typecode)

fn = L.fn

_NAME_SYM = b"KLU SYM D FACTOR"      # klu.c:36-39
_NAME_NUM = b"KLU NUM D FACTOR"
_NAME_SYM_Z = b"KLU SYM Z FACTOR"
_NAME_NUM_Z = b"KLU NUM Z FACTOR"


@C.CFUNCTYPE(None, C.c_void_p)
def _free_symbolic(capsule_addr):            # free_klu_d_symbolic, klu.c:51-61
    try:
        ptr = _raw_GetPointer(capsule_addr, _NAME_SYM)
        if ptr:
            fn["b200s_klu_free_symbolic"](ptr)
    except Exception:
        pass


@C.CFUNCTYPE(None, C.c_void_p)
def _free_symbolic_z(capsule_addr):
    try:
        ptr = _raw_GetPointer(capsule_addr, _NAME_SYM_Z)
        if ptr:
            fn["b200s_klu_free_symbolic"](ptr)
    except Exception:
        pass


@C.CFUNCTYPE(None, C.c_void_p)
def _free_numeric_z(capsule_addr):           # free_klu_z_numeric, klu.c:74-81
    try:
        ptr = _raw_GetPointer(capsule_addr, _NAME_NUM_Z)
        if ptr:
            fn["b200s_klu_free_numeric"](ptr)
    except Exception:
        pass


@C.CFUNCTYPE(None, C.c_void_p)
def _free_numeric(capsule_addr):             # free_klu_d_numeric, klu.c:63-72
    try:
        ptr = _raw_GetPointer(capsule_addr, _NAME_NUM)
        if ptr:
            _pending.pop(int(ptr), None)         # batches begun and never ended die with the object
            fn["b200s_klu_free_numeric"](ptr)
    except Exception:
        pass


def _raise_status(st):
    if st == L.OUT_OF_MEMORY:
        raise MemoryError()
    if st == L.SINGULAR:
        raise ArithmeticError("singular matrix")
    if st in (L.NO_DEVICE, L.CUDA_ERROR):
        raise RuntimeError("kvxopt_b200.klu: %s (%s)" % (L.strerror(st), L.last_error()))
    raise ValueError("KLU ERROR %d" % st)


def _check_A(A, msg):
    if not _is_spmatrix(A) or _size(A)[0] != _size(A)[1]:
        raise TypeError(msg)
    if _typecode(A) not in ("d", "z"):
        raise TypeError(msg)
    return _typecode(A) == "z"


def _capsule_ptr(F, name, msg, argname):
    if type(F).__name__ != "PyCapsule":
        raise TypeError("%s is not a Capsule" % argname)
    got = _py.PyCapsule_GetName(F)
    if got != name:
        raise TypeError(msg)
    return _py.PyCapsule_GetPointer(F, name)


def _analyze(cp, ri, n):
    h = C.c_void_p()
    st = fn["b200s_klu_analyze"](n, L.ptr_i64(cp), L.ptr_i64(ri), C.byref(h))
    if st != L.OK:
        _raise_status(st)
    return h


def _factor(hs, cp, ri, vx, z=False):
    h = C.c_void_p()
    st = fn["b200s_klu_factor_z" if z else "b200s_klu_factor"](hs, L.ptr_i64(cp), L.ptr_i64(ri), L.ptr_f64(vx), C.byref(h))
    if st != L.OK:
        _raise_status(st)
    return h


def _dense_args(B, n, nrhs, ldB, offsetB, z=False):
    flat, nrows, ncols = _dense_view(B, z)
    if nrhs < 0:
        nrhs = ncols
    if n == 0 or nrhs == 0:
        return None
    if ldB == 0:
        ldB = max(1, nrows)
    if ldB < max(1, n):
        raise ValueError("ldB must be at least max(1,n)")
    if offsetB < 0:
        raise ValueError("offsetB must be a nonnegative integer")
    if offsetB + (nrhs - 1) * ldB + n > flat.size:
        raise TypeError("length of B is too small")
    return (flat[offsetB:].view(np.float64) if z else flat[offsetB:]), nrhs, ldB


def _trans_flag(trans, z=False):
    if trans not in ("N", "T", "C"):
        raise ValueError("possible values of trans are: 'N', 'T', 'C'")
    if z:
        return {"N": 0, "T": 1, "C": 2}[trans]       # klu.c:661-668: klu_zl_tsolve with conj_solve for 'C'
    return 0 if trans == "N" else 1      # real matrices: 'C' == 'T'


def _solve_call(hn, z, t, args):
    st = fn["b200s_klu_solve_z" if z else "b200s_klu_solve"](hn, t, L.ptr_f64(args[0]), args[1], args[2])
    if st != L.OK:
        _raise_status(st)


def linsolve(A, B, trans="N", nrhs=-1, ldB=0, offsetB=0):
    """linsolve(A, B, trans='N', nrhs, ldB, offsetB): solves A X = B (or A^T X = B) in place -- klu.c:94-230.
    Returns 0 on empty input like the reference (klu.c:126), None otherwise."""
    z = _check_A(A, "A must be a square sparse matrix")
    n = _size(A)[0]
    if not _is_dense(B) or _typecode(B) != _typecode(A):
        raise TypeError("B must a dense matrix of the same numeric type as A")
    args = _dense_args(B, n, nrhs, ldB, offsetB, z)
    if args is None:
        return 0
    t = _trans_flag(trans, z)
    cp, ri, vx = _ccs(A)
    hs = _analyze(cp, ri, n)
    try:
        hn = _factor(hs, cp, ri, vx, z)
        try:
            _solve_call(hn, z, t, args)
        finally:
            fn["b200s_klu_free_numeric"](hn)
    finally:
        fn["b200s_klu_free_symbolic"](hs)
    return None


def symbolic(A):
    """Fs = symbolic(A): BTF + per-block AMD ordering -- klu.c:242-291"""
    z = _check_A(A, "A must be a square sparse matrix")
    cp, ri, _ = _ccs(A)
    h = _analyze(cp, ri, _size(A)[0])
    if z:       # the analysis is the same for both types (klu.c:266); the capsule name carries the type (klu.c:284-288)
        return _py.PyCapsule_New(h, _NAME_SYM_Z, C.cast(_free_symbolic_z, C.c_void_p))
    return _py.PyCapsule_New(h, _NAME_SYM, C.cast(_free_symbolic, C.c_void_p))


def numeric(A, Fs):
    """Fn = numeric(A, Fs): numeric LU with partial pivoting -- klu.c:310-379.
    Raises ArithmeticError("singular matrix") for singular input."""
    z = _check_A(A, "A must a square sparse matrix")
    hs = _capsule_ptr(Fs, _NAME_SYM_Z if z else _NAME_SYM, "Fs is not the KLU symbolic factor of a '%s' matrix" % ("z" if z else "d"), "Fs")
    cp, ri, vx = _ccs(A)
    h = _factor(hs, cp, ri, vx, z)
    if z:
        return _py.PyCapsule_New(h, _NAME_NUM_Z, C.cast(_free_numeric_z, C.c_void_p))
    return _py.PyCapsule_New(h, _NAME_NUM, C.cast(_free_numeric, C.c_void_p))


def solve(A, Fs, F, B, trans="N", nrhs=-1, ldB=0, offsetB=0):
    """solve(A, Fs, F, B, trans='N', nrhs, ldB, offsetB): B overwritten by the solution -- klu.c:593-690"""
    z = _check_A(A, "A must a square sparse matrix")
    n = _size(A)[0]
    tc = "z" if z else "d"
    hn = _capsule_ptr(F, _NAME_NUM_Z if z else _NAME_NUM, "F is not the KLU numeric factor of a '%s' matrix" % tc, "F")
    _capsule_ptr(Fs, _NAME_SYM_Z if z else _NAME_SYM, "F is not the KLU symbolic factor of a '%s' matrix" % tc, "Fs")
    if not _is_dense(B) or _typecode(B) != _typecode(A):
        raise TypeError("B must a dense matrix of the same numeric type as A")
    args = _dense_args(B, n, nrhs, ldB, offsetB, z)
    if args is None:
        return
    _solve_call(hn, z, _trans_flag(trans, z), args)


def _extract_z(hn):
    inf = L.KluInfo()
    fn["b200s_klu_info"](hn, C.byref(inf))
    n = inf.n
    Lp = np.zeros(n + 1, np.int64); Up = np.zeros(n + 1, np.int64); Fp = np.zeros(n + 1, np.int64)
    Li = np.zeros(max(inf.nnz_L, 1), np.int64); Ui = np.zeros(max(inf.nnz_U, 1), np.int64); Fi = np.zeros(max(inf.nnz_F, 1), np.int64)
    Lx = np.zeros(2 * max(inf.nnz_L, 1)); Ux = np.zeros(2 * max(inf.nnz_U, 1)); Fx = np.zeros(2 * max(inf.nnz_F, 1))
    P = np.zeros(max(n, 1), np.int64); Q = np.zeros(max(n, 1), np.int64); Rs = np.zeros(max(n, 1)); R = np.zeros(inf.nblocks + 1, np.int64)
    st = fn["b200s_klu_extract_z"](hn, L.ptr_i64(Lp), L.ptr_i64(Li), L.ptr_f64(Lx), L.ptr_i64(Up), L.ptr_i64(Ui), L.ptr_f64(Ux),
                                   L.ptr_i64(Fp), L.ptr_i64(Fi), L.ptr_f64(Fx), L.ptr_i64(P), L.ptr_i64(Q), L.ptr_f64(Rs), L.ptr_i64(R))
    if st != L.OK:
        _raise_status(st)
    cz = lambda a, k: a.view(np.complex128)[:k]
    return dict(n=n, Lp=Lp, Li=Li[:inf.nnz_L], Lx=cz(Lx, inf.nnz_L), Up=Up, Ui=Ui[:inf.nnz_U], Ux=cz(Ux, inf.nnz_U),
                Fp=Fp, Fi=Fi[:inf.nnz_F], Fx=cz(Fx, inf.nnz_F), P=P[:n], Q=Q[:n], Rs=Rs[:n], R=R)


def _extract(hn, batch_index=None):
    inf = L.KluInfo()
    fn["b200s_klu_info"](hn, C.byref(inf))
    n = inf.n
    Lp = np.zeros(n + 1, np.int64); Up = np.zeros(n + 1, np.int64); Fp = np.zeros(n + 1, np.int64)
    Li = np.zeros(max(inf.nnz_L, 1), np.int64); Ui = np.zeros(max(inf.nnz_U, 1), np.int64); Fi = np.zeros(max(inf.nnz_F, 1), np.int64)
    Lx = np.zeros(max(inf.nnz_L, 1)); Ux = np.zeros(max(inf.nnz_U, 1)); Fx = np.zeros(max(inf.nnz_F, 1))
    P = np.zeros(max(n, 1), np.int64); Q = np.zeros(max(n, 1), np.int64); Rs = np.zeros(max(n, 1)); R = np.zeros(inf.nblocks + 1, np.int64)
    st = fn["b200s_klu_extract"](hn, L.ptr_i64(Lp), L.ptr_i64(Li), L.ptr_f64(Lx), L.ptr_i64(Up), L.ptr_i64(Ui), L.ptr_f64(Ux),
                                 L.ptr_i64(Fp), L.ptr_i64(Fi), L.ptr_f64(Fx), L.ptr_i64(P), L.ptr_i64(Q), L.ptr_f64(Rs), L.ptr_i64(R))
    if st != L.OK:
        _raise_status(st)
    if batch_index is not None:
        st = fn["b200s_klu_extract_batch"](hn, batch_index, L.ptr_f64(Lx), L.ptr_f64(Ux), L.ptr_f64(Fx), L.ptr_f64(Rs))
        if st != L.OK:
            _raise_status(st)
    return dict(n=n, Lp=Lp, Li=Li[:inf.nnz_L], Lx=Lx[:inf.nnz_L], Up=Up, Ui=Ui[:inf.nnz_U], Ux=Ux[:inf.nnz_U],
                Fp=Fp, Fi=Fi[:inf.nnz_F], Fx=Fx[:inf.nnz_F], P=P[:n], Q=Q[:n], Rs=Rs[:n], R=R)


def get_numeric(A, Fs, Fn):
    """L, U, P, Q, R, F, r = get_numeric(A, Fs, Fn) with R*P*A*Q = L*U + F -- klu.c:392-566
    (R is returned already inverted, klu.c:516-522)."""
    z = _check_A(A, "A must a square sparse matrix")
    tc = "z" if z else "d"
    hn = _capsule_ptr(Fn, _NAME_NUM_Z if z else _NAME_NUM, "F is not the KLU numeric factor of a '%s' matrix" % tc, "F")
    _capsule_ptr(Fs, _NAME_SYM_Z if z else _NAME_SYM, "F is not the KLU symbolic factor of a '%s' matrix" % tc, "Fs")
    e = _extract_z(hn) if z else _extract(hn)
    n = e["n"]
    ar = np.arange(n + 1, dtype=np.int64)
    Lm = _make_spmatrix(A, e["Lx"], e["Li"], e["Lp"], (n, n), tc)          # complex L, U, F for 'z' (klu.c:468-479, 486-512)
    Um = _make_spmatrix(A, e["Ux"], e["Ui"], e["Up"], (n, n), tc)
    Fm = _make_spmatrix(A, e["Fx"], e["Fi"], e["Fp"], (n, n), tc)
    prow = np.zeros(n, dtype=np.int64)
    prow[e["P"]] = np.arange(n)                     # P[i, Pnum[i]] = 1  (klu.c:526-532)
    Pm = _make_spmatrix(A, np.ones(n), prow, ar, (n, n))
    Qm = _make_spmatrix(A, np.ones(n), e["Q"], ar, (n, n))      # Q[Q[i], i] = 1  (klu.c:535-541)
    Rm = _make_spmatrix(A, 1.0 / e["Rs"], np.arange(n, dtype=np.int64), ar, (n, n))
    return Lm, Um, Pm, Qm, Rm, Fm, [int(v) for v in e["R"]]


def _perm_swaps(p):
    """number of swaps that sort the permutation (parity as computed in klu.c:785-806)"""
    w = np.array(p, dtype=np.int64)
    npiv = 0
    for i in range(len(w)):
        while w[i] != i:
            j = w[i]
            w[i], w[j] = w[j], w[i]
            npiv += 1
    return npiv


def get_det(A, Fs, Fn):
    """d = get_det(A, Fs, Fn): determinant from the factors -- klu.c:707-828"""
    z = _check_A(A, "A must a square sparse matrix")
    tc = "z" if z else "d"
    hn = _capsule_ptr(Fn, _NAME_NUM_Z if z else _NAME_NUM, "F is not the UMFPACK numeric factor of a '%s' matrix" % tc, "F")
    _capsule_ptr(Fs, _NAME_SYM_Z if z else _NAME_SYM, "F is not the UMFPACK symbolic factor of a '%s' matrix" % tc, "Fs")
    e = _extract_z(hn) if z else _extract(hn)
    n = e["n"]
    det = 1.0 + 0.0j if z else 1.0
    for k in range(n):
        det *= e["Ux"][e["Up"][k + 1] - 1] * e["Rs"][k]
    sign = -1.0 if (_perm_swaps(e["P"]) + _perm_swaps(e["Q"])) % 2 else 1.0
    return det * sign


# ---- batched extension ------------------------------------------------------------------------------

def refactor_batch(Fn, values, check=True):
    """Refactor `batch` matrices that share the pattern and pivot sequence held by Fn.
    values: float64 array of shape (batch, nnz), row b = the CCS value array of matrix b.
    Returns an int array of per-matrix status (0 ok, 2 singular); raises ArithmeticError when
    check=True and any matrix hit a zero pivot."""
    hn = _capsule_ptr(Fn, _NAME_NUM, "F is not the KLU numeric factor of a 'd' matrix", "F")
    v = np.ascontiguousarray(values, dtype=np.float64)
    if v.ndim != 2:
        raise TypeError("values must have shape (batch, nnz)")
    inf = L.KluInfo()
    fn["b200s_klu_info"](hn, C.byref(inf))
    if v.shape[1] != inf.nnz_A:
        raise ValueError("values must have nnz(A) = %d columns" % inf.nnz_A)
    status = np.zeros(v.shape[0], dtype=np.int32)
    st = fn["b200s_klu_refactor_batch"](hn, L.ptr_f64(v), v.shape[0], v.shape[1], status.ctypes.data_as(L.p_int))
    if st != L.OK:
        _raise_status(st)
    if check and status.any():
        raise ArithmeticError("singular matrix")
    return status


def refactor_batch_begin(Fn, values):
    """Pipelined form of refactor_batch: enqueue the upload of `values` (float64, shape (batch, nnz), C-contiguous,
    ideally pinned host memory; keep it alive and unmodified until the matching refactor_batch_end) and the
    refactorization kernels, and return at once.  Up to two batches may be in flight; the upload of the second
    overlaps the kernels of the first."""
    hn = _capsule_ptr(Fn, _NAME_NUM, "F is not the KLU numeric factor of a 'd' matrix", "F")
    if not isinstance(values, np.ndarray) or values.dtype != np.float64 or values.ndim != 2 or not values.flags.c_contiguous:
        raise TypeError("values must be a C-contiguous float64 array of shape (batch, nnz)")
    inf = L.KluInfo()
    fn["b200s_klu_info"](hn, C.byref(inf))
    if values.shape[1] != inf.nnz_A:
        raise ValueError("values must have nnz(A) = %d columns" % inf.nnz_A)
    if values.shape[0] == 0 or inf.n == 0:
        return                    # the library queues nothing for an empty batch: nothing to end either
    st = fn["b200s_klu_refactor_batch_begin"](hn, L.ptr_f64(values), values.shape[0], values.shape[1])
    if st != L.OK:
        _raise_status(st)
    _pending.setdefault(_key(hn), []).append(values.shape[0])


def refactor_batch_end(Fn, check=True):
    """Wait for the oldest batch begun with refactor_batch_begin; returns its per-matrix status array."""
    hn = _capsule_ptr(Fn, _NAME_NUM, "F is not the KLU numeric factor of a 'd' matrix", "F")
    q = _pending.get(_key(hn), [])
    if not q:
        raise ValueError("no batch in flight")
    batch = q.pop(0)
    status = np.zeros(batch, dtype=np.int32)
    st = fn["b200s_klu_refactor_batch_end"](hn, status.ctypes.data_as(L.p_int))
    if st != L.OK:
        _raise_status(st)
    if check and status.any():
        raise ArithmeticError("singular matrix")
    return status


_pending = {}


def _key(h):
    return int(h.value) if hasattr(h, "value") else int(h)


def solve_batch(Fn, B, trans="N"):
    """Solve with every matrix of the last refactored batch.  B: float64 array (batch, nrhs, n) or
    (batch, n), overwritten in place (must be C-contiguous)."""
    hn = _capsule_ptr(Fn, _NAME_NUM, "F is not the KLU numeric factor of a 'd' matrix", "F")
    if not isinstance(B, np.ndarray) or B.dtype != np.float64 or not B.flags.c_contiguous:
        raise TypeError("B must be a C-contiguous float64 array of shape (batch, nrhs, n)")
    b3 = B.reshape(B.shape[0], 1, B.shape[1]) if B.ndim == 2 else B
    batch, nrhs, n = b3.shape
    st = fn["b200s_klu_solve_batch"](hn, _trans_flag(trans), L.ptr_f64(b3), nrhs, n, batch)
    if st != L.OK:
        _raise_status(st)


def factor_info(Fn):
    hn = _capsule_ptr(Fn, _NAME_NUM, "F is not the KLU numeric factor of a 'd' matrix", "F")
    inf = L.KluInfo()
    fn["b200s_klu_info"](hn, C.byref(inf))
    return inf.asdict()


def install(kvxopt_module=None):
    """Register this module as `<kvxopt>.klu`."""
    if kvxopt_module is None:
        import kvxopt as kvxopt_module
    name = kvxopt_module.__name__
    sys.modules[name + ".klu"] = sys.modules[__name__]
    setattr(kvxopt_module, "klu", sys.modules[__name__])
    for sub in ("misc", "coneprog", "cvxprog", "solvers"):       # rebind where `klu` was bound at import time
        m = sys.modules.get(name + "." + sub)
        if m is not None and hasattr(m, "klu"):
            setattr(m, "klu", sys.modules[__name__])
    return sys.modules[__name__]
