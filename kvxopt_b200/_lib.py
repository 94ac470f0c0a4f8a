"""ctypes binding of libb200sparse.so (include/b200sparse.h).

The library is the product: there is no Python or CPU implementation behind these calls.  Importing
this module fails loudly when the shared library has not been built (python -m kvxopt_b200.build).
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libb200sparse.so")

if not os.path.exists(LIB_PATH):
    raise ImportError(
        "kvxopt_b200: %s is missing; build it with `python -m kvxopt_b200.build` "
        "(there is no CPU fallback)" % LIB_PATH)

lib = C.CDLL(LIB_PATH)

i64 = C.c_int64
p_i64 = C.POINTER(C.c_int64)
p_f64 = C.POINTER(C.c_double)
p_int = C.POINTER(C.c_int)
vp = C.c_void_p

OK, NOT_POSDEF, SINGULAR = 0, 1, 2
OUT_OF_MEMORY, TOO_LARGE, INVALID, NO_DEVICE, CUDA_ERROR = -2, -3, -4, -5, -6


from ._lib_types import CholOpts, CholInfo, KluInfo, KluPlanView, KktInfo, KktdInfo  # noqa: E402,F401


def _sig(name, restype, *argtypes):
    f = getattr(lib, name)
    f.restype = restype
    f.argtypes = list(argtypes)
    return f


# names here must match include/b200sparse.h exactly (tests/test_abi.py checks the header against this list)
SIGNATURES = {
    "b200s_strerror": (C.c_char_p, C.c_int),
    "b200s_last_error": (C.c_char_p,),
    "b200s_version": (C.c_char_p,),
    "b200s_device_count": (C.c_int,),
    "b200s_set_device": (C.c_int, C.c_int),
    "b200s_chol_default_opts": (None, C.POINTER(CholOpts)),
    "b200s_chol_analyze": (C.c_int, i64, p_i64, p_i64, C.c_char, p_i64, C.POINTER(CholOpts), C.POINTER(vp)),
    "b200s_chol_factorize": (C.c_int, vp, p_i64, p_i64, p_f64, p_i64),
    "b200s_chol_factorize_dev": (C.c_int, vp, vp, p_i64),
    "b200s_chol_solve": (C.c_int, vp, C.c_int, p_f64, i64, i64),
    "b200s_chol_solve_dev": (C.c_int, vp, C.c_int, vp, i64, i64),
    "b200s_chol_spsolve": (C.c_int, vp, C.c_int, i64, i64, p_i64, p_i64, p_f64,
                           C.POINTER(p_i64), C.POINTER(p_i64), C.POINTER(p_f64)),
    "b200s_chol_diag": (C.c_int, vp, p_f64),
    "b200s_chol_analyze_z": (C.c_int, i64, p_i64, p_i64, C.c_char, p_i64, C.POINTER(CholOpts), C.POINTER(vp)),
    "b200s_chol_factorize_z": (C.c_int, vp, p_i64, p_i64, p_f64, p_i64),
    "b200s_chol_spsolve_z": (C.c_int, vp, C.c_int, i64, i64, p_i64, p_i64, p_f64, C.POINTER(p_i64), C.POINTER(p_i64), C.POINTER(p_f64)),
    "b200s_chol_diag_z": (C.c_int, vp, p_f64),
    "b200s_chol_get_L_z": (C.c_int, vp, C.POINTER(p_i64), C.POINTER(p_i64), C.POINTER(p_f64)),
    "b200s_chol_get_L": (C.c_int, vp, C.POINTER(p_i64), C.POINTER(p_i64), C.POINTER(p_f64)),
    "b200s_chol_info": (C.c_int, vp, C.POINTER(CholInfo)),
    "b200s_chol_set_profiling": (C.c_int, vp, C.c_int),
    "b200s_chol_get_perm": (C.c_int, vp, p_i64),
    "b200s_chol_get_super": (C.c_int, vp, p_i64, p_i64, p_i64),
    "b200s_chol_set_owned": (C.c_int, vp, C.c_char_p),
    "b200s_chol_factor_begin": (C.c_int, vp, vp, C.c_int),
    "b200s_chol_factor_level": (C.c_int, vp, i64),
    "b200s_chol_factor_level_phase": (C.c_int, vp, i64, C.c_int),
    "b200s_chol_solve_dist_begin": (C.c_int, vp, vp),
    "b200s_chol_solve_dist_level": (C.c_int, vp, C.c_int, i64),
    "b200s_chol_solve_dist_end": (C.c_int, vp, vp),
    "b200s_chol_solve_buffers": (C.c_int, vp, C.POINTER(vp), C.POINTER(vp)),
    "b200s_chol_front_layout2": (C.c_int, vp, p_i64, p_i64),
    "b200s_chol_set_syrk_split": (C.c_int, vp, C.c_char_p, C.POINTER(C.c_int32), C.POINTER(C.c_int32), p_i64, vp),
    "b200s_chol_factor_end": (C.c_int, vp, p_i64),
    "b200s_chol_sync": (C.c_int, vp),
    "b200s_chol_front_layout": (C.c_int, vp, p_i64, p_i64, p_i64, p_i64, p_i64, p_i64, p_i64, p_i64),
    "b200s_chol_device_buffers": (C.c_int, vp, C.POINTER(vp), C.POINTER(vp)),
    "b200s_chol_set_numeric": (C.c_int, vp, C.c_int, i64),
    "b200s_chol_free": (None, vp),
    "b200s_free": (None, vp),
    "b200s_grid_nd_perm": (C.c_int, i64, i64, i64, i64, p_i64),
    "b200s_persist_schedule_check": (C.c_int, i64, i64, i64),
    "b200s_chol_child_lists_check": (C.c_int, vp),
    "b200s_chol_set_solve_sweeps": (C.c_int, vp, C.c_int),
    "b200s_amd_order": (C.c_int, i64, p_i64, p_i64, C.c_char, p_i64),
    "b200s_klu_analyze": (C.c_int, i64, p_i64, p_i64, C.POINTER(vp)),
    "b200s_klu_factor": (C.c_int, vp, p_i64, p_i64, p_f64, C.POINTER(vp)),
    "b200s_klu_pivot_host": (C.c_int, vp, p_i64, p_i64, p_f64, C.POINTER(vp)),
    "b200s_klu_extract_host": (C.c_int, vp, p_f64, p_f64, p_f64, p_f64),
    "b200s_klu_refactor_batch": (C.c_int, vp, p_f64, i64, i64, p_int),
    "b200s_klu_refactor_batch_dev": (C.c_int, vp, vp, i64, i64, p_int),
    "b200s_klu_refactor_batch_begin": (C.c_int, vp, p_f64, i64, i64),
    "b200s_klu_refactor_batch_end": (C.c_int, vp, p_int),
    "b200s_klu_solve_batch": (C.c_int, vp, C.c_int, p_f64, i64, i64, i64),
    "b200s_klu_solve_batch_dev": (C.c_int, vp, C.c_int, vp, i64, i64, i64),
    "b200s_klu_solve": (C.c_int, vp, C.c_int, p_f64, i64, i64),
    "b200s_klu_info": (C.c_int, vp, C.POINTER(KluInfo)),
    "b200s_klu_extract": (C.c_int, vp, p_i64, p_i64, p_f64, p_i64, p_i64, p_f64, p_i64, p_i64, p_f64,
                          p_i64, p_i64, p_f64, p_i64),
    "b200s_klu_extract_batch": (C.c_int, vp, i64, p_f64, p_f64, p_f64, p_f64),
    "b200s_klu_plan_view": (C.c_int, vp, C.POINTER(KluPlanView)),
    "b200s_klu_symbolic_perm": (C.c_int, vp, p_i64, p_i64, p_i64, p_i64),
    "b200s_klu_factor_z": (C.c_int, vp, p_i64, p_i64, p_f64, C.POINTER(vp)),
    "b200s_klu_solve_z": (C.c_int, vp, C.c_int, p_f64, i64, i64),
    "b200s_klu_extract_z": (C.c_int, vp, p_i64, p_i64, p_f64, p_i64, p_i64, p_f64, p_i64, p_i64, p_f64, p_i64, p_i64, p_f64, p_i64),
    "b200s_klu_embedded": (vp, vp),
    "b200s_klu_plan_emulate_host": (C.c_int, vp, p_f64, p_f64, p_f64, p_f64, p_f64),
    "b200s_klu_solve_tape_host": (C.c_int, vp, C.c_int, p_f64, i64, i64),
    "b200s_kkt_create": (C.c_int, i64, i64, i64, p_i64, p_i64, p_f64, p_i64, p_i64, p_f64, p_i64, p_i64, C.POINTER(vp)),
    "b200s_kkt_set_singular": (C.c_int, vp, C.c_int),
    "b200s_kkt_factor": (C.c_int, vp, p_f64, p_f64, p_i64),
    "b200s_kkt_solve": (C.c_int, vp, p_f64, p_f64, p_f64),
    "b200s_kkt_info": (C.c_int, vp, C.POINTER(KktInfo)),
    "b200s_kkt_plan_check_host": (C.c_int, vp, p_f64, p_f64, p_f64),
    "b200s_kkt_free": (None, vp),
    "b200s_spmv_create": (C.c_int, i64, i64, p_i64, p_i64, p_f64, C.POINTER(vp)),
    "b200s_spmv_apply": (C.c_int, vp, p_f64, C.POINTER(vp)),
    "b200s_spmv_get": (C.c_int, vp, p_f64),
    "b200s_spmv_free": (None, vp),
    "b200s_kktd_create": (C.c_int, i64, i64, i64, p_f64, p_f64, C.POINTER(vp)),
    "b200s_kktd_factor": (C.c_int, vp, p_f64, p_f64, p_i64),
    "b200s_kktd_solve": (C.c_int, vp, p_f64, p_f64, p_f64),
    "b200s_kktd_info": (C.c_int, vp, C.POINTER(KktdInfo)),
    "b200s_kktd_get_qr": (C.c_int, vp, p_f64, p_f64, p_f64),
    "b200s_kktd_free": (None, vp),
    "b200s_klu_free_symbolic": (None, vp),
    "b200s_klu_free_numeric": (None, vp),
}

fn = {}
_missing = []
for _name, _s in SIGNATURES.items():
    try:
        fn[_name] = _sig(_name, _s[0], *_s[1:])
    except AttributeError:
        _missing.append(_name)
if _missing:
    raise ImportError("libb200sparse.so does not export: " + ", ".join(_missing))


def as_i64(a):
    return np.ascontiguousarray(a, dtype=np.int64)


def as_f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def ptr_i64(a):
    return a.ctypes.data_as(p_i64) if a is not None else None


def ptr_f64(a):
    return a.ctypes.data_as(p_f64) if a is not None else None


def strerror(st):
    return fn["b200s_strerror"](st).decode()


def last_error():
    return fn["b200s_last_error"]().decode()


def device_count():
    return fn["b200s_device_count"]()


def take_array(ptr, count, dtype):
    """copy a library-allocated array into numpy and release it"""
    if count == 0:
        out = np.zeros(0, dtype=dtype)
    else:
        out = np.ctypeslib.as_array(ptr, shape=(count,)).astype(dtype, copy=True)
    fn["b200s_free"](C.cast(ptr, vp))
    return out
