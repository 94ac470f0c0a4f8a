"""The C-ABI library loads on a CPU-only box and exports every symbol include/b200sparse.h declares."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

from conftest import ROOT


def header_symbols():
    txt = open(os.path.join(ROOT, "include", "b200sparse.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(b200s_[A-Za-z0-9_]+)\s*\(", txt)))


def test_header_symbols_exported():
    from kvxopt_b200 import _lib
    syms = header_symbols()
    assert len(syms) >= 35
    out = subprocess.check_output(["nm", "-D", "--defined-only", _lib.LIB_PATH]).decode()
    exported = set(l.split()[-1] for l in out.splitlines() if " T " in l)
    missing = [s for s in syms if s not in exported]
    assert not missing, missing
    # every declared function has a ctypes signature and vice versa
    assert sorted(_lib.SIGNATURES) == syms


def test_library_is_sm100a():
    from kvxopt_b200 import _lib
    out = subprocess.run(["cuobjdump", "-lelf", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out


def test_status_strings_and_version():
    from kvxopt_b200 import _lib
    assert _lib.fn["b200s_version"]().decode().startswith("b200sparse")
    assert "positive definite" in _lib.strerror(1)
    assert "no CUDA device" in _lib.strerror(-5)


def test_numeric_refuses_without_gpu_or_runs_with_one():
    """no CPU fallback: on a box without a GPU the numeric entry points return B200S_NO_DEVICE"""
    from kvxopt_b200 import _lib
    import scipy.sparse as sp
    A = sp.csc_matrix(np.array([[4.0, 0], [1.0, 3.0]]))
    cp, ri, vx = A.indptr.astype(np.int64), A.indices.astype(np.int64), A.data.copy()
    F = _lib.vp()
    assert _lib.fn["b200s_chol_analyze"](2, _lib.ptr_i64(cp), _lib.ptr_i64(ri), b"L", None, None, C.byref(F)) == 0
    st = _lib.fn["b200s_chol_factorize"](F, None, None, _lib.ptr_f64(vx), None)
    if _lib.device_count() == 0:
        assert st == _lib.NO_DEVICE
    else:
        assert st == 0
    _lib.fn["b200s_chol_free"](F)


def test_python_mirror_raises_without_gpu():
    from kvxopt_b200 import _lib, cholmod, klu
    import scipy.sparse as sp
    if _lib.device_count() > 0:
        pytest.skip("GPU present")
    A = sp.csc_matrix(np.array([[4.0, 0], [1.0, 3.0]]))
    B = np.ones((2, 1), order="F")
    with pytest.raises(RuntimeError):
        cholmod.linsolve(A, B)
    with pytest.raises(RuntimeError):
        klu.linsolve(A, B)
