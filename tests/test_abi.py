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


def test_compiled_reference_wrappers_load_and_refuse_to_run_without_a_gpu():
    """kvxopt.cholmod / kvxopt.klu as compiled extension modules (the reference's src/C/cholmod.c and klu.c against
    include/suitesparse_shim, tools/build_kvxopt_ext.sh): they import in a fresh interpreter, expose the reference's
    function table, do the host-side symbolic work, and every numeric call fails loudly when there is no GPU."""
    import subprocess
    import sys
    ref = os.path.join(ROOT, "oracle", "_ref")
    if not any(f.startswith("cholmod.") and f.endswith(".so") for f in os.listdir(os.path.join(ref, "kvxopt"))):
        pytest.fail("oracle/_ref/kvxopt/cholmod*.so is missing: run tools/build_kvxopt_ext.sh in the build container")
    code = r'''
import sys, json
sys.path.insert(0, %r)
from kvxopt import cholmod, klu, spmatrix, matrix
from kvxopt_b200 import _lib
out = {"compiled": cholmod.__file__.endswith(".so") and klu.__file__.endswith(".so"),
       "funcs": all(hasattr(cholmod, f) for f in ("options", "symbolic", "numeric", "solve", "spsolve", "linsolve", "splinsolve", "diag", "getfactor"))
                and all(hasattr(klu, f) for f in ("linsolve", "symbolic", "numeric", "solve", "get_numeric", "get_det"))}
A = spmatrix([10, 3, 5, -2, 5, 2], [0, 2, 1, 3, 2, 3], [0, 0, 1, 1, 2, 3])
F = cholmod.symbolic(A)
out["capsule"] = repr(F)
Fs = klu.symbolic(A)
if _lib.device_count() == 0:
    for name, call in (("cholmod.numeric", lambda: cholmod.numeric(A, F)), ("klu.numeric", lambda: klu.numeric(A, Fs)),
                       ("cholmod.linsolve", lambda: cholmod.linsolve(A, matrix(1.0, (4, 1))))):
        try:
            call(); out[name] = "ran"
        except (ValueError, ArithmeticError) as e:
            out[name] = "refused"
print("OUT " + json.dumps(out))
''' % ref
    p = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert p.returncode == 0, p.stderr[-2000:]
    import json
    out = json.loads([l for l in p.stdout.splitlines() if l.startswith("OUT ")][-1][4:])
    assert out["compiled"] and out["funcs"] and "CHOLMOD SYM D FACTOR L" in out["capsule"]
    for k in ("cholmod.numeric", "klu.numeric", "cholmod.linsolve"):
        assert out.get(k, "refused") == "refused"          # no silent CPU path behind the compiled modules
