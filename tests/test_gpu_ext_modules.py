"""`kvxopt.cholmod` and `kvxopt.klu` as COMPILED extension modules -- the reference's
own wrapper sources (src/C/cholmod.c, src/C/klu.c) built against include/suitesparse_shim/*.h and linked to
libb200sparse.so (tools/build_kvxopt_ext.sh) -- exercised in a fresh interpreter with an unpatched kvxopt package:
documented known answers (doc/source/spsolvers.rst:333-345, 555-585, 759-772), the reference's TestKLU and doc examples
verbatim, error contract, and boeing2 through misc.kkt_chol2 (29 iterations, objective to 1e-8)."""
import json
import os
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ext():
    p = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "ext_runner.py")], capture_output=True, text=True, timeout=900)
    for ln in reversed(p.stdout.splitlines()):
        if ln.startswith("EXTRUNNER "):
            return json.loads(ln[len("EXTRUNNER "):])
    pytest.fail("ext_runner.py failed (rc %d):\n%s\n%s" % (p.returncode, p.stdout[-2000:], p.stderr[-3000:]))


def test_modules_are_the_compiled_reference_wrappers(ext):
    assert ext["compiled"], (ext["cholmod_file"], ext["klu_file"])
    assert any("libb200sparse" in s for s in ext["native"])


def test_cholmod_documented_answers(ext):
    assert ext["linsolve_err"] < 1e-12 and ext["solve_err"] < 1e-12 and ext["splinsolve_err"] < 1e-12 and ext["spsolve_err"] < 1e-12
    assert abs(ext["logdet_diag"] - 5.505331535932363) < 1e-12          # spsolvers.rst:759-772, through diag()
    assert abs(ext["logdet_ldl"] - 5.505331535932363) < 1e-12           # ... and through sys = 6 with supernodal = 0
    assert ext["diag_ldl_refused"] and ext["getfactor_nnz"] >= 6


def test_cholmod_complex_hermitian_through_the_compiled_wrapper(ext):
    assert ext["cholmod_z_linsolve_err"] < 1e-12 and ext["cholmod_z_diag_err"] < 1e-12 and ext["cholmod_z_typecode"] == "z"


def test_cholmod_error_contract(ext):
    assert ext["npd"].startswith("ArithmeticError(")                     # documented: numeric raises with the failing column
    assert ext["npd_solve"] == "singular matrix"                         # cholmod.c:456
    assert ext["symbolic_solve"] == "called with symbolic factor"        # cholmod.c:452-453


def test_klu_documented_answers(ext):
    assert ext["klu_linsolve_err"] < 1e-12
    assert abs(ext["klu_det"] - 114.0) < 1e-9
    assert ext["klu_identity"] < 1e-12 and ext["klu_tsolve_res"] < 1e-12
    assert ext["klu_singular"] == "singular matrix"                      # klu.c:370-371


@pytest.mark.parametrize("name", ["test_lu", "test_linsolve", "test_solve", "test_get_det", "test_ch9_acent", "test_ch8_lp", "test_ch8_coneqp"])
def test_reference_suite_verbatim_on_compiled_modules(ext, name):
    assert ext["ref_tests"][name]["ok"], ext["ref_tests"][name]["detail"]


def test_boeing2_through_compiled_cholmod(ext):
    b = ext["boeing2"]
    assert b["status"] == "optimal" and b["iterations"] == 29
    assert abs(b["objective"] - (-315.0187296452)) <= 1e-8 * 315.02
