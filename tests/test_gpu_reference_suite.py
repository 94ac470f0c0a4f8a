"""The reference's OWN acceptance tests, executed verbatim from the copy oracle/build_ref.sh places
under oracle/_ref/tests (git-ignored build output; the files are not part of this repository) against the B200 modules
registered as kvxopt.klu / kvxopt.cholmod:
  * tests/test_sparse_solvers.py::TestKLU  (reference tests/test_sparse_solvers.py:216-323) -- the real ('d') cases; the
    complex half of each loop is dropped by replacing the module-level `product` (complex KLU is SURVEY section 8f-4 and
    the engine rejects 'z' input with TypeError, checked here as well);
  * tests/test_examples.py: the doc examples that reach cholmod through the IPM (test_ch9_acent: solvers.cp ->
    misc.kkt_chol2 -> cholmod.solve with 50 right-hand sides; test_ch8_lp; test_ch10_lp ...)."""
import importlib.util
import os
import unittest

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu
REF_TESTS = os.path.join(ROOT, "oracle", "_ref", "tests")


def load_ref_module(name):
    path = os.path.join(REF_TESTS, name + ".py")
    if not os.path.exists(path):
        pytest.fail("oracle/_ref/tests/%s.py is missing: run oracle/build_ref.sh in the build container" % name)
    spec = importlib.util.spec_from_file_location("ref_" + name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def run_case(cls, method):
    suite = unittest.TestSuite([cls(method)])
    res = unittest.TestResult()
    suite.run(res)
    return res


@pytest.mark.parametrize("method", ["test_lu", "test_linsolve", "test_solve"])
def test_reference_TestKLU_real_cases(kvx, method):
    mod = load_ref_module("test_sparse_solvers")
    # the reference loops over product(cases, [True, False]) with True = complex: keep the real half
    mod.product = lambda cases, flags: ((c, f) for c in cases for f in flags if not f)
    res = run_case(mod.TestKLU, method)
    assert res.testsRun == 1 and not res.skipped, res.skipped
    assert not res.failures and not res.errors, (res.failures, res.errors)


def test_reference_TestKLU_get_det(kvx):
    """the real determinant (= 114) passes; the complex half of the same reference test is refused with TypeError"""
    mod = load_ref_module("test_sparse_solvers")
    res = run_case(mod.TestKLU, "test_get_det")
    assert res.testsRun == 1 and not res.failures
    assert len(res.errors) == 1 and "TypeError" in res.errors[0][1] and "Ac" in res.errors[0][1], res.errors


def test_reference_TestKLU_complex_is_refused_not_wrong(kvx):
    mod = load_ref_module("test_sparse_solvers")
    mod.product = lambda cases, flags: ((c, f) for c in cases[:1] for f in flags if f)
    res = run_case(mod.TestKLU, "test_lu")
    assert len(res.errors) == 1 and "TypeError" in res.errors[0][1]


@pytest.mark.parametrize("method", ["test_ch9_acent", "test_ch8_lp", "test_ch8_coneqp", "test_ch10_lp", "test_ch9_acent2", "test_ch9_l2ac"])
def test_reference_doc_examples(kvx, method):
    mod = load_ref_module("test_examples")
    from kvxopt import cholmod
    calls = {"numeric": 0, "solve": 0}
    on, osv = cholmod.numeric, cholmod.solve

    def numeric(*a, **k):
        calls["numeric"] += 1
        return on(*a, **k)

    def solve(*a, **k):
        calls["solve"] += 1
        return osv(*a, **k)
    cholmod.numeric, cholmod.solve = numeric, solve
    try:
        res = run_case(mod.TestExamples, method)
    finally:
        cholmod.numeric, cholmod.solve = on, osv
    assert res.testsRun == 1 and not res.skipped
    assert not res.failures and not res.errors, (res.failures, res.errors)
    if method == "test_ch9_acent":
        assert calls["numeric"] > 0 and calls["solve"] > 0          # the cholmod path really ran
