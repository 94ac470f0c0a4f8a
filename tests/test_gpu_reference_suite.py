"""The reference's OWN acceptance tests, executed verbatim from the copy oracle/build_ref.sh places
under oracle/_ref/tests (git-ignored build output; the files are not part of this repository) against the B200 modules
registered as kvxopt.klu / kvxopt.cholmod:
  * tests/test_sparse_solvers.py::TestKLU  (reference tests/test_sparse_solvers.py:216-323) -- all four methods, real AND
    complex cases, exactly as the reference loops over them;
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


@pytest.mark.parametrize("method", ["test_lu", "test_linsolve", "test_solve", "test_get_det"])
def test_reference_TestKLU_verbatim(kvx, method):
    """the reference's own loops: product(cases, [True, False]) with True = complex ('z') matrices, trans in N, T, C"""
    mod = load_ref_module("test_sparse_solvers")
    res = run_case(mod.TestKLU, method)
    assert res.testsRun == 1 and not res.skipped, res.skipped
    assert not res.failures and not res.errors, (res.failures, res.errors)


def test_reference_TestKLU_complex_cases_really_run(kvx):
    """guard against a silently shortened loop: count the complex factorizations the reference test makes"""
    mod = load_ref_module("test_sparse_solvers")
    from kvxopt import klu
    seen = {"z": 0, "d": 0}
    orig = klu.numeric

    def numeric(A, Fs):
        seen[A.typecode] += 1
        return orig(A, Fs)
    klu.numeric = numeric
    try:
        res = run_case(mod.TestKLU, "test_lu")
    finally:
        klu.numeric = orig
    assert not res.failures and not res.errors, (res.failures, res.errors)
    assert seen["z"] >= 1 and seen["z"] == seen["d"]


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
