"""The reference's own interior-point code (probe build oracle/_ref: coneprog.py + misc.py, unmodified) with the
B200 cholmod module plugged in as kvxopt.cholmod: identical iteration counts, objectives within 1e-8
(BASELINE.json north_star) against the reference's dense LAPACK KKT solvers (tests/golden)."""
import numpy as np
import pytest
import scipy.sparse as sp

from conftest import GOLD

pytestmark = pytest.mark.gpu


def to_spmatrix(kvx, M):
    M = sp.coo_matrix(M)
    return kvx.spmatrix(M.data.tolist(), M.row.tolist(), M.col.tolist(), M.shape)


def test_boeing2_lp_chol2_sparse_through_cuda_cholmod(kvx):
    """BASELINE config 3 (default sparse 'chol2' => misc.kkt_chol2 => cholmod.symbolic/numeric/solve/spsolve)"""
    from kvxopt import matrix, spmatrix, solvers
    z = np.load(GOLD + "/boeing2_lp.npz")
    G = sp.csc_matrix((z["Gx"], z["Gi"], z["Gp"]), shape=tuple(z["G_size"]))
    A = sp.csc_matrix((z["Ax"], z["Ai"], z["Ap"]), shape=tuple(z["A_size"]))
    c, h, b = matrix(z["c"]), matrix(z["h"]), matrix(z["b"])
    calls = {"numeric": 0, "solve": 0}
    from kvxopt import cholmod
    orig_numeric, orig_solve = cholmod.numeric, cholmod.solve
    def numeric(*a, **k):
        calls["numeric"] += 1
        return orig_numeric(*a, **k)
    def solve(*a, **k):
        calls["solve"] += 1
        return orig_solve(*a, **k)
    cholmod.numeric, cholmod.solve = numeric, solve
    try:
        sol = solvers.lp(c, to_spmatrix(kvx, G), h, to_spmatrix(kvx, A), b)
    finally:
        cholmod.numeric, cholmod.solve = orig_numeric, orig_solve
    assert sol["status"] == "optimal"
    assert sol["iterations"] == int(z["iters_chol2"]) == 29
    assert abs(sol["primal objective"] - float(z["pobj_chol2"])) <= 1e-8 * abs(float(z["pobj_chol2"]))
    assert calls["numeric"] >= 2 * 29 and calls["solve"] >= 5 * 29      # the GPU path really ran
    np.testing.assert_allclose(np.array(sol["x"]).ravel(), z["x_ref"], rtol=1e-5, atol=1e-6)


def test_small_sparse_lp_reference_test(kvx):
    """reference tests/test_osqp.py:33-44 (the LP part: sparse G => kkt_chol2 => cholmod): optimal, x = [1, 1]"""
    from kvxopt import matrix, spmatrix, solvers
    c = matrix([-4.0, -5.0])
    G = spmatrix([2.0, 1.0, -1.0, 1.0, 2.0, -1.0], [0, 1, 2, 0, 1, 3], [0, 0, 0, 1, 1, 1])
    h = matrix([3.0, 3.0, 0.0, 0.0])
    sol = solvers.lp(c, G, h)
    assert sol["status"] == "optimal"
    np.testing.assert_allclose(np.array(sol["x"]).ravel(), [1.0, 1.0], atol=1e-6)


def test_qp_mini_coneqp(kvx):
    """BASELINE config 5 generator at reduced size through coneqp + kkt_chol2 (sparse P, G; no equalities, so
    the 0x0 / n x 0 paths of misc.py:1483-1487,1545 are exercised)"""
    import sys
    sys.path.insert(0, GOLD)
    from generators import qp_instance
    z = np.load(GOLD + "/qp_mini.npz")
    P, q, G, h = qp_instance(int(z["nx"]), int(z["ny"]), int(z["nrand"]))
    from kvxopt import matrix, solvers
    sol = solvers.qp(to_spmatrix(kvx, sp.tril(P)), matrix(q), to_spmatrix(kvx, G), matrix(h))
    assert sol["status"] == "optimal"
    assert sol["iterations"] == int(z["iters"])
    assert abs(sol["primal objective"] - float(z["pobj"])) <= 1e-8 * abs(float(z["pobj"]))
