"""Pins the CPU oracle (oracle/) to the reference: doc known answers (reference doc/source/spsolvers.rst),
LAPACK solutions computed by the reference's own lapack.posv (tests/golden/posv_*.npz), the determinant
of reference tests/test_sparse_solvers.py:298-313, and dense numpy factorizations."""
import os
import sys

import numpy as np
import pytest
import scipy.sparse as sp
import scipy.sparse.linalg as spla

from conftest import GOLD, load_matrix, lower_ccs, rand_spd, sym_from_lower
from oracle import CholOracle, KluOracle

# reference doc/source/spsolvers.rst:555-563 (cholmod.linsolve example)
DOC_A = sp.csc_matrix(([10.0, 3, 5, -2, 5, 2], ([0, 2, 1, 3, 2, 3], [0, 0, 1, 1, 2, 3])), shape=(4, 4))
DOC_X = np.array([[-0.146341463414634, 0.048780487804878], [1.333333333333333, 4.0],
                  [0.487804878048781, 1.170731707317073], [2.833333333333333, 7.5]])


def test_chol_doc_linsolve_known_answer():
    A = DOC_A.copy(); A.sort_indices()
    B = np.arange(8, dtype=float).reshape(4, 2, order="F")
    O = CholOracle(4, A.indptr, A.indices, "L")
    O.factorize(A.data)
    np.testing.assert_allclose(O.solve(B), DOC_X, rtol=1e-13, atol=1e-14)
    # spsolvers.rst:759-772: log det via diag = 5.50533153593236
    assert abs(2.0 * np.sum(np.log(O.diag())) - 5.505331535932363) < 1e-13


@pytest.mark.parametrize("name", ["bcsstk13", "bcsstk24"])
def test_chol_matches_reference_lapack_posv(name):
    Al = load_matrix(name)
    n = Al.shape[0]
    B = np.random.default_rng(0).standard_normal((n, 3))
    Xref = np.load(GOLD + "/posv_%s.npz" % name)["X"]
    O = CholOracle(n, Al.indptr, Al.indices, "L")            # natural order, like a dense LAPACK solve
    O.factorize(Al.data)
    X = O.solve(B)
    assert np.linalg.norm(X - Xref) / np.linalg.norm(Xref) < 1e-10
    A = sym_from_lower(Al)
    berr = np.linalg.norm(A @ X - B, axis=0) / (spla.norm(A, 1) * np.linalg.norm(X, axis=0) + np.linalg.norm(B, axis=0))
    assert berr.max() < 1e-12


@pytest.mark.parametrize("n,dens,seed", [(1, 1.0, 0), (7, 0.5, 1), (60, 0.1, 2), (400, 0.02, 3)])
def test_chol_against_dense(n, dens, seed):
    A = rand_spd(n, dens, seed)
    Al = lower_ccs(A)
    rng = np.random.default_rng(seed)
    perm = rng.permutation(n)
    O = CholOracle(n, Al.indptr, Al.indices, "L", perm)
    O.factorize(Al.data)
    Ad = A.toarray()
    p = O.perm()
    assert np.array_equal(p, perm)
    Lref = np.linalg.cholesky(Ad[np.ix_(p, p)])
    np.testing.assert_allclose(O.dense_L(), Lref, atol=1e-12 * np.abs(Lref).max())
    b = rng.standard_normal((n, 2))
    np.testing.assert_allclose(O.solve(b, 4), np.linalg.solve(Lref, b), rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(O.solve(b, 5), np.linalg.solve(Lref.T, b), rtol=1e-9, atol=1e-12)
    np.testing.assert_array_equal(O.solve(b, 7), b[p])
    x8 = np.zeros_like(b); x8[p] = b
    np.testing.assert_array_equal(O.solve(b, 8), x8)
    np.testing.assert_array_equal(O.solve(b, 6), b)
    # upper-triangle input gives the same factorization
    Au = sp.triu(A).tocsc(); Au.sort_indices()
    O2 = CholOracle(n, Au.indptr, Au.indices, "U", perm)
    O2.factorize(Au.data)
    np.testing.assert_allclose(O2.dense_L(), O.dense_L(), atol=1e-13 * np.abs(Lref).max())


def test_chol_not_positive_definite_reports_column():
    A = sp.csc_matrix(np.array([[4.0, 1, 0], [1, 0.1, 0], [0, 0, -1]]))
    Al = lower_ccs(A)
    O = CholOracle(3, Al.indptr, Al.indices)
    with pytest.raises(ArithmeticError) as e:
        O.factorize(Al.data)
    assert e.value.args[0] == 1


def test_chol_pack_ignores_other_triangle():
    """pack (reference cholmod.c:132-181) reads only the uplo triangle"""
    A = rand_spd(30, 0.2, 5)
    junk = A.copy().tolil()
    junk[0, 29] = 1e9                 # upper-triangle garbage must be ignored with uplo='L'
    junk = junk.tocsc(); junk.sort_indices()
    O = CholOracle(30, junk.indptr, junk.indices, "L")
    O.factorize(junk.data)
    b = np.ones(30)
    np.testing.assert_allclose(A @ O.solve(b), b, atol=1e-10)


def test_chol_zero_size():
    O = CholOracle(0, np.zeros(1, dtype=np.int64), np.zeros(0, dtype=np.int64))
    O.factorize(np.zeros(0))
    assert O.solve(np.zeros((0, 1))).shape == (0, 1)


# reference doc/source/spsolvers.rst:333-345 klu.linsolve example and :420-439
KLU_V = [2, 3, 3, -1, 4, 4, -3, 1, 2, 2, 6, 1]
KLU_I = [0, 1, 0, 2, 4, 1, 2, 3, 4, 2, 1, 4]
KLU_J = [0, 0, 1, 1, 1, 2, 2, 2, 2, 3, 4, 4]


def klu_doc_matrix():
    A = sp.csc_matrix((np.array(KLU_V, dtype=float), (KLU_I, KLU_J)), shape=(5, 5))
    A.sort_indices()
    return A


def test_klu_doc_known_answers():
    A = klu_doc_matrix()
    O = KluOracle(5, A.indptr, A.indices, A.data)
    # spsolvers.rst:333-345: B = [1,...,1]... the documented solve uses B = A^T-style chain; check x = A^-T B^-1 A^-1 1
    B = sp.csc_matrix((np.array([4, 3, 3, -1, 4, 4, -3, 1, 2, 2, 6, 2], dtype=float), (KLU_I, KLU_J)), shape=(5, 5)).tocsc()
    B.sort_indices()
    OB = KluOracle(5, B.indptr, B.indices, B.data)
    x = O.solve(np.ones(5))
    x = OB.solve(x)
    x = O.solve(x, "T")
    np.testing.assert_allclose(x, [0.580654371385528, -0.236595065688228, 1.628000923361034, 8.06557280782751,
                                   -0.13075278223259], rtol=1e-12)
    # reference tests/test_sparse_solvers.py:298-313
    assert abs(O.det() - 114.0) < 1e-10


@pytest.mark.parametrize("name", ["bp_800", "bcsstk13", "ACTIVSg2000"])
def test_klu_oracle_vs_superlu(name):
    A = load_matrix(name)
    n = A.shape[0]
    # a fill-reducing column order from SuperLU keeps the unordered oracle fast enough
    lu = spla.splu(A.tocsc())
    O = KluOracle(n, A.indptr, A.indices, A.data, P0=None, Q=lu.perm_c.argsort() if False else None) if n < 1000 else \
        KluOracle(n, A.indptr, A.indices, A.data, Q=np.argsort(lu.perm_c))
    B = np.random.default_rng(0).standard_normal((n, 3))
    for trans, M in (("N", A), ("T", A.T)):
        X = O.solve(B, trans)
        Xref = spla.splu(M.tocsc()).solve(B)
        assert np.linalg.norm(X - Xref) / np.linalg.norm(Xref) < 1e-9
        assert np.abs(M @ X - B).max() < 1e-7          # the reference's own criterion (7 places)
    v2 = A.data * (1 + 1e-3 * np.random.default_rng(1).uniform(-1, 1, A.nnz))
    O.refactor(v2)
    A2 = sp.csc_matrix((v2, A.indices, A.indptr), shape=A.shape)
    X = O.solve(B)
    assert np.abs(A2 @ X - B).max() < 1e-7


def test_klu_singular():
    A = sp.csc_matrix(np.array([[1.0, 2], [2, 4]]))
    with pytest.raises(ArithmeticError):
        KluOracle(2, A.indptr, A.indices, A.data)


def test_ldl_oracle_against_numpy_and_reference_sytrf():
    """oracle/ldl_oracle.py (LDL' without pivoting, misc.kkt_ldl for componentwise cones): L D L' = P A P', inertia,
    numpy solutions, and the UNMODIFIED reference's sytrf/sytrs KKT answers (tests/golden/kkt_ldl_ref.npz)."""
    from oracle import ldl_oracle
    sys.path.insert(0, GOLD)
    from generators import qp_instance
    rng = np.random.default_rng(5)
    E = rng.standard_normal((30, 30)); E = E @ E.T + 30 * np.eye(30)
    Fm = rng.standard_normal((20, 20)); Fm = Fm @ Fm.T + 20 * np.eye(20)
    B = rng.standard_normal((20, 30))
    K = np.block([[E, B.T], [B, -Fm]])
    perm = rng.permutation(50)
    Lo, d, minor = ldl_oracle.ldl_nopivot(K[np.ix_(perm, perm)])
    assert minor == 50 and int((d < 0).sum()) == 20
    assert np.abs(Lo @ np.diag(d) @ Lo.T - K[np.ix_(perm, perm)]).max() <= 1e-12 * np.abs(K).max()
    rhs = rng.standard_normal((50, 2))
    assert np.linalg.norm(ldl_oracle.solve_sys(Lo, d, perm, rhs, 0) - np.linalg.solve(K, rhs)) <= 1e-12 * np.linalg.norm(rhs)
    assert ldl_oracle.ldl_nopivot(np.array([[2.0, 1, 0], [1, 0.5, 1], [0, 1, 3.0]]))[2] == 1
    z = np.load(os.path.join(GOLD, "kkt_ldl_ref.npz"))
    b2 = np.load(os.path.join(GOLD, "boeing2_lp.npz"))
    G = sp.csc_matrix((b2["Gx"], b2["Gi"], b2["Gp"]), shape=tuple(b2["G_size"])).toarray()
    A = sp.csc_matrix((b2["Ax"], b2["Ai"], b2["Ap"]), shape=tuple(b2["A_size"])).toarray()
    ux, uy, uz = ldl_oracle.kkt_ldl_lcone(G, A, None, 1.0 / z["lp_d"], z["lp_bx"], z["lp_by"], z["lp_bz"])
    for got, want in ((ux, z["lp_ux"]), (uy, z["lp_uy"]), (uz, z["lp_uz"])):
        assert np.linalg.norm(got - want) <= 1e-10 * np.linalg.norm(want)
    P, _, Gq, _ = qp_instance(50, 40, 50)
    ux, uy, uz = ldl_oracle.kkt_ldl_lcone(Gq.toarray(), np.zeros((0, 2000)), P.toarray(), 1.0 / z["qp_d"], z["qp_bx"],
                                          z["qp_by"], z["qp_bz"])
    for got, want in ((ux, z["qp_ux"]), (uz, z["qp_uz"])):
        assert np.linalg.norm(got - want) <= 1e-10 * np.linalg.norm(want)
