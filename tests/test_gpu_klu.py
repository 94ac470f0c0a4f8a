"""Parity of the CUDA KLU path against the reference's own test criteria (reference
tests/test_sparse_solvers.py:216-323), its documented known answers, the CPU oracle and SuperLU."""
import numpy as np
import pytest
import scipy.sparse as sp
import scipy.sparse.linalg as spla

from conftest import load_matrix

pytestmark = pytest.mark.gpu
MATRICES = ["ACTIVSg2000", "bcsstk13", "bcsstk24", "bp_800"]      # reference tests/test_sparse_solvers.py:29-30

KLU_V = [2, 3, 3, -1, 4, 4, -3, 1, 2, 2, 6, 1]
KLU_I = [0, 1, 0, 2, 4, 1, 2, 3, 4, 2, 1, 4]
KLU_J = [0, 0, 1, 1, 1, 2, 2, 2, 2, 3, 4, 4]


@pytest.fixture(scope="module")
def klu():
    from kvxopt_b200 import klu as m, _lib
    assert _lib.device_count() > 0, "GPU tests need a CUDA device; there is no CPU fallback"
    return m


@pytest.mark.parametrize("name", MATRICES)
def test_lu_identity(klu, name):
    """TestKLU.test_lu: || R P A Q - (L U + F) ||_1 == 0 to 7 places"""
    A = load_matrix(name)
    Fs = klu.symbolic(A)
    Fn = klu.numeric(A, Fs)
    Lm, Um, P, Q, R, Fm, r = klu.get_numeric(A, Fs, Fn)
    err = abs(R @ P @ A @ Q - (Lm @ Um + Fm)).sum(axis=0).max()
    assert round(err, 7) == 0
    assert r[0] == 0 and r[-1] == A.shape[0]


@pytest.mark.parametrize("name", MATRICES)
@pytest.mark.parametrize("trans", ["N", "T"])
def test_linsolve_and_solve(klu, name, trans):
    """TestKLU.test_linsolve / test_solve: A x == b elementwise to 7 places, 3 random columns"""
    A = load_matrix(name)
    n = A.shape[0]
    B = np.random.default_rng(0).standard_normal((n, 3))
    M = A if trans == "N" else A.T
    X = np.asfortranarray(B.copy())
    assert klu.linsolve(A, X, trans=trans) is None
    np.testing.assert_array_almost_equal(M @ X, B, decimal=7)
    Fs = klu.symbolic(A); Fn = klu.numeric(A, Fs)
    X2 = np.asfortranarray(B.copy())
    klu.solve(A, Fs, Fn, X2, trans=trans)
    assert np.array_equal(X, X2)
    Xref = spla.splu(M.tocsc()).solve(B)
    assert np.linalg.norm(X - Xref) / np.linalg.norm(Xref) < 1e-10


def test_doc_known_answers_and_det(klu, kvx):
    """reference doc/source/spsolvers.rst:333-345, 420-439 and tests/test_sparse_solvers.py:288-323, reference types"""
    from kvxopt import matrix, spmatrix, klu as kk
    assert kk is klu
    A = spmatrix(KLU_V, KLU_I, KLU_J)
    B = matrix(1.0, (5, 1))
    kk.linsolve(A, B)
    np.testing.assert_allclose(np.array(B).ravel(), [0.57894737, -0.05263158, 1.0, 1.97368421, -0.78947368], rtol=1e-7)
    VB = [4, 3, 3, -1, 4, 4, -3, 1, 2, 2, 6, 2]
    Bm = spmatrix(VB, KLU_I, KLU_J)
    x = matrix(1.0, (5, 1))
    Fa = kk.symbolic(A); FA = kk.numeric(A, Fa)
    Fb = kk.symbolic(Bm); FB = kk.numeric(Bm, Fb)
    kk.solve(A, Fa, FA, x)
    kk.solve(Bm, Fb, FB, x)
    kk.solve(A, Fa, FA, x, trans="T")
    np.testing.assert_allclose(np.array(x).ravel(), [0.580654371385528, -0.236595065688228, 1.628000923361034,
                                                     8.06557280782751, -0.13075278223259], rtol=1e-12)
    assert abs(kk.get_det(A, Fa, FA) - 114.0) < 1e-9
    Lm, Um, P, Q, R, Fm, r = kk.get_numeric(A, Fa, FA)
    assert type(Lm).__name__ == "spmatrix" and isinstance(r, list)
    assert max(abs(R * P * A * Q - (Lm * Um + Fm))) < 1e-12          # spsolvers.rst:462-484
    assert kk.linsolve(spmatrix([], [], [], (0, 0)), matrix(0.0, (0, 1))) == 0      # klu.c:126


def test_singular_and_type_errors(klu):
    with pytest.raises(ArithmeticError, match="singular"):
        klu.linsolve(sp.csc_matrix(np.array([[1.0, 2], [2, 4]])), np.ones((2, 1), order="F"))
    with pytest.raises(TypeError):
        klu.symbolic(sp.csc_matrix(np.ones((2, 3))))
    A = load_matrix("bp_800")
    Fs = klu.symbolic(A)
    with pytest.raises(TypeError):
        klu.solve(A, Fs, Fs, np.ones((822, 1), order="F"))
    with pytest.raises(ValueError):
        klu.linsolve(A, np.ones((822, 1), order="F"), trans="X")


@pytest.mark.parametrize("name,batch", [("bp_800", 70), ("ACTIVSg2000", 96)])
def test_batched_refactor_vs_oracle(klu, name, batch):
    """BASELINE config 2 at test size: same-pattern value perturbations a_k (1 + 1e-3 u_k); every matrix of the
    batch is compared with the CPU oracle's refactorization (same pivot order) and with SuperLU"""
    from oracle import KluOracle
    A = load_matrix(name)
    n = A.shape[0]
    Fs = klu.symbolic(A); Fn = klu.numeric(A, Fs)
    rng = np.random.default_rng(0)
    vals = A.data[None, :] * (1 + 1e-3 * rng.uniform(-1, 1, size=(batch, A.nnz)))
    status = klu.refactor_batch(Fn, vals)
    assert not status.any()
    B = rng.standard_normal((batch, 2, n))
    X = B.copy()
    klu.solve_batch(Fn, X)
    Lm, Um, P, Q, R, Fm, r = klu.get_numeric(A, Fs, Fn)
    Pnum = np.asarray(P.tocsc().indices)                  # P[i, Pnum[i]] = 1 -> column of the single entry in row i
    Pnum = np.asarray(P.tocsr().indices)
    Qv = np.asarray(Q.tocsc().indices)
    O = KluOracle(n, A.indptr, A.indices, A.data, P0=Pnum, Q=Qv)
    for b in list(range(0, batch, max(1, batch // 6))) + [batch - 1]:
        Ab = sp.csc_matrix((vals[b], A.indices, A.indptr), shape=(n, n))
        O.refactor(vals[b])
        Xo = O.solve(B[b].T)
        assert np.linalg.norm(X[b].T - Xo) / np.linalg.norm(Xo) < 1e-10
        assert np.abs(Ab @ X[b].T - B[b].T).max() < 1e-7
    XT = B.copy()
    klu.solve_batch(Fn, XT, trans="T")
    b = batch // 2
    Ab = sp.csc_matrix((vals[b], A.indices, A.indptr), shape=(n, n))
    assert np.abs(Ab.T @ XT[b].T - B[b].T).max() < 1e-7
    # a matrix with a zeroed pivot column is flagged, the others are unaffected
    vals2 = vals.copy(); vals2[3, :] = 0.0
    st = klu.refactor_batch(Fn, vals2, check=False)
    assert st[3] != 0 and st[:3].sum() == 0 and st[4:].sum() == 0
    with pytest.raises(ArithmeticError):
        klu.refactor_batch(Fn, vals2)


def test_pipelined_refactor_matches_synchronous_call(klu):
    """refactor_batch_begin / refactor_batch_end (two batches in flight) leave the same factors and status as the
    synchronous refactor_batch, in order, also with a singular matrix in one of the batches."""
    A = load_matrix("bp_800")
    Fs = klu.symbolic(A)
    Fn = klu.numeric(A, Fs)
    rng = np.random.default_rng(11)
    n = A.shape[0]
    batches = [np.ascontiguousarray(A.data[None, :] * (1 + 1e-3 * rng.uniform(-1, 1, size=(bsz, A.nnz)))) for bsz in (70, 33, 70)]
    batches[1][5, :] = 0.0                       # singular matrix in the middle batch
    B = rng.standard_normal((70, 1, n))
    ref = []
    for v in batches:
        st = klu.refactor_batch(Fn, v, check=False)
        X = B[:v.shape[0]].copy()
        klu.solve_batch(Fn, X)
        ref.append((st.copy(), X))
    assert ref[1][0][5] != 0 and not ref[0][0].any()
    klu.refactor_batch_begin(Fn, batches[0])
    klu.refactor_batch_begin(Fn, batches[1])
    with pytest.raises(ValueError):
        klu.refactor_batch_begin(Fn, batches[2])           # at most two in flight
    st0 = klu.refactor_batch_end(Fn, check=False)
    st1 = klu.refactor_batch_end(Fn, check=False)
    assert (st0 == ref[0][0]).all() and (st1 == ref[1][0]).all()
    klu.refactor_batch_begin(Fn, batches[2])
    st2 = klu.refactor_batch_end(Fn)
    X = B.copy()
    klu.solve_batch(Fn, X)                                 # factors of the most recently begun batch
    assert (st2 == ref[2][0]).all() and np.array_equal(X, ref[2][1])
    with pytest.raises(ValueError):
        klu.refactor_batch_end(Fn)


def _variety(case):
    rng = np.random.default_rng(100 + case)
    if case == 0:                      # block upper triangular: several BTF blocks with off-diagonal (F) entries
        blocks = [sp.random(m, m, density=0.3, random_state=rng) + 3 * sp.identity(m) for m in (7, 1, 12, 30, 1, 5)]
        A = sp.block_diag(blocks).tolil()
        n = A.shape[0]
        for _ in range(60):
            i, j = sorted(rng.integers(0, n, 2))
            if i != j:
                A[i, j] = rng.standard_normal()
        perm = rng.permutation(n)
        return A.tocsc()[perm][:, rng.permutation(n)].tocsc()
    if case == 1:                      # fully dense: every column of the factor is longer than a wave's shared memory
        n = 480
        return sp.csc_matrix(rng.standard_normal((n, n)) + n * np.eye(n))
    if case == 2:                      # wide band + random fill: long dense tail, many waves
        n = 900
        return (sp.diags([rng.standard_normal(n - abs(k)) for k in range(-6, 7)], list(range(-6, 7))) +
                sp.random(n, n, density=0.001, random_state=rng) + 8 * sp.identity(n)).tocsc()
    if case == 3:                      # diagonal and 1 x 1
        return sp.csc_matrix(np.array([[2.5]]))
    if case == 4:                      # unsymmetric pattern, zero diagonal entries that need the row permutation
        n = 300
        A = sp.random(n, n, density=0.02, random_state=rng).tolil()
        p = rng.permutation(n)
        for i in range(n):
            A[i, p[i]] = 4 + rng.uniform()
        return A.tocsc()
    n = 64                             # arrow matrix: one dense row and column
    A = sp.identity(n).tolil() * 3
    A[0, :] = 1.0; A[:, 0] = 1.0; A[0, 0] = n
    return A.tocsc()


@pytest.mark.parametrize("early_minw", [None, 4])
@pytest.mark.parametrize("case", range(6))
def test_structure_variety_single_and_batched(klu, case, early_minw, monkeypatch):
    """patterns the reference matrices do not cover: many BTF blocks with F entries, columns longer than the shared-memory
    wave (level-schedule kernel), long dense tails, 1 x 1, permuted diagonals, arrow -- LU identity, both solves, and a
    perturbed batch against SuperLU"""
    if early_minw is not None:      # the level-by-level early-column kernel (k_klu_early) on these small patterns as well
        monkeypatch.setenv("B200S_KLU_EARLY_MINW", str(early_minw))
    A = _variety(case)
    A.sort_indices()
    n = A.shape[0]
    Fs = klu.symbolic(A); Fn = klu.numeric(A, Fs)
    Lm, Um, P, Q, R, Fm, r = klu.get_numeric(A, Fs, Fn)
    scale = abs(A).sum(axis=0).max()
    assert round(abs(R @ P @ A @ Q - (Lm @ Um + Fm)).sum(axis=0).max() / scale, 7) == 0          # the reference's criterion
    if case == 0:
        assert len(r) - 1 >= 6 and Fm.nnz > 0
    rng = np.random.default_rng(case)
    B = rng.standard_normal((n, 2))
    for trans, M in (("N", A), ("T", A.T)):
        X = np.asfortranarray(B.copy())
        klu.solve(A, Fs, Fn, X, trans=trans)
        assert np.abs(M @ X - B).max() <= 1e-9 * max(1.0, np.abs(X).max()) * scale
    batch = 37
    vals = A.data[None, :] * (1 + 1e-3 * rng.uniform(-1, 1, size=(batch, A.nnz)))
    assert not klu.refactor_batch(Fn, vals).any()
    Bb = rng.standard_normal((batch, 1, n))
    Xb = Bb.copy()
    klu.solve_batch(Fn, Xb)
    for b in (0, 18, 36):
        Ab = sp.csc_matrix((vals[b], A.indices, A.indptr), shape=(n, n))
        xref = spla.splu(Ab.tocsc()).solve(Bb[b, 0])
        assert np.linalg.norm(Xb[b, 0] - xref) <= 1e-9 * np.linalg.norm(xref)


def test_solve_ldB_buffer_ends_with_the_last_column(klu):
    """klu.c:593-690: B needs offsetB + (nrhs-1)*ldB + n entries, not nrhs*ldB: nothing past the last column's n entries
    may be read or written (regression: the host copies used nrhs*ldB doubles)"""
    A = load_matrix("bp_800")
    n = A.shape[0]
    Fs = klu.symbolic(A); Fn = klu.numeric(A, Fs)
    rng = np.random.default_rng(5)
    ld, off = n + 7, 3
    buf = rng.standard_normal(off + ld + n)            # exactly the documented minimum for nrhs = 2
    keep = buf.copy()
    klu.solve(A, Fs, Fn, buf, nrhs=2, ldB=ld, offsetB=off)
    for j in range(2):
        b = keep[off + j * ld: off + j * ld + n]
        x = buf[off + j * ld: off + j * ld + n]
        assert np.abs(A @ x - b).max() <= 1e-9 * max(1.0, np.abs(x).max()) * abs(A).sum(axis=0).max()
    mask = np.ones(buf.size, bool)
    for j in range(2):
        mask[off + j * ld: off + j * ld + n] = False
    assert np.array_equal(buf[mask], keep[mask])


@pytest.mark.parametrize("name", ["ACTIVSg2000", "bp_800", "bcsstk13", "arrow"])
@pytest.mark.parametrize("trans", ["N", "T"])
def test_one_matrix_solve_kernel_matches_level_kernel(klu, name, trans, monkeypatch):
    """klu.solve on ONE matrix runs the sequential operation tape in one CTA per right-hand side (k_klu_solve_one); with
    B200S_KLU_SOLVE_ONE=0 the same call goes through the batched level kernel (k_klu_solve_lvl).  Same factor, two summation
    orders: the solutions agree to rounding and with SuperLU; 'arrow' has a column longer than one chunk of the tape."""
    if name == "arrow":
        n = 2600
        rng = np.random.default_rng(3)
        A = sp.lil_matrix((n, n)); A.setdiag(rng.uniform(2, 3, n))
        A[n - 1, :] = rng.uniform(-1e-2, 1e-2, n); A[:, n - 1] = rng.uniform(-1e-2, 1e-2, (n, 1)); A[n - 1, n - 1] = 4.0
        A = A.tocsc(); A.sort_indices()
    else:
        A = load_matrix(name)
    n = A.shape[0]
    B = np.random.default_rng(1).standard_normal((n, 5))
    sols = []
    for flag in ("1", "0"):
        monkeypatch.setenv("B200S_KLU_SOLVE_ONE", flag)        # read when the numeric object is created
        Fs = klu.symbolic(A); Fn = klu.numeric(A, Fs)
        X = np.asfortranarray(B.copy())
        klu.solve(A, Fs, Fn, X, trans=trans)
        sols.append(X)
    M = (A if trans == "N" else A.T).tocsc()
    Xref = spla.splu(M).solve(B)
    for X in sols:
        assert np.linalg.norm(X - Xref) / np.linalg.norm(Xref) < 1e-10
    assert np.linalg.norm(sols[0] - sols[1]) / np.linalg.norm(sols[1]) < 1e-12
    assert not np.array_equal(sols[0], sols[1]) or name == "bcsstk13"      # two kernels really ran (different rounding)
