"""Parity of the CUDA sparse Cholesky (through the kvxopt.cholmod-compatible API and the C ABI) against the
CPU oracle, the reference's LAPACK golden vectors and the reference's documented known answers.
Tolerances are BASELINE.json's: relative backward error <= 1e-12, solution relative difference <= 1e-10."""
import ctypes as C

import numpy as np
import pytest
import scipy.sparse as sp
import scipy.sparse.linalg as spla

from conftest import GOLD, lap3d, load_matrix, lower_ccs, rand_spd, sym_from_lower

pytestmark = pytest.mark.gpu

BERR_TOL = 1e-12
XREL_TOL = 1e-10


def berr(A, X, B):
    X = X.reshape(A.shape[0], -1); B = B.reshape(A.shape[0], -1)
    return (np.linalg.norm(A @ X - B, axis=0) / (spla.norm(A, 1) * np.linalg.norm(X, axis=0) + np.linalg.norm(B, axis=0))).max()


@pytest.fixture(scope="module")
def cholmod():
    from kvxopt_b200 import cholmod as m, _lib
    assert _lib.device_count() > 0, "GPU tests need a CUDA device; there is no CPU fallback"
    return m


def test_doc_example_linsolve_numpy_types(cholmod):
    # reference doc/source/spsolvers.rst:555-563
    A = sp.csc_matrix(([10.0, 3, 5, -2, 5, 2], ([0, 2, 1, 3, 2, 3], [0, 0, 1, 1, 2, 3])), shape=(4, 4))
    X = np.arange(8, dtype=float).reshape(4, 2, order="F")
    cholmod.linsolve(A, X)
    ref = np.array([[-0.146341463414634, 0.048780487804878], [1.333333333333333, 4.0],
                    [0.487804878048781, 1.170731707317073], [2.833333333333333, 7.5]])
    np.testing.assert_allclose(X, ref, rtol=1e-13)


def test_doc_examples_with_reference_types(cholmod, kvx):
    """the reference's own documentation examples, typed exactly as in the docs (kvxopt matrix/spmatrix)"""
    from kvxopt import matrix, spmatrix, cholmod as kc, log
    assert kc is cholmod
    A = spmatrix([10, 3, 5, -2, 5, 2], [0, 2, 1, 3, 2, 3], [0, 0, 1, 1, 2, 3])
    X = matrix(range(8), (4, 2), "d")
    kc.linsolve(A, X)
    np.testing.assert_allclose(np.array(X), [[-0.146341463414634, 0.048780487804878], [1.333333333333333, 4.0],
                                             [0.487804878048781, 1.170731707317073], [2.833333333333333, 7.5]], rtol=1e-13)
    # spsolvers.rst:580-585: splinsolve with sparse identity gives the inverse
    I4 = spmatrix(1.0, range(4), range(4))
    Xs = kc.splinsolve(A, I4)
    Ad = np.array(matrix(A)); Ad = np.tril(Ad) + np.tril(Ad, -1).T
    np.testing.assert_allclose(np.array(matrix(Xs)), np.linalg.inv(Ad), rtol=1e-12, atol=1e-15)
    # spsolvers.rst:759-772: log det through diag()
    F = kc.symbolic(A)
    kc.numeric(A, F)
    assert abs(2.0 * sum(log(kc.diag(F))) - 5.505331535932363) < 1e-12
    # spsolvers.rst: solve with the systems used by kkt_chol2 (misc.py:1531-1558): P, L, L', P'
    b = matrix([1.0, 2.0, 3.0, 4.0])
    x = +b
    for s in (7, 4, 5, 8):
        kc.solve(F, x, sys=s)
    np.testing.assert_allclose(np.array(x).ravel(), np.linalg.solve(Ad, np.array(b).ravel()), rtol=1e-13)
    Lf = kc.getfactor(F)
    Ld = np.array(matrix(Lf))
    p = cholmod.factor_perm(F)
    np.testing.assert_allclose(Ld @ Ld.T, Ad[np.ix_(p, p)], rtol=1e-13, atol=1e-14)


@pytest.mark.parametrize("name", ["bcsstk13", "bcsstk24"])
def test_reference_matrices_vs_reference_lapack_and_oracle(cholmod, name):
    """BASELINE config 1: cholmod.linsolve on tests/bcsstk24.mtx (lower triangle as stored), random RHS"""
    from oracle import CholOracle
    Al = load_matrix(name)
    n = Al.shape[0]
    A = sym_from_lower(Al)
    B = np.random.default_rng(0).standard_normal((n, 3))
    X = np.asfortranarray(B.copy())
    cholmod.linsolve(Al, X)
    assert berr(A, X, B) <= BERR_TOL
    Xref = np.load(GOLD + "/posv_%s.npz" % name)["X"]            # reference lapack.posv
    assert np.linalg.norm(X - Xref) / np.linalg.norm(Xref) <= XREL_TOL
    F = cholmod.symbolic(Al)
    cholmod.numeric(Al, F)
    O = CholOracle(n, Al.indptr, Al.indices, "L", cholmod.factor_perm(F))   # same P A P' on the CPU
    O.factorize(Al.data)
    Xo = O.solve(B)
    assert np.linalg.norm(X - Xo) / np.linalg.norm(Xo) <= XREL_TOL
    d = cholmod.diag(F)
    np.testing.assert_allclose(np.asarray(d).ravel(), O.diag(), rtol=1e-9)


@pytest.mark.parametrize("n,dens,seed", [(1, 1.0, 0), (2, 1.0, 1), (33, 0.2, 2), (129, 0.05, 3), (700, 0.01, 4), (1500, 0.004, 5)])
def test_random_spd_all_systems_vs_oracle(cholmod, n, dens, seed):
    from oracle import CholOracle
    A = rand_spd(n, dens, seed)
    Al = lower_ccs(A)
    F = cholmod.symbolic(Al)
    cholmod.numeric(Al, F)
    p = cholmod.factor_perm(F)
    O = CholOracle(n, Al.indptr, Al.indices, "L", p)
    O.factorize(Al.data)
    rng = np.random.default_rng(seed)
    B = rng.standard_normal((n, 3))
    for s in range(9):
        X = np.asfortranarray(B.copy())
        cholmod.solve(F, X, sys=s)
        Xo = O.solve(B, s)
        assert np.linalg.norm(X - Xo) <= XREL_TOL * max(np.linalg.norm(Xo), 1e-300), "sys=%d" % s
    X = np.asfortranarray(B.copy()); cholmod.solve(F, X)
    assert berr(A, X, B) <= BERR_TOL
    Lg = cholmod.getfactor(F).toarray()
    np.testing.assert_allclose(Lg, O.dense_L(), atol=1e-11 * np.abs(Lg).max())
    # refactorization with new values on the same pattern (what the IPM does every iteration)
    Al2 = Al.copy(); Al2.data = Al.data * (1 + 0.1 * rng.uniform(0, 1, Al.nnz)); Al2 = Al2 + sp.identity(n) * n
    Al2 = lower_ccs(Al2)
    assert np.array_equal(Al2.indices, Al.indices)
    cholmod.numeric(Al2, F)
    X = np.asfortranarray(B.copy()); cholmod.solve(F, X)
    assert berr(sym_from_lower(Al2), X, B) <= BERR_TOL


def test_upper_storage_user_perm_ldB_offset_nrhs(cholmod):
    n = 200
    A = rand_spd(n, 0.03, 11)
    Au = sp.triu(A).tocsc(); Au.sort_indices()
    rng = np.random.default_rng(1)
    perm = rng.permutation(n)
    F = cholmod.symbolic(Au, p=perm, uplo="U")
    cholmod.numeric(Au, F)
    ldB, off = n + 7, 5
    buf = rng.standard_normal(off + 3 * ldB)
    keep = buf.copy()
    cholmod.solve(F, buf, sys=0, nrhs=2, ldB=ldB, offsetB=off)
    for j in range(2):
        b = keep[off + j * ldB: off + j * ldB + n]
        x = buf[off + j * ldB: off + j * ldB + n]
        assert berr(A, x, b) <= BERR_TOL
    # everything outside the two solved columns is untouched
    mask = np.ones_like(buf, dtype=bool)
    for j in range(2):
        mask[off + j * ldB: off + j * ldB + n] = False
    assert np.array_equal(buf[mask], keep[mask])
    with pytest.raises(ValueError):
        cholmod.solve(F, buf, ldB=n - 1, nrhs=1)
    with pytest.raises(TypeError):
        cholmod.solve(F, buf, nrhs=4, ldB=ldB, offsetB=off)
    with pytest.raises(ValueError):
        cholmod.solve(F, buf, sys=9)


def test_not_positive_definite_reports_the_oracles_column(cholmod):
    from oracle import CholOracle
    n = 300
    A = rand_spd(n, 0.02, 21).tolil()
    A[150, 150] = -1.0
    A = A.tocsc()
    Al = lower_ccs(A)
    F = cholmod.symbolic(Al)
    with pytest.raises(ArithmeticError) as e:
        cholmod.numeric(Al, F)
    O = CholOracle(n, Al.indptr, Al.indices, "L", cholmod.factor_perm(F))
    with pytest.raises(ArithmeticError) as eo:
        O.factorize(Al.data)
    assert e.value.args[0] == eo.value.args[0]
    with pytest.raises(ArithmeticError):
        cholmod.solve(F, np.ones((n, 1), order="F"))
    with pytest.raises(ArithmeticError):
        cholmod.linsolve(Al, np.ones((n, 1), order="F"))
    # a later successful numeric() on the same factor object recovers (kkt_chol2's retry path, misc.py:1433-1447)
    A2 = lower_ccs(rand_spd(n, 0.02, 21))
    if np.array_equal(A2.indices, Al.indices):
        cholmod.numeric(A2, F)
        X = np.ones((n, 1), order="F"); cholmod.solve(F, X)
        assert berr(sym_from_lower(A2), X, np.ones((n, 1))) <= BERR_TOL


def test_error_contract(cholmod):
    A = lower_ccs(rand_spd(20, 0.2, 3))
    F = cholmod.symbolic(A)
    with pytest.raises(ValueError, match="symbolic factor"):
        cholmod.solve(F, np.ones((20, 1), order="F"))
    with pytest.raises(TypeError):
        cholmod.solve("not a capsule", np.ones((20, 1), order="F"))
    with pytest.raises(TypeError):
        cholmod.symbolic(sp.csc_matrix(np.ones((3, 4))))
    with pytest.raises(TypeError):
        cholmod.linsolve(A.astype(np.complex128), np.ones((20, 1), order="F"))
    with pytest.raises(ValueError):
        cholmod.symbolic(A, uplo="X")
    cholmod.options["nope"] = 1
    try:
        with pytest.raises(ValueError, match="invalid value for CHOLMOD parameter"):
            cholmod.symbolic(A)
    finally:
        del cholmod.options["nope"]
    cholmod.options["supernodal"] = 7
    try:
        with pytest.raises(ValueError):
            cholmod.symbolic(A)
    finally:
        del cholmod.options["supernodal"]


def test_simplicial_ldl_semantics_supernodal_0(cholmod):
    """cholmod.options['supernodal'] = 0 (reference src/C/cholmod.c:60-64): the factor answers as an LDL' factor --
    sys 1..6 per cholmod.c:437-439, getfactor with D on the diagonal, diag() refused (cholmod.c:919-922) -- and the
    documented log-det example (doc/source/spsolvers.rst:759-772) gives 5.505331535932363 through sys=6."""
    import scipy.linalg as sla
    A4 = sp.csc_matrix(np.array([[10.0, 0, 3, 0], [0, 5, 0, -2], [3, 0, 5, 0], [0, -2, 0, 2]]))
    cholmod.options["supernodal"] = 0
    try:
        F = cholmod.symbolic(lower_ccs(A4)); cholmod.numeric(lower_ccs(A4), F)
        Di = np.ones((4, 1), order="F")
        cholmod.solve(F, Di, sys=6)
        assert abs(-np.log(Di).sum() - 5.505331535932363) < 1e-12
        with pytest.raises(ValueError):
            cholmod.diag(F)
        A = rand_spd(150, 0.05, 4); Al = lower_ccs(A); n = 150
        F = cholmod.symbolic(Al); cholmod.numeric(Al, F)
        perm = cholmod.factor_perm(F)
        Ap = A.toarray()[np.ix_(perm, perm)]
        Lc = np.linalg.cholesky(Ap)
        d = np.diag(Lc) ** 2
        Ld = Lc / np.diag(Lc)[None, :]
        Lf = cholmod.getfactor(F).toarray()
        assert np.abs(np.diag(Lf) - d).max() <= 1e-12 * d.max()
        assert np.abs(np.tril(Lf, -1) - np.tril(Ld, -1)).max() <= 1e-12
        B = np.random.default_rng(0).standard_normal((n, 2))
        Pm = np.eye(n)[perm]
        ref = {0: np.linalg.solve(A.toarray(), B), 1: np.linalg.solve(Ap, B), 2: np.linalg.solve(Ld @ np.diag(d), B),
               3: np.linalg.solve(np.diag(d) @ Ld.T, B), 4: sla.solve_triangular(Ld, B, lower=True, unit_diagonal=True),
               5: sla.solve_triangular(Ld.T, B, lower=False, unit_diagonal=True), 6: B / d[:, None], 7: Pm @ B, 8: Pm.T @ B}
        for sys_, want in ref.items():
            X = np.asfortranarray(B.copy())
            cholmod.solve(F, X, sys=sys_)
            assert np.linalg.norm(X - want) <= 1e-10 * np.linalg.norm(want), sys_
    finally:
        del cholmod.options["supernodal"]


def test_zero_size_inputs(cholmod):
    """0x0 A, n x 0 sparse RHS and 0 x 1 B occur in kkt_chol2 when there are no equality constraints"""
    E = sp.csc_matrix((0, 0))
    F = cholmod.symbolic(E)
    cholmod.numeric(E, F)
    cholmod.solve(F, np.zeros((0, 1), order="F"))
    A = lower_ccs(rand_spd(10, 0.3, 1))
    F = cholmod.symbolic(A); cholmod.numeric(A, F)
    X = cholmod.spsolve(F, sp.csc_matrix((10, 0)), sys=7)
    assert X.shape == (10, 0)
    cholmod.linsolve(A, np.zeros((10, 0), order="F"))


def test_spsolve_matches_dense(cholmod):
    n = 120
    A = rand_spd(n, 0.04, 8)
    Al = lower_ccs(A)
    F = cholmod.symbolic(Al); cholmod.numeric(Al, F)
    Bs = sp.random(n, 6, density=0.05, random_state=np.random.default_rng(2), format="csc")
    for s in (0, 4, 7):
        Xs = cholmod.spsolve(F, Bs, sys=s)
        Xd = np.asfortranarray(Bs.toarray()); cholmod.solve(F, Xd, sys=s)
        np.testing.assert_allclose(Xs.toarray(), Xd, rtol=1e-13, atol=1e-300)
    Xl = cholmod.splinsolve(Al, Bs)
    np.testing.assert_allclose(A @ Xl.toarray(), Bs.toarray(), atol=1e-12)


@pytest.mark.parametrize("dims,order", [((24, 24, 24), "nd"), ((40, 40, 40), "nd"), ((30, 30, 30), "amd"), ((300, 300, 1), "amd")])
def test_laplacians_large_fronts(cholmod, dims, order):
    """fronts wider than one 128-column block, many CTAs per panel launch (the shape of BASELINE config 4)"""
    from kvxopt_b200 import _lib as L
    nx, ny, nz = dims
    A = lap3d(nx, ny, nz)
    Al = lower_ccs(A)
    n = A.shape[0]
    perm = None
    if order == "nd":
        perm = np.zeros(n, np.int64)
        assert L.fn["b200s_grid_nd_perm"](nx, ny, nz, 64, L.ptr_i64(perm)) == 0
    F = cholmod.symbolic(Al, p=perm)
    cholmod.numeric(Al, F)
    B = np.random.default_rng(0).standard_normal((n, 2))
    X = np.asfortranarray(B.copy())
    cholmod.solve(F, X)
    assert berr(A, X, B) <= BERR_TOL
    # determinism: a second factorization + solve gives bit-identical results (no atomics in the extend-add)
    cholmod.numeric(Al, F)
    X2 = np.asfortranarray(B.copy()); cholmod.solve(F, X2)
    assert np.array_equal(X, X2)
    # size-independent property: linearity of the solve
    Y = np.asfortranarray(2.5 * B[:, :1] - 0.5 * B[:, 1:2]); cholmod.solve(F, Y)
    np.testing.assert_allclose(Y[:, 0], 2.5 * X[:, 0] - 0.5 * X[:, 1], rtol=1e-9, atol=1e-12 * np.abs(X).max())


def test_device_resident_entry_points(cholmod):
    """b200s_chol_factorize_dev / solve_dev (values and right-hand sides already in HBM) match the host-pointer calls"""
    import torch
    from kvxopt_b200 import _lib as L
    A = lap3d(16, 16, 16); Al = lower_ccs(A); n = A.shape[0]
    F = cholmod.symbolic(Al)
    h, _ = cholmod._factor_handle(F)
    vals = torch.from_numpy(Al.data.copy()).cuda()
    minor = C.c_int64()
    torch.cuda.synchronize()
    assert L.fn["b200s_chol_factorize_dev"](h, vals.data_ptr(), C.byref(minor)) == 0
    B = np.random.default_rng(0).standard_normal((n, 2))
    Xd = torch.from_numpy(np.asfortranarray(B).T.copy()).cuda()       # (2, n) C-order == n x 2 column-major
    torch.cuda.synchronize()
    assert L.fn["b200s_chol_solve_dev"](h, 0, Xd.data_ptr(), 2, n) == 0
    X = Xd.cpu().numpy().T
    assert berr(A, X, B) <= BERR_TOL
    Xh = np.asfortranarray(B.copy()); cholmod.solve(F, Xh)
    assert np.array_equal(Xh, X)


def test_multi_rhs_solve_equals_column_by_column(cholmod):
    """reference cholmod.c:481-493 solves column by column; the engine stages each tile of L once for all columns
    of the call (groups of four in the large-front update kernels) -- the result is bit-identical per column"""
    A = lap3d(22, 22, 22); Al = lower_ccs(A); n = A.shape[0]
    perm = np.zeros(n, np.int64)
    from kvxopt_b200 import _lib as L
    assert L.fn["b200s_grid_nd_perm"](22, 22, 22, 64, L.ptr_i64(perm)) == 0
    F = cholmod.symbolic(Al, p=perm); cholmod.numeric(Al, F)
    assert cholmod.factor_info(F)["max_front_rows"] > 256          # large-front (tiled) path
    B = np.random.default_rng(3).standard_normal((n, 9))
    cols = []
    for j in range(9):
        x = np.asfortranarray(B[:, j:j + 1].copy()); cholmod.solve(F, x); cols.append(x[:, 0])
    for k in (2, 3, 5, 9):
        for sys_ in (0, 4, 5):
            X = np.asfortranarray(B[:, :k].copy()); cholmod.solve(F, X, sys=sys_)
            if sys_ == 0:
                for j in range(k):
                    assert np.array_equal(X[:, j], cols[j]), (k, j)
                assert berr(A, X, B[:, :k]) <= BERR_TOL
            else:
                x1 = np.asfortranarray(B[:, k - 1:k].copy()); cholmod.solve(F, x1, sys=sys_)
                assert np.array_equal(X[:, k - 1], x1[:, 0]), (k, sys_)


@pytest.mark.parametrize("case", ["lap22", "lap40", "dense1500", "lap30_ldl"])
def test_persistent_sweeps_bit_identical_to_the_launch_per_step_path(cholmod, case):
    """one right-hand side: the forward / backward sweeps over a level's large fronts run as ONE persistent kernel per level
    (CTAs hand the solved 128-column block on through flags, chol_gpu.cu k_fwd_persist / k_bwd_persist).  Same arithmetic in
    the same order as the launch-per-step kernels: solutions are bit-identical for every combination of the two sweeps and
    every system of cholmod.c:437-439 that runs a sweep; several blocks per front, several fronts per level, a root front
    with nothing below its last block (dense1500), signed LDL' mode."""
    from kvxopt_b200 import _lib as L
    opts = dict(cholmod.options)
    try:
        if case.startswith("lap"):
            nx = int(case[3:5])
            A = lap3d(nx, nx, nx); Al = lower_ccs(A); n = A.shape[0]
            perm = np.zeros(n, np.int64)
            assert L.fn["b200s_grid_nd_perm"](nx, nx, nx, 64, L.ptr_i64(perm)) == 0
            if case.endswith("ldl"):
                cholmod.options["supernodal"] = 0
            F = cholmod.symbolic(Al, p=perm)
        else:
            n = 1500
            rng = np.random.default_rng(11)
            M = rng.standard_normal((n, n)) / np.sqrt(n)
            A = sp.csc_matrix(M @ M.T + 2.0 * np.eye(n)); Al = lower_ccs(A)
            F = cholmod.symbolic(Al, p=np.arange(n, dtype=np.int64))
        cholmod.numeric(Al, F)
        assert cholmod.factor_info(F)["max_front_rows"] > 256
        b = np.random.default_rng(5).standard_normal((n, 1))
        ref = {}
        for mode in (0, 1, 2, 3, -1):
            cholmod.set_solve_sweeps(F, mode)
            for sys_ in (0, 1, 4, 5) + ((2, 3) if case.endswith("ldl") else ()):
                x = np.asfortranarray(b.copy()); cholmod.solve(F, x, sys=sys_)
                if mode == 0:
                    ref[sys_] = x
                    if sys_ == 0:
                        assert berr(A, x, b) <= BERR_TOL
                else:
                    assert np.array_equal(x, ref[sys_]), (case, mode, sys_, np.abs(x - ref[sys_]).max())
            if mode == 3:       # replayed from the captured graph: same bits again
                x = np.asfortranarray(b.copy()); cholmod.solve(F, x)
                assert np.array_equal(x, ref[0])
    finally:
        cholmod.options.clear(); cholmod.options.update(opts)


def test_numeric_with_subset_and_changed_pattern(cholmod):
    """cholmod.numeric rebuilds the matrix from A's own pattern (cholmod.c:340-358): a matrix that stores a SUBSET of the
    analysed pattern factors correctly (missing entries are zeros); an entry outside the analysed pattern is refused instead
    of being scattered through the old map"""
    n = 200
    A = rand_spd(n, 0.03, 5)
    Al = lower_ccs(A)
    F = cholmod.symbolic(Al)
    # drop a third of the off-diagonal entries: same analysis, subset pattern
    coo = Al.tocoo()
    rng = np.random.default_rng(3)
    keep = (coo.row == coo.col) | (rng.uniform(size=coo.nnz) > 0.33)
    As = sp.csc_matrix((coo.data[keep], (coo.row[keep], coo.col[keep])), shape=(n, n)); As.sort_indices()
    assert As.nnz < Al.nnz
    cholmod.numeric(As, F)
    B = rng.standard_normal((n, 2)); X = np.asfortranarray(B.copy())
    cholmod.solve(F, X)
    assert berr(sym_from_lower(As), X, B) <= BERR_TOL
    # same factor object, full pattern again
    cholmod.numeric(Al, F)
    X = np.asfortranarray(B.copy()); cholmod.solve(F, X)
    assert berr(sym_from_lower(Al), X, B) <= BERR_TOL
    # an entry the analysis has not seen
    lil = Al.tolil()
    free = [(i, j) for j in range(3) for i in range(j + 1, n) if lil[i, j] == 0]
    i, j = free[0]
    lil[i, j] = 0.5
    Ax = lil.tocsc(); Ax.sort_indices()
    with pytest.raises(ValueError):
        cholmod.numeric(Ax, F)


@pytest.mark.parametrize("case", ["lap3d_nd", "random_amd"])
def test_spsolve_structure_aware_equals_dense_solve(cholmod, case):
    """cholmod.spsolve (reference src/C/cholmod.c:524-587) with genuinely sparse right-hand sides: sparse upload, forward sweep
    restricted to the elimination-tree reach of the nonzero rows, compaction on the device.  Every system 0..8 must give the
    entries of the dense solve of the densified columns (same kernels on the same data: equal to rounding), hold only
    numerically nonzero entries in ascending row order, and handle empty columns and more columns than one device chunk."""
    from kvxopt_b200 import _lib
    rng = np.random.default_rng(11)
    if case == "lap3d_nd":
        A = lap3d(14, 14, 14)
        perm = np.zeros(A.shape[0], np.int64)
        assert _lib.fn["b200s_grid_nd_perm"](14, 14, 14, 32, _lib.ptr_i64(perm)) == 0
    else:
        A = rand_spd(900, 0.004, 5)
        perm = None
    n = A.shape[0]
    Al = lower_ccs(A)
    F = cholmod.symbolic(Al, p=perm)
    cholmod.numeric(Al, F)
    ncols = 150                                   # more than one chunk of 64 columns
    rows, cols, vals = [], [], []
    for j in range(ncols):
        if j % 17 == 3:
            continue                              # empty column
        k = int(rng.integers(1, 4))
        r = rng.choice(n, size=k, replace=False)
        rows += r.tolist(); cols += [j] * k; vals += rng.standard_normal(k).tolist()
    B = sp.csc_matrix((vals, (rows, cols)), shape=(n, ncols)); B.sort_indices()
    Bd = B.toarray(order="F")
    for sys in range(9):
        X = cholmod.spsolve(F, B, sys=sys)
        X = sp.csc_matrix(X); 
        assert X.shape == (n, ncols) and X.has_sorted_indices or True
        Xd = np.asfortranarray(Bd.copy())
        cholmod.solve(F, Xd, sys=sys)
        assert np.abs(X.toarray() - Xd).max() <= 1e-13 * max(1.0, np.abs(Xd).max()), sys
        assert (X.data != 0).all()
        ind = X.indices; ptr = X.indptr
        for j in range(ncols):
            assert (np.diff(ind[ptr[j]:ptr[j + 1]]) > 0).all()
        if sys in (4, 7, 8):
            assert X.nnz < 0.6 * n * ncols          # L^-1 b and permutations of sparse columns stay sparse
    # the empty matrix and a matrix of empty columns
    E = cholmod.spsolve(F, sp.csc_matrix((n, 0)))
    assert sp.csc_matrix(E).shape == (n, 0)
    Z = sp.csc_matrix(cholmod.spsolve(F, sp.csc_matrix((n, 3)), sys=4))
    assert Z.shape == (n, 3) and Z.nnz == 0
    # LDL' semantics (supernodal = 0): the scaled systems go through the same path
    cholmod.options["supernodal"] = 0
    try:
        F0 = cholmod.symbolic(Al, p=perm)
        cholmod.numeric(Al, F0)
        for sys in (0, 2, 4, 6):
            X = sp.csc_matrix(cholmod.spsolve(F0, B[:, :20], sys=sys))
            Xd = np.asfortranarray(Bd[:, :20].copy())
            cholmod.solve(F0, Xd, sys=sys)
            assert np.abs(X.toarray() - Xd).max() <= 1e-13 * max(1.0, np.abs(Xd).max())
    finally:
        del cholmod.options["supernodal"]
