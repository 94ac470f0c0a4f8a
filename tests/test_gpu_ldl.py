"""Sparse LDL' without pivoting for symmetric indefinite (quasi-definite) matrices -- what
cholmod.options['supernodal'] = 0 selects in the reference (src/C/cholmod.c:60-64, sys 1..6 at :437-439) -- and the
sparse counterpart of the 'ldl' KKT solver (src/python/misc.py:1055-1130).  Parity against oracle/ldl_oracle.py with
the same permutation, numpy, and the reference's own sytrf/sytrs answers (tests/golden/kkt_ldl_ref.npz).
Tolerances are BASELINE.json's: relative backward error <= 1e-12, solution relative difference <= 1e-10."""
import os
import sys

import numpy as np
import pytest
import scipy.sparse as sp
import scipy.sparse.linalg as spla

from conftest import GOLD, lap3d, lower_ccs, rand_spd

pytestmark = pytest.mark.gpu

BERR_TOL = 1e-12
XREL_TOL = 1e-10


@pytest.fixture(scope="module")
def cholmod():
    from kvxopt_b200 import cholmod as m, _lib
    assert _lib.device_count() > 0, "GPU tests need a CUDA device; there is no CPU fallback"
    m.options["supernodal"] = 0
    yield m
    del m.options["supernodal"]


def quasi_definite(E, F, dens, seed):
    """[[E, B'], [B, -F]] with E, F positive definite"""
    rng = np.random.default_rng(seed)
    B = sp.random(F.shape[0], E.shape[0], density=dens, random_state=rng, format="csc")
    return sp.bmat([[E, B.T], [B, -F]]).tocsc()


def berr(A, X, B):
    X = X.reshape(A.shape[0], -1); B = B.reshape(A.shape[0], -1)
    return (np.linalg.norm(A @ X - B, axis=0) / (spla.norm(A, 1) * np.linalg.norm(X, axis=0) + np.linalg.norm(B, axis=0))).max()


CASES = {
    # small fronts only (one CTA per front in shared memory)
    "small": lambda: quasi_definite(rand_spd(120, 0.05, 1), rand_spd(80, 0.05, 2), 0.03, 3),
    # nested-dissection-like fill: fronts above 128 rows => panel kernel + DMMA updates, signs mixed inside blocks
    "grid": lambda: quasi_definite((lap3d(9, 9, 9) + sp.identity(729)).tocsc(), (lap3d(8, 8, 4) + sp.identity(256)).tocsc(), 0.004, 4),
    # one dense front of 700 columns: every block-column step, the near (K = 128) and far (K = 512) updates
    "dense": lambda: quasi_definite(rand_spd(420, 0.6, 5), rand_spd(280, 0.6, 6), 0.5, 7),
}


@pytest.mark.parametrize("name", list(CASES))
def test_ldl_factor_and_systems_vs_oracle(cholmod, name):
    from oracle import ldl_oracle
    K = CASES[name]()
    n = K.shape[0]
    Kl = lower_ccs(K)
    F = cholmod.symbolic(Kl); cholmod.numeric(Kl, F)
    perm = cholmod.factor_perm(F)
    Kd = K.toarray()
    Lo, do, minor = ldl_oracle.ldl_nopivot(Kd[np.ix_(perm, perm)])
    assert minor == n
    nneg = int((np.linalg.eigvalsh(Kd) < 0).sum())
    assert int((do < 0).sum()) == nneg                     # Sylvester: the oracle's inertia is the matrix's
    Lf = cholmod.getfactor(F).toarray()                    # unit lower triangle with D on the diagonal
    d = np.diag(Lf).copy()
    assert int((d < 0).sum()) == nneg and nneg > 0
    Lg = np.tril(Lf, -1) + np.eye(n)
    scale = np.abs(do).max()
    assert np.abs(d - do).max() <= 1e-10 * scale
    assert np.abs(Lg - Lo).max() <= 1e-10 * max(1.0, np.abs(Lo).max())
    R = Lg @ np.diag(d) @ Lg.T - Kd[np.ix_(perm, perm)]
    assert np.abs(R).max() <= 1e-12 * np.abs(Kd).max() * max(1.0, np.abs(Lo).max() ** 2)
    B = np.random.default_rng(0).standard_normal((n, 3))
    for sys_ in range(9):
        X = np.asfortranarray(B.copy())
        cholmod.solve(F, X, sys=sys_)
        want = ldl_oracle.solve_sys(Lo, do, perm, B, sys_)
        assert np.linalg.norm(X - want) <= XREL_TOL * np.linalg.norm(want), (name, sys_)
    X = np.asfortranarray(B.copy())
    cholmod.linsolve(Kl, X)
    assert berr(K, X, B) <= BERR_TOL
    assert np.linalg.norm(X - np.linalg.solve(Kd, B)) <= XREL_TOL * np.linalg.norm(X)
    with pytest.raises(ValueError):
        cholmod.diag(F)                                    # cholmod.c:919-922
    # refactorization with new values on the same symbolic object, bit-identical repeat
    cholmod.numeric(Kl, F)
    X2 = np.asfortranarray(B.copy()); cholmod.solve(F, X2)
    X3 = np.asfortranarray(B.copy()); cholmod.solve(F, X3)
    assert np.array_equal(X2, X3)


def test_ldl_zero_pivot_raises(cholmod):
    """a zero pivot is the one failure of LDL' without pivoting: ArithmeticError(k) (cholmod.c:376-380)"""
    # second pivot 1 - 2 * 2 / 4 = 0 exactly, also in the engine's signed square-root form (sqrt(4) and 1/2 are exact)
    K = sp.csc_matrix(np.array([[4.0, 2, 0], [2, 1.0, 1], [0, 1, 3.0]]))
    with pytest.raises(ArithmeticError) as e:
        F = cholmod.symbolic(lower_ccs(K), p=np.arange(3)); cholmod.numeric(lower_ccs(K), F)
    assert e.value.args[0] == 1


def test_ldl_sparse_rhs_and_ownership_refused(cholmod):
    """spsolve / splinsolve go through the same signed sweeps; front ownership (multi-GPU building block) is refused in
    LDL' mode instead of silently dropping the pivot signs of remote fronts"""
    import ctypes as C
    from kvxopt_b200 import _lib as L
    K = CASES["small"]()
    n = K.shape[0]
    Kl = lower_ccs(K)
    Bs = sp.random(n, 4, density=0.1, random_state=np.random.default_rng(2), format="csc")
    X = cholmod.splinsolve(Kl, Bs)
    want = np.linalg.solve(K.toarray(), Bs.toarray())
    assert np.linalg.norm(X.toarray() - want) <= XREL_TOL * np.linalg.norm(want)
    F = cholmod.symbolic(Kl)
    h, _ = cholmod._factor_handle(F)
    owned = (C.c_ubyte * 100000)(*([1] * 100000))
    assert L.fn["b200s_chol_set_owned"](h, C.cast(owned, C.c_char_p)) == L.INVALID


def test_ldl_large_quasi_definite_properties(cholmod):
    """beyond the dense oracle's reach: 30^3 Laplacian block + constraints (n = 31 000), checked by backward error,
    linearity and inertia (number of negative pivots = order of the negative block)"""
    E = (lap3d(30, 30, 30) + 0.5 * sp.identity(27000)).tocsc()
    Fm = (lap3d(20, 20, 10) + sp.identity(4000)).tocsc()
    K = quasi_definite(E, Fm, 2e-4, 9)
    Kl = lower_ccs(K)
    n = K.shape[0]
    F = cholmod.symbolic(Kl); cholmod.numeric(Kl, F)
    B = np.random.default_rng(1).standard_normal((n, 2))
    X = np.asfortranarray(B.copy()); cholmod.solve(F, X)
    assert berr(K, X, B) <= BERR_TOL
    Y = np.asfortranarray((2.0 * B[:, :1] - 3.0 * B[:, 1:2]).copy()); cholmod.solve(F, Y)
    assert np.linalg.norm(Y - (2.0 * X[:, :1] - 3.0 * X[:, 1:2])) <= 1e-10 * np.linalg.norm(Y)
    D = np.asfortranarray(np.ones((n, 1))); cholmod.solve(F, D, sys=6)
    assert int((D < 0).sum()) == 4000


def _kkt_case(prefix):
    z = np.load(os.path.join(GOLD, "kkt_ldl_ref.npz"))
    g = {k[len(prefix):]: z[k] for k in z.files if k.startswith(prefix)}
    if prefix == "lp_":
        b = np.load(os.path.join(GOLD, "boeing2_lp.npz"))
        G = sp.csc_matrix((b["Gx"], b["Gi"], b["Gp"]), shape=tuple(b["G_size"]))
        A = sp.csc_matrix((b["Ax"], b["Ai"], b["Ap"]), shape=tuple(b["A_size"]))
        H = None
    else:
        sys.path.insert(0, GOLD)
        from generators import qp_instance
        H, _, G, _ = qp_instance(50, 40, 50)
        A = sp.csc_matrix((0, H.shape[0]))
    return G, A, H, g


@pytest.mark.parametrize("prefix", ["lp_", "qp_"])
def test_kkt_ldl_matches_reference_sytrf(prefix):
    """kkt.ldl against the UNMODIFIED reference's misc.kkt_ldl (dense sytrf/sytrs), fixtures made by
    tests/golden/make_kkt_ldl_fixtures.py"""
    from kvxopt_b200 import kkt
    G, A, H, g = _kkt_case(prefix)
    m, n = G.shape
    factor = kkt.ldl(G, {"l": m, "q": [], "s": []}, A)
    W = {"d": g["d"].copy(), "di": 1.0 / g["d"]}
    for _ in range(2):                                     # second pass: values-only refactorization
        solve = factor(W, H)
        x, y, zz = g["bx"].copy(), g["by"].copy(), g["bz"].copy()
        solve(x, y, zz)
        for got, want in ((x, g["ux"]), (y, g["uy"]), (zz, g["uz"])):
            if len(want):
                assert np.linalg.norm(got - want) <= XREL_TOL * np.linalg.norm(want)
    # regularised variant (coneprog.py:430-434): unconstrained fill-reducing order on the quasi-definite matrix
    from oracle import ldl_oracle
    reg = 1e-6
    solve = kkt.ldl(G, {"l": m, "q": [], "s": []}, A, kktreg=reg)(W, H)
    x, y, zz = g["bx"].copy(), g["by"].copy(), g["bz"].copy()
    solve(x, y, zz)
    p = A.shape[0]
    Kd = np.zeros((n + p + m, n + p + m))
    if H is not None:
        Kd[:n, :n] = H.toarray()
    Kd[n:n + p, :n] = A.toarray(); Kd[n + p:, :n] = (1.0 / g["d"])[:, None] * G.toarray()
    Kd = np.tril(Kd) + np.tril(Kd, -1).T
    Kd[np.arange(n), np.arange(n)] += reg
    Kd[np.arange(n, n + p + m), np.arange(n, n + p + m)] -= reg
    Kd[np.arange(n + p, n + p + m), np.arange(n + p, n + p + m)] -= 1.0
    want = np.linalg.solve(Kd, np.concatenate([g["bx"], g["by"], g["bz"] / g["d"]]))
    got = np.concatenate([x, y, zz])
    assert np.linalg.norm(got - want) <= 1e-9 * np.linalg.norm(want)


def test_boeing2_ipm_with_sparse_ldl(kvx):
    """BASELINE configs[2] through the 'ldl'-style solver: same iteration count and objective as the reference's
    dense 'ldl' (tests/golden/boeing2_lp.npz: 29 iterations, objective to 1e-8)"""
    from kvxopt import matrix, spmatrix, solvers
    from kvxopt_b200 import kkt
    z = np.load(os.path.join(GOLD, "boeing2_lp.npz"))
    def spm(p, i, x, size):
        cols = np.repeat(np.arange(size[1]), np.diff(p))
        return spmatrix(x.tolist(), i.tolist(), cols.tolist(), tuple(int(s) for s in size))
    G = spm(z["Gp"], z["Gi"], z["Gx"], z["G_size"]); A = spm(z["Ap"], z["Ai"], z["Ax"], z["A_size"])
    c, h, b = matrix(z["c"]), matrix(z["h"]), matrix(z["b"])
    dims = {"l": G.size[0], "q": [], "s": []}
    sol = solvers.conelp(c, G, h, dims, A, b, kktsolver=kkt.ldl(G, dims, A))
    assert sol["status"] == "optimal"
    assert sol["iterations"] == int(z["iters_ldl"])
    assert abs(sol["primal objective"] - float(z["pobj_ldl"])) <= 1e-8 * abs(float(z["pobj_ldl"]))


def test_ldl_edge_cases(cholmod):
    """1 x 1 negative matrix, a negative definite matrix (every pivot negative) and an empty one"""
    K1 = sp.csc_matrix(np.array([[-2.0]]))
    x = np.array([[3.0]], order="F"); cholmod.linsolve(K1, x)
    assert abs(x[0, 0] + 1.5) <= 4e-16 * 1.5          # square-root form: l = sqrt(2), x = -(3 / l) / l
    A = rand_spd(200, 0.05, 8)
    Kn = (-A).tocsc()
    B = np.random.default_rng(4).standard_normal((200, 2))
    X = np.asfortranarray(B.copy()); cholmod.linsolve(lower_ccs(Kn), X)
    assert berr(Kn, X, B) <= BERR_TOL
    F = cholmod.symbolic(lower_ccs(Kn)); cholmod.numeric(lower_ccs(Kn), F)
    D = np.asfortranarray(np.ones((200, 1))); cholmod.solve(F, D, sys=6)
    assert (D < 0).all()
    E = sp.csc_matrix((0, 0))
    F0 = cholmod.symbolic(E); cholmod.numeric(E, F0)
    cholmod.solve(F0, np.zeros((0, 1), order="F"))


@pytest.mark.parametrize("kktreg", [None, 1e-10])
def test_qp_mini_coneqp_with_sparse_ldl(kvx, kktreg):
    """BASELINE config 5 generator at reduced size through the reference coneqp with kkt.ldl: the iteration count and
    objective of the reference's own run (tests/golden/qp_mini.npz, dense 'chol')"""
    sys.path.insert(0, GOLD)
    from generators import qp_instance
    from kvxopt import matrix, spmatrix, solvers
    from kvxopt_b200 import kkt
    z = np.load(os.path.join(GOLD, "qp_mini.npz"))
    P, q, G, h = qp_instance(int(z["nx"]), int(z["ny"]), int(z["nrand"]))
    def spm(M):
        M = sp.coo_matrix(M)
        return spmatrix(M.data.tolist(), M.row.tolist(), M.col.tolist(), M.shape)
    Pk, Gk = spm(sp.tril(P)), spm(G)
    dims = {"l": Gk.size[0], "q": [], "s": []}
    f3 = kkt.ldl(Gk, dims, spmatrix([], [], [], (0, Pk.size[0])), kktreg=kktreg)
    sol = solvers.coneqp(Pk, matrix(q), Gk, matrix(h), dims, kktsolver=lambda W: f3(W, Pk))
    assert sol["status"] == "optimal"
    assert sol["iterations"] == int(z["iters"])
    assert abs(sol["primal objective"] - float(z["pobj"])) <= 1e-8 * abs(float(z["pobj"]))
    assert f3.info()["factorizations"] >= sol["iterations"]


@pytest.mark.parametrize("prefix", ["lp_", "qp_"])
def test_kkt_ldl2_matches_reference_sytrf(prefix):
    """kkt.ldl2 (2 x 2 system, misc.py:1128-1210) returns the same (ux, uy, W uz) as the reference's dense solvers: the golden
    vectors of misc.kkt_ldl, which solves the same KKT system"""
    from kvxopt_b200 import kkt
    G, A, H, g = _kkt_case(prefix)
    m, n = G.shape
    factor = kkt.ldl2(G, {"l": m, "q": [], "s": []}, A)
    W = {"d": g["d"].copy(), "di": 1.0 / g["d"]}
    for _ in range(2):
        solve = factor(W, H)
        x, y, zz = g["bx"].copy(), g["by"].copy(), g["bz"].copy()
        solve(x, y, zz)
        for got, want in ((x, g["ux"]), (y, g["uy"]), (zz, g["uz"])):
            if len(want):
                assert np.linalg.norm(got - want) <= XREL_TOL * np.linalg.norm(want)


def test_boeing2_ipm_with_sparse_ldl2(kvx):
    from kvxopt import matrix, spmatrix, solvers
    from kvxopt_b200 import kkt
    z = np.load(os.path.join(GOLD, "boeing2_lp.npz"))
    def spm(p, i, x, size):
        cols = np.repeat(np.arange(size[1]), np.diff(p))
        return spmatrix(x.tolist(), i.tolist(), cols.tolist(), tuple(int(s) for s in size))
    G = spm(z["Gp"], z["Gi"], z["Gx"], z["G_size"]); A = spm(z["Ap"], z["Ai"], z["Ax"], z["A_size"])
    c, h, b = matrix(z["c"]), matrix(z["h"]), matrix(z["b"])
    dims = {"l": G.size[0], "q": [], "s": []}
    sol = solvers.conelp(c, G, h, dims, A, b, kktsolver=kkt.ldl2(G, dims, A))
    assert sol["status"] == "optimal"
    assert sol["iterations"] == int(z["iters_ldl"])
    assert abs(sol["primal objective"] - float(z["pobj_ldl"])) <= 1e-8 * abs(float(z["pobj_ldl"]))


def test_value_assembler_matches_scipy():
    """b200s_spmv_*: y = M w on the device (the linear map from [di^2; H; A] to the stored entries of K that kkt.ldl2 uses),
    against scipy; rows without terms give 0; bad indices are rejected."""
    import ctypes as C
    import scipy.sparse as sp
    from kvxopt_b200 import _lib as L
    fn = L.fn
    rng = np.random.default_rng(4)
    M = sp.random(5000, 700, density=0.004, random_state=rng, format="csr"); M.sort_indices()
    rp = M.indptr.astype(np.int64); ci = M.indices.astype(np.int64); vx = M.data.astype(np.float64)
    h = C.c_void_p()
    assert fn["b200s_spmv_create"](M.shape[0], M.shape[1], L.ptr_i64(rp), L.ptr_i64(ci), L.ptr_f64(vx), C.byref(h)) == 0
    for _ in range(2):
        w = rng.standard_normal(M.shape[1])
        yd = C.c_void_p()
        assert fn["b200s_spmv_apply"](h, L.ptr_f64(w), C.byref(yd)) == 0 and yd.value
        y = np.full(M.shape[0], np.nan)
        assert fn["b200s_spmv_get"](h, L.ptr_f64(y)) == 0
        ref = M @ w
        assert np.abs(y - ref).max() <= 1e-13 * max(1.0, np.abs(ref).max())
        assert (y[np.diff(rp) == 0] == 0).all()
    fn["b200s_spmv_free"](h)
    bad = ci.copy(); bad[3] = M.shape[1]
    h2 = C.c_void_p()
    assert fn["b200s_spmv_create"](M.shape[0], M.shape[1], L.ptr_i64(rp), L.ptr_i64(bad), L.ptr_f64(vx), C.byref(h2)) == L.INVALID
