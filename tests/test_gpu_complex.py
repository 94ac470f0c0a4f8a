"""Complex Hermitian ('z') matrices through the cholmod mirror (reference src/C/cholmod.c:144,153,286-290,343-357,460-464:
CHOLMOD_COMPLEX, capsule names "CHOLMOD SYM Z FACTOR L/U").  The CUDA engine factors the real symmetric embedding
a + ib -> [[a, -b], [b, a]]; parity here is against numpy's dense complex Cholesky / solves on the same matrix.
Tolerances as for the real case: backward error <= 1e-12, solution relative difference <= 1e-10."""
import numpy as np
import pytest
import scipy.sparse as sp

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def cholmod():
    from kvxopt_b200 import cholmod as m, _lib
    assert _lib.device_count() > 0, "GPU tests need a CUDA device; there is no CPU fallback"
    return m


def rand_hpd(n, dens, seed):
    rng = np.random.default_rng(seed)
    M = sp.random(n, n, density=dens, random_state=rng, format="csc")
    N = sp.random(n, n, density=dens, random_state=rng, format="csc")
    Z = (M + 1j * N).tocsc()
    A = (Z + Z.getH()).tocsc()
    d = np.asarray(abs(A).sum(axis=1)).ravel() + 1.0
    return (A + sp.diags(d)).tocsc()


def lower(A):
    Al = sp.tril(A).tocsc(); Al.sort_indices()
    return Al


def upper(A):
    Au = sp.triu(A).tocsc(); Au.sort_indices()
    return Au


@pytest.mark.parametrize("n,dens,seed", [(1, 1.0, 0), (7, 0.5, 1), (60, 0.08, 2), (400, 0.01, 3), (1500, 0.004, 4)])
@pytest.mark.parametrize("uplo", ["L", "U"])
def test_linsolve_and_factor_against_numpy(cholmod, n, dens, seed, uplo):
    A = rand_hpd(n, dens, seed)
    At = lower(A) if uplo == "L" else upper(A)
    rng = np.random.default_rng(seed + 10)
    B = np.asfortranarray(rng.standard_normal((n, 3)) + 1j * rng.standard_normal((n, 3)))
    X = B.copy(order="F")
    cholmod.linsolve(At, X, uplo=uplo)
    Ad = A.toarray()
    Xr = np.linalg.solve(Ad, B)
    assert np.linalg.norm(X - Xr) / np.linalg.norm(Xr) < 1e-10
    res = np.linalg.norm(Ad @ X - B) / (np.abs(Ad).sum(axis=0).max() * np.linalg.norm(X) + np.linalg.norm(B))
    assert res < 1e-12
    # symbolic / numeric / getfactor / diag: P A P' = L L^H with a real positive diagonal
    F = cholmod.symbolic(At, uplo=uplo)
    cholmod.numeric(At, F)
    p = cholmod.factor_perm(F)
    assert sorted(p.tolist()) == list(range(n))
    Lf = cholmod.getfactor(F)
    assert Lf.dtype == np.complex128
    Ld = Lf.toarray()
    assert np.allclose(np.triu(Ld, 1), 0)
    PAP = Ad[np.ix_(p, p)]
    assert np.linalg.norm(Ld @ Ld.conj().T - PAP) / np.linalg.norm(PAP) < 1e-13
    d = cholmod.diag(F)
    assert d.dtype == np.complex128 and np.all(d.imag == 0) and np.all(d.real > 0)
    np.testing.assert_allclose(d.real.ravel(), np.diag(Ld).real, rtol=1e-13)
    # entrywise against numpy's dense complex Cholesky of the permuted matrix (unique factor with positive diagonal)
    Lnp = np.linalg.cholesky(PAP)
    assert np.abs(Ld - Lnp).max() / np.abs(Lnp).max() < 1e-11


def test_all_nine_systems(cholmod):
    n = 300
    A = rand_hpd(n, 0.02, 7)
    Al = lower(A)
    F = cholmod.symbolic(Al)
    cholmod.numeric(Al, F)
    p = cholmod.factor_perm(F)
    Ld = cholmod.getfactor(F).toarray()
    P = np.eye(n)[p]
    rng = np.random.default_rng(8)
    B = np.asfortranarray(rng.standard_normal((n, 2)) + 1j * rng.standard_normal((n, 2)))
    LH = Ld.conj().T
    ref = {0: np.linalg.solve(A.toarray(), B), 1: np.linalg.solve(Ld @ LH, B), 2: np.linalg.solve(Ld, B),
           3: np.linalg.solve(LH, B), 4: np.linalg.solve(Ld, B), 5: np.linalg.solve(LH, B), 6: B.copy(),
           7: P @ B, 8: P.T @ B}
    for sys, Xr in ref.items():
        X = B.copy(order="F")
        cholmod.solve(F, X, sys=sys)
        assert np.linalg.norm(X - Xr) / np.linalg.norm(Xr) < 1e-10, sys


def test_ldB_offset_nrhs(cholmod):
    n = 50
    A = rand_hpd(n, 0.1, 11)
    Al = lower(A)
    F = cholmod.symbolic(Al)
    cholmod.numeric(Al, F)
    rng = np.random.default_rng(12)
    ld, off = n + 5, 3
    buf = rng.standard_normal(off + 2 * ld + n) + 1j * rng.standard_normal(off + 2 * ld + n)
    keep = buf.copy()
    cholmod.solve(F, buf, sys=0, nrhs=2, ldB=ld, offsetB=off)
    Ad = A.toarray()
    for j in range(2):
        b = keep[off + j * ld: off + j * ld + n]
        x = buf[off + j * ld: off + j * ld + n]
        assert np.linalg.norm(x - np.linalg.solve(Ad, b)) / np.linalg.norm(x) < 1e-10
    mask = np.ones(buf.size, bool)
    for j in range(2):
        mask[off + j * ld: off + j * ld + n] = False
    assert np.array_equal(buf[mask], keep[mask])          # nothing outside the addressed columns is touched


def test_spsolve_and_splinsolve(cholmod):
    n = 200
    A = rand_hpd(n, 0.02, 21)
    Al = lower(A)
    rng = np.random.default_rng(22)
    Bs = sp.random(n, 4, density=0.03, random_state=rng, format="csc")
    Bs = (Bs + 1j * sp.random(n, 4, density=0.03, random_state=rng, format="csc")).tocsc()
    Bs.sort_indices()
    X = cholmod.splinsolve(Al, Bs)
    Xr = np.linalg.solve(A.toarray(), Bs.toarray())
    assert X.dtype == np.complex128
    assert np.linalg.norm(X.toarray() - Xr) / np.linalg.norm(Xr) < 1e-10
    F = cholmod.symbolic(Al)
    cholmod.numeric(Al, F)
    Ld = cholmod.getfactor(F).toarray()
    Y = cholmod.spsolve(F, Bs, sys=4)
    Yr = np.linalg.solve(Ld, Bs.toarray())
    assert np.linalg.norm(Y.toarray() - Yr) / np.linalg.norm(Yr) < 1e-10


def test_type_contract(cholmod):
    # cholmod.c:343-357 / 460-464: a 'z' factor takes 'z' matrices and right-hand sides only, and the other way round
    n = 20
    A = rand_hpd(n, 0.2, 31)
    Al = lower(A)
    Fz = cholmod.symbolic(Al)
    Ar = lower((A.real + sp.identity(n)).tocsc())
    Fd = cholmod.symbolic(Ar)
    with pytest.raises(TypeError):
        cholmod.numeric(Ar, Fz)
    with pytest.raises(TypeError):
        cholmod.numeric(Al, Fd)
    cholmod.numeric(Al, Fz)
    cholmod.numeric(Ar, Fd)
    with pytest.raises(TypeError):
        cholmod.solve(Fz, np.zeros(n))
    with pytest.raises(TypeError):
        cholmod.solve(Fd, np.zeros(n, dtype=complex))
    with pytest.raises(TypeError):
        cholmod.linsolve(Al, np.zeros(n))
    # not positive definite: ArithmeticError carries the (complex) column index
    Bad = Al.copy().tolil(); Bad[5, 5] = -1.0; Bad = Bad.tocsc(); Bad.sort_indices()
    with pytest.raises(ArithmeticError) as e:
        cholmod.linsolve(Bad, np.zeros(n, dtype=complex), p=np.arange(n))
    assert e.value.args[0] == 5


def test_reference_types_complex(cholmod, kvx):
    """kvxopt's own 'z' spmatrix / matrix through the mirror: capsule name and result types as at cholmod.c:286-290, 923-924, 973"""
    from kvxopt import matrix, spmatrix, cholmod as kc
    A = spmatrix([10, 3 + 1j, 5, -2 - 2j, 5, 2], [0, 2, 1, 3, 2, 3], [0, 0, 1, 1, 2, 3], (4, 4), "z")
    Ad = np.array(matrix(A))
    Ad = np.tril(Ad) + np.tril(Ad, -1).conj().T
    b = matrix([1 + 1j, 2, 3 - 2j, 4j])
    x = +b
    kc.linsolve(A, x)
    np.testing.assert_allclose(np.array(x).ravel(), np.linalg.solve(Ad, np.array(b).ravel()), rtol=1e-12)
    F = kc.symbolic(A)
    assert "CHOLMOD SYM Z FACTOR L" in repr(F)
    kc.numeric(A, F)
    d = kc.diag(F)
    assert d.typecode == "z" and d.size == (4, 1)
    Lf = kc.getfactor(F)
    assert Lf.typecode == "z"
    x = +b
    kc.solve(F, x)
    np.testing.assert_allclose(np.array(x).ravel(), np.linalg.solve(Ad, np.array(b).ravel()), rtol=1e-12)
