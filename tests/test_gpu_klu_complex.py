"""Complex ('z') matrices through the klu mirror (reference src/C/klu.c:161-162,348-355,468-479,661-668,754-813: klu_zl_*).
The complex threshold-pivoting factorization of the host pivot search is what get_numeric / get_det return; solves run on the
device through the real embedding of order 2n.  Checked against numpy / scipy on the same matrices; the reference's own
complex test loops run verbatim in test_gpu_reference_suite.py."""
import numpy as np
import pytest
import scipy.sparse as sp
import scipy.sparse.linalg as spla

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def klu():
    from kvxopt_b200 import klu as m, _lib
    assert _lib.device_count() > 0, "GPU tests need a CUDA device; there is no CPU fallback"
    return m


def rand_z(n, dens, seed, blocks=False):
    rng = np.random.default_rng(seed)
    M = sp.random(n, n, density=dens, random_state=rng, format="csc") + 1j * sp.random(n, n, density=dens, random_state=rng, format="csc")
    D = sp.diags(rng.uniform(1, 2, n) * np.exp(1j * rng.uniform(0, 2 * np.pi, n)))      # pivots with arbitrary phase
    A = (M + D).tocsc()
    if blocks:       # block upper triangular: several BTF blocks and an F part
        A = sp.triu(A, 1).tocsc() * 0.5 + sp.block_diag([A[i:i + n // 4, i:i + n // 4] for i in range(0, n - n % (n // 4), n // 4)] +
                                                       ([A[n - n % (n // 4):, n - n % (n // 4):]] if n % (n // 4) else [])).tocsc()
    A = A.tocsc(); A.sort_indices()
    return A


@pytest.mark.parametrize("n,dens,seed,blocks", [(1, 1.0, 0, False), (8, 0.4, 1, False), (120, 0.05, 2, False), (600, 0.01, 3, False),
                                                (200, 0.04, 4, True)])
def test_factor_identity_solves_and_determinant(klu, n, dens, seed, blocks):
    A = rand_z(n, dens, seed, blocks)
    Fs = klu.symbolic(A)
    Fn = klu.numeric(A, Fs)
    Lm, Um, P, Q, R, Fm, r = klu.get_numeric(A, Fs, Fn)
    assert Lm.dtype == np.complex128 and Um.dtype == np.complex128
    scale = abs(A).sum(axis=0).max()
    assert abs(R @ P @ A @ Q - (Lm @ Um + Fm)).sum(axis=0).max() <= 1e-13 * scale          # R P A Q = L U + F (klu.c:392-566)
    assert np.allclose(Lm.diagonal(), 1.0) and abs(sp.triu(Lm, 1)).sum() == 0 and abs(sp.tril(Um, -1)).sum() == 0
    if blocks:
        assert len(r) - 1 >= 2
    # |L| <= 1 / tol: threshold partial pivoting on |z|
    assert abs(Lm).max() <= 1e3 * (1 + 1e-12)
    rng = np.random.default_rng(seed + 100)
    B = np.asfortranarray(rng.standard_normal((n, 2)) + 1j * rng.standard_normal((n, 2)))
    Ad = A.toarray()
    for trans, M in (("N", Ad), ("T", Ad.T), ("C", Ad.conj().T)):
        X = B.copy(order="F")
        klu.solve(A, Fs, Fn, X, trans=trans)
        Xr = np.linalg.solve(M, B)
        assert np.linalg.norm(X - Xr) <= 1e-9 * np.linalg.norm(Xr), trans
        X2 = B.copy(order="F")
        klu.linsolve(A, X2, trans=trans)
        assert np.linalg.norm(X2 - Xr) <= 1e-9 * np.linalg.norm(Xr), trans
    d = klu.get_det(A, Fs, Fn)
    dr = np.linalg.det(Ad)
    assert abs(d - dr) <= 1e-9 * abs(dr)


def test_ldB_offset_and_type_contract(klu):
    n = 40
    A = rand_z(n, 0.1, 9)
    Fs = klu.symbolic(A); Fn = klu.numeric(A, Fs)
    rng = np.random.default_rng(10)
    ld, off = n + 3, 2
    buf = rng.standard_normal(off + ld + n) + 1j * rng.standard_normal(off + ld + n)
    keep = buf.copy()
    klu.solve(A, Fs, Fn, buf, nrhs=2, ldB=ld, offsetB=off)
    Ad = A.toarray()
    for j in range(2):
        b = keep[off + j * ld: off + j * ld + n]
        assert np.linalg.norm(buf[off + j * ld: off + j * ld + n] - np.linalg.solve(Ad, b)) <= 1e-9 * np.linalg.norm(b)
    mask = np.ones(buf.size, bool)
    for j in range(2):
        mask[off + j * ld: off + j * ld + n] = False
    assert np.array_equal(buf[mask], keep[mask])
    # klu.c:348-355, 422-425: capsules carry the numerical type
    Ar = sp.csc_matrix(A.real + sp.identity(n) * 3); Ar.sort_indices()
    Fsr = klu.symbolic(Ar); Fnr = klu.numeric(Ar, Fsr)
    with pytest.raises(TypeError):
        klu.numeric(A, Fsr)
    with pytest.raises(TypeError):
        klu.numeric(Ar, Fs)
    with pytest.raises(TypeError):
        klu.solve(A, Fs, Fnr, np.zeros(n, dtype=complex))
    with pytest.raises(TypeError):
        klu.solve(A, Fs, Fn, np.zeros(n))
    with pytest.raises(TypeError):
        klu.get_numeric(Ar, Fs, Fn)
    # singular complex matrix
    S = A.tolil(); S[:, 3] = 0; S = S.tocsc(); S.sort_indices()
    with pytest.raises(ArithmeticError):
        klu.numeric(S, klu.symbolic(S))
