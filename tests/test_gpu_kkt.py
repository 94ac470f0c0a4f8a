"""Device-side reduced KKT solver (kvxopt_b200.kkt) against (i) the dense solution of the
3 x 3 block system it is documented to solve (reference src/python/misc.py:1367-1379), (ii) the reference's own
misc.kkt_chol2 (probe build oracle/_ref, CPU oracle behind kvxopt.cholmod) on the same W, and (iii) inside the
reference's interior-point drivers: identical iteration counts, objectives within 1e-8 (BASELINE north_star)."""
import sys

import numpy as np
import pytest
import scipy.sparse as sp

from conftest import GOLD

pytestmark = pytest.mark.gpu


def to_spmatrix(kvx, M):
    M = sp.coo_matrix(M)
    return kvx.spmatrix(M.data.tolist(), M.row.tolist(), M.col.tolist(), M.shape)


def problem(seed, n, ml, p, with_H, dens=0.15):
    rng = np.random.default_rng(seed)
    G = sp.vstack([sp.identity(n), sp.random(ml - n, n, density=dens, random_state=rng)]).tocsc()
    A = sp.random(p, n, density=0.4, random_state=rng).tocsc() if p else sp.csc_matrix((0, n))
    if p:
        A = (A + sp.csc_matrix((np.ones(p), (np.arange(p), np.arange(p))), shape=(p, n))).tocsc()      # full row rank
    H = None
    if with_H:
        M = sp.random(n, n, density=0.1, random_state=rng)
        H = (M @ M.T + 0.1 * sp.identity(n)).tocsc()
    d = rng.uniform(0.3, 3.0, ml)
    return G, A, H, d, rng


@pytest.mark.parametrize("n,ml,p,with_H", [(30, 70, 0, False), (30, 70, 4, False), (50, 120, 0, True), (50, 120, 7, True),
                                            (400, 900, 12, True)])
def test_block_system_solution(kvx, n, ml, p, with_H):
    from kvxopt import matrix
    from kvxopt_b200 import kkt
    G, A, H, d, rng = problem(n + p, n, ml, p, with_H)
    W = {"d": matrix(d), "di": matrix(1.0 / d)}
    factor = kkt.chol2(to_spmatrix(kvx, G), {"l": ml, "q": [], "s": []}, to_spmatrix(kvx, A))
    Hk = to_spmatrix(kvx, sp.tril(H)) if with_H else None
    for rep in range(2):                      # second factor call re-uses the pattern with new scalings
        solve = factor(W, Hk)
        bx, by, bz = rng.standard_normal(n), rng.standard_normal(p), rng.standard_normal(ml)
        x, y, z = matrix(bx), matrix(by) if p else matrix(0.0, (0, 1)), matrix(bz)
        solve(x, y, z)
        Hd = H.toarray() if with_H else np.zeros((n, n))
        Kd = np.block([[Hd, A.toarray().T, G.toarray().T],
                       [A.toarray(), np.zeros((p, p)), np.zeros((p, ml))],
                       [G.toarray(), np.zeros((ml, p)), -np.diag(d * d)]])
        u = np.linalg.solve(Kd, np.concatenate([bx, by, bz]))
        ux, uy, uz = u[:n], u[n:n + p], u[n + p:]
        scale = np.linalg.norm(u)
        assert np.linalg.norm(np.array(x).ravel() - ux) <= 1e-10 * scale
        if p:
            assert np.linalg.norm(np.array(y).ravel() - uy) <= 1e-10 * scale
        assert np.linalg.norm(np.array(z).ravel() - d * uz) <= 1e-10 * np.linalg.norm(d * uz) + 1e-10 * scale
        d = rng.uniform(0.3, 3.0, ml)
        W = {"d": matrix(d), "di": matrix(1.0 / d)}


def test_matches_reference_kkt_chol2_and_singular_first_call(kvx):
    """same W, same right-hand sides: reference misc.kkt_chol2 (CPU oracle behind kvxopt.cholmod) vs the device solver,
    including the branch where the first S is singular and both switch to S + A'A (misc.py:1427-1447)"""
    from kvxopt import matrix
    import kvxopt.misc as misc
    from kvxopt_b200 import kkt
    from oracle import cholmod_cpu
    n, ml, p = 24, 20, 6
    rng = np.random.default_rng(3)
    Gd = np.zeros((ml, n)); Gd[np.arange(ml), np.arange(ml)] = rng.uniform(1, 2, ml)       # columns ml..n-1 of G are empty: S singular
    Ad = rng.standard_normal((p, n)) * (rng.uniform(size=(p, n)) < 0.5)
    Ad[:, ml:] += np.eye(p, n - ml) + 0.3                                              # A covers the null space of G'G
    G, A = sp.csc_matrix(Gd), sp.csc_matrix(Ad)
    d = rng.uniform(0.5, 2.0, ml)
    W = {"d": matrix(d), "di": matrix(1.0 / d), "dnl": matrix(0.0, (0, 1)), "dnli": matrix(0.0, (0, 1)), "r": [], "rti": [],
         "v": [], "beta": []}
    saved = misc.cholmod
    misc.cholmod = cholmod_cpu
    try:
        fref = misc.kkt_chol2(to_spmatrix(kvx, G), {"l": ml, "q": [], "s": []}, to_spmatrix(kvx, A))
        sref = fref(W)
        x0, y0, z0 = matrix(rng.standard_normal(n)), matrix(rng.standard_normal(p)), matrix(rng.standard_normal(ml))
        xr, yr, zr = +x0, +y0, +z0
        sref(xr, yr, zr)
    finally:
        misc.cholmod = saved
    fgpu = kkt.chol2(to_spmatrix(kvx, G), {"l": ml, "q": [], "s": []}, to_spmatrix(kvx, A))
    sgpu = fgpu(W)
    assert fgpu.info()["singular_mode"] == 1
    xg, yg, zg = +x0, +y0, +z0
    sgpu(xg, yg, zg)
    for a, b in ((xg, xr), (yg, yr), (zg, zr)):
        a, b = np.array(a).ravel(), np.array(b).ravel()
        assert np.linalg.norm(a - b) <= 1e-10 * max(np.linalg.norm(b), 1.0)


def test_boeing2_lp_with_device_kkt_solver(kvx):
    """BASELINE config 3 with the KKT plug-in API: same 29 iterations and objective as the reference's solvers"""
    from kvxopt import matrix, solvers
    from kvxopt_b200 import kkt
    z = np.load(GOLD + "/boeing2_lp.npz")
    G = sp.csc_matrix((z["Gx"], z["Gi"], z["Gp"]), shape=tuple(z["G_size"]))
    A = sp.csc_matrix((z["Ax"], z["Ai"], z["Ap"]), shape=tuple(z["A_size"]))
    c, h, b = matrix(z["c"]), matrix(z["h"]), matrix(z["b"])
    Gk, Ak = to_spmatrix(kvx, G), to_spmatrix(kvx, A)
    sol = solvers.lp(c, Gk, h, Ak, b, kktsolver=kkt.lp_kktsolver(Gk, Ak))
    assert sol["status"] == "optimal"
    assert sol["iterations"] == int(z["iters_chol2"]) == 29
    assert abs(sol["primal objective"] - float(z["pobj_chol2"])) <= 1e-8 * abs(float(z["pobj_chol2"]))
    np.testing.assert_allclose(np.array(sol["x"]).ravel(), z["x_ref"], rtol=1e-5, atol=1e-6)


def test_qp_mini_with_device_kkt_solver(kvx):
    """BASELINE config 5 generator at reduced size through coneqp with the device KKT solver"""
    sys.path.insert(0, GOLD)
    from generators import qp_instance
    from kvxopt import matrix, solvers
    from kvxopt_b200 import kkt
    z = np.load(GOLD + "/qp_mini.npz")
    P, q, G, h = qp_instance(int(z["nx"]), int(z["ny"]), int(z["nrand"]))
    Pk, Gk = to_spmatrix(kvx, sp.tril(P)), to_spmatrix(kvx, G)
    sol = solvers.qp(Pk, matrix(q), Gk, matrix(h), kktsolver=kkt.qp_kktsolver(Pk, Gk))
    assert sol["status"] == "optimal"
    assert sol["iterations"] == int(z["iters"])
    assert abs(sol["primal objective"] - float(z["pobj"])) <= 1e-8 * abs(float(z["pobj"]))


def test_errors_follow_the_reference(kvx):
    from kvxopt import matrix
    from kvxopt_b200 import kkt
    G = to_spmatrix(kvx, sp.identity(3).tocsc())
    A = kvx.spmatrix([], [], [], (0, 3), "d")
    with pytest.raises(ValueError):
        kkt.chol2(G, {"l": 3, "q": [2], "s": []}, A)                 # misc.py:1380-1383
    with pytest.raises(ValueError):
        kkt.chol2(G, {"l": 3, "q": [], "s": []}, A, mnl=1)
    factor = kkt.chol2(G, {"l": 3, "q": [], "s": []}, A)
    W = {"d": matrix([1.0, 1.0, 1.0]), "di": matrix([1.0, 1.0, 1.0])}
    Hneg = kvx.spmatrix([-5.0, -5.0, -5.0], [0, 1, 2], [0, 1, 2], (3, 3))
    with pytest.raises(ArithmeticError):
        factor(W, Hneg)                                             # S = -5 I + I: not positive definite, A is empty


def test_solve_outlives_the_factor_closure(kvx):
    """kkt.chol2(...)(W) used as a temporary: the returned solve() must keep the device object alive"""
    import gc
    from kvxopt import matrix
    from kvxopt_b200 import kkt
    G = to_spmatrix(kvx, sp.identity(4).tocsc() * 2.0)
    A = kvx.spmatrix([], [], [], (0, 4), "d")
    W = {"d": matrix([1.0] * 4), "di": matrix([1.0] * 4)}
    solve = kkt.chol2(G, {"l": 4, "q": [], "s": []}, A)(W)
    gc.collect()
    x, y, z = matrix([1.0, 2.0, 3.0, 4.0]), matrix(0.0, (0, 1)), matrix([0.0] * 4)
    solve(x, y, z)                                    # S = 4 I, bz = 0  =>  ux = bx / 4, W uz = G ux
    np.testing.assert_allclose(np.array(x).ravel(), [0.25, 0.5, 0.75, 1.0], rtol=1e-14)
    np.testing.assert_allclose(np.array(z).ravel(), [0.5, 1.0, 1.5, 2.0], rtol=1e-14)
