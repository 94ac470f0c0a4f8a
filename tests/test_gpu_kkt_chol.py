"""kvxopt_b200.kkt.chol -- the device counterpart of the reference's dense 'chol' KKT solver misc.kkt_chol
(reference src/python/misc.py:1213-1349; named by BASELINE.json configs[2]) -- against (i) the dense solution of the
3 x 3 block system, (ii) the UNMODIFIED reference misc.kkt_chol (probe build oracle/_ref, LAPACK) on the same W and
right-hand sides, (iii) boeing2 through the reference's conelp: 29 iterations, objective within 1e-8 (north_star)."""
import numpy as np
import pytest
import scipy.sparse as sp

from conftest import GOLD

pytestmark = pytest.mark.gpu


def problem(seed, n, ml, p, with_H):
    rng = np.random.default_rng(seed)
    G = np.vstack([np.eye(n), rng.standard_normal((ml - n, n)) * (rng.uniform(size=(ml - n, n)) < 0.3)])
    A = rng.standard_normal((p, n))
    H = None
    if with_H:
        M = rng.standard_normal((n, n)) * (rng.uniform(size=(n, n)) < 0.2)
        H = M @ M.T + 0.1 * np.eye(n)
    d = rng.uniform(0.3, 3.0, ml)
    return G, A, H, d, rng


@pytest.mark.parametrize("n,ml,p,with_H", [(30, 70, 0, False), (30, 70, 4, False), (50, 120, 7, True), (143, 352, 4, False),
                                            (300, 700, 20, True), (64, 64, 64, True)])
def test_block_system_and_reference_kkt_chol(kvx, n, ml, p, with_H):
    from kvxopt import matrix, misc
    from kvxopt_b200 import kkt
    G, A, H, d, rng = problem(n + 7 * p, n, ml, p, with_H)
    dims = {"l": ml, "q": [], "s": []}
    Gk, Ak = matrix(G), matrix(A) if p else matrix(0.0, (0, n))
    Hk = matrix(H) if with_H else None
    factor = kkt.chol(Gk, dims, Ak)
    ref_factor = misc.kkt_chol(Gk, dims, Ak)
    for rep in range(2):
        W = {"d": matrix(d), "di": matrix(1.0 / d), "r": [], "rti": [], "v": [], "beta": []}
        solve = factor(W, Hk)
        ref_solve = ref_factor(W, Hk)
        bx, by, bz = rng.standard_normal(n), rng.standard_normal(p), rng.standard_normal(ml)
        x, y, z = matrix(bx), matrix(by) if p else matrix(0.0, (0, 1)), matrix(bz)
        solve(x, y, z)
        xr, yr, zr = matrix(bx), matrix(by) if p else matrix(0.0, (0, 1)), matrix(bz)
        ref_solve(xr, yr, zr)
        got = np.concatenate([np.array(x).ravel(), np.array(y).ravel(), np.array(z).ravel()])
        ref = np.concatenate([np.array(xr).ravel(), np.array(yr).ravel(), np.array(zr).ravel()])
        assert np.linalg.norm(got - ref) <= 1e-10 * np.linalg.norm(ref)          # solution relative difference (north_star)
        Hd = H if with_H else np.zeros((n, n))
        Kd = np.block([[Hd, A.T, G.T], [A, np.zeros((p, p)), np.zeros((p, ml))], [G, np.zeros((ml, p)), -np.diag(d * d)]])
        u = np.concatenate([got[:n + p], got[n + p:] / d])
        rhs = np.concatenate([bx, by, bz])
        berr = np.linalg.norm(Kd @ u - rhs) / (np.linalg.norm(Kd, 1) * np.linalg.norm(u) + np.linalg.norm(rhs))
        assert berr <= 1e-12
        d = rng.uniform(0.3, 3.0, ml)


def test_not_positive_definite_raises_like_potrf(kvx):
    from kvxopt import matrix
    from kvxopt_b200 import kkt
    n, ml = 20, 20
    G = np.eye(n)
    H = -3.0 * np.eye(n)        # H + G'D^2G indefinite for d = 1
    factor = kkt.chol(matrix(G), {"l": ml, "q": [], "s": []}, matrix(0.0, (0, n)))
    with pytest.raises(ArithmeticError):
        factor({"di": matrix(1.0, (ml, 1))}, matrix(H))


def test_graph_replayed_factor_and_solve_report_and_recover(kvx):
    """from its second call on factor(W) (H = None) and solve() are replayed CUDA graphs (kktd_gpu.cu): a scaling that makes
    K22 singular inside the replayed graph still raises ArithmeticError (the pivot word travels back through the staging
    buffer), the next factorization works again, and replayed solves give the same answers as the first, plainly launched one"""
    from kvxopt import matrix
    from kvxopt_b200 import kkt
    G, A, H, d, rng = problem(5, 143, 352, 4, False)
    n, ml, p = 143, 352, 4
    factor = kkt.chol(matrix(G), {"l": ml, "q": [], "s": []}, matrix(A))
    bx, by, bz = rng.standard_normal(n), rng.standard_normal(p), rng.standard_normal(ml)

    def run(dd):
        solve = factor({"di": matrix(1.0 / dd)})
        outs = []
        for _ in range(3):
            x, y, z = matrix(bx), matrix(by), matrix(bz)
            solve(x, y, z)
            outs.append(np.concatenate([np.array(x).ravel(), np.array(y).ravel(), np.array(z).ravel()]))
        return outs
    first = run(d)                                   # plain launches, then the captured solve
    assert np.array_equal(first[0], first[1]) and np.array_equal(first[1], first[2])
    second = run(d)                                  # the captured factorization
    assert np.array_equal(second[0], first[0])
    with pytest.raises(ArithmeticError):
        factor({"di": matrix(0.0, (ml, 1))})         # K = 0: zero pivot inside the replayed graph
    third = run(d)
    assert np.array_equal(third[0], first[0])


def test_rank_deficient_A_raises(kvx):
    from kvxopt import matrix
    from kvxopt_b200 import kkt
    A = np.zeros((2, 5)); A[0, 0] = 1.0          # second row zero
    with pytest.raises(ArithmeticError):
        kkt.chol(matrix(np.eye(5)), {"l": 5, "q": [], "s": []}, matrix(A))


def test_boeing2_lp_with_device_chol_kktsolver(kvx):
    """BASELINE configs[2]: solvers.lp on boeing2 via conelp with the 'chol' KKT solver -- here kkt.chol through the
    reference's kktsolver= plug-in API; same iteration count and objective as the reference's own 'chol'."""
    from kvxopt import matrix, solvers
    z = np.load(GOLD + "/boeing2_lp.npz")
    G = sp.csc_matrix((z["Gx"], z["Gi"], z["Gp"]), shape=tuple(z["G_size"]))
    A = sp.csc_matrix((z["Ax"], z["Ai"], z["Ap"]), shape=tuple(z["A_size"]))
    c, h, b = matrix(z["c"]), matrix(z["h"]), matrix(z["b"])
    from kvxopt_b200 import kkt
    Gd, Ad = matrix(G.toarray()), matrix(A.toarray())
    dims = {"l": G.shape[0], "q": [], "s": []}
    factor = kkt.chol(Gd, dims, Ad)
    sol = solvers.conelp(c, Gd, h, dims, Ad, b, kktsolver=factor)
    ref = solvers.conelp(c, Gd, h, dims, Ad, b, kktsolver="chol")
    assert sol["status"] == "optimal" and ref["status"] == "optimal"
    assert sol["iterations"] == ref["iterations"] == 29
    assert abs(sol["primal objective"] - ref["primal objective"]) <= 1e-8 * abs(ref["primal objective"])
    assert abs(sol["primal objective"] - (-315.0187296452)) <= 1e-8 * 315.02
    assert factor.info()["launches"] > 0
