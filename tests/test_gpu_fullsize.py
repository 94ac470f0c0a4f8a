"""BASELINE.json's full sizes, checked through size-independent properties (the oracle does not reach them in
seconds): configs[1] -- 4096 same-pattern ACTIVSg2000 matrices; configs[3] -- the 100^3 7-point Laplacian."""
import numpy as np
import pytest
import scipy.sparse as sp
import scipy.sparse.linalg as spla

from conftest import load_matrix, lap3d, lower_ccs

pytestmark = pytest.mark.gpu


def test_klu_batch_4096_properties():
    from kvxopt_b200 import klu
    A = load_matrix("ACTIVSg2000"); A.sort_indices()
    n, batch = A.shape[0], 4096
    Fs = klu.symbolic(A); Fn = klu.numeric(A, Fs)
    rng = np.random.default_rng(7)
    vals = np.empty((batch, A.nnz))
    for b0 in range(0, batch, 512):
        vals[b0:b0 + 512] = A.data[None, :] * (1 + 1e-3 * rng.uniform(-1, 1, size=(512, A.nnz)))
    assert not klu.refactor_batch(Fn, vals).any()
    B = rng.standard_normal((batch, 2, n))
    X = B.copy()
    klu.solve_batch(Fn, X)
    # residual of a sample of the batch (first, last, and across the 128 groups of 32 matrices) + SuperLU on two of them
    for b in [0, 31, 32, 1000, 2047, 2048, 3333, 4095]:
        Ab = sp.csc_matrix((vals[b], A.indices, A.indptr), shape=(n, n))
        R = Ab @ X[b].T - B[b].T
        assert np.abs(R).max() <= 1e-9 * max(1.0, np.abs(X[b]).max())
    for b in (17, 4095):
        Ab = sp.csc_matrix((vals[b], A.indices, A.indptr), shape=(n, n))
        xr = spla.splu(Ab).solve(B[b, 0])
        assert np.linalg.norm(X[b, 0] - xr) <= 1e-10 * np.linalg.norm(xr)
    # determinism: the same values give bit-identical factors (fixed schedule, no atomics)
    klu.refactor_batch(Fn, vals)
    X2 = B.copy(); klu.solve_batch(Fn, X2)
    assert np.array_equal(X, X2)
    # linearity of the batched solve
    Y = (2.0 * B[:, :1, :] - 3.0 * B[:, 1:2, :]).copy()
    klu.solve_batch(Fn, Y)
    ref = 2.0 * X[:, 0, :] - 3.0 * X[:, 1, :]
    assert np.abs(Y[:, 0, :] - ref).max() <= 1e-9 * np.abs(ref).max()
    # transpose solve: A' x = b on a sample
    XT = B[:, :1, :].copy(); klu.solve_batch(Fn, XT, trans="T")
    for b in (5, 4000):
        Ab = sp.csc_matrix((vals[b], A.indices, A.indptr), shape=(n, n))
        assert np.abs(Ab.T @ XT[b, 0] - B[b, 0]).max() <= 1e-9 * max(1.0, np.abs(XT[b]).max())
    # the streaming form leaves the same factors
    klu.refactor_batch_begin(Fn, vals); assert not klu.refactor_batch_end(Fn).any()
    X3 = B.copy(); klu.solve_batch(Fn, X3)
    assert np.array_equal(X, X3)


def test_cholesky_100cubed_properties():
    from kvxopt_b200 import cholmod, _lib as L
    nx = 100
    A = lap3d(nx, nx, nx); Al = lower_ccs(A); n = A.shape[0]
    perm = np.zeros(n, np.int64)
    assert L.fn["b200s_grid_nd_perm"](nx, nx, nx, 64, L.ptr_i64(perm)) == 0
    F = cholmod.symbolic(Al, p=perm)
    cholmod.numeric(Al, F)
    rng = np.random.default_rng(0)
    B = rng.standard_normal((n, 2))
    X = np.asfortranarray(B.copy()); cholmod.solve(F, X)
    res = np.linalg.norm(A @ X - B) / (abs(A).sum(axis=0).max() * np.linalg.norm(X) + np.linalg.norm(B))
    assert res <= 1e-12                                                     # north_star backward-error bound
    # linearity
    Y = np.asfortranarray(0.5 * B[:, :1] + 4.0 * B[:, 1:2]); cholmod.solve(F, Y)
    ref = 0.5 * X[:, 0] + 4.0 * X[:, 1]
    assert np.abs(Y[:, 0] - ref).max() <= 1e-9 * np.abs(ref).max()
    # sys 7 then 8 is the identity; sys 4 then 5 (with the permutations) is sys 0
    Z = np.asfortranarray(B[:, :1].copy()); cholmod.solve(F, Z, sys=7); cholmod.solve(F, Z, sys=8)
    assert np.array_equal(Z[:, 0], B[:, 0])
    W = np.asfortranarray(B[:, :1].copy())
    for s in (7, 4, 5, 8):
        cholmod.solve(F, W, sys=s)
    assert np.abs(W[:, 0] - X[:, 0]).max() <= 1e-10 * np.abs(X[:, 0]).max()
    # product of the diagonal of L against the known determinant of the Dirichlet Laplacian (sum of log eigenvalues)
    d = np.asarray(cholmod.diag(F)).ravel()
    k = np.arange(1, nx + 1)
    lam1 = 2.0 - 2.0 * np.cos(np.pi * k / (nx + 1))
    logdet = np.log(lam1[:, None, None] + lam1[None, :, None] + lam1[None, None, :]).sum()
    assert abs(2.0 * np.log(d).sum() - logdet) <= 1e-9 * abs(logdet)
    # determinism
    cholmod.numeric(Al, F)
    X2 = np.asfortranarray(B.copy()); cholmod.solve(F, X2)
    assert np.array_equal(X, X2)


_C5_CHILD = r"""
import json, os, sys
import numpy as np, scipy.sparse as sp
root = sys.argv[1]
sys.path.insert(0, root); sys.path.insert(0, os.path.join(root, "oracle", "_ref")); sys.path.insert(0, os.path.join(root, "tests", "golden"))
import kvxopt
from kvxopt import matrix, spmatrix, solvers
from kvxopt_b200 import kkt
from generators import qp_instance
solvers.options["show_progress"] = False
P, q, G, h = qp_instance(500, 400, 5000)
def tosp(M):
    M = sp.coo_matrix(M)
    return spmatrix(M.data.tolist(), M.row.tolist(), M.col.tolist(), M.shape)
Pk, Gk = tosp(sp.tril(P)), tosp(G)
sol = solvers.qp(Pk, matrix(q), Gk, matrix(h), kktsolver=kkt.qp_kktsolver(Pk, Gk))
x = np.array(sol["x"]).ravel()
print("C5RESULT " + json.dumps({"status": sol["status"], "iterations": sol["iterations"], "pobj": sol["primal objective"],
                                 "max_violation": float((G @ x - h).max())}))
"""


def test_config5_qp_200k_variables_through_coneqp():
    """BASELINE configs[4] at full size: the 200 000-variable sparse QP through the UNMODIFIED reference coneqp with the device
    KKT solver plugged in through kktsolver=; golden = the same IPM with the CPU oracle as its factorization
    (tests/golden/qp_config5_golden.json): same iteration count, objective to 1e-8 relative.  Runs in a child process with
    OPENBLAS_NUM_THREADS=1 set before the BLAS loads: the reference's own misc_solvers.scale crashes under multi-threaded
    OpenBLAS at m >= 4e5 rows."""
    import json
    import os
    import subprocess
    import sys
    from conftest import GOLD, ROOT
    gold = json.load(open(os.path.join(GOLD, "qp_config5_golden.json")))
    env = dict(os.environ, OPENBLAS_NUM_THREADS="1")
    r = subprocess.run([sys.executable, "-c", _C5_CHILD, ROOT], env=env, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    line = [l for l in r.stdout.splitlines() if l.startswith("C5RESULT ")][-1]
    sol = json.loads(line[len("C5RESULT "):])
    assert sol["status"] == gold["status"]
    assert sol["iterations"] == gold["iterations"]
    assert abs(sol["pobj"] - gold["primal_objective"]) <= 1e-8 * abs(gold["primal_objective"])
    assert sol["max_violation"] <= 1e-6          # G x <= h at the returned point
