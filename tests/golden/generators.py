"""Synthetic workload generators shared by the fixture scripts, the tests and bench.py (no reference needed)."""
import numpy as np
import scipy.sparse as sp


def qp_instance(nx, ny, nrand, seed=1):
    """BASELINE config 5 generator at reduced size: P = 5-point Laplacian + 1e-2 I,
    G = [I; -I; R] with R random 3-nnz rows, h = [1; 1; |R| 1 + 1], q ~ N(0,1)."""
    n = nx * ny
    T = lambda m: sp.diags([-np.ones(m - 1), 2 * np.ones(m), -np.ones(m - 1)], [-1, 0, 1])
    P = (sp.kron(sp.identity(ny), T(nx)) + sp.kron(T(ny), sp.identity(nx)) + 1e-2 * sp.identity(n)).tocsc()
    rng = np.random.default_rng(seed)
    cols = rng.integers(0, n, size=(nrand, 3))
    vals = rng.standard_normal((nrand, 3))
    R = sp.csc_matrix((vals.reshape(-1), (np.repeat(np.arange(nrand), 3), cols.reshape(-1))), shape=(nrand, n))
    G = sp.vstack([sp.identity(n), -sp.identity(n), R]).tocsc()
    h = np.concatenate([np.ones(2 * n), np.abs(R).sum(axis=1).A1 + 1.0])
    q = np.random.default_rng(2).standard_normal(n)
    return P, q, G, h
