"""TEST INFRASTRUCTURE: CPU restatement of CHOLMOD's simplicial LDL' factorization without pivoting -- what
`cholmod.options['supernodal'] = 0` selects in the reference (src/C/cholmod.c:60-64; numeric at :362, `sys` 1..6 at
:437-439) -- and of the dense 3 x 3 KKT solver `misc.kkt_ldl` (src/python/misc.py:1055-1130) for componentwise cones.

SuiteSparse 7.8.2 is a third-party dependency absent from /root/reference; the algorithm restated here is
the published up-looking row LDL' (Davis, "Algorithm 849: a concise sparse Cholesky factorization package", and the
simplicial path of Algorithm 887 CHOLMOD): row k of L solves L(0:k,0:k) D(0:k) l_k' = A(0:k,k), then
d_k = a_kk - sum_j l_kj^2 d_j; a zero pivot stops the factorization (CHOLMOD reports it as "not positive definite" with
`minor` = k), a negative one does not.  Pinned in tests/test_oracle.py against numpy (L D L' = P A P' entrywise, inertia =
number of negative eigenvalues, solutions of numpy.linalg.solve) and in tests/test_gpu_ldl.py against the reference's own
`lapack.sytrf/sytrs` KKT solutions (misc.kkt_ldl run from oracle/_ref, fixtures in tests/golden/).

Only tests/ may import this module; the product never does.
"""
import numpy as np
import scipy.linalg as sla


def ldl_nopivot(Ap):
    """Dense up-looking LDL' of the symmetric matrix Ap (already permuted).  Returns (L unit lower, d, minor) with
    minor = n on success or the index of the first zero pivot."""
    Ap = np.asarray(Ap, dtype=np.float64)
    n = Ap.shape[0]
    L = np.eye(n)
    d = np.zeros(n)
    for k in range(n):
        if k:
            y = sla.solve_triangular(L[:k, :k], Ap[:k, k], lower=True, unit_diagonal=True, check_finite=False)
            lk = y / d[:k]
            L[k, :k] = lk
            d[k] = Ap[k, k] - lk @ y
        else:
            d[0] = Ap[0, 0]
        if d[k] == 0.0 or not np.isfinite(d[k]):
            return L, d, k
    return L, d, n


def solve_sys(L, d, perm, B, sys):
    """The nine systems of cholmod.solve for an LDL' factor (reference numbering, src/C/cholmod.c:437-439)."""
    B = np.asarray(B, dtype=np.float64)
    fwd = lambda R: sla.solve_triangular(L, R, lower=True, unit_diagonal=True, check_finite=False)
    bwd = lambda R: sla.solve_triangular(L.T, R, lower=False, unit_diagonal=True, check_finite=False)
    D = d[:, None] if B.ndim == 2 else d
    if sys == 0:
        X = np.empty_like(B)
        X[perm] = bwd(fwd(B[perm]) / D)
        return X
    if sys == 1:
        return bwd(fwd(B) / D)
    if sys == 2:
        return fwd(B) / D
    if sys == 3:
        return bwd(B / D)
    if sys == 4:
        return fwd(B)
    if sys == 5:
        return bwd(B)
    if sys == 6:
        return B / D
    if sys == 7:
        return B[perm]
    if sys == 8:
        X = np.empty_like(B)
        X[perm] = B
        return X
    raise ValueError("sys")


def kkt_ldl_lcone(G, A, H, di, bx, by, bz):
    """misc.kkt_ldl (src/python/misc.py:1055-1130) for dims = {'l': m}: solves
        [ H  A'  G' ] [ux]   [bx]
        [ A  0   0  ] [uy] = [by]        and returns ux, uy, W uz  (W = diag(d), di = 1/d)
        [ G  0 -W'W ] [uz]   [bz]
    through the scaled system with -I in the (3,3) block, densely (numpy.linalg.solve stands in for sytrf/sytrs)."""
    G = np.asarray(G, dtype=np.float64); A = np.asarray(A, dtype=np.float64).reshape(-1, G.shape[1])
    m, n = G.shape
    p = A.shape[0]
    K = np.zeros((n + p + m, n + p + m))
    if H is not None:
        K[:n, :n] = H
    K[n:n + p, :n] = A; K[:n, n:n + p] = A.T
    Gs = di[:, None] * G
    K[n + p:, :n] = Gs; K[:n, n + p:] = Gs.T
    K[n + p:, n + p:] = -np.eye(m)
    u = np.linalg.solve(K, np.concatenate([bx, by, di * bz]))
    return u[:n], u[n:n + p], u[n + p:]
