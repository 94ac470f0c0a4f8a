"""Cholesky solve timing on the 3-D Laplacian nx^3 (nested dissection): one factorization, then repeated solves."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import lap3d_lower
from kvxopt_b200 import _lib as L, cholmod
nx = int(sys.argv[1]) if len(sys.argv) > 1 else 64
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
nrhs = int(sys.argv[3]) if len(sys.argv) > 3 else 1
Al = lap3d_lower(nx); n = Al.shape[0]
perm = np.zeros(n, np.int64); L.fn["b200s_grid_nd_perm"](nx, nx, nx, 64, L.ptr_i64(perm))
F = cholmod.symbolic(Al, p=perm)
cholmod.numeric(Al, F)
d = cholmod.factor_info(F)
print("n=%d nnz(L)=%.3g factor %.2f ms" % (n, d["nnz_L"], d["ms_factor"]), flush=True)
rng = np.random.default_rng(0)
for r in range(reps):
    b = rng.standard_normal((n, nrhs)); x = np.asfortranarray(b.copy())
    t0 = time.perf_counter(); cholmod.solve(F, x); t1 = time.perf_counter()
    A = Al + Al.T - __import__("scipy.sparse", fromlist=["diags"]).diags(Al.diagonal())
    res = np.abs(A @ x - b).max() / (np.abs(A).sum(axis=1).max() * np.abs(x).max() + np.abs(b).max())
    print("solve nrhs=%d: device %.3f ms, wall %.3f ms, backward error %.2e, L bytes/solve-time = %.0f GB/s" % (
        nrhs, cholmod.factor_info(F)["ms_solve"], (t1 - t0) * 1e3, res, 2 * 8 * d["nnz_L"] / cholmod.factor_info(F)["ms_solve"] / 1e6), flush=True)
