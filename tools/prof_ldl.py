"""Signed (LDL', cholmod.options['supernodal'] = 0) against plain (LL') factorization of the same matrices:
3-D Laplacian nx^3 (all pivots positive: isolates the cost of the sign handling) and a quasi-definite system
[[Lap + I, B'], [B, -(Lap2 + I)]] (mixed signs)."""
import os, sys
import numpy as np, scipy.sparse as sp
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import lap3d_lower
from kvxopt_b200 import _lib as L, cholmod
nx = int(sys.argv[1]) if len(sys.argv) > 1 else 64
Al = lap3d_lower(nx); n = Al.shape[0]
perm = np.zeros(n, np.int64); L.fn["b200s_grid_nd_perm"](nx, nx, nx, 64, L.ptr_i64(perm))
b = np.random.default_rng(0).standard_normal((n, 1))
A = Al + sp.tril(Al, -1).T
for mode in (2, 0):
    cholmod.options["supernodal"] = mode
    F = cholmod.symbolic(Al, p=perm)
    for rep in range(3):
        cholmod.numeric(Al, F)
        d = cholmod.factor_info(F)
        x = np.asfortranarray(b.copy()); cholmod.solve(F, x)
        r = np.abs(A @ x - b).max() / (np.abs(A).sum(axis=1).max() * np.abs(x).max() + np.abs(b).max())
        print("lap %d^3 supernodal=%d: factor %.2f ms (%.2f TFLOP/s) solve %.2f ms backward error %.1e" % (
            nx, mode, d["ms_factor"], d["flops"] / d["ms_factor"] / 1e9, cholmod.factor_info(F)["ms_solve"], r), flush=True)
    del F
del cholmod.options["supernodal"]
