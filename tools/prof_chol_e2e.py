"""Host-clock split of the host-buffer Cholesky path on the nx^3 Laplacian (what bench.py's cholesky.e2e measures):
cholmod.numeric(A, F) and cholmod.solve(F, X) call by call, first calls and repeats, next to the device times."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import lap3d_lower
from kvxopt_b200 import _lib as L, cholmod
nx = int(sys.argv[1]) if len(sys.argv) > 1 else 100
Al = lap3d_lower(nx); n = Al.shape[0]
perm = np.zeros(n, np.int64); L.fn["b200s_grid_nd_perm"](nx, nx, nx, 64, L.ptr_i64(perm))
F = cholmod.symbolic(Al, p=perm)
B = np.random.default_rng(0).standard_normal((n, 1))
for rep in range(4):
    X = np.asfortranarray(B.copy())
    t0 = time.perf_counter(); cholmod.numeric(Al, F); t1 = time.perf_counter(); cholmod.solve(F, X); t2 = time.perf_counter()
    d = cholmod.factor_info(F)
    print("rep %d: numeric %.1f ms (device total %.1f, factor %.1f, h2d %.1f) solve %.1f ms (device %.2f)" % (
        rep, (t1 - t0) * 1e3, d["ms_total"], d["ms_factor"], d["ms_h2d"], (t2 - t1) * 1e3, d["ms_solve"]), flush=True)
