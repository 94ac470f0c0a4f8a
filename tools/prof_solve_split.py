"""Forward / backward split of the supernodal solve on the 3-D Laplacian nx^3 (sys 4 = L x = b, sys 5 = L'x = b, sys 0 = all)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import lap3d_lower
from kvxopt_b200 import _lib as L, cholmod
nx = int(sys.argv[1]) if len(sys.argv) > 1 else 64
Al = lap3d_lower(nx); n = Al.shape[0]
perm = np.zeros(n, np.int64); L.fn["b200s_grid_nd_perm"](nx, nx, nx, 64, L.ptr_i64(perm))
F = cholmod.symbolic(Al, p=perm)
cholmod.numeric(Al, F)
d = cholmod.factor_info(F)
rng = np.random.default_rng(0)
for nrhs in (1, 4):
    for sys_ in (4, 5, 0):
        for rep in range(3):
            x = np.asfortranarray(rng.standard_normal((n, nrhs))); cholmod.solve(F, x, sys=sys_)
        ms = cholmod.factor_info(F)["ms_solve"]
        passes = 2 if sys_ == 0 else 1
        print("nrhs=%d sys=%d: %.3f ms  => %.0f GB/s of L" % (nrhs, sys_, ms, passes * 8 * d["nnz_L"] / ms / 1e6), flush=True)
