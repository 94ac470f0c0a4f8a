"""Solve timing of the 3-D Laplacian nx^3 with the persistent sweeps switched on/off (cholmod.set_solve_sweeps):
mode 0 = launch per block step, 1 = persistent forward, 2 = persistent backward, 3 = both; bitwise comparison with mode 0."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import lap3d_lower
from kvxopt_b200 import _lib as L, cholmod
nx = int(sys.argv[1]) if len(sys.argv) > 1 else 64
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
Al = lap3d_lower(nx); n = Al.shape[0]
perm = np.zeros(n, np.int64); L.fn["b200s_grid_nd_perm"](nx, nx, nx, 64, L.ptr_i64(perm))
F = cholmod.symbolic(Al, p=perm)
cholmod.numeric(Al, F)
d = cholmod.factor_info(F)
print("nx %d factor %.2f ms nnz(L) %.3e" % (nx, d["ms_factor"], d["nnz_L"]), flush=True)
b = np.random.default_rng(0).standard_normal((n, 1))
ref = None
for mode in (0, 1, 2, 3):
    cholmod.set_solve_sweeps(F, mode)
    best = {}
    for sys_ in (0, 4, 5):
        ts = []
        for r in range(reps + 1):
            x = np.asfortranarray(b.copy()); cholmod.solve(F, x, sys=sys_)
            ts.append(cholmod.factor_info(F)["ms_solve"])
        best[sys_] = min(ts[1:])
        if sys_ == 0:
            if ref is None: ref = x
            same = np.array_equal(x, ref)
    print("mode %d: solve %.3f ms (forward %.3f, backward %.3f)  bitwise==mode0 %s  maxdiff %.2e" % (
        mode, best[0], best[4], best[5], same, np.abs(x - ref).max() if sys_ == 5 else 0), flush=True)
