"""One factorization of the nx^3 Laplacian and two one-column solves with the persistent sweeps, for an ncu launch list
(ncu -k regex:persist --metrics gpu__time_duration.sum): duration of every k_fwd_persist / k_bwd_persist launch."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import lap3d_lower
from kvxopt_b200 import _lib as L, cholmod
nx = int(sys.argv[1]) if len(sys.argv) > 1 else 100
Al = lap3d_lower(nx); n = Al.shape[0]
perm = np.zeros(n, np.int64); L.fn["b200s_grid_nd_perm"](nx, nx, nx, 64, L.ptr_i64(perm))
F = cholmod.symbolic(Al, p=perm)
cholmod.numeric(Al, F)
b = np.ones((n, 1))
for r in range(2):
    x = np.asfortranarray(b.copy()); cholmod.solve(F, x)
    print("solve %.3f ms" % cholmod.factor_info(F)["ms_solve"], flush=True)
