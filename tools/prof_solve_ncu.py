"""One factorization + one solve (sys given) for an ncu launch list of the solve kernels (run with B200S_NO_GRAPH=1)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import lap3d_lower
from kvxopt_b200 import _lib as L, cholmod
nx = int(sys.argv[1]); nrhs = int(sys.argv[2]); sys_ = int(sys.argv[3])
Al = lap3d_lower(nx); n = Al.shape[0]
perm = np.zeros(n, np.int64); L.fn["b200s_grid_nd_perm"](nx, nx, nx, 64, L.ptr_i64(perm))
F = cholmod.symbolic(Al, p=perm)
cholmod.numeric(Al, F)
x = np.asfortranarray(np.random.default_rng(0).standard_normal((n, nrhs)))
cholmod.solve(F, x, sys=sys_)
print("ms_solve", cholmod.factor_info(F)["ms_solve"])
