"""For an ncu launch list of ONE factorization: analyse + warm-up factorization of the nx^3 Laplacian, then cudaProfilerStart /
one factorization / cudaProfilerStop (run under `ncu --profile-from-start off`)."""
import os, sys, ctypes
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import lap3d_lower
from kvxopt_b200 import _lib as L, cholmod
nx = int(sys.argv[1]) if len(sys.argv) > 1 else 100
rt = ctypes.CDLL("libcudart.so.12")
Al = lap3d_lower(nx); n = Al.shape[0]
perm = np.zeros(n, np.int64); L.fn["b200s_grid_nd_perm"](nx, nx, nx, 64, L.ptr_i64(perm))
F = cholmod.symbolic(Al, p=perm)
cholmod.numeric(Al, F)
rt.cudaProfilerStart()
cholmod.numeric(Al, F)
rt.cudaProfilerStop()
print("factor %.2f ms" % cholmod.factor_info(F)["ms_factor"], flush=True)
