"""Short Cholesky run for ncu: 3-D Laplacian nx^3 with nested dissection, one warm-up + one factorization + solve."""
import ctypes as C, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import lap3d_lower
from kvxopt_b200 import _lib as L, cholmod
nx = int(sys.argv[1]) if len(sys.argv) > 1 else 64
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
Al = lap3d_lower(nx); n = Al.shape[0]
perm = np.zeros(n, np.int64); L.fn["b200s_grid_nd_perm"](nx, nx, nx, 64, L.ptr_i64(perm))
if os.environ.get("CHOL_LDL"):
    cholmod.options["supernodal"] = 0      # signed (LDL') instantiation of the kernels
F = cholmod.symbolic(Al, p=perm)
h, _ = cholmod._factor_handle(F)
prof = 0 if (len(sys.argv) > 3 and sys.argv[3] == "noprof") else 1
L.fn["b200s_chol_set_profiling"](h, prof)
for r in range(reps):
    cholmod.numeric(Al, F)
    d = cholmod.factor_info(F)
    print("factor %.2f ms  %.2f TF/s | ext %.2f small %.2f panel %.2f upd %.2f (%.2f TF/s in k_update)" % (
        d["ms_factor"], d["flops"] / d["ms_factor"] / 1e9, d["ms_extend"], d["ms_potrf"], d["ms_trsm"], d["ms_dense_update"],
        d["flops_update"] / max(d["ms_dense_update"], 1e-9) / 1e9), flush=True)
x = np.ones((n, 1), order="F"); cholmod.solve(F, x)
print("solve %.2f ms" % cholmod.factor_info(F)["ms_solve"])
