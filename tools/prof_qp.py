import sys, os, time, ctypes as C, numpy as np, scipy.sparse as sp
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
from generators import qp_instance
from kvxopt_b200 import _lib as L, cholmod
nx, ny, nr = (int(a) for a in sys.argv[1:4])
P, q, G, h = qp_instance(nx, ny, nr)
S = sp.tril(P + G.T @ G).tocsc(); S.sort_indices()
F = cholmod.symbolic(S)
hh, _ = cholmod._factor_handle(F)
for prof in (1, 0, 0):
    L.fn["b200s_chol_set_profiling"](hh, prof)
    t = time.perf_counter(); cholmod.numeric(S, F); wall = time.perf_counter() - t
    d = cholmod.factor_info(F)
    print("prof=%d wall %.1f ms | total %.2f h2d %.2f asm %.2f factor %.2f | ext %.2f small %.2f panel %.2f upd %.2f | levels %d nsuper %d flops %.3g" % (
        prof, wall * 1e3, d["ms_total"], d["ms_h2d"], d["ms_assemble"], d["ms_factor"], d["ms_extend"], d["ms_potrf"], d["ms_trsm"], d["ms_dense_update"], d["nlevels"], d["nsuper"], d["flops"]), flush=True)
b = np.ones((S.shape[0], 1), order="F"); t = time.perf_counter(); cholmod.solve(F, b); print("solve wall %.1f ms (device %.2f)" % ((time.perf_counter() - t) * 1e3, cholmod.factor_info(F)["ms_solve"]))
