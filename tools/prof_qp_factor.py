"""Per-kernel-class device time of ONE factorization of the reduced KKT matrix S = P + G' diag(d)^2 G of BASELINE config 5
(200k-variable QP; AMD ordering as the KKT solver uses), through cholmod.numeric with the profiling events on."""
import os, sys
import numpy as np, scipy.sparse as sp
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
from generators import qp_instance
from kvxopt_b200 import _lib as L, cholmod
P, q, G, h = qp_instance(500, 400, 5000)
d = np.random.default_rng(0).uniform(0.5, 2.0, G.shape[0])
S = (sp.csc_matrix(P) + (G.T @ sp.diags(d * d) @ G)).tocsc()
Sl = sp.tril(S).tocsc(); Sl.sort_indices()
F = cholmod.symbolic(Sl)
hF, _ = cholmod._factor_handle(F)
cholmod.numeric(Sl, F)
L.fn["b200s_chol_set_profiling"](hF, 1)
for _ in range(2):
    cholmod.numeric(Sl, F)
    i = cholmod.factor_info(F)
    print("n %d nnz(L) %.3e flops %.3e | factor %.2f ms %.2f TF/s | ext %.2f small %.2f panel %.2f upd %.2f (%.2f TF/s in k_update) | levels %d max front %d x %d" % (
        i["n"], i["nnz_L"], i["flops"], i["ms_factor"], i["flops"] / i["ms_factor"] / 1e9, i["ms_extend"], i["ms_potrf"], i["ms_trsm"],
        i["ms_dense_update"], i["flops_update"] / max(i["ms_dense_update"], 1e-9) / 1e9, i["nlevels"], i["max_front_rows"], i["max_front_cols"]), flush=True)
