"""Per-kernel-class device time of ONE factorization of the 3 x 3 KKT matrix of BASELINE config 5 through kkt.ldl (signed LDL',
order 605 000, z -> x elimination order), next to the reduced system of kkt.chol2 (tools/prof_qp_factor.py)."""
import os, sys, ctypes as C
import numpy as np, scipy.sparse as sp
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests", "golden")); sys.path.insert(0, os.path.join(ROOT, "oracle", "_ref"))
from generators import qp_instance
from kvxopt import matrix, spmatrix
from kvxopt_b200 import _lib as L, kkt
nx, ny, nrand = (int(a) for a in sys.argv[1:4]) if len(sys.argv) > 3 else (500, 400, 5000)
P, q, G, h = qp_instance(nx, ny, nrand)
def tosp(M):
    M = sp.coo_matrix(M); return spmatrix(M.data.tolist(), M.row.tolist(), M.col.tolist(), M.shape)
Pk, Gk = tosp(sp.tril(P)), tosp(G)
m = Gk.size[0]
f3 = kkt.ldl(Gk, {"l": m, "q": [], "s": []}, spmatrix([], [], [], (0, Pk.size[0])))
W = {"di": matrix(np.random.default_rng(0).uniform(0.5, 2.0, m))}
f3(W, Pk)
hF = f3._state["handle"].h
for prof in (0, 1):
    L.fn["b200s_chol_set_profiling"](hF, prof)
    for _ in range(2):
        f3(W, Pk)
        i = L.CholInfo(); L.fn["b200s_chol_info"](hF, C.byref(i))
        print("profiling %d: n %d nsuper %d nnz(L) %.3e flops %.3e | total %.2f ms (h2d %.2f assemble %.2f factor %.2f) | ext %.2f small %.2f panel %.2f upd %.2f | levels %d" % (
            prof, i.n, i.nsuper, i.nnz_L, i.flops, i.ms_total, i.ms_h2d, i.ms_assemble, i.ms_factor, i.ms_extend, i.ms_potrf, i.ms_trsm, i.ms_dense_update, i.nlevels), flush=True)
L.fn["b200s_chol_set_profiling"](hF, 0)
solve = f3(W, Pk)
n = Pk.size[0]
rng = np.random.default_rng(1)
for _ in range(3):
    x, y, z = matrix(rng.standard_normal(n)), matrix(0.0, (0, 1)), matrix(rng.standard_normal(m))
    solve(x, y, z)
    i = L.CholInfo(); L.fn["b200s_chol_info"](hF, C.byref(i))
    print("solve: device %.2f ms" % i.ms_solve, flush=True)
if os.environ.get("PROF_SOLVE_NCU"):      # ncu --profile-from-start off: the launch list of ONE solve (B200S_NO_GRAPH=1)
    import ctypes
    rt = ctypes.CDLL("libcudart.so.12")
    x, y, z = matrix(rng.standard_normal(n)), matrix(0.0, (0, 1)), matrix(rng.standard_normal(m))
    rt.cudaProfilerStart(); solve(x, y, z); rt.cudaProfilerStop()
