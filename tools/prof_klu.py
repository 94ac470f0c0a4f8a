"""Short KLU refactor run for ncu: ACTIVSg2000, one batch size, a few launches."""
import ctypes as C, os, sys
import numpy as np, scipy.sparse as sp
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from kvxopt_b200 import _lib as L
fn = L.fn
batch = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden", "ACTIVSg2000.npz"))
n = int(z["n"]); A = sp.csc_matrix((z["values"], z["rowind"].astype(np.int64), z["colptr"]), shape=(n, n)); A.sort_indices()
cp = A.indptr.astype(np.int64); ri = A.indices.astype(np.int64); vx = A.data.astype(np.float64); nnz = len(vx)
S = L.vp(); assert fn["b200s_klu_analyze"](n, L.ptr_i64(cp), L.ptr_i64(ri), C.byref(S)) == 0
N = L.vp(); assert fn["b200s_klu_factor"](S, L.ptr_i64(cp), L.ptr_i64(ri), L.ptr_f64(vx), C.byref(N)) == 0
rng = np.random.default_rng(0)
vals = np.ascontiguousarray(vx[None, :] * (1 + 1e-3 * rng.uniform(-1, 1, size=(batch, nnz))))
inf = L.KluInfo()
for r in range(reps):
    assert fn["b200s_klu_refactor_batch"](N, L.ptr_f64(vals), batch, nnz, None) == 0
    fn["b200s_klu_info"](N, C.byref(inf))
    print("batch %d: refactor %.3f ms, kernel %.3f ms, h2d %.3f ms" % (batch, inf.ms_refactor, inf.ms_kernel, inf.ms_h2d), flush=True)
if len(sys.argv) > 3:      # argv[3] = nrhs: also time the batched solve ('N' then 'T')
    nrhs = int(sys.argv[3])
    Bb = np.ascontiguousarray(rng.standard_normal((batch, nrhs, n)))
    for tr in (0, 1):
        X = Bb.copy()
        assert fn["b200s_klu_solve_batch"](N, tr, L.ptr_f64(X), nrhs, n, batch) == 0
        fn["b200s_klu_info"](N, C.byref(inf))
        Ab = sp.csc_matrix((vals[0], A.indices, A.indptr), shape=(n, n))
        M = Ab.T if tr else Ab
        res = np.abs(M @ X[0].T - Bb[0].T).max()
        print("solve_batch trans=%d nrhs=%d: %.3f ms (host buffers), residual %.2e" % (tr, nrhs, inf.ms_solve, res), flush=True)
