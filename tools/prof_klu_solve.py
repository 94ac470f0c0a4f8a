"""Device-resident batched KLU solve timing (ACTIVSg2000, batch 4096): ms per solve_batch_dev call, 'N' and 'T'."""
import ctypes as C, os, sys
import numpy as np, scipy.sparse as sp, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from kvxopt_b200 import _lib as L
fn = L.fn
batch = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
nrhs = int(sys.argv[2]) if len(sys.argv) > 2 else 1
z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden", "ACTIVSg2000.npz"))
n = int(z["n"]); A = sp.csc_matrix((z["values"], z["rowind"].astype(np.int64), z["colptr"]), shape=(n, n)); A.sort_indices()
cp = A.indptr.astype(np.int64); ri = A.indices.astype(np.int64); vx = A.data.astype(np.float64); nnz = len(vx)
S = L.vp(); assert fn["b200s_klu_analyze"](n, L.ptr_i64(cp), L.ptr_i64(ri), C.byref(S)) == 0
N = L.vp(); assert fn["b200s_klu_factor"](S, L.ptr_i64(cp), L.ptr_i64(ri), L.ptr_f64(vx), C.byref(N)) == 0
rng = np.random.default_rng(0)
vals = np.ascontiguousarray(vx[None, :] * (1 + 1e-3 * rng.uniform(-1, 1, size=(batch, nnz))))
assert fn["b200s_klu_refactor_batch"](N, L.ptr_f64(vals), batch, nnz, None) == 0
Bh = rng.standard_normal((batch, nrhs, n))
inf = L.KluInfo()
for tr in (0, 1):
    ts = []
    for rep in range(4):
        Bd = torch.from_numpy(Bh).cuda()
        torch.cuda.synchronize()
        assert fn["b200s_klu_solve_batch_dev"](N, tr, Bd.data_ptr(), nrhs, n, batch) == 0
        fn["b200s_klu_info"](N, C.byref(inf)); ts.append(inf.ms_solve)
    X = Bd.cpu().numpy()
    Ab = sp.csc_matrix((vals[7], A.indices, A.indptr), shape=(n, n)); M = Ab.T if tr else Ab
    print("solve_batch_dev trans=%d nrhs=%d batch=%d: %s ms, residual %.2e" % (tr, nrhs, batch, ["%.3f" % t for t in ts], np.abs(M @ X[7].T - Bh[7].T).max()), flush=True)
