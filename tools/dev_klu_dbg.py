import ctypes as C, numpy as np, scipy.sparse as sp, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from kvxopt_b200 import _lib as L
fn = L.fn
rng = np.random.default_rng(1)
for n, dens in ((1, 1), (6, 0.4), (50, 0.08), (300, 0.01)):
    A = (sp.random(n, n, density=dens, random_state=rng, format="csc") + sp.identity(n) * 0.5).tocsc(); A.sort_indices()
    if n < 50: continue
    cp = A.indptr.astype(np.int64); ri = A.indices.astype(np.int64); vx = A.data.astype(np.float64)
    S = L.vp(); assert fn["b200s_klu_analyze"](n, L.ptr_i64(cp), L.ptr_i64(ri), C.byref(S)) == 0
    N = L.vp(); st = fn["b200s_klu_factor"](S, L.ptr_i64(cp), L.ptr_i64(ri), L.ptr_f64(vx), C.byref(N)); assert st == 0
    inf = L.KluInfo(); fn["b200s_klu_info"](N, C.byref(inf)); d = inf.asdict()
    hL = np.zeros(d["nnz_L"]); hU = np.zeros(d["nnz_U"]); hF = np.zeros(max(d["nnz_F"], 1)); hR = np.zeros(n)
    fn["b200s_klu_extract_host"](N, L.ptr_f64(hL), L.ptr_f64(hU), L.ptr_f64(hF), L.ptr_f64(hR))
    dL = np.zeros(d["nnz_L"]); dU = np.zeros(d["nnz_U"]); dF = np.zeros(max(d["nnz_F"], 1)); dR = np.zeros(n)
    st = fn["b200s_klu_extract_batch"](N, 0, L.ptr_f64(dL), L.ptr_f64(dU), L.ptr_f64(dF), L.ptr_f64(dR)); assert st == 0, st
    print(n, "Rs", np.abs(hR - dR).max(), "U", np.abs(hU - dU).max(), "L", np.abs(hL - dL).max(), "F", np.abs(hF - dF).max())
    bad = np.nonzero(np.abs(hU - dU) > 1e-9)[0]; print(" bad U idx", bad[:10], hU[bad[:5]], dU[bad[:5]])
    bad = np.nonzero(np.abs(hL - dL) > 1e-9)[0]; print(" bad L idx", bad[:10], hL[bad[:5]], dL[bad[:5]])
    bad = np.nonzero(np.abs(hR - dR) > 1e-9)[0]; print(" bad R idx", bad[:10], hR[bad[:5]], dR[bad[:5]])
    # solve with the extracted factors in numpy and compare with the device solve
    Lp = np.zeros(n + 1, np.int64); Up = np.zeros(n + 1, np.int64); Fp = np.zeros(n + 1, np.int64)
    Li = np.zeros(d["nnz_L"], np.int64); Ui = np.zeros(d["nnz_U"], np.int64); Fi = np.zeros(max(d["nnz_F"], 1), np.int64)
    P = np.zeros(n, np.int64); Q = np.zeros(n, np.int64); R = np.zeros(d["nblocks"] + 1, np.int64)
    fn["b200s_klu_extract"](N, L.ptr_i64(Lp), L.ptr_i64(Li), None, L.ptr_i64(Up), L.ptr_i64(Ui), None, L.ptr_i64(Fp), L.ptr_i64(Fi), None, L.ptr_i64(P), L.ptr_i64(Q), None, L.ptr_i64(R))
    Lm = sp.csc_matrix((dL, Li, Lp), shape=(n, n)).toarray(); Um = sp.csc_matrix((dU, Ui, Up), shape=(n, n)).toarray()
    Fm = sp.csc_matrix((dF[:d["nnz_F"]], Fi[:d["nnz_F"]], Fp), shape=(n, n)).toarray()
    b = rng.standard_normal(n)
    y = np.linalg.solve(Lm, b[P] / dR); z = np.linalg.solve(Um + Fm, y); xref = np.zeros(n); xref[Q] = z
    print("   numpy-from-factors residual", np.abs(A @ xref - b).max(), "blocks", d["nblocks"])
    x = b.copy(); st = fn["b200s_klu_solve"](N, 0, L.ptr_f64(x), 1, n); assert st == 0
    print("   device solve vs numpy-from-factors", np.abs(x - xref).max(), "device residual", np.abs(A @ x - b).max())
    ident = np.abs(np.diag(1 / dR) @ A.toarray()[np.ix_(P, Q)] - (Lm @ Um + Fm)).max(); print("   identity", ident)
