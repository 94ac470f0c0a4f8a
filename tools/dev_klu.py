"""Development check of the CUDA KLU path (runs on the GPU box via gpurun)."""
import ctypes as C, os, sys, time
import numpy as np, scipy.sparse as sp, scipy.sparse.linalg as spla
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from kvxopt_b200 import _lib as L
fn = L.fn
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")

def load(name):
    z = np.load(os.path.join(GOLD, name + ".npz")); n = int(z["n"])
    return sp.csc_matrix((z["values"], z["rowind"], z["colptr"]), shape=(n, n))

def check(name, A, batch=64, nrhs=2, big=0):
    A = A.tocsc(); A.sort_indices(); n = A.shape[0]
    cp = A.indptr.astype(np.int64); ri = A.indices.astype(np.int64); vx = A.data.astype(np.float64); nnz = len(vx)
    S = L.vp(); st = fn["b200s_klu_analyze"](n, L.ptr_i64(cp), L.ptr_i64(ri), C.byref(S)); assert st == 0
    N = L.vp(); st = fn["b200s_klu_factor"](S, L.ptr_i64(cp), L.ptr_i64(ri), L.ptr_f64(vx), C.byref(N))
    assert st == 0, (st, L.last_error())
    rng = np.random.default_rng(0)
    ok = True
    msg = "%-14s n=%d nnz=%d" % (name, n, nnz)
    for trans in (0, 1):
        B = rng.standard_normal((n, nrhs)); X = np.asfortranarray(B.copy())
        st = fn["b200s_klu_solve"](N, trans, L.ptr_f64(X), nrhs, n); assert st == 0, (st, L.last_error())
        M = A.T if trans else A
        res = np.abs(M @ X - B).max() / max(1.0, np.abs(B).max())
        Xref = spla.splu(M.tocsc()).solve(B)
        rel = np.linalg.norm(X - Xref) / np.linalg.norm(Xref)
        msg += " t%d res=%.1e rel=%.1e" % (trans, res, rel)
        if not (rel < 1e-8): ok = False
    # batch
    vals = vx[None, :] * (1 + 1e-3 * rng.uniform(-1, 1, size=(batch, nnz)))
    vals = np.ascontiguousarray(vals)
    status = np.zeros(batch, dtype=np.int32)
    st = fn["b200s_klu_refactor_batch"](N, L.ptr_f64(vals), batch, nnz, status.ctypes.data_as(L.p_int)); assert st == 0, (st, L.last_error())
    inf = L.KluInfo(); fn["b200s_klu_info"](N, C.byref(inf)); d = inf.asdict()
    Bb = rng.standard_normal((batch, nrhs, n)); Xb = Bb.copy()
    st = fn["b200s_klu_solve_batch"](N, 0, L.ptr_f64(Xb), nrhs, n, batch); assert st == 0, (st, L.last_error())
    worst = 0
    for b in range(min(batch, 8)):
        Ab = sp.csc_matrix((vals[b], A.indices, A.indptr), shape=(n, n))
        Xref = spla.splu(Ab).solve(Bb[b].T)
        worst = max(worst, np.linalg.norm(Xb[b].T - Xref) / np.linalg.norm(Xref))
    msg += " batch%d: status=%d refactor=%.3fms h2d=%.3fms solve=%.3fms worst_rel=%.1e" % (batch, int(status.max()), d["ms_refactor"], d["ms_h2d"], inf.ms_solve, worst)
    if not (worst < 1e-8) or status.max() != 0: ok = False
    if big:
        vals = vx[None, :] * (1 + 1e-3 * rng.uniform(-1, 1, size=(big, nnz))); vals = np.ascontiguousarray(vals)
        for rep in range(3):
            st = fn["b200s_klu_refactor_batch"](N, L.ptr_f64(vals), big, nnz, None); assert st == 0
            fn["b200s_klu_info"](N, C.byref(inf))
        msg += " | batch%d refactor=%.3fms (%.0f refactors/s device, h2d %.2fms) levels=%d flops=%.3g bytes=%d" % (
            big, inf.ms_refactor, big / inf.ms_refactor * 1e3, inf.ms_h2d, inf.nlevels, inf.flops, inf.bytes_per_refactor)
        Bb = rng.standard_normal((big, 1, n)); Xb = Bb.copy()
        st = fn["b200s_klu_solve_batch"](N, 0, L.ptr_f64(Xb), 1, n, big); assert st == 0
        fn["b200s_klu_info"](N, C.byref(inf)); msg += " solve=%.2fms" % inf.ms_solve
        b = big - 1
        Ab = sp.csc_matrix((vals[b], A.indices, A.indptr), shape=(n, n)); Xref = spla.splu(Ab).solve(Bb[b, 0])
        msg += " last_rel=%.1e" % (np.linalg.norm(Xb[b, 0] - Xref) / np.linalg.norm(Xref))
    print(("OK   " if ok else "FAIL ") + msg, flush=True)
    fn["b200s_klu_free_numeric"](N); fn["b200s_klu_free_symbolic"](S)
    return ok

if __name__ == "__main__":
    rng = np.random.default_rng(1)
    allok = True
    for n, dens in ((1, 1), (6, 0.4), (50, 0.08), (300, 0.01), (1000, 0.003)):
        A = sp.random(n, n, density=dens, random_state=rng, format="csc") + sp.identity(n) * 0.5
        allok &= check("rand%d" % n, A, batch=33)
    allok &= check("bp_800", load("bp_800"))
    allok &= check("bcsstk13(tri)", load("bcsstk13"), batch=8)
    allok &= check("ACTIVSg2000", load("ACTIVSg2000"), batch=64, big=4096)
    print("ALL OK" if allok else "SOME FAILED")
    sys.exit(0 if allok else 1)
