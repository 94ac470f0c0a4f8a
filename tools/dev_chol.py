"""Development check of the CUDA Cholesky path against scipy (runs on the GPU box via gpurun)."""
import ctypes as C, os, sys, time
import numpy as np, scipy.sparse as sp, scipy.sparse.linalg as spla
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
lib = C.CDLL(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "kvxopt_b200", "libb200sparse.so"))
i64 = C.c_int64; pi = C.POINTER(i64); pd = C.POINTER(C.c_double); vp = C.c_void_p
from kvxopt_b200 import _lib_types as T  # noqa
lib.b200s_last_error.restype = C.c_char_p
def P(a, t): return a.ctypes.data_as(t)

def analyze(A, perm=None, opts=None):
    A = sp.tril(A).tocsc(); A.sort_indices()
    cp = A.indptr.astype(np.int64); ri = A.indices.astype(np.int64)
    F = vp()
    pp = perm.astype(np.int64) if perm is not None else None
    st = lib.b200s_chol_analyze(i64(A.shape[0]), P(cp, pi), P(ri, pi), C.c_char(b'L'), P(pp, pi) if pp is not None else None, opts, C.byref(F))
    assert st == 0, (st, lib.b200s_last_error())
    return F, A

def info(F):
    inf = T.CholInfo(); lib.b200s_chol_info(F, C.byref(inf)); return inf.asdict()

def check(name, A, perm=None, nrhs=2, getL=False, reps=1, prof=False):
    n = A.shape[0]
    t0 = time.time(); F, Al = analyze(A, perm); ta = time.time() - t0
    val = np.ascontiguousarray(Al.data, dtype=np.float64)
    minor = i64(-1)
    if prof: lib.b200s_chol_set_profiling(F, 1)
    for _ in range(reps):
        st = lib.b200s_chol_factorize(F, P(val, pd), C.byref(minor))
    inf = info(F)
    if st != 0:
        print(name, "FACTOR FAILED", st, minor.value, lib.b200s_last_error()); return False
    rng = np.random.default_rng(0)
    B = rng.standard_normal((n, nrhs)); X = np.asfortranarray(B.copy())
    st = lib.b200s_chol_solve(F, 0, P(X, pd), i64(nrhs), i64(n))
    assert st == 0, (st, lib.b200s_last_error())
    inf2 = info(F)
    Afull = (Al + sp.tril(Al, -1).T).tocsc()
    R = Afull @ X - B
    berr = np.max(np.linalg.norm(R, axis=0) / (spla.norm(Afull, 1) * np.linalg.norm(X, axis=0) + np.linalg.norm(B, axis=0)))
    msg = "%-22s n=%-8d nnzL=%-10d nsup=%-6d lev=%-3d maxfront=%d/%d flops=%.3g analyze=%.0fms factor=%.2fms (h2d %.2f asm %.2f num %.2f) solve=%.2fms berr=%.2e" % (
        name, n, inf["nnz_L"], inf["nsuper"], inf["nlevels"], inf["max_front_rows"], inf["max_front_cols"], inf["flops"], ta * 1e3,
        inf["ms_total"], inf["ms_h2d"], inf["ms_assemble"], inf["ms_factor"], inf2["ms_solve"], berr)
    if inf["ms_factor"] > 0: msg += " %.2f TF/s" % (inf["flops"] / inf["ms_factor"] / 1e9)
    if prof: msg += " | ext %.2f small %.2f panel %.2f upd %.2f" % (inf["ms_extend"], inf["ms_potrf"], inf["ms_trsm"], inf["ms_dense_update"])
    ok = berr < 1e-12
    if n <= 3000:
        Xref = np.linalg.solve(Afull.toarray(), B)
        rel = np.linalg.norm(X - Xref) / np.linalg.norm(Xref)
        msg += " relx=%.2e" % rel
    if getL and n <= 2000:
        Lp = pi(); Li = pi(); Lx = pd()
        st = lib.b200s_chol_get_L(F, C.byref(Lp), C.byref(Li), C.byref(Lx)); assert st == 0
        lp = np.ctypeslib.as_array(Lp, shape=(n + 1,)).copy(); nnz = lp[-1]
        li = np.ctypeslib.as_array(Li, shape=(nnz,)).copy(); lx = np.ctypeslib.as_array(Lx, shape=(nnz,)).copy()
        L = sp.csc_matrix((lx, li, lp), shape=(n, n)).toarray()
        perm_o = np.zeros(n, dtype=np.int64); lib.b200s_chol_get_perm(F, P(perm_o, pi))
        Ad = Afull.toarray()[np.ix_(perm_o, perm_o)]
        Lref = np.linalg.cholesky(Ad)
        msg += " |L-Lref|=%.2e" % (np.abs(L - Lref).max() / np.abs(Lref).max())
        # sys 4/5/7/8
        b = rng.standard_normal(n)
        for sys_, ref in ((4, np.linalg.solve(Lref, b)), (5, np.linalg.solve(Lref.T, b)), (7, b[perm_o]), (8, None)):
            x = b.copy(); st = lib.b200s_chol_solve(F, sys_, P(x, pd), i64(1), i64(n)); assert st == 0
            if sys_ == 8: ref = np.zeros(n); ref[perm_o] = b
            e = np.linalg.norm(x - ref) / max(np.linalg.norm(ref), 1e-300)
            if e > 1e-9: ok = False
            msg += " s%d=%.1e" % (sys_, e)
        d = np.zeros(n); lib.b200s_chol_diag(F, P(d, pd)); msg += " diag=%.1e" % np.abs(d - np.diag(Lref)).max()
    print(("OK   " if ok else "FAIL ") + msg, flush=True)
    lib.b200s_chol_free(F)
    return ok

def lap(nx, ny, nz):
    def T1(n): return sp.diags([-np.ones(n - 1), 2 * np.ones(n), -np.ones(n - 1)], [-1, 0, 1])
    Ix, Iy, Iz = sp.identity(nx), sp.identity(ny), sp.identity(nz)
    A = sp.kron(Iz, sp.kron(Iy, T1(nx))) + sp.kron(Iz, sp.kron(T1(ny), Ix))
    if nz > 1: A = A + sp.kron(T1(nz), sp.kron(Iy, Ix))
    return A.tocsc()

def nd(nx, ny, nz, leaf=64):
    p = np.zeros(nx * ny * nz, dtype=np.int64)
    st = lib.b200s_grid_nd_perm(i64(nx), i64(ny), i64(nz), i64(leaf), P(p, pi)); assert st == 0
    return p

def rand_spd(n, dens, seed):
    rng = np.random.default_rng(seed)
    M = sp.random(n, n, density=dens, random_state=rng, format="csc")
    A = M + M.T + sp.identity(n) * (n * dens * 2 + 1)
    return A.tocsc()

if __name__ == "__main__":
    big = "--big" in sys.argv
    allok = True
    for n, d in ((1, 1.0), (5, 0.5), (40, 0.2), (150, 0.05), (300, 0.9), (700, 0.01), (1500, 0.004)):
        allok &= check("rand n=%d" % n, rand_spd(n, d, n), getL=True)
    allok &= check("dense 400", sp.csc_matrix(np.cov(np.random.default_rng(1).standard_normal((400, 900))) + np.eye(400)), getL=True)
    allok &= check("lap2d 40x40 amd", lap(40, 40, 1), getL=True)
    allok &= check("lap2d 300x300 amd", lap(300, 300, 1))
    allok &= check("lap3d 20 nd", lap(20, 20, 20), nd(20, 20, 20))
    allok &= check("lap3d 40 nd", lap(40, 40, 40), nd(40, 40, 40), reps=2, prof=True)
    allok &= check("lap3d 40 amd", lap(40, 40, 40), reps=2, prof=True)
    if big:
        allok &= check("lap3d 64 nd", lap(64, 64, 64), nd(64, 64, 64), reps=2, prof=True)
        allok &= check("lap3d 100 nd", lap(100, 100, 100), nd(100, 100, 100), reps=2, nrhs=1, prof=True)
    print("ALL OK" if allok else "SOME FAILED")
    sys.exit(0 if allok else 1)
