"""Host (CPU, integer) logic of the engine through the C ABI: orderings, supernode plan, KLU pivot search
and the static refactorization plan.  No GPU needed."""
import ctypes as C

import numpy as np
import pytest
import scipy.sparse as sp

from conftest import lap3d, load_matrix, lower_ccs, rand_spd
from kvxopt_b200 import _lib as L
from oracle import CholOracle

fn = L.fn


def analyze(A, perm=None, uplo="L", opts=None):
    cp, ri = A.indptr.astype(np.int64), A.indices.astype(np.int64)
    F = L.vp()
    p = np.ascontiguousarray(perm, dtype=np.int64) if perm is not None else None
    st = fn["b200s_chol_analyze"](A.shape[0], L.ptr_i64(cp), L.ptr_i64(ri), uplo.encode(), L.ptr_i64(p),
                                  C.byref(opts) if opts is not None else None, C.byref(F))
    return st, F


def plan_of(F):
    inf = L.CholInfo(); fn["b200s_chol_info"](F, C.byref(inf))
    n, ns = inf.n, inf.nsuper
    perm = np.zeros(n, np.int64); fn["b200s_chol_get_perm"](F, L.ptr_i64(perm))
    sup = np.zeros(ns + 1, np.int64); rp = np.zeros(ns + 1, np.int64)
    fn["b200s_chol_get_super"](F, L.ptr_i64(sup), L.ptr_i64(rp), None)
    rows = np.zeros(max(int(rp[-1]), 1), np.int64)
    fn["b200s_chol_get_super"](F, L.ptr_i64(sup), L.ptr_i64(rp), L.ptr_i64(rows))
    return inf, perm, sup, rp, rows


@pytest.mark.parametrize("n,dens,seed", [(1, 1.0, 0), (9, 0.4, 1), (80, 0.06, 2), (500, 0.01, 3)])
def test_supernode_structure_covers_exact_fill(n, dens, seed):
    A = rand_spd(n, dens, seed)
    Al = lower_ccs(A)
    st, F = analyze(Al)
    assert st == 0
    inf, perm, sup, rp, rows = plan_of(F)
    assert sorted(perm) == list(range(n))
    # exact fill pattern from a dense Cholesky of the permuted matrix
    Ad = A.toarray()[np.ix_(perm, perm)] + n * np.eye(n)
    Lc = np.linalg.cholesky(Ad)
    covered = np.zeros((n, n), dtype=bool)
    for s in range(inf.nsuper):
        c0, c1 = sup[s], sup[s + 1]
        r = rows[rp[s]:rp[s + 1]]
        assert list(r[:c1 - c0]) == list(range(c0, c1))
        assert np.all(np.diff(r) > 0)
        for j in range(c0, c1):
            covered[r[r >= j], j] = True
    assert np.all(covered[np.abs(Lc) > 1e-14])
    assert inf.nnz_L == covered.sum()
    # the independent oracle symbolic with the same permutation predicts the same exact fill count or less
    O = CholOracle(n, Al.indptr, Al.indices, "L", perm)
    assert O.nnzL <= inf.nnz_L
    fn["b200s_chol_free"](F)


def test_no_relaxation_gives_exact_fill():
    A = lap3d(8, 8, 4)
    Al = lower_ccs(A)
    o = L.CholOpts(); fn["b200s_chol_default_opts"](C.byref(o))
    for i in range(3):
        o.nrelax[i] = 0; o.zrelax[i] = 0.0
    st, F = analyze(Al, opts=o)
    inf, perm, *_ = plan_of(F)
    O = CholOracle(A.shape[0], Al.indptr, Al.indices, "L", perm)
    # oracle counts exact fill; with relaxation off the plan may only add 'free' merges (no explicit zeros)
    assert inf.nnz_L == O.nnzL
    fn["b200s_chol_free"](F)


def test_amd_reduces_fill_on_grid():
    A = lap3d(30, 30, 1)
    Al = lower_ccs(A)
    st, F = analyze(Al)
    inf_amd = plan_of(F)[0]
    o = L.CholOpts(); fn["b200s_chol_default_opts"](C.byref(o)); o.ordering = 1
    st, F2 = analyze(Al, opts=o)
    inf_nat = plan_of(F2)[0]
    assert inf_amd.nnz_L < 0.6 * inf_nat.nnz_L
    assert inf_amd.flops < inf_nat.flops
    fn["b200s_chol_free"](F); fn["b200s_chol_free"](F2)


def test_grid_nd_and_user_perm():
    nx, ny, nz = 12, 10, 7
    p = np.zeros(nx * ny * nz, np.int64)
    assert fn["b200s_grid_nd_perm"](nx, ny, nz, 16, L.ptr_i64(p)) == 0
    assert sorted(p) == list(range(nx * ny * nz))
    Al = lower_ccs(lap3d(nx, ny, nz))
    st, F = analyze(Al, perm=p)
    assert st == 0
    inf, perm, *_ = plan_of(F)
    assert sorted(perm) == list(range(nx * ny * nz))
    assert inf.nlevels >= 3
    fn["b200s_chol_free"](F)
    bad = p.copy(); bad[0] = bad[1]
    st, F = analyze(Al, perm=bad)
    assert st == L.INVALID and "permutation" in L.last_error()


def test_invalid_inputs():
    A = lower_ccs(rand_spd(10, 0.3, 0))
    st, _ = analyze(A, uplo="X")
    assert st == L.INVALID
    cp, ri = A.indptr.astype(np.int64), A.indices.astype(np.int64)
    ri2 = ri.copy(); ri2[0] = 99
    F = L.vp()
    assert fn["b200s_chol_analyze"](10, L.ptr_i64(cp), L.ptr_i64(ri2), b"L", None, None, C.byref(F)) == L.INVALID
    o = L.CholOpts(); fn["b200s_chol_default_opts"](C.byref(o)); o.supernodal = 3
    st, _ = analyze(A, opts=o)
    assert st == L.INVALID         # cholmod.options['supernodal'] is 0 (LDL' semantics), 1 or 2
    o.supernodal = 0
    st, F0 = analyze(A, opts=o)
    assert st == 0


def test_zero_size_and_upper():
    E = sp.csc_matrix((0, 0))
    st, F = analyze(E)
    assert st == 0
    assert fn["b200s_chol_factorize"](F, None, None, None, None) == 0          # n = 0 succeeds without a device
    assert fn["b200s_chol_solve"](F, 0, None, 0, 1) == 0
    fn["b200s_chol_free"](F)
    A = rand_spd(40, 0.1, 4)
    Au = sp.triu(A).tocsc(); Au.sort_indices()
    stU, FU = analyze(Au, uplo="U")
    stL, FL = analyze(lower_ccs(A))
    assert stU == 0 and stL == 0
    assert plan_of(FU)[0].nnz_L == plan_of(FL)[0].nnz_L
    fn["b200s_chol_free"](FU); fn["b200s_chol_free"](FL)


def test_numeric_rejects_entry_outside_analysed_pattern():
    """cholmod.c:322-398 rebuilds the matrix from A's own pattern; the engine compares it with the analysed one: a
    same-nnz but different pattern must not be scattered through the old map"""
    A = lower_ccs(rand_spd(30, 0.15, 7))
    st, F = analyze(A)
    assert st == 0
    cp, ri, vx = A.indptr.astype(np.int64), A.indices.astype(np.int64), A.data.copy()
    # move one off-diagonal entry of column 0 to a row the analysis has not seen (same nnz, different pattern)
    col0 = ri[cp[0]:cp[1]]
    free = [r for r in range(1, 30) if r not in set(col0.tolist())]
    assert len(col0) > 1 and free
    ri2 = ri.copy()
    ri2[cp[1] - 1] = max(free)
    seg = np.sort(ri2[cp[0]:cp[1]]); ri2[cp[0]:cp[1]] = seg
    st = fn["b200s_chol_factorize"](F, L.ptr_i64(cp), L.ptr_i64(ri2), L.ptr_f64(vx), None)
    if len(np.unique(seg)) == len(seg):
        assert st == L.INVALID and "outside the pattern" in L.last_error()
    # unsorted rows are refused as well
    ri3 = ri.copy(); ri3[cp[0]], ri3[cp[0] + 1] = ri[cp[0] + 1], ri[cp[0]]
    assert fn["b200s_chol_factorize"](F, L.ptr_i64(cp), L.ptr_i64(ri3), L.ptr_f64(vx), None) == L.INVALID
    fn["b200s_chol_free"](F)


# ---- KLU host ----------------------------------------------------------------------------------------

def klu_pivot(A):
    A = A.tocsc(); A.sort_indices()
    n = A.shape[0]
    cp, ri, vx = A.indptr.astype(np.int64), A.indices.astype(np.int64), A.data.astype(np.float64)
    S = L.vp()
    assert fn["b200s_klu_analyze"](n, L.ptr_i64(cp), L.ptr_i64(ri), C.byref(S)) == 0
    N = L.vp()
    st = fn["b200s_klu_pivot_host"](S, L.ptr_i64(cp), L.ptr_i64(ri), L.ptr_f64(vx), C.byref(N))
    return st, S, N, A


def klu_pattern(N, n):
    inf = L.KluInfo(); fn["b200s_klu_info"](N, C.byref(inf))
    Lp = np.zeros(n + 1, np.int64); Up = np.zeros(n + 1, np.int64); Fp = np.zeros(n + 1, np.int64)
    Li = np.zeros(max(inf.nnz_L, 1), np.int64); Ui = np.zeros(max(inf.nnz_U, 1), np.int64); Fi = np.zeros(max(inf.nnz_F, 1), np.int64)
    P = np.zeros(max(n, 1), np.int64); Q = np.zeros(max(n, 1), np.int64); R = np.zeros(inf.nblocks + 1, np.int64)
    assert fn["b200s_klu_extract"](N, L.ptr_i64(Lp), L.ptr_i64(Li), None, L.ptr_i64(Up), L.ptr_i64(Ui), None,
                                   L.ptr_i64(Fp), L.ptr_i64(Fi), None, L.ptr_i64(P), L.ptr_i64(Q), None, L.ptr_i64(R)) == 0
    return inf, Lp, Li[:inf.nnz_L], Up, Ui[:inf.nnz_U], Fp, Fi[:inf.nnz_F], P[:n], Q[:n], R


def identity_error(A, vals, Rs, pat, Lx, Ux, Fx):
    inf, Lp, Li, Up, Ui, Fp, Fi, P, Q, R = pat
    n = A.shape[0]
    Lm = sp.csc_matrix((Lx, Li, Lp), shape=(n, n)); Um = sp.csc_matrix((Ux, Ui, Up), shape=(n, n))
    Fm = sp.csc_matrix((Fx, Fi, Fp), shape=(n, n))
    A2 = sp.csc_matrix((vals, A.indices, A.indptr), shape=(n, n))
    return abs(sp.diags(1.0 / Rs) @ A2[P, :][:, Q] - (Lm @ Um + Fm)).sum(axis=0).max()


@pytest.mark.parametrize("name", ["bp_800", "bcsstk13", "bcsstk24", "ACTIVSg2000"])
def test_klu_pivot_identity_on_reference_matrices(name):
    """R P A Q = L U + F, the check of reference tests/test_sparse_solvers.py:216-237, on the pivot search"""
    st, S, N, A = klu_pivot(load_matrix(name))
    assert st == 0
    n = A.shape[0]
    pat = klu_pattern(N, n)
    inf = pat[0]
    Lx = np.zeros(max(inf.nnz_L, 1)); Ux = np.zeros(max(inf.nnz_U, 1)); Fx = np.zeros(max(inf.nnz_F, 1)); Rs = np.zeros(n)
    fn["b200s_klu_extract_host"](N, L.ptr_f64(Lx), L.ptr_f64(Ux), L.ptr_f64(Fx), L.ptr_f64(Rs))
    assert identity_error(A, A.data, Rs, pat, Lx[:inf.nnz_L], Ux[:inf.nnz_U], Fx[:inf.nnz_F]) < 1e-7
    assert sorted(pat[7]) == list(range(n)) and sorted(pat[8]) == list(range(n))
    if name == "ACTIVSg2000":
        assert inf.nblocks == 1                     # irreducible
    if name.startswith("bcsstk"):
        assert inf.nblocks == n                     # stored lower triangle => n singleton blocks
    fn["b200s_klu_free_numeric"](N); fn["b200s_klu_free_symbolic"](S)


@pytest.mark.parametrize("n,dens,seed", [(200, 0.03, 1), (600, 0.008, 2), (1500, 0.003, 3)])
def test_klu_pruned_pivot_search_on_unsymmetric_patterns(n, dens, seed):
    """The host pivot search prunes its depth-first search (Eisenstat-Liu, as KLU's kernel does): on unsymmetric patterns
    with a weak diagonal (off-diagonal pivots, several BTF blocks) the factors still satisfy R P A Q = L U + F, every column
    of L and U is sorted by row, L has a unit diagonal and |L| <= 1/tol."""
    rng = np.random.default_rng(seed)
    A = sp.random(n, n, dens, random_state=rng, format="csc") + sp.diags(0.05 * rng.standard_normal(n))
    st, S, N, A = klu_pivot(A)
    assert st == 0
    pat = klu_pattern(N, n)
    inf, Lp, Li, Up, Ui, Fp, Fi, P, Q, R = pat
    Lx = np.zeros(max(inf.nnz_L, 1)); Ux = np.zeros(max(inf.nnz_U, 1)); Fx = np.zeros(max(inf.nnz_F, 1)); Rs = np.zeros(n)
    fn["b200s_klu_extract_host"](N, L.ptr_f64(Lx), L.ptr_f64(Ux), L.ptr_f64(Fx), L.ptr_f64(Rs))
    # threshold pivoting with tol = 1e-3 lets the entries of U grow (1e6 - 1e8 here): the bound is relative to that growth
    assert identity_error(A, A.data, Rs, pat, Lx[:inf.nnz_L], Ux[:inf.nnz_U], Fx[:inf.nnz_F]) < 1e-13 * max(1.0, np.abs(Ux).max())
    for k in range(n):
        li = Li[Lp[k]:Lp[k + 1]]; ui = Ui[Up[k]:Up[k + 1]]
        assert li[0] == k and Lx[Lp[k]] == 1.0 and np.all(np.diff(li) > 0)
        assert ui[-1] == k and np.all(np.diff(ui) > 0)
    assert np.abs(Lx[:inf.nnz_L]).max() <= 1e3 * (1 + 1e-12)
    assert sorted(P) == list(range(n)) and sorted(Q) == list(range(n))
    fn["b200s_klu_free_numeric"](N); fn["b200s_klu_free_symbolic"](S)


def emulate_plan(N, A, vals):
    """replays the static refactorization plan with numpy exactly as the CUDA kernels do"""
    n = A.shape[0]
    v = L.KluPlanView(); assert fn["b200s_klu_plan_view"](N, C.byref(v)) == 0
    arr = lambda p, m: np.ctypeslib.as_array(p, shape=(m,)).copy() if m > 0 else np.zeros(0, np.int64)
    cbeg = arr(v.cbeg, n + 1); rowptr = arr(v.rowptr, n + 1); upd_ptr = arr(v.upd_ptr, n + 1); upd_dest = arr(v.upd_dest, v.nupd)
    udiag = arr(v.udiag_slot, n); ssrc = arr(v.slot_src, v.nslots); srow = arr(v.slot_row, v.nslots); rowent = arr(v.rowent, v.nnz_A)
    lp = arr(v.level_ptr, v.nlevels + 1); lc = arr(v.level_cols, n)
    uus = arr(v.upd_uslot, v.nupd); uls = arr(v.upd_lslot, v.nupd); ucnt = arr(v.upd_cnt, v.nupd)
    dest = arr(v.dest, v.ndest); lslot0 = arr(v.lslot0, n); fslot0 = arr(v.fslot0, n)
    Rs = np.ones(n)
    for i in range(n):
        e = rowent[rowptr[i]:rowptr[i + 1]]
        if len(e):
            Rs[i] = np.abs(vals[e]).max()
    LU = np.where(ssrc >= 0, vals[np.maximum(ssrc, 0)] / Rs[srow], 0.0)
    done_level = np.full(n, -1)
    for l in range(v.nlevels):
        for k in lc[lp[l]:lp[l + 1]]:
            for u in range(upd_ptr[k], upd_ptr[k + 1]):
                for t in range(ucnt[u]):
                    LU[dest[upd_dest[u] + t]] -= LU[uls[u] + t] * LU[uus[u]]
            LU[lslot0[k]:cbeg[k + 1]] /= LU[udiag[k]]
            done_level[k] = l
    return LU, Rs, cbeg, lslot0, fslot0


@pytest.mark.parametrize("case", ["rand60", "rand300", "bp_800"])
def test_klu_static_plan_reproduces_factorization_of_perturbed_values(case):
    if case.startswith("rand"):
        n = int(case[4:])
        rng = np.random.default_rng(n)
        A = sp.random(n, n, density=3.0 / n, random_state=rng, format="csc") + sp.identity(n) * 0.5
    else:
        A = load_matrix(case)
    st, S, N, A = klu_pivot(A)
    assert st == 0
    n = A.shape[0]
    pat = klu_pattern(N, n)
    inf, Lp, Li, Up, Ui, Fp, Fi, P, Q, R = pat
    vals = A.data * (1 + 1e-3 * np.random.default_rng(7).uniform(-1, 1, A.nnz))
    LU, Rs, cbeg, lslot0, fslot0 = emulate_plan(N, A, vals)
    Lx = np.zeros(inf.nnz_L); Ux = np.zeros(inf.nnz_U); Fx = np.zeros(inf.nnz_F)
    for k in range(n):
        nu = Up[k + 1] - Up[k]
        Ux[Up[k]:Up[k + 1]] = LU[cbeg[k]:cbeg[k] + nu]
        Lx[Lp[k]] = 1.0
        Lx[Lp[k] + 1:Lp[k + 1]] = LU[lslot0[k]:lslot0[k] + (Lp[k + 1] - Lp[k] - 1)]
        Fx[Fp[k]:Fp[k + 1]] = LU[fslot0[k]:fslot0[k] + Fp[k + 1] - Fp[k]]
    assert identity_error(A, vals, Rs, pat, Lx, Ux, Fx) < 1e-9
    fn["b200s_klu_free_numeric"](N); fn["b200s_klu_free_symbolic"](S)


def test_klu_structurally_and_numerically_singular():
    st, S, N, _ = klu_pivot(sp.csc_matrix(np.array([[1.0, 2], [2, 4]])))
    assert st == L.SINGULAR
    st, S, N, _ = klu_pivot(sp.csc_matrix(np.array([[1.0, 0, 0], [0, 0, 0], [0, 3, 1]])))
    assert st == L.SINGULAR


def test_klu_zero_diagonal_is_permuted_away():
    A = sp.csc_matrix(np.array([[0, 2.0, 0, 0], [3, 0, 0, 1], [0, 0, 0, 5], [1, 0, 4, 0]]))
    st, S, N, A = klu_pivot(A)
    assert st == 0
    fn["b200s_klu_free_numeric"](N); fn["b200s_klu_free_symbolic"](S)


# ---- device-side KKT solver: host-built pattern of S and assembly term lists (kkt_gpu.cu build_plan) ----------------

@pytest.mark.parametrize("seed,n,ml,p", [(0, 40, 90, 3), (1, 25, 25, 0), (2, 60, 200, 10)])
def test_kkt_assembly_plan_reproduces_H_plus_GtDG(seed, n, ml, p):
    """S = tril(H + G' diag(di)^2 G [+ A'A]) entry by entry from the constant term lists (misc.py:1418-1455 restated
    as one fixed-pattern weighted sum)"""
    rng = np.random.default_rng(seed)
    G = sp.random(ml, n, density=0.08, random_state=rng, format="csc"); G.sort_indices()
    A = sp.random(p, n, density=0.3, random_state=rng, format="csc"); A.sort_indices()
    H = sp.random(n, n, density=0.05, random_state=rng); H = (H + H.T + sp.identity(n)).tocsc()
    Hl = sp.tril(H).tocsc(); Hl.sort_indices()
    i64 = lambda a: np.ascontiguousarray(a, dtype=np.int64)
    arrs = [i64(G.indptr), i64(G.indices), G.data.copy(), i64(A.indptr), i64(A.indices), A.data.copy(), i64(Hl.indptr), i64(Hl.indices)]
    h = C.c_void_p()
    st = fn["b200s_kkt_create"](n, ml, p, L.ptr_i64(arrs[0]), L.ptr_i64(arrs[1]), L.ptr_f64(arrs[2]), L.ptr_i64(arrs[3]),
                                L.ptr_i64(arrs[4]), L.ptr_f64(arrs[5]), L.ptr_i64(arrs[6]), L.ptr_i64(arrs[7]), C.byref(h))
    assert st == 0, L.last_error()
    for sing in (0, 1):
        assert fn["b200s_kkt_set_singular"](h, sing) == 0
        inf = L.KktInfo(); fn["b200s_kkt_info"](h, C.byref(inf))
        assert inf.singular_mode == sing and inf.n == n and inf.ml == ml and inf.p == p
        Sp = np.ctypeslib.as_array(inf.Sp, shape=(n + 1,)).copy()
        Si = np.ctypeslib.as_array(inf.Si, shape=(inf.nnz_S,)).copy()
        di = rng.uniform(0.5, 2, ml); Sx = np.zeros(inf.nnz_S)
        assert fn["b200s_kkt_plan_check_host"](h, L.ptr_f64(di), L.ptr_f64(Hl.data.copy()), L.ptr_f64(Sx)) == 0
        S = sp.csc_matrix((Sx, Si, Sp), shape=(n, n))
        ref = H + G.T @ sp.diags(di ** 2) @ G
        if sing and p:
            ref = ref + A.T @ A
        assert abs(S - sp.tril(ref)).max() < 1e-13
        assert all(np.all(np.diff(Si[Sp[j]:Sp[j + 1]]) > 0) and (Sp[j] == Sp[j + 1] or Si[Sp[j]] >= j) for j in range(n))
    fn["b200s_kkt_free"](h)


def test_kkt_rejects_bad_input_and_needs_a_device_for_numeric_work():
    h = C.c_void_p()
    cp = np.array([0, 1, 2], dtype=np.int64); ri = np.array([0, 5], dtype=np.int64); vx = np.ones(2)
    st = fn["b200s_kkt_create"](2, 2, 0, L.ptr_i64(cp), L.ptr_i64(ri), L.ptr_f64(vx), None, None, None, None, None, C.byref(h))
    assert st == L.INVALID and "out of range" in L.last_error()
    ri = np.array([0, 1], dtype=np.int64)
    st = fn["b200s_kkt_create"](2, 2, 0, L.ptr_i64(cp), L.ptr_i64(ri), L.ptr_f64(vx), None, None, None, None, None, C.byref(h))
    assert st == 0
    if L.device_count() == 0:
        di = np.ones(2); minor = C.c_int64(0)
        assert fn["b200s_kkt_factor"](h, L.ptr_f64(di), None, C.byref(minor)) == L.NO_DEVICE       # no CPU fallback
        x = np.ones(2)
        assert fn["b200s_kkt_solve"](h, L.ptr_f64(x), None, L.ptr_f64(x)) == L.INVALID           # not factored
    fn["b200s_kkt_free"](h)


def test_kkt_ldl_assembles_the_reference_matrix_and_fails_loudly_without_gpu():
    """kkt.ldl (counterpart of misc.kkt_ldl, misc.py:1055-1130): the CCS values sent to the engine are the lower
    triangle of [[H + reg, A', G'W^-1], [A, -reg, 0], [W^-T G, 0, -I - reg]]; without a GPU factor() raises (no fallback)"""
    import scipy.sparse as sp
    from kvxopt_b200 import kkt, _lib
    rng = np.random.default_rng(3)
    n, p, m = 12, 3, 20
    G = sp.random(m, n, density=0.3, random_state=rng, format="csc")
    A = sp.random(p, n, density=0.5, random_state=rng, format="csc")
    Hh = sp.random(n, n, density=0.2, random_state=rng, format="csc"); H = (Hh + Hh.T + 5 * sp.identity(n)).tocsc()
    di = np.exp(rng.standard_normal(m))
    for reg in (None, 1e-3):
        for Hm in (H, None):
            factor = kkt.ldl(G, {"l": m, "q": [], "s": []}, A, kktreg=reg)
            try:
                factor({"di": di.copy()}, Hm)
                assert _lib.device_count() > 0
            except RuntimeError:
                assert _lib.device_count() == 0
            st = factor._state
            K = sp.csc_matrix((st["kv"], st["ki"], st["kp"]), shape=(n + p + m, n + p + m)).toarray()
            r = reg or 0.0
            want = np.zeros_like(K)
            if Hm is not None:
                want[:n, :n] = np.tril(H.toarray())
            want[np.arange(n), np.arange(n)] += r
            want[n:n + p, :n] = A.toarray()
            want[n + p:, :n] = di[:, None] * G.toarray()
            want[np.arange(n, n + p), np.arange(n, n + p)] = -r
            want[np.arange(n + p, n + p + m), np.arange(n + p, n + p + m)] = -1.0 - r
            assert np.abs(K - want).max() == 0.0


def test_kkt_ldl2_assembles_the_reduced_matrix():
    """kkt.ldl2 (counterpart of misc.kkt_ldl2, misc.py:1128-1210): the values the engine factors are the lower triangle of
    [[H + G' W^-2 G, A'], [A, 0]], a fixed sparse linear map (the list of products g_ki g_kj, H, A) of w = [di^2; H; A] that
    the device evaluates (b200s_spmv_apply); here the map and w are taken from the solver's state and applied with scipy, and
    on a GPU box the device's own result is read back and compared as well"""
    import ctypes as C
    import scipy.sparse as sp
    from kvxopt_b200 import kkt, _lib
    rng = np.random.default_rng(3)
    n, p, m = 12, 3, 20
    G = sp.random(m, n, density=0.3, random_state=rng, format="csc")
    A = sp.random(p, n, density=0.5, random_state=rng, format="csc")
    Hh = sp.random(n, n, density=0.2, random_state=rng, format="csc"); H = (Hh + Hh.T + 5 * sp.identity(n)).tocsc()
    di = np.exp(rng.standard_normal(m))
    for Hm in (H, None):
        factor = kkt.ldl2(G, {"l": m, "q": [], "s": []}, A)
        try:
            factor({"di": di.copy()}, Hm)
            assert _lib.device_count() > 0
        except RuntimeError:
            assert _lib.device_count() == 0
        st = factor._state
        mp, mc, mv = st["asm_csr"]
        kv = sp.csr_matrix((mv, mc, mp), shape=(st["nk"], len(st["w"]))) @ st["w"]
        if _lib.device_count() > 0:
            kd = np.zeros(st["nk"])
            assert _lib.fn["b200s_spmv_get"](st["asm"].h, _lib.ptr_f64(kd)) == 0
            assert np.abs(kd - kv).max() <= 1e-13 * np.abs(kv).max()
        K = sp.csc_matrix((kv, st["ki"], st["kp"]), shape=(n + p, n + p)).toarray()
        S = (G.T @ sp.diags(di * di) @ G).toarray() + (H.toarray() if Hm is not None else 0.0)
        want = np.zeros((n + p, n + p)); want[:n, :n] = np.tril(S); want[n:, :n] = A.toarray()
        assert np.abs(K - want).max() <= 1e-13 * np.abs(want).max()


def _synthetic_unsym(kind, seed=0):
    rng = np.random.default_rng(seed)
    if kind == "banded_dense_tail":
        n = 600
        M = sp.diags([rng.uniform(1, 2, n - 1), rng.uniform(4, 5, n), rng.uniform(1, 2, n - 1), rng.uniform(0.1, 1, n - 7)], [-1, 0, 1, 7]).tolil()
        M[n - 90:, n - 90:] = rng.uniform(-1, 1, (90, 90)) + 12 * np.eye(90)
        M[:, n - 3] = rng.uniform(0.1, 1, (n, 1))
        return M.tocsc()
    if kind == "blocks2x2":          # 2 x 2 blocked, structurally symmetric, like a power-flow Jacobian
        nb = 400
        G = sp.random(nb, nb, density=0.006, random_state=rng)
        G = (G + G.T + sp.identity(nb)).tocsr()
        G.data[:] = 1.0
        M = sp.kron(G, np.ones((2, 2))).tocsc()
        M.data[:] = rng.uniform(-1, 1, M.nnz)
        return (M + sp.identity(2 * nb) * 6).tocsc()
    if kind == "random":
        n = 500
        return (sp.random(n, n, density=0.01, random_state=rng) + sp.identity(n) * 3).tocsc()
    raise ValueError(kind)


@pytest.mark.parametrize("early_minw", [None, 8])
@pytest.mark.parametrize("name", ["ACTIVSg2000", "bp_800", "bcsstk13", "banded_dense_tail", "blocks2x2", "random"])
def test_klu_wave_plan_replayed_on_the_host_matches_the_pivoting_factorization(name, early_minw, monkeypatch):
    """The wave schedule the CUDA refactorization kernel executes (supernode source blocks -> staged pieces -> per-column
    records, in-wave updates, dense trailing block) replayed by the host interpreter b200s_klu_plan_emulate_host: same L, U,
    F and row scales as the pivoting Gilbert-Peierls factorization of the same matrix, and -- with perturbed values -- as an
    independent dense check P R^-1 A Q = L U + F."""
    from conftest import load_matrix
    if early_minw is not None:      # narrower "wide level" threshold: the early-column path (k_klu_early's tables) on small patterns too
        monkeypatch.setenv("B200S_KLU_EARLY_MINW", str(early_minw))
    A = (load_matrix(name) if name[0].isupper() or name.startswith("b") and name[1] in "pc" else _synthetic_unsym(name)).tocsc()
    A.sort_indices()
    n = A.shape[0]
    cp, ri, vx = A.indptr.astype(np.int64), A.indices.astype(np.int64), A.data.astype(np.float64)
    S = L.vp(); assert fn["b200s_klu_analyze"](n, L.ptr_i64(cp), L.ptr_i64(ri), C.byref(S)) == 0
    N = L.vp(); assert fn["b200s_klu_pivot_host"](S, L.ptr_i64(cp), L.ptr_i64(ri), L.ptr_f64(vx), C.byref(N)) == 0
    inf = L.KluInfo(); fn["b200s_klu_info"](N, C.byref(inf))
    v = L.KluPlanView(); fn["b200s_klu_plan_view"](N, C.byref(v))
    if name == "ACTIVSg2000":
        assert v.nearly > 2000 and v.nearly_levels >= 4
    if early_minw is not None and name in ("bp_800", "bcsstk13", "blocks2x2"):
        assert v.nearly > 0
    if not v.wave_ok:       # pattern outside the wave kernel's budget: the level-schedule kernel serves it
        assert fn["b200s_klu_plan_emulate_host"](N, L.ptr_f64(vx), None, None, None, None) == L.INVALID
        fn["b200s_klu_free_numeric"](N); fn["b200s_klu_free_symbolic"](S)
        return
    assert v.npiece_users >= v.npieces and (v.npieces > 0) == (v.nbatches > 0)
    Lx0, Ux0, Fx0, Rs0 = np.zeros(inf.nnz_L), np.zeros(inf.nnz_U), np.zeros(max(inf.nnz_F, 1)), np.zeros(n)
    assert fn["b200s_klu_extract_host"](N, L.ptr_f64(Lx0), L.ptr_f64(Ux0), L.ptr_f64(Fx0), L.ptr_f64(Rs0)) == 0
    Lx, Ux, Fx, Rs = np.zeros(inf.nnz_L), np.zeros(inf.nnz_U), np.zeros(max(inf.nnz_F, 1)), np.zeros(n)
    st = fn["b200s_klu_plan_emulate_host"](N, L.ptr_f64(vx), L.ptr_f64(Lx), L.ptr_f64(Ux), L.ptr_f64(Fx), L.ptr_f64(Rs))
    assert st == 0, L.last_error()
    assert np.array_equal(Rs, Rs0)
    scale = np.abs(Ux0).max()
    assert np.abs(Ux - Ux0).max() <= 1e-12 * scale and np.abs(Lx - Lx0).max() <= 1e-11
    if inf.nnz_F:
        assert np.abs(Fx[:inf.nnz_F] - Fx0[:inf.nnz_F]).max() <= 1e-14 * max(1.0, np.abs(Fx0).max())
    # perturbed values, same pattern and pivots (klu_refactor semantics): residual of the factorization identity
    rng = np.random.default_rng(3)
    v2 = vx * (1 + 1e-3 * rng.uniform(-1, 1, vx.size))
    assert fn["b200s_klu_plan_emulate_host"](N, L.ptr_f64(v2), L.ptr_f64(Lx), L.ptr_f64(Ux), L.ptr_f64(Fx), L.ptr_f64(Rs)) == 0
    Lp, Up, Fp = np.zeros(n + 1, np.int64), np.zeros(n + 1, np.int64), np.zeros(n + 1, np.int64)
    Li, Ui, Fi = np.zeros(inf.nnz_L, np.int64), np.zeros(inf.nnz_U, np.int64), np.zeros(max(inf.nnz_F, 1), np.int64)
    P, Q = np.zeros(n, np.int64), np.zeros(n, np.int64)
    fn["b200s_klu_extract"](N, L.ptr_i64(Lp), L.ptr_i64(Li), None, L.ptr_i64(Up), L.ptr_i64(Ui), None, L.ptr_i64(Fp), L.ptr_i64(Fi), None,
                            L.ptr_i64(P), L.ptr_i64(Q), None, None)
    Lm = sp.csc_matrix((Lx, Li, Lp), shape=(n, n)); Um = sp.csc_matrix((Ux, Ui, Up), shape=(n, n))
    Fm = sp.csc_matrix((Fx[:inf.nnz_F], Fi[:inf.nnz_F], Fp), shape=(n, n))
    A2 = sp.csc_matrix((v2, A.indices, A.indptr), shape=(n, n))
    lhs = sp.diags(1.0 / Rs) @ A2[P, :][:, Q]
    res = abs(lhs - (Lm @ Um + Fm)).max()
    assert res <= 1e-12 * max(1.0, abs(Um).max())
    fn["b200s_klu_free_numeric"](N); fn["b200s_klu_free_symbolic"](S)


@pytest.mark.parametrize("name", ["ACTIVSg2000", "bp_800", "bcsstk13", "unsym_offdiag", "random", "arrow"])
@pytest.mark.parametrize("trans", [0, 1])
def test_klu_one_matrix_solve_tape_replayed_on_the_host(name, trans):
    """The operation tape of the one-matrix solve kernel (k_klu_solve_one: klu_solve / klu_tsolve of klu.c:593-690 as column
    operations in execution order, cut into shared-memory chunks) replayed by b200s_klu_solve_tape_host with the values of
    the pivot search: A x = b ('N') and A' x = b ('T') to SuperLU's solution, several BTF blocks and off-diagonal pivots
    included, a right-hand side with a leading dimension > n."""
    import scipy.sparse.linalg as spla
    if name == "unsym_offdiag":
        rng = np.random.default_rng(11)
        n = 400
        A = (sp.random(n, n, density=0.02, random_state=rng) + sp.identity(n) * 1e-6 +
             sp.csc_matrix((rng.uniform(1, 2, n), (rng.permutation(n), np.arange(n))), shape=(n, n))).tocsc()
    elif name == "random":
        A = _synthetic_unsym("random")
    elif name == "arrow":             # a column of U and a row of L longer than one chunk of the tape: split operations
        n = 2600
        rng = np.random.default_rng(3)
        A = sp.lil_matrix((n, n)); A.setdiag(rng.uniform(2, 3, n))
        A[n - 1, :] = rng.uniform(-1e-2, 1e-2, n); A[:, n - 1] = rng.uniform(-1e-2, 1e-2, (n, 1)); A[n - 1, n - 1] = 4.0
        A = A.tocsc()
    else:
        A = load_matrix(name).tocsc()
    st, S, N, A = klu_pivot(A)
    assert st == 0
    n = A.shape[0]
    if name == "arrow":
        inf = L.KluInfo(); fn["b200s_klu_info"](N, C.byref(inf))
        assert inf.nblocks == 1 and inf.nnz_U >= 2 * n - 1
    ld = n + 3
    rng = np.random.default_rng(5)
    B = np.full((2, ld), np.nan); B[:, :n] = rng.standard_normal((2, n))
    X = B.copy()
    assert fn["b200s_klu_solve_tape_host"](N, trans, L.ptr_f64(X), 2, ld) == 0
    assert np.isnan(X[:, n:]).all()                      # the padding rows are not touched
    M = (A.T if trans else A).tocsc()
    Xref = spla.splu(M).solve(B[:, :n].T).T
    assert np.linalg.norm(X[:, :n] - Xref) / np.linalg.norm(Xref) < 1e-10
    assert fn["b200s_klu_solve_tape_host"](N, 2, L.ptr_f64(X), 2, ld) == L.INVALID
    assert fn["b200s_klu_solve_tape_host"](N, trans, L.ptr_f64(X), 2, n - 1) == L.INVALID
    fn["b200s_klu_free_numeric"](N); fn["b200s_klu_free_symbolic"](S)


@pytest.mark.parametrize("name", ["ACTIVSg2000", "bp_800", "unsym_offdiag", "random"])
def test_klu_pivot_rule_against_an_independent_run(name):
    """The pivot RULE of the host pivot search (threshold partial pivoting, tol 1e-3, diagonal preferred, on the row-scaled
    matrix; klu_defaults) against an independent run: the CPU oracle is given only the SYMBOLIC pre-ordering (BTF + AMD, before
    any numeric pivoting) and chooses its pivots itself; its pivot sequence must be the product's, row by row."""
    from conftest import load_matrix
    from oracle import KluOracle
    if name == "unsym_offdiag":       # small diagonal entries: most pivots are off the diagonal
        rng = np.random.default_rng(11)
        n = 400
        A = (sp.random(n, n, density=0.02, random_state=rng) + sp.identity(n) * 1e-6 +
             sp.csc_matrix((rng.uniform(1, 2, n), (rng.permutation(n), np.arange(n))), shape=(n, n))).tocsc()
    elif name == "random":
        A = _synthetic_unsym("random")
    else:
        A = load_matrix(name).tocsc()
    A.sort_indices()
    n = A.shape[0]
    cp, ri, vx = A.indptr.astype(np.int64), A.indices.astype(np.int64), A.data.astype(np.float64)
    S = L.vp(); assert fn["b200s_klu_analyze"](n, L.ptr_i64(cp), L.ptr_i64(ri), C.byref(S)) == 0
    N = L.vp(); assert fn["b200s_klu_pivot_host"](S, L.ptr_i64(cp), L.ptr_i64(ri), L.ptr_f64(vx), C.byref(N)) == 0
    P0, Q0 = np.zeros(n, np.int64), np.zeros(n, np.int64)
    assert fn["b200s_klu_symbolic_perm"](S, L.ptr_i64(P0), L.ptr_i64(Q0), None, None) == 0
    Pn, Qn = np.zeros(n, np.int64), np.zeros(n, np.int64)
    fn["b200s_klu_extract"](N, None, None, None, None, None, None, None, None, None, L.ptr_i64(Pn), L.ptr_i64(Qn), None, None)
    assert np.array_equal(Qn, Q0)                       # numeric pivoting permutes rows only
    O = KluOracle(n, cp, ri, vx, P0=P0, Q=Q0)            # the oracle's own pivoting from the symbolic pre-ordering
    assert np.array_equal(O.pnum, Pn)
    if name == "unsym_offdiag":
        assert (Pn != P0).sum() > n // 4                # the case really pivots off the diagonal
    inf = L.KluInfo(); fn["b200s_klu_info"](N, C.byref(inf))
    # same pivots => same L pattern (L is block diagonal in BTF form); with several blocks the one-block oracle fills the
    # off-diagonal part of U that KLU keeps unfactored in F
    assert O.nnz_L == inf.nnz_L and (inf.nblocks > 1 or O.nnz_U == inf.nnz_U)
    fn["b200s_klu_free_numeric"](N); fn["b200s_klu_free_symbolic"](S)


def test_persistent_solve_schedules_replayed_on_the_host():
    """k_fwd_persist / k_bwd_persist run every block step of a level's large fronts in ONE kernel whose CTAs wait for each other
    through flags: the static work lists and the waits are replayed on the CPU (b200s_persist_schedule_check, the same
    host/device iterators the kernels use) -- every (block, tile) exactly once and in block order, every block solved once
    after the rows it reads, no CTA left waiting -- over the shapes of the 100^3 root front, fronts with nr == nc (nothing
    below the last block), odd pivot counts, more CTAs than blocks, and random shapes."""
    import random
    from kvxopt_b200 import _lib
    chk = _lib.fn["b200s_persist_schedule_check"]
    cases = [(15000, 14900, 148), (15000, 14900, 1), (10000, 10000, 79), (5000, 2500, 37), (129, 1, 1), (129, 129, 2),
             (300, 200, 3), (256, 256, 2), (257, 257, 3), (1000, 999, 8), (1000, 1000, 8), (640, 130, 5), (193, 65, 2),
             (2563, 1283, 21), (385, 383, 4)]
    rng = random.Random(7)
    for _ in range(600):
        nr = rng.randint(129, 4000)
        cases.append((nr, rng.randint(1, nr), rng.randint(1, (nr + 127) // 128)))
    for nr, nc, g in cases:
        assert chk(nr, nc, g) == 0, (nr, nc, g, chk(nr, nc, g))
    assert chk(100, 200, 1) == -1 and chk(100, 50, 0) == -1


@pytest.mark.parametrize("case", ["bcsstk24", "lap24_nd", "kkt3x3"])
def test_assembly_and_gather_child_lists_match_their_definition(case):
    """k_extend_add and k_fwd_gather visit, per column range / 512-row chunk of a front, only the children listed for it
    (b200s_chol_child_lists_check rebuilds the lists for every front of the plan and compares them with the definition: a child
    is listed exactly when one of its relative indices falls into the range, in child order, one-row children flagged).
    'kkt3x3' is the unreduced KKT matrix of a box-constrained QP in the z -> x order of kkt.ldl: fronts with hundreds of
    one-entry children, the structure that made every CTA of the root front walk 18 800 children at full size."""
    from bench import lap3d_lower
    perm = None
    if case == "bcsstk24":
        A = sp.tril(load_matrix("bcsstk24")).tocsc()
    elif case == "lap24_nd":
        A = lap3d_lower(24)
        perm = np.zeros(A.shape[0], np.int64)
        fn["b200s_grid_nd_perm"](24, 24, 24, 64, L.ptr_i64(perm))
    else:
        import os, sys
        sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
        from generators import qp_instance
        P, q, G, h = qp_instance(60, 50, 300)
        n, m = P.shape[0], G.shape[0]
        K = sp.bmat([[sp.tril(P), None], [G, -sp.identity(m)]]).tocsc()
        A = sp.tril(K).tocsc()
        perm = np.concatenate([n + np.arange(m), np.arange(n)]).astype(np.int64)       # z first, then x
    A.sort_indices()
    N = A.shape[0]
    o = L.CholOpts(); fn["b200s_chol_default_opts"](C.byref(o))
    if case == "kkt3x3":
        o.supernodal = 0
    h = C.c_void_p()
    assert fn["b200s_chol_analyze"](N, L.ptr_i64(A.indptr.astype(np.int64)), L.ptr_i64(A.indices.astype(np.int64)), b"L",
                                    L.ptr_i64(perm) if perm is not None else None, C.byref(o), C.byref(h)) == 0
    inf = L.CholInfo(); fn["b200s_chol_info"](h, C.byref(inf))
    assert fn["b200s_chol_child_lists_check"](h) == 0, L.last_error()
    if case == "kkt3x3":
        ns = inf.nsuper
        parent = np.zeros(ns, np.int64); nrows = np.zeros(ns, np.int64); ncols = np.zeros(ns, np.int64)
        assert fn["b200s_chol_front_layout"](h, L.ptr_i64(parent), None, L.ptr_i64(ncols), L.ptr_i64(nrows), None, None, None, None) == 0
        kids = np.bincount(parent[parent >= 0], minlength=ns)
        assert kids.max() > 100                                # a many-children front (lap24_nd has the multi-chunk fronts)
    fn["b200s_chol_free"](h)
    assert fn["b200s_chol_child_lists_check"](None) == L.INVALID
