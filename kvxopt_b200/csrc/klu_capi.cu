// extern "C" boundary, KLU half (include/b200sparse.h).
#include "../../include/b200sparse.h"
#include "gpu.hpp"
#include "nvtx_range.hpp"
#include "klu_host.hpp"
#include <cstring>
#include <new>
#include <stdexcept>
#include <vector>
#include <chrono>
#include <cstdio>

using namespace b200s;

namespace b200s {
class KluDevice;
KluDevice* klu_device_create(const KluPlan& P, const KluNumeric& N, const KluSymbolic& S, int device, int* status);
void klu_device_destroy(KluDevice* d);
int klu_device_init_refactor(KluDevice* d, const KluPlan& P);
int klu_device_load_host_factor(KluDevice* d, const double* slots_host, const double* rs_host);
int klu_device_refactor(KluDevice* d, const double* vals, bool on_device, long long batch, long long ldv, int* status);
int klu_device_refactor_begin(KluDevice* d, const double* vals, long long batch, long long ldv);
int klu_device_refactor_end(KluDevice* d, int* status);
int klu_device_solve(KluDevice* d, int trans, double* B, long long nrhs, long long ldB, long long batch, bool on_device);
int klu_device_extract(KluDevice* d, long long b, double* slots_host, double* rs_host);
int klu_solve_tape_host(const KluSymbolic& S, const KluNumeric& N, const KluPlan& P, int trans, const double* slots, double* B,
                        long long nrhs, long long ldB);
void klu_device_times(const KluDevice* d, double* h2d, double* refactor, double* solve, double* kernel, double* dense, long long* launches);
}  // namespace b200s

struct b200s_klu_sym {
    KluSymbolic S;
};
struct b200s_klu_num {
    const b200s_klu_sym* sym = nullptr;
    KluSymbolic S;         // copy: the numeric object must outlive nothing but itself
    KluNumeric N;
    KluPlan P;
    KluDevice* dev = nullptr;
    int device = 0;
    // complex matrices (b200s_klu_factor_z): the complex factor of the host pivot search (get_numeric / get_det) and the
    // factorization of the real embedding a + ib -> [[a, -b], [b, a]] of order 2n that serves the solves on the device
    KluNumericZ* zN = nullptr;
    b200s_klu_sym* emb_sym = nullptr;
    b200s_klu_num* emb = nullptr;
};

extern "C" {

b200s_status b200s_klu_analyze(b200s_int n, const b200s_int* colptr, const b200s_int* rowind, b200s_klu_sym** out) {
    B200S_NVTX("b200s_klu_analyze");
    if (!out) return B200S_INVALID;
    *out = nullptr;
    if (n < 0 || (n > 0 && (!colptr || (colptr[n] > 0 && !rowind)))) return B200S_INVALID;
    b200s_klu_sym* S = new (std::nothrow) b200s_klu_sym();
    if (!S) return B200S_OUT_OF_MEMORY;
    try {
        static const b200s_int zero = 0;
        klu_analyze(n, n > 0 ? colptr : &zero, rowind, S->S);
    } catch (const std::bad_alloc&) {
        delete S; return B200S_OUT_OF_MEMORY;
    } catch (const std::exception& e) {
        set_last_error(e.what()); delete S; return B200S_INVALID;
    }
    *out = S;
    return B200S_OK;
}

static b200s_status factor_impl(b200s_klu_sym* S, const b200s_int* colptr, const b200s_int* rowind, const double* val,
                                b200s_klu_num** out, bool with_device) {
    B200S_NVTX("factor_impl");
    if (!S || !out) return B200S_INVALID;
    *out = nullptr;
    const i32 n = S->S.n;
    if (n > 0) {
        if (!colptr || !val) return B200S_INVALID;
        if (colptr[n] != S->S.nnz) { set_last_error("matrix pattern differs from the analysed pattern"); return B200S_INVALID; }
        for (i32 j = 0; j <= n; j++) if (colptr[j] != S->S.Ap[j]) { set_last_error("matrix pattern differs from the analysed pattern"); return B200S_INVALID; }
        for (i64 p = 0; p < S->S.nnz; p++) if (rowind[p] != S->S.Ai[p]) { set_last_error("matrix pattern differs from the analysed pattern"); return B200S_INVALID; }
    }
    b200s_klu_num* N = new (std::nothrow) b200s_klu_num();
    if (!N) return B200S_OUT_OF_MEMORY;
    N->S = S->S;
    N->device = current_device();
    try {
        const bool tdbg = getenv("B200S_DEBUG") != nullptr;
        auto tk0 = std::chrono::steady_clock::now();
        int st = klu_factor(N->S, val, N->N);
        if (tdbg) fprintf(stderr, "[b200s klu] pivoting factorization (host) %.1f ms\n", std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - tk0).count());
        if (st != ST_OK) { delete N; return (b200s_status)st; }
        // with a device, klu.numeric() serves its own matrix from the values of the pivot search and only needs the slot
        // layout; the refactorization tables are built when a batch is first refactored (ensure_refactor_tables)
        klu_build_plan(N->S, N->N, N->P, !with_device || getenv("B200S_KLU_EAGER_PLAN") != nullptr);
    } catch (const std::bad_alloc&) {
        delete N; return B200S_OUT_OF_MEMORY;
    } catch (const std::exception& e) {
        set_last_error(e.what()); delete N; return B200S_INVALID;
    }
    if (n > 0 && with_device) {
        // the numeric values used by solve/extract come from the device refactorization (batch of one)
        int st = ST_OK;
        const bool tdbg = getenv("B200S_DEBUG") != nullptr;
        auto td0 = std::chrono::steady_clock::now();
        N->dev = klu_device_create(N->P, N->N, N->S, N->device, &st);
        if (!N->dev) { delete N; return (b200s_status)st; }
        auto td1 = std::chrono::steady_clock::now();
        {
            // the factor of THIS matrix: the values the pivoting factorization just computed, in the plan's slot layout
            const KluPlan& P = N->P; const KluNumeric& M = N->N;
            std::vector<double> slots((size_t)std::max<i64>(P.nslots, 1), 0.0);
            for (i32 k = 0; k < n; k++) {
                for (i64 p = M.Up[k]; p < M.Up[k + 1]; p++) slots[P.cbeg[k] + (p - M.Up[k])] = M.Ux[p];
                for (i64 p = M.Lp[k] + 1; p < M.Lp[k + 1]; p++) slots[P.lslot0[k] + (p - M.Lp[k] - 1)] = M.Lx[p];
                for (i64 p = M.Fp[k]; p < M.Fp[k + 1]; p++) slots[P.fslot0[k] + (p - M.Fp[k])] = M.Fx[p];
            }
            st = klu_device_load_host_factor(N->dev, slots.data(), M.Rs.data());
        }
        if (tdbg) fprintf(stderr, "[b200s klu] device tables %.1f ms, factor values to the device %.1f ms\n",
                          std::chrono::duration<double, std::milli>(td1 - td0).count(),
                          std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - td1).count());
        if (st != ST_OK) { klu_device_destroy(N->dev); delete N; return (b200s_status)st; }
    }
    *out = N;
    return B200S_OK;
}

b200s_status b200s_klu_factor(b200s_klu_sym* S, const b200s_int* colptr, const b200s_int* rowind, const double* val,
                              b200s_klu_num** out) {
    return factor_impl(S, colptr, rowind, val, out, true);
}
b200s_status b200s_klu_pivot_host(b200s_klu_sym* S, const b200s_int* colptr, const b200s_int* rowind, const double* val,
                                  b200s_klu_num** out) {
    return factor_impl(S, colptr, rowind, val, out, false);
}
b200s_status b200s_klu_extract_host(const b200s_klu_num* N, double* Lx, double* Ux, double* Fx, double* Rs) {
    if (!N) return B200S_INVALID;
    if (Lx) for (size_t p = 0; p < N->N.Lx.size(); p++) Lx[p] = N->N.Lx[p];
    if (Ux) for (size_t p = 0; p < N->N.Ux.size(); p++) Ux[p] = N->N.Ux[p];
    if (Fx) for (size_t p = 0; p < N->N.Fx.size(); p++) Fx[p] = N->N.Fx[p];
    if (Rs) for (size_t p = 0; p < N->N.Rs.size(); p++) Rs[p] = N->N.Rs[p];
    return B200S_OK;
}

// update lists + kernel schedules of the refactorization, built (and uploaded) the first time they are needed
static b200s_status ensure_refactor_tables(b200s_klu_num* N) {
    if (N->P.have_refactor) return B200S_OK;
    B200S_NVTX("klu_refactor_tables");
    try {
        klu_build_plan(N->S, N->N, N->P, true);
    } catch (const std::bad_alloc&) {
        return B200S_OUT_OF_MEMORY;
    } catch (const std::exception& e) {
        set_last_error(e.what()); return B200S_INVALID;
    }
    if (N->dev) return (b200s_status)klu_device_init_refactor(N->dev, N->P);
    return B200S_OK;
}

static b200s_status refactor_impl(b200s_klu_num* N, const double* vals, bool on_device, b200s_int batch, b200s_int ldv,
                                  int* status_per_matrix) {
    B200S_NVTX("refactor_impl");
    if (!N || batch < 0) return B200S_INVALID;
    if (N->N.n == 0 || batch == 0) return B200S_OK;
    if (!vals || ldv < N->S.nnz || batch > 0x7fffff00) return B200S_INVALID;
    if (!N->dev) return B200S_NO_DEVICE;
    if (b200s_status st = ensure_refactor_tables(N)) return st;
    return (b200s_status)klu_device_refactor(N->dev, vals, on_device, batch, ldv, status_per_matrix);
}
b200s_status b200s_klu_refactor_batch(b200s_klu_num* N, const double* vals, b200s_int batch, b200s_int ldv, int* status_per_matrix) {
    return refactor_impl(N, vals, false, batch, ldv, status_per_matrix);
}
b200s_status b200s_klu_refactor_batch_dev(b200s_klu_num* N, const double* vals_dev, b200s_int batch, b200s_int ldv, int* status_per_matrix) {
    return refactor_impl(N, vals_dev, true, batch, ldv, status_per_matrix);
}

b200s_status b200s_klu_refactor_batch_begin(b200s_klu_num* N, const double* vals, b200s_int batch, b200s_int ldv) {
    B200S_NVTX("b200s_klu_refactor_batch_begin");
    if (!N || batch < 0) return B200S_INVALID;
    if (N->N.n == 0 || batch == 0) return B200S_OK;
    if (!vals || ldv < N->S.nnz || batch > 0x7fffff00) return B200S_INVALID;
    if (!N->dev) return B200S_NO_DEVICE;
    if (b200s_status st = ensure_refactor_tables(N)) return st;
    return (b200s_status)klu_device_refactor_begin(N->dev, vals, batch, ldv);
}
b200s_status b200s_klu_refactor_batch_end(b200s_klu_num* N, int* status_per_matrix) {
    B200S_NVTX("b200s_klu_refactor_batch_end");
    if (!N) return B200S_INVALID;
    if (N->N.n == 0) return B200S_OK;
    if (!N->dev) return B200S_NO_DEVICE;
    return (b200s_status)klu_device_refactor_end(N->dev, status_per_matrix);
}

static b200s_status solve_impl(b200s_klu_num* N, int trans, double* B, b200s_int nrhs, b200s_int ldB, b200s_int batch, bool on_device) {
    B200S_NVTX("solve_impl");
    if (!N || nrhs < 0 || batch < 0 || (trans != 0 && trans != 1)) return B200S_INVALID;
    if (N->N.n == 0 || nrhs == 0 || batch == 0) return B200S_OK;
    if (!B || ldB < N->N.n) return B200S_INVALID;
    if (!N->dev) return B200S_NO_DEVICE;
    return (b200s_status)klu_device_solve(N->dev, trans, B, nrhs, ldB, batch, on_device);
}
b200s_status b200s_klu_solve_batch(b200s_klu_num* N, int trans, double* B, b200s_int nrhs, b200s_int ldB, b200s_int batch) {
    return solve_impl(N, trans, B, nrhs, ldB, batch, false);
}
b200s_status b200s_klu_solve_batch_dev(b200s_klu_num* N, int trans, double* B_dev, b200s_int nrhs, b200s_int ldB, b200s_int batch) {
    return solve_impl(N, trans, B_dev, nrhs, ldB, batch, true);
}
b200s_status b200s_klu_solve(b200s_klu_num* N, int trans, double* B, b200s_int nrhs, b200s_int ldB) {
    return solve_impl(N, trans, B, nrhs, ldB, 1, false);
}

b200s_status b200s_klu_info(const b200s_klu_num* N, b200s_klu_info_t* info) {
    if (!N || !info) return B200S_INVALID;
    memset(info, 0, sizeof *info);
    if (N->zN) {          // complex factor object: the counts of the complex factor, the timings of the embedded device object
        const i32 nz = N->zN->n;
        info->n = nz; info->nblocks = N->S.nblocks; info->nnz_A = N->S.nnz; info->max_block = N->S.maxblock;
        info->nnz_L = N->zN->Lp[nz]; info->nnz_U = N->zN->Up[nz]; info->nnz_F = N->zN->Fp[nz];
        info->flops = 4.0 * N->zN->flops;
        info->bytes_per_refactor = 16 * (info->nnz_A + info->nnz_L + info->nnz_U + info->nnz_F) + 8 * 2 * (i64)nz;
        if (N->emb && N->emb->dev) {
            long long nl = 0;
            klu_device_times(N->emb->dev, &info->ms_h2d, &info->ms_refactor, &info->ms_solve, &info->ms_kernel, &info->ms_dense, &nl);
            info->launches = nl;
        }
        return B200S_OK;
    }
    const i32 n = N->N.n;
    info->n = n; info->nblocks = N->S.nblocks; info->nnz_A = N->S.nnz;
    info->nnz_L = N->N.Lp[n]; info->nnz_U = N->N.Up[n]; info->nnz_F = N->N.Fp[n];
    info->nlevels = N->P.nlevels; info->max_block = N->S.maxblock;
    info->flops = N->N.flops;
    info->bytes_per_refactor = 8 * (info->nnz_A + info->nnz_L + info->nnz_U + info->nnz_F + 2 * (i64)n);
    if (N->dev) {
        long long nl = 0;
        klu_device_times(N->dev, &info->ms_h2d, &info->ms_refactor, &info->ms_solve, &info->ms_kernel, &info->ms_dense, &nl);
        info->launches = nl;
    }
    return B200S_OK;
}

static b200s_status extract_values(b200s_klu_num* N, b200s_int b, double* Lx, double* Ux, double* Fx, double* Rs) {
    const i32 n = N->N.n;
    if (n == 0) return B200S_OK;
    if (!N->dev) return B200S_NO_DEVICE;
    std::vector<double> slots((size_t)N->P.nslots);
    std::vector<double> rs((size_t)n);
    int st = klu_device_extract(N->dev, b, slots.data(), rs.data());
    if (st != ST_OK) return (b200s_status)st;
    for (i32 k = 0; k < n; k++) {
        const i64 nu = N->N.Up[k + 1] - N->N.Up[k];
        if (Ux) for (i64 p = 0; p < nu; p++) Ux[N->N.Up[k] + p] = slots[N->P.cbeg[k] + p];
        if (Lx) {
            Lx[N->N.Lp[k]] = 1.0;
            for (i64 p = N->N.Lp[k] + 1; p < N->N.Lp[k + 1]; p++) Lx[p] = slots[N->P.lslot0[k] + (p - N->N.Lp[k] - 1)];
        }
        if (Fx) for (i64 p = N->N.Fp[k]; p < N->N.Fp[k + 1]; p++) Fx[p] = slots[N->P.fslot0[k] + (p - N->N.Fp[k])];
    }
    if (Rs) for (i32 k = 0; k < n; k++) Rs[k] = rs[k];
    return B200S_OK;
}

b200s_status b200s_klu_extract(const b200s_klu_num* Nc, b200s_int* Lp, b200s_int* Li, double* Lx, b200s_int* Up, b200s_int* Ui,
                               double* Ux, b200s_int* Fp, b200s_int* Fi, double* Fx, b200s_int* P, b200s_int* Q, double* Rs,
                               b200s_int* R) {
    if (!Nc || Nc->zN) return B200S_INVALID;
    b200s_klu_num* N = const_cast<b200s_klu_num*>(Nc);
    const i32 n = N->N.n;
    for (i32 k = 0; k <= n; k++) {
        if (Lp) Lp[k] = N->N.Lp[k];
        if (Up) Up[k] = N->N.Up[k];
        if (Fp) Fp[k] = N->N.Fp[k];
    }
    if (Li) for (size_t p = 0; p < N->N.Li.size(); p++) Li[p] = N->N.Li[p];
    if (Ui) for (size_t p = 0; p < N->N.Ui.size(); p++) Ui[p] = N->N.Ui[p];
    if (Fi) for (size_t p = 0; p < N->N.Fi.size(); p++) Fi[p] = N->N.Fi[p];
    for (i32 k = 0; k < n; k++) {
        if (P) P[k] = N->N.Pnum[k];
        if (Q) Q[k] = N->S.Q[k];
    }
    if (R) for (i32 b = 0; b <= N->S.nblocks; b++) R[b] = N->S.R[b];
    if (Lx || Ux || Fx || Rs) return extract_values(N, 0, Lx, Ux, Fx, Rs);
    return B200S_OK;
}
b200s_status b200s_klu_extract_batch(b200s_klu_num* N, b200s_int b, double* Lx, double* Ux, double* Fx, double* Rs) {
    if (!N) return B200S_INVALID;
    return extract_values(N, b, Lx, Ux, Fx, Rs);
}

/* Test hook: replays the wave-schedule tables of the refactorization plan on the HOST for one matrix (klu_plan_emulate) and
 * returns the factor in the layout of b200s_klu_extract_host.  It verifies the host-built plan without a GPU; nothing in
 * the Python mirrors calls it and it is not a factorization path. */
b200s_status b200s_klu_plan_emulate_host(const b200s_klu_num* N, const double* val, double* Lx, double* Ux, double* Fx, double* Rs) {
    if (!N || !val) return B200S_INVALID;
    if (b200s_status st0 = ensure_refactor_tables(const_cast<b200s_klu_num*>(N))) return st0;       // (lazily built)
    const i32 n = N->N.n;
    std::vector<double> slots((size_t)std::max<i64>(N->P.nslots, 1)), rs((size_t)std::max<i32>(n, 1));
    int st;
    try { st = klu_plan_emulate(N->S, N->P, val, slots.data(), rs.data()); }
    catch (const std::exception& e) { set_last_error(e.what()); return B200S_INVALID; }
    for (i32 k = 0; k < n; k++) {
        const i64 nu = N->N.Up[k + 1] - N->N.Up[k];
        if (Ux) for (i64 p = 0; p < nu; p++) Ux[N->N.Up[k] + p] = slots[N->P.cbeg[k] + p];
        if (Lx) {
            Lx[N->N.Lp[k]] = 1.0;
            for (i64 p = N->N.Lp[k] + 1; p < N->N.Lp[k + 1]; p++) Lx[p] = slots[N->P.lslot0[k] + (p - N->N.Lp[k] - 1)];
        }
        if (Fx) for (i64 p = N->N.Fp[k]; p < N->N.Fp[k + 1]; p++) Fx[p] = slots[N->P.fslot0[k] + (p - N->N.Fp[k])];
        if (Rs) Rs[k] = rs[k];
    }
    return (b200s_status)st;
}

/* Test hook: the operation tape of the one-matrix solve kernel (k_klu_solve_one) replayed on the HOST with the values of the
 * pivot search, in place on B (n x nrhs, leading dimension ldB).  Verifies the tape without a GPU; not a solve path. */
b200s_status b200s_klu_solve_tape_host(const b200s_klu_num* N, int trans, double* B, b200s_int nrhs, b200s_int ldB) {
    if (!N || nrhs < 0 || (trans != 0 && trans != 1) || N->zN) return B200S_INVALID;
    const i32 n = N->N.n;
    if (n == 0 || nrhs == 0) return B200S_OK;
    if (!B || ldB < n) return B200S_INVALID;
    const KluPlan& P = N->P; const KluNumeric& M = N->N;
    try {
        std::vector<double> slots((size_t)std::max<i64>(P.nslots, 1), 0.0);
        for (i32 k = 0; k < n; k++) {
            for (i64 p = M.Up[k]; p < M.Up[k + 1]; p++) slots[P.cbeg[k] + (p - M.Up[k])] = M.Ux[p];
            for (i64 p = M.Lp[k] + 1; p < M.Lp[k + 1]; p++) slots[P.lslot0[k] + (p - M.Lp[k] - 1)] = M.Lx[p];
            for (i64 p = M.Fp[k]; p < M.Fp[k + 1]; p++) slots[P.fslot0[k] + (p - M.Fp[k])] = M.Fx[p];
        }
        return (b200s_status)klu_solve_tape_host(N->S, M, P, trans, slots.data(), B, nrhs, ldB);
    } catch (const std::bad_alloc&) {
        return B200S_OUT_OF_MEMORY;
    }
}

b200s_status b200s_klu_plan_view(const b200s_klu_num* N, b200s_klu_plan_view_t* v) {
    if (!N || !v) return B200S_INVALID;
    if (b200s_status st0 = ensure_refactor_tables(const_cast<b200s_klu_num*>(N))) return st0;       // (lazily built)
    const KluPlan& P = N->P;
    v->n = P.n; v->nlevels = P.nlevels; v->nslots = P.nslots; v->lu_slots = P.lu_slots; v->nnz_A = P.nnzA;
    v->nupd = (b200s_int)P.upd_uslot.size(); v->ndest = (b200s_int)P.dest.size();
    v->cbeg = P.cbeg.data(); v->udiag_slot = P.udiag_slot.data(); v->slot_src = P.slot_src.data();
    v->slot_row = P.slot_row.data(); v->rowptr = P.rowptr.data(); v->rowent = P.rowent.data();
    v->level_ptr = P.level_ptr.data(); v->level_cols = P.level_cols.data(); v->upd_ptr = P.upd_ptr.data();
    v->upd_uslot = P.upd_uslot.data(); v->upd_lslot = P.upd_lslot.data(); v->upd_cnt = P.upd_cnt.data();
    v->upd_dest = P.upd_dest.data(); v->dest = P.dest.data(); v->lslot0 = P.lslot0.data(); v->fslot0 = P.fslot0.data();
    v->nwaves = (b200s_int)P.wave_col0.size() - 1; v->nwaves_with_deps = 0;
    for (int d : P.wave_hasdep) v->nwaves_with_deps += d;
    v->nbatches = (b200s_int)P.bseg_ptr.size() - 1; v->nsegments = (b200s_int)P.seg_src.size(); v->staged_rows = 0;
    for (int c : P.seg_cnt) v->staged_rows += c;
    v->npieces = (b200s_int)P.pc_j0.size(); v->npiece_users = (b200s_int)P.pc_user_col.size();
    v->wave_ok = P.wave_ok ? 1 : 0; v->nearly = (b200s_int)P.ecols.size(); v->nearly_levels = (b200s_int)P.elevel_ptr.size() - 1;
    return B200S_OK;
}

b200s_status b200s_klu_symbolic_perm(const b200s_klu_sym* S, b200s_int* P, b200s_int* Q, b200s_int* R, b200s_int* nblocks) {
    if (!S) return B200S_INVALID;
    for (i32 k = 0; k < S->S.n; k++) { if (P) P[k] = S->S.P[k]; if (Q) Q[k] = S->S.Q[k]; }
    if (R) for (i32 b = 0; b <= S->S.nblocks; b++) R[b] = S->S.R[b];
    if (nblocks) *nblocks = S->S.nblocks;
    return B200S_OK;
}
void b200s_klu_free_symbolic(b200s_klu_sym* S) { delete S; }
void b200s_klu_free_numeric(b200s_klu_num* N) {
    if (!N) return;
    if (N->dev) klu_device_destroy(N->dev);
    if (N->emb) b200s_klu_free_numeric(N->emb);
    if (N->emb_sym) b200s_klu_free_symbolic(N->emb_sym);
    delete N->zN;
    delete N;
}

/* ---- complex matrices ('z' spmatrix; klu_zl_*: src/C/klu.c:161-162,348-355,468-479,661-668,754-813) -----------------------
 * val: nnz (re, im) pairs.  The threshold-pivoting factorization runs on the host in complex arithmetic (|z| = hypot, row
 * scale = max |z| of the row: KLU's rules) and IS the factor get_numeric / get_det return; the solves run on the device
 * through the real embedding a + ib -> [[a, -b], [b, a]] of order 2n -- a complex vector is its own embedding (interleaved
 * (re, im)), so b200s_klu_solve_z works on the caller's buffer. */
b200s_status b200s_klu_factor_z(b200s_klu_sym* S, const b200s_int* colptr, const b200s_int* rowind, const double* val,
                                b200s_klu_num** out) {
    B200S_NVTX("b200s_klu_factor_z");
    if (!S || !out) return B200S_INVALID;
    *out = nullptr;
    const i32 n = S->S.n;
    if (n > 0) {
        if (!colptr || !val) return B200S_INVALID;
        if (colptr[n] != S->S.nnz) { set_last_error("matrix pattern differs from the analysed pattern"); return B200S_INVALID; }
        for (i32 j = 0; j <= n; j++) if (colptr[j] != S->S.Ap[j]) { set_last_error("matrix pattern differs from the analysed pattern"); return B200S_INVALID; }
        for (i64 p = 0; p < S->S.nnz; p++) if (rowind[p] != S->S.Ai[p]) { set_last_error("matrix pattern differs from the analysed pattern"); return B200S_INVALID; }
    }
    b200s_klu_num* N = new (std::nothrow) b200s_klu_num();
    if (!N) return B200S_OUT_OF_MEMORY;
    N->S = S->S;
    N->device = current_device();
    b200s_status rs = B200S_OK;
    try {
        N->zN = new KluNumericZ();
        int st = klu_factor_z(N->S, reinterpret_cast<const std::complex<double>*>(val), *N->zN);
        if (st != ST_OK) { b200s_klu_free_numeric(N); return (b200s_status)st; }
        N->N.n = 0;           // the real members stay empty: every entry point dispatches on zN
        if (n > 0) {
            // embedded real matrix: column j -> columns 2j, 2j+1; entry (i, a + ib) -> rows 2i, 2i+1
            const i64 nnz = S->S.nnz;
            std::vector<i64> ecp((size_t)2 * n + 1), eri((size_t)4 * nnz);
            std::vector<double> ev((size_t)4 * nnz);
            i64 q = 0;
            ecp[0] = 0;
            for (i32 j = 0; j < n; j++) {
                for (int half = 0; half < 2; half++) {
                    for (i64 p = colptr[j]; p < colptr[j + 1]; p++) {
                        const double a = val[2 * p], b = val[2 * p + 1];
                        eri[q] = 2 * rowind[p];     ev[q++] = half == 0 ? a : -b;
                        eri[q] = 2 * rowind[p] + 1; ev[q++] = half == 0 ? b : a;
                    }
                    ecp[2 * j + half + 1] = q;
                }
            }
            rs = b200s_klu_analyze(2 * (i64)n, ecp.data(), eri.data(), &N->emb_sym);
            if (rs == B200S_OK) rs = b200s_klu_factor(N->emb_sym, ecp.data(), eri.data(), ev.data(), &N->emb);
        }
    } catch (const std::bad_alloc&) {
        b200s_klu_free_numeric(N); return B200S_OUT_OF_MEMORY;
    } catch (const std::exception& e) {
        set_last_error(e.what()); b200s_klu_free_numeric(N); return B200S_INVALID;
    }
    if (rs != B200S_OK) { b200s_klu_free_numeric(N); return rs; }
    *out = N;
    return B200S_OK;
}

/* trans: 0 = A x = b, 1 = A^T x = b, 2 = A^H x = b (klu.c:651-668: klu_zl_solve / klu_zl_tsolve with conj_solve).
 * B: nrhs columns of (re, im) pairs, leading dimension ldB (in complex numbers), overwritten by the solution. */
b200s_status b200s_klu_solve_z(b200s_klu_num* N, int trans, double* B, b200s_int nrhs, b200s_int ldB) {
    B200S_NVTX("b200s_klu_solve_z");
    if (!N || !N->zN || trans < 0 || trans > 2 || nrhs < 0) return B200S_INVALID;
    const i64 n = N->zN->n;
    if (n == 0 || nrhs == 0) return B200S_OK;
    if (!B || ldB < n || !N->emb) return B200S_INVALID;
    // emb(A)^T = emb(A^H): the real transposed solve is the conjugate-transposed complex one; A^T x = b is
    // conj(A^H conj(x)) = b, i.e. conjugate the right-hand side, solve with A^H, conjugate the result
    auto conj_cols = [&]() {
        for (i64 c = 0; c < nrhs; c++)
            for (i64 i = 0; i < n; i++) B[2 * (c * ldB + i) + 1] = -B[2 * (c * ldB + i) + 1];
    };
    if (trans == 1) conj_cols();
    b200s_status st = b200s_klu_solve(N->emb, trans == 0 ? 0 : 1, B, nrhs, 2 * ldB);
    if (trans == 1) conj_cols();
    return st;
}

/* the factor object of the real embedding (order 2n; column j -> columns 2j, 2j+1; per complex entry (i, a + ib) the
 * embedded CCS holds, in column 2j: (2i, a), (2i+1, b), in column 2j+1: (2i, -b), (2i+1, a), entries in the caller's order):
 * the batched entry points (b200s_klu_refactor_batch*, b200s_klu_solve_batch*) take it with embedded values / interleaved
 * (re, im) right-hand sides.  Owned by N. */
b200s_klu_num* b200s_klu_embedded(b200s_klu_num* N) { return N ? N->emb : nullptr; }

/* klu_zl_extract (klu.c:468-479): the complex factors; Lx, Ux, Fx receive (re, im) pairs */
b200s_status b200s_klu_extract_z(const b200s_klu_num* N, b200s_int* Lp, b200s_int* Li, double* Lx, b200s_int* Up, b200s_int* Ui,
                                 double* Ux, b200s_int* Fp, b200s_int* Fi, double* Fx, b200s_int* P, b200s_int* Q, double* Rs,
                                 b200s_int* R) {
    if (!N || !N->zN) return B200S_INVALID;
    const KluNumericZ& Z = *N->zN;
    const i32 n = Z.n;
    for (i32 k = 0; k <= n; k++) {
        if (Lp) Lp[k] = Z.Lp[k];
        if (Up) Up[k] = Z.Up[k];
        if (Fp) Fp[k] = Z.Fp[k];
    }
    if (Li) for (size_t p = 0; p < Z.Li.size(); p++) Li[p] = Z.Li[p];
    if (Ui) for (size_t p = 0; p < Z.Ui.size(); p++) Ui[p] = Z.Ui[p];
    if (Fi) for (size_t p = 0; p < Z.Fi.size(); p++) Fi[p] = Z.Fi[p];
    if (Lx) for (size_t p = 0; p < Z.Lx.size(); p++) { Lx[2 * p] = Z.Lx[p].real(); Lx[2 * p + 1] = Z.Lx[p].imag(); }
    if (Ux) for (size_t p = 0; p < Z.Ux.size(); p++) { Ux[2 * p] = Z.Ux[p].real(); Ux[2 * p + 1] = Z.Ux[p].imag(); }
    if (Fx) for (size_t p = 0; p < Z.Fx.size(); p++) { Fx[2 * p] = Z.Fx[p].real(); Fx[2 * p + 1] = Z.Fx[p].imag(); }
    for (i32 k = 0; k < n; k++) {
        if (P) P[k] = Z.Pnum[k];
        if (Q) Q[k] = N->S.Q[k];
        if (Rs) Rs[k] = Z.Rs[k];
    }
    if (R) for (i32 b = 0; b <= N->S.nblocks; b++) R[b] = N->S.R[b];
    return B200S_OK;
}

}  // extern "C"
