// extern "C" boundary of libb200sparse.so (include/b200sparse.h): plain pointers and sizes, no C++ or
// torch types.  Cholesky half; the KLU half lives in klu_capi.cu.
#include "../../include/b200sparse.h"
#include "gpu.hpp"
#include "nvtx_range.hpp"
#include <cstdlib>
#include <cstring>
#include <new>
#include <stdexcept>
#include <algorithm>
#include <memory>
#include <mutex>
#include <vector>

using namespace b200s;

struct b200s_chol {
    // The symbolic plan is immutable after analyze and may be shared by several factor objects: kvxopt re-analyses the same
    // pattern over and over (cholmod.linsolve frees its factor at the end of every call, cholmod.c:750; misc.kkt_chol2 calls
    // cholmod.symbolic(K) in every interior-point iteration, misc.py:1486), so analyze keeps the last few small plans keyed
    // by (n, uplo, options, pattern, user permutation) and hands them out again after comparing the pattern itself.
    std::shared_ptr<const CholPlan> plan_sp;
    const CholPlan& plan() const { return *plan_sp; }
    CholOpts opts;
    CholDevice* dev = nullptr;
    int device = 0;
    CholTimes times;
    i64 minor = 0;
    bool numeric = false, profiling = false, ldl = false, custom_owned = false;
    char uplo = 'L';
    std::vector<i64> Ap, Ai;      // the pattern given to analyze: numeric() compares the caller's pattern with it (cholmod.c:322-398
                                  // rebuilds the cholmod_sparse from A's own colptr/rowind on every call)
    std::vector<double> remap;    // scratch of the slow path (pattern differs: values re-mapped through A's own indices)
    // complex Hermitian matrices (b200s_chol_*_z): factored through the real symmetric embedding a + ib -> [[a, -b], [b, a]] of
    // order 2 zn, rows/columns (2i, 2i+1) per complex index i
    i64 zn = 0;                   // complex order (0: real factor object)
    std::vector<i64> zAp, zAi;    // the complex pattern given to analyze_z
    std::vector<i64> zsrc;        // per entry of the embedded pattern: complex entry it comes from (-1: structural zero)
    std::vector<signed char> zcode;   // 0: real part, 1: imaginary part, 2: minus the imaginary part
    std::vector<double> zval;     // embedded values (scratch of factorize_z)
};

static_assert(int(B200S_OK) == ST_OK && int(B200S_NOT_POSDEF) == ST_NOT_POSDEF && int(B200S_SINGULAR) == ST_SINGULAR &&
              int(B200S_OUT_OF_MEMORY) == ST_OOM && int(B200S_TOO_LARGE) == ST_TOO_LARGE && int(B200S_INVALID) == ST_INVALID &&
              int(B200S_NO_DEVICE) == ST_NO_DEVICE && int(B200S_CUDA_ERROR) == ST_CUDA, "status codes out of sync");

// ---- plan cache (see b200s_chol::plan_sp) -------------------------------------------------------------------------------
namespace {
struct PlanKey {
    i64 n = 0, nnz = 0;
    char uplo = 'L';
    CholOpts opts;
    bool has_perm = false;
    unsigned long long hash = 0;
    std::vector<i64> Ap, Ai, perm;           // the pattern itself: a hit is confirmed by comparing it, not by the hash
    std::shared_ptr<const CholPlan> plan;
    unsigned long long stamp = 0;
    // device objects (uploaded plan tables, schedules, L and W storage, stream) of factor objects that were freed: the next
    // factor object of this plan on the same GPU takes one over instead of uploading everything again
    std::vector<std::pair<int, CholDevice*>> idle;
    ~PlanKey() { for (auto& d : idle) chol_device_destroy(d.second); }
    PlanKey() = default;
    PlanKey(PlanKey&&) = default;
    PlanKey& operator=(PlanKey&& o) {
        if (this != &o) {
            for (auto& d : idle) chol_device_destroy(d.second);
            idle.clear();
            n = o.n; nnz = o.nnz; uplo = o.uplo; opts = o.opts; has_perm = o.has_perm; hash = o.hash;
            Ap = std::move(o.Ap); Ai = std::move(o.Ai); perm = std::move(o.perm); plan = std::move(o.plan); stamp = o.stamp;
            idle = std::move(o.idle); o.idle.clear();
        }
        return *this;
    }
};
constexpr size_t PLAN_CACHE_IDLE_DEVS = 2;
constexpr i64 PLAN_CACHE_IDLE_BYTES = 256ll << 20;
std::mutex g_plan_mu;
std::vector<PlanKey>& g_plans = *new std::vector<PlanKey>();      // never destroyed: no CUDA calls during static destruction at exit
unsigned long long g_plan_stamp = 0;
constexpr size_t PLAN_CACHE_ENTRIES = 8;
constexpr i64 PLAN_CACHE_MAX_NNZ = 4 << 20;  // only patterns whose copy is cheap to keep (the 100^3 case re-analyses)

unsigned long long fnv(const void* p, size_t bytes, unsigned long long h) {
    const unsigned long long* w = (const unsigned long long*)p;
    for (size_t i = 0; i < bytes / 8; i++) { h ^= w[i]; h *= 1099511628211ull; }
    return h;
}
bool same_opts(const CholOpts& a, const CholOpts& b) {
    return a.supernodal == b.supernodal && a.nmethods == b.nmethods && a.postorder == b.postorder && a.ordering == b.ordering &&
           a.dbound == b.dbound && a.block == b.block && a.max_merge_cols == b.max_merge_cols && !memcmp(a.nrelax, b.nrelax, sizeof a.nrelax) && !memcmp(a.zrelax, b.zrelax, sizeof a.zrelax);
}
bool plan_cache_enabled() { static const bool on = getenv("B200S_NO_PLAN_CACHE") == nullptr; return on; }
unsigned long long pattern_hash(i64 n, const b200s_int* cp, const b200s_int* ri, const b200s_int* perm) {
    unsigned long long h = 1469598103934665603ull;
    h = fnv(cp, sizeof(i64) * (size_t)(n + 1), h);
    h = fnv(ri, sizeof(i64) * (size_t)cp[n], h);
    if (perm) h = fnv(perm, sizeof(i64) * (size_t)n, h);
    return h;
}
std::shared_ptr<const CholPlan> plan_cache_lookup(i64 n, const b200s_int* cp, const b200s_int* ri, char uplo, const b200s_int* perm,
                                                  const CholOpts& o) {
    if (!plan_cache_enabled() || n <= 0 || cp[n] > PLAN_CACHE_MAX_NNZ) return nullptr;
    const unsigned long long h = pattern_hash(n, cp, ri, perm);
    std::lock_guard<std::mutex> g(g_plan_mu);
    for (PlanKey& k : g_plans)
        if (k.hash == h && k.n == n && k.nnz == cp[n] && k.uplo == uplo && k.has_perm == (perm != nullptr) && same_opts(k.opts, o) &&
            !memcmp(k.Ap.data(), cp, sizeof(i64) * (size_t)(n + 1)) && !memcmp(k.Ai.data(), ri, sizeof(i64) * (size_t)cp[n]) &&
            (!perm || !memcmp(k.perm.data(), perm, sizeof(i64) * (size_t)n))) {
            k.stamp = ++g_plan_stamp;
            return k.plan;
        }
    return nullptr;
}
void plan_cache_store(i64 n, const b200s_int* cp, const b200s_int* ri, char uplo, const b200s_int* perm, const CholOpts& o,
                      const std::shared_ptr<const CholPlan>& plan) {
    if (!plan_cache_enabled() || n <= 0 || cp[n] > PLAN_CACHE_MAX_NNZ) return;
    PlanKey k;
    k.n = n; k.nnz = cp[n]; k.uplo = uplo; k.opts = o; k.has_perm = perm != nullptr; k.hash = pattern_hash(n, cp, ri, perm);
    k.Ap.assign(cp, cp + n + 1); k.Ai.assign(ri, ri + cp[n]);
    if (perm) k.perm.assign(perm, perm + n);
    k.plan = plan;
    std::lock_guard<std::mutex> g(g_plan_mu);
    k.stamp = ++g_plan_stamp;
    if (g_plans.size() < PLAN_CACHE_ENTRIES) { g_plans.push_back(std::move(k)); return; }
    size_t old = 0;
    for (size_t i = 1; i < g_plans.size(); i++) if (g_plans[i].stamp < g_plans[old].stamp) old = i;
    g_plans[old] = std::move(k);
}
CholDevice* plan_cache_take_device(const CholPlan* plan, int device) {
    std::lock_guard<std::mutex> g(g_plan_mu);
    for (PlanKey& k : g_plans)
        if (k.plan.get() == plan)
            for (size_t i = 0; i < k.idle.size(); i++)
                if (k.idle[i].first == device) { CholDevice* d = k.idle[i].second; k.idle.erase(k.idle.begin() + i); return d; }
    return nullptr;
}
bool plan_cache_park_device(const CholPlan* plan, int device, CholDevice* d) {
    if ((plan->lsize + plan->wsize) * 8 > PLAN_CACHE_IDLE_BYTES) return false;
    std::lock_guard<std::mutex> g(g_plan_mu);
    for (PlanKey& k : g_plans)
        if (k.plan.get() == plan && k.idle.size() < PLAN_CACHE_IDLE_DEVS) { k.idle.emplace_back(device, d); return true; }
    return false;
}
}  // namespace

extern "C" {

const char* b200s_strerror(b200s_status s) {
    switch (s) {
        case B200S_OK: return "ok";
        case B200S_NOT_POSDEF: return "matrix is not positive definite";
        case B200S_SINGULAR: return "singular matrix";
        case B200S_OUT_OF_MEMORY: return "out of memory";
        case B200S_TOO_LARGE: return "problem too large";
        case B200S_INVALID: return "invalid input";
        case B200S_NO_DEVICE: return "no CUDA device (the numeric phase has no CPU fallback)";
        case B200S_CUDA_ERROR: return "CUDA error";
    }
    return "unknown status";
}
const char* b200s_last_error(void) { return get_last_error(); }
const char* b200s_version(void) { return "b200sparse 0.1 (sm_100a)"; }
int b200s_device_count(void) { return device_count(); }
b200s_status b200s_set_device(int dev) {
    if (dev < 0 || dev >= device_count()) return B200S_INVALID;
    set_current_device(dev);
    return B200S_OK;
}

void b200s_chol_default_opts(b200s_chol_opts* o) {
    if (!o) return;
    CholOpts d;
    o->supernodal = d.supernodal; o->nmethods = d.nmethods; o->postorder = d.postorder; o->dbound = d.dbound;
    o->ordering = d.ordering; o->block = d.block; o->max_merge_cols = d.max_merge_cols;
    for (int i = 0; i < 3; i++) { o->nrelax[i] = d.nrelax[i]; o->zrelax[i] = d.zrelax[i]; }
}

b200s_status b200s_chol_analyze(b200s_int n, const b200s_int* colptr, const b200s_int* rowind, char uplo,
                                const b200s_int* perm, const b200s_chol_opts* opts, b200s_chol** out) {
    B200S_NVTX("b200s_chol_analyze");
    if (!out) return B200S_INVALID;
    *out = nullptr;
    if (n < 0 || (n > 0 && (!colptr || (colptr[n] > 0 && !rowind)))) return B200S_INVALID;
    b200s_chol* F = new (std::nothrow) b200s_chol();
    if (!F) return B200S_OUT_OF_MEMORY;
    if (opts) {
        F->opts.supernodal = opts->supernodal; F->opts.nmethods = opts->nmethods; F->opts.postorder = opts->postorder;
        F->opts.dbound = opts->dbound; F->opts.ordering = opts->ordering;
        if (opts->block > 0) F->opts.block = opts->block;
        F->opts.max_merge_cols = opts->max_merge_cols > 0 ? opts->max_merge_cols : 0;
        for (int i = 0; i < 3; i++) { F->opts.nrelax[i] = opts->nrelax[i]; F->opts.zrelax[i] = opts->zrelax[i]; }
    }
    // supernodal = 0 asks CHOLMOD for a simplicial LDL' factorization (no pivoting; any symmetric matrix with nonzero
    // pivots, e.g. quasi-definite KKT systems).  The engine always factors supernodally; in this mode the kernels compute
    // the signed square-root form P A P' = Lt S Lt', S = diag(+-1), so L_ldl = Lt diag(Lt)^-1 and D = S diag(Lt)^2, and the
    // factor object answers with LDL' semantics (sys 1..6 scaled by the diagonal, getfactor with D on the diagonal,
    // diag() refused) -- see chol_gpu.cu.
    if (F->opts.supernodal < 0 || F->opts.supernodal > 2) {
        set_last_error("cholmod.options['supernodal'] must be 0, 1 or 2");
        delete F;
        return B200S_INVALID;
    }
    F->ldl = F->opts.supernodal == 0;
    F->device = current_device();
    try {
        static const b200s_int zero = 0;
        std::shared_ptr<const CholPlan> cached = plan_cache_lookup(n, colptr, rowind, uplo, perm, F->opts);
        if (cached) F->plan_sp = cached;
        else {
            auto fresh = std::make_shared<CholPlan>();
            chol_analyze(n, n > 0 ? colptr : &zero, rowind, uplo, perm, F->opts, *fresh);
            F->plan_sp = fresh;
            plan_cache_store(n, colptr, rowind, uplo, perm, F->opts, F->plan_sp);
        }
    } catch (const std::bad_alloc&) {
        delete F;
        return B200S_OUT_OF_MEMORY;
    } catch (const std::invalid_argument& e) {
        set_last_error(e.what());
        delete F;
        return B200S_INVALID;
    } catch (const std::exception& e) {
        set_last_error(e.what());
        delete F;
        return B200S_INVALID;
    }
    F->minor = n;
    F->uplo = uplo;
    try {
        if (n > 0) { F->Ap.assign(colptr, colptr + n + 1); F->Ai.assign(rowind, rowind + colptr[n]); }
    } catch (const std::bad_alloc&) { delete F; return B200S_OUT_OF_MEMORY; }
    *out = F;
    return B200S_OK;
}

static b200s_status ensure_device(b200s_chol* F) {
    if (!F->dev) {
        int st = ST_OK;
        F->dev = plan_cache_take_device(F->plan_sp.get(), F->device);
        if (!F->dev) F->dev = chol_device_create(F->plan(), F->opts, F->device, &st);
        if (!F->dev) return (b200s_status)st;
        chol_device_set_profiling(F->dev, F->profiling);
        chol_device_set_ldl(F->dev, F->ldl);
    }
    return B200S_OK;
}

static b200s_status factorize_impl(b200s_chol* F, const double* val, bool on_device, b200s_int* minor_out) {
    B200S_NVTX("factorize_impl");
    if (!F) return B200S_INVALID;
    F->numeric = false;
    if (F->plan().n == 0) { F->numeric = true; if (minor_out) *minor_out = 0; return B200S_OK; }
    if (!val && F->plan().nnzA > 0) return B200S_INVALID;
    { b200s_status es = ensure_device(F); if (es != B200S_OK) return es; }
    i64 minor = F->plan().n;
    int st = chol_device_factorize(F->dev, val, on_device, &minor, &F->times);
    F->minor = minor;
    if (minor_out) *minor_out = minor;
    F->numeric = (st == ST_OK);
    return (b200s_status)st;
}
// Values of a matrix whose pattern differs from the analysed one, re-mapped onto the analysed entries: entries of the
// analysed pattern that A does not have become zero (CHOLMOD accepts a subset pattern); an entry of A inside the
// referenced triangle that the analysis has not seen is an error.  Row indices must be sorted within each column.
static b200s_status remap_values(b200s_chol* F, const b200s_int* colptr, const b200s_int* rowind, const double* val) {
    const i64 n = F->plan().n;
    try { F->remap.assign((size_t)F->plan().nnzA, 0.0); } catch (const std::bad_alloc&) { return B200S_OUT_OF_MEMORY; }
    if (colptr[0] != 0) return B200S_INVALID;
    for (i64 j = 0; j < n; j++) {
        i64 q = F->Ap[j];
        const i64 qe = F->Ap[j + 1];
        i64 last = -1;
        if (colptr[j + 1] < colptr[j]) return B200S_INVALID;
        for (i64 k = colptr[j]; k < colptr[j + 1]; k++) {
            const i64 r = rowind[k];
            if (r < 0 || r >= n || r <= last) { set_last_error("numeric: row indices of A must be sorted and in range"); return B200S_INVALID; }
            last = r;
            const bool referenced = F->uplo == 'L' ? r >= j : r <= j;
            while (q < qe && F->Ai[q] < r) q++;
            if (q < qe && F->Ai[q] == r) F->remap[(size_t)q] = val[k];
            else if (referenced) { set_last_error("numeric: A has an entry outside the pattern analysed by symbolic()"); return B200S_INVALID; }
        }
    }
    return B200S_OK;
}
b200s_status b200s_chol_factorize(b200s_chol* F, const b200s_int* colptr, const b200s_int* rowind, const double* val,
                                  b200s_int* minor_out) {
    if (F && colptr && F->plan().n > 0) {
        const i64 n = F->plan().n, nnz = F->plan().nnzA;
        const bool same = colptr[n] == nnz && memcmp(colptr, F->Ap.data(), sizeof(i64) * (size_t)(n + 1)) == 0 &&
                          (nnz == 0 || (rowind && memcmp(rowind, F->Ai.data(), sizeof(i64) * (size_t)nnz) == 0));
        if (!same) {
            if (!rowind && colptr[n] > 0) return B200S_INVALID;
            if (!val && colptr[n] > 0) return B200S_INVALID;
            b200s_status st = remap_values(F, colptr, rowind, val);
            if (st != B200S_OK) return st;
            return factorize_impl(F, F->remap.data(), false, minor_out);
        }
    }
    return factorize_impl(F, val, false, minor_out);
}
b200s_status b200s_chol_factorize_dev(b200s_chol* F, const double* val_dev, b200s_int* minor_out) {
    return factorize_impl(F, val_dev, true, minor_out);
}

// ---- level-stepped factorization and front ownership (multi-GPU subtree-to-subcube building blocks) ----------
b200s_status b200s_chol_set_owned(b200s_chol* F, const unsigned char* owned) {
    if (!F) return B200S_INVALID;
    if (F->plan().n == 0) return B200S_OK;
    if (F->ldl && owned) {   // the pivot signs of fronts factored elsewhere are not exchanged: one GPU per factor in LDL' mode
        set_last_error("front ownership (subtree-to-subcube) is not available with supernodal = 0 (LDL')");
        return B200S_INVALID;
    }
    b200s_status es = ensure_device(F);
    if (es != B200S_OK) return es;
    F->custom_owned = true;
    return (b200s_status)chol_device_set_owned(F->dev, owned);
}
b200s_status b200s_chol_factor_begin(b200s_chol* F, const double* val, int val_on_device) {
    B200S_NVTX("b200s_chol_factor_begin");
    if (!F) return B200S_INVALID;
    F->numeric = false;
    if (F->plan().n == 0) return B200S_OK;
    if (!val && F->plan().nnzA > 0) return B200S_INVALID;
    b200s_status es = ensure_device(F);
    if (es != B200S_OK) return es;
    return (b200s_status)chol_device_factor_begin(F->dev, val, val_on_device != 0);
}
b200s_status b200s_chol_factor_level(b200s_chol* F, b200s_int level) {
    B200S_NVTX("b200s_chol_factor_level");
    if (!F || !F->dev) return B200S_INVALID;
    return (b200s_status)chol_device_factor_level(F->dev, (int)level);
}
b200s_status b200s_chol_factor_level_phase(b200s_chol* F, b200s_int level, int phase) {
    B200S_NVTX("b200s_chol_factor_level_phase");
    if (!F || !F->dev || phase < 1 || phase > 3) return B200S_INVALID;
    return (b200s_status)chol_device_factor_level_phase(F->dev, (int)level, phase);
}
b200s_status b200s_chol_set_syrk_split(b200s_chol* F, const unsigned char* own, const int* tile_lo, const int* tile_hi,
                                       const long long* base, double* scratch_dev) {
    if (!F) return B200S_INVALID;
    if (F->plan().n == 0) return B200S_OK;
    if (F->ldl && own) { set_last_error("shared Schur complements are not available with supernodal = 0 (LDL')"); return B200S_INVALID; }
    b200s_status es = ensure_device(F);
    if (es != B200S_OK) return es;
    return (b200s_status)chol_device_set_syrk_split(F->dev, own, tile_lo, tile_hi, base, scratch_dev);
}
b200s_status b200s_chol_factor_end(b200s_chol* F, b200s_int* minor_out) {
    B200S_NVTX("b200s_chol_factor_end");
    if (!F) return B200S_INVALID;
    if (F->plan().n == 0) { F->numeric = true; if (minor_out) *minor_out = 0; return B200S_OK; }
    if (!F->dev) return B200S_INVALID;
    i64 minor = F->plan().n;
    int st = chol_device_factor_end(F->dev, &minor, &F->times);
    F->minor = minor;
    if (minor_out) *minor_out = minor;
    F->numeric = (st == ST_OK);
    return (b200s_status)st;
}
b200s_status b200s_chol_sync(b200s_chol* F) {
    if (!F) return B200S_INVALID;
    if (!F->dev) return B200S_OK;
    return (b200s_status)chol_device_sync(F->dev);
}
b200s_status b200s_chol_front_layout(const b200s_chol* F, b200s_int* parent, b200s_int* level, b200s_int* ncols, b200s_int* nrows,
                                     b200s_int* loff, b200s_int* lsize, b200s_int* uoff, b200s_int* usize) {
    if (!F) return B200S_INVALID;
    const CholPlan& P = F->plan();
    for (size_t s = 0; s < P.fronts.size(); s++) {
        const Front& f = P.fronts[s];
        const i64 m = f.nr - f.nc, mu = m + (f.nc & 1), ldu = (mu + 1) & ~(i64)1;
        if (parent) parent[s] = f.parent;
        if (level) level[s] = f.level;
        if (ncols) ncols[s] = f.nc;
        if (nrows) nrows[s] = f.nr;
        if (loff) loff[s] = f.loff;
        if (lsize) lsize[s] = (i64)f.ld * f.nc;
        if (uoff) uoff[s] = f.uoff;
        if (usize) usize[s] = m > 0 ? ldu * mu : 0;
    }
    return B200S_OK;
}
/* distributed solves: one right-hand side, level by level over the fronts marked by b200s_chol_set_owned */
b200s_status b200s_chol_solve_dist_begin(b200s_chol* F, const double* b_dev) {
    if (!F || !F->dev || !b_dev) return B200S_INVALID;
    if (!F->numeric) { set_last_error("called with symbolic factor"); return B200S_INVALID; }
    return (b200s_status)chol_device_solve_dist_begin(F->dev, b_dev);
}
b200s_status b200s_chol_solve_dist_level(b200s_chol* F, int backward, b200s_int level) {
    if (!F || !F->dev) return B200S_INVALID;
    return (b200s_status)chol_device_solve_dist_level(F->dev, backward, (int)level);
}
b200s_status b200s_chol_solve_dist_end(b200s_chol* F, double* x_dev) {
    if (!F || !F->dev || !x_dev) return B200S_INVALID;
    return (b200s_status)chol_device_solve_dist_end(F->dev, x_dev);
}
b200s_status b200s_chol_solve_buffers(b200s_chol* F, double** T_dev, double** X_dev) {
    if (!F || !T_dev || !X_dev) return B200S_INVALID;
    b200s_status es = ensure_device(F);
    if (es != B200S_OK) return es;
    return (b200s_status)chol_device_solve_buffers(F->dev, T_dev, X_dev);
}
b200s_status b200s_chol_front_layout2(const b200s_chol* F, b200s_int* rowptr, b200s_int* col0) {
    if (!F) return B200S_INVALID;
    const CholPlan& P = F->plan();
    for (size_t s = 0; s < P.fronts.size(); s++) {
        if (rowptr) rowptr[s] = P.fronts[s].rowptr;
        if (col0) col0[s] = P.fronts[s].col0;
    }
    return B200S_OK;
}
b200s_status b200s_chol_device_buffers(b200s_chol* F, double** L_dev, double** W_dev) {
    if (!F || !L_dev || !W_dev) return B200S_INVALID;
    b200s_status es = ensure_device(F);
    if (es != B200S_OK) return es;
    chol_device_buffers(F->dev, L_dev, W_dev);
    return B200S_OK;
}
b200s_status b200s_chol_set_numeric(b200s_chol* F, int numeric, b200s_int minor) {
    if (!F || !F->dev) return B200S_INVALID;
    F->numeric = numeric != 0;
    F->minor = minor;
    chol_device_mark_numeric(F->dev, F->numeric);
    return B200S_OK;
}

static b200s_status solve_impl(b200s_chol* F, int sys, double* B, b200s_int nrhs, b200s_int ldB, bool on_device) {
    B200S_NVTX("solve_impl");
    if (!F || sys < 0 || sys > 8 || nrhs < 0) return B200S_INVALID;
    if (F->plan().n == 0 || nrhs == 0) return B200S_OK;
    if (!B || ldB < F->plan().n) return B200S_INVALID;
    if (!F->numeric || !F->dev) { set_last_error("called with symbolic factor"); return B200S_INVALID; }
    return (b200s_status)chol_device_solve(F->dev, sys, B, nrhs, ldB, on_device, &F->times);
}
b200s_status b200s_chol_solve(b200s_chol* F, int sys, double* B, b200s_int nrhs, b200s_int ldB) {
    return solve_impl(F, sys, B, nrhs, ldB, false);
}
b200s_status b200s_chol_solve_dev(b200s_chol* F, int sys, double* B_dev, b200s_int nrhs, b200s_int ldB) {
    return solve_impl(F, sys, B_dev, nrhs, ldB, true);
}

b200s_status b200s_chol_spsolve(b200s_chol* F, int sys, b200s_int nrows, b200s_int ncols, const b200s_int* Bp,
                                const b200s_int* Bi, const double* Bx, b200s_int** Xp, b200s_int** Xi, double** Xx) {
    B200S_NVTX("b200s_chol_spsolve");
    if (!F || !Xp || !Xi || !Xx || nrows != F->plan().n || ncols < 0) return B200S_INVALID;
    *Xp = nullptr; *Xi = nullptr; *Xx = nullptr;
    const i64 n = nrows;
    const CholPlan& P = F->plan();
    for (i64 j = 0; j < ncols; j++) {
        if (Bp[j + 1] < Bp[j]) return B200S_INVALID;
        for (i64 k = Bp[j]; k < Bp[j + 1]; k++)
            if (Bi[k] < 0 || Bi[k] >= n || (k > Bp[j] && Bi[k] <= Bi[k - 1])) { set_last_error("spsolve: row indices of B must be sorted and in range"); return B200S_INVALID; }
    }
    std::vector<i64> vp, vi;
    std::vector<double> vx;
    try {
        if (n > 0 && ncols > 0 && (!F->numeric || !F->dev)) { set_last_error("called with symbolic factor"); return B200S_INVALID; }
        if (sys == 7 || sys == 8 || (sys == 6 && !F->ldl)) {
            // x = P b, x = P' b (and D = I for LL'): a permutation of the stored entries, done where they are
            vp.assign((size_t)ncols + 1, 0);
            std::vector<std::pair<i64, double>> col;
            for (i64 j = 0; j < ncols; j++) {
                col.clear();
                for (i64 k = Bp[j]; k < Bp[j + 1]; k++) {
                    if (Bx[k] == 0.0) continue;                 // the result holds the numerically nonzero entries
                    const i64 r = sys == 7 ? P.iperm[Bi[k]] : (sys == 8 ? P.perm[Bi[k]] : Bi[k]);
                    col.emplace_back(r, Bx[k]);
                }
                std::sort(col.begin(), col.end());
                for (auto& e : col) { vi.push_back(e.first); vx.push_back(e.second); }
                vp[j + 1] = (i64)vi.size();
            }
        } else {
            // structure-aware device solve: sparse upload, forward sweep restricted to the elimination-tree reach of the
            // nonzero rows, numerically nonzero entries compacted on the device (chol_gpu.cu, CholDevice::spsolve)
            int st = chol_device_spsolve(F->dev, sys, ncols, Bp, Bi, Bx, vp, vi, vx, &F->times);
            if (st != ST_OK) return (b200s_status)st;
        }
    } catch (const std::bad_alloc&) { return B200S_OUT_OF_MEMORY; }
    const i64 nnz = (i64)vi.size();
    b200s_int* xp = (b200s_int*)malloc(sizeof(b200s_int) * (size_t)(ncols + 1));
    b200s_int* xi = (b200s_int*)malloc(sizeof(b200s_int) * (size_t)std::max<i64>(nnz, 1));
    double* xx = (double*)malloc(sizeof(double) * (size_t)std::max<i64>(nnz, 1));
    if (!xp || !xi || !xx) { free(xp); free(xi); free(xx); return B200S_OUT_OF_MEMORY; }
    for (i64 j = 0; j <= ncols; j++) xp[j] = vp[j];
    for (i64 k = 0; k < nnz; k++) { xi[k] = vi[k]; xx[k] = vx[k]; }
    *Xp = xp; *Xi = xi; *Xx = xx;
    return B200S_OK;
}

// ---- complex Hermitian matrices through the real symmetric embedding ---------------------------------------------------
// z = a + ib  ->  [[a, -b], [b, a]]: a *-homomorphism, so the Cholesky factor of the embedded matrix (pairs (2i, 2i+1) kept
// adjacent and in this order by the elimination ordering) IS the embedding of the complex Cholesky factor: L_r = emb(L_c),
// with a real positive diagonal.  Complex vectors embed as interleaved (re, im) pairs -- exactly the memory layout of a
// kvxopt 'z' matrix -- so solves run on the caller's buffer as 2n real unknowns.
b200s_status b200s_chol_analyze_z(b200s_int n, const b200s_int* colptr, const b200s_int* rowind, char uplo, const b200s_int* perm,
                                  const b200s_chol_opts* opts, b200s_chol** out) {
    B200S_NVTX("b200s_chol_analyze_z");
    if (!out) return B200S_INVALID;
    *out = nullptr;
    if (n < 0 || (n > 0 && (!colptr || (colptr[n] > 0 && !rowind))) || (uplo != 'L' && uplo != 'U') || n > 0x3ffffff0) return B200S_INVALID;
    try {
        // ordering of the complex pattern (user permutation, or AMD as for real matrices), expanded to pairs
        std::vector<i64> p1((size_t)n), p2((size_t)2 * n);
        if (perm) {
            std::vector<char> seen((size_t)n, 0);
            for (i64 k = 0; k < n; k++) {
                if (perm[k] < 0 || perm[k] >= n || seen[perm[k]]) { set_last_error("invalid permutation"); return B200S_INVALID; }
                seen[perm[k]] = 1; p1[k] = perm[k];
            }
        } else if (opts && opts->ordering == 1) {
            for (i64 k = 0; k < n; k++) p1[k] = k;
        } else {
            std::vector<i32> a = amd_order(sym_pattern_from_triangle(n, colptr, rowind, uplo));
            for (i64 k = 0; k < n; k++) p1[k] = a[k];
        }
        for (i64 k = 0; k < n; k++) { p2[2 * k] = 2 * p1[k]; p2[2 * k + 1] = 2 * p1[k] + 1; }
        // embedded pattern of the referenced triangle; the diagonal pair always has its three slots (the off-diagonal one of
        // the pair is a structural zero: it makes 2j the only child of 2j+1 in the elimination tree, so any postorder keeps
        // the pair adjacent and the two columns fall into one supernode)
        std::vector<i64> ecp((size_t)2 * n + 1, 0), eri, zsrc;
        std::vector<signed char> zc;
        auto push = [&](i64 r, i64 src, int code) { eri.push_back(r); zsrc.push_back(src); zc.push_back((signed char)code); };
        for (i64 j = 0; j < n; j++) {
            i64 dsrc = -1;
            for (i64 k = colptr[j]; k < colptr[j + 1]; k++) {
                if (rowind[k] < 0 || rowind[k] >= n || (k > colptr[j] && rowind[k] <= rowind[k - 1])) { set_last_error("row indices must be sorted and in range"); return B200S_INVALID; }
                if (rowind[k] == j) dsrc = k;
            }
            for (int half = 0; half < 2; half++) {          // embedded columns 2j and 2j+1
                if (uplo == 'U') {
                    for (i64 k = colptr[j]; k < colptr[j + 1] && rowind[k] < j; k++) {
                        push(2 * rowind[k], k, half == 0 ? 0 : 2);
                        push(2 * rowind[k] + 1, k, half == 0 ? 1 : 0);
                    }
                    if (half == 0) push(2 * j, dsrc, 0);
                    else { push(2 * j, -1, 0); push(2 * j + 1, dsrc, 0); }
                } else {
                    if (half == 0) { push(2 * j, dsrc, 0); push(2 * j + 1, -1, 0); }
                    else push(2 * j + 1, dsrc, 0);
                    for (i64 k = colptr[j]; k < colptr[j + 1]; k++) {
                        if (rowind[k] <= j) continue;
                        push(2 * rowind[k], k, half == 0 ? 0 : 2);
                        push(2 * rowind[k] + 1, k, half == 0 ? 1 : 0);
                    }
                }
                ecp[2 * j + half + 1] = (i64)eri.size();
            }
        }
        b200s_chol_opts o2;
        if (opts) o2 = *opts; else b200s_chol_default_opts(&o2);
        o2.nmethods = 1;
        b200s_chol* F = nullptr;
        b200s_status st = b200s_chol_analyze(2 * n, ecp.data(), eri.data(), uplo, n > 0 ? p2.data() : nullptr, &o2, &F);
        if (st != B200S_OK) return st;
        for (i64 k = 0; k < n; k++)
            if ((F->plan().perm[2 * k] & 1) || F->plan().perm[2 * k + 1] != F->plan().perm[2 * k] + 1) {
                set_last_error("analyze_z: the ordering separated a (re, im) pair");
                b200s_chol_free(F);
                return B200S_INVALID;
            }
        F->zn = n;
        if (n > 0) { F->zAp.assign(colptr, colptr + n + 1); F->zAi.assign(rowind, rowind + colptr[n]); }
        F->zsrc = std::move(zsrc); F->zcode = std::move(zc);
        *out = F;
        return B200S_OK;
    } catch (const std::bad_alloc&) {
        return B200S_OUT_OF_MEMORY;
    } catch (const std::exception& e) {
        set_last_error(e.what());
        return B200S_INVALID;
    }
}

/* val: nnz complex numbers as (re, im) pairs, the values of the SAME (colptr, rowind) given to analyze_z; imaginary parts of
 * the diagonal are ignored (Hermitian).  *minor_out is a complex column index. */
b200s_status b200s_chol_factorize_z(b200s_chol* F, const b200s_int* colptr, const b200s_int* rowind, const double* val, b200s_int* minor_out) {
    if (!F) return B200S_INVALID;
    const i64 n = F->zn;
    if (F->plan().n == 0) return factorize_impl(F, nullptr, false, minor_out);
    if (n == 0) { set_last_error("factorize_z on a real factor object"); return B200S_INVALID; }
    if (!val) return B200S_INVALID;
    if (colptr) {
        const i64 nnz = F->zAp[n];
        if (colptr[n] != nnz || memcmp(colptr, F->zAp.data(), sizeof(i64) * (size_t)(n + 1)) ||
            (nnz > 0 && (!rowind || memcmp(rowind, F->zAi.data(), sizeof(i64) * (size_t)nnz)))) {
            set_last_error("numeric: the pattern of a complex matrix must be the one analysed by symbolic()");
            return B200S_INVALID;
        }
    }
    try { F->zval.resize(F->zsrc.size()); } catch (const std::bad_alloc&) { return B200S_OUT_OF_MEMORY; }
    for (size_t e = 0; e < F->zsrc.size(); e++) {
        const i64 k = F->zsrc[e];
        F->zval[e] = k < 0 ? 0.0 : (F->zcode[e] == 0 ? val[2 * k] : (F->zcode[e] == 1 ? val[2 * k + 1] : -val[2 * k + 1]));
    }
    b200s_int m = 0;
    b200s_status st = factorize_impl(F, F->zval.data(), false, &m);
    if (minor_out) *minor_out = m / 2;
    return st;
}

/* sparse complex right-hand sides: n x ncols CCS with (re, im) values; result likewise (library-allocated, b200s_free) */
b200s_status b200s_chol_spsolve_z(b200s_chol* F, int sys, b200s_int nrows, b200s_int ncols, const b200s_int* Bp, const b200s_int* Bi,
                                  const double* Bx, b200s_int** Xp, b200s_int** Xi, double** Xx) {
    if (!F || !Xp || !Xi || !Xx || nrows != F->zn || ncols < 0) return B200S_INVALID;
    *Xp = nullptr; *Xi = nullptr; *Xx = nullptr;
    std::vector<i64> rp((size_t)ncols + 1, 0), ri;
    std::vector<double> rx;
    try {
        for (i64 j = 0; j < ncols; j++) {
            for (i64 k = Bp[j]; k < Bp[j + 1]; k++) {
                ri.push_back(2 * Bi[k]); rx.push_back(Bx[2 * k]);
                ri.push_back(2 * Bi[k] + 1); rx.push_back(Bx[2 * k + 1]);
            }
            rp[j + 1] = (i64)ri.size();
        }
    } catch (const std::bad_alloc&) { return B200S_OUT_OF_MEMORY; }
    static const i64 dummy_i = 0; static const double dummy_x = 0.0;
    b200s_int *xp = nullptr, *xi = nullptr;
    double* xx = nullptr;
    b200s_status st = b200s_chol_spsolve(F, sys, 2 * nrows, ncols, rp.data(), ri.empty() ? &dummy_i : ri.data(), rx.empty() ? &dummy_x : rx.data(), &xp, &xi, &xx);
    if (st != B200S_OK) return st;
    // pairs (2i, 2i+1) -> one complex entry
    const i64 rn = xp[ncols];
    b200s_int* zp = (b200s_int*)malloc(sizeof(b200s_int) * (size_t)(ncols + 1));
    b200s_int* zi = (b200s_int*)malloc(sizeof(b200s_int) * (size_t)std::max<i64>(rn, 1));
    double* zx = (double*)malloc(sizeof(double) * 2 * (size_t)std::max<i64>(rn, 1));
    if (!zp || !zi || !zx) { free(zp); free(zi); free(zx); free(xp); free(xi); free(xx); return B200S_OUT_OF_MEMORY; }
    i64 q = 0;
    zp[0] = 0;
    for (i64 j = 0; j < ncols; j++) {
        for (i64 k = xp[j]; k < xp[j + 1]; k++) {
            const i64 i = xi[k] >> 1;
            if (q > zp[j] && zi[q - 1] == i) { zx[2 * (q - 1) + (xi[k] & 1)] = xx[k]; continue; }
            zi[q] = i; zx[2 * q] = 0.0; zx[2 * q + 1] = 0.0; zx[2 * q + (xi[k] & 1)] = xx[k]; q++;
        }
        zp[j + 1] = q;
    }
    free(xp); free(xi); free(xx);
    *Xp = zp; *Xi = zi; *Xx = zx;
    return B200S_OK;
}

/* diagonal of the complex Cholesky factor as n (re, im) pairs (real and positive: im = 0) */
b200s_status b200s_chol_diag_z(b200s_chol* F, double* d_out) {
    if (!F || !d_out) return B200S_INVALID;
    const i64 n = F->zn;
    std::vector<double> d((size_t)2 * n + 1);
    b200s_status st = b200s_chol_diag(F, d.data());
    if (st != B200S_OK) return st;
    for (i64 k = 0; k < n; k++) { d_out[2 * k] = d[2 * k]; d_out[2 * k + 1] = 0.0; }
    return B200S_OK;
}

/* the complex factor as CCS with (re, im) values (cholmod.getfactor of a 'z' factor) */
b200s_status b200s_chol_get_L_z(b200s_chol* F, b200s_int** Lp, b200s_int** Li, double** Lx) {
    if (!F || !Lp || !Li || !Lx) return B200S_INVALID;
    *Lp = nullptr; *Li = nullptr; *Lx = nullptr;
    const i64 n = F->zn;
    b200s_int *rp = nullptr, *ri = nullptr;
    double* rx = nullptr;
    b200s_status st = b200s_chol_get_L(F, &rp, &ri, &rx);
    if (st != B200S_OK) return st;
    i64 nnz = 0;
    for (i64 j = 0; j < n; j++) nnz += rp[2 * j + 1] - rp[2 * j];
    b200s_int* zp = (b200s_int*)malloc(sizeof(b200s_int) * (size_t)(n + 1));
    b200s_int* zi = (b200s_int*)malloc(sizeof(b200s_int) * (size_t)std::max<i64>(nnz, 1));
    double* zx = (double*)malloc(sizeof(double) * 2 * (size_t)std::max<i64>(nnz, 1));
    if (!zp || !zi || !zx) { free(zp); free(zi); free(zx); free(rp); free(ri); free(rx); return B200S_OUT_OF_MEMORY; }
    i64 q = 0;
    zp[0] = 0;
    for (i64 j = 0; j < n; j++) {
        for (i64 k = rp[2 * j]; k < rp[2 * j + 1]; k++) {          // embedded column 2j: rows 2i carry re(l_ij), rows 2i+1 im(l_ij)
            const i64 i = ri[k] >> 1;
            if (q > zp[j] && zi[q - 1] == i) { zx[2 * (q - 1) + (ri[k] & 1)] = rx[k]; continue; }
            zi[q] = i; zx[2 * q] = 0.0; zx[2 * q + 1] = 0.0; zx[2 * q + (ri[k] & 1)] = rx[k]; q++;
        }
        zp[j + 1] = q;
    }
    free(rp); free(ri); free(rx);
    *Lp = zp; *Li = zi; *Lx = zx;
    return B200S_OK;
}

b200s_status b200s_chol_diag(b200s_chol* F, double* d_out) {
    B200S_NVTX("b200s_chol_diag");
    if (!F) return B200S_INVALID;
    if (F->plan().n == 0) return B200S_OK;
    if (!d_out) return B200S_INVALID;
    if (!F->numeric || !F->dev) { set_last_error("called with symbolic factor"); return B200S_INVALID; }
    if (F->ldl) { set_last_error("F must be a nonsingular supernodal Cholesky factor"); return B200S_INVALID; }    /* cholmod.c:919-922 */
    return (b200s_status)chol_device_diag(F->dev, d_out);
}

b200s_status b200s_chol_get_L(b200s_chol* F, b200s_int** Lp, b200s_int** Li, double** Lx) {
    B200S_NVTX("b200s_chol_get_L");
    if (!F || !Lp || !Li || !Lx) return B200S_INVALID;
    *Lp = nullptr; *Li = nullptr; *Lx = nullptr;
    const CholPlan& P = F->plan();
    const i64 n = P.n;
    if (n > 0 && (!F->numeric || !F->dev)) { set_last_error("called with symbolic factor"); return B200S_INVALID; }
    std::vector<double> raw;
    try { raw.resize((size_t)std::max<i64>(P.lsize, 1)); } catch (const std::bad_alloc&) { return B200S_OUT_OF_MEMORY; }
    if (n > 0) {
        int st = chol_device_download_L(F->dev, raw.data());
        if (st != ST_OK) return (b200s_status)st;
    }
    std::vector<double> sign;
    if (F->ldl && n > 0) {
        try { sign.resize((size_t)n); } catch (const std::bad_alloc&) { return B200S_OUT_OF_MEMORY; }
        int st = chol_device_download_sign(F->dev, sign.data());
        if (st != ST_OK) return (b200s_status)st;
    }
    i64 nnz = 0;
    for (const Front& f : P.fronts)
        for (i32 c = 0; c < f.nc; c++)
            for (i32 r = c; r < f.nr; r++)
                if (r == c || raw[f.loff + (i64)c * f.ld + r] != 0.0) nnz++;
    b200s_int* lp = (b200s_int*)malloc(sizeof(b200s_int) * (size_t)(n + 1));
    b200s_int* li = (b200s_int*)malloc(sizeof(b200s_int) * (size_t)std::max<i64>(nnz, 1));
    double* lx = (double*)malloc(sizeof(double) * (size_t)std::max<i64>(nnz, 1));
    if (!lp || !li || !lx) { free(lp); free(li); free(lx); return B200S_OUT_OF_MEMORY; }
    i64 p = 0;
    lp[0] = 0;
    for (const Front& f : P.fronts)
        for (i32 c = 0; c < f.nc; c++) {
            // LDL' semantics (supernodal = 0): cholmod_factor_to_sparse returns L with D on its diagonal; from the signed
            // square-root panels (A = Lt S Lt') that is Lt(:,c) / l_cc below the diagonal and s_c l_cc^2 on it
            const double lcc = raw[f.loff + (i64)c * f.ld + c];
            for (i32 r = c; r < f.nr; r++) {
                double v = raw[f.loff + (i64)c * f.ld + r];
                if (r == c || v != 0.0) {
                    li[p] = P.rows[f.rowptr + r];
                    lx[p] = !F->ldl ? v : (r == c ? sign[f.col0 + c] * lcc * lcc : v / lcc);
                    p++;
                }
            }
            lp[f.col0 + c + 1] = p;
        }
    *Lp = lp; *Li = li; *Lx = lx;
    return B200S_OK;
}

b200s_status b200s_chol_info(const b200s_chol* F, b200s_chol_info_t* info) {
    if (!F || !info) return B200S_INVALID;
    const CholPlan& P = F->plan();
    memset(info, 0, sizeof *info);
    info->n = P.n; info->nsuper = (b200s_int)P.fronts.size(); info->nnz_L = P.nnzL; info->nnz_A = P.nnzA;
    info->nlevels = P.nlevels; info->max_front_rows = P.max_nr; info->max_front_cols = P.max_nc;
    info->factor_bytes = P.lsize * 8; info->workspace_bytes = P.wsize * 8;
    info->flops = P.flops; info->flops_potrf = P.flops_potrf; info->flops_trsm = P.flops_trsm; info->flops_syrk = P.flops_syrk;
    info->is_numeric = F->numeric ? 1 : 0; info->minor = F->minor;
    info->ms_h2d = F->times.ms_h2d; info->ms_assemble = F->times.ms_assemble; info->ms_factor = F->times.ms_factor;
    info->ms_total = F->times.ms_total; info->ms_solve = F->times.ms_solve; info->ms_analyze = P.ms_analyze;
    info->ms_dense_update = F->times.ms_dense_update; info->ms_potrf = F->times.ms_potrf;
    info->ms_trsm = F->times.ms_trsm; info->ms_extend = F->times.ms_extend;
    info->flops_update = P.flops_update;
    info->zn = F->zn;
    return B200S_OK;
}
b200s_status b200s_chol_set_profiling(b200s_chol* F, int on) {
    if (!F) return B200S_INVALID;
    F->profiling = on != 0;
    if (F->dev) chol_device_set_profiling(F->dev, F->profiling);
    return B200S_OK;
}
b200s_status b200s_chol_get_perm(const b200s_chol* F, b200s_int* perm_out) {
    if (!F || (!perm_out && F->plan().n > 0)) return B200S_INVALID;
    for (i32 k = 0; k < F->plan().n; k++) perm_out[k] = F->plan().perm[k];
    return B200S_OK;
}
b200s_status b200s_chol_get_super(const b200s_chol* F, b200s_int* super_out, b200s_int* rowptr_out, b200s_int* rows_out) {
    if (!F) return B200S_INVALID;
    const CholPlan& P = F->plan();
    const size_t ns = P.fronts.size();
    for (size_t s = 0; s < ns; s++) {
        if (super_out) super_out[s] = P.fronts[s].col0;
        if (rowptr_out) rowptr_out[s] = P.fronts[s].rowptr;
    }
    if (super_out) super_out[ns] = P.n;
    if (rowptr_out) rowptr_out[ns] = (b200s_int)P.rows.size();
    if (rows_out) for (size_t k = 0; k < P.rows.size(); k++) rows_out[k] = P.rows[k];
    return B200S_OK;
}
void b200s_chol_free(b200s_chol* F) {
    if (!F) return;
    if (F->dev) {
        chol_device_sync(F->dev);
        // an object whose front ownership was changed (multi-GPU driver) is not handed on
        chol_device_set_solve_sweeps(F->dev, -1);
        if (F->custom_owned || !plan_cache_park_device(F->plan_sp.get(), F->device, F->dev)) chol_device_destroy(F->dev);
    }
    delete F;
}
void b200s_free(void* p) { free(p); }

b200s_status b200s_grid_nd_perm(b200s_int nx, b200s_int ny, b200s_int nz, b200s_int leaf, b200s_int* perm_out) {
    if (nx < 1 || ny < 1 || nz < 1 || !perm_out || nx * ny * nz > 0x7fffffff) return B200S_INVALID;
    std::vector<i32> p = grid_nd(nx, ny, nz, leaf);
    for (size_t k = 0; k < p.size(); k++) perm_out[k] = p[k];
    return B200S_OK;
}
b200s_status b200s_chol_set_solve_sweeps(b200s_chol* F, int mode) {
    if (!F || mode < -1 || mode > 3) return B200S_INVALID;
    if (F->plan().n == 0) return B200S_OK;
    b200s_status es = ensure_device(F);
    if (es != B200S_OK) return es;
    chol_device_set_solve_sweeps(F->dev, mode);
    return B200S_OK;
}
b200s_status b200s_chol_child_lists_check(const b200s_chol* F) {
    if (!F) return B200S_INVALID;
    try { return (b200s_status)chol_child_lists_check(F->plan()); }
    catch (const std::bad_alloc&) { return B200S_OUT_OF_MEMORY; }
}
int b200s_persist_schedule_check(b200s_int nr, b200s_int nc, b200s_int nctas) {
    if (nr < 1 || nr > 0x3fffffff || nc < 1 || nc > nr || nctas < 1 || nctas > 4096) return -1;
    return persist_schedule_check((int)nr, (int)nc, (int)nctas);
}
b200s_status b200s_amd_order(b200s_int n, const b200s_int* colptr, const b200s_int* rowind, char uplo, b200s_int* perm_out) {
    if (n < 0 || n > 0x7fffffff - 16 || (n > 0 && (!colptr || !perm_out))) return B200S_INVALID;
    try {
        std::vector<i32> p = amd_order(sym_pattern_from_triangle(n, colptr, rowind, uplo));
        for (size_t k = 0; k < p.size(); k++) perm_out[k] = p[k];
    } catch (const std::exception& e) {
        set_last_error(e.what());
        return B200S_INVALID;
    }
    return B200S_OK;
}

}  // extern "C"
