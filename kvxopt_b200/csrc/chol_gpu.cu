// Numeric phase of the multifrontal supernodal Cholesky on B200 (sm_100a) and the supernodal
// triangular solves.  This is what runs in place of cholmod_l_factorize (src/C/cholmod.c:362,677,824)
// and cholmod_l_solve (src/C/cholmod.c:483,735).
//
// Data layout in HBM (all FP64, column-major):
//   L   : one nr x nc panel per front, leading dimension ld (even), 128-byte aligned; rows follow the
//         front's sorted row list, the first nc rows are the pivots (dense nc x nc lower triangle on top).
//   W   : update (Schur complement) matrices, one per front with rows below the pivots, placed by
//         lifetime (host.hpp Front::uoff); only the lower triangle is ever touched.
// Schedule: fronts are grouped by level of the supernodal elimination tree (leaves first).  Per level:
//   k_extend_add  zeroes the update matrix and adds the children's update matrices into the front
//                 (parent-driven, children in fixed order => deterministic sums, no atomics);
//   k_small_front fronts with <= 128 rows: whole front factored in shared memory by one CTA;
//   k_panel       large fronts, per 128-column block: diagonal-block Cholesky in shared memory +
//                 triangular solve of 64-row tiles (FP64 DMMA for the GEMM part, register substitution);
//   k_update      C -= A_i A_j^T on 128x128 tiles with FP64 tensor-core DMMA (mma.sync m8n8k4.f64),
//                 operands staged by a 2-stage cp.async pipeline of 32-column k-tiles; used for the in-panel trailing update
//                 (K = 128) and for the Schur complement (K = nc).
// Solves: k_fwd / k_bwd (small fronts and, in the solve phase, one-block fronts with few rows: one CTA per front);
//   k_fwd_gather + k_fwd_diag / k_fwd_upd and k_bwd_gather + k_bwd_upd / k_bwd_diag (large fronts, two launches per
//   128-column block step: several right-hand sides, ownership-masked distributed sweeps, wide levels);
//   k_fwd_persist / k_bwd_persist / k_bwd_rect (one right-hand side, levels with <= 32 large fronts: ONE cooperative kernel
//   per level and direction, CTAs hand the solved block on through release / acquire flags; bit-identical to the above).
#include "gpu.hpp"
#include "devpool.hpp"
#include <cuda_runtime.h>
#include <climits>
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <vector>
#include <mutex>

namespace b200s {

// ---------------------------------------------------------------------------------------------------
// error plumbing
// ---------------------------------------------------------------------------------------------------
static thread_local std::string g_last_error;
static int g_device = 0;
void set_last_error(const std::string& s) { g_last_error = s; }
const char* get_last_error() { return g_last_error.c_str(); }
int current_device() { return g_device; }
void set_current_device(int d) { g_device = d; }
int device_count() {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

#define CUDA_TRY(expr)                                                                               \
    do {                                                                                             \
        cudaError_t e__ = (expr);                                                                    \
        if (e__ != cudaSuccess) {                                                                    \
            char buf__[512];                                                                         \
            snprintf(buf__, sizeof buf__, "%s failed at %s:%d: %s", #expr, __FILE__, __LINE__,       \
                     cudaGetErrorString(e__));                                                       \
            set_last_error(buf__);                                                                   \
            return e__ == cudaErrorMemoryAllocation ? ST_OOM : ST_CUDA;                              \
        }                                                                                            \
    } while (0)

// ---------------------------------------------------------------------------------------------------
// device structures
// ---------------------------------------------------------------------------------------------------
struct FrontD {
    long long loff, uoff, reloff, rowptr;
    long long ioff;               // large fronts: offset of the inverted diagonal sub-blocks in Minv (MINV_BLK per block column)
    int col0, nc, nr, ld;
    int nchild, childptr, parent, level;
};
// one CTA of k_extend_add: the destination columns [c0, c1) of a front; [q_off, q_off + q_cnt) in the item-child list = the
// positions (ascending: the order of summation) of the front's children that have an update column in that range
struct EAItem { int front, c0, c1, q_off, q_cnt; };

constexpr int NB = PLAN_NB;    // block-column width inside large fronts
constexpr int TR = 64;         // rows per CTA in the panel triangular solve
constexpr int LDL = NB + 4;    // smem stride of the diagonal block   (stride % 16 == 4 -> conflict-free DMMA fragment loads)
constexpr int LDX = TR + 4;    // smem stride of the row tile
constexpr int SMALL_NR = PLAN_SMALL_NR;  // fronts with nr <= SMALL_NR are factored by one CTA in shared memory
constexpr int BT = 128;        // update tile is BT x BT
constexpr int BK = 32;         // k-depth per pipeline stage
constexpr int STAGES = 2;      // 2 x 32 columns: half the CTA barriers of 4 x 16 at the same shared memory (216.4 -> 212.2 ms on 100^3)
constexpr int LDT = BT + 4;    // smem stride of operand tiles

__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem, int src_bytes) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(s), "l"(gmem), "r"(src_bytes));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;"); }
// TMA bulk copies (cp.async.bulk, 1-D) completing on an mbarrier: the staging path of k_update's operand tiles
__device__ __forceinline__ void mbar_init(unsigned long long* b, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(b)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* b, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* b, unsigned parity) {
    const unsigned a = (unsigned)__cvta_generic_to_shared(b);
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(a), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* smem, const void* gmem, unsigned bytes, unsigned long long* b) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
                 "r"((unsigned)__cvta_generic_to_shared(smem)), "l"(gmem), "r"(bytes), "r"((unsigned)__cvta_generic_to_shared(b)) : "memory");
}
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N)); }

__device__ __forceinline__ int lower_bound_dev(const int* a, int n, int v) {
    int lo = 0, hi = n;
    while (lo < hi) { int mid = (lo + hi) >> 1; if (a[mid] < v) lo = mid + 1; else hi = mid; }
    return lo;
}
// group lookup: prefix[0..ng] increasing, returns g with prefix[g] <= b < prefix[g+1]
__device__ __forceinline__ int find_group(const int* prefix, int ng, int b) {
    int lo = 0, hi = ng;
    while (hi - lo > 1) { int mid = (lo + hi) >> 1; if (prefix[mid] <= b) lo = mid; else hi = mid; }
    return lo;
}

// Launch metadata of the panel / update / solve kernels passed BY VALUE (constant bank) when a block step covers few fronts -- which is
// every step of the top levels, where the steps are many and short: no dependent global loads (group prefix -> front id
// -> front record) ahead of the first data access.  ng == 0: read the schedule arrays in global memory instead.
struct FrontS { long long loff, rowptr, ioff, uoff; int nc, nr, ld, col0, id, pad_; };
constexpr int MAXG = 32;
struct SolveGroups { int ng; int prefix[MAXG + 1]; FrontS fr[MAXG]; };
__device__ __forceinline__ FrontS load_front(const SolveGroups& sg, const int* gfront, const FrontD* F, int g) {
    if (sg.ng) return sg.fr[g];
    const FrontD fd = F[gfront[g]];
    FrontS f;
    f.loff = fd.loff; f.rowptr = fd.rowptr; f.ioff = fd.ioff; f.uoff = fd.uoff; f.nc = fd.nc; f.nr = fd.nr; f.ld = fd.ld;
    f.col0 = fd.col0; f.id = gfront[g]; f.pad_ = 0;
    return f;
}
__device__ __forceinline__ int locate_group(const SolveGroups& sg, const int* gprefix, int ngroups, int b, int& tile) {
    if (sg.ng) {
        int g = 0;
        while (g + 1 < sg.ng && sg.prefix[g + 1] <= b) g++;
        tile = b - sg.prefix[g];
        return g;
    }
    const int g = find_group(gprefix, ngroups, b);
    tile = b - gprefix[g];
    return g;
}

// ---------------------------------------------------------------------------------------------------
// K1: scatter the caller's CCS values into the (zeroed) panels
// ---------------------------------------------------------------------------------------------------
__global__ void k_set_int(int* p, int v) { *p = v; }
__global__ void k_scatter_A(const double* __restrict__ val, const long long* __restrict__ amap, long long nnz,
                            double* __restrict__ L) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x, st = (long long)gridDim.x * blockDim.x;
    for (; i < nnz; i += st) {
        long long m = amap[i];
        if (m >= 0) L[m] = val[i];
    }
}

// ---------------------------------------------------------------------------------------------------
// K5: extend-add.  One CTA per (front, column slab): zero the slab of the update matrix, then add every
// child's update matrix entries that fall into the slab, children in ascending order.
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_extend_add(const EAItem* __restrict__ items, const FrontD* __restrict__ F,
                                                    const int* __restrict__ child_idx, const int* __restrict__ rel,
                                                    double* __restrict__ L, double* __restrict__ W,
                                                    const unsigned char* __restrict__ owned, const int* __restrict__ item_child) {
    const EAItem it = items[blockIdx.x];
    if (!owned[it.front]) return;
    const FrontD fp = F[it.front];
    const int nc = fp.nc, nr = fp.nr, uo = nc & 1;
    const int ldu = ((nr - nc + uo) + 1) & ~1;
    double* Wp = W + fp.uoff;
    double* Lp = L + fp.loff;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int c = max(it.c0, nc) + warp; c < it.c1; c += 8) {
        double* col = Wp + (long long)(c - nc + uo) * ldu + (uo - nc);
        for (int r = c + lane; r < nr; r += 32) col[r] = 0.0;
    }
    __syncthreads();
    // only the children that reach into [c0, c1) (a front of a KKT matrix can have thousands of one-entry children: every CTA
    // of the root front of BASELINE config 5's 3 x 3 system looped over 18 800 of them, 44 of the 77 ms of that factorization)
    for (int t = 0; t < it.q_cnt; t++) {
        const FrontD fc = F[child_idx[fp.childptr + item_child[it.q_off + t]]];
        const int mc = fc.nr - fc.nc, uoc = fc.nc & 1;
        const int ldc = ((mc + uoc) + 1) & ~1;
        const int* rl = rel + fc.reloff;
        const double* Wc = W + fc.uoff;
        const int j0 = lower_bound_dev(rl, mc, it.c0), j1 = lower_bound_dev(rl, mc, it.c1);
        for (int j = j0 + warp; j < j1; j += 8) {
            const int pc = rl[j];
            const double* src = Wc + (long long)(j + uoc) * ldc + uoc;
            double* dst = (pc < nc) ? Lp + (long long)pc * fp.ld : Wp + (long long)(pc - nc + uo) * ldu + (uo - nc);
            // (four independent read-modify-writes per lane and trip were tried: 20.6 -> 23.2 ms on 100^3, the loop is not latency-bound)
            for (int i = j + lane; i < mc; i += 32) dst[rl[i]] += src[i];
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------------------
// right-looking Cholesky of the first nc columns of an nr x nr lower-triangular matrix held in shared
// memory (column-major, stride lds): potrf + trsm + syrk of a small front in one routine.
// ---------------------------------------------------------------------------------------------------
template <int THREADS, bool SGN>
__device__ __forceinline__ void smem_partial_chol(double* S, int lds, int nr, int nc, int gcol0, int* minor,
                                                  double dbound, bool record, double* __restrict__ sgn) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int NW = THREADS / 32;
    for (int j = 0; j < nc; j++) {
        __syncthreads();
        const double d = S[j * lds + j];
        // SGN: signed factorization A = L S L', S = diag(+-1) (LDL' without pivoting: D = S diag(L)^2); only an exactly
        // zero (or NaN) pivot stops it, as in CHOLMOD's simplicial LDL'
        const double sj = (SGN && d < 0.0) ? -1.0 : 1.0;
        const double ad = SGN ? fabs(d) : d;
        if (!(ad > 0.0) && record && tid == 0) atomicMin(minor, gcol0 + j);
        double l = sqrt(ad);
        if (dbound > 0.0 && l < dbound) l = dbound;
        const double inv = SGN ? sj / l : 1.0 / l;
        for (int i = j + 1 + tid; i < nr; i += THREADS) S[j * lds + i] *= inv;
        __syncthreads();
        if (tid == 0) { S[j * lds + j] = l; if (SGN) sgn[gcol0 + j] = sj; }
        for (int c = j + 1 + warp; c < nr; c += NW) {
            const double lc = SGN ? sj * S[j * lds + c] : S[j * lds + c];
            if (lc != 0.0)
                for (int i = c + lane; i < nr; i += 32) S[c * lds + i] -= S[j * lds + i] * lc;
        }
    }
    __syncthreads();
}

// Blocked Cholesky of a w x w (w <= 128, padded to wpad = multiple of 16 with an identity tail) lower-triangular
// block held in shared memory with stride LDL, 256 threads, 16 columns per step:
//  (1) the 16 x 16 diagonal chunk is factored by warp 0 in registers (one lane per row, pivot column broadcast by
//      shuffles, rsqrt instead of sqrt + divide);
//  (2) one thread per row solves the rows below against it;
//  (3) all warps apply the rank-16 update to the trailing lower triangle with FP64 DMMA, 8 x 32 strips of tiles handed
//      out through a counter; warp 0 first updates the next diagonal chunk and factors it while the others work on the
//      rest (look-ahead), so step (1) is off the critical path after the first chunk.
// Only the lower triangle of S is meaningful afterwards.
template <bool SGN>
__device__ __forceinline__ void chol_diag16(double* S, int d0, int lane, double* rdiag, double dbound, int w, int gcol0,
                                            int* minor, bool record, double* ssign) {
    const int r = lane & 15;
    double a[16];
#pragma unroll
    for (int c = 0; c < 16; c++) a[c] = S[(d0 + c) * LDL + d0 + r];
    int bad = 16;
#pragma unroll
    for (int j = 0; j < 16; j++) {
        const double d = __shfl_sync(0xffffffffu, a[j], j);
        // SGN: A = L S L' with S = diag(+-1); rdiag carries the sign (1 / (s_j l_jj)), ssign[] the signs of the block
        const double sj = (SGN && d < 0.0) ? -1.0 : 1.0;
        const double ad = SGN ? fabs(d) : d;
        if (!(ad > 0.0) && j < bad) bad = j;
        double inv = rsqrt(ad);
        double l = ad * inv;
        if (dbound > 0.0 && l < dbound) { l = dbound; inv = 1.0 / dbound; }
        if (SGN) inv *= sj;
        const double lij = (r == j) ? l : a[j] * inv;
        if (r >= j) a[j] = lij;
        const double mlij = SGN ? sj * lij : lij;
#pragma unroll
        for (int c = 0; c < 16; c++)
            if (c > j) {                 // static after unrolling: keeps a[] in registers
                const double u = __shfl_sync(0xffffffffu, lij, c);
                const double t = fma(-mlij, u, a[c]);
                a[c] = (r >= c) ? t : a[c];
            }
        if (lane == j) { rdiag[j] = inv; if (SGN) ssign[d0 + j] = sj; }
    }
    if (lane < 16) {
#pragma unroll
        for (int c = 0; c < 16; c++)
            if (r >= c) S[(d0 + c) * LDL + d0 + r] = a[c];
    }
    if (record && lane == 0 && bad < 16 && d0 + bad < w) atomicMin(minor, gcol0 + d0 + bad);
}

// C(8 x 32 strip: tile row ti, tile columns tj0..tj0+3, those in qmask) -= L(rows, c0..c0+15) L(cols, c0..c0+15)^T
template <bool SGN>
__device__ __forceinline__ void chol_strip(double* S, int c0, int t0, int ti, int tj0, int qmask, int lane, const double* ssign) {
    double* Cp = S + (t0 + 8 * tj0 + 2 * (lane & 3)) * LDL + t0 + 8 * ti + (lane >> 2);
    const double* Ap = S + (c0 + (lane & 3)) * LDL + t0 + 8 * ti + (lane >> 2);
    const double* Bp = S + (c0 + (lane & 3)) * LDL + t0 + 8 * tj0 + (lane >> 2);
    double acc[4][2];
#pragma unroll
    for (int q = 0; q < 4; q++)
        if (qmask >> q & 1) { acc[q][0] = Cp[q * 8 * LDL]; acc[q][1] = Cp[q * 8 * LDL + LDL]; }
#pragma unroll
    for (int k4 = 0; k4 < 16; k4 += 4) {
        const double av = SGN ? -Ap[k4 * LDL] * ssign[c0 + k4 + (lane & 3)] : -Ap[k4 * LDL];
#pragma unroll
        for (int q = 0; q < 4; q++)
            if (qmask >> q & 1) dmma884(acc[q][0], acc[q][1], av, Bp[k4 * LDL + 8 * q]);
    }
#pragma unroll
    for (int q = 0; q < 4; q++)
        if (qmask >> q & 1) { Cp[q * 8 * LDL] = acc[q][0]; Cp[q * 8 * LDL + LDL] = acc[q][1]; }
}

template <bool SGN>
__device__ __forceinline__ void smem_potrf_blocked(double* S, int w, int wpad, int gcol0, int* minor, double dbound,
                                                   bool record, double* rdiag, double* ssign) {
    __shared__ int tctr;
    __shared__ unsigned char strip_ti[32], strip_tj[32];      // strips of the lower triangle of <= 14 x 14 tiles, by row
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) {
        int n = 0;
        for (int ti = 0; ti < 14; ti++)
            for (int tj0 = 0; tj0 <= ti; tj0 += 4) { strip_ti[n] = (unsigned char)ti; strip_tj[n] = (unsigned char)tj0; n++; }
    }
    __syncthreads();
    if (warp == 0) chol_diag16<SGN>(S, 0, lane, rdiag, dbound, w, gcol0, minor, record, ssign);
    for (int c0 = 0; c0 < wpad; c0 += 16) {
        __syncthreads();                       // diagonal chunk c0 factored, trailing update of the previous chunk done
        const int t0 = c0 + 16;
        if (tid == 0) tctr = 0;
        for (int r = t0 + tid; r < wpad; r += 256) {
            double xv[16];
#pragma unroll
            for (int q = 0; q < 16; q++) xv[q] = S[(c0 + q) * LDL + r];
#pragma unroll
            for (int q = 0; q < 16; q++) {
                double v = xv[q];
#pragma unroll
                for (int p = 0; p < q; p++) v = fma(-xv[p], S[(c0 + p) * LDL + c0 + q], v);
                xv[q] = SGN ? v * fabs(rdiag[q]) : v * rdiag[q];     // SGN: s_q l_rq = v / l_qq, which is what the sums need
            }
            if (SGN) {
#pragma unroll
                for (int q = 0; q < 16; q++) xv[q] *= ssign[c0 + q];       // back to l_rq
            }
#pragma unroll
            for (int q = 0; q < 16; q++) S[(c0 + q) * LDL + r] = xv[q];
        }
        __syncthreads();
        const int nt = (wpad - t0) >> 3;
        if (nt == 0) break;
        int nstrips = 0;                       // strips with ti < nt: a prefix of the table
        for (int ti = 0; ti < nt; ti++) nstrips += (ti >> 2) + 1;
        if (warp == 0) {
            chol_strip<SGN>(S, c0, t0, 0, 0, 1, lane, ssign);
            chol_strip<SGN>(S, c0, t0, 1, 0, 3, lane, ssign);
            __syncwarp();
            chol_diag16<SGN>(S, t0, lane, rdiag, dbound, w, gcol0, minor, record, ssign);
        }
        int mt = 0;
        if (lane == 0) mt = atomicAdd(&tctr, 1);
        mt = __shfl_sync(0xffffffffu, mt, 0);
        while (mt < nstrips) {
            int nxt = 0;
            if (lane == 0) nxt = atomicAdd(&tctr, 1);      // next ticket fetched under the DMMA work of this strip
            const int ti = strip_ti[mt], tj0 = strip_tj[mt];
            int qmask = (ti - tj0 >= 3) ? 15 : ((1 << (ti - tj0 + 1)) - 1);
            if (ti < 2 && tj0 == 0) qmask = 0;             // the next diagonal chunk: done by warp 0 above
            if (qmask) chol_strip<SGN>(S, c0, t0, ti, tj0, qmask, lane, ssign);
            mt = __shfl_sync(0xffffffffu, nxt, 0);
        }
    }
    __syncthreads();
}

// K6: one CTA per small front (nr <= SMALL_NR): load panel + update matrix, factor, write back.
template <int THREADS, bool SGN>
__global__ void __launch_bounds__(THREADS) k_small_front(const int* __restrict__ list, const FrontD* __restrict__ F,
                                                         double* __restrict__ L, double* __restrict__ W, int* minor,
                                                         double dbound, const unsigned char* __restrict__ owned,
                                                         double* __restrict__ sgn) {
    extern __shared__ double S[];
    if (!owned[list[blockIdx.x]]) return;
    const FrontD f = F[list[blockIdx.x]];
    const int nr = f.nr, nc = f.nc, m = nr - nc, uo = nc & 1, lds = nr | 1;
    const int ldu = ((m + uo) + 1) & ~1;
    double* P = L + f.loff;
    double* U = W + f.uoff;
    const int tid = threadIdx.x;
    for (int idx = tid; idx < nr * nc; idx += THREADS) {
        int c = idx / nr, r = idx - c * nr;
        S[c * lds + r] = (r >= c) ? P[(long long)c * f.ld + r] : 0.0;
    }
    for (int idx = tid; idx < m * m; idx += THREADS) {
        int c = idx / m, r = idx - c * m;
        if (r >= c) S[(nc + c) * lds + nc + r] = U[(long long)(c + uo) * ldu + r + uo];
    }
    smem_partial_chol<THREADS, SGN>(S, lds, nr, nc, f.col0, minor, dbound, true, sgn);
    for (int idx = tid; idx < nr * nc; idx += THREADS) {
        int c = idx / nr, r = idx - c * nr;
        if (r >= c) P[(long long)c * f.ld + r] = S[c * lds + r];
    }
    for (int idx = tid; idx < m * m; idx += THREADS) {
        int c = idx / m, r = idx - c * m;
        if (r >= c) U[(long long)(c + uo) * ldu + r + uo] = S[(nc + c) * lds + nc + r];
    }
}

// ---------------------------------------------------------------------------------------------------
// K2+K3: panel kernel for block column kb of large fronts.  CTA 0 of a front factors the diagonal block
// and writes it back; every other CTA factors the same block redundantly in shared memory (no
// inter-CTA dependency, no extra launch) and solves X L11^T = B for its 64-row tile.
// ---------------------------------------------------------------------------------------------------
template <bool SGN>
__global__ void __launch_bounds__(256, 1) k_panel(const __grid_constant__ SolveGroups sg, const int* __restrict__ gfront, const int* __restrict__ gprefix,
                                                  int ngroups, int kb, const FrontD* __restrict__ F,
                                                  double* __restrict__ L, double* __restrict__ diag_scratch,
                                                  int* minor, double dbound, const unsigned char* __restrict__ owned,
                                                  double* __restrict__ sgn) {
    extern __shared__ double sm[];
    double* Ls = sm;
    double* Xs = Ls + NB * LDL;
    double* rinv = Xs + NB * LDX;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    int r;
    const int g = locate_group(sg, gprefix, ngroups, blockIdx.x, r);
    const FrontS f = load_front(sg, gfront, F, g);
    if (!owned[f.id]) return;
    const int k0 = kb * NB;
    const int w = min(NB, f.nc - k0);
    const int wpad = (w + 15) & ~15;
    const int ld = f.ld;
    double* P = L + f.loff;
    // diagonal block -> shared memory, 16-byte LDGSTS all in flight: rows below w and row pairs above the diagonal are
    // zero-filled (src-size 0 / 8), pad columns get the identity
    for (int idx = tid; idx < NB * (NB / 2); idx += 256) {
        const int c = idx >> 6, rr = (idx & 63) * 2;
        if (c >= wpad || rr >= wpad) continue;
        if (c < w) {
            const int bytes = (rr + 1 < c) ? 0 : ((rr + 1 < w) ? 16 : ((rr < w) ? 8 : 0));
            const double* src = bytes ? P + (long long)(k0 + c) * ld + k0 + rr : P;
            cp_async16(Ls + c * LDL + rr, src, bytes);
        } else {
            Ls[c * LDL + rr] = (rr == c) ? 1.0 : 0.0;
            Ls[c * LDL + rr + 1] = (rr + 1 == c) ? 1.0 : 0.0;
        }
    }
    cp_async_commit();
    cp_async_wait<0>();
    __shared__ double rdiag16[16];
    __shared__ double ssign[NB];
    smem_potrf_blocked<SGN>(Ls, w, wpad, f.col0 + k0, minor, dbound, r == 0, rdiag16, ssign);
    if (r == 0) {
        if (SGN && tid < w) sgn[f.col0 + k0 + tid] = ssign[tid];
        // The factored block goes to scratch, not to the panel: CTAs of this launch that start later still
        // have to read the UNfactored block.  k_diag_writeback copies it into the panel after this launch.
        double* dst = diag_scratch + (size_t)g * NB * NB;
        for (int idx = tid; idx < w * w; idx += 256) {
            int c = idx / w, rr = idx - c * w;
            if (rr >= c) dst[c * NB + rr] = Ls[c * LDL + rr];
        }
        return;
    }
    if (SGN) {
        // signed factorization A = L S L': the tile solve is X (L11 S)^T = B, i.e. the same substitution against the
        // column-signed block
        for (int idx = tid; idx < wpad * wpad; idx += 256) {
            const int c = idx / wpad, rr = idx - c * wpad;
            if (rr >= c) Ls[c * LDL + rr] *= ssign[c];
        }
        __syncthreads();
    }
    if (tid < wpad) rinv[tid] = 1.0 / Ls[tid * LDL + tid];
    // this CTA solves the 64-row tiles r-1, r-1+nsolve, ... of the rows below the diagonal block with the factor
    // it holds in shared memory (the redundant diagonal factorization is paid once per CTA, not once per tile)
    const int nsolve = (sg.ng ? sg.prefix[g + 1] - sg.prefix[g] : gprefix[g + 1] - gprefix[g]) - 1;
    const int ntiles = (f.nr - (k0 + w) + TR - 1) / TR;
    const int q = lane & 3, rw = warp * 8 + (lane >> 2);
    for (int tile = r - 1; tile < ntiles; tile += nsolve) {
        const int row0 = k0 + w + tile * TR;
        const int nrows = min(TR, f.nr - row0);
        __syncthreads();                 // previous tile fully written back / rinv visible
#pragma unroll 8
        for (int idx = tid; idx < wpad * TR; idx += 256) {
            int c = idx / TR, i = idx - c * TR;
            Xs[c * LDX + i] = (c < w && i < nrows) ? P[(long long)(k0 + c) * ld + row0 + i] : 0.0;
        }
        __syncthreads();
        // ---- X L11^T = B, 16 columns at a time; warp owns rows [8*warp, 8*warp+8), 4 lanes per row
        for (int c0 = 0; c0 < wpad; c0 += 16) {
            double b00 = Xs[(c0 + 2 * q) * LDX + rw], b01 = Xs[(c0 + 2 * q + 1) * LDX + rw];
            double b10 = Xs[(c0 + 8 + 2 * q) * LDX + rw], b11 = Xs[(c0 + 8 + 2 * q + 1) * LDX + rw];
            for (int k4 = 0; k4 < c0; k4 += 4) {
                const double a = -Xs[(k4 + q) * LDX + rw];
                const double l0 = Ls[(k4 + q) * LDL + c0 + (lane >> 2)];
                const double l1 = Ls[(k4 + q) * LDL + c0 + 8 + (lane >> 2)];
                dmma884(b00, b01, a, l0);
                dmma884(b10, b11, a, l1);
            }
            // substitution against the 16x16 diagonal chunk, values stay in registers
#pragma unroll
            for (int p = 0; p < 16; p++) {
                const int owner = (p & 7) >> 1;
                double cand = (p < 8) ? ((p & 1) ? b01 : b00) : ((p & 1) ? b11 : b10);
                cand *= rinv[c0 + p];
                const double xp = __shfl_sync(0xffffffffu, cand, (lane & ~3) | owner);
                if (q == owner) { if (p < 8) { if (p & 1) b01 = xp; else b00 = xp; } else { if (p & 1) b11 = xp; else b10 = xp; } }
                const double* lp = Ls + (c0 + p) * LDL + c0;
                if (2 * q > p) b00 -= xp * lp[2 * q];
                if (2 * q + 1 > p) b01 -= xp * lp[2 * q + 1];
                if (8 + 2 * q > p) b10 -= xp * lp[8 + 2 * q];
                if (8 + 2 * q + 1 > p) b11 -= xp * lp[8 + 2 * q + 1];
            }
            Xs[(c0 + 2 * q) * LDX + rw] = b00;
            Xs[(c0 + 2 * q + 1) * LDX + rw] = b01;
            Xs[(c0 + 8 + 2 * q) * LDX + rw] = b10;
            Xs[(c0 + 8 + 2 * q + 1) * LDX + rw] = b11;
            __syncwarp();
        }
        __syncthreads();
        for (int idx = tid; idx < w * TR; idx += 256) {
            int c = idx / TR, i = idx - c * TR;
            if (i < nrows) P[(long long)(k0 + c) * ld + row0 + i] = Xs[c * LDX + i];
        }
    }
}

__global__ void __launch_bounds__(256) k_diag_writeback(const int* __restrict__ gfront, int kb,
                                                        const FrontD* __restrict__ F, double* __restrict__ L,
                                                        const double* __restrict__ diag_scratch,
                                                        const unsigned char* __restrict__ owned) {
    if (!owned[gfront[blockIdx.x]]) return;
    const FrontD f = F[gfront[blockIdx.x]];
    const int k0 = kb * NB, w = min(NB, f.nc - k0);
    double* P = L + f.loff;
    const double* src = diag_scratch + (size_t)blockIdx.x * NB * NB;
    for (int idx = threadIdx.x; idx < w * w; idx += 256) {
        int c = idx / w, rr = idx - c * w;
        if (rr >= c) P[(long long)(k0 + c) * f.ld + k0 + rr] = src[c * NB + rr];
    }
}

// ---------------------------------------------------------------------------------------------------
// K4: C(128x128 tile) -= A_i A_j^T with FP64 DMMA.  A_i, A_j are 128-row slices of the front's panel
// (column-major, leading dimension ld), K columns starting at k0.
//   mode 0: in-panel trailing update after block column kb (target columns inside the panel)
//   mode 1: Schur complement into the update matrix (K = nc)
// ---------------------------------------------------------------------------------------------------
constexpr int SUPER_NB = 4;        // block columns per super-block of the two-level trailing update
constexpr int UPD_THREADS = 256;   // 8 warps (4 x 2), each a 32 x 32 piece of a 128 x 64 tile; 2 CTAs per SM
constexpr int BTN = 64;            // tile columns
constexpr int LDTB = BTN + 4;      // smem stride of the column-block operand

template <int ROWS, int LDS_>
__device__ __forceinline__ void load_tile_async(double* dst, const double* __restrict__ src, int ld, int rows_valid,
                                                int kcols_valid, int tid) {
    // ROWS rows x BK columns in 16-byte chunks
    constexpr int CPC = ROWS / 2;                      // chunks per column
#pragma unroll
    for (int i = 0; i < CPC * BK / UPD_THREADS; i++) {
        const int ch = tid + i * UPD_THREADS;
        const int c = ch / CPC, r = (ch % CPC) * 2;
        int bytes = 0;
        if (c < kcols_valid) bytes = (r + 1 < rows_valid) ? 16 : ((r < rows_valid) ? 8 : 0);
        const double* s = (bytes > 0) ? src + (long long)c * ld + r : src;
        cp_async16(dst + c * LDS_ + r, s, bytes);
    }
}

// tile decode shared by host counting and the kernel: column tiles of 64, row tiles of 128, lower triangle only:
// column tile cj pairs with row tiles ti >= cj/2
// TMA = true: the operand tiles are staged by TMA bulk copies (one cp.async.bulk per tile column: 1 KB of the row block,
// 512 B of the column block, issued by the lanes of warp 0, completion counted in bytes on one mbarrier per stage) instead of
// 16-byte LDGSTS from all 256 threads.  Rows past the tile edge are simply not copied (they only feed accumulators that are
// never stored); k-columns past K are zeroed by plain stores.
// Schur complement of one front shared by several GPUs (multi-GPU factorization, kvxopt_b200/dist.py): every participant has
// the front's factored panel and computes the column tiles [lo, hi) of C = -L21 L21^T; the front's owner in place in its
// update matrix (which already holds the children's contributions), a helper into a zeroed scratch buffer that the owner of
// the parent front adds to the update matrix it receives.  All arrays are indexed by front; own == nullptr: not in use.
struct SyrkSplit {
    const unsigned char* own;     // this rank computes a slab of the front's Schur complement
    const int *lo, *hi;           // its column tiles (of BTN columns)
    const long long* base;        // LLONG_MIN: in place in W; else the slab starts at scratch[base] (column tile lo)
    double* scratch;
};

template <bool SGN, bool TMA>
__global__ void __launch_bounds__(UPD_THREADS, 2) k_update(const __grid_constant__ SolveGroups sg, const int* __restrict__ gfront, const int* __restrict__ gprefix,
                                                           int ngroups, int mode, int kb, const FrontD* __restrict__ F,
                                                           double* __restrict__ L, double* __restrict__ W,
                                                           const unsigned char* __restrict__ owned, const SyrkSplit sp,
                                                           const double* __restrict__ sgn) {
    extern __shared__ double sm[];
    double* As = sm;                          // [STAGES][BK][LDT]    rows of the tile's row block (128)
    double* Bs = sm + STAGES * BK * LDT;      // [STAGES][BK][LDTB]   rows of the tile's column block (64)
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    int t;
    const int g = locate_group(sg, gprefix, ngroups, blockIdx.x, t);
    const FrontS f = load_front(sg, gfront, F, g);
    if (mode == 1 && sp.own) { if (!sp.own[f.id]) return; }
    else if (!owned[f.id]) return;
    const int nr = f.nr, nc = f.nc, ld = f.ld;
    const double* P = L + f.loff;
    int rowI, rowJ, k0, K, ldc, crows, ccols, lo;
    double* C;
    if (mode != 1) {
        // mode 0: every column tile right of block kb; mode 2: only the next block column (lookahead part A);
        // mode 3: the column tiles right of the next block column (part B, overlapped with the next panel kernel)
        // mode 4: "near" update inside the super-block of SUPER_NB block columns (K = 128, column tiles up to the end
        // of the super-block); mode 5: "far" update after the last panel of a super-block: all column tiles right of
        // it with K = the whole super-block (up to 512), so C is read and written once per super-block
        const int nrt = (nr + BT - 1) / BT;
        // mode 6: far update without the column tiles of the NEXT super-block (those are mode 5 with a short grid: they
        // gate the next panels and run first; the rest overlaps the next super-block's latency-bound panel kernels)
        int cj = 2 * (kb + 1) + (mode == 3 ? 2 : 0) + (mode == 6 ? 2 * SUPER_NB : 0);
        while (t >= nrt - (cj >> 1)) { t -= nrt - (cj >> 1); cj++; }
        const int ti = (cj >> 1) + t;
        rowI = ti * BT; rowJ = cj * BTN;
        k0 = kb * NB; K = min(NB, nc - k0);
        if (mode == 5 || mode == 6) { k0 = (kb / SUPER_NB) * SUPER_NB * NB; K = (kb + 1) * NB - k0; }
        C = L + f.loff + rowI + (long long)rowJ * ld; ldc = ld;
        crows = min(BT, nr - rowI); ccols = min(BTN, nc - rowJ);
        lo = 0;
    } else {
        const int r0 = nc - (nc & 1);
        const int mu = nr - r0, T = (mu + BT - 1) / BT;
        const int ldu = (mu + 1) & ~1;
        int cj = 0;
        while (t >= T - (cj >> 1)) { t -= T - (cj >> 1); cj++; }
        const int ti = (cj >> 1) + t;
        rowI = r0 + ti * BT; rowJ = r0 + cj * BTN;
        k0 = 0; K = nc;
        C = W + f.uoff + (long long)ti * BT + (long long)cj * BTN * ldu; ldc = ldu;
        if (sp.own) {            // a slab of a shared Schur complement
            const int tlo = sp.lo[f.id];
            if (cj < tlo || cj >= sp.hi[f.id]) return;
            const long long base = sp.base[f.id];
            if (base != LLONG_MIN) C = sp.scratch + base + (long long)ti * BT + (long long)(cj - tlo) * BTN * ldu;
        }
        crows = min(BT, nr - rowI); ccols = min(BTN, nr - rowJ);
        lo = nc;
    }
    const int dshift = rowJ - rowI;            // entry (r, c) of the tile is on/below the diagonal iff r >= c + dshift
    const double* A = P + rowI + (long long)k0 * ld;
    const double* B = P + rowJ + (long long)k0 * ld;
    const int brows = ccols;

    // warp tile: 32 C-rows (MMA N dimension, 4 tiles) x 32 C-columns (MMA M dimension, 4 tiles)
    const int wr = (warp >> 1) * 32, wc = (warp & 1) * 32;
    // Accumulators start from C (all loads issued up front, they land while the cp.async prologue runs) and the MMA
    // adds (-A_j)(A_i)^T, so the epilogue is stores only: a load-subtract-store epilogue serialises dependent
    // global round trips (64 % of the stall samples of the first version, profiles/).
    double acc[4][4][2];
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const int c = wc + i * 8 + (lane >> 2);
        const bool cv = c < ccols && rowJ + c >= lo;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const int r = wr + j * 8 + 2 * (lane & 3);
            const double* p = C + (long long)c * ldc + r;
            const bool v0 = cv && r < crows && rowI + r >= lo && r >= c + dshift;
            const bool v1 = cv && r + 1 < crows && rowI + r + 1 >= lo && r + 1 >= c + dshift;
            if (v0 && v1) {
                const double2 x = *reinterpret_cast<const double2*>(p);
                acc[i][j][0] = x.x; acc[i][j][1] = x.y;
            } else {
                acc[i][j][0] = v0 ? p[0] : 0.0;
                acc[i][j][1] = v1 ? p[1] : 0.0;
            }
        }
    }

    const int nkt = (K + BK - 1) / BK;
    __shared__ unsigned long long full_bar[STAGES];
    const unsigned abytes = (unsigned)((crows + 1) & ~1) * 8u, bbytes = (unsigned)((brows + 1) & ~1) * 8u;
    auto issue_tma = [&](int tile, int st) {
        const int kc = min(BK, K - tile * BK);
        if (kc < BK) {                 // k-columns beyond K: zeros (generic stores, ordered by the CTA barrier before they are read)
            for (int idx = tid; idx < (BK - kc) * BT; idx += UPD_THREADS) As[(st * BK + kc + idx / BT) * LDT + idx % BT] = 0.0;
            for (int idx = tid; idx < (BK - kc) * BTN; idx += UPD_THREADS) Bs[(st * BK + kc + idx / BTN) * LDTB + idx % BTN] = 0.0;
        }
        if (warp == 0) {
            if (lane == 0) {
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // generic reads of this stage precede the async writes
                mbar_expect_tx(&full_bar[st], (unsigned)kc * (abytes + bbytes));
            }
            __syncwarp();
            if (lane < kc) {
                bulk_g2s(As + (st * BK + lane) * LDT, A + (long long)(tile * BK + lane) * ld, abytes, &full_bar[st]);
                bulk_g2s(Bs + (st * BK + lane) * LDTB, B + (long long)(tile * BK + lane) * ld, bbytes, &full_bar[st]);
            }
        }
    };
    if (TMA) {
        if (tid == 0) {
#pragma unroll
            for (int s = 0; s < STAGES; s++) mbar_init(&full_bar[s], 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncthreads();
#pragma unroll
        for (int s = 0; s < STAGES - 1; s++) if (s < nkt) issue_tma(s, s);
    } else {
#pragma unroll
        for (int s = 0; s < STAGES - 1; s++) {
            if (s < nkt) {
                load_tile_async<BT, LDT>(As + s * BK * LDT, A + (long long)s * BK * ld, ld, crows, K - s * BK, tid);
                load_tile_async<BTN, LDTB>(Bs + s * BK * LDTB, B + (long long)s * BK * ld, ld, brows, K - s * BK, tid);
            }
            cp_async_commit();
        }
    }
    for (int kt = 0; kt < nkt; kt++) {
        if (TMA) mbar_wait(&full_bar[kt % STAGES], (unsigned)((kt / STAGES) & 1));
        else cp_async_wait<STAGES - 2>();
        __syncthreads();
        const int nk = kt + STAGES - 1;
        if (TMA) {
            if (nk < nkt) issue_tma(nk, nk % STAGES);
        } else {
            if (nk < nkt) {
                const int s = nk % STAGES;
                load_tile_async<BT, LDT>(As + s * BK * LDT, A + (long long)nk * BK * ld, ld, crows, K - nk * BK, tid);
                load_tile_async<BTN, LDTB>(Bs + s * BK * LDTB, B + (long long)nk * BK * ld, ld, brows, K - nk * BK, tid);
            }
            cp_async_commit();
        }
        const double* as = As + (kt % STAGES) * BK * LDT;
        const double* bs = Bs + (kt % STAGES) * BK * LDTB;
        // SGN (A = L S L'): C -= A_i S A_j^T, the column signs are applied to the A_j fragment.  The sign array is padded by
        // BK entries, columns beyond K meet zero-filled operands.
        int flip[BK / 4];                 // sign bit of s_k: XORed into the fragment on the integer pipe, the FP64 pipe is DMMA's
        if (SGN) {
#pragma unroll
            for (int kk = 0; kk < BK; kk += 4) flip[kk / 4] = __double2hiint(sgn[f.col0 + k0 + kt * BK + kk + (lane & 3)]) & 0x80000000;
        }
#pragma unroll
        for (int kk = 0; kk < BK; kk += 4) {
            double am[4], bn[4];
#pragma unroll
            for (int i = 0; i < 4; i++) {
                double v = bs[(kk + (lane & 3)) * LDTB + wc + i * 8 + (lane >> 2)];
                if (SGN) v = __hiloint2double(__double2hiint(v) ^ flip[kk / 4], __double2loint(v));
                am[i] = -v;               // the negation is a DMMA operand modifier
            }
#pragma unroll
            for (int j = 0; j < 4; j++) bn[j] = as[(kk + (lane & 3)) * LDT + wr + j * 8 + (lane >> 2)];
#pragma unroll
            for (int i = 0; i < 4; i++)
#pragma unroll
                for (int j = 0; j < 4; j++) dmma884(acc[i][j][0], acc[i][j][1], am[i], bn[j]);
        }
    }
    if (!TMA) cp_async_wait<0>();
    // epilogue: thread holds C rows (r, r+1) of one column per accumulator pair
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const int c = wc + i * 8 + (lane >> 2);
        if (c >= ccols || rowJ + c < lo) continue;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const int r = wr + j * 8 + 2 * (lane & 3);
            double* p = C + (long long)c * ldc + r;
            const bool v0 = r < crows && rowI + r >= lo && r >= c + dshift;
            const bool v1 = r + 1 < crows && rowI + r + 1 >= lo && r + 1 >= c + dshift;
            if (v0 && v1) *reinterpret_cast<double2*>(p) = make_double2(acc[i][j][0], acc[i][j][1]);
            else if (v0) p[0] = acc[i][j][0];
            else if (v1) p[1] = acc[i][j][1];
        }
    }
}

// ---------------------------------------------------------------------------------------------------
// K7: supernodal triangular solves, one CTA per (front, right-hand side), level by level.
// Work vector of a front: T[rowptr .. rowptr+nr) (first nc entries = solution block, rest = the update
// vector handed to the parent through the same relative indices as the factorization).
// ---------------------------------------------------------------------------------------------------
constexpr int CB = 32;
// THREADS = 256, or 32 for the fronts with <= 32 rows (one warp per front, 32 fronts resident per SM instead of 8: the leaf
// levels of a KKT matrix hold hundreds of thousands of two- and three-row fronts); same arithmetic per entry either way.
template <int THREADS>
__global__ void __launch_bounds__(THREADS) k_fwd(const int* __restrict__ list, const FrontD* __restrict__ F,
                                             const int* __restrict__ child_idx, const int* __restrict__ rel,
                                             const double* __restrict__ L, double* __restrict__ T, long long tstride,
                                             double* __restrict__ X, long long xstride,
                                             const unsigned char* __restrict__ owned = nullptr) {
    __shared__ double D[CB][CB + 1];
    __shared__ double xs[CB];
    if (owned && !owned[list[blockIdx.x]]) return;       // distributed solves: another GPU's front
    const FrontD f = F[list[blockIdx.x]];
    double* t = T + blockIdx.y * tstride + f.rowptr;
    double* x = X + blockIdx.y * xstride + f.col0;
    const double* P = L + f.loff;
    const int nr = f.nr, nc = f.nc, ld = f.ld, tid = threadIdx.x, lane = tid & 31;
    for (int i = tid; i < nr; i += THREADS) t[i] = (i < nc) ? x[i] : 0.0;
    __syncthreads();
    for (int q = 0; q < f.nchild; q++) {
        const FrontD fc = F[child_idx[f.childptr + q]];
        const int mc = fc.nr - fc.nc;
        const int* rl = rel + fc.reloff;
        const double* tc = T + blockIdx.y * tstride + fc.rowptr + fc.nc;
        for (int i = tid; i < mc; i += THREADS) t[rl[i]] += tc[i];
        __syncthreads();
    }
    for (int b0 = 0; b0 < nc; b0 += CB) {
        const int wb = min(CB, nc - b0);
        for (int idx = tid; idx < wb * wb; idx += THREADS) {
            int c = idx / wb, r = idx - c * wb;
            D[r][c] = (r >= c) ? P[(long long)(b0 + c) * ld + b0 + r] : 0.0;
        }
        __syncthreads();
        if (tid < 32) {
            double v = (lane < wb) ? t[b0 + lane] : 0.0;
            const double rd = (lane < wb) ? 1.0 / D[lane][lane] : 1.0;      // one division per lane, off the 32-step chain
            for (int qq = 0; qq < wb; qq++) {
                double xq = __shfl_sync(0xffffffffu, v, qq) * __shfl_sync(0xffffffffu, rd, qq);
                if (lane == qq) v = xq;
                else if (lane > qq && lane < wb) v -= xq * D[lane][qq];
            }
            if (lane < wb) { xs[lane] = v; t[b0 + lane] = v; x[b0 + lane] = v; }
        }
        __syncthreads();
        for (int r = b0 + wb + tid; r < nr; r += THREADS) {
            double acc = t[r];
            const double* col = P + (long long)b0 * ld + r;
            for (int qq = 0; qq < wb; qq++) acc -= col[(long long)qq * ld] * xs[qq];
            t[r] = acc;
        }
        __syncthreads();
    }
}

template <int THREADS>
__global__ void __launch_bounds__(THREADS) k_bwd(const int* __restrict__ list, const FrontD* __restrict__ F,
                                             const int* __restrict__ rows, const double* __restrict__ L,
                                             double* __restrict__ T, long long tstride, double* __restrict__ X,
                                             long long xstride, const unsigned char* __restrict__ owned = nullptr) {
    __shared__ double D[CB][CB + 1];
    __shared__ double zs[CB];
    if (owned && !owned[list[blockIdx.x]]) return;
    const FrontD f = F[list[blockIdx.x]];
    double* t = T + blockIdx.y * tstride + f.rowptr;
    double* xg = X + blockIdx.y * xstride;
    const double* P = L + f.loff;
    const int* rw = rows + f.rowptr;
    const int nr = f.nr, nc = f.nc, ld = f.ld, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < nr; i += THREADS) t[i] = xg[rw[i]];
    __syncthreads();
    const int nblk = (nc + CB - 1) / CB;
    for (int b = nblk - 1; b >= 0; b--) {
        const int b0 = b * CB, wb = min(CB, nc - b0);
        for (int idx = tid; idx < wb * wb; idx += THREADS) {
            int c = idx / wb, r = idx - c * wb;
            D[r][c] = (r >= c) ? P[(long long)(b0 + c) * ld + b0 + r] : 0.0;
        }
        for (int qq = warp; qq < wb; qq += THREADS / 32) {
            const double* col = P + (long long)(b0 + qq) * ld;
            double s = 0.0;
            for (int r = b0 + wb + lane; r < nr; r += 32) s += col[r] * t[r];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
            if (lane == 0) zs[qq] = t[b0 + qq] - s;
        }
        __syncthreads();
        if (tid < 32) {
            double v = (lane < wb) ? zs[lane] : 0.0;
            const double rd = (lane < wb) ? 1.0 / D[lane][lane] : 1.0;
            for (int qq = wb - 1; qq >= 0; qq--) {
                double xq = __shfl_sync(0xffffffffu, v, qq) * __shfl_sync(0xffffffffu, rd, qq);
                if (lane == qq) v = xq;
                else if (lane < qq) v -= xq * D[qq][lane];
            }
            if (lane < wb) { t[b0 + lane] = v; xg[f.col0 + b0 + lane] = v; }
        }
        __syncthreads();
    }
}

// ---- large fronts: the same solves split over many CTAs, one 128-column block step at a time ------------------
// The 32 x 32 diagonal sub-blocks of every 128-column block are inverted once per factorization (k_diag_inverse), so
// a block step is four small matrix-vector products instead of a 128-step substitution chain.
constexpr int SOLVE_FT = 64;     // rows per CTA in the forward update  (256 threads = 64 rows x 4 column quarters)
constexpr int SOLVE_BT = 128;    // rows per CTA in the backward (transposed) update: the two 64-row slices of an ABSOLUTE 128-row pair
constexpr int SB = 32;           // inverted diagonal sub-block
constexpr int MINV_HALF = (NB / SB) * SB * SB;  // the four inverse sub-blocks of a 128-column block, column-major
constexpr int MINV_BLK = 2 * MINV_HALF;         // ... followed by their transposes (backward solve)
constexpr int LDM = SB;          // smem stride of a staged inverse sub-block (always read column-wise)
// Staged diagonal block, PACKED: only the part below the inverted 32 x 32 sub-blocks is ever read -- for the 32 columns of
// sub-block s the rows (s + 1) * 32 .. 127.  Column strides 104 / 72 / 40 (= rows + 8: stride % 16 == 8 keeps the
// 8-rows-per-column access of the backward block solve at the two-wavefront minimum; even: 16-byte LDGSTS rows).
constexpr int DPK_R0 = 104, DPK_R1 = 72, DPK_R2 = 40;
constexpr int DPK_B1 = 32 * DPK_R0, DPK_B2 = DPK_B1 + 32 * DPK_R1, DIAG_PACK = DPK_B2 + 32 * DPK_R2;
__device__ __forceinline__ int dpk_base(int s) { return s == 0 ? 0 : (s == 1 ? DPK_B1 : DPK_B2); }
__device__ __forceinline__ int dpk_rows(int s) { return s == 0 ? DPK_R0 : (s == 1 ? DPK_R1 : DPK_R2); }
// element (row r, column c) of the block, r >= ((c >> 5) + 1) * 32
__device__ __forceinline__ int dpk(int c, int r) { const int s = c >> 5; return dpk_base(s) + (c & 31) * dpk_rows(s) + (r - 32 * (s + 1)); }
static constexpr size_t SMEM_SDIAG = (size_t)(DIAG_PACK + MINV_HALF + NB) * sizeof(double);

// inverse of the SB x SB lower-triangular diagonal sub-blocks: one CTA per 128-column block, one warp per sub-block,
// lane j solves L x = e_j by substitution
__global__ void __launch_bounds__(128) k_diag_inverse(const int* __restrict__ blkfront, const int* __restrict__ blkkb,
                                                      const FrontD* __restrict__ F, const double* __restrict__ L,
                                                      double* __restrict__ Minv) {
    __shared__ double D[NB / SB][SB][SB + 1];
    const FrontD f = F[blkfront[blockIdx.x]];
    const int kb = blkkb[blockIdx.x], k0 = kb * NB, w = min(NB, f.nc - k0);
    const int s = threadIdx.x >> 5, lane = threadIdx.x & 31, b0 = s * SB, wb = min(SB, w - b0);
    if (wb <= 0) return;
    const double* P = L + f.loff + (long long)(k0 + b0) * f.ld + k0 + b0;
    for (int c = 0; c < SB; c++) {
        double v = (lane == c) ? 1.0 : 0.0;
        if (c < wb && lane < wb && lane >= c) v = P[(long long)c * f.ld + lane];
        D[s][lane][c] = v;
    }
    __syncwarp();
    double x[SB];
#pragma unroll
    for (int i = 0; i < SB; i++) {
        double acc = (i == lane) ? 1.0 : 0.0;
#pragma unroll
        for (int k = 0; k < i; k++) acc = fma(-D[s][i][k], x[k], acc);
        x[i] = acc / D[s][i][i];
    }
    __syncwarp();
#pragma unroll
    for (int i = 0; i < SB; i++) D[s][i][lane] = x[i];      // element (i, lane) of the inverse
    __syncwarp();
    double* out = Minv + f.ioff + (long long)kb * MINV_BLK + s * SB * SB;
    for (int c = 0; c < SB; c++) { out[c * SB + lane] = D[s][lane][c]; out[MINV_HALF + c * SB + lane] = D[s][c][lane]; }
}

// forward: t = [x(cols); 0] + children's update vectors.  One CTA per GATHER_ROWS-row chunk of a front: a child's relative
// indices ascend, so the entries that fall into the chunk are a contiguous run found by binary search; children are
// applied one after the other (deterministic sums).
constexpr int GATHER_SINGLE = 0x40000000;   // chunk-child entry: the child has exactly one update row
constexpr int GATHER_ROWS = 512;       // (2048: the top levels ran 8-40 CTAs of 8 dependent rounds per child, ~30 us per launch)
__global__ void __launch_bounds__(256) k_fwd_gather(const int* __restrict__ list, const int* __restrict__ cprefix, int nfronts,
                                                    const FrontD* __restrict__ F,
                                                    const int* __restrict__ child_idx, const int* __restrict__ rel,
                                                    double* __restrict__ T, long long tstride, const double* __restrict__ X,
                                                    long long xstride, const int2* __restrict__ chunk_q, const int* __restrict__ chunk_child,
                                                    const unsigned char* __restrict__ owned = nullptr) {
    const int g = find_group(cprefix, nfronts, blockIdx.x);
    if (owned && !owned[list[g]]) return;
    const FrontD f = F[list[g]];
    const int a = (blockIdx.x - cprefix[g]) * GATHER_ROWS, b = min(f.nr, a + GATHER_ROWS);
    double* t = T + blockIdx.y * tstride + f.rowptr;
    const double* x = X + blockIdx.y * xstride + f.col0;
    const int nc = f.nc, tid = threadIdx.x;
    for (int i = a + tid; i < b; i += 256) t[i] = (i < nc) ? x[i] : 0.0;
    __syncthreads();
    // chunk_q[CTA] = {first, count} in chunk_child: the positions (ascending) of the children with an update row in [a, b) -- the
    // root front of a 3 x 3 KKT matrix has thousands of one-row children, each of which concerns one chunk.  Children with ONE
    // update row carry GATHER_SINGLE: a run of up to 256 of them is fetched by 256 threads at once (one latency instead of one
    // per child and no barrier per child), then every thread adds, in child order, the values that land on the two rows it
    // owns; the other children are added by the whole CTA, one after the other.  Sums stay in child order either way.
    __shared__ int s_first, s_code, s_row[256];
    __shared__ double s_val[256];
    const int2 qc = chunk_q[blockIdx.x];
    int t_ = 0;
    while (t_ < qc.y) {
        if (tid == 0) s_first = 256;
        __syncthreads();
        const int myq = t_ + tid;
        const int code = myq < qc.y ? chunk_child[qc.x + myq] : 0;
        if (!(myq < qc.y && (code & GATHER_SINGLE))) atomicMin(&s_first, tid);
        if (tid == 0) s_code = code;
        __syncthreads();
        const int nrun = s_first;
        if (nrun > 0) {
            if (tid < nrun) {
                const FrontD fc = F[child_idx[f.childptr + (code & ~GATHER_SINGLE)]];
                s_row[tid] = rel[fc.reloff];
                s_val[tid] = T[blockIdx.y * tstride + fc.rowptr + fc.nc];
            }
            __syncthreads();
            for (int k = 0; k < nrun; k++) {
                const int d = s_row[k];
                if (((d - a) & 255) == tid) t[d] += s_val[k];
            }
            t_ += nrun;
        } else {
            const FrontD fc = F[child_idx[f.childptr + (s_code & ~GATHER_SINGLE)]];
            const int mc = fc.nr - fc.nc;
            const int* rl = rel + fc.reloff;
            const double* tc = T + blockIdx.y * tstride + fc.rowptr + fc.nc;
            int lo = 0, hi = mc;                       // first entry with rl >= a
            while (lo < hi) { const int mid = (lo + hi) >> 1; if (rl[mid] < a) lo = mid + 1; else hi = mid; }
            for (int i = lo + tid; i < mc; i += 256) {
                const int d = rl[i];
                if (d >= b) break;
                t[d] += tc[i];
            }
            t_ += 1;
            __syncthreads();        // s_first / s_code are rewritten at the top of the loop
        }
    }
}
// backward: t = x(rows of the front)
__global__ void __launch_bounds__(256) k_bwd_gather(const int* __restrict__ list, const FrontD* __restrict__ F,
                                                    const int* __restrict__ rows, double* __restrict__ T, long long tstride,
                                                    const double* __restrict__ X, long long xstride,
                                                    const unsigned char* __restrict__ owned = nullptr) {
    if (owned && !owned[list[blockIdx.x]]) return;
    const FrontD f = F[list[blockIdx.x]];
    double* t = T + blockIdx.y * tstride + f.rowptr;
    const double* xg = X + blockIdx.y * xstride;
    const int* rw = rows + f.rowptr;
    for (int i = threadIdx.x; i < f.nr; i += 256) t[i] = xg[rw[i]];
}

// stage the inverse sub-blocks (or their transposes) and the part of the 128-block below them in shared memory:
// 16-byte LDGSTS, all in flight.  Row pairs may reach one row past w (w odd): inside the padded panel, never used.
__device__ __forceinline__ void cp_async16(unsigned dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src));
}
__device__ __forceinline__ void stage_diag_issue(const double* __restrict__ P, int ld, int k0, int w,
                                                 const double* __restrict__ minv, double* Ls, double* Ms, int tid) {
    const unsigned lbase = (unsigned)__cvta_generic_to_shared(Ls), mbase = (unsigned)__cvta_generic_to_shared(Ms);
    const int nm = min(NB / SB, (w + SB - 1) / SB) * SB * SB;
    for (int i = tid * 2; i < nm; i += 512) cp_async16(mbase + 8u * (unsigned)i, minv + i);
    for (int idx = tid; idx < NB * (NB / 2); idx += 256) {
        const int c = idx >> 6, r = (idx & 63) * 2;
        if (c < w && r < w && r >= ((c >> 5) + 1) * SB) cp_async16(lbase + 8u * (unsigned)dpk(c, r), P + (long long)(k0 + c) * ld + k0 + r);
    }
}
__device__ __forceinline__ void stage_diag_block(const double* __restrict__ P, int ld, int k0, int w,
                                                 const double* __restrict__ minv, double* Ls, double* Ms, int tid) {
    stage_diag_issue(P, ld, k0, w, minv, Ls, Ms, tid);
    asm volatile("cp.async.commit_group;\n cp.async.wait_group 0;" ::: "memory");
}
// Forward block solve out of shared memory: ts[0..w) <- inv(L11) ts.  All 256 threads; ends with a CTA barrier.
__device__ __forceinline__ void fwd_diag_core(const double* Ls, const double* Ms, double* ts, int w, int tid) {
    const int lane = tid & 31;
    for (int b0 = 0; b0 < w; b0 += SB) {
        const double* M = Ms + (b0 / SB) * SB * SB;
        double xv = 0.0;
        if (tid < 32) {                    // x_sub = inv(L_sub) t_sub
            double a0 = 0, a1 = 0;
#pragma unroll
            for (int c = 0; c < SB; c += 2) { a0 = fma(M[c * LDM + lane], ts[b0 + c], a0); a1 = fma(M[(c + 1) * LDM + lane], ts[b0 + c + 1], a1); }
            xv = a0 + a1;
        }
        __syncthreads();
        if (tid < 32) ts[b0 + lane] = xv;
        __syncthreads();
        const int r = b0 + SB + tid;       // rows of this 128-block below the sub-block
        if (r < w) {
            const int s = b0 / SB, rows = dpk_rows(s);
            const double* lc = Ls + dpk_base(s) + tid;      // (row r, column b0): r - 32 (s + 1) = tid
            double a0 = 0, a1 = 0;
#pragma unroll
            for (int c = 0; c < SB; c += 2) { a0 = fma(lc[c * rows], ts[b0 + c], a0); a1 = fma(lc[(c + 1) * rows], ts[b0 + c + 1], a1); }
            ts[r] -= a0 + a1;
        }
        __syncthreads();
    }
}
// Backward block solve: zs[0..w) <- inv(L11)^T zs (Ms holds the transposed inverses).  All 256 threads; ends with a CTA barrier.
__device__ __forceinline__ void bwd_diag_core(const double* Ls, const double* Ms, double* zs, int w, int tid) {
    const int lane = tid & 31;
    const int nsb = (w + SB - 1) / SB;
    for (int sbk = nsb - 1; sbk >= 0; sbk--) {
        const int b0 = sbk * SB;
        // contributions of the already solved rows of this 128-block (below the sub-block): 8 threads per column
        {
            const int q = tid >> 3, gq = tid & 7;
            double sv = 0.0;
            if (sbk < NB / SB - 1) {
                const double* lc = Ls + dpk_base(sbk) + q * dpk_rows(sbk) - (b0 + SB);     // + r: element (row r, column b0 + q)
                for (int r = b0 + SB + gq; r < w; r += 8) sv = fma(lc[r], zs[r], sv);
            }
            sv += __shfl_xor_sync(0xffffffffu, sv, 1);
            sv += __shfl_xor_sync(0xffffffffu, sv, 2);
            sv += __shfl_xor_sync(0xffffffffu, sv, 4);
            if (gq == 0) zs[b0 + q] -= sv;
        }
        __syncthreads();
        double xv = 0.0;
        if (tid < 32) {                    // x_sub = inv(L_sub)^T z_sub
            const double* M = Ms + sbk * SB * SB;          // transposed inverse, column-major
            double a0 = 0, a1 = 0;
#pragma unroll
            for (int c = 0; c < SB; c += 2) { a0 = fma(M[c * LDM + lane], zs[b0 + c], a0); a1 = fma(M[(c + 1) * LDM + lane], zs[b0 + c + 1], a1); }
            xv = a0 + a1;
        }
        __syncthreads();
        if (tid < 32) zs[b0 + lane] = xv;
        __syncthreads();
    }
}

// forward, block kb: solve the (<=128)^2 diagonal block for t[k0..k0+w)
__global__ void __launch_bounds__(256) k_fwd_diag(const __grid_constant__ SolveGroups sg, const int* __restrict__ gfront, int kb, const FrontD* __restrict__ F,
                                                  const double* __restrict__ L, const double* __restrict__ Minv,
                                                  double* __restrict__ T, long long tstride,
                                                  double* __restrict__ X, long long xstride,
                                                  const unsigned char* __restrict__ owned = nullptr) {
    extern __shared__ double sm[];
    double* Ls = sm;                       // packed part of the block below the inverted sub-blocks (dpk)
    double* Ms = Ls + DIAG_PACK;           // [sub-block][col][row], stride LDM
    double* ts = Ms + MINV_HALF;
    const FrontS f = load_front(sg, gfront, F, blockIdx.x);
    if (owned && !owned[f.id]) return;
    double* t = T + blockIdx.y * tstride + f.rowptr;
    double* x = X + blockIdx.y * xstride + f.col0;
    const int tid = threadIdx.x;
    const int k0 = kb * NB, w = min(NB, f.nc - k0);
    stage_diag_block(L + f.loff, f.ld, k0, w, Minv + f.ioff + (long long)kb * MINV_BLK, Ls, Ms, tid);
    if (tid < NB) ts[tid] = (tid < w) ? t[k0 + tid] : 0.0;
    __syncthreads();
    fwd_diag_core(Ls, Ms, ts, w, tid);
    if (tid < w) { t[k0 + tid] = ts[tid]; x[k0 + tid] = ts[tid]; }
}
// stage a 64-row x w-column slice of the panel (rows r0.. with r0 even, columns k0..) in shared memory as S[col][64]:
// thread = (row pair, group of 16 columns) issues its 16 16-byte LDGSTS back to back, so the whole 64 KB slice is in
// flight at once.  A pair may reach row nr (nr odd): inside the padded panel (ld even), masked by the callers.
__device__ __forceinline__ void stage_rows64(const double* __restrict__ P, int ld, int nr, int k0, int w, int r0, double* S, int tid) {
    const int rp = tid & 31, cg = tid >> 5, r = r0 + 2 * rp;
    if (r >= nr) return;
    const int wq = min(16, w - cg * 16);
    const double* col = P + (long long)(k0 + cg * 16) * ld + r;
    const unsigned sb = (unsigned)__cvta_generic_to_shared(S) + 8u * (unsigned)(cg * 16 * 64 + 2 * rp);
#pragma unroll 8
    for (int j = 0; j < wq; j++) cp_async16(sb + 8u * 64u * (unsigned)j, col + (long long)j * ld);
}

// forward, block kb: t[r] -= L[r, blk] * x_blk for a 64-row tile of the rows below the block.  The tile of L is staged
// once and applied to all right-hand sides of the call, CG columns at a time (one shared-memory read feeds CG sums).
template <int CG>
__global__ void __launch_bounds__(256) k_fwd_upd(const __grid_constant__ SolveGroups sg, const int* __restrict__ gfront, const int* __restrict__ gprefix, int ngroups,
                                                 int kb, const FrontD* __restrict__ F, const double* __restrict__ L,
                                                 double* __restrict__ T, long long tstride, int ncols,
                                                 const unsigned char* __restrict__ owned = nullptr) {
    extern __shared__ double sm[];
    double* S = sm;                        // [128 columns][64 rows]
    __shared__ double xs[CG][NB];
    __shared__ double red[CG][4][SOLVE_FT];
    int tile;
    const int g = locate_group(sg, gprefix, ngroups, blockIdx.x, tile);
    const FrontS f = load_front(sg, gfront, F, g);
    if (owned && !owned[f.id]) return;
    const int k0 = kb * NB, w = min(NB, f.nc - k0), tid = threadIdx.x;
    const int rr = tid & (SOLVE_FT - 1), cq = tid >> 6;
    const int rb = k0 + w, r0 = (rb & ~1) + tile * SOLVE_FT;      // tiles start at an even row (16-byte LDGSTS)
    stage_rows64(L + f.loff, f.ld, f.nr, k0, w, r0, S, tid);
    asm volatile("cp.async.commit_group;" ::: "memory");
    const bool rowok = r0 + rr < f.nr && r0 + rr >= rb;
    const int wq = min(32, w - cq * 32);
    const double* sp = S + cq * 32 * 64 + rr;
    for (int c0 = 0; c0 < ncols; c0 += CG) {
        const int cn = min(CG, ncols - c0);
        if (c0) __syncthreads();           // xs / red of the previous column group are consumed
        for (int i = tid; i < CG * NB; i += 256) {
            const int c = i / NB, q = i - c * NB;
            xs[c][q] = (c < cn && q < w) ? T[(c0 + c) * tstride + f.rowptr + k0 + q] : 0.0;
        }
        if (c0 == 0) asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncthreads();
        double a0[CG], a1[CG];
#pragma unroll
        for (int c = 0; c < CG; c++) a0[c] = a1[c] = 0.0;
        if (rowok) {
            int j = 0;
            for (; j + 1 < wq; j += 2) {
                const double s0 = sp[j * 64], s1 = sp[(j + 1) * 64];
#pragma unroll
                for (int c = 0; c < CG; c++) { a0[c] = fma(s0, xs[c][cq * 32 + j], a0[c]); a1[c] = fma(s1, xs[c][cq * 32 + j + 1], a1[c]); }
            }
            if (j < wq) {
                const double s0 = sp[j * 64];
#pragma unroll
                for (int c = 0; c < CG; c++) a0[c] = fma(s0, xs[c][cq * 32 + j], a0[c]);
            }
        }
#pragma unroll
        for (int c = 0; c < CG; c++) red[c][cq][rr] = a0[c] + a1[c];
        __syncthreads();
        for (int i = tid; i < cn * SOLVE_FT; i += 256) {
            const int c = i / SOLVE_FT, r = i - c * SOLVE_FT;
            if (r0 + r < f.nr && r0 + r >= rb)
                T[(c0 + c) * tstride + f.rowptr + r0 + r] -= (red[c][0][r] + red[c][1][r]) + (red[c][2][r] + red[c][3][r]);
        }
    }
}
// backward, block kb: partial[q] = sum over the rows r >= rb of ONE 128-row pair (rows 128 p .. 128 p + 127, p = rb / 128 +
// tile: the same absolute pairs as the forward sweep) of L[r, k0+q] * t[r].  Two 64-row slices, one shared-memory buffer each
// (128 KB in flight); thread = (row, column quarter) accumulates slice 1 then slice 0; the 32 per-lane column sums of a warp are
// combined by a halving butterfly.  CG right-hand sides share the staged tile (blockIdx.y = group of CG columns).
// column sums of one pair: red[c][2 cq .. 2 cq + 1][lane] hold, after the closing barrier, the two half sums of column
// cq * 32 + lane of right-hand side c (added by the callers)
template <int CG>
__device__ __forceinline__ void bwd_pair_sums(const FrontS& f, int kb, int r0, const double* __restrict__ L, const double* __restrict__ t,
                                              long long tstride, int cn, double* sm, double (&red)[CG][8][32]) {
    const double* P = L + f.loff;
    const int k0 = kb * NB, w = min(NB, f.nc - k0), tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int rr = tid & 63, cq = tid >> 6;
    const int wq = min(32, w - cq * 32);
    const int rb = k0 + w;
    constexpr int NSUB = SOLVE_BT / 64;
    double p[CG][32];
#pragma unroll
    for (int c = 0; c < CG; c++)
#pragma unroll
        for (int j = 0; j < 32; j++) p[c][j] = 0.0;
    static_assert(NSUB == 2, "one buffer per slice");
    double tv[CG][NSUB];
    // slices in DESCENDING row order (1, 0): the order of the persistent sweep
#pragma unroll
    for (int s_ = 0; s_ < NSUB; s_++) {
        const int sub = NSUB - 1 - s_;
        stage_rows64(P, f.ld, f.nr, k0, w, r0 + sub * 64, sm + sub * NB * 64, tid);
        asm volatile("cp.async.commit_group;" ::: "memory");
        const int r = r0 + sub * 64 + rr;
#pragma unroll
        for (int c = 0; c < CG; c++) tv[c][sub] = (c < cn && r < f.nr && r >= rb) ? t[c * tstride + r] : 0.0;
    }
#pragma unroll
    for (int s_ = 0; s_ < NSUB; s_++) {
        const int sub = NSUB - 1 - s_;
        if (s_ == 0) asm volatile("cp.async.wait_group 1;" ::: "memory");
        if (s_ == 1) asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncthreads();
        const int r = r0 + sub * 64 + rr;
        if (r < f.nr && r >= rb) {
            const double* sp = sm + sub * NB * 64 + cq * 32 * 64 + rr;
#pragma unroll
            for (int j = 0; j < 32; j++)
                if (j < wq) {
                    const double lv = sp[j * 64];
#pragma unroll
                    for (int c = 0; c < CG; c++) p[c][j] = fma(lv, tv[c][sub], p[c][j]);
                }
        }
    }
    // after the step with offset o the lanes with bit o set hold the upper half of the surviving columns: lane l ends
    // with the sum of column l
#pragma unroll
    for (int c = 0; c < CG; c++) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const bool up = (lane & o) != 0;
#pragma unroll
            for (int j = 0; j < o; j++) {
                const double send = up ? p[c][j] : p[c][j + o];
                const double keep = up ? p[c][j + o] : p[c][j];
                p[c][j] = keep + __shfl_xor_sync(0xffffffffu, send, o);
            }
        }
        red[c][warp][lane] = p[c][0];
    }
    __syncthreads();
}
template <int CG>
__global__ void __launch_bounds__(256) k_bwd_upd(const __grid_constant__ SolveGroups sg, const int* __restrict__ gfront, const int* __restrict__ gprefix, int ngroups,
                                                 int kb, const FrontD* __restrict__ F, const double* __restrict__ L,
                                                 const double* __restrict__ T, long long tstride, double* __restrict__ part,
                                                 long long pstride, int ncols, const unsigned char* __restrict__ owned = nullptr) {
    extern __shared__ double sm[];         // 2 buffers of [128 columns][64 rows]
    __shared__ double red[CG][8][32];
    int tile;
    const int g = locate_group(sg, gprefix, ngroups, blockIdx.x, tile);
    const FrontS f = load_front(sg, gfront, F, g);
    if (owned && !owned[f.id]) return;
    const int c0 = blockIdx.y * CG, cn = min(CG, ncols - c0), tid = threadIdx.x;
    const int rb = min(f.nc, (kb + 1) * NB);
    bwd_pair_sums<CG>(f, kb, (rb / NB + tile) * SOLVE_BT, L, T + c0 * tstride + f.rowptr, tstride, cn, sm, red);
    for (int i = tid; i < cn * NB; i += 256) {
        const int c = i / NB, q = i - c * NB;
        part[(c0 + c) * pstride + (long long)blockIdx.x * NB + q] = red[c][2 * (q >> 5)][q & 31] + red[c][2 * (q >> 5) + 1][q & 31];
    }
}
// backward, block kb: z = t_blk - sum of the tile partials (fixed order), then solve L11^T x = z
constexpr int BWD_CG = 4;        // right-hand sides per CTA in the backward update (multi-column solves)
constexpr int PT_CHUNK = 48;     // partial rows staged per pass
static constexpr size_t SMEM_FUPD = (size_t)NB * 64 * sizeof(double), SMEM_BUPD = 2 * SMEM_FUPD;
static constexpr size_t SMEM_BDIAG = SMEM_SDIAG + (size_t)PT_CHUNK * NB * sizeof(double);
__global__ void __launch_bounds__(256) k_bwd_diag(const __grid_constant__ SolveGroups sg, const int* __restrict__ gfront, const int* __restrict__ gprefix, int kb,
                                                  const FrontD* __restrict__ F, const double* __restrict__ L,
                                                  const double* __restrict__ Minv,
                                                  double* __restrict__ T, long long tstride, double* __restrict__ X,
                                                  long long xstride, const double* __restrict__ part, long long pstride,
                                                  const unsigned char* __restrict__ owned = nullptr) {
    extern __shared__ double sm[];
    double* Ls = sm;
    double* Ms = Ls + DIAG_PACK;
    double* zs = Ms + MINV_HALF;
    double* Ps = zs + NB;                  // [PT_CHUNK][128] staged partial sums
    const FrontS f = load_front(sg, gfront, F, blockIdx.x);
    if (owned && !owned[f.id]) return;
    double* t = T + blockIdx.y * tstride + f.rowptr;
    double* x = X + blockIdx.y * xstride + f.col0;
    const int tid = threadIdx.x;
    const int k0 = kb * NB, w = min(NB, f.nc - k0);
    const int tile0 = sg.ng ? sg.prefix[blockIdx.x] : gprefix[blockIdx.x], tile1 = sg.ng ? sg.prefix[blockIdx.x + 1] : gprefix[blockIdx.x + 1];
    const double* pp = part + blockIdx.y * pstride + (long long)tile0 * NB;
    const unsigned pbase = (unsigned)__cvta_generic_to_shared(Ps);
    const int nt = tile1 - tile0;
    {   // the partials are subtracted one after the other from the LAST pair down (the order in which the persistent sweep
        // meets them: the rows below the pivots first, then block after block as they are solved): top chunk first
        const int c0 = max(0, nt - PT_CHUNK);
        for (int i = tid * 2; i < (nt - c0) * NB; i += 512)
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(pbase + 8u * (unsigned)i), "l"(pp + (long long)c0 * NB + i));
    }
    stage_diag_block(L + f.loff, f.ld, k0, w, Minv + f.ioff + (long long)kb * MINV_BLK + MINV_HALF, Ls, Ms, tid);   // commits + waits for all
    __syncthreads();
    {
        double z = (tid < w) ? t[k0 + tid] : 0.0;
        for (int c1 = nt; c1 > 0; c1 -= PT_CHUNK) {
            const int c0 = max(0, c1 - PT_CHUNK);
            if (c1 < nt) {
                __syncthreads();
                for (int i = tid * 2; i < (c1 - c0) * NB; i += 512)
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(pbase + 8u * (unsigned)i), "l"(pp + (long long)c0 * NB + i));
                asm volatile("cp.async.commit_group;\n cp.async.wait_group 0;" ::: "memory");
                __syncthreads();
            }
            if (tid < NB)
                for (int tl = c1 - 1; tl >= c0; tl--) z -= Ps[(tl - c0) * NB + tid];
        }
        if (tid < NB) zs[tid] = (tid < w) ? z : 0.0;
    }
    __syncthreads();
    bwd_diag_core(Ls, Ms, zs, w, tid);
    if (tid < w) { t[k0 + tid] = zs[tid]; x[k0 + tid] = zs[tid]; }
}

// ---- persistent sweeps over the large fronts of one level (one right-hand side) -----------------------------------------
// The level-by-level launches above pay one launch + one dependent load chain per 128-column block step (two kernels per
// step, ~230 steps per sweep of the 100^3 factor: 0.25 of the HBM roofline).  Here ONE kernel per level runs all block steps
// of all its large fronts: the CTAs of a front (one per SM) hand the solved block from one to the next through flags in
// global memory (release / acquire), and stream their slices of the panel through a two-stage LDGSTS ring whose loads do not
// depend on the solve chain -- so the chain per block step is one L2 round trip plus the block solve, and the panel is read at
// memory speed underneath it.  Arithmetic and summation order are exactly those of k_fwd_diag / k_fwd_upd / k_bwd_upd /
// k_bwd_diag (same device functions, same thread mapping): the results are bit-identical to the launch-per-step path, which
// remains for several right-hand sides and for the ownership-masked distributed solves.
struct PersistFront { FrontS f; int cta0, ncta, sync0, rect0; };     // first CTA, CTAs, first flag, first tile of the rectangle pre-pass
// A level's launch: `nf` records in global memory (a persistent kernel starts once per level: one dependent load is nothing).
// nf <= CTAs: front i runs on the CTAs cta0 .. cta0 + ncta - 1 (`cmap` gives the front of every CTA); more fronts than SMs:
// every front runs on ONE CTA (ncta == 1), CTA b takes the fronts b, b + gridDim.x, ...
struct PersistLevel { const PersistFront* fr; const int* cmap; int nf, shared; };
static constexpr size_t SMEM_PERSIST = (size_t)(2 * NB * 64 + DIAG_PACK + MINV_HALF + 4 * NB + 256) * sizeof(double);
constexpr unsigned PERSIST_SPIN_LIMIT = 1u << 25;      // polls before a wait gives up and reports (never reached by a correct schedule)

__device__ __forceinline__ int ld_acquire(const int* p) {
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release(int* p, int v) { asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
// wait until *p >= need; false when the wait was abandoned (error word set: the host reports ST_CUDA instead of hanging)
__device__ __forceinline__ bool persist_wait(const int* p, int need, int* err) {
    unsigned spins = 0;
    while (ld_acquire(p) < need) {
        if ((++spins & 1023u) == 0 && (spins > PERSIST_SPIN_LIMIT || *(volatile int*)err)) { *(volatile int*)err = 1; return false; }
    }
    return true;
}

// Static work lists of the persistent sweeps (host + device: persist_schedule_check replays them on the CPU).
// Forward: item = (block kb, 64-row tile m), m in a 128-row pair owned by CTA c (pair p -> CTA p % G), rows at or below
// rb(kb) = min(nc, (kb + 1) * 128); per CTA in (kb, m) ascending order.
struct FwdSched {
    int nr, nc, c, G;
    __host__ __device__ int nblk() const { return (nc + NB - 1) / NB; }
    __host__ __device__ int ntile() const { return (nr + SOLVE_FT - 1) / SOLVE_FT; }
    __host__ __device__ int rb_of(int kb) const { return nc < (kb + 1) * NB ? nc : (kb + 1) * NB; }
    __host__ __device__ bool first_tile(int kb, int& m) const {      // first item of block kb at or after tile m; false: none
        const int rb = rb_of(kb), pmin = rb / NB;
        int pr = (m >> 1) > pmin ? (m >> 1) : pmin;
        pr += ((c - pr) % G + G) % G;            // next pair owned by this CTA
        if (m < 2 * pr) m = 2 * pr;
        if (m * SOLVE_FT + SOLVE_FT - 1 < rb) m++;       // first tile of the pair entirely above rb (last block only)
        return m < ntile();
    }
    __host__ __device__ bool next_item(int& kb, int& m) const {      // advances (kb, m) to the item after it; false: done
        if ((m & 1) == 0 && m + 1 < ntile()) { m++; return true; }
        m = (m | 1) + 1 + 2 * (G - 1);           // first tile of the next owned pair
        const int npairs = (nr + NB - 1) / NB;
        const int pmax = c < npairs ? c + ((npairs - 1 - c) / G) * G : -1;     // last pair of this CTA
        while (true) {
            if (m < ntile() && first_tile(kb, m)) return true;
            // (the first pair below a block only moves down: once it is past pmax the CTA is done -- no scan over the
            // remaining blocks, which sat in the middle of the solve chain)
            if (++kb >= nblk() || rb_of(kb) / NB > pmax) return false;
            m = 0;
            if (first_tile(kb, m)) return true;
            m = ntile();
        }
    }
};
// Backward (the mirror image): COLUMN blocks are dealt to the CTAs (block kb -> CTA kb % G, which also solves it); the row
// pairs are met from the last one down -- the rows below the pivots first, then pair p as soon as block p is solved -- and pair
// p is applied to every owned block kb whose rows below the pivots reach into it (pmin(kb) = rb(kb) / 128 <= p), the highest
// block first: kb = p - 1 is the one the chain waits for.  Item = (pair p, block kb, 64-row slice sub), slice 1 before 0.
// rect != 0: the pairs that hold only rows below the pivots (p >= nblk) were summed by the level-wide pre-pass (k_bwd_rect)
// and are not items: pend() bounds the pairs of the work list.
struct BwdSched {
    int nr, nc, c, G, rect;
    __host__ __device__ int nblk() const { return (nc + NB - 1) / NB; }
    __host__ __device__ int npairs() const { return (nr + NB - 1) / NB; }
    __host__ __device__ int pend() const { return rect ? nblk() : npairs(); }
    __host__ __device__ int rb_of(int kb) const { return nc < (kb + 1) * NB ? nc : (kb + 1) * NB; }
    __host__ __device__ int pmin_of(int kb) const { return rb_of(kb) / NB; }
    __host__ __device__ int subs_of(int p) const { return p * NB + 64 < nr ? 2 : 1; }       // 64-row slices of pair p inside the front
    // highest block owned by this CTA that pair p is applied to; < 0: none
    __host__ __device__ int first_kb(int p) const {
        const int last = nblk() - 1;
        int kmax = p >= pmin_of(last) ? last : (last - 1 < p - 1 ? last - 1 : p - 1);
        if (kmax < c) return -1;
        return kmax - (kmax - c) % G;
    }
    __host__ __device__ bool first_item(int& p, int& kb, int& sub) const {
        for (p = pend() - 1; p >= 0; p--) {
            kb = first_kb(p);
            if (kb >= 0) { sub = subs_of(p) - 1; return true; }
            if (p < nblk()) break;               // below the last pivot pair first_kb only shrinks
        }
        return false;
    }
    __host__ __device__ bool next_slice(int& p, int& kb, int& sub) const {
        if (sub > 0) { sub--; return true; }
        if (kb - G >= 0) { kb -= G; sub = subs_of(p) - 1; return true; }
        if (--p < 0) return false;
        kb = first_kb(p);
        if (kb < 0) return false;                // (p < nblk here: first_kb(p) only shrinks with p)
        sub = subs_of(p) - 1;
        return true;
    }
};

// CPU replay of both persistent schedules for ONE front (nr rows, nc pivot columns) on G CTAs: every CTA's program is followed
// with the waits the kernels make, the CTAs advanced round robin.  Returns 0 when every (block, tile) is applied exactly once
// and in the order of the launch-per-step kernels, every block is solved once and only after everything it reads, and no CTA
// is left waiting (deadlock); a positive code otherwise (tests/test_host.py).
int persist_schedule_check(int nr, int nc, int G) {
    if (nr < 1 || nc < 1 || nc > nr || G < 1) return 1;
    const int nblk = (nc + NB - 1) / NB;
    {   // ---- forward
        const int ntile = (nr + SOLVE_FT - 1) / SOLVE_FT;
        std::vector<int> flag(nblk, 0), applied((size_t)nblk * ntile, 0), last_kb(ntile, -1);
        struct St { int kb, m; bool have, started; };
        std::vector<St> st(G);
        auto diag = [&](int kb) -> int {          // all earlier blocks applied to the rows of block kb
            if (flag[kb]) return 10;
            for (int j = 0; j < kb; j++)
                for (int m = 2 * kb; m < std::min(ntile, 2 * kb + 2); m++)
                    if (applied[(size_t)j * ntile + m] != 1) return 11;
            flag[kb] = 1;
            return 0;
        };
        for (int c = 0; c < G; c++) {
            const FwdSched S{nr, nc, c, G};
            St& q = st[c];
            q.kb = 0; q.m = 0; q.started = false;
            q.have = S.first_tile(0, q.m);
            if (!q.have) { q.m = ntile; q.have = S.next_item(q.kb, q.m); }
        }
        bool progress = true, all_done = false;
        while (progress && !all_done) {
            progress = false; all_done = true;
            for (int c = 0; c < G; c++) {
                const FwdSched S{nr, nc, c, G};
                St& q = st[c];
                if (!q.started) { q.started = true; progress = true; if (c == 0) { int e = diag(0); if (e) return e; } }
                while (q.have) {
                    if (!flag[q.kb]) break;      // persist_wait(flag + kb)
                    const int rb = S.rb_of(q.kb);
                    if (q.m >= ntile || q.m * SOLVE_FT + SOLVE_FT - 1 < rb) return 12;      // not a tile below the block
                    if ((q.m >> 1) % G != c) return 13;                                     // not this CTA's pair
                    if (applied[(size_t)q.kb * ntile + q.m]++) return 14;
                    if (last_kb[q.m] >= q.kb) return 15;
                    last_kb[q.m] = q.kb;
                    const bool pair_done = (q.m & 1) || q.m + 1 >= ntile;
                    if (pair_done && (q.m >> 1) == q.kb + 1 && q.kb + 1 < nblk) { int e = diag(q.kb + 1); if (e) return e; }
                    q.have = S.next_item(q.kb, q.m);
                    progress = true;
                }
                if (q.have) all_done = false;
            }
        }
        if (!all_done) return 16;                // deadlock
        for (int kb = 0; kb < nblk; kb++) {
            if (!flag[kb]) return 17;
            const int rb = std::min(nc, (kb + 1) * NB);
            for (int m = 0; m < ntile; m++)
                if (applied[(size_t)kb * ntile + m] != (m * SOLVE_FT + SOLVE_FT - 1 >= rb ? 1 : 0)) return 18;
        }
    }
    for (int rect = 0; rect < 2; rect++) {   // ---- backward, without / with the rectangle pre-pass
        const BwdSched S0{nr, nc, 0, G, rect};
        const int np = S0.pend();
        std::vector<int> flag(nblk, 0), applied((size_t)nblk * np, 0), last_p(nblk, np);
        struct St { int p, kb, sub; bool have, started; };
        std::vector<St> st(G);
        auto solve = [&](int kb, int c) -> int {
            if (flag[kb]) return 20;
            if (kb % G != c) return 21;
            for (int p = S0.pmin_of(kb); p < np; p++) if (applied[(size_t)kb * np + p] != 1) return 22;
            for (int j = kb + 1; j < nblk; j++) if (!flag[j]) return 23;
            flag[kb] = 1;
            return 0;
        };
        bool progress = true, all_done = false;
        while (progress && !all_done) {
            progress = false; all_done = true;
            for (int c = 0; c < G; c++) {
                const BwdSched S{nr, nc, c, G, rect};
                St& q = st[c];
                if (!q.started) {
                    q.started = true; progress = true;
                    q.have = S.first_item(q.p, q.kb, q.sub);
                    // a block with no row below it at all (last block of a front with nr == nc, nc % 128 == 0): solved up front
                    const int last = nblk - 1;
                    if (last % G == c && S.pmin_of(last) >= np) { int e = solve(last, c); if (e) return e; }
                }
                while (q.have) {
                    if (q.p < nblk && q.p > q.kb && !flag[q.p]) break;             // persist_wait(flag + p)
                    if (q.kb % G != c || q.kb < 0 || q.kb >= nblk || q.p < S.pmin_of(q.kb) || q.p >= np) return 24;
                    if (q.sub < 0 || q.sub >= S.subs_of(q.p)) return 25;
                    {   // every pivot row of the pair that the item reads (r >= rb of its block) is solved
                        const int lo = std::max(q.p * NB, S.rb_of(q.kb)), hi = std::min(q.p * NB + NB, nc);
                        if (lo < hi && !flag[q.p]) return 32;
                    }
                    if (q.sub == 0) {
                        if (applied[(size_t)q.kb * np + q.p]++) return 26;
                        if (last_p[q.kb] != q.p + 1) return 27;      // pairs of a block from the last one down, none skipped
                        last_p[q.kb] = q.p;
                        if (q.p == S.pmin_of(q.kb)) { int e = solve(q.kb, c); if (e) return e; }
                    }
                    int np_ = q.p, nkb = q.kb, nsub = q.sub;
                    const bool hnext = S.next_slice(np_, nkb, nsub);
                    if (q.sub > 0 && !(hnext && np_ == q.p && nkb == q.kb && nsub == q.sub - 1)) return 28;
                    q.have = hnext; q.p = np_; q.kb = nkb; q.sub = nsub;
                    progress = true;
                }
                if (q.have) all_done = false;
            }
        }
        if (!all_done) return 29;                // deadlock
        for (int kb = 0; kb < nblk; kb++) {
            if (!flag[kb]) return 30;
            for (int p = 0; p < np; p++) if (applied[(size_t)kb * np + p] != (p >= S0.pmin_of(kb) ? 1 : 0)) return 31;
        }
    }
    return 0;
}

// CTA-wide wait: ONE thread polls (a front's CTAs all wait for the same word: 256 pollers per CTA saturate its L2 slice and
// delay the very store they wait for), the barrier hands the acquired state to the others
__device__ __forceinline__ void persist_wait_cta(const int* p, int need, int* err) {
    if (threadIdx.x == 0) persist_wait(p, need, err);
    __syncthreads();
}
// forward sweep of the large fronts of one level.  Rows are dealt in PAIRS of 64-row tiles (128 rows: pair p -> CTA p % G of
// the front); the CTA that owns pair kb solves diagonal block kb right after it has applied block kb - 1 to that pair.
__global__ void __launch_bounds__(256, 1) k_fwd_persist(const PersistLevel pl, const double* __restrict__ L,
                                                        const double* __restrict__ Minv, double* T, double* X, int* flags, int* err,
                                                        long long* dbg) {
    extern __shared__ double sm[];
    double* ring = sm;                     // 2 x [128 columns][64 rows]
    double* Ls = ring + 2 * NB * 64;
    double* Ms = Ls + DIAG_PACK;
    double* ts = Ms + MINV_HALF;
    double* xs = ts + NB;
    double* red = xs + 2 * NB;             // [4][64]
    for (int fi = pl.shared ? pl.cmap[blockIdx.x] : (int)blockIdx.x; fi < pl.nf; fi += pl.shared ? pl.nf : (int)gridDim.x) {
    const PersistFront pfr = pl.fr[fi];
    const FrontS f = pfr.f;
    const int c = pl.shared ? (int)blockIdx.x - pfr.cta0 : 0, G = pfr.ncta;
    int* flag = flags + pfr.sync0;
    double* t = T + f.rowptr;
    double* x = X + f.col0;
    const double* P = L + f.loff;
    const double* minv = Minv + f.ioff;
    const int tid = threadIdx.x, rr = tid & (SOLVE_FT - 1), cq = tid >> 6;
    const int nblk = (f.nc + NB - 1) / NB, ntile = (f.nr + SOLVE_FT - 1) / SOLVE_FT;
    const FwdSched S{f.nr, f.nc, c, G};
    auto first_tile = [&](int kb, int& m) { return S.first_tile(kb, m); };
    auto next_item = [&](int& kb, int& m) { return S.next_item(kb, m); };
    auto issue = [&](int kb, int m, int buf) {
        const int k0 = kb * NB;
        stage_rows64(P, f.ld, f.nr, k0, min(NB, f.nc - k0), m * SOLVE_FT, ring + buf * NB * 64, tid);
    };
    auto diag = [&](int kb, bool ts_ready) {     // staged data of block kb complete and visible (caller); ts_ready: so is ts
        const int k0 = kb * NB, w = min(NB, f.nc - k0);
        if (!ts_ready) {
            if (tid < NB) ts[tid] = (tid < w) ? __ldcg(t + k0 + tid) : 0.0;
            __syncthreads();
        }
        fwd_diag_core(Ls, Ms, ts, w, tid);
        if (dbg && tid == 0) dbg[c * 8 + 4] = clock64();
        if (tid < w) { __stcg(t + k0 + tid, ts[tid]); __stcg(x + k0 + tid, ts[tid]); }
        __syncthreads();
        if (tid == 0) { __threadfence(); st_release(flag + kb, 1); }
        if (dbg && tid == 0) { long long gt; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt)); dbg[c * 8 + 6] = gt; dbg[c * 8 + 5] = clock64(); }
        if (kb + G < nblk) {
            const int k1 = (kb + G) * NB;
            stage_diag_issue(P, f.ld, k1, min(NB, f.nc - k1), minv + (long long)(kb + G) * MINV_BLK, Ls, Ms, tid);
        }
        cp_async_commit();
    };
    if (c < nblk) {
        const int k1 = c * NB;
        stage_diag_issue(P, f.ld, k1, min(NB, f.nc - k1), minv + (long long)c * MINV_BLK, Ls, Ms, tid);
    }
    cp_async_commit();
    int kb = 0, m = 0;
    bool have = first_tile(0, m);
    if (!have) { m = ntile; have = next_item(kb, m); }
    if (have) issue(kb, m, 0);
    cp_async_commit();
    if (c == 0) {
        cp_async_wait<1>();
        __syncthreads();
        diag(0, false);
    }
    int xs_kb = -1, it = 0;
    while (have) {
        int nkb = kb, nm = m;
        const bool hnext = next_item(nkb, nm);
        if (hnext) issue(nkb, nm, (it + 1) & 1);
        cp_async_commit();
        const int k0 = kb * NB, w = min(NB, f.nc - k0), rb = k0 + w;
        if (kb != xs_kb) {                       // x of block kb: published by the owner of pair kb
            if (dbg && tid == 0 && kb + 1 == c) dbg[c * 8 + 7] = clock64();
            persist_wait_cta(flag + kb, 1, err);
            if (dbg && tid == 0 && kb + 1 == c) { long long gt; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt)); dbg[c * 8 + 0] = gt; dbg[c * 8 + 1] = clock64(); }
            if (tid < NB) xs[tid] = (tid < w) ? __ldcg(t + k0 + tid) : 0.0;
            xs_kb = kb;
        }
        const int r0 = m * SOLVE_FT;
        const bool wr = tid < SOLVE_FT && r0 + tid < f.nr && r0 + tid >= rb;
        const double tprev = wr ? __ldcg(t + r0 + tid) : 0.0;
        cp_async_wait<1>();
        __syncthreads();
        if (dbg && tid == 0 && kb + 1 == c && !(m & 1)) dbg[c * 8 + 2] = clock64();
        {
            const bool rowok = r0 + rr < f.nr && r0 + rr >= rb;
            const int wq = min(32, w - cq * 32);
            const double* sp = ring + (it & 1) * NB * 64 + cq * 32 * 64 + rr;
            double a0 = 0.0, a1 = 0.0;
            if (rowok) {
                int j = 0;
                for (; j + 1 < wq; j += 2) {
                    const double s0 = sp[j * 64], s1 = sp[(j + 1) * 64];
                    a0 = fma(s0, xs[cq * 32 + j], a0); a1 = fma(s1, xs[cq * 32 + j + 1], a1);
                }
                if (j < wq) a0 = fma(sp[j * 64], xs[cq * 32 + j], a0);
            }
            red[cq * SOLVE_FT + rr] = a0 + a1;
        }
        __syncthreads();
        // pair kb + 1: its owner solves the next diagonal block as soon as the pair is complete (the chain), then goes on with
        // block kb; the right-hand side of that solve goes straight to shared memory (same values as written to T)
        const bool chain = (m >> 1) == kb + 1 && kb + 1 < nblk;
        const bool pair_done = (m & 1) || m + 1 >= ntile;
        if (tid < SOLVE_FT) {
            const double v = tprev - ((red[tid] + red[SOLVE_FT + tid]) + (red[2 * SOLVE_FT + tid] + red[3 * SOLVE_FT + tid]));
            if (wr) __stcg(t + r0 + tid, v);
            if (chain) {
                ts[(m & 1) * SOLVE_FT + tid] = (wr && r0 + tid < f.nc) ? v : 0.0;
                if (pair_done && !(m & 1)) ts[SOLVE_FT + tid] = 0.0;      // the pair has one tile only
            }
        }
        if (chain && pair_done) {
            __syncthreads();
            if (dbg && tid == 0) dbg[c * 8 + 3] = clock64();
            diag(kb + 1, true);
        }
        have = hnext; kb = nkb; m = nm; it++;
    }
    cp_async_wait<0>();
    __syncthreads();                             // (several fronts per CTA: shared memory is reused by the next one)
    }
}

// Rectangle pre-pass of the backward sweep: for every front of the level, every column block kb and every pair p that holds
// only rows below the pivots (p >= nblk: their x is known before the sweep starts), the column sums of L[pair p, block kb]^T x
// -- the tile of k_bwd_upd, same code -- into rect[(rect0 + kb * nR + p - nblk) * 128 ..].  One launch at memory speed over
// all (front, block, pair) instead of the block owners streaming their rectangles one 64 KB slice at a time.
__global__ void __launch_bounds__(256) k_bwd_rect(const PersistLevel pl, const double* __restrict__ L, const double* __restrict__ T,
                                                  double* __restrict__ rect) {
    extern __shared__ double sm[];
    __shared__ double red[1][8][32];
    int lo = 0, hi = pl.nf - 1;                  // last front with rect0 <= blockIdx.x
    while (lo < hi) { const int mid = (lo + hi + 1) >> 1; if (pl.fr[mid].rect0 <= (int)blockIdx.x) lo = mid; else hi = mid - 1; }
    const PersistFront pfr = pl.fr[lo];
    const FrontS f = pfr.f;
    const int nblk = (f.nc + NB - 1) / NB, nR = (f.nr + NB - 1) / NB - nblk;
    const int idx = blockIdx.x - pfr.rect0, kb = idx / nR, p = nblk + idx - kb * nR;
    bwd_pair_sums<1>(f, kb, p * SOLVE_BT, L, T + f.rowptr, 0, 1, sm, red);
    const int tid = threadIdx.x;
    if (tid < NB) rect[(long long)blockIdx.x * NB + tid] = red[0][2 * (tid >> 5)][tid & 31] + red[0][2 * (tid >> 5) + 1][tid & 31];
}

// backward sweep of the large fronts of one level, the mirror image of the forward one: the CTA that owns column block kb keeps
// z_kb = t_kb - sum over the pairs p (from the last one down) of L[pair p, block kb]^T x[pair p] in place in T, applies pair
// kb + 1 the moment block kb + 1 is published, and solves L11^T x = z at once.  No partial sums travel between CTAs.
__global__ void __launch_bounds__(256, 1) k_bwd_persist(const PersistLevel pl, const double* __restrict__ L,
                                                        const double* __restrict__ Minv, double* T, double* X, int* flags,
                                                        const double* __restrict__ rect, int* err, long long* dbg) {
    extern __shared__ double sm[];
    double* ring = sm;
    double* Ls = ring + 2 * NB * 64;
    double* Ms = Ls + DIAG_PACK;
    double* zs = Ms + MINV_HALF;
    double* red = zs + NB;                 // [8][32]
    for (int fi = pl.shared ? pl.cmap[blockIdx.x] : (int)blockIdx.x; fi < pl.nf; fi += pl.shared ? pl.nf : (int)gridDim.x) {
    const PersistFront pfr = pl.fr[fi];
    const FrontS f = pfr.f;
    const int c = pl.shared ? (int)blockIdx.x - pfr.cta0 : 0, G = pfr.ncta;
    int* flag = flags + pfr.sync0;
    double* t = T + f.rowptr;
    double* x = X + f.col0;
    const double* P = L + f.loff;
    const double* minv = Minv + f.ioff + MINV_HALF;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, rr = tid & 63, cq = tid >> 6;
    const int nblk = (f.nc + NB - 1) / NB;
    const BwdSched S{f.nr, f.nc, c, G, rect != nullptr};
    auto issue = [&](int p, int kb, int sub, int buf) {
        const int k0 = kb * NB;
        stage_rows64(P, f.ld, f.nr, k0, min(NB, f.nc - k0), p * NB + sub * 64, ring + buf * NB * 64, tid);
    };
    // zs holds z of block kb_ (zero past w): solve, publish, stage the diagonal block of the next owned block
    auto solve_block = [&](int kb_) {
        const int k0 = kb_ * NB, w = min(NB, f.nc - k0);
        bwd_diag_core(Ls, Ms, zs, w, tid);
        if (tid < w) { __stcg(t + k0 + tid, zs[tid]); __stcg(x + k0 + tid, zs[tid]); }
        __syncthreads();
        if (tid == 0) { __threadfence(); st_release(flag + kb_, 1); }
        if (kb_ - G >= 0) {
            const int k1 = (kb_ - G) * NB;
            stage_diag_issue(P, f.ld, k1, min(NB, f.nc - k1), minv + (long long)(kb_ - G) * MINV_BLK, Ls, Ms, tid);
        }
        cp_async_commit();
    };
    // diagonal blocks this CTA solves: kb % G == c, descending
    const int dfirst = nblk - 1 - (((nblk - 1 - c) % G) + G) % G;
    if (dfirst >= 0) {
        const int k1 = dfirst * NB;
        stage_diag_issue(P, f.ld, k1, min(NB, f.nc - k1), minv + (long long)dfirst * MINV_BLK, Ls, Ms, tid);
    }
    cp_async_commit();
    int p = 0, kb = 0, sub = 0;
    bool have = S.first_item(p, kb, sub);
    if (have) issue(p, kb, sub, 0);
    cp_async_commit();
    if (rect) {
        // the rows below the pivots, summed by k_bwd_rect: z_kb = t_kb - R[last pair] - ... - R[nblk] for every owned block, the
        // order of the launch-per-step kernels (sixteen loads in flight, subtracted one after the other)
        const int nR = S.npairs() - nblk;
        if (nR > 0 && tid < NB)
            for (int kq = dfirst; kq >= 0; kq -= G) {
                const int k0 = kq * NB, w = min(NB, f.nc - k0);
                if (tid >= w) continue;
                const double* rp = rect + ((long long)pfr.rect0 + (long long)kq * nR) * NB + tid;
                double z = __ldcg(t + k0 + tid);
                int q = nR - 1;
                for (; q >= 15; q -= 16) {
                    double v[16];
#pragma unroll
                    for (int e = 0; e < 16; e++) v[e] = __ldg(rp + (long long)(q - e) * NB);
#pragma unroll
                    for (int e = 0; e < 16; e++) z -= v[e];
                }
                for (; q >= 0; q--) z -= __ldg(rp + (long long)q * NB);
                __stcg(t + k0 + tid, z);
            }
        __syncthreads();
    }
    if (dfirst == nblk - 1 && S.pmin_of(nblk - 1) >= S.pend()) {        // no pair left to apply to the last block: z = t
        const int k0 = dfirst * NB, w = min(NB, f.nc - k0);
        if (tid < NB) zs[tid] = (tid < w) ? __ldcg(t + k0 + tid) : 0.0;
        cp_async_wait<1>();
        __syncthreads();
        solve_block(dfirst);
    }
    int acq = nblk;                              // blocks >= acq are known to be published
    int it = 0;
    double pj[32];
    double zprev = 0.0;
    while (have) {
        int np = p, nkb = kb, nsub = sub;
        const bool hnext = S.next_slice(np, nkb, nsub);
        if (hnext) issue(np, nkb, nsub, (it + 1) & 1);
        cp_async_commit();
        const int k0 = kb * NB, w = min(NB, f.nc - k0), rb = k0 + w;
        const bool first_slice = sub + 1 == S.subs_of(p);
        if (first_slice) {
#pragma unroll
            for (int j = 0; j < 32; j++) pj[j] = 0.0;
            if (tid < NB) zprev = (tid < w) ? __ldcg(t + k0 + tid) : 0.0;      // (written by this thread when the pair before was applied)
        }
        if (p < nblk && p > kb && p < acq) {     // the pivot rows of pair p at or below rb: block p (p == kb: only rows below the pivots)
            if (dbg && tid == 0 && kb == p - 1) dbg[2048 + c * 16 + 0] = clock64();
            persist_wait_cta(flag + p, 1, err);
            acq = p;
            if (dbg && tid == 0 && kb == p - 1) { long long gt; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt)); dbg[2048 + c * 16 + 8] = gt; dbg[2048 + c * 16 + 1] = clock64(); }
        }
        const int r = p * NB + sub * 64 + rr;
        const double tv = (r < f.nr && r >= rb) ? __ldcg(t + r) : 0.0;
        cp_async_wait<1>();
        __syncthreads();
        if (r < f.nr && r >= rb) {
            const int wq = min(32, w - cq * 32);
            const double* sp = ring + (it & 1) * NB * 64 + cq * 32 * 64 + rr;
#pragma unroll
            for (int j = 0; j < 32; j++)
                if (j < wq) pj[j] = fma(sp[j * 64], tv, pj[j]);
        }
        if (sub == 0) {                          // pair complete: column sums over its 128 rows
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const bool up = (lane & o) != 0;
#pragma unroll
                for (int j = 0; j < o; j++) {
                    const double send = up ? pj[j] : pj[j + o];
                    const double keep = up ? pj[j + o] : pj[j];
                    pj[j] = keep + __shfl_xor_sync(0xffffffffu, send, o);
                }
            }
            red[warp * 32 + lane] = pj[0];
            __syncthreads();
            const bool last_pair = p == S.pmin_of(kb);       // every pair of block kb applied: solve it
            if (tid < NB) {
                const double z = zprev - (red[(2 * (tid >> 5)) * 32 + (tid & 31)] + red[(2 * (tid >> 5) + 1) * 32 + (tid & 31)]);
                if (last_pair) zs[tid] = (tid < w) ? z : 0.0;
                else if (tid < w) __stcg(t + k0 + tid, z);
            }
            if (last_pair) {
                __syncthreads();
                if (dbg && tid == 0) dbg[2048 + c * 16 + 2] = clock64();
                solve_block(kb);
                if (dbg && tid == 0) { long long gt; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt)); dbg[2048 + c * 16 + 9] = gt; dbg[2048 + c * 16 + 3] = clock64(); }
            }
        }
        __syncthreads();                         // ring buffer / red free for the slice after the next
        have = hnext; p = np; kb = nkb; sub = nsub; it++;
    }
    cp_async_wait<0>();
    __syncthreads();
    }
}

__global__ void k_perm_gather(const double* __restrict__ B, long long ldB, const int* __restrict__ perm, int n,
                              double* __restrict__ X, long long ldX) {
    const double* b = B + blockIdx.y * ldB;
    double* x = X + blockIdx.y * ldX;
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) x[k] = b[perm[k]];
}
__global__ void k_perm_scatter(double* __restrict__ B, long long ldB, const int* __restrict__ perm, int n,
                               const double* __restrict__ X, long long ldX) {
    double* b = B + blockIdx.y * ldB;
    const double* x = X + blockIdx.y * ldX;
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) b[perm[k]] = x[k];
}
__global__ void k_copy_cols(const double* __restrict__ S, long long lds, double* __restrict__ Dst, long long ldd, int n) {
    const double* s = S + blockIdx.y * lds;
    double* d = Dst + blockIdx.y * ldd;
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) d[k] = s[k];
}
// LDL' semantics on top of the LL' factor (cholmod.options['supernodal'] = 0, reference src/C/cholmod.c:60-64,437-439):
// the engine factors P A P' = Lt S Lt' with S = diag(+-1) (signed square-root form of LDL' without pivoting, any symmetric
// matrix whose pivots are nonzero: positive definite, quasi-definite KKT systems, ...).  L_ldl = Lt diag(Lt)^-1 and
// D = S diag(Lt)^2, so the LDL' systems are the Lt sweeps with a diagonal scaling before, between or after.
// mode 1: x *= l, 2: x *= s / l, 3: x *= s / l^2, 4: x *= s   (l = diag(Lt), s = signs, both in permuted order)
__global__ void k_scale_by_diag(double* __restrict__ x, long long ldx, const double* __restrict__ dg,
                                const double* __restrict__ sg, int n, int mode) {
    double* col = x + (long long)blockIdx.y * ldx;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        if (mode == 4) { col[i] *= sg[i]; continue; }
        const double l = dg[i];
        col[i] = mode == 1 ? col[i] * l : mode == 2 ? sg[i] * col[i] / l : sg[i] * col[i] / (l * l);
    }
}

__global__ void k_diag(const FrontD* __restrict__ F, int ns, const double* __restrict__ L, double* __restrict__ d) {
    for (int s = blockIdx.x; s < ns; s += gridDim.x) {
        const FrontD f = F[s];
        for (int c = threadIdx.x; c < f.nc; c += blockDim.x) d[f.col0 + c] = L[f.loff + (long long)c * f.ld + c];
    }
}

// ---------------------------------------------------------------------------------------------------
// host-side driver
// ---------------------------------------------------------------------------------------------------
struct Launch {            // one grouped launch: groups [goff, goff+ng) in the schedule arrays, `ctas` CTAs
    int goff = 0, ng = 0, ctas = 0;
    int sgi = -1;          // solve launches with few groups: index of the by-value metadata in CholDevice::sgroups
};
struct LevelSched {
    int ea_off = 0, ea_cnt = 0;
    // small fronts by rows: <= 32, <= 64, then three classes of the 256-thread kernel by how many CTAs fit an SM with the whole
    // front in shared memory: <= 96 rows (74.5 KB: three), <= 118 (112 KB: two), <= 128 (one).  One class for 65..128 rows sized
    // every launch for its largest front: a level of 8082 fronts of <= 96 rows and 400 of up to 128 ran at one CTA per SM.
    static constexpr int NSMALL = 5;
    int small_off[NSMALL] = {0, 0, 0, 0, 0}, small_cnt[NSMALL] = {0, 0, 0, 0, 0}, small_maxnr[NSMALL] = {0, 0, 0, 0, 0};
    std::vector<Launch> panel, upd;   // per block step kb (upd = all trailing column tiles)
    std::vector<Launch> updA, updB;   // lookahead split of upd: next block column / the rest
    std::vector<Launch> updN, updF;   // two-level update: near (inside the super-block, K = 128) / far (K = super-block)
    std::vector<Launch> updFA, updFB; // far split: column tiles of the next super-block / the rest (look-ahead)
    Launch syrk;
    std::vector<Launch> sfwd, sbwd;   // triangular solves of large fronts: row tiles per block step
    Launch gfwd;                      // forward gather of large fronts: GATHER_ROWS-row chunks
    int gq_off = 0;                   // first chunk of this level in the chunk-child table
    Launch sbig;                      // the fronts the SOLVES treat as large (one entry each): panel[0] minus the one-block fronts with few rows
    int small_all_off = 0, small_all_cnt = 0;
    int pgi = -1, pctas = 0;          // persistent sweeps (one right-hand side): index in CholDevice::plevels, CTAs
    int prect = 0;                    // tiles of the backward rectangle pre-pass (k_bwd_rect)
};

class CholDevice {
public:
    const CholPlan* plan = nullptr;
    CholOpts opts;
    int device = 0;
    cudaStream_t stream = nullptr, stream2 = nullptr;   // stream2: lookahead (trailing update overlapped with the next panel)
    cudaEvent_t evP = nullptr, evB = nullptr;
    double *dL = nullptr, *dW = nullptr, *dval = nullptr, *dT = nullptr, *dX = nullptr, *dBstage = nullptr;
    i64 bstage_cap = 0;            // doubles
    long long* damap = nullptr;
    FrontD* dF = nullptr;
    int *drows = nullptr, *drel = nullptr, *dchild = nullptr, *dperm = nullptr, *dlevel_fronts = nullptr;
    int *dsched = nullptr, *dminor = nullptr;
    std::vector<int> hsched;
    double* ddiag = nullptr;       // scratch for factored diagonal blocks of one panel launch
    EAItem* dea = nullptr;
    int* dea_child = nullptr;      // item-child list of the extend-add items
    int2* dgq = nullptr;           // chunk-child lists of the forward gather (k_fwd_gather)
    int* dgq_child = nullptr;
    std::vector<LevelSched> levels;
    std::vector<SolveGroups> sgroups;   // sgroups[0] = empty (use the schedule arrays)
    std::map<int, cudaGraphExec_t> solve_graphs;   // key: columns * 4 + forward * 2 + backward
    bool use_graphs = true;
    bool upd_tma = false;          // operand tiles of k_update staged by TMA bulk copies (B200S_UPDATE_TMA=0/1 overrides the default)
    void drop_graphs() { for (auto& kv : solve_graphs) cudaGraphExecDestroy(kv.second); solve_graphs.clear(); }
    i64 solve_cols = 0;            // capacity (columns) of dT / dX
    int max_solve_ctas = 1;        // most CTAs of one backward-update launch (sizes the partial-sum buffer)
    std::vector<PersistLevel> plevels;    // persistent sweeps: launch records of the levels with 1 .. 2 x SMs large fronts
    PersistFront* dpfront = nullptr;      // their front records and CTA -> front maps (device)
    int* dpcmap = nullptr;
    double* drect = nullptr;              // column sums of the rectangle pre-pass of one level (reused level after level)
    bool persist_coop = true;             // B200S_PERSIST_COOP=0: plain launches
    bool persist_rect = true;             // B200S_PERSIST_RECT=0: the block owners stream the rows below the pivots themselves
    int solve_medium_nr = 320;     // solves: one-block large fronts with at most this many rows run with the small fronts (B200S_SOLVE_MEDIUM_NR)
    int persist_mode = 3;          // bit 0: forward, bit 1: backward sweep by k_fwd_persist / k_bwd_persist (B200S_SOLVE_PERSIST)
    int persist_default = 3, persist_hw = 3;      // B200S_SOLVE_PERSIST or 3; 0 when a persistent CTA does not fit an SM
    int* dsync = nullptr;          // [2][nsync]: forward flags, backward flags (zeroed per sweep pair)
    int nsync = 0;
    long long* persist_dbg = nullptr;   // B200S_PERSIST_DBG: phase clocks of the root front's forward chain (mapped host memory)
    int* herr = nullptr;           // mapped host word: a persistent kernel gave up a wait (schedule error instead of a hang)
    double* dpart = nullptr;
    double* dMinv = nullptr;       // inverted 32 x 32 diagonal sub-blocks of the large fronts (solve phase)
    int *dinv_front = nullptr, *dinv_kb = nullptr;
    int ninvblk = 0;
    bool minv_valid = false;
    bool numeric = false, profiling = false;
    cudaEvent_t ev[8] = {};
    std::vector<cudaEvent_t> pev;  // profiling event pool
    struct Span { int cls; size_t e0, e1; };
    std::vector<Span> spans;
    size_t pe = 0;
    unsigned char* downed = nullptr;   // per front: 1 = this device factors it (multi-GPU subtree ownership)
    i64 total_bytes = 0;

    ~CholDevice() {
        cudaSetDevice(device);
        if (stream2) cudaStreamSynchronize(stream2);
        if (stream) cudaStreamSynchronize(stream);      // blocks go back to the caching allocator: nothing may still use them
        drop_graphs();
        pool_free(dL); pool_free(dW); pool_free(dval); pool_free(dT); pool_free(dX); pool_free(dBstage); pool_free(damap); pool_free(dF);
        pool_free(drows); pool_free(drel); pool_free(dchild); pool_free(dperm); pool_free(dlevel_fronts);
        pool_free(dsched); pool_free(dminor); pool_free(dea); pool_free(dea_child); pool_free(dgq); pool_free(dgq_child); pool_free(ddiag); pool_free(dpart); pool_free(downed);
        pool_free(dreach); pool_free(dsp_i); pool_free(dsp_x); pool_free(dsp_cnt); pool_free(dsp_oi); pool_free(dsp_ox); pool_free(dsp_meta);
        for (auto& e : ev_sp) if (e) cudaEventDestroy(e);
        pool_free(dMinv); pool_free(dinv_front); pool_free(dinv_kb); pool_free(ddiagL); pool_free(dsgn);
        pool_free(dsync); pool_free(dpfront); pool_free(dpcmap); pool_free(drect);
        for (auto& e : ev) if (e) cudaEventDestroy(e);
        for (auto& e : pev) cudaEventDestroy(e);
        if (evP) cudaEventDestroy(evP);
        if (evB) cudaEventDestroy(evB);
        if (stream2) cudaStreamDestroy(stream2);
        if (stream) cudaStreamDestroy(stream);
    }
    template <class T> int upload(T** dst, const T* src, size_t count) {
        size_t bytes = std::max<size_t>(count, 1) * sizeof(T);
        CUDA_TRY(pool_malloc((void**)dst, bytes));
        total_bytes += bytes;
        if (count) CUDA_TRY(cudaMemcpy(*dst, src, count * sizeof(T), cudaMemcpyHostToDevice));
        return ST_OK;
    }
    int init();
    int factorize(const double* val, bool on_device, i64* minor, CholTimes* times);
    int factor_begin(const double* val, bool on_device);
    int factor_level(int l, int phase = 3);     // phase bit 0: everything up to the panels' in-panel updates, bit 1: the Schur complements
    int set_syrk_split(const unsigned char* own, const int* lo, const int* hi, const long long* base, double* scratch);
    SyrkSplit syrk_split = {nullptr, nullptr, nullptr, nullptr, nullptr};
    unsigned char* dsy_own = nullptr; int *dsy_lo = nullptr, *dsy_hi = nullptr; long long* dsy_base = nullptr;
    int factor_end(i64* minor, CholTimes* times);
    int factor_enqueue(const double* val_dev);
    int set_owned(const unsigned char* owned_host);
    int solve(int sys, double* B, i64 nrhs, i64 ldB, bool on_device, CholTimes* times, bool async = false,
              const unsigned char* active_fronts = nullptr);
    // sparse right-hand sides: reach-restricted forward sweep, sparse upload, compacted download (cholmod.spsolve)
    int spsolve(int sys, i64 ncols, const i64* Bp, const i64* Bi, const double* Bx, std::vector<i64>& Xp, std::vector<i64>& Xi,
                std::vector<double>& Xx, CholTimes* times);
    int* dreach = nullptr; i64 reach_cap = 0;        // filtered per-level lists of small fronts (reach of the current right-hand sides)
    i64 *dsp_i = nullptr, *dsp_meta = nullptr; double* dsp_x = nullptr; int* dsp_cnt = nullptr; i64 sp_cap = 0, spout_cap = 0, spmeta_cap = 0;
    cudaEvent_t ev_sp[2] = {nullptr, nullptr};
    i64* dsp_oi = nullptr; double* dsp_ox = nullptr;
    bool ldl = false;             // LDL' semantics of sys 2..6 (supernodal = 0)
    double* ddiagL = nullptr;     // diagonal of L (permuted order), valid while diagL_valid
    double* dsgn = nullptr;       // ldl: sign of every pivot (+-1, permuted order, padded), written by the factorization kernels
    bool diagL_valid = false;
    int ensure_solve_ws(i64 cols);
    // distributed solves (kvxopt_b200/dist.py): one right-hand side, level by level, only the fronts marked by set_owned
    int solve_dist_begin(const double* b_dev);
    int solve_dist_level(int backward, int l);
    int solve_dist_end(double* x_dev);
};

static constexpr size_t SMEM_PANEL = (size_t)(NB * LDL + NB * LDX + NB) * sizeof(double);
static constexpr size_t SMEM_UPDATE = (size_t)(STAGES * BK * (LDT + LDTB)) * sizeof(double);

// ---- child lists of the assembly items and of the forward-gather chunks (host; verified by chol_child_lists_check) -------------
// extend-add items of front s: its destination columns cut into ranges of about EA_TARGET entries, and per item the children
// with an update column inside it -- two passes (count, fill) over the children in order, each child hopping from item to item
// through its sorted relative indices
static void ea_items_of_front(const CholPlan& P, int s, std::vector<EAItem>& ea, std::vector<int>& ea_child) {
    const long long EA_TARGET = 16384;
    const Front& f = P.fronts[s];
    const int nchild = P.child_ptr[s + 1] - P.child_ptr[s];
    int cbeg = nchild > 0 ? 0 : f.nc;
    long long acc = 0;
    int c0 = cbeg;
    const size_t item0 = ea.size();
    for (int c = cbeg; c < f.nr; c++) {
        acc += f.nr - c;
        if (acc >= EA_TARGET || c == f.nr - 1) { ea.push_back({s, c0, c + 1, 0, 0}); c0 = c + 1; acc = 0; }
    }
    const int nit = (int)(ea.size() - item0);
    if (nchild == 0 || nit == 0) return;
    auto item_of = [&](int c) {          // the item whose range holds destination column c (ranges are consecutive from cbeg = 0)
        int lo = 0, hi = nit - 1;
        while (lo < hi) { const int mid = (lo + hi) >> 1; if (ea[item0 + mid].c1 <= c) lo = mid + 1; else hi = mid; }
        return lo;
    };
    for (int pass = 0; pass < 2; pass++) {
        if (pass == 1) {
            int off = (int)ea_child.size();
            for (int i = 0; i < nit; i++) { ea[item0 + i].q_off = off; off += ea[item0 + i].q_cnt; ea[item0 + i].q_cnt = 0; }
            ea_child.resize(off);
        }
        for (int q = 0; q < nchild; q++) {
            const Front& fc = P.fronts[P.child_idx[P.child_ptr[s] + q]];
            const int mc = fc.nr - fc.nc;
            const i32* rl = P.rel.data() + fc.reloff;
            int j = 0;
            while (j < mc) {
                EAItem& e = ea[item0 + item_of(rl[j])];
                if (pass == 1) ea_child[e.q_off + e.q_cnt] = q;
                e.q_cnt++;
                j = (int)(std::lower_bound(rl + j, rl + mc, e.c1) - rl);
            }
        }
    }
}
// gather chunks of front s (nch = ceil(nr / GATHER_ROWS) CTAs of k_fwd_gather): per chunk {first, count} in gq_child = the children
// with an update row inside the chunk, ascending; children with exactly one update row carry GATHER_SINGLE
static void gather_chunks_of_front(const CholPlan& P, int s, int nch, std::vector<int2>& gq, std::vector<int>& gq_child) {
    const int nchild = P.child_ptr[s + 1] - P.child_ptr[s];
    const size_t q0 = gq.size();
    gq.resize(q0 + nch, make_int2(0, 0));
    for (int pass = 0; pass < 2; pass++) {
        if (pass == 1) {
            int off = (int)gq_child.size();
            for (int i = 0; i < nch; i++) { gq[q0 + i].x = off; off += gq[q0 + i].y; gq[q0 + i].y = 0; }
            gq_child.resize(off);
        }
        for (int q = 0; q < nchild; q++) {
            const Front& fc = P.fronts[P.child_idx[P.child_ptr[s] + q]];
            const int mc = fc.nr - fc.nc;
            const i32* rl = P.rel.data() + fc.reloff;
            int j = 0;
            while (j < mc) {
                const int ch = rl[j] / GATHER_ROWS;
                int2& e = gq[q0 + ch];
                if (pass == 1) gq_child[e.x + e.y] = mc == 1 ? (q | GATHER_SINGLE) : q;
                e.y++;
                j = (int)(std::lower_bound(rl + j, rl + mc, (ch + 1) * GATHER_ROWS) - rl);
            }
        }
    }
}
// Test hook (b200s_chol_child_lists_check): both tables rebuilt for EVERY front of the plan and compared with the definition --
// a child is listed for an item / a chunk exactly when one of its relative indices falls into the item's column range / the
// chunk's row range, lists ascend, items tile [0 or nc, nr), the one-row flag is right.  Host only.
int chol_child_lists_check(const CholPlan& P) {
    for (int s = 0; s < (int)P.fronts.size(); s++) {
        const Front& f = P.fronts[s];
        const int nchild = P.child_ptr[s + 1] - P.child_ptr[s];
        std::vector<EAItem> ea; std::vector<int> ec;
        ea_items_of_front(P, s, ea, ec);
        int expect = nchild > 0 ? 0 : f.nc;
        for (const EAItem& e : ea) {
            if (e.front != s || e.c0 != expect || e.c1 <= e.c0) { set_last_error("child lists: items do not tile the columns"); return ST_INVALID; }
            expect = e.c1;
            int t = 0;
            for (int q = 0; q < nchild; q++) {
                const Front& fc = P.fronts[P.child_idx[P.child_ptr[s] + q]];
                const i32* rl = P.rel.data() + fc.reloff;
                bool hit = false;
                for (int i = 0; i < fc.nr - fc.nc; i++) hit |= rl[i] >= e.c0 && rl[i] < e.c1;
                if (hit) {
                    if (t >= e.q_cnt || ec[e.q_off + t] != q) { set_last_error("child lists: extend-add item misses a child"); return ST_INVALID; }
                    t++;
                }
            }
            if (t != e.q_cnt) { set_last_error("child lists: extend-add item lists a child without a column in its range"); return ST_INVALID; }
        }
        if (expect != f.nr && !(ea.empty() && expect >= f.nr)) { set_last_error("child lists: items do not reach the last column"); return ST_INVALID; }
        const int nch = (f.nr + GATHER_ROWS - 1) / GATHER_ROWS;
        std::vector<int2> gq; std::vector<int> gc;
        gather_chunks_of_front(P, s, nch, gq, gc);
        for (int ch = 0; ch < nch; ch++) {
            int t = 0;
            for (int q = 0; q < nchild; q++) {
                const Front& fc = P.fronts[P.child_idx[P.child_ptr[s] + q]];
                const i32* rl = P.rel.data() + fc.reloff;
                const int mc = fc.nr - fc.nc;
                bool hit = false;
                for (int i = 0; i < mc; i++) hit |= rl[i] / GATHER_ROWS == ch;
                if (hit) {
                    if (t >= gq[ch].y || gc[gq[ch].x + t] != (mc == 1 ? (q | GATHER_SINGLE) : q)) { set_last_error("child lists: gather chunk misses a child"); return ST_INVALID; }
                    t++;
                }
            }
            if (t != gq[ch].y) { set_last_error("child lists: gather chunk lists a child without a row in it"); return ST_INVALID; }
        }
    }
    return ST_OK;
}

int CholDevice::init() {
    const CholPlan& P = *plan;
    const bool dbg = getenv("B200S_DEBUG") != nullptr;
    auto t_start = std::chrono::steady_clock::now();
    auto lap = [&](const char* what) {
        if (!dbg) return;
        auto now = std::chrono::steady_clock::now();
        fprintf(stderr, "[b200s chol init] %-28s %8.1f ms\n", what, std::chrono::duration<double, std::milli>(now - t_start).count());
        t_start = now;
    };
    CUDA_TRY(cudaSetDevice(device));
    CUDA_TRY(cudaFree(0));
    lap("context");
    use_graphs = getenv("B200S_NO_GRAPH") == nullptr;
    if (const char* e = getenv("B200S_SOLVE_MEDIUM_NR")) solve_medium_nr = atoi(e);
    if (const char* e = getenv("B200S_UPDATE_TMA")) upd_tma = atoi(e) != 0;
    {   // the main stream carries the latency-bound panel chain: its CTAs must get the SM slots that the bulk update
        // kernels on stream2 free up, ahead of that kernel's own queued CTAs
        int prio_lo = 0, prio_hi = 0;
        CUDA_TRY(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
        CUDA_TRY(cudaStreamCreateWithPriority(&stream, cudaStreamNonBlocking, prio_hi));
        CUDA_TRY(cudaStreamCreateWithPriority(&stream2, cudaStreamNonBlocking, prio_lo));
    }
    CUDA_TRY(cudaEventCreateWithFlags(&evP, cudaEventDisableTiming));
    CUDA_TRY(cudaEventCreateWithFlags(&evB, cudaEventDisableTiming));
    for (auto& e : ev) CUDA_TRY(cudaEventCreate(&e));
    const int ns = (int)P.fronts.size();
    std::vector<FrontD> hf(ns);
    std::vector<int> inv_front, inv_kb;
    for (int s = 0; s < ns; s++) {
        const Front& f = P.fronts[s];
        FrontD d;
        d.loff = f.loff; d.uoff = f.uoff; d.reloff = f.reloff; d.rowptr = f.rowptr;
        d.col0 = f.col0; d.nc = f.nc; d.nr = f.nr; d.ld = f.ld;
        d.nchild = P.child_ptr[s + 1] - P.child_ptr[s]; d.childptr = P.child_ptr[s];
        d.parent = f.parent; d.level = f.level;
        d.ioff = -1;
        if (f.nr > SMALL_NR) {
            d.ioff = (long long)inv_front.size() * MINV_BLK;
            for (int kb = 0; kb < (f.nc + NB - 1) / NB; kb++) { inv_front.push_back(s); inv_kb.push_back(kb); }
        }
        hf[s] = d;
    }
    ninvblk = (int)inv_front.size();
    int rc;
    if ((rc = upload(&dF, hf.data(), hf.size()))) return rc;
    if ((rc = upload(&dinv_front, inv_front.data(), inv_front.size()))) return rc;
    if ((rc = upload(&dinv_kb, inv_kb.data(), inv_kb.size()))) return rc;
    if ((rc = upload(&drows, P.rows.data(), P.rows.size()))) return rc;
    if ((rc = upload(&drel, P.rel.data(), P.rel.size()))) return rc;
    if ((rc = upload(&dchild, P.child_idx.data(), P.child_idx.size()))) return rc;
    if ((rc = upload(&dperm, P.perm.data(), P.perm.size()))) return rc;
    if ((rc = upload(&dlevel_fronts, P.level_fronts.data(), P.level_fronts.size()))) return rc;
    {
        std::vector<long long> am(P.amap.begin(), P.amap.end());
        if ((rc = upload(&damap, am.data(), am.size()))) return rc;
    }
    lap("plan upload");
    CUDA_TRY(pool_malloc((void**)&dL, std::max<i64>(P.lsize, 1) * sizeof(double)));
    CUDA_TRY(pool_malloc((void**)&dW, std::max<i64>(P.wsize, 1) * sizeof(double)));
    CUDA_TRY(pool_malloc((void**)&dval, std::max<i64>(P.nnzA, 1) * sizeof(double)));
    CUDA_TRY(pool_malloc((void**)&dminor, sizeof(int)));
    CUDA_TRY(pool_malloc((void**)&dMinv, std::max<size_t>((size_t)ninvblk * MINV_BLK, 1) * sizeof(double)));
    total_bytes += (i64)ninvblk * MINV_BLK * sizeof(double);
    CUDA_TRY(pool_malloc((void**)&downed, std::max<size_t>(hf.size(), 1)));
    CUDA_TRY(cudaMemset(downed, 1, std::max<size_t>(hf.size(), 1)));
    total_bytes += (P.lsize + P.wsize + P.nnzA) * sizeof(double);

    lap("cudaMalloc L/W/val");
    // ---- schedule
    std::vector<int>& sched = hsched;  // group arrays: [front ids...][prefix...] (host copy kept: the reach-restricted solve filters it)
    std::vector<EAItem> ea;
    std::vector<int> ea_child;
    std::vector<int2> gq;          // k_fwd_gather: per chunk {first, count} in gq_child
    std::vector<int> gq_child;
    levels.resize(P.nlevels);
    sgroups.assign(1, SolveGroups());
    memset(&sgroups[0], 0, sizeof(SolveGroups));
    for (int l = 0; l < P.nlevels; l++) {
        LevelSched& LS = levels[l];
        std::vector<int> smalls[LevelSched::NSMALL], bigs;
        LS.ea_off = (int)ea.size();
        for (int q = P.level_ptr[l]; q < P.level_ptr[l + 1]; q++) {
            const int s = P.level_fronts[q];
            const Front& f = P.fronts[s];
            ea_items_of_front(P, s, ea, ea_child);
            if (f.nr <= SMALL_NR) {
                int cls = f.nr <= 32 ? 0 : (f.nr <= 64 ? 1 : (f.nr <= 96 ? 2 : (f.nr <= 118 ? 3 : 4)));
                smalls[cls].push_back(s);
                LS.small_maxnr[cls] = std::max(LS.small_maxnr[cls], f.nr);
            } else bigs.push_back(s);
        }
        LS.ea_cnt = (int)ea.size() - LS.ea_off;
        for (int c = 0; c < LevelSched::NSMALL; c++) {
            LS.small_off[c] = (int)sched.size();
            LS.small_cnt[c] = (int)smalls[c].size();
            sched.insert(sched.end(), smalls[c].begin(), smalls[c].end());
        }
        // SOLVE phase only: large fronts with one block column and few rows (the deep levels of a 3-D problem hold thousands of
        // them: 100^3 has 3653 in one level) go with the small fronts -- one CTA per front (k_fwd / k_bwd take any nr) instead of a
        // 128 KB tile CTA per 128 rows at one CTA per SM (their two block-step launches took 135 / 278 us per level and sweep).
        // Row limit swept on the B200 (100^3 / 64^3 solve through the host API): none 8.77 / 2.63 ms, 200: 8.31 / 2.58, 256: 8.16 /
        // 2.48, 320: 8.12 / 2.50, 448: 8.28 / 2.58, 640: 8.37 / 2.91 (beyond ~320 rows the one-CTA kernels are the slower ones).
        // Their ids follow the small fronts' in the schedule array, so the small-front list of the solves simply grows.
        std::vector<int> mediums, sbigs;
        for (int s : bigs) {
            const Front& f = P.fronts[s];
            if (f.nc <= NB && f.nr <= solve_medium_nr) mediums.push_back(s); else sbigs.push_back(s);
        }
        sched.insert(sched.end(), mediums.begin(), mediums.end());
        int maxblk = 0;
        for (int s : bigs) maxblk = std::max(maxblk, (P.fronts[s].nc + NB - 1) / NB);
        LS.panel.resize(maxblk);
        LS.upd.resize(maxblk);
        LS.updA.resize(maxblk);
        LS.updB.resize(maxblk);
        LS.updN.resize(maxblk);
        LS.updF.resize(maxblk);
        LS.updFA.resize(maxblk);
        LS.updFB.resize(maxblk);
        auto emit = [&](Launch& la, const std::vector<int>& fr, const std::vector<int>& cnt) {
            la.goff = (int)sched.size();
            la.ng = (int)fr.size();
            sched.insert(sched.end(), fr.begin(), fr.end());
            int run = 0;
            for (int c : cnt) { sched.push_back(run); run += c; }
            sched.push_back(run);
            la.ctas = run;
            la.sgi = -1;
            if (!fr.empty() && (int)fr.size() <= MAXG) {
                SolveGroups G;
                memset(&G, 0, sizeof(G));
                G.ng = (int)fr.size();
                int acc = 0;
                for (int i = 0; i < G.ng; i++) {
                    const FrontD& d = hf[fr[i]];
                    G.prefix[i] = acc; acc += cnt[i];
                    G.fr[i] = FrontS{d.loff, d.rowptr, d.ioff, d.uoff, d.nc, d.nr, d.ld, d.col0, fr[i], 0};
                }
                G.prefix[G.ng] = acc;
                la.sgi = (int)sgroups.size();
                sgroups.push_back(G);
            }
        };
        if (!sbigs.empty()) {
            std::vector<int> cnt;
            for (int s : sbigs) cnt.push_back((P.fronts[s].nr + GATHER_ROWS - 1) / GATHER_ROWS);
            emit(LS.gfwd, sbigs, cnt);
            // per gather chunk (= CTA of k_fwd_gather, in launch order) the children with an update row inside it
            LS.gq_off = (int)gq.size();
            for (size_t fi = 0; fi < sbigs.size(); fi++) gather_chunks_of_front(P, sbigs[fi], cnt[fi], gq, gq_child);
            emit(LS.sbig, sbigs, std::vector<int>(sbigs.size(), 1));
        }
        LS.small_all_off = LS.small_off[0];
        LS.small_all_cnt = (int)mediums.size();
        for (int c = 0; c < LevelSched::NSMALL; c++) LS.small_all_cnt += LS.small_cnt[c];
        int maxblk_s = 0;
        for (int s : sbigs) maxblk_s = std::max(maxblk_s, (P.fronts[s].nc + NB - 1) / NB);
        LS.sfwd.resize(maxblk_s);
        LS.sbwd.resize(maxblk_s);
        for (int kb = 0; kb < maxblk_s; kb++) {
            std::vector<int> fs, cf, cb2;
            for (int s : sbigs) {
                const Front& f = P.fronts[s];
                const int nblk = (f.nc + NB - 1) / NB;
                if (kb >= nblk) continue;
                const int w = std::min(NB, f.nc - kb * NB);
                const int below = f.nr - ((kb * NB + w) & ~1);      // row tiles of the forward update start at an even row
                fs.push_back(s);
                cf.push_back((below + SOLVE_FT - 1) / SOLVE_FT);
                cb2.push_back(std::max(0, (f.nr + NB - 1) / NB - (kb * NB + w) / NB));      // backward: the 128-row pairs that hold rows >= rb
            }
            emit(LS.sfwd[kb], fs, cf);
            emit(LS.sbwd[kb], fs, cb2);
            max_solve_ctas = std::max(max_solve_ctas, LS.sbwd[kb].ctas);
        }
        for (int kb = 0; kb < maxblk; kb++) {
            std::vector<int> fr, cp, fu, cu, fuA, cuA, fuB, cuB, fuN, cuN, fuF, cuF, fuFA, cuFA, fuFB, cuFB;
            // panel CTAs: one per front for the diagonal block + solver CTAs that each factor the block redundantly
            // and then walk over several 64-row tiles; the solver CTAs of a launch are capped near one wave (148 SMs)
            long long tiles_total = 0;
            int nact = 0;
            for (int s : bigs) {
                const Front& f = P.fronts[s];
                if (kb >= (f.nc + NB - 1) / NB) continue;
                const int w = std::min(NB, f.nc - kb * NB);
                tiles_total += (f.nr - (kb * NB + w) + TR - 1) / TR;
                nact++;
            }
            const long long budget = std::max<long long>(148 - nact, nact);
            for (int s : bigs) {
                const Front& f = P.fronts[s];
                const int nblk = (f.nc + NB - 1) / NB;
                if (kb >= nblk) continue;
                const int w = std::min(NB, f.nc - kb * NB);
                const int below = f.nr - (kb * NB + w);
                const int ntiles = (below + TR - 1) / TR;
                int nsolve = ntiles;
                if (tiles_total > budget && ntiles > 0)
                    nsolve = (int)std::min<long long>(ntiles, std::max<long long>(1, (budget * ntiles + tiles_total - 1) / tiles_total));
                fr.push_back(s);
                cp.push_back(1 + nsolve);
                const int nrt = (f.nr + BT - 1) / BT, ncolt = (f.nc + BTN - 1) / BTN;
                long long tiles = 0, tilesA = 0;
                for (int cj = 2 * (kb + 1); cj < ncolt; cj++) {
                    tiles += nrt - (cj >> 1);
                    if (cj < 2 * (kb + 1) + 2) tilesA += nrt - (cj >> 1);
                }
                if (tiles > 0) { fu.push_back(s); cu.push_back((int)tiles); }
                if (tilesA > 0) { fuA.push_back(s); cuA.push_back((int)tilesA); }
                if (tiles - tilesA > 0) { fuB.push_back(s); cuB.push_back((int)(tiles - tilesA)); }
                // two-level: near = column tiles up to the end of kb's super-block, far (only after the super-block's
                // last panel) = everything right of the super-block
                const int sb_end = (kb / SUPER_NB + 1) * SUPER_NB;          // first block column outside
                long long tilesN = 0, tilesF = 0;
                for (int cj = 2 * (kb + 1); cj < std::min(ncolt, 2 * sb_end); cj++) tilesN += nrt - (cj >> 1);
                long long tilesFA = 0;
                if (kb + 1 == sb_end)
                    for (int cj = 2 * sb_end; cj < ncolt; cj++) {
                        tilesF += nrt - (cj >> 1);
                        if (cj < 2 * (sb_end + SUPER_NB)) tilesFA += nrt - (cj >> 1);
                    }
                if (tilesN > 0) { fuN.push_back(s); cuN.push_back((int)tilesN); }
                if (tilesF > 0) { fuF.push_back(s); cuF.push_back((int)tilesF); }
                if (tilesFA > 0) { fuFA.push_back(s); cuFA.push_back((int)tilesFA); }
                if (tilesF - tilesFA > 0) { fuFB.push_back(s); cuFB.push_back((int)(tilesF - tilesFA)); }
            }
            emit(LS.panel[kb], fr, cp);
            emit(LS.upd[kb], fu, cu);
            emit(LS.updA[kb], fuA, cuA);
            emit(LS.updB[kb], fuB, cuB);
            emit(LS.updN[kb], fuN, cuN);
            emit(LS.updF[kb], fuF, cuF);
            emit(LS.updFA[kb], fuFA, cuFA);
            emit(LS.updFB[kb], fuFB, cuFB);
        }
        {
            std::vector<int> fr, cnt;
            for (int s : bigs) {
                const Front& f = P.fronts[s];
                if (f.nr == f.nc) continue;
                const int r0 = f.nc - (f.nc & 1), mu = f.nr - r0;
                const int T = (mu + BT - 1) / BT, ncolt = (mu + BTN - 1) / BTN;
                long long tiles = 0;
                for (int cj = 0; cj < ncolt; cj++) tiles += T - (cj >> 1);
                fr.push_back(s);
                cnt.push_back((int)tiles);
            }
            emit(LS.syrk, fr, cnt);
        }
    }
    {   // persistent sweeps: CTAs of a level dealt to its large fronts by panel area, one CTA per SM
        if (const char* e = getenv("B200S_SOLVE_PERSIST")) persist_default = atoi(e) & 3;
        int nsm = 0, occ_f = 0, occ_b = 0;
        CUDA_TRY(cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, device));
        CUDA_TRY(cudaFuncSetAttribute(k_fwd_persist, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_PERSIST));
        CUDA_TRY(cudaFuncSetAttribute(k_bwd_persist, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_PERSIST));
        CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_f, k_fwd_persist, 256, SMEM_PERSIST));
        CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_b, k_bwd_persist, 256, SMEM_PERSIST));
        if (occ_f < 1 || occ_b < 1 || nsm < 1) persist_hw = 0;        // every CTA of a launch must be resident (they wait for each other)
        persist_default &= persist_hw;
        persist_mode = persist_default;
        // levels with more large fronts than this keep the launch-per-step kernels: their launches are wide (many fronts) and few
        // (short fronts), which the level-wide kernels handle at memory speed, while 1-2 CTAs per front would serialise
        // each front's steps (measured on 100^3: 32 -> 9.4 ms per solve, 296 -> 9.8 ms)
        int maxf = 32;
        if (const char* e = getenv("B200S_PERSIST_MAXF")) maxf = std::max(0, std::min(atoi(e), 2 * nsm));
        if (const char* e = getenv("B200S_PERSIST_RECT")) persist_rect = atoi(e) != 0;
        if (const char* e = getenv("B200S_PERSIST_COOP")) persist_coop = atoi(e) != 0;
        { int coop = 0; cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, device); if (!coop) persist_coop = false; }
        int max_rect = 0;
        std::vector<PersistFront> hpf;
        std::vector<int> hcmap;
        std::vector<std::pair<size_t, size_t>> offs;       // per persistent level: first record, first map entry
        for (int l = 0; l < P.nlevels && persist_hw; l++) {
            LevelSched& LS = levels[l];
            if (LS.sbig.ng < 1 || LS.sbig.ng > maxf) continue;
            const int ng = LS.sbig.ng;
            const int* fr = sched.data() + LS.sbig.goff;          // the fronts of the level that the solves treat as large
            const bool shared = ng <= nsm;                        // else: one CTA per front, several fronts per CTA
            double area = 0;
            for (int i = 0; i < ng; i++) area += (double)hf[fr[i]].nr * hf[fr[i]].nc;
            int cta = 0, rtiles = 0;
            offs.push_back({hpf.size(), hcmap.size()});
            for (int i = 0; i < ng; i++) {
                const FrontD& d = hf[fr[i]];
                const int npairs = (d.nr + NB - 1) / NB, nblk = (d.nc + NB - 1) / NB;
                const int extra = shared ? (int)((double)(nsm - ng) * ((double)d.nr * d.nc / area)) : 0;
                PersistFront pf;
                memset(&pf, 0, sizeof(pf));
                pf.f = FrontS{d.loff, d.rowptr, d.ioff, d.uoff, d.nc, d.nr, d.ld, d.col0, fr[i], 0};
                pf.cta0 = cta;
                pf.ncta = 1 + std::max(0, std::min(npairs - 1, extra));
                pf.sync0 = nsync;
                pf.rect0 = rtiles;
                rtiles += (npairs - nblk) * nblk;
                if (shared) for (int q = 0; q < pf.ncta; q++) hcmap.push_back(i);
                cta += pf.ncta;
                nsync += nblk;
                hpf.push_back(pf);
            }
            PersistLevel pl;
            pl.fr = nullptr; pl.cmap = nullptr; pl.nf = ng; pl.shared = shared ? 1 : 0;
            LS.pgi = (int)plevels.size();
            LS.pctas = shared ? cta : nsm;
            LS.prect = persist_rect ? rtiles : 0;
            max_rect = std::max(max_rect, LS.prect);
            plevels.push_back(pl);
        }
        if (max_rect > 0) {
            CUDA_TRY(pool_malloc((void**)&drect, (size_t)max_rect * NB * sizeof(double)));
            total_bytes += (size_t)max_rect * NB * sizeof(double);
            CUDA_TRY(cudaFuncSetAttribute(k_bwd_rect, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BUPD));
        }
        if (!plevels.empty()) {
            if ((rc = upload(&dpfront, hpf.data(), hpf.size()))) return rc;
            if ((rc = upload(&dpcmap, hcmap.data(), hcmap.size()))) return rc;
            for (size_t i = 0; i < plevels.size(); i++) { plevels[i].fr = dpfront + offs[i].first; plevels[i].cmap = dpcmap + offs[i].second; }
        }
        if (!plevels.empty()) {
            CUDA_TRY(pool_malloc((void**)&dsync, (size_t)2 * nsync * sizeof(int)));
            // one mapped error word per process (a pinned allocation per factor object costs milliseconds, which the small
            // configurations -- a new KKT object per LP -- would pay every time); a set word is a schedule bug, whoever sees it
            {
                static std::mutex mu;
                static int* g_err = nullptr;
                std::lock_guard<std::mutex> lk(mu);
                if (!g_err) {
                    CUDA_TRY(cudaHostAlloc((void**)&g_err, sizeof(int), cudaHostAllocMapped | cudaHostAllocPortable));
                    *g_err = 0;
                }
                herr = g_err;
            }
            if (getenv("B200S_PERSIST_DBG")) { CUDA_TRY(cudaHostAlloc((void**)&persist_dbg, 8192 * sizeof(long long), cudaHostAllocMapped)); memset(persist_dbg, 0, 8192 * sizeof(long long)); }
            total_bytes += (size_t)2 * nsync * sizeof(int);
        }
    }
    {
        int maxng = 1;
        for (auto& LS : levels) for (auto& la : LS.panel) maxng = std::max(maxng, la.ng);
        CUDA_TRY(pool_malloc((void**)&ddiag, (size_t)maxng * NB * NB * sizeof(double)));
        total_bytes += (size_t)maxng * NB * NB * sizeof(double);
    }
    lap("schedule build");
    if ((rc = upload(&dsched, sched.data(), sched.size()))) return rc;
    if ((rc = upload(&dea, ea.data(), ea.size()))) return rc;
    if ((rc = upload(&dea_child, ea_child.data(), ea_child.size()))) return rc;
    if ((rc = upload(&dgq, gq.data(), gq.size()))) return rc;
    if ((rc = upload(&dgq_child, gq_child.data(), gq_child.size()))) return rc;
    CUDA_TRY(cudaFuncSetAttribute(k_panel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_PANEL));
    CUDA_TRY(cudaFuncSetAttribute(k_update<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_UPDATE));
    CUDA_TRY(cudaFuncSetAttribute(k_update<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_UPDATE));
    CUDA_TRY(cudaFuncSetAttribute(k_update<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_UPDATE));
    CUDA_TRY(cudaFuncSetAttribute(k_panel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_PANEL));
    CUDA_TRY(cudaFuncSetAttribute(k_update<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_UPDATE));
    CUDA_TRY(cudaFuncSetAttribute(k_fwd_diag, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_SDIAG));
    CUDA_TRY(cudaFuncSetAttribute(k_bwd_diag, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BDIAG));
    CUDA_TRY(cudaFuncSetAttribute(k_fwd_upd<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_FUPD));
    CUDA_TRY(cudaFuncSetAttribute(k_fwd_upd<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_FUPD));
    CUDA_TRY(cudaFuncSetAttribute(k_bwd_upd<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BUPD));
    CUDA_TRY(cudaFuncSetAttribute(k_bwd_upd<BWD_CG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BUPD));
    CUDA_TRY(cudaFuncSetAttribute(k_small_front<256, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)((SMALL_NR | 1) * SMALL_NR * sizeof(double))));
    CUDA_TRY(cudaFuncSetAttribute(k_small_front<128, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65 * 64 * 8));
    CUDA_TRY(cudaFuncSetAttribute(k_small_front<256, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)((SMALL_NR | 1) * SMALL_NR * sizeof(double))));
    CUDA_TRY(cudaFuncSetAttribute(k_small_front<128, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65 * 64 * 8));
    lap("schedule upload + attributes");
    return ST_OK;
}

int CholDevice::factor_begin(const double* val, bool on_device) {
    const CholPlan& P = *plan;
    CUDA_TRY(cudaSetDevice(device));
    numeric = false;
    minv_valid = false;
    diagL_valid = false;
    if (ldl && !dsgn) {
        const size_t cnt = (size_t)P.n + 2 * NB;          // k_update reads up to BK entries past a front's last column
        std::vector<double> ones(cnt, 1.0);
        CUDA_TRY(pool_malloc((void**)&dsgn, cnt * sizeof(double)));
        CUDA_TRY(cudaMemcpy(dsgn, ones.data(), cnt * sizeof(double), cudaMemcpyHostToDevice));
    }
    CUDA_TRY(cudaEventRecord(ev[0], stream));
    const double* dv = val;
    if (!on_device) {
        if (P.nnzA) CUDA_TRY(cudaMemcpyAsync(dval, val, P.nnzA * sizeof(double), cudaMemcpyHostToDevice, stream));
        dv = dval;
    }
    CUDA_TRY(cudaEventRecord(ev[1], stream));
    k_set_int<<<1, 1, 0, stream>>>(dminor, 0x7fffffff);      // (a kernel, not a copy from the stack: this sequence may be recorded into a caller's graph)
    if (P.lsize) CUDA_TRY(cudaMemsetAsync(dL, 0, P.lsize * sizeof(double), stream));
    if (P.nnzA) {
        int blocks = (int)std::min<i64>((P.nnzA + 255) / 256, 148 * 16);
        k_scatter_A<<<blocks, 256, 0, stream>>>(dv, damap, P.nnzA, dL);
    }
    CUDA_TRY(cudaEventRecord(ev[2], stream));
    pe = 0;
    spans.clear();
    return ST_OK;
}

// launch sites shared by the LL' kernels and their signed (LDL') instantiations, which also take the pivot-sign array
#define LAUNCH_SGN(KERN, GRID, BLOCK, SMEM, STREAM, ...)                                         \
    do {                                                                                          \
        if (ldl) KERN<true><<<GRID, BLOCK, SMEM, STREAM>>>(__VA_ARGS__, dsgn);                    \
        else KERN<false><<<GRID, BLOCK, SMEM, STREAM>>>(__VA_ARGS__, nullptr);                    \
    } while (0)
#define LAUNCH_UPD(GRID, BLOCK, SMEM, STREAM, ...)                                                \
    do {                                                                                          \
        if (ldl) { if (upd_tma) k_update<true, true><<<GRID, BLOCK, SMEM, STREAM>>>(__VA_ARGS__, syrk_split, dsgn);       \
                   else k_update<true, false><<<GRID, BLOCK, SMEM, STREAM>>>(__VA_ARGS__, syrk_split, dsgn); }            \
        else { if (upd_tma) k_update<false, true><<<GRID, BLOCK, SMEM, STREAM>>>(__VA_ARGS__, syrk_split, nullptr);       \
               else k_update<false, false><<<GRID, BLOCK, SMEM, STREAM>>>(__VA_ARGS__, syrk_split, nullptr); }            \
    } while (0)
#define LAUNCH_SMALL(T, GRID, BLOCK, SMEM, STREAM, ...)                                          \
    do {                                                                                          \
        if (ldl) k_small_front<T, true><<<GRID, BLOCK, SMEM, STREAM>>>(__VA_ARGS__, dsgn);        \
        else k_small_front<T, false><<<GRID, BLOCK, SMEM, STREAM>>>(__VA_ARGS__, nullptr);        \
    } while (0)

int CholDevice::set_syrk_split(const unsigned char* own, const int* lo, const int* hi, const long long* base, double* scratch) {
    CUDA_TRY(cudaSetDevice(device));
    const size_t ns = std::max<size_t>(plan->fronts.size(), 1);
    if (!own) { syrk_split = {nullptr, nullptr, nullptr, nullptr, nullptr}; return ST_OK; }
    if (!lo || !hi || !base) return ST_INVALID;
    if (!dsy_own) {
        CUDA_TRY(cudaMalloc((void**)&dsy_own, ns));
        CUDA_TRY(cudaMalloc((void**)&dsy_lo, ns * sizeof(int)));
        CUDA_TRY(cudaMalloc((void**)&dsy_hi, ns * sizeof(int)));
        CUDA_TRY(cudaMalloc((void**)&dsy_base, ns * sizeof(long long)));
    }
    CUDA_TRY(cudaMemcpy(dsy_own, own, plan->fronts.size(), cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(dsy_lo, lo, plan->fronts.size() * sizeof(int), cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(dsy_hi, hi, plan->fronts.size() * sizeof(int), cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(dsy_base, base, plan->fronts.size() * sizeof(long long), cudaMemcpyHostToDevice));
    syrk_split = {dsy_own, dsy_lo, dsy_hi, dsy_base, scratch};
    return ST_OK;
}

int CholDevice::factor_level(int l, int phase) {
    const CholPlan& P = *plan;
    if (l < 0 || l >= P.nlevels) return ST_INVALID;
    CUDA_TRY(cudaSetDevice(device));
    if (!(phase & 1)) {          // the Schur complements alone (the panels were factored by an earlier call, here or elsewhere)
        const LevelSched& LS2 = levels[l];
        if ((phase & 2) && LS2.syrk.ctas)
            LAUNCH_UPD(LS2.syrk.ctas, UPD_THREADS, SMEM_UPDATE, stream,
                       sgroups[std::max(LS2.syrk.sgi, 0)], dsched + LS2.syrk.goff, dsched + LS2.syrk.goff + LS2.syrk.ng,
                       LS2.syrk.ng, 1, 0, dF, dL, dW, downed);
        CUDA_TRY(cudaGetLastError());
        return ST_OK;
    }
    auto prof_begin = [&](int cls) {
        if (!profiling) return;
        while (pev.size() < pe + 2) { cudaEvent_t e; cudaEventCreate(&e); pev.push_back(e); }
        cudaEventRecord(pev[pe], stream);
        spans.push_back({cls, pe, pe + 1});
        pe += 2;
    };
    auto prof_end = [&]() { if (profiling) cudaEventRecord(pev[spans.back().e1], stream); };
    const LevelSched& LS = levels[l];
    if (LS.ea_cnt) {
        prof_begin(0);
        k_extend_add<<<LS.ea_cnt, 256, 0, stream>>>(dea + LS.ea_off, dF, dchild, drel, dL, dW, downed, dea_child);
        prof_end();
    }
    prof_begin(1);
    if (LS.small_cnt[0])
        LAUNCH_SMALL(64, LS.small_cnt[0], 64, (size_t)(LS.small_maxnr[0] | 1) * LS.small_maxnr[0] * 8, stream,
                         dsched + LS.small_off[0], dF, dL, dW, dminor, opts.dbound, downed);
    if (LS.small_cnt[1])
        LAUNCH_SMALL(128, LS.small_cnt[1], 128, (size_t)(LS.small_maxnr[1] | 1) * LS.small_maxnr[1] * 8, stream,
                         dsched + LS.small_off[1], dF, dL, dW, dminor, opts.dbound, downed);
    for (int c = 2; c < LevelSched::NSMALL; c++)
        if (LS.small_cnt[c])
            LAUNCH_SMALL(256, LS.small_cnt[c], 256, (size_t)(LS.small_maxnr[c] | 1) * LS.small_maxnr[c] * 8, stream,
                             dsched + LS.small_off[c], dF, dL, dW, dminor, opts.dbound, downed);
    prof_end();
    // Lookahead of depth one: after panel kb, part A of its trailing update (the next block column only) runs on
    // the main stream, then panel kb+1; part B (all other column tiles) runs on stream2 concurrently with panel
    // kb+1, whose few latency-bound CTAs would otherwise leave the GPU idle.  Part A of step kb+1 waits for part B
    // of step kb (same C tiles).  With profiling on everything is serialised on the main stream.
    const bool lookahead = !profiling;
    const bool two_level = getenv("B200S_ONE_LEVEL") == nullptr;
    bool pendingB = false;
    for (size_t kb = 0; kb < LS.panel.size(); kb++) {
        const Launch& lp = LS.panel[kb];
        if (lp.ctas) {
            prof_begin(2);
            LAUNCH_SGN(k_panel, lp.ctas, 256, SMEM_PANEL, stream,
                       sgroups[std::max(lp.sgi, 0)], dsched + lp.goff, dsched + lp.goff + lp.ng, lp.ng, (int)kb,
                       dF, dL, ddiag, dminor, opts.dbound, downed);
            k_diag_writeback<<<lp.ng, 256, 0, stream>>>(dsched + lp.goff, (int)kb, dF, dL, ddiag, downed);
            prof_end();
        }
        if (two_level) {
            const Launch& ln = LS.updN[kb];
            const Launch& lf = LS.updF[kb];
            prof_begin(3);
            if (ln.ctas)
                LAUNCH_UPD(ln.ctas, UPD_THREADS, SMEM_UPDATE, stream,
                       sgroups[std::max(ln.sgi, 0)], dsched + ln.goff, dsched + ln.goff + ln.ng, ln.ng, 4,
                       (int)kb, dF, dL, dW, downed);
            if (lf.ctas && !lookahead)
                LAUNCH_UPD(lf.ctas, UPD_THREADS, SMEM_UPDATE, stream,
                       sgroups[std::max(lf.sgi, 0)], dsched + lf.goff, dsched + lf.goff + lf.ng, lf.ng, 5,
                       (int)kb, dF, dL, dW, downed);
            if (lf.ctas && lookahead) {
                // Far update with look-ahead: the column tiles of the NEXT super-block (part A) run here and gate its
                // panels; all other column tiles (part B, the bulk of the flops) run on stream2 under the next
                // super-block's four panel + near-update steps, whose few latency-bound CTAs leave the GPU idle.
                // Part A of the next far update touches tiles that this part B writes, so it waits for it.
                const Launch& lfa = LS.updFA[kb];
                const Launch& lfb = LS.updFB[kb];
                if (lfb.ctas) CUDA_TRY(cudaEventRecord(evP, stream));            // panels and near updates of this super-block
                if (pendingB) { CUDA_TRY(cudaStreamWaitEvent(stream, evB, 0)); pendingB = false; }
                if (lfa.ctas)
                    LAUNCH_UPD(lfa.ctas, UPD_THREADS, SMEM_UPDATE, stream,
                       sgroups[std::max(lfa.sgi, 0)], dsched + lfa.goff, dsched + lfa.goff + lfa.ng, lfa.ng, 5,
                       (int)kb, dF, dL, dW, downed);
                if (lfb.ctas) {
                    CUDA_TRY(cudaStreamWaitEvent(stream2, evP, 0));
                    LAUNCH_UPD(lfb.ctas, UPD_THREADS, SMEM_UPDATE, stream2,
                       sgroups[std::max(lfb.sgi, 0)], dsched + lfb.goff, dsched + lfb.goff + lfb.ng, lfb.ng, 6,
                       (int)kb, dF, dL, dW, downed);
                    CUDA_TRY(cudaEventRecord(evB, stream2));
                    pendingB = true;
                }
            }
            prof_end();
            continue;
        }
        if (!lookahead) {
            const Launch& lu = LS.upd[kb];
            if (lu.ctas) {
                prof_begin(3);
                LAUNCH_UPD(lu.ctas, UPD_THREADS, SMEM_UPDATE, stream,
                       sgroups[std::max(lu.sgi, 0)], dsched + lu.goff, dsched + lu.goff + lu.ng, lu.ng, 0,
                       (int)kb, dF, dL, dW, downed);
                prof_end();
            }
            continue;
        }
        const Launch& la = LS.updA[kb];
        const Launch& lb = LS.updB[kb];
        if (lb.ctas) CUDA_TRY(cudaEventRecord(evP, stream));             // panel kb is complete
        if (pendingB) { CUDA_TRY(cudaStreamWaitEvent(stream, evB, 0)); pendingB = false; }   // part B of step kb-1
        if (la.ctas)
            LAUNCH_UPD(la.ctas, UPD_THREADS, SMEM_UPDATE, stream,
                       sgroups[std::max(la.sgi, 0)], dsched + la.goff, dsched + la.goff + la.ng, la.ng, 2,
                       (int)kb, dF, dL, dW, downed);
        if (lb.ctas) {
            CUDA_TRY(cudaStreamWaitEvent(stream2, evP, 0));
            LAUNCH_UPD(lb.ctas, UPD_THREADS, SMEM_UPDATE, stream2,
                       sgroups[std::max(lb.sgi, 0)], dsched + lb.goff, dsched + lb.goff + lb.ng, lb.ng, 3,
                       (int)kb, dF, dL, dW, downed);
            CUDA_TRY(cudaEventRecord(evB, stream2));
            pendingB = true;
        }
    }
    if (pendingB) CUDA_TRY(cudaStreamWaitEvent(stream, evB, 0));
    if (LS.syrk.ctas && (phase & 2)) {
        prof_begin(3);
        LAUNCH_UPD(LS.syrk.ctas, UPD_THREADS, SMEM_UPDATE, stream,
                       sgroups[std::max(LS.syrk.sgi, 0)], dsched + LS.syrk.goff, dsched + LS.syrk.goff + LS.syrk.ng,
                       LS.syrk.ng, 1, 0, dF, dL, dW, downed);
        prof_end();
    }
    CUDA_TRY(cudaGetLastError());
    return ST_OK;
}

int CholDevice::factor_end(i64* minor, CholTimes* times) {
    const CholPlan& P = *plan;
    CUDA_TRY(cudaSetDevice(device));
    const int big = 0x7fffffff;
    int hminor = 0;
    CUDA_TRY(cudaMemcpyAsync(&hminor, dminor, sizeof(int), cudaMemcpyDeviceToHost, stream));
    CUDA_TRY(cudaEventRecord(ev[3], stream));
    CUDA_TRY(cudaStreamSynchronize(stream));
    if (times) {
        float ms;
        cudaEventElapsedTime(&ms, ev[0], ev[1]); times->ms_h2d = ms;
        cudaEventElapsedTime(&ms, ev[1], ev[2]); times->ms_assemble = ms;
        cudaEventElapsedTime(&ms, ev[2], ev[3]); times->ms_factor = ms;
        cudaEventElapsedTime(&ms, ev[0], ev[3]); times->ms_total = ms;
        times->ms_extend = times->ms_potrf = times->ms_trsm = times->ms_dense_update = 0;
        for (auto& sp : spans) {
            cudaEventElapsedTime(&ms, pev[sp.e0], pev[sp.e1]);
            if (sp.cls == 0) times->ms_extend += ms;
            else if (sp.cls == 1) times->ms_potrf += ms;       // small fronts (all three phases fused)
            else if (sp.cls == 2) times->ms_trsm += ms;        // panel kernels (diag block + triangular solves)
            else times->ms_dense_update += ms;
        }
    }
    if (hminor != big) {
        if (minor) *minor = hminor;
        return ST_NOT_POSDEF;
    }
    if (minor) *minor = P.n;
    numeric = true;
    return ST_OK;
}

int CholDevice::factorize(const double* val, bool on_device, i64* minor, CholTimes* times) {
    int rc = factor_begin(val, on_device);
    if (rc) return rc;
    for (int l = 0; l < plan->nlevels; l++)
        if ((rc = factor_level(l))) return rc;
    return factor_end(minor, times);
}

// the whole factorization enqueued on the streams, nothing read back: for callers that record it into a CUDA graph (dense KKT
// solver); `minor` stays on the device (chol_device_minor_ptr: 0x7fffffff = positive definite), the caller marks the factor
// numeric after it has looked at it
int CholDevice::factor_enqueue(const double* val_dev) {
    int rc = factor_begin(val_dev, true);
    if (rc) return rc;
    for (int l = 0; l < plan->nlevels; l++)
        if ((rc = factor_level(l))) return rc;
    return ST_OK;
}

int CholDevice::set_owned(const unsigned char* owned_host) {
    CUDA_TRY(cudaSetDevice(device));
    const size_t ns = plan->fronts.size();
    std::vector<unsigned char> all(std::max<size_t>(ns, 1), 1);
    CUDA_TRY(cudaMemcpy(downed, owned_host ? owned_host : all.data(), ns, cudaMemcpyHostToDevice));
    return ST_OK;
}

int CholDevice::ensure_solve_ws(i64 cols) {
    if (cols <= solve_cols) return ST_OK;
    drop_graphs();                 // the captured sweeps hold the old workspace pointers
    pool_free(dT); pool_free(dX); dT = dX = nullptr; solve_cols = 0;
    const CholPlan& P = *plan;
    CUDA_TRY(pool_malloc((void**)&dT, std::max<i64>((i64)P.rows.size() * cols, 1) * sizeof(double)));
    CUDA_TRY(pool_malloc((void**)&dX, std::max<i64>((i64)P.n * cols, 1) * sizeof(double)));
    pool_free(dpart); dpart = nullptr;
    CUDA_TRY(pool_malloc((void**)&dpart, (size_t)max_solve_ctas * NB * cols * sizeof(double)));
    solve_cols = cols;
    return ST_OK;
}

int CholDevice::solve(int sys, double* B, i64 nrhs, i64 ldB, bool on_device, CholTimes* times, bool async, const unsigned char* active_fronts) {
    const CholPlan& P = *plan;
    const int n = P.n;
    if (n == 0 || nrhs == 0) return ST_OK;
    CUDA_TRY(cudaSetDevice(device));
    // chunk of right-hand sides processed per pass (bounded workspace)
    const i64 maxcols = std::max<i64>(1, std::min<i64>(nrhs, (i64)(256ll << 20) / std::max<i64>(1, (i64)P.rows.size() * 8)));
    int rc = ensure_solve_ws(std::min<i64>(maxcols, 64));
    if (rc) return rc;
    const i64 chunk = solve_cols;
    // called inside a caller's stream capture (the dense KKT solver records its whole solve into one graph): no nested capture,
    // and the inverses of the diagonal blocks are recomputed in the graph (it is replayed after later factorizations)
    cudaStreamCaptureStatus cap_st = cudaStreamCaptureStatusNone;
    cudaStreamIsCapturing(stream, &cap_st);
    const bool capturing = cap_st == cudaStreamCaptureStatusActive;
    if (!capturing) CUDA_TRY(cudaEventRecord(ev[4], stream));
    if ((capturing || !minv_valid) && ninvblk > 0) {
        k_diag_inverse<<<ninvblk, 128, 0, stream>>>(dinv_front, dinv_kb, dF, dL, dMinv);
        minv_valid = true;
    }
    double* dB = B;
    if (!on_device) {
        if ((i64)n * nrhs > bstage_cap) {
            pool_free(dBstage); dBstage = nullptr; bstage_cap = 0;
            CUDA_TRY(pool_malloc((void**)&dBstage, (size_t)n * nrhs * sizeof(double)));
            bstage_cap = (i64)n * nrhs;
        }
        CUDA_TRY(cudaMemcpy2DAsync(dBstage, (size_t)n * 8, B, (size_t)ldB * 8, (size_t)n * 8, nrhs, cudaMemcpyHostToDevice, stream));
        dB = dBstage;
    }
    const i64 ldd = on_device ? ldB : n;
    int pre = 0, post = 0;        // LDL' semantics: diagonal scaling before / after the LL' sweeps
    if (ldl) {
        if (sys == 2) post = 2;             // L D x = b   : x = D^-1/2 (L_ll^-1 b)
        else if (sys == 3) pre = 2;         // D L' x = b  : x = L_ll^-T (D^-1/2 b)
        else if (sys == 4) post = 1;        // L x = b     : x = D^1/2 (L_ll^-1 b)
        else if (sys == 5) pre = 1;         // L' x = b    : x = L_ll^-T (D^1/2 b)
        else if (sys == 6) pre = 3;         // D x = b
        if ((pre || post) && !diagL_valid) {
            if (!ddiagL) CUDA_TRY(pool_malloc((void**)&ddiagL, (size_t)n * sizeof(double)));
            k_diag<<<std::min<int>((int)P.fronts.size(), 148 * 8), 128, 0, stream>>>(dF, (int)P.fronts.size(), dL, ddiagL);
            diagL_valid = true;
        }
    }
    // sys 0..8: reference numbering (src/C/cholmod.c:437-439).  Internal extras for the device-side KKT solver
    // (kkt_gpu.cu): 9 = L x = P b (sys 7 then 4 in one pass), 10 = x = P' L^-T b (sys 5 then 8); with `async` the
    // call returns without synchronising the stream.
    const bool perm_in = (sys == 0 || sys == 9), perm_out = (sys == 0 || sys == 10),
               do_fwd = (sys == 0 || sys == 1 || sys == 2 || sys == 4 || sys == 9),
               do_bwd = (sys == 0 || sys == 1 || sys == 3 || sys == 5 || sys == 10);
    const int gx = std::min((n + 255) / 256, 148 * 8);
    const i64 tstride = (i64)P.rows.size();
    // Reach-restricted forward sweep (sparse right-hand sides): only the fronts on the elimination-tree paths
    // from the nonzero rows to the root can carry nonzeros of L^-1 b.  The small fronts (the bulk of the tree) run from
    // per-level lists filtered on the host; the work vectors are zeroed first so that a skipped child contributes zeros to
    // its parent; the few large fronts near the root are all on some path and run as usual.
    std::vector<int> reach_off, reach_cnt;
    if (active_fronts && do_fwd) {
        std::vector<int> lists;
        reach_off.assign(P.nlevels, 0); reach_cnt.assign(P.nlevels, 0);
        for (int l = 0; l < P.nlevels; l++) {
            const LevelSched& LS = levels[l];
            reach_off[l] = (int)lists.size();
            for (int q = 0; q < LS.small_all_cnt; q++) {
                const int f = hsched[LS.small_all_off + q];
                if (active_fronts[f]) lists.push_back(f);
            }
            reach_cnt[l] = (int)lists.size() - reach_off[l];
        }
        if ((i64)lists.size() > reach_cap) {
            pool_free(dreach); dreach = nullptr; reach_cap = 0;
            CUDA_TRY(pool_malloc((void**)&dreach, std::max<size_t>(lists.size(), 1) * sizeof(int)));
            reach_cap = (i64)lists.size();
        }
        if (!lists.empty()) CUDA_TRY(cudaMemcpyAsync(dreach, lists.data(), lists.size() * sizeof(int), cudaMemcpyHostToDevice, stream));
        CUDA_TRY(cudaStreamSynchronize(stream));       // `lists` is a local: the copy must have left it
    }
    for (i64 j0 = 0; j0 < nrhs; j0 += chunk) {
        const int nc = (int)std::min<i64>(chunk, nrhs - j0);
        double* b = dB + j0 * ldd;
        if (sys == 7) {          // x = P b
            k_perm_gather<<<dim3(gx, nc), 256, 0, stream>>>(b, ldd, dperm, n, dX, n);
            k_copy_cols<<<dim3(gx, nc), 256, 0, stream>>>(dX, n, b, ldd, n);
            continue;
        }
        if (sys == 8) {          // x = P' b
            k_perm_scatter<<<dim3(gx, nc), 256, 0, stream>>>(dX, n, dperm, n, b, ldd);
            k_copy_cols<<<dim3(gx, nc), 256, 0, stream>>>(dX, n, b, ldd, n);
            continue;
        }
        if (sys == 6) {          // D x = b: D = I for LL', diag(L)^2 with LDL' semantics
            if (pre) k_scale_by_diag<<<dim3(gx, nc), 256, 0, stream>>>(b, ldd, ddiagL, dsgn, n, pre);
            continue;
        }
        if (perm_in) k_perm_gather<<<dim3(gx, nc), 256, 0, stream>>>(b, ldd, dperm, n, dX, n);
        else k_copy_cols<<<dim3(gx, nc), 256, 0, stream>>>(b, ldd, dX, n, n);
        if (pre) k_scale_by_diag<<<dim3(gx, nc), 256, 0, stream>>>(dX, n, ddiagL, dsgn, n, pre);
        const long long pstride = (long long)max_solve_ctas * NB;
        // the level sweeps are a fixed launch sequence (hundreds of short dependent kernels): captured once per
        // (columns, directions) into a CUDA graph and replayed
        const bool reach = active_fronts && do_fwd;
        if (reach) CUDA_TRY(cudaMemsetAsync(dT, 0, (size_t)tstride * nc * sizeof(double), stream));
        const bool pfwd = nc == 1 && (persist_mode & 1) && !plevels.empty(), pbwd = nc == 1 && (persist_mode & 2) && !plevels.empty();
        // The CTAs of a persistent sweep wait for each other: launched COOPERATIVELY, so that the grid only starts when all of
        // it is resident (a second persistent sweep on another stream of the process cannot interleave and starve both).
        auto launch_persist = [&](int ctas, auto kernel, auto... args) {
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3((unsigned)ctas); cfg.blockDim = dim3(256); cfg.dynamicSmemBytes = SMEM_PERSIST; cfg.stream = stream;
            cudaLaunchAttribute at[1];
            at[0].id = cudaLaunchAttributeCooperative; at[0].val.cooperative = persist_coop ? 1 : 0;
            cfg.attrs = at; cfg.numAttrs = 1;
            cudaLaunchKernelEx(&cfg, kernel, args...);
        };
        auto sweeps = [&]() {
            if ((pfwd && do_fwd) || (pbwd && do_bwd)) cudaMemsetAsync(dsync, 0, (size_t)2 * nsync * sizeof(int), stream);
            if (do_fwd)
                for (int l = 0; l < P.nlevels; l++) {
                    const LevelSched& LS = levels[l];
                    if (reach) {
                        if (reach_cnt[l])
                            k_fwd<256><<<dim3(reach_cnt[l], nc), 256, 0, stream>>>(dreach + reach_off[l], dF, dchild, drel, dL, dT, tstride, dX, n);
                    } else if (LS.small_all_cnt) {
                        const int tiny = LS.small_cnt[0];          // nr <= 32: first in the small-front list
                        if (tiny) k_fwd<32><<<dim3(tiny, nc), 32, 0, stream>>>(dsched + LS.small_all_off, dF, dchild, drel, dL, dT, tstride, dX, n);
                        if (LS.small_all_cnt > tiny)
                            k_fwd<256><<<dim3(LS.small_all_cnt - tiny, nc), 256, 0, stream>>>(dsched + LS.small_all_off + tiny, dF, dchild, drel, dL, dT, tstride, dX, n);
                    }
                    if (LS.sbig.ng) {
                        k_fwd_gather<<<dim3(LS.gfwd.ctas, nc), 256, 0, stream>>>(dsched + LS.gfwd.goff, dsched + LS.gfwd.goff + LS.gfwd.ng, LS.gfwd.ng, dF, dchild, drel, dT, tstride, dX, n, dgq + LS.gq_off, dgq_child);
                        if (pfwd && LS.pgi >= 0)
                            launch_persist(LS.pctas, k_fwd_persist, plevels[LS.pgi], (const double*)dL, (const double*)dMinv, dT, dX, dsync, herr,
                                           (persist_dbg && plevels[LS.pgi].nf == 1 && l == P.nlevels - 1) ? persist_dbg : (long long*)nullptr);
                        else
                        for (size_t kb = 0; kb < LS.sfwd.size(); kb++) {
                            const Launch& la = LS.sfwd[kb];
                            k_fwd_diag<<<dim3(la.ng, nc), 256, SMEM_SDIAG, stream>>>(sgroups[std::max(la.sgi, 0)], dsched + la.goff, (int)kb, dF, dL, dMinv, dT, tstride, dX, n);
                            if (la.ctas)
                                (nc == 1 ? k_fwd_upd<1> : k_fwd_upd<4>)<<<la.ctas, 256, SMEM_FUPD, stream>>>(sgroups[std::max(la.sgi, 0)], dsched + la.goff, dsched + la.goff + la.ng, la.ng, (int)kb, dF, dL, dT, tstride, nc, (const unsigned char*)nullptr);
                        }
                    }
                }
            if (ldl && do_fwd && do_bwd)       // Lt S Lt' x = b: the signs sit between the two sweeps
                k_scale_by_diag<<<dim3(gx, nc), 256, 0, stream>>>(dX, n, nullptr, dsgn, n, 4);
            if (do_bwd)
                for (int l = P.nlevels - 1; l >= 0; l--) {
                    const LevelSched& LS = levels[l];
                    if (LS.sbig.ng) {
                        k_bwd_gather<<<dim3(LS.sbig.ng, nc), 256, 0, stream>>>(dsched + LS.sbig.goff, dF, drows, dT, tstride, dX, n);
                        if (pbwd && LS.pgi >= 0) {
                            if (LS.prect) k_bwd_rect<<<LS.prect, 256, SMEM_BUPD, stream>>>(plevels[LS.pgi], dL, dT, drect);
                            launch_persist(LS.pctas, k_bwd_persist, plevels[LS.pgi], (const double*)dL, (const double*)dMinv, dT, dX, dsync + nsync,
                                           (const double*)(LS.prect ? drect : nullptr), herr,
                                           (persist_dbg && plevels[LS.pgi].nf == 1 && l == P.nlevels - 1) ? persist_dbg : (long long*)nullptr);
                        } else
                        for (int kb = (int)LS.sbwd.size() - 1; kb >= 0; kb--) {
                            const Launch& la = LS.sbwd[kb];
                            if (la.ctas)
                                (nc == 1 ? k_bwd_upd<1> : k_bwd_upd<BWD_CG>)<<<dim3(la.ctas, nc == 1 ? 1 : (nc + BWD_CG - 1) / BWD_CG), 256, SMEM_BUPD, stream>>>(
                                    sgroups[std::max(la.sgi, 0)], dsched + la.goff, dsched + la.goff + la.ng, la.ng, kb, dF, dL, dT, tstride, dpart, pstride, nc, (const unsigned char*)nullptr);
                            k_bwd_diag<<<dim3(la.ng, nc), 256, SMEM_BDIAG, stream>>>(sgroups[std::max(la.sgi, 0)], dsched + la.goff, dsched + la.goff + la.ng, kb, dF, dL, dMinv, dT, tstride, dX, n, dpart, pstride);
                        }
                    }
                    if (LS.small_all_cnt) {
                        const int tiny = LS.small_cnt[0];
                        if (LS.small_all_cnt > tiny)
                            k_bwd<256><<<dim3(LS.small_all_cnt - tiny, nc), 256, 0, stream>>>(dsched + LS.small_all_off + tiny, dF, drows, dL, dT, tstride, dX, n);
                        if (tiny) k_bwd<32><<<dim3(tiny, nc), 32, 0, stream>>>(dsched + LS.small_all_off, dF, drows, dL, dT, tstride, dX, n);
                    }
                }
        };
        const int gkey = nc * 32 + (pfwd ? 16 : 0) + (pbwd ? 8 : 0) + (ldl ? 4 : 0) + (do_fwd ? 2 : 0) + (do_bwd ? 1 : 0);
        if (!use_graphs || reach || capturing) sweeps();      // the filtered lists change from call to call: not captured
        else {
            auto it = solve_graphs.find(gkey);
            if (it == solve_graphs.end()) {
                cudaGraph_t gr = nullptr;
                cudaGraphExec_t ge = nullptr;
                CUDA_TRY(cudaStreamBeginCapture(stream, cudaStreamCaptureModeThreadLocal));
                sweeps();
                CUDA_TRY(cudaStreamEndCapture(stream, &gr));
                cudaError_t ie = cudaGraphInstantiate(&ge, gr, 0);
                cudaGraphDestroy(gr);
                CUDA_TRY(ie);
                it = solve_graphs.emplace(gkey, ge).first;
            }
            CUDA_TRY(cudaGraphLaunch(it->second, stream));
        }
        if (post) k_scale_by_diag<<<dim3(gx, nc), 256, 0, stream>>>(dX, n, ddiagL, dsgn, n, post);
        if (perm_out) k_perm_scatter<<<dim3(gx, nc), 256, 0, stream>>>(b, ldd, dperm, n, dX, n);
        else k_copy_cols<<<dim3(gx, nc), 256, 0, stream>>>(dX, n, b, ldd, n);
    }
    cudaError_t le = cudaGetLastError();
    if (!on_device && le == cudaSuccess)
        le = cudaMemcpy2DAsync(B, (size_t)ldB * 8, dBstage, (size_t)n * 8, (size_t)n * 8, nrhs, cudaMemcpyDeviceToHost, stream);
    if (!capturing) cudaEventRecord(ev[5], stream);
    if (async && on_device) { CUDA_TRY(le); return ST_OK; }
    cudaError_t se = cudaStreamSynchronize(stream);
    CUDA_TRY(le);
    CUDA_TRY(se);
    if (persist_dbg) {
        double seen_to_item = 0, items = 0, dg = 0, pub = 0, prop = 0, idle = 0; int cnt = 0;
        for (int c = 2; c < 116; c++) {
            const long long* d = persist_dbg + c * 8; const long long* dp = persist_dbg + (c - 1) * 8;
            if (!d[0] || !dp[6]) continue;
            prop += (double)(d[0] - dp[6]); seen_to_item += (double)(d[2] - d[1]); items += (double)(d[3] - d[2]); dg += (double)(d[4] - d[3]); pub += (double)(d[5] - d[4]);
            idle += (double)(d[1] - d[7]); cnt++;
        }
        if (cnt) fprintf(stderr, "[persist dbg] steps %d: publish->seen %.0f ns | cycles: seen->item1 %.0f, items %.0f, diag %.0f, publish %.0f, waited %.0f\n", cnt, prop / cnt, seen_to_item / cnt, items / cnt, dg / cnt, pub / cnt, idle / cnt);
        {
            double a[4] = {0}; int n2 = 0;
            for (int c = 2; c < 110; c++) {
                const long long* d = persist_dbg + 2048 + c * 16; const long long* dn = persist_dbg + 2048 + (c + 1) * 16;
                if (!d[8] || !dn[9]) continue;
                a[0] += (double)(d[1] - d[0]); a[1] += (double)(d[2] - d[1]); a[2] += (double)(d[3] - d[2]); a[3] += (double)(d[8] - dn[9]); n2++;
            }
            if (n2) fprintf(stderr, "[persist dbg bwd] steps %d: cycles waited %.0f, seen->solve %.0f, solve+publish %.0f | publish->seen %.0f ns\n",
                            n2, a[0] / n2, a[1] / n2, a[2] / n2, a[3] / n2);
        }
        memset(persist_dbg, 0, 8192 * sizeof(long long));
    }
    if (herr && *herr) {
        *herr = 0;
        set_last_error("persistent solve sweep: a CTA gave up waiting for another one (schedule error); set B200S_SOLVE_PERSIST=0");
        return ST_CUDA;
    }
    if (times) { float ms; cudaEventElapsedTime(&ms, ev[4], ev[5]); times->ms_solve = ms; }
    return ST_OK;
}

// ---- level-stepped solve for factors whose fronts live on several GPUs -------------------------------------------------
// begin: X = P b (every rank has b); level (forward, leaves to root): the owned fronts of one level gather their pivots
// from X and their children's update vectors from T -- the caller has copied the update vectors of children owned by another
// GPU into T before (identical layout on every rank) --, solve, and leave y in X and their update vector in T; level
// (backward, root to leaves): the owned fronts gather x at their row lists from X -- the caller has copied the solution
// entries of ancestors owned elsewhere into X --, and write x(columns) to X; end: x = P' X.  LL' factors only.
int CholDevice::solve_dist_begin(const double* b_dev) {
    const int n = plan->n;
    CUDA_TRY(cudaSetDevice(device));
    if (n == 0) return ST_OK;
    if (ldl) return ST_INVALID;
    int rc = ensure_solve_ws(1);
    if (rc) return rc;
    if (ninvblk > 0 && !minv_valid) {
        k_diag_inverse<<<ninvblk, 128, 0, stream>>>(dinv_front, dinv_kb, dF, dL, dMinv);
        minv_valid = true;
    }
    const int gx = std::min((n + 255) / 256, 148 * 8);
    k_perm_gather<<<dim3(gx, 1), 256, 0, stream>>>(b_dev, n, dperm, n, dX, n);
    CUDA_TRY(cudaGetLastError());
    return ST_OK;
}

int CholDevice::solve_dist_level(int backward, int l) {
    const CholPlan& P = *plan;
    if (l < 0 || l >= P.nlevels) return ST_INVALID;
    CUDA_TRY(cudaSetDevice(device));
    const int n = P.n, nc = 1;
    const i64 tstride = (i64)P.rows.size();
    const long long pstride = (long long)max_solve_ctas * NB;
    const LevelSched& LS = levels[l];
    if (!backward) {
        if (LS.small_all_cnt)
            k_fwd<256><<<dim3(LS.small_all_cnt, nc), 256, 0, stream>>>(dsched + LS.small_all_off, dF, dchild, drel, dL, dT, tstride, dX, n, downed);
        if (LS.sbig.ng) {
            k_fwd_gather<<<dim3(LS.gfwd.ctas, nc), 256, 0, stream>>>(dsched + LS.gfwd.goff, dsched + LS.gfwd.goff + LS.gfwd.ng, LS.gfwd.ng, dF, dchild, drel, dT, tstride, dX, n, dgq + LS.gq_off, dgq_child, downed);
            for (size_t kb = 0; kb < LS.sfwd.size(); kb++) {
                const Launch& la = LS.sfwd[kb];
                k_fwd_diag<<<dim3(la.ng, nc), 256, SMEM_SDIAG, stream>>>(sgroups[std::max(la.sgi, 0)], dsched + la.goff, (int)kb, dF, dL, dMinv, dT, tstride, dX, n, downed);
                if (la.ctas)
                    k_fwd_upd<1><<<la.ctas, 256, SMEM_FUPD, stream>>>(sgroups[std::max(la.sgi, 0)], dsched + la.goff, dsched + la.goff + la.ng, la.ng, (int)kb, dF, dL, dT, tstride, nc, downed);
            }
        }
    } else {
        if (LS.sbig.ng) {
            k_bwd_gather<<<dim3(LS.sbig.ng, nc), 256, 0, stream>>>(dsched + LS.sbig.goff, dF, drows, dT, tstride, dX, n, downed);
            for (int kb = (int)LS.sbwd.size() - 1; kb >= 0; kb--) {
                const Launch& la = LS.sbwd[kb];
                if (la.ctas)
                    k_bwd_upd<1><<<dim3(la.ctas, 1), 256, SMEM_BUPD, stream>>>(
                        sgroups[std::max(la.sgi, 0)], dsched + la.goff, dsched + la.goff + la.ng, la.ng, kb, dF, dL, dT, tstride, dpart, pstride, nc, downed);
                k_bwd_diag<<<dim3(la.ng, nc), 256, SMEM_BDIAG, stream>>>(sgroups[std::max(la.sgi, 0)], dsched + la.goff, dsched + la.goff + la.ng, kb, dF, dL, dMinv, dT, tstride, dX, n, dpart, pstride, downed);
            }
        }
        if (LS.small_all_cnt)
            k_bwd<256><<<dim3(LS.small_all_cnt, nc), 256, 0, stream>>>(dsched + LS.small_all_off, dF, drows, dL, dT, tstride, dX, n, downed);
    }
    CUDA_TRY(cudaGetLastError());
    return ST_OK;
}

int CholDevice::solve_dist_end(double* x_dev) {
    const int n = plan->n;
    CUDA_TRY(cudaSetDevice(device));
    if (n == 0) return ST_OK;
    const int gx = std::min((n + 255) / 256, 148 * 8);
    k_perm_scatter<<<dim3(gx, 1), 256, 0, stream>>>(x_dev, n, dperm, n, dX, n);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaStreamSynchronize(stream));
    return ST_OK;
}

// ---- sparse right-hand sides (cholmod.spsolve, reference src/C/cholmod.c:524-587) ------------------------------------
// dense block B (n x nc, zeroed) <- the nonzeros of nc sparse columns: cp[c]..cp[c+1] index (ri, vx)
__global__ void k_sp_scatter(const i64* __restrict__ cp, const i64* __restrict__ ri, const double* __restrict__ vx, int nc,
                             long long n, double* __restrict__ B) {
    const int c = blockIdx.y;
    for (long long p = cp[c] + blockIdx.x * (long long)blockDim.x + threadIdx.x; p < cp[c + 1]; p += (long long)gridDim.x * blockDim.x)
        B[(long long)c * n + ri[p]] = vx[p];
}
// numerically nonzero entries per column
__global__ void __launch_bounds__(256) k_sp_count(const double* __restrict__ B, long long n, int* __restrict__ cnt) {
    __shared__ int red[256];
    const double* b = B + (long long)blockIdx.x * n;
    int c = 0;
    for (long long i = threadIdx.x; i < n; i += 256) c += b[i] != 0.0;
    red[threadIdx.x] = c;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) { if ((int)threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o]; __syncthreads(); }
    if (threadIdx.x == 0) cnt[blockIdx.x] = red[0];
}
// ordered compaction of column blockIdx.x into (oi, ox) starting at off[blockIdx.x]: tiles of 1024 rows, ballot + warp-sum scan
__global__ void __launch_bounds__(1024) k_sp_compact(const double* __restrict__ B, long long n, const i64* __restrict__ off,
                                                     i64* __restrict__ oi, double* __restrict__ ox) {
    __shared__ int wsum[32];
    __shared__ long long base;
    const double* b = B + (long long)blockIdx.x * n;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) base = off[blockIdx.x];
    __syncthreads();
    for (long long i0 = 0; i0 < n; i0 += 1024) {
        const long long i = i0 + threadIdx.x;
        const double v = i < n ? b[i] : 0.0;
        const unsigned m = __ballot_sync(0xffffffffu, v != 0.0);
        if (lane == 0) wsum[warp] = __popc(m);
        __syncthreads();
        int before = 0, total = 0;
        for (int w = 0; w < 32; w++) { const int c = wsum[w]; if (w < warp) before += c; total += c; }
        if (v != 0.0) {
            const long long d = base + before + __popc(m & ((1u << lane) - 1u));
            oi[d] = i; ox[d] = v;
        }
        __syncthreads();
        if (threadIdx.x == 0) base += total;
        __syncthreads();
    }
}

int CholDevice::spsolve(int sys, i64 ncols, const i64* Bp, const i64* Bi, const double* Bx, std::vector<i64>& Xp, std::vector<i64>& Xi,
                        std::vector<double>& Xx, CholTimes* times) {
    const CholPlan& P = *plan;
    const i64 n = P.n;
    Xp.assign((size_t)ncols + 1, 0); Xi.clear(); Xx.clear();
    if (n == 0 || ncols == 0) return ST_OK;
    CUDA_TRY(cudaSetDevice(device));
    const i64 maxcols = std::max<i64>(1, std::min<i64>(ncols, (i64)(256ll << 20) / std::max<i64>(1, (i64)P.rows.size() * 8)));
    int rc = ensure_solve_ws(std::min<i64>(maxcols, 64));
    if (rc) return rc;
    const i64 chunk = solve_cols;
    // the systems whose forward sweep starts from b itself (4: L x = b; 2: LD x = b) or from P b (0, 1 have the backward sweep too)
    const bool fwd = sys == 0 || sys == 1 || sys == 2 || sys == 4, perm_in = sys == 0;
    const size_t nf = P.fronts.size();
    std::vector<unsigned char> active(nf);
    std::vector<i64> cpl((size_t)chunk + 1), offs((size_t)chunk);
    std::vector<int> cnt((size_t)chunk);
    float ms_acc = 0.f;
    for (i64 j0 = 0; j0 < ncols; j0 += chunk) {
        const int nc = (int)std::min<i64>(chunk, ncols - j0);
        const i64 e0 = Bp[j0], ne = Bp[j0 + nc] - e0;
        // reach of the chunk: fronts on the tree paths from its nonzero rows (permuted for the systems that start with P b)
        if (fwd) {
            std::fill(active.begin(), active.end(), 0);
            for (i64 p = e0; p < e0 + ne; p++) {
                const i64 r = perm_in ? P.iperm[Bi[p]] : Bi[p];
                for (i32 f = P.sn_of_col[r]; f >= 0 && !active[f]; f = P.fronts[f].parent) active[f] = 1;
            }
        }
        if (ne > sp_cap || !dsp_i) {
            pool_free(dsp_i); pool_free(dsp_x); dsp_i = nullptr; dsp_x = nullptr; sp_cap = 0;
            CUDA_TRY(pool_malloc((void**)&dsp_i, (size_t)std::max<i64>(ne, 1) * sizeof(i64)));
            CUDA_TRY(pool_malloc((void**)&dsp_x, (size_t)std::max<i64>(ne, 1) * sizeof(double)));
            sp_cap = std::max<i64>(ne, 1);
        }
        if (2 * chunk + 2 > spmeta_cap) {
            pool_free(dsp_meta); pool_free(dsp_cnt); dsp_meta = nullptr; dsp_cnt = nullptr; spmeta_cap = 0;
            CUDA_TRY(pool_malloc((void**)&dsp_meta, (size_t)(2 * chunk + 2) * sizeof(i64)));
            CUDA_TRY(pool_malloc((void**)&dsp_cnt, (size_t)chunk * sizeof(int)));
            spmeta_cap = 2 * chunk + 2;
        }
        if ((i64)n * nc > bstage_cap) {
            pool_free(dBstage); dBstage = nullptr; bstage_cap = 0;
            CUDA_TRY(pool_malloc((void**)&dBstage, (size_t)n * chunk * sizeof(double)));
            bstage_cap = (i64)n * chunk;
        }
        i64* d_cp = dsp_meta;                  // chunk column pointers (relative to the chunk's first entry)
        i64* d_off = dsp_meta + chunk + 1;     // output offsets of the chunk's columns
        for (auto& e : ev_sp) if (!e) CUDA_TRY(cudaEventCreate(&e));
        for (int c = 0; c <= nc; c++) cpl[c] = Bp[j0 + c] - e0;
        cudaEvent_t e_a = ev_sp[0], e_b = ev_sp[1];
        CUDA_TRY(cudaEventRecord(e_a, stream));
        CUDA_TRY(cudaMemsetAsync(dBstage, 0, (size_t)n * nc * sizeof(double), stream));
        if (ne) {
            CUDA_TRY(cudaMemcpyAsync(dsp_i, Bi + e0, (size_t)ne * sizeof(i64), cudaMemcpyHostToDevice, stream));
            CUDA_TRY(cudaMemcpyAsync(dsp_x, Bx + e0, (size_t)ne * sizeof(double), cudaMemcpyHostToDevice, stream));
        }
        CUDA_TRY(cudaMemcpyAsync(d_cp, cpl.data(), (size_t)(nc + 1) * sizeof(i64), cudaMemcpyHostToDevice, stream));
        if (ne) k_sp_scatter<<<dim3(8, nc), 256, 0, stream>>>(d_cp, dsp_i, dsp_x, nc, n, dBstage);
        CUDA_TRY(cudaGetLastError());
        CUDA_TRY(cudaStreamSynchronize(stream));       // cpl is reused by the next chunk
        rc = solve(sys, dBstage, nc, n, true, nullptr, true, fwd ? active.data() : nullptr);
        if (rc) return rc;
        k_sp_count<<<nc, 256, 0, stream>>>(dBstage, n, dsp_cnt);
        CUDA_TRY(cudaMemcpyAsync(cnt.data(), dsp_cnt, (size_t)nc * sizeof(int), cudaMemcpyDeviceToHost, stream));
        CUDA_TRY(cudaStreamSynchronize(stream));
        i64 tot = 0;
        for (int c = 0; c < nc; c++) { offs[c] = tot; tot += cnt[c]; Xp[j0 + c + 1] = Xp[j0 + c] + cnt[c]; }
        if (tot > spout_cap) {
            pool_free(dsp_oi); pool_free(dsp_ox); dsp_oi = nullptr; dsp_ox = nullptr; spout_cap = 0;
            CUDA_TRY(pool_malloc((void**)&dsp_oi, (size_t)tot * sizeof(i64)));
            CUDA_TRY(pool_malloc((void**)&dsp_ox, (size_t)tot * sizeof(double)));
            spout_cap = tot;
        }
        if (tot) {
            CUDA_TRY(cudaMemcpyAsync(d_off, offs.data(), (size_t)nc * sizeof(i64), cudaMemcpyHostToDevice, stream));
            k_sp_compact<<<nc, 1024, 0, stream>>>(dBstage, n, d_off, dsp_oi, dsp_ox);
            const size_t old = Xi.size();
            Xi.resize(old + (size_t)tot); Xx.resize(old + (size_t)tot);
            CUDA_TRY(cudaMemcpyAsync(Xi.data() + old, dsp_oi, (size_t)tot * sizeof(i64), cudaMemcpyDeviceToHost, stream));
            CUDA_TRY(cudaMemcpyAsync(Xx.data() + old, dsp_ox, (size_t)tot * sizeof(double), cudaMemcpyDeviceToHost, stream));
        }
        CUDA_TRY(cudaEventRecord(e_b, stream));
        CUDA_TRY(cudaStreamSynchronize(stream));
        float ms; cudaEventElapsedTime(&ms, e_a, e_b); ms_acc += ms;
    }
    if (times) times->ms_solve = ms_acc;
    return ST_OK;
}

CholDevice* chol_device_create(const CholPlan& plan, const CholOpts& opts, int device, int* status) {
    if (device_count() <= 0) { *status = ST_NO_DEVICE; set_last_error("no CUDA device available"); return nullptr; }
    CholDevice* d = new CholDevice();
    d->plan = &plan;
    d->opts = opts;
    d->device = device;
    *status = d->init();
    if (*status != ST_OK) { delete d; return nullptr; }
    return d;
}
void chol_device_destroy(CholDevice* d) { delete d; }
int chol_device_factorize(CholDevice* d, const double* val, bool val_on_device, i64* minor, CholTimes* times) {
    return d->factorize(val, val_on_device, minor, times);
}
int chol_device_solve(CholDevice* d, int sys, double* B, i64 nrhs, i64 ldB, bool on_device, CholTimes* times) {
    return d->solve(sys, B, nrhs, ldB, on_device, times, false);
}
int chol_device_spsolve(CholDevice* d, int sys, i64 ncols, const i64* Bp, const i64* Bi, const double* Bx, std::vector<i64>& Xp,
                        std::vector<i64>& Xi, std::vector<double>& Xx, CholTimes* times) {
    return d->spsolve(sys, ncols, Bp, Bi, Bx, Xp, Xi, Xx, times);
}
int chol_device_solve_async(CholDevice* d, int sys, double* B_dev, i64 nrhs, i64 ldB) {
    return d->solve(sys, B_dev, nrhs, ldB, true, nullptr, true);
}
void* chol_device_stream(CholDevice* d) { return (void*)d->stream; }
int chol_device_diag(CholDevice* d, double* diag_host) {
    const CholPlan& P = *d->plan;
    if (P.n == 0) return ST_OK;
    CUDA_TRY(cudaSetDevice(d->device));
    int rc = d->ensure_solve_ws(1);
    if (rc) return rc;
    k_diag<<<std::min<int>((int)P.fronts.size(), 148 * 8), 128, 0, d->stream>>>(d->dF, (int)P.fronts.size(), d->dL, d->dX);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaMemcpyAsync(diag_host, d->dX, P.n * sizeof(double), cudaMemcpyDeviceToHost, d->stream));
    CUDA_TRY(cudaStreamSynchronize(d->stream));
    return ST_OK;
}
int chol_device_download_L(CholDevice* d, double* L_host) {
    CUDA_TRY(cudaSetDevice(d->device));
    CUDA_TRY(cudaStreamSynchronize(d->stream));
    if (d->plan->lsize) CUDA_TRY(cudaMemcpy(L_host, d->dL, d->plan->lsize * sizeof(double), cudaMemcpyDeviceToHost));
    return ST_OK;
}
void chol_device_set_profiling(CholDevice* d, bool on) { d->profiling = on; }
void chol_device_set_ldl(CholDevice* d, bool on) { d->ldl = on; }
int chol_device_download_sign(CholDevice* d, double* sign_host) {
    const CholPlan& P = *d->plan;
    if (P.n == 0) return ST_OK;
    if (!d->ldl || !d->dsgn) { for (i64 i = 0; i < P.n; i++) sign_host[i] = 1.0; return ST_OK; }
    CUDA_TRY(cudaSetDevice(d->device));
    CUDA_TRY(cudaStreamSynchronize(d->stream));
    CUDA_TRY(cudaMemcpy(sign_host, d->dsgn, P.n * sizeof(double), cudaMemcpyDeviceToHost));
    return ST_OK;
}
int chol_device_set_owned(CholDevice* d, const unsigned char* owned) { return d->set_owned(owned); }
int chol_device_factor_begin(CholDevice* d, const double* val, bool on_device) { return d->factor_begin(val, on_device); }
int chol_device_factor_level(CholDevice* d, int level) { return d->factor_level(level); }
int chol_device_factor_level_phase(CholDevice* d, int level, int phase) { return d->factor_level(level, phase); }
int chol_device_set_syrk_split(CholDevice* d, const unsigned char* own, const int* lo, const int* hi, const long long* base, double* scratch) { return d->set_syrk_split(own, lo, hi, base, scratch); }
int chol_device_factor_end(CholDevice* d, i64* minor, CholTimes* times) { return d->factor_end(minor, times); }
int chol_device_sync(CholDevice* d) {
    CUDA_TRY(cudaSetDevice(d->device));
    CUDA_TRY(cudaStreamSynchronize(d->stream));
    return ST_OK;
}
void chol_device_buffers(CholDevice* d, double** L, double** W) { *L = d->dL; *W = d->dW; }
int chol_device_solve_dist_begin(CholDevice* d, const double* b_dev) { return d->solve_dist_begin(b_dev); }
int chol_device_solve_dist_level(CholDevice* d, int backward, int level) { return d->solve_dist_level(backward, level); }
int chol_device_solve_dist_end(CholDevice* d, double* x_dev) { return d->solve_dist_end(x_dev); }
int chol_device_solve_buffers(CholDevice* d, double** T, double** X) {
    int rc = d->ensure_solve_ws(1);
    *T = d->dT; *X = d->dX;
    return rc;
}
int chol_device_factor_enqueue(CholDevice* d, const double* val_dev) { return d->factor_enqueue(val_dev); }
const int* chol_device_minor_ptr(CholDevice* d) { return d->dminor; }
void chol_device_set_solve_sweeps(CholDevice* d, int mode) { d->persist_mode = mode < 0 ? d->persist_default : (mode & d->persist_hw); }
void chol_device_mark_numeric(CholDevice* d, bool numeric) { d->numeric = numeric; d->minv_valid = false; }
i64 chol_device_workspace_bytes(const CholDevice* d) { return d->total_bytes; }

}  // namespace b200s
