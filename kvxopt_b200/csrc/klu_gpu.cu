// Batched KLU numeric refactorization and solves on B200 (sm_100a).
// Replaces klu_l_factor on same-pattern inputs (reference src/C/klu.c:337, klu_refactor semantics) and
// klu_l_solve / klu_l_tsolve (src/C/klu.c:651-657) for a batch of matrices that share one sparsity
// pattern and one pivot sequence.
//
// Layout in HBM: every per-matrix quantity is batch-interleaved, value v of matrix b at [v*Bp + b]
// (Bp = batch rounded up to 32), so that the 32 lanes of a warp -- 32 different matrices executing the
// same static schedule -- read and write 256 contiguous bytes.  No divergence: the pattern, the pivot
// order and therefore the instruction stream are identical for all matrices.
//   Axt [nnzA ][Bp]  transposed copy of the caller's values
//   Rs  [n    ][Bp]  row scale factors (max |row|), pivotal row order
//   LU  [Bp/32][slots][32]  GROUP-MAJOR: the slots of a group of 32 matrices are contiguous (value v of matrix b at
//       (b/32)*slots*32 + v*32 + b%32), so a run of consecutive slots -- the L part of a column -- is ONE contiguous
//       block that a single cp.async.bulk (TMA) moves.  Per column: U above the diagonal, U diagonal, L below; then F
// Kernels: k_klu_transpose (tiled), k_klu_rowscale (row maxima + scaling), k_klu_scatter, k_klu_early (the wide first levels of
// the dependency graph: one warp per (column, group), one launch per level), k_klu_refactor_wave (the remaining columns: TMA
// producer warp + 16 consumer warps per group of 32 matrices), k_klu_dense_pack / k_klu_dense_lu (dense trailing block),
// k_klu_refactor (level-schedule kernel for patterns outside the wave kernel's budget), k_klu_load_one (klu.numeric: the host
// pivot search's values), k_klu_solve_lvl.
#include "gpu.hpp"
#include "devpool.hpp"
#include "klu_host.hpp"
#include <cuda_runtime.h>
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <string>
#include <chrono>

namespace b200s {

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

// in[b*ldv + k] -> out[k*Bp + b]
__global__ void k_klu_transpose(const double* __restrict__ in, long long ldv, long long nnz, int batch, int Bp,
                                double* __restrict__ out) {
    __shared__ double tile[32][33];
    const long long k0 = (long long)blockIdx.x * 32;
    const int b0 = blockIdx.y * 32;
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        const int b = b0 + r;
        const long long k = k0 + threadIdx.x;
        tile[r][threadIdx.x] = (b < batch && k < nnz) ? in[(long long)b * ldv + k] : 0.0;
    }
    __syncthreads();
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        const long long k = k0 + r;
        const int b = b0 + threadIdx.x;
        if (k < nnz) out[k * Bp + b] = tile[threadIdx.x][r];
    }
}

// Rs[i][b] = max_k |A(i,k)| (1 when the row is empty or zero)
// prescale != 0: the row's entries are divided by the scale right away (they were just read: the second sweep hits L1/L2), which
// replaces the separate k_klu_prescale pass over all of A (one 0.96 GB read and the 0.13 GB of scales less per batch of 4096)
__global__ void k_klu_rowscale(const long long* __restrict__ rowptr, const int* __restrict__ rowent, int n, int Bp,
                               double* __restrict__ Axt, double* __restrict__ Rs, int prescale) {
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int ngroups = Bp >> 5;
    const long long total = (long long)n * ngroups;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    for (long long w = warp; w < total; w += nwarps) {
        const int i = (int)(w / ngroups), g = (int)(w - (long long)i * ngroups);
        const int b = g * 32 + lane;
        double m = 0.0;
        const long long p0 = rowptr[i], p1 = rowptr[i + 1];
        for (long long p = p0; p < p1; p++) m = fmax(m, fabs(Axt[(long long)rowent[p] * Bp + b]));
        const double rs = (m > 0.0) ? m : 1.0;
        Rs[(long long)i * Bp + b] = rs;
        if (prescale)
            for (long long p = p0; p < p1; p++) {
                double* a = Axt + (long long)rowent[p] * Bp + b;
                *a = *a / rs;
            }
    }
}

// LU[v][b] = A(src(v))[b] / Rs[row(v)][b], zero for fill-in slots
__global__ void k_klu_scatter(const int* __restrict__ slot_src, const int* __restrict__ slot_row, long long slot0,
                              long long nslots, int Bp, const double* __restrict__ Axt, const double* __restrict__ Rs,
                              double* __restrict__ LU, int prescaled, long long gstride) {
    const long long total = nslots * Bp, first = slot0 * Bp;
    for (long long t = first + blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        const long long v = t / Bp;
        const int b = (int)(t - v * Bp);
        const int src = slot_src[v];
        LU[(long long)(b >> 5) * gstride + v * 32 + (b & 31)] = (src >= 0) ? (prescaled ? Axt[(long long)src * Bp + b] : Axt[(long long)src * Bp + b] / Rs[(long long)slot_row[v] * Bp + b]) : 0.0;
    }
}

// Early columns (KluPlan::early): one warp per (column, group of 32 matrices), lane = matrix.  The column lives in the warp's
// shared-memory scratch ([row][32]); its sources are finished columns of earlier levels (earlier launches), read straight from
// LU with coalesced 256-byte rows.  No inter-warp dependency, thousands of columns per level: the launch fills the GPU and runs
// at memory speed, instead of the one-wave-at-a-time latency of the wave kernel (these columns hold ~85 % of the columns and
// ~5 % of the multiply-adds of a power-flow Jacobian).
struct KluEarlyCol { int cb, len, diag, l0, upd0, nupd; };      // first slot, slots, pivot row, first L row, updates
template <int ROWS, int WARPS>
__global__ void __launch_bounds__(WARPS * 32) k_klu_early(const KluEarlyCol* __restrict__ cols, int c0, int c1,
                                                          const int4* __restrict__ eupd, const unsigned short* __restrict__ edest,
                                                          const int* __restrict__ slot_src, long long gstride, int Bp,
                                                          const double* __restrict__ Axs, double* __restrict__ LU,
                                                          int* __restrict__ status) {
    extern __shared__ double smem_early[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int ci = c0 + blockIdx.x * WARPS + warp;
    if (ci >= c1) return;
    const KluEarlyCol C = cols[ci];
    double* x = smem_early + warp * ROWS * 32 + lane;
    const double* ax = Axs + (long long)blockIdx.y * 32 + lane;
    double* lu = LU + (long long)blockIdx.y * gstride + lane;
    // The source rows of all updates are requested from HBM up front (L2 prefetch: no registers, no waiting); the update loop
    // below is a chain of dependent accesses per update (descriptor -> L rows -> shared-memory read-modify-write) and would
    // otherwise pay the HBM latency once per update with only a few warps per SM to hide it.
    for (int u = 0; u < C.nupd; u++) {
        const int4 U = eupd[C.upd0 + u];
        const double* lcol = lu + (long long)U.y * 32;
        for (int t = 0; t < U.z; t++) asm volatile("prefetch.global.L2 [%0];" ::"l"(lcol + t * 32));
    }
    for (int r = 0; r < C.len; r++) {          // pre-scaled input values (k_klu_rowscale), zero in fill-in slots
        const int src = slot_src[C.cb + r];
        x[r * 32] = src >= 0 ? ax[(long long)src * Bp] : 0.0;
    }
    int4 Un = C.nupd > 0 ? eupd[C.upd0] : make_int4(0, 0, 0, 0);
    for (int u = 0; u < C.nupd; u++) {
        const int4 U = Un;                     // {row of u_jk, first L slot of the source, rows, first destination}
        if (u + 1 < C.nupd) Un = eupd[C.upd0 + u + 1];
        const double ujk = x[U.x * 32];
        const double* lcol = lu + (long long)U.y * 32;
        const unsigned short* d = edest + U.w;
        int t = 0;
        for (; t + 4 <= U.z; t += 4) {
            const double l0 = lcol[t * 32], l1 = lcol[(t + 1) * 32], l2 = lcol[(t + 2) * 32], l3 = lcol[(t + 3) * 32];
            double* p0 = x + d[t] * 32; double* p1 = x + d[t + 1] * 32; double* p2 = x + d[t + 2] * 32; double* p3 = x + d[t + 3] * 32;
            *p0 = fma(-l0, ujk, *p0); *p1 = fma(-l1, ujk, *p1); *p2 = fma(-l2, ujk, *p2); *p3 = fma(-l3, ujk, *p3);
        }
        for (; t < U.z; t++) { double* p0 = x + d[t] * 32; *p0 = fma(-lcol[t * 32], ujk, *p0); }
    }
    const double piv = x[C.diag * 32];
    const bool bad = !(fabs(piv) > 0.0);       // zero or NaN pivot
    const double rpiv = 1.0 / piv;
    for (int sl = C.l0; sl < C.len; sl++) x[sl * 32] *= rpiv;
    for (int r = 0; r < C.len; r++) lu[(long long)(C.cb + r) * 32] = x[r * 32];
    if (bad) status[blockIdx.y * 32 + lane] = ST_SINGULAR;
}

struct KluPlanD {
    int n, nlevels;
    long long gstride;          // doubles between the LU blocks of consecutive groups of 32 matrices (= slots * 32)
    const int *level_ptr, *level_cols, *udiag_slot, *lslot0, *upd_uslot, *upd_lslot, *upd_cnt, *dest;
    const long long *cbeg, *upd_ptr, *upd_dest;
};

// One CTA per group of 32 matrices.  Levels of the column dependency graph are separated by
// __syncthreads(); inside a level each warp takes whole columns.
constexpr int KLU_WARPS = 16;
__global__ void __launch_bounds__(KLU_WARPS * 32) k_klu_refactor(KluPlanD P, long long gstride, double* __restrict__ LU,
                                                                 int* __restrict__ status) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int b = blockIdx.x * 32 + lane;
    double* lu = LU + (long long)blockIdx.x * gstride + lane;
    int bad = 0;
    for (int l = 0; l < P.nlevels; l++) {
        for (int c = P.level_ptr[l] + warp; c < P.level_ptr[l + 1]; c += KLU_WARPS) {
            const int k = P.level_cols[c];
            for (long long u = P.upd_ptr[k]; u < P.upd_ptr[k + 1]; u++) {
                const double ujk = lu[(long long)P.upd_uslot[u] * 32];
                const int cnt = P.upd_cnt[u];
                const double* lcol = lu + (long long)P.upd_lslot[u] * 32;
                const int* d = P.dest + P.upd_dest[u];
                int t = 0;
                for (; t + 4 <= cnt; t += 4) {
                    const double l0 = lcol[(long long)t * 32], l1 = lcol[(long long)(t + 1) * 32];
                    const double l2 = lcol[(long long)(t + 2) * 32], l3 = lcol[(long long)(t + 3) * 32];
                    double* p0 = lu + (long long)d[t] * 32; double* p1 = lu + (long long)d[t + 1] * 32;
                    double* p2 = lu + (long long)d[t + 2] * 32; double* p3 = lu + (long long)d[t + 3] * 32;
                    const double x0 = *p0, x1 = *p1, x2 = *p2, x3 = *p3;
                    *p0 = x0 - l0 * ujk; *p1 = x1 - l1 * ujk; *p2 = x2 - l2 * ujk; *p3 = x3 - l3 * ujk;
                }
                for (; t < cnt; t++) {
                    double* p0 = lu + (long long)d[t] * 32;
                    *p0 -= lcol[(long long)t * 32] * ujk;
                }
            }
            const double piv = lu[(long long)P.udiag_slot[k] * 32];
            if (!(fabs(piv) > 0.0)) bad = 1;            // zero or NaN pivot
            const long long l0 = P.lslot0[k], l1 = P.cbeg[k + 1];
            for (long long s = l0; s < l1; s++) lu[s * 32] /= piv;
        }
        __syncthreads();
    }
    if (bad) status[b] = ST_SINGULAR;
}

// Fast path: wave schedule.  One CTA per group of 32 matrices, one warp per column of the current wave.  The
// column's slots live in shared memory ([row][32 matrices], every lane touches only its own matrix, so no
// intra-warp synchronisation is needed); L columns of finished columns stream from global memory; the scatter of
// the scaled input values is fused into the column initialisation.
struct KluWaveD {
    int nwaves, spine0;
    const int *wave_col0, *col_roff, *batch_rowslot, *wave_rowsrc;
    const int *ne_cols, *wave_rows, *wrun_ptr;      // wave_col0 holds positions in ne_cols (the columns k_klu_early does not factor)
    const int4* wrun;           // per run of consecutive columns of a wave: {first slot, first shared-memory row, rows, 0}
    const unsigned* bentry;     // per batch [KLU_WAVE_WARPS][KLU_CHUNK_ROWS] staged-row actions
    const unsigned* wblob;      // in-wave update blobs
    const long long *wbatch_ptr, *wblob_ptr;
    const int* bseg_ptr;        // per batch: its staged segments [bseg_ptr[g], bseg_ptr[g+1])
    const int2* segd;           // per segment: {first LU slot, first stage row | rows << 8}: rows consecutive slots = one bulk copy
};

__device__ __forceinline__ void klu_cp_async16(void* smem, const void* gmem) {
    unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sa), "l"(gmem));
}
constexpr int KLU_ENTRY_DOUBLES = (KLU_WAVE_WARPS * KLU_REC_U32 + KLU_CHUNK_ROWS) / 2;   // column records + next row->slot table (uint32), in doubles
constexpr int KLU_STAGE_DOUBLES = KLU_CHUNK_ROWS * 32 + KLU_ENTRY_DOUBLES;      // L rows + per-warp row actions
constexpr size_t KLU_WAVE_SMEM =
    (size_t)(KLU_WAVE_ROWS * 32 + KLU_STAGES * KLU_STAGE_DOUBLES) * sizeof(double) + KLU_BLOB_BYTES;

// mbarrier / TMA-bulk helpers (shared::cta addresses)
__device__ __forceinline__ void klu_mbar_init(unsigned long long* b, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(b)), "r"(count));
}
__device__ __forceinline__ void klu_mbar_expect_tx(unsigned long long* b, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void klu_mbar_arrive(unsigned long long* b) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"((unsigned)__cvta_generic_to_shared(b)) : "memory");
}
__device__ __forceinline__ void klu_mbar_wait(unsigned long long* b, unsigned parity) {
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
__device__ __forceinline__ void klu_bulk_g2s(void* smem, const void* gmem, unsigned bytes, unsigned long long* b) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
                 "r"((unsigned)__cvta_generic_to_shared(smem)), "l"(gmem), "r"(bytes), "r"((unsigned)__cvta_generic_to_shared(b)) : "memory");
}
__device__ __forceinline__ void klu_mbar_arrive_n(unsigned long long* b, unsigned count) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(b)), "r"(count) : "memory");
}
// One staged piece applied to one column of the wave by one warp of the column's team (lane = matrix).
//   x     : the column in shared memory (this lane), rows 32 doubles apart
//   lb    : staged rectangle, this user's first source column: entry (a, i) at lb[(a * nrows + i) * 32]
//   tb    : staged strictly lower triangle of the block (column b's g-1-b entries consecutively)
//   d     : destination row in x of every rectangle row
//   u     : the block's multipliers u_a = x[uloc + a] after the triangle; kept in registers for continuation pieces
// The team's warps take 4-row chunks round robin (sub, T); the main loop is unpredicated, the last partial chunk predicated.
template <int GU>
__device__ __forceinline__ void klu_piece(double* __restrict__ x, const double* __restrict__ lb, const double* __restrict__ tb,
                                          const unsigned short* __restrict__ d, int nrows, int uloc, int g, int s0, bool hastri,
                                          bool cont, int sub, int T, double (&u)[KLU_SN_MAX], double (&pend)[KLU_SN_MAX - 1],
                                          int& pend_row, int& pend_n) {
    if (!cont) {
#pragma unroll
        for (int a = 0; a < GU; a++) u[a] = x[(uloc + a) * 32];
    }
    if (GU > 1 && hastri) {
        // u_a -= sum_{b < a} L(j0+s0+a, j0+s0+b) u_b: computed by every warp of the team (identical values); warp `sub == 0`
        // writes the rows back -- at once when it owns the column alone, else after the team's NEXT barrier, so that no warp
        // of the team can read an already updated row (nobody reads these rows again before that: each row of U belongs to
        // one source column)
#pragma unroll
        for (int b = 0; b < GU - 1; b++) {
            const int B = s0 + b;
            const double* tcol = tb + (B * (g - 1) - (B * (B - 1)) / 2) * 32;     // column B of the triangle: rows B+1 ..
#pragma unroll
            for (int a = b + 1; a < GU; a++) u[a] = fma(-tcol[(a - b - 1) * 32], u[b], u[a]);
        }
        if (sub == 0) {
            if (T == 1) {
#pragma unroll
                for (int a = 1; a < GU; a++) x[(uloc + a) * 32] = u[a];
            } else {
                pend_row = uloc + 1; pend_n = GU - 1;
#pragma unroll
                for (int a = 1; a < GU; a++) pend[a - 1] = u[a];
            }
        }
    }
    int i = 4 * sub;
    const int step = 4 * T;
    for (; i + 4 <= nrows; i += step) {
        // all loads of the chunk are issued before the first multiply-add (the compiler otherwise chains every shared-memory
        // load with its use through one temporary: 4 GU dependent load latencies per chunk)
        const int e0 = d[i], e1 = d[i + 1], e2 = d[i + 2], e3 = d[i + 3];
        double lv[GU][4];
        const double* l = lb + i * 32;
#pragma unroll
        for (int a = 0; a < GU; a++) {
            lv[a][0] = l[0]; lv[a][1] = l[32]; lv[a][2] = l[64]; lv[a][3] = l[96];
            l += nrows * 32;
        }
        double* p0 = x + e0 * 32; double* p1 = x + e1 * 32; double* p2 = x + e2 * 32; double* p3 = x + e3 * 32;
        double a0 = *p0, a1 = *p1, a2 = *p2, a3 = *p3;
        asm volatile("" ::"d"(lv[0][0]), "d"(lv[GU - 1][3]), "d"(a0), "d"(a3));      // scheduling fence: loads above, arithmetic below
#pragma unroll
        for (int a = 0; a < GU; a++) {
            a0 = fma(-lv[a][0], u[a], a0); a1 = fma(-lv[a][1], u[a], a1); a2 = fma(-lv[a][2], u[a], a2); a3 = fma(-lv[a][3], u[a], a3);
        }
        *p0 = a0; *p1 = a1; *p2 = a2; *p3 = a3;
    }
    if (i < nrows) {            // last, partial chunk (1..3 rows): rows past the end repeat row i and are not stored
        const int left = nrows - i;
        const int o1 = left > 1 ? 1 : 0, o2 = left > 2 ? 2 : 0;
        const int e0 = d[i], e1 = d[i + o1], e2 = d[i + o2];
        double lv[GU][3];
        const double* l = lb + i * 32;
#pragma unroll
        for (int a = 0; a < GU; a++) {
            lv[a][0] = l[0]; lv[a][1] = l[o1 * 32]; lv[a][2] = l[o2 * 32];
            l += nrows * 32;
        }
        double* p0 = x + e0 * 32; double* p1 = x + e1 * 32; double* p2 = x + e2 * 32;
        double a0 = *p0, a1 = *p1, a2 = *p2;
        asm volatile("" ::"d"(lv[0][0]), "d"(lv[GU - 1][2]), "d"(a0), "d"(a2));
#pragma unroll
        for (int a = 0; a < GU; a++) {
            a0 = fma(-lv[a][0], u[a], a0); a1 = fma(-lv[a][1], u[a], a1); a2 = fma(-lv[a][2], u[a], a2);
        }
        *p0 = a0;
        if (left > 1) *p1 = a1;
        if (left > 2) *p2 = a2;
    }
}
constexpr int KLU_CONS_BAR = 9;      // named barrier of the 16 consumer warps (team barriers use 1..8)

// Fast path: wave schedule.  One CTA per group of 32 matrices (lane = matrix), one warp per column of
// the wave, plus ONE PRODUCER WARP that streams the staged batches with cp.async.bulk (256-byte rows + the batch's
// records) into a KLU_STAGES-deep ring guarded by full/empty mbarriers: the consumer warps never issue a copy and
// there is no CTA-wide barrier per batch.
//   xs    : the columns of the wave, [row][32 matrices] doubles -- every lane touches only its own matrix
//   stage : KLU_STAGES-deep cp.async ring; a batch = 64 rows of finished L columns (read once from HBM/L2 and
//           consumed by every column of the wave) + for each warp and row what to do with it (destination row, row
//           holding u_jk) so that the update loop reads NO metadata from global memory
//   blob  : the updates between columns of the same wave (applied in rounds from the source's xs region)
// The input values arrive pre-scaled (k_klu_rowscale), gathered straight into xs by cp.async with zero fill.
__global__ void __launch_bounds__((KLU_WAVE_WARPS + 1) * 32, 1) k_klu_refactor_wave(KluPlanD P, KluWaveD W, int Bp,
                                                                              const double* __restrict__ Axs,
                                                                              double* __restrict__ LU, int* __restrict__ status,
                                                                              long long* __restrict__ dbg) {
    extern __shared__ double smem_klu[];
    double* xs = smem_klu;
    double* stage = smem_klu + KLU_WAVE_ROWS * 32;
    double* blob = stage + KLU_STAGES * KLU_STAGE_DOUBLES;
    __shared__ int done_round[KLU_WAVE_WARPS];
    __shared__ unsigned long long full_bar[KLU_STAGES], empty_bar[KLU_STAGES];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int b = blockIdx.x * 32 + lane;
    const double* lug = LU + (long long)blockIdx.x * P.gstride;   // group base for the cooperative copies
    const double* axg = Axs + (long long)blockIdx.x * 32;
    int bad = 0;
    const int srow = tid >> 4, spc = (tid & 15) * 2;
    long long t_init = 0, t_p1 = 0, t_p2 = 0, n_rounds = 0, tA = 0, t_w = 0, t_w0 = 0;
    if (tid == 0) {
        for (int s = 0; s < KLU_STAGES; s++) { klu_mbar_init(&full_bar[s], 1); klu_mbar_init(&empty_bar[s], KLU_WAVE_WARPS); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (warp == KLU_WAVE_WARPS) {
        // ---------------- producer warp: lane r issues staged rows r and r + 32 of every batch
        constexpr unsigned META_BYTES = KLU_WAVE_WARPS * KLU_REC_U32 * 4;
        for (int w = 0; w < W.nwaves; w++) {
            const long long c0 = W.wbatch_ptr[w], c1 = W.wbatch_ptr[w + 1];
            int s0 = 0, s1 = 0;
            int2 sd = make_int2(0, 0);
            if (c0 < c1) { s0 = W.bseg_ptr[c0]; s1 = W.bseg_ptr[c0 + 1]; if (s0 + lane < s1) sd = W.segd[s0 + lane]; }
            for (long long g = c0; g < c1; g++) {
                const int slot = (int)(g % KLU_STAGES);
                const unsigned par = (unsigned)((g / KLU_STAGES) & 1);
                const int a0 = s0, a1 = s1;
                const int2 cur = sd;
                if (g + 1 < c1) {           // the next batch's descriptors are fetched while this one waits for its slot
                    s0 = a1; s1 = W.bseg_ptr[g + 2];
                    sd = make_int2(0, 0);
                    if (s0 + lane < s1) sd = W.segd[s0 + lane];
                }
                klu_mbar_wait(&empty_bar[slot], par ^ 1u);
                double* dst = stage + (long long)slot * KLU_STAGE_DOUBLES;
                // lane q takes segments q (descriptor prefetched), q + 32, ... (a batch holds at most KLU_CHUNK_ROWS one-row segments)
                constexpr int XSEG = (KLU_CHUNK_ROWS + 31) / 32 - 1;
                int2 ex[XSEG > 0 ? XSEG : 1];
                int rows = a0 + lane < a1 ? (cur.y >> 8) : 0;
#pragma unroll
                for (int q = 0; q < XSEG; q++) {
                    ex[q] = make_int2(0, 0);
                    if (a0 + 32 * (q + 1) + lane < a1) { ex[q] = W.segd[a0 + 32 * (q + 1) + lane]; rows += ex[q].y >> 8; }
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) rows += __shfl_xor_sync(0xffffffffu, rows, o);
                if (lane == 0) klu_mbar_expect_tx(&full_bar[slot], (unsigned)rows * 256u + META_BYTES);
                __syncwarp();
                if (a0 + lane < a1)
                    klu_bulk_g2s(dst + (cur.y & 0xff) * 32, lug + (long long)cur.x * 32, (unsigned)(cur.y >> 8) * 256u, &full_bar[slot]);
#pragma unroll
                for (int q = 0; q < XSEG; q++)
                    if (a0 + 32 * (q + 1) + lane < a1)
                        klu_bulk_g2s(dst + (ex[q].y & 0xff) * 32, lug + (long long)ex[q].x * 32, (unsigned)(ex[q].y >> 8) * 256u, &full_bar[slot]);
                if (lane == 0) klu_bulk_g2s(dst + KLU_CHUNK_ROWS * 32, W.bentry + g * KLU_ENTRY_DOUBLES * 2, META_BYTES, &full_bar[slot]);
            }
            asm volatile("bar.sync 0;" ::: "memory");       // end of wave: the consumers stored (and fenced) the wave's columns
        }
        return;
    }
    // Per-wave parameters and the row -> input-entry table of the NEXT wave are fetched while the current wave runs
    // (part 1 right after the gather is issued, the dependent part 2 after the staged phase), so that a wave starts
    // without a chain of dependent global loads.
    __shared__ int rs_tab[KLU_WAVE_ROWS];
    // raw values only: nothing is computed from a prefetched word until the next wave starts (no stall on the loads)
    int n_k0 = W.wave_col0[0], n_k1 = W.wave_col0[1];
    long long n_c0 = W.wbatch_ptr[0], n_c1 = W.wbatch_ptr[1], n_bp0 = W.wblob_ptr[0], n_bp1 = W.wblob_ptr[1];
    int n_cb = 0, n_len = 0, n_roff = 0, n_diag = 0, n_l0 = 0, n_wrows = 0;
    auto team_size = [](int wc) { return wc <= 8 ? KLU_WAVE_WARPS / wc : 1; };      // warps per column (any size, not only 2^k)
    int n_k = 0;
    auto load_part2 = [&](int wn) {
        const int n_wc = n_k1 - n_k0;
        const int c = warp / team_size(n_wc);
        const int k = W.ne_cols[n_k0 + (c < n_wc ? c : 0)];
        n_k = k;
        n_cb = (int)P.cbeg[k];
        n_len = (int)P.cbeg[k + 1] - n_cb;
        n_roff = W.col_roff[k];
        n_diag = P.udiag_slot[k] - n_cb;
        n_l0 = P.lslot0[k] - n_cb;
        n_wrows = W.wave_rows[wn];
    };
    load_part2(0);
    for (int r = tid; r < KLU_WAVE_ROWS; r += KLU_WAVE_WARPS * 32) rs_tab[r] = W.wave_rowsrc[r];
    asm volatile("bar.sync %0, %1;" ::"n"(KLU_CONS_BAR), "n"(KLU_WAVE_WARPS * 32) : "memory");
    for (int w = 0; w < W.nwaves; w++) {
        if (dbg) tA = clock64();
        const int wc = n_k1 - n_k0;
        // team = the warps that share one column: floor(16 / wc) warps
        const int T = team_size(wc), col = warp / T, sub = warp - col * T;
        const bool active = col < wc;
        const int k = n_k;
        const int len = n_len;
        double* x = xs + n_roff * 32 + lane;
        const int diag = n_diag, l0 = n_l0;
        const long long c0 = n_c0;
        const int nb = (int)(n_c1 - n_c0);
        auto team_sync = [&]() { if (T > 1) asm volatile("bar.sync %0, %1;" ::"r"(col + 1), "r"(T * 32) : "memory"); };
        // ---- group 0: gather the (pre-scaled) input values of the wave's columns into xs, and the in-wave blob
        const int wrows = n_wrows;
        constexpr int RS_PER = (KLU_WAVE_ROWS + KLU_WAVE_WARPS * 32 - 1) / (KLU_WAVE_WARPS * 32);
        int rs_next[RS_PER];
        {
            for (int row = srow; row < wrows; row += 32) {
                const int src = rs_tab[row];
                if (src >= 0) klu_cp_async16(xs + row * 32 + spc, axg + (long long)src * Bp + spc);
                else *reinterpret_cast<double2*>(xs + row * 32 + spc) = make_double2(0.0, 0.0);      // fill-in slot
            }
            const long long bp0 = n_bp0;
            const int pieces = (int)(n_bp1 - n_bp0);
            for (int q = tid; q < pieces; q += KLU_WAVE_WARPS * 32) klu_cp_async16(blob + q * 2, W.wblob + (bp0 + q) * 4);
            asm volatile("cp.async.commit_group;");
            if (w + 1 < W.nwaves) {          // part 1 of the next wave's parameters
                n_k0 = W.wave_col0[w + 1]; n_k1 = W.wave_col0[w + 2];
                n_c0 = W.wbatch_ptr[w + 1]; n_c1 = W.wbatch_ptr[w + 2];
                n_bp0 = W.wblob_ptr[w + 1]; n_bp1 = W.wblob_ptr[w + 2];
#pragma unroll
                for (int q = 0; q < RS_PER; q++) {
                    const int r = tid + q * KLU_WAVE_WARPS * 32;
                    rs_next[q] = r < KLU_WAVE_ROWS ? W.wave_rowsrc[(long long)(w + 1) * KLU_WAVE_ROWS + r] : -1;
                }
            }
        }
        asm volatile("cp.async.wait_group 0;");
        asm volatile("bar.sync %0, %1;" ::"n"(KLU_CONS_BAR), "n"(KLU_WAVE_WARPS * 32) : "memory");
        if (w + 1 < W.nwaves) {                                // every thread is done with this wave's table
#pragma unroll
            for (int q = 0; q < RS_PER; q++) {
                const int r = tid + q * KLU_WAVE_WARPS * 32;
                if (r < KLU_WAVE_ROWS) rs_tab[r] = rs_next[q];
            }
        }
        if (dbg) { long long tB = clock64(); t_init += tB - tA; tA = tB; }
        double pend[KLU_SN_MAX - 1], u[KLU_SN_MAX];
        int pend_row = 0, pend_n = 0;
#pragma unroll
        for (int a = 0; a < KLU_SN_MAX; a++) u[a] = 0.0;
#pragma unroll
        for (int a = 0; a < KLU_SN_MAX - 1; a++) pend[a] = 0.0;
        // Only the warps that own (a share of) a column walk the staged batches; the 16 arrivals a ring slot waits for are
        // dealt among them (mbarrier.arrive with a count), so the other warps go straight to the end-of-phase barrier
        // instead of spinning through every batch.
        if (active && nb > 0) {
            const int nact = wc * T;
            const unsigned arrive_cnt = (unsigned)(KLU_WAVE_WARPS / nact + (warp < KLU_WAVE_WARPS % nact ? 1 : 0));
            for (int c = 0; c < nb; c++) {
                const long long g = c0 + c;
                const int buf = (int)(g % KLU_STAGES);
                long long q0 = 0;
                if (dbg) q0 = clock64();
                klu_mbar_wait(&full_bar[buf], (unsigned)((g / KLU_STAGES) & 1));
                if (dbg) { t_w += clock64() - q0; if (c == 0) t_w0 += clock64() - q0; }
                const double* sb = stage + (long long)buf * KLU_STAGE_DOUBLES + lane;
                const unsigned* rec = reinterpret_cast<const unsigned*>(stage + (long long)buf * KLU_STAGE_DOUBLES + KLU_CHUNK_ROWS * 32) +
                                      col * KLU_REC_U32;
                const unsigned short* dd = reinterpret_cast<const unsigned short*>(rec + KLU_REC_HDR);
                const int nseg = (int)rec[0];
                // A piece = rows [i, i + nrows) below a source BLOCK of g <= KLU_SN_MAX consecutive columns of one L supernode, of
                // which this column uses the last gu = g - s0 (fill closure).  The block's own rows (u_jk, consecutive rows of
                // x) first go through the block's unit lower triangle in registers; then every destination row is read and
                // written ONCE for gu multiply-adds.  The code is specialised on gu (no predicated-off multiply-adds) and the
                // next piece's record words are fetched before the current piece runs.
                unsigned w0 = rec[1], w1 = rec[2];
                for (int sgi = 0; sgi < nseg; sgi++) {
                    const unsigned nw0 = rec[3 + 2 * sgi], nw1 = rec[4 + 2 * sgi];      // (one pair past the last piece: inside the record)
                    const int r0 = w0 & 0xffu, nrows = (w0 >> 8) & 0xffu, uloc = (int)(w0 >> 16);
                    const int gg = w1 & 0xfu, s0 = (w1 >> 4) & 0xfu, tri0 = (w1 >> 8) & 0xffu;
                    const bool hastri = (w1 >> 16) & 1u, cont = (w1 >> 17) & 1u;
                    const int gu = gg - s0;
                    team_sync();          // the previous piece of every warp of the team is complete
                    if (pend_n) {         // rows of the previous block's triangle (see klu_piece)
#pragma unroll
                        for (int a = 0; a < KLU_SN_MAX - 1; a++) if (a < pend_n) x[(pend_row + a) * 32] = pend[a];
                        pend_n = 0;
                    }
                    const double* lb = sb + (r0 + s0 * nrows) * 32;
                    const double* tb = sb + tri0 * 32;
                    switch (gu) {
                        case 1: klu_piece<1>(x, lb, tb, dd + r0, nrows, uloc, gg, s0, hastri, cont, sub, T, u, pend, pend_row, pend_n); break;
                        case 2: klu_piece<2>(x, lb, tb, dd + r0, nrows, uloc, gg, s0, hastri, cont, sub, T, u, pend, pend_row, pend_n); break;
                        case 3: klu_piece<3>(x, lb, tb, dd + r0, nrows, uloc, gg, s0, hastri, cont, sub, T, u, pend, pend_row, pend_n); break;
                        default: klu_piece<4>(x, lb, tb, dd + r0, nrows, uloc, gg, s0, hastri, cont, sub, T, u, pend, pend_row, pend_n); break;
                    }
                    w0 = nw0; w1 = nw1;
                }
                __syncwarp();
                if (lane == 0) klu_mbar_arrive_n(&empty_bar[buf], arrive_cnt);
            }
        }
        if (active && T > 1) team_sync();      // every warp of the team is past the reads of its last piece
        if (pend_n) {             // (T > 1, warp sub == 0) the last triangle rows of this wave
#pragma unroll
            for (int a = 0; a < KLU_SN_MAX - 1; a++) if (a < pend_n) x[(pend_row + a) * 32] = pend[a];
            pend_n = 0;
        }
        if (tid < KLU_WAVE_WARPS) done_round[tid] = 0x7fffffff;
        if (w + 1 < W.nwaves) load_part2(w + 1);      // part 2 of the next wave's parameters (depends on part 1)
        asm volatile("bar.sync %0, %1;" ::"n"(KLU_CONS_BAR), "n"(KLU_WAVE_WARPS * 32) : "memory");
        if (dbg) { long long tB = clock64(); t_p1 += tB - tA; tA = tB; }
        // ---- sources inside the wave: rounds.  In round r a column consumes (in pivot order) the in-wave sources
        // finalized in rounds < r from their xs regions, and finalizes itself once all its updates are applied.
        const unsigned* bl = reinterpret_cast<const unsigned*>(blob);
        int ui = active ? (int)bl[2 * col] : 0;
        const int ue = active ? ui + (int)bl[2 * col + 1] : 0;
        const int nupd_wave = (int)(bl[2 * (wc - 1)] + bl[2 * (wc - 1) + 1]);
        const unsigned* updl = bl + 2 * KLU_WAVE_WARPS;
        const unsigned short* bdst = reinterpret_cast<const unsigned short*>(updl + 4 * nupd_wave);
        bool fin = !active;
        for (int r = 0;; r++) {
            if (!fin) {
                while (ui < ue) {
                    const unsigned w0 = updl[4 * ui];
                    if (done_round[w0 & 0xffu] >= r) break;
                    team_sync();                  // the previous update (or the staged phase) of every team warp is done
                    const double uj = x[updl[4 * ui + 1] * 32];
                    const int cnt = (int)updl[4 * ui + 2];
                    const unsigned short* d = bdst + updl[4 * ui + 3];
                    const double* lsrc = xs + (w0 >> 8) * 32 + lane;
                    for (int t = sub * 4; t < cnt; t += 4 * T) {
                        int e[4]; double m[4], xv[4];
#pragma unroll
                        for (int q = 0; q < 4; q++) {
                            const bool on = t + q < cnt;
                            e[q] = on ? d[t + q] * 32 : 0;
                            m[q] = on ? lsrc[(t + q) * 32] : 0.0;
                        }
#pragma unroll
                        for (int q = 0; q < 4; q++) xv[q] = x[e[q]];
#pragma unroll
                        for (int q = 0; q < 4; q++) if (t + q < cnt) x[e[q]] = xv[q] - m[q] * uj;
                    }
                    ui++;
                }
                if (ui == ue && k >= W.spine0) {
                    // column of the dense trailing block: only the updates from columns < spine0 were applied here;
                    // store it unfinished, k_klu_dense_lu factors the block
                    fin = true;
                    if (lane == 0 && sub == 0) done_round[col] = r;
                } else if (ui == ue) {
                    team_sync();
                    const double piv = x[diag * 32];
                    if (!(fabs(piv) > 0.0)) bad = 1;
                    team_sync();                  // everyone has read the pivot before the column is rewritten
                    const double rpiv = 1.0 / piv;    // one division per column; L(:,k) = x * (1/pivot)
#pragma unroll 4
                    for (int sl = l0 + sub; sl < len; sl += T) x[sl * 32] *= rpiv;
                    fin = true;
                    if (lane == 0 && sub == 0) done_round[col] = r;
                }
            }
            n_rounds++;
            {
                unsigned left;
                asm volatile("{\n.reg .pred p;\nsetp.ne.u32 p, %1, 0;\nbar.red.popc.u32 %0, %2, %3, p;\n}\n"
                             : "=r"(left) : "r"((unsigned)!fin), "n"(KLU_CONS_BAR), "n"(KLU_WAVE_WARPS * 32) : "memory");
                if (left == 0) break;
            }
        }
        if (dbg) { long long tB = clock64(); t_p2 += tB - tA; tA = tB; if (tid == 0 && blockIdx.x == 0) { dbg[8 + 2 * w] = t_init + t_p1; dbg[9 + 2 * w] = t_p2; } }
        // a run of consecutive columns of the wave owns consecutive slots and consecutive xs rows: one bulk store (TMA) per run
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("bar.sync %0, %1;" ::"n"(KLU_CONS_BAR), "n"(KLU_WAVE_WARPS * 32) : "memory");
        if (tid == 0) {
            for (int rn = W.wrun_ptr[w]; rn < W.wrun_ptr[w + 1]; rn++) {
                const int4 run = W.wrun[rn];
                asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(lug + (long long)run.x * 32),
                             "r"((unsigned)__cvta_generic_to_shared(xs + run.y * 32)), "r"((unsigned)run.z * 256u) : "memory");
            }
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
        }
        asm volatile("bar.sync 0;" ::: "memory");      // the producer may now read the wave's columns; xs is free
    }
    if (dbg && tid == 0 && blockIdx.x == 0) { dbg[0] = t_init; dbg[1] = t_p1; dbg[2] = t_p2; dbg[3] = n_rounds; dbg[4] = t_w; dbg[5] = t_w0; dbg[6] = 0; }
    if (bad) status[b] = ST_SINGULAR;
}

// Dense trailing block: one CTA per matrix gathers the nd x nd block (pattern entries from the LU slots, zeros
// elsewhere) into shared memory, factors it without pivoting (the pivot order is the frozen one) by 16-column
// chunks -- diagonal chunk by one warp, L21 = A21 U11^-1 and U12 = L11^-1 A12 one thread per row / column, trailing
// update with FP64 DMMA on 8x8 tiles -- and scatters the pattern entries back.  Entries outside the symbolic
// pattern stay exactly zero (the pattern is closed under fill).
__device__ __forceinline__ void klu_dmma884(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
// diagonal 16 x 16 chunk at (d0, d0) factored in registers by one warp: lane r (and r + 16) owns row d0 + r
__device__ __forceinline__ int klu_dense_diag(double* S, int lds, int d0, int lane, double* rdiag) {
    const int r = lane & 15;
    double a[16];
#pragma unroll
    for (int c = 0; c < 16; c++) a[c] = S[(d0 + c) * lds + d0 + r];
    int flag = 0;
#pragma unroll
    for (int j = 0; j < 16; j++) {
        const double piv = __shfl_sync(0xffffffffu, a[j], j);
        if (!(fabs(piv) > 0.0)) flag = 1;
        const double rp = __drcp_rn(piv);
        const bool below = r > j;
        const double lij = a[j] * rp;
        if (below) a[j] = lij;
#pragma unroll
        for (int c = 0; c < 16; c++)
            if (c > j) {                 // static after unrolling: keeps a[] in registers
                const double u = __shfl_sync(0xffffffffu, a[c], j);
                const double t = fma(-lij, u, a[c]);
                a[c] = below ? t : a[c];
            }
        if (lane == j) rdiag[j] = rp;
    }
    if (lane < 16) {
#pragma unroll
        for (int c = 0; c < 16; c++) S[(d0 + c) * lds + d0 + r] = a[c];
    }
    return flag;
}

// C(8 x 32 strip at tile row ti, tile columns tj0..tj0+3, those in qmask) -= L(:, c0..c0+15) U(c0..c0+15, :)
template <bool FULL>
__device__ __forceinline__ void klu_dense_strip(double* S, int lds, int c0, int t0, int ti, int tj0, int qmask, int lane) {
    double* Cp = S + (t0 + 8 * tj0 + 2 * (lane & 3)) * lds + t0 + 8 * ti + (lane >> 2);      // C(row, col), col += 8 per q
    const double* Ap = S + (c0 + (lane & 3)) * lds + t0 + 8 * ti + (lane >> 2);               // L(row, c0 + k)
    const double* Bp = S + (t0 + 8 * tj0 + (lane >> 2)) * lds + c0 + (lane & 3);              // U(c0 + k, col)
    const int q8 = 8 * lds;
    double acc[4][2];
#pragma unroll
    for (int q = 0; q < 4; q++)
        if (FULL || (qmask >> q & 1)) {
            acc[q][0] = Cp[q * q8];
            acc[q][1] = Cp[q * q8 + lds];
        }
#pragma unroll
    for (int k4 = 0; k4 < 16; k4 += 4) {
        const double av = -Ap[k4 * lds];
#pragma unroll
        for (int q = 0; q < 4; q++)
            if (FULL || (qmask >> q & 1)) klu_dmma884(acc[q][0], acc[q][1], av, Bp[q * q8 + k4]);
    }
#pragma unroll
    for (int q = 0; q < 4; q++)
        if (FULL || (qmask >> q & 1)) {
            Cp[q * q8] = acc[q][0];
            Cp[q * q8 + lds] = acc[q][1];
        }
}

// LU[slot][matrix] (dir 0) -> D[matrix][entry] for the entries of the dense trailing block, and back (dir 1): 32 x 32
// tiles through shared memory so both sides move full 256-byte rows
__global__ void __launch_bounds__(256) k_klu_dense_pack(const int* __restrict__ dslot, int ndmap, int ndp, long long gstride, double* __restrict__ LU,
                                                        double* __restrict__ D, int dir) {
    __shared__ double tile[32][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int e0 = blockIdx.x * 32, b0 = blockIdx.y * 32;
    if (dir == 0) {
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const int e = e0 + ty + 8 * i;
            if (e < ndmap) tile[ty + 8 * i][tx] = LU[(long long)(b0 >> 5) * gstride + (long long)dslot[e] * 32 + tx];
        }
        __syncthreads();
#pragma unroll
        for (int i = 0; i < 4; i++)
            if (e0 + tx < ndmap) D[(long long)(b0 + ty + 8 * i) * ndp + e0 + tx] = tile[tx][ty + 8 * i];
    } else {
#pragma unroll
        for (int i = 0; i < 4; i++)
            if (e0 + tx < ndmap) tile[tx][ty + 8 * i] = D[(long long)(b0 + ty + 8 * i) * ndp + e0 + tx];
        __syncthreads();
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const int e = e0 + ty + 8 * i;
            if (e < ndmap) LU[(long long)(b0 >> 5) * gstride + (long long)dslot[e] * 32 + tx] = tile[ty + 8 * i][tx];
        }
    }
}

// D[matrix][entry]: the block entries of one matrix, contiguous (k_klu_dense_pack), so gather and scatter are coalesced.
// dmeta: per column of the block KLU_DENSE_META ints = {index of its first entry, row bitmap (KLU_DENSE_MAX bits)}
constexpr int KLU_DENSE_THREADS = 256;
// two CTAs per SM (64 registers per thread): the latency-bound chunk chains of two matrices overlap
constexpr int KLU_DENSE_OCC = 2;
__global__ void __launch_bounds__(KLU_DENSE_THREADS, KLU_DENSE_OCC) k_klu_dense_lu(int nd, const int* __restrict__ dmeta, int ndp, int batch,
                                                         double* __restrict__ D, int* __restrict__ status) {
    extern __shared__ double S[];
    __shared__ double rdiag[16];
    __shared__ int bad, tctr;
    const int b = blockIdx.x;
    if (b >= batch) return;
    const int lds = nd + 4, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    int* meta = reinterpret_cast<int*>(S + nd * lds);
    double* dm = D + (long long)b * ndp;      // this matrix' block entries, column by column, rows ascending
    {
        double2* S2 = reinterpret_cast<double2*>(S);
        for (int i = tid; i < (nd * lds) >> 1; i += KLU_DENSE_THREADS) S2[i] = make_double2(0.0, 0.0);
        for (int i = tid; i < nd * KLU_DENSE_META; i += KLU_DENSE_THREADS) meta[i] = dmeta[i];
    }
    if (tid == 0) bad = 0;
    __syncthreads();
    {   // gather straight into shared memory (LDGSTS): every request of the CTA is in flight at once
        const unsigned sbase = (unsigned)__cvta_generic_to_shared(S);
        for (int c = warp; c < nd; c += KLU_DENSE_THREADS / 32) {
            const int* m = meta + c * KLU_DENSE_META;
            int slot = m[0];
#pragma unroll
            for (int w = 0; w < KLU_DENSE_META - 1; w++) {
                const unsigned bits = (unsigned)m[1 + w];
                if (bits >> lane & 1) {
                    const double* src = dm + slot + __popc(bits & ((1u << lane) - 1u));
                    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(sbase + 8u * (unsigned)(c * lds + 32 * w + lane)), "l"(src));
                }
                slot += __popc(bits);
            }
        }
        asm volatile("cp.async.commit_group;\n cp.async.wait_group 0;" ::: "memory");
    }
    __syncthreads();
    if (warp == 0 && klu_dense_diag(S, lds, 0, lane, rdiag)) bad = 1;
    for (int c0 = 0; c0 < nd; c0 += 16) {
        __syncthreads();                       // diagonal chunk c0 factored, trailing update of the previous chunk done
        const int t0 = c0 + 16, rem = nd - t0;
        if (tid == 0) tctr = 0;
        for (int idx = tid; idx < 2 * rem; idx += KLU_DENSE_THREADS) {
            double v[16];
            if (idx < rem) {                   // row r of L21: x U11 = a
                const int r = t0 + idx;
#pragma unroll
                for (int q = 0; q < 16; q++) v[q] = S[(c0 + q) * lds + r];
#pragma unroll
                for (int q = 0; q < 16; q++) {
                    double acc = v[q];
#pragma unroll
                    for (int pp = 0; pp < q; pp++) acc = fma(-v[pp], S[(c0 + q) * lds + c0 + pp], acc);
                    v[q] = acc * rdiag[q];
                }
#pragma unroll
                for (int q = 0; q < 16; q++) S[(c0 + q) * lds + r] = v[q];
            } else {                           // column c of U12: L11 u = a (unit lower)
                const int c = t0 + idx - rem;
#pragma unroll
                for (int q = 0; q < 16; q++) v[q] = S[c * lds + c0 + q];
#pragma unroll
                for (int q = 0; q < 16; q++) {
                    double acc = v[q];
#pragma unroll
                    for (int pp = 0; pp < q; pp++) acc = fma(-S[(c0 + pp) * lds + c0 + q], v[pp], acc);
                    v[q] = acc;
                }
#pragma unroll
                for (int q = 0; q < 16; q++) S[c * lds + c0 + q] = v[q];
            }
        }
        __syncthreads();
        // trailing update by 8 x 32 strips handed out through a counter; warp 0 first updates the next diagonal chunk
        // and factors it while the other warps work on the rest (look-ahead)
        const int nt = rem >> 3, ntj = (nt + 3) >> 2;
        if (warp == 0 && rem > 0) {
            klu_dense_strip<false>(S, lds, c0, t0, 0, 0, 3, lane);
            klu_dense_strip<false>(S, lds, c0, t0, 1, 0, 3, lane);
            __syncwarp();
            if (klu_dense_diag(S, lds, t0, lane, rdiag)) bad = 1;
        }
        const int nstrips = nt * ntj, rcp = (65536 + ntj - 1) / max(ntj, 1);
        int mt = 0;
        if (lane == 0) mt = atomicAdd(&tctr, 1);
        mt = __shfl_sync(0xffffffffu, mt, 0);
        while (mt < nstrips) {
            int nxt = 0;
            if (lane == 0) nxt = atomicAdd(&tctr, 1);      // next ticket fetched under the DMMA work of this strip
            const int ti = (mt * rcp) >> 16, tj0 = (mt - ti * ntj) * 4;
            int qmask = (nt - tj0 >= 4) ? 15 : ((1 << (nt - tj0)) - 1);
            if (ti < 2 && tj0 == 0) qmask &= ~3;
            if (qmask == 15) klu_dense_strip<true>(S, lds, c0, t0, ti, tj0, 15, lane);
            else if (qmask) klu_dense_strip<false>(S, lds, c0, t0, ti, tj0, qmask, lane);
            mt = __shfl_sync(0xffffffffu, nxt, 0);
        }
    }
    __syncthreads();
    for (int c = warp; c < nd; c += KLU_DENSE_THREADS / 32) {
        const int* m = meta + c * KLU_DENSE_META;
        int slot = m[0];
#pragma unroll
        for (int w = 0; w < KLU_DENSE_META - 1; w++) {
            const unsigned bits = (unsigned)m[1 + w];
            if (bits >> lane & 1) dm[slot + __popc(bits & ((1u << lane) - 1u))] = S[c * lds + 32 * w + lane];
            slot += __popc(bits);
        }
    }
    if (tid == 0 && bad) status[b] = ST_SINGULAR;
}

// ---- batched triangular solves: level-scheduled gather tasks ----------------------------------------
// klu_solve / klu_tsolve (reference src/C/klu.c:593-690) as a DAG of 2n tasks over the work vector V = [Y; Z]
// (2n rows, interleaved [row][matrix]):  V[t] = (V[i] - sum_terms LU[slot] * V[src]) / LU[diag]   with i = t mod n.
//   'N' (row gather):    task i   = row i of L  (Y[i] = b~[i] - L(i,:) Y - F(i,:) Z, F couples to the later BTF blocks)
//                        task n+i = row i of U  (Z[i] = (Y[i] - U(i,i+1:) Z) / U(i,i))
//   'T' (column gather): task k   = column k of U and F (Y[k] = (b~[k] - U(:k,k)' Y - F(:,k)' Z) / U(k,k))
//                        task n+k = column k of L       (Z[k] = Y[k] - L(k+1:,k)' Z)
// Tasks are sorted by DAG level; one CTA (16 warps) serves 32 matrices (lane = matrix), the tasks of a level are
// dealt to the warps, one block barrier per level.  Task records and the first 32 terms of a warp's next task are
// prefetched before the barrier (they are static), so a level costs about one round trip for the values.
struct KluSolveLvlD {
    int n, nlev;
    const int* lvl_ptr;     // nlev + 1
    const int4* rec;        // per task in level order: {task id t, first term, end term, diagonal slot or -1}
    const int* tslot;       // value slot of a term
    const int* tsrc;        // row of V it multiplies
    const int* extra;       // per part: split levels: head = number of further parts of its task, others = -1
    const int* mode;        // per level: 1 = split mode
};
constexpr int KLU_SOLVE_WARPS = 16;
constexpr int KLU_SOLVE_PART = 16;        // terms per part of a split level = value pairs in flight per round
constexpr int KLU_SOLVE_MAXPARTS = 192;   // parts of one split level (shared-memory partial sums: 192 x 32 doubles)
struct KluSolveLvlH {
    std::vector<int> lvl_ptr, tslot, tsrc, extra, mode;
    std::vector<int4> rec;
    int nlev = 0;
};

static void build_solve_levels(const KluSymbolic& S, const KluNumeric& N, const KluPlan& P, bool trans, KluSolveLvlH& H) {
    const int n = P.n;
    // per task t (Y rows 0..n-1, Z rows n..2n-1): terms [tp[t], tp[t+1]) = (slot, src row of V), in column order; built
    // in two passes over the factor (count, fill) -- no per-task containers
    std::vector<int> tp(2 * (size_t)n + 1, 0), tsl, tsr;
    std::vector<int> diag(2 * (size_t)n, -1);
    auto sweep = [&](auto&& emit) {
        for (int k = 0; k < n; k++) {
            const long long u0 = N.Up[k], u1 = N.Up[k + 1] - 1, l0 = N.Lp[k] + 1, l1 = N.Lp[k + 1], f0 = N.Fp[k], f1 = N.Fp[k + 1];
            const int us = (int)P.cbeg[k], ls = P.lslot0[k], fs = P.fslot0[k];
            if (!trans) {
                for (long long p = l0; p < l1; p++) emit(N.Li[p], ls + (int)(p - l0), k);                  // L(i,k) Y[k]
                for (long long p = f0; p < f1; p++) emit(N.Fi[p], fs + (int)(p - f0), n + k);              // F(i,k) Z[k]
                for (long long p = u0; p < u1; p++) emit(n + N.Ui[p], us + (int)(p - u0), n + k);          // U(i,k) Z[k]
            } else {
                for (long long p = u0; p < u1; p++) emit(k, us + (int)(p - u0), N.Ui[p]);                  // U(p,k) Y[p]
                for (long long p = f0; p < f1; p++) emit(k, fs + (int)(p - f0), n + N.Fi[p]);              // F(p,k) Z[p]
                for (long long p = l0; p < l1; p++) emit(n + k, ls + (int)(p - l0), n + N.Li[p]);          // L(p,k) Z[p]
            }
        }
    };
    sweep([&](int t, int, int) { tp[t + 1]++; });
    for (int t = 0; t < 2 * n; t++) tp[t + 1] += tp[t];
    tsl.resize(tp[2 * (size_t)n]); tsr.resize(tp[2 * (size_t)n]);
    {
        std::vector<int> cur(tp.begin(), tp.end() - 1);
        sweep([&](int t, int slot, int src) { const int q = cur[t]++; tsl[q] = slot; tsr[q] = src; });
    }
    for (int k = 0; k < n; k++) diag[(trans ? 0 : (size_t)n) + k] = P.udiag_slot[k];
    auto nterms = [&](int t) { return tp[t + 1] - tp[t]; };
    // levels in a topological order of the sequential algorithm
    std::vector<int> level(2 * (size_t)n, 0);
    auto visit = [&](int t) {
        int lv = 0;
        if (t >= n) lv = level[t - n] + 1;                        // Z-task of i needs Y[i]
        for (int q = tp[t]; q < tp[t + 1]; q++) lv = std::max(lv, level[tsr[q]] + 1);
        level[t] = lv;
    };
    if (!trans)
        for (int blk = S.nblocks - 1; blk >= 0; blk--) {
            for (int i = S.R[blk]; i < S.R[blk + 1]; i++) visit(i);
            for (int i = S.R[blk + 1] - 1; i >= S.R[blk]; i--) visit(n + i);
        }
    else
        for (int blk = 0; blk < S.nblocks; blk++) {
            for (int k = S.R[blk]; k < S.R[blk + 1]; k++) visit(k);
            for (int k = S.R[blk + 1] - 1; k >= S.R[blk]; k--) visit(n + k);
        }
    int nlev = 0;
    for (int t = 0; t < 2 * n; t++) nlev = std::max(nlev, level[t] + 1);
    H.nlev = nlev;
    H.lvl_ptr.assign(nlev + 1, 0);
    for (int t = 0; t < 2 * n; t++) H.lvl_ptr[level[t] + 1]++;
    for (int l = 0; l < nlev; l++) H.lvl_ptr[l + 1] += H.lvl_ptr[l];
    std::vector<int> pos(H.lvl_ptr.begin(), H.lvl_ptr.end() - 1), order(2 * (size_t)n);
    // long tasks first inside a level: the warps that get them start early
    std::vector<int> ids(2 * (size_t)n);
    for (int t = 0; t < 2 * n; t++) ids[t] = t;
    std::stable_sort(ids.begin(), ids.end(), [&](int x, int y) {
        if (level[x] != level[y]) return level[x] < level[y];
        return nterms(x) > nterms(y);
    });
    // Emit the work items ("parts") level by level.  A level with fewer tasks than warps whose tasks are long is
    // run in split mode: every task is cut into parts of <= 32 terms that different warps sum concurrently into
    // shared memory; after a barrier the head part's warp combines them (fixed order) and stores the row.
    H.rec.clear(); H.extra.clear(); H.mode.assign(nlev, 0);
    H.tslot.clear(); H.tsrc.clear();
    std::vector<int> lp2(nlev + 1, 0);
    for (int l = 0; l < nlev; l++) {
        const int q0 = H.lvl_ptr[l], q1 = H.lvl_ptr[l + 1];
        int parts = 0;
        for (int q = q0; q < q1; q++) parts += std::max<int>(1, (nterms(ids[q]) + KLU_SOLVE_PART - 1) / KLU_SOLVE_PART);
        const bool split = (q1 - q0) < KLU_SOLVE_WARPS && parts > (q1 - q0) && parts <= KLU_SOLVE_MAXPARTS;
        H.mode[l] = split ? 1 : 0;
        for (int q = q0; q < q1; q++) {
            const int t = ids[q];
            const int base = (int)H.tslot.size();
            H.tslot.insert(H.tslot.end(), tsl.begin() + tp[t], tsl.begin() + tp[t + 1]);
            H.tsrc.insert(H.tsrc.end(), tsr.begin() + tp[t], tsr.begin() + tp[t + 1]);
            const int cnt = nterms(t);
            if (!split) {
                H.rec.push_back(make_int4(t, base, base + cnt, diag[t]));
                H.extra.push_back(0);
            } else {
                const int np = std::max(1, (cnt + KLU_SOLVE_PART - 1) / KLU_SOLVE_PART);
                for (int pi = 0; pi < np; pi++) {
                    H.rec.push_back(make_int4(t, base + KLU_SOLVE_PART * pi, std::min(base + cnt, base + KLU_SOLVE_PART * (pi + 1)), pi == 0 ? diag[t] : -1));
                    H.extra.push_back(pi == 0 ? np - 1 : -1);
                }
            }
        }
        lp2[l + 1] = (int)H.rec.size();
    }
    H.lvl_ptr = lp2;
    (void)pos; (void)order;
}

// V[k][b] <- B_b[perm[k]] (/ Rs[k][b] for 'N'), and back: 32 x 32 tiles through shared memory
__global__ void __launch_bounds__(256) k_klu_solve_load(int n, int batch, int Bp, const int* __restrict__ perm, const double* __restrict__ Rs,
                                                        const double* __restrict__ B, long long ldB, long long bstride, double* __restrict__ V) {
    __shared__ double tile[32][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5, k0 = blockIdx.x * 32, b0 = blockIdx.y * 32, rhs = blockIdx.z;
    double* v = V + (long long)rhs * 2 * n * Bp;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const int b = b0 + ty + 8 * i, k = k0 + tx;
        tile[ty + 8 * i][tx] = (b < batch && k < n) ? B[(long long)b * bstride + (long long)rhs * ldB + perm[k]] : 0.0;
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const int k = k0 + ty + 8 * i, b = b0 + tx;
        if (k < n) {
            double x = tile[tx][ty + 8 * i];
            if (Rs) x = (b < batch) ? x / Rs[(long long)k * Bp + b] : 0.0;
            v[(long long)k * Bp + b] = x;
        }
    }
}
__global__ void __launch_bounds__(256) k_klu_solve_store(int n, int batch, int Bp, const int* __restrict__ perm, const double* __restrict__ Rs,
                                                         double* __restrict__ B, long long ldB, long long bstride, const double* __restrict__ V) {
    __shared__ double tile[32][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5, k0 = blockIdx.x * 32, b0 = blockIdx.y * 32, rhs = blockIdx.z;
    const double* z = V + (long long)rhs * 2 * n * Bp + (long long)n * Bp;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const int k = k0 + ty + 8 * i, b = b0 + tx;
        double x = 0.0;
        if (k < n && b < batch) { x = z[(long long)k * Bp + b]; if (Rs) x /= Rs[(long long)k * Bp + b]; }
        tile[ty + 8 * i][tx] = x;
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const int b = b0 + ty + 8 * i, k = k0 + tx;
        if (b < batch && k < n) B[(long long)b * bstride + (long long)rhs * ldB + perm[k]] = tile[tx][ty + 8 * i];
    }
}

// sum of LU[slot] * V[src] over the terms [p0, p1); have: the metadata of the first 32 terms is already in ps / pj.
// 16 value pairs are in flight per round (a part of a split level is one round).
__device__ __forceinline__ double klu_solve_terms(const int* __restrict__ tslot, const int* __restrict__ tsrc, int p0, int p1,
                                                  int ps, int pj, bool have, int lane,
                                                  const double* __restrict__ lu, const double* __restrict__ v, int Bp) {
    double s0 = 0.0, s1 = 0.0;
    for (int p = p0; p < p1; p += 32) {
        if (p != p0 || !have) { ps = 0; pj = 0; if (p + lane < p1) { ps = tslot[p + lane]; pj = tsrc[p + lane]; } }
        const int cnt = min(32, p1 - p);
        for (int u = 0; u < cnt; u += KLU_SOLVE_PART) {
            double a[KLU_SOLVE_PART], x[KLU_SOLVE_PART];
#pragma unroll
            for (int j = 0; j < KLU_SOLVE_PART; j++) {
                const int sl = __shfl_sync(0xffffffffu, ps, (u + j) & 31), sr = __shfl_sync(0xffffffffu, pj, (u + j) & 31);
                const bool on = u + j < cnt;
                a[j] = on ? lu[(long long)sl * 32] : 0.0;
                x[j] = on ? v[(long long)sr * Bp] : 0.0;
            }
#pragma unroll
            for (int j = 0; j < KLU_SOLVE_PART; j += 2) { s0 = fma(a[j], x[j], s0); s1 = fma(a[j + 1], x[j + 1], s1); }
        }
    }
    return s0 + s1;
}

__global__ void __launch_bounds__(KLU_SOLVE_WARPS * 32) k_klu_solve_lvl(KluSolveLvlD S, int Bp, long long gstride, const double* __restrict__ LU,
                                                                         double* __restrict__ V) {
    extern __shared__ double smd[];
    double* part = smd;                                            // [KLU_SOLVE_MAXPARTS][32] partial sums of a split level
    int* lp = reinterpret_cast<int*>(smd + KLU_SOLVE_MAXPARTS * 32);   // level pointers, then modes
    int* md = lp + S.nlev + 2;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, n = S.n, nlev = S.nlev;
    const int4* __restrict__ rec = S.rec;
    const int* __restrict__ tslot = S.tslot;
    const int* __restrict__ tsrc = S.tsrc;
    const int* __restrict__ extra = S.extra;
    for (int i = threadIdx.x; i <= nlev; i += KLU_SOLVE_WARPS * 32) lp[i] = S.lvl_ptr[i];
    for (int i = threadIdx.x; i < nlev; i += KLU_SOLVE_WARPS * 32) md[i] = S.mode[i];
    if (threadIdx.x == 0) lp[nlev + 1] = lp[nlev];
    __syncthreads();
    const int b = blockIdx.x * 32 + lane;
    const double* lu = LU + (long long)blockIdx.x * gstride + lane;
    double* v = V + (long long)blockIdx.y * 2 * n * Bp + b;
    // static metadata of this warp's first part of a level is fetched one level ahead (rN..), under the previous level
    int4 r = make_int4(0, 0, 0, -1), rN = r;
    int ps = 0, pj = 0, ex = 0, psN = 0, pjN = 0, exN = 0;
    bool have = false, haveN = false;
    if (nlev > 0 && lp[0] + warp < lp[1]) {
        const int q = lp[0] + warp;
        r = rec[q]; ex = extra[q];
        if (r.y + lane < r.z) { ps = tslot[r.y + lane]; pj = tsrc[r.y + lane]; }
        have = true;
    }
    for (int lev = 0; lev < nlev; lev++) {
        const int q0 = lp[lev], qe = lp[lev + 1];
        haveN = false;
        if (lev + 1 < nlev && qe + warp < lp[lev + 2]) {
            const int q = qe + warp;
            rN = rec[q]; exN = extra[q];
            psN = 0; pjN = 0;
            if (rN.y + lane < rN.z) { psN = tslot[rN.y + lane]; pjN = tsrc[rN.y + lane]; }
            haveN = true;
        }
        if (!md[lev]) {
            for (int q = q0 + warp; q < qe; q += KLU_SOLVE_WARPS) {
                const bool mine = have && q == q0 + warp;
                if (!mine) r = rec[q];
                const int t = r.x, i = (t < n) ? t : t - n;
                const double y = v[(long long)i * Bp];
                double dg = 1.0;
                if (r.w >= 0) dg = lu[(long long)r.w * 32];
                double acc = y - klu_solve_terms(tslot, tsrc, r.y, r.z, ps, pj, mine, lane, lu, v, Bp);
                if (r.w >= 0) acc /= dg;
                v[(long long)t * Bp] = acc;
            }
        } else {
            // pass 1: every part (<= KLU_SOLVE_PART terms) -> its partial sum; a task's head also fetches V[i], diagonal
            double y = 0.0, dg = 1.0;
            int ht = -1, hq = 0, hex = 0, hw = -1;
            for (int q = q0 + warp; q < qe; q += KLU_SOLVE_WARPS) {
                const bool mine = have && q == q0 + warp;
                if (!mine) { r = rec[q]; ex = extra[q]; }
                if (ex >= 0 && ht < 0) {       // first head of this warp: keep its operands in registers
                    const int i = (r.x < n) ? r.x : r.x - n;
                    y = v[(long long)i * Bp];
                    if (r.w >= 0) dg = lu[(long long)r.w * 32];
                    ht = r.x; hq = q; hex = ex; hw = r.w;
                }
                part[(q - q0) * 32 + lane] = klu_solve_terms(tslot, tsrc, r.y, r.z, ps, pj, mine, lane, lu, v, Bp);
            }
            __syncthreads();
            // pass 2: heads combine their parts in order and store the row
            for (int q = q0 + warp; q < qe; q += KLU_SOLVE_WARPS) {
                int t, e, w;
                if (q == hq && ht >= 0) { t = ht; e = hex; w = hw; }
                else {
                    e = extra[q];
                    if (e < 0) continue;
                    const int4 rr = rec[q];
                    t = rr.x; w = rr.w;
                    const int i = (t < n) ? t : t - n;
                    y = v[(long long)i * Bp];
                    dg = (w >= 0) ? lu[(long long)w * 32] : 1.0;
                }
                double ssum = 0.0;
                for (int k = 0; k <= e; k++) ssum += part[(q - q0 + k) * 32 + lane];
                double acc = y - ssum;
                if (w >= 0) acc /= dg;
                v[(long long)t * Bp] = acc;
            }
        }
        // the factor rows of this warp's first task of the next level are static: pull them into L2 while the level's
        // barrier is pending (lane q holds term q's slot), so that the dependent chain of the next level starts from L2
        if (haveN && rN.y + lane < rN.z) {
            const double* lrow = lu - lane + (long long)psN * 32;
            asm volatile("prefetch.global.L2 [%0];" ::"l"(lrow));
            asm volatile("prefetch.global.L2 [%0];" ::"l"(lrow + 16));
            const double* vrow = v - lane + (long long)pjN * Bp;          // the right-hand-side rows the terms read
            asm volatile("prefetch.global.L2 [%0];" ::"l"(vrow));
            asm volatile("prefetch.global.L2 [%0];" ::"l"(vrow + 16));
        }
        if (haveN && lane == 0 && rN.w >= 0) {
            const double* drow = lu + (long long)rN.w * 32;
            asm volatile("prefetch.global.L2 [%0];" ::"l"(drow));
            asm volatile("prefetch.global.L2 [%0];" ::"l"(drow + 16));
        }
        r = rN; ps = psN; pj = pjN; ex = exN; have = haveN;
        __syncthreads();
    }
}

// ---- one matrix, one CTA per right-hand side (klu.solve / klu.linsolve: batch of one) ----------------------------------------
// A batch of one leaves 31 lanes of every warp of k_klu_solve_lvl idle and pays its level barriers and its dependent
// global loads for one matrix (ACTIVSg2000: 492 levels, 2 ms), and the level schedule itself costs more host time than the
// solve.  Here the SEQUENTIAL algorithm of klu_solve / klu_tsolve (klu.c:593-690) is written down once per pattern as a
// tape of column operations in execution order -- 'N': x_k (/= u_kk), x[rows] -= column * x_k (L columns ascending, U
// columns descending, then the block's F columns; BTF blocks last to first); 'T': x_k = (x_k - column . x[rows]) (/ u_kk)
// (U+F columns ascending, L columns descending; blocks first to last) -- O(nnz) to build, no sorting.  The work vector
// lives in shared memory; the tape is cut into chunks of whole (or split) operations that warps 1..7 stage into a
// double buffer (row index + factor value of the first matrix of the handle + reciprocal pivots: all static, so nothing
// on the dependent chain touches global memory) while warp 0 runs the previous chunk, lanes across the entries of a column.
constexpr int KLU_ONE_ENT = 2048, KLU_ONE_OPS = 512, KLU_ONE_THREADS = 256, KLU_ONE_MAXN = 16384;
struct KluSolveOneD {
    int n = 0, nchunks = 0;
    const int4* ops = nullptr;      // {k, diagonal slot or -1, first entry relative to its chunk, entries}
    const int2* ent = nullptr;      // {row, value slot}
    const int2* chunk = nullptr;    // nchunks + 1: {first op, first entry}
};
struct KluSolveOneH { std::vector<int4> ops; std::vector<int2> ent, chunk; };

static void build_solve_tape(const KluSymbolic& S, const KluNumeric& N, const KluPlan& P, bool trans, KluSolveOneH& H) {
    H.ops.clear(); H.ent.clear(); H.chunk.assign(1, make_int2(0, 0));
    int cops = 0, cent = 0;          // operations / entries of the open chunk
    std::vector<int2> e;
    // an operation longer than a chunk is cut into parts: 'N' parts re-read the finished x_k, 'T' parts subtract their partial
    // dot products one after the other; the pivot goes with the first ('N') / last ('T') part
    auto add_op = [&](int k, int diag) {
        if (e.empty() && diag < 0) return;
        const int total = (int)e.size();
        int done = 0;
        do {
            const int cnt = std::min(total - done, KLU_ONE_ENT);
            if (cops + 1 > KLU_ONE_OPS || cent + cnt > KLU_ONE_ENT) {
                H.chunk.push_back(make_int2((int)H.ops.size(), (int)H.ent.size()));
                cops = 0; cent = 0;
            }
            const bool first = done == 0, last = done + cnt == total;
            H.ops.push_back(make_int4(k, (trans ? last : first) ? diag : -1, cent, cnt));
            H.ent.insert(H.ent.end(), e.begin() + done, e.begin() + done + cnt);
            cops++; cent += cnt; done += cnt;
        } while (done < total);
        e.clear();
    };
    auto Lcol = [&](int k) { for (long long p = N.Lp[k] + 1; p < N.Lp[k + 1]; p++) e.push_back(make_int2(N.Li[p], P.lslot0[k] + (int)(p - N.Lp[k] - 1))); };
    auto Ucol = [&](int k) { for (long long p = N.Up[k]; p < N.Up[k + 1] - 1; p++) e.push_back(make_int2(N.Ui[p], (int)P.cbeg[k] + (int)(p - N.Up[k]))); };
    auto Fcol = [&](int k) { for (long long p = N.Fp[k]; p < N.Fp[k + 1]; p++) e.push_back(make_int2(N.Fi[p], P.fslot0[k] + (int)(p - N.Fp[k]))); };
    if (!trans)
        for (int blk = S.nblocks - 1; blk >= 0; blk--) {
            const int k1 = S.R[blk], k2 = S.R[blk + 1];
            for (int k = k1; k < k2; k++) { Lcol(k); add_op(k, -1); }
            for (int k = k2 - 1; k >= k1; k--) { Ucol(k); add_op(k, P.udiag_slot[k]); }
            for (int k = k1; k < k2; k++) { Fcol(k); add_op(k, -1); }
        }
    else
        for (int blk = 0; blk < S.nblocks; blk++) {
            const int k1 = S.R[blk], k2 = S.R[blk + 1];
            for (int k = k1; k < k2; k++) { Ucol(k); Fcol(k); add_op(k, P.udiag_slot[k]); }
            for (int k = k2 - 1; k >= k1; k--) { Lcol(k); add_op(k, -1); }
        }
    H.chunk.push_back(make_int2((int)H.ops.size(), (int)H.ent.size()));
}

// x / d with the reciprocal taken off the dependent chain: one residual correction of q = x * (1/d)
__device__ __forceinline__ double klu_div_rcp(double x, double d, double rd) {
    const double q = x * rd;
    return fma(fma(-q, d, x), rd, q);
}

template <bool TRANS>
__global__ void __launch_bounds__(KLU_ONE_THREADS) k_klu_solve_one(KluSolveOneD T, int Bp, const double* __restrict__ LU, const double* __restrict__ Rs,
                                                                    const int* __restrict__ pin, const int* __restrict__ pout,
                                                                    double* __restrict__ B, long long ldB) {
    extern __shared__ double smo[];
    const int n = T.n, np = (n + 1) & ~1;
    double* x = smo;                                              // n
    double* sval = x + np;                                        // [2][KLU_ONE_ENT]
    double* sdg = sval + 2 * KLU_ONE_ENT;                         // [2][KLU_ONE_OPS] pivots
    double* srd = sdg + 2 * KLU_ONE_OPS;                          // [2][KLU_ONE_OPS] their reciprocals
    int4* sop = reinterpret_cast<int4*>(srd + 2 * KLU_ONE_OPS);   // [2][KLU_ONE_OPS]
    int* sidx = reinterpret_cast<int*>(sop + 2 * KLU_ONE_OPS);    // [2][KLU_ONE_ENT]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    double* b = B + (long long)blockIdx.x * ldB;
    auto stage = [&](int c, int t, int nt) {
        const int bf = c & 1;
        const int2 c0 = T.chunk[c], c1 = T.chunk[c + 1];
        const int nops = c1.x - c0.x, nent = c1.y - c0.y;
        for (int i = t; i < nent; i += nt) {
            const int2 en = T.ent[c0.y + i];
            sidx[bf * KLU_ONE_ENT + i] = en.x;
            sval[bf * KLU_ONE_ENT + i] = LU[(long long)en.y * 32];
        }
        for (int i = t; i < nops; i += nt) {
            const int4 o = T.ops[c0.x + i];
            const double d = o.y >= 0 ? LU[(long long)o.y * 32] : 1.0;
            sop[bf * KLU_ONE_OPS + i] = o;
            sdg[bf * KLU_ONE_OPS + i] = d;
            srd[bf * KLU_ONE_OPS + i] = 1.0 / d;
        }
    };
    for (int k = tid; k < n; k += KLU_ONE_THREADS) {
        double v = b[pin[k]];
        if (!TRANS) v /= Rs[(long long)k * Bp];
        x[k] = v;
    }
    if (T.nchunks > 0) stage(0, tid, KLU_ONE_THREADS);
    __syncthreads();
    for (int c = 0; c < T.nchunks; c++) {
        if (warp != 0) {
            if (c + 1 < T.nchunks) stage(c + 1, tid - 32, KLU_ONE_THREADS - 32);
        } else {
            const int bf = c & 1;
            const int nops = T.chunk[c + 1].x - T.chunk[c].x;
            const int4* ops = sop + bf * KLU_ONE_OPS;
            const double *dg = sdg + bf * KLU_ONE_OPS, *rd = srd + bf * KLU_ONE_OPS, *val = sval + bf * KLU_ONE_ENT;
            const int* idx = sidx + bf * KLU_ONE_ENT;
            // (fetching the next operation's first 32 entries ahead as well was measured slower: one warp is bound by the
            // number of instructions on the chain, not by the shared-memory latency: 1.12 -> 1.22 ms on ACTIVSg2000)
            int4 o = ops[0];
            double d = dg[0], r = rd[0];
            for (int i = 0; i < nops; i++) {
                int4 on = o; double dn = d, rn = r;
                if (i + 1 < nops) { on = ops[i + 1]; dn = dg[i + 1]; rn = rd[i + 1]; }
                const int k = o.x, e0 = o.z, cnt = o.w;
                if (!TRANS) {
                    double xk = x[k];
                    if (o.y >= 0) {
                        xk = klu_div_rcp(xk, d, r);
                        __syncwarp();                              // every lane has read x_k before lane 0 replaces it
                        if (lane == 0) x[k] = xk;
                    }
                    for (int e = lane; e < cnt; e += 32) {
                        const int row = idx[e0 + e];
                        x[row] = fma(-val[e0 + e], xk, x[row]);
                    }
                } else {
                    double s = 0.0;
                    for (int e = lane; e < cnt; e += 32) s = fma(val[e0 + e], x[idx[e0 + e]], s);
#pragma unroll
                    for (int off = 16; off; off >>= 1)
                        if (cnt > off) s += __shfl_xor_sync(0xffffffffu, s, off);
                    if (lane == 0) {
                        double xk = x[k] - s;
                        if (o.y >= 0) xk = klu_div_rcp(xk, d, r);
                        x[k] = xk;
                    }
                }
                __syncwarp();
                o = on; d = dn; r = rn;
            }
        }
        __syncthreads();
    }
    for (int k = tid; k < n; k += KLU_ONE_THREADS) {
        double v = x[k];
        if (TRANS) v /= Rs[(long long)k * Bp];
        b[pout[k]] = v;
    }
}
// Host replay of the tape (test hook behind b200s_klu_solve_tape_host: the tape is verified without a GPU): the operations in
// order on one right-hand side, exactly as warp 0 of k_klu_solve_one applies them (the sums of a 'T' operation in index order).
int klu_solve_tape_host(const KluSymbolic& S, const KluNumeric& N, const KluPlan& P, int trans, const double* slots, double* B,
                        long long nrhs, long long ldB) {
    KluSolveOneH H;
    build_solve_tape(S, N, P, trans != 0, H);
    const int n = P.n;
    std::vector<double> x((size_t)std::max(n, 1));
    const int* pin = trans ? S.Q.data() : N.Pnum.data();
    const int* pout = trans ? N.Pnum.data() : S.Q.data();
    for (long long r = 0; r < nrhs; r++) {
        double* b = B + r * ldB;
        for (int k = 0; k < n; k++) x[k] = trans ? b[pin[k]] : b[pin[k]] / N.Rs[k];
        for (size_t c = 0; c + 1 < H.chunk.size(); c++) {
            if (H.chunk[c + 1].x - H.chunk[c].x > KLU_ONE_OPS || H.chunk[c + 1].y - H.chunk[c].y > KLU_ONE_ENT) return ST_INVALID;
            for (int q = H.chunk[c].x; q < H.chunk[c + 1].x; q++) {
                const int4 o = H.ops[q];
                const int2* e = H.ent.data() + H.chunk[c].y + o.z;
                if (H.chunk[c].y + o.z + o.w > H.chunk[c + 1].y) return ST_INVALID;
                if (!trans) {
                    if (o.y >= 0) x[o.x] /= slots[o.y];
                    for (int i = 0; i < o.w; i++) x[e[i].x] -= slots[e[i].y] * x[o.x];
                } else {
                    double s = 0.0;
                    for (int i = 0; i < o.w; i++) s += slots[e[i].y] * x[e[i].x];
                    x[o.x] -= s;
                    if (o.y >= 0) x[o.x] /= slots[o.y];
                }
            }
        }
        for (int k = 0; k < n; k++) b[pout[k]] = trans ? x[k] / N.Rs[k] : x[k];
    }
    return ST_OK;
}

static size_t klu_solve_one_smem(int n) {
    return (size_t)((n + 1) & ~1) * 8 + 2 * (size_t)KLU_ONE_ENT * 12 + 2 * (size_t)KLU_ONE_OPS * 32;
}

__global__ void k_klu_gather_slots(const double* __restrict__ LU, int Bp, int b, long long nslots, double* __restrict__ out) {
    for (long long v = blockIdx.x * (long long)blockDim.x + threadIdx.x; v < nslots; v += (long long)gridDim.x * blockDim.x)
        out[v] = LU[(long long)(b >> 5) * nslots * 32 + v * 32 + (b & 31)];
}

class KluDevice {
public:
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev[7] = {};
    KluPlanD PD{};
    KluWaveD WD{};
    KluSolveLvlD SL[2] = {};          // [0] = 'N', [1] = 'T'
    KluSolveOneD SO[2] = {};          // the tapes of the one-matrix solve (k_klu_solve_one)
    bool solve_one_ok = true;         // B200S_KLU_SOLVE_ONE=0: a batch of one runs through the level kernel as well
    const int *d_Pnum = nullptr, *d_Q = nullptr;
    bool use_wave = false;
    long long lu_slots = 0;
    std::vector<void*> owned;
    int *d_slot_src = nullptr, *d_slot_row = nullptr, *d_rowent = nullptr, *d_status = nullptr;
    const int *d_dense_meta = nullptr, *d_dense_slot = nullptr;
    double* dD = nullptr;
    int spine_nd = 0, ndmap = 0, ndp = 0;
    size_t dense_smem = 0;
    std::vector<int> h_wave_col0;
    // pipelined host-buffer path (refactor_begin / refactor_end)
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t ev_h2d[2] = {}, ev_free[2] = {}, ev_done[2] = {};
    double* dAp[2] = {nullptr, nullptr};
    long long capAp[2] = {0, 0}, pend_batch[2] = {0, 0};
    int* h_status_pinned[2] = {nullptr, nullptr};
    std::vector<int> h_status[2];
    bool buf_used[2] = {false, false};
    int npending = 0, next_buf = 0;
    long long* ddbg = nullptr;     // optional phase timers of the wave kernel (B200S_KLU_DEBUG=1)
    // early columns (k_klu_early): per level the [short | long] column ranges in d_ecols
    const KluEarlyCol* d_ecols = nullptr;
    const int4* d_eupd = nullptr;
    const unsigned short* d_edest = nullptr;
    std::vector<int> h_elevel_ptr;
    long long* d_rowptr = nullptr;
    long long nslots = 0, nnzA = 0;
    int n = 0;
    // batch buffers
    int Bp = 0, batch = 0;
    double *dA = nullptr, *dAxt = nullptr, *dRs = nullptr, *dLU = nullptr, *dX = nullptr, *dB = nullptr;
    long long capA = 0, capX = 0, capB = 0;
    double ms_h2d = 0, ms_refactor = 0, ms_solve = 0, ms_kernel = 0, ms_dense = 0;
    long long launches = 0;

    ~KluDevice() {
        cudaSetDevice(device);
        if (copy_stream) cudaStreamSynchronize(copy_stream);
        if (stream) cudaStreamSynchronize(stream);      // blocks go back to the caching allocator
        for (void* p : owned) pool_free(p);
        pool_free(dA); pool_free(dAxt); pool_free(dRs); pool_free(dLU); pool_free(dX); pool_free(dB); pool_free(d_status); pool_free(dD);
        for (auto& e : ev) if (e) cudaEventDestroy(e);
        for (int q = 0; q < 2; q++) {
            pool_free(dAp[q]);
            if (h_status_pinned[q]) cudaFreeHost(h_status_pinned[q]);
            if (ev_h2d[q]) cudaEventDestroy(ev_h2d[q]);
            if (ev_free[q]) cudaEventDestroy(ev_free[q]);
            if (ev_done[q]) cudaEventDestroy(ev_done[q]);
        }
        if (copy_stream) cudaStreamDestroy(copy_stream);
        if (stream) cudaStreamDestroy(stream);
    }
    template <class T> int up(const T** dst, const std::vector<T>& src) {
        T* p = nullptr;
        CUDA_TRY(pool_malloc((void**)&p, std::max<size_t>(src.size(), 1) * sizeof(T)));
        owned.push_back(p);
        if (!src.empty()) CUDA_TRY(cudaMemcpy(p, src.data(), src.size() * sizeof(T), cudaMemcpyHostToDevice));
        *dst = p;
        return ST_OK;
    }
    int init(const KluPlan& P, const KluNumeric& N, const KluSymbolic& S);
    int ensure_solve_levels(int tr);
    int ensure_solve_tape(int tr);
    const KluPlan* hP = nullptr; const KluNumeric* hN = nullptr; const KluSymbolic* hS = nullptr;   // owned by the numeric object
    int ensure_batch(int b);
    int enqueue_refactor(const double* dv, long long ldv, cudaEvent_t after_transpose);
    int refactor(const double* vals, bool on_device, long long batch_, long long ldv, int* status_host);
    int init_refactor(const KluPlan& P);
    int load_host_factor(const double* slots_host, const double* rs_host);
    bool refactor_ready = false;
    int refactor_begin(const double* vals, long long batch_, long long ldv);
    int refactor_end(int* status_host);
    int solve(int trans, double* B, long long nrhs, long long ldB, long long batch_, bool on_device);
};

int KluDevice::init(const KluPlan& P, const KluNumeric& N, const KluSymbolic& S) {
    CUDA_TRY(cudaSetDevice(device));
    CUDA_TRY(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
    for (auto& e : ev) CUDA_TRY(cudaEventCreate(&e));
    n = P.n; nslots = P.nslots; nnzA = P.nnzA;
    int rc;
    PD.n = P.n; PD.nlevels = P.nlevels; PD.gstride = (long long)P.nslots * 32;
    std::vector<long long> cbeg(P.cbeg.begin(), P.cbeg.end()), rowptr(P.rowptr.begin(), P.rowptr.end());
    if ((rc = up(&PD.udiag_slot, P.udiag_slot))) return rc;
    if ((rc = up(&PD.lslot0, P.lslot0))) return rc;
    if ((rc = up(&PD.cbeg, cbeg))) return rc;
    lu_slots = P.lu_slots;
    const int* tmp_i; const long long* tmp_l;
    if ((rc = up(&tmp_i, P.slot_src))) return rc; d_slot_src = (int*)tmp_i;
    if ((rc = up(&tmp_i, P.slot_row))) return rc; d_slot_row = (int*)tmp_i;
    if ((rc = up(&tmp_i, P.rowent))) return rc; d_rowent = (int*)tmp_i;
    if ((rc = up(&tmp_l, rowptr))) return rc; d_rowptr = (long long*)tmp_l;
    if ((rc = up(&d_Pnum, N.Pnum))) return rc;
    if ((rc = up(&d_Q, S.Q))) return rc;
    hP = &P; hN = &N; hS = &S;       // the solve schedules are built at the first solve of each kind (ensure_solve_levels)
    { const char* e1 = getenv("B200S_KLU_SOLVE_ONE"); solve_one_ok = !(e1 && e1[0] == '0'); }
    if (P.have_refactor) return init_refactor(P);
    return ST_OK;
}

// The tables of the refactorization kernels (update lists or wave schedule, early columns, dense block).  Separate from init():
// klu.numeric() takes the values of its factor from the host pivot search, so a caller that never refactors a batch (one
// matrix, klu.linsolve) never pays for building and uploading these tables.
int KluDevice::init_refactor(const KluPlan& P) {
    CUDA_TRY(cudaSetDevice(device));
    if (refactor_ready) return ST_OK;
    int rc;
    if (getenv("B200S_KLU_DEBUG")) { CUDA_TRY(pool_malloc((void**)&ddbg, (8 + 2 * P.wave_col0.size()) * sizeof(long long))); owned.push_back(ddbg); h_wave_col0 = P.wave_col0; cudaMemset(ddbg, 0, (8 + 2 * P.wave_col0.size()) * sizeof(long long)); }
    std::vector<long long> updp(P.upd_ptr.begin(), P.upd_ptr.end()), updd(P.upd_dest.begin(), P.upd_dest.end());
    use_wave = P.wave_ok && P.max_col_len <= KLU_WAVE_ROWS;
    if (!use_wave) {          // update lists and level schedule: only the level-schedule kernel reads them
        if ((rc = up(&PD.level_ptr, P.level_ptr))) return rc;
        if ((rc = up(&PD.level_cols, P.level_cols))) return rc;
        if ((rc = up(&PD.upd_uslot, P.upd_uslot))) return rc;
        if ((rc = up(&PD.upd_lslot, P.upd_lslot))) return rc;
        if ((rc = up(&PD.upd_cnt, P.upd_cnt))) return rc;
        if ((rc = up(&PD.dest, P.dest))) return rc;
        if ((rc = up(&PD.upd_ptr, updp))) return rc;
        if ((rc = up(&PD.upd_dest, updd))) return rc;
    }
    {
        WD.nwaves = (int)P.wave_col0.size() - 1;
        if ((rc = up(&WD.wave_col0, P.wave_col0))) return rc;
        if ((rc = up(&WD.col_roff, P.col_roff))) return rc;
        std::vector<long long> wbp(P.wbatch_ptr.begin(), P.wbatch_ptr.end()), wlp(P.wblob_ptr.begin(), P.wblob_ptr.end());
        if ((rc = up(&WD.wbatch_ptr, wbp))) return rc;
        if ((rc = up(&WD.wblob_ptr, wlp))) return rc;
        if ((rc = up(&WD.wave_rowsrc, P.wave_rowsrc))) return rc;
        {
            std::vector<int> bsp(P.bseg_ptr.begin(), P.bseg_ptr.end());
            std::vector<int2> sgd(P.seg_src.size());
            for (size_t q = 0; q < sgd.size(); q++)
                sgd[q] = make_int2(P.lslot0[P.seg_src[q]] + P.seg_off[q], P.seg_row[q] | (P.seg_cnt[q] << 8));
            bsp.push_back(bsp.back());       // the producer reads bseg_ptr[g + 2] one batch ahead
            if ((rc = up(&WD.bseg_ptr, bsp))) return rc;
            if ((rc = up(&WD.segd, sgd))) return rc;
        }
        if ((rc = up(&WD.ne_cols, P.ne_cols))) return rc;
        if ((rc = up(&WD.wave_rows, P.wave_rows))) return rc;
        if ((rc = up(&WD.wrun_ptr, P.wrun_ptr))) return rc;
        {
            std::vector<int4> runs(P.wrun_slot.size());
            for (size_t q = 0; q < runs.size(); q++) runs[q] = make_int4(P.wrun_slot[q], P.wrun_row[q], P.wrun_cnt[q], 0);
            if ((rc = up(&WD.wrun, runs))) return rc;
        }
        if (use_wave && !P.ecols.empty()) {
            // compact tables of the early columns, in launch order
            std::vector<KluEarlyCol> ec(P.ecols.size());
            std::vector<int4> eu;
            std::vector<unsigned short> ed;
            for (size_t e = 0; e < P.ecols.size(); e++) {
                const int k = P.ecols[e];
                const long long cb = P.cbeg[k];
                ec[e].cb = (int)cb; ec[e].len = (int)(P.cbeg[k + 1] - cb);
                ec[e].diag = P.udiag_slot[k] - (int)cb; ec[e].l0 = P.lslot0[k] - (int)cb;
                ec[e].upd0 = (int)eu.size(); ec[e].nupd = (int)(P.upd_ptr[k + 1] - P.upd_ptr[k]);
                for (long long u = P.upd_ptr[k]; u < P.upd_ptr[k + 1]; u++) {
                    eu.push_back(make_int4(P.upd_uslot[u] - (int)cb, P.upd_lslot[u], P.upd_cnt[u], (int)ed.size()));
                    for (int t = 0; t < P.upd_cnt[u]; t++) ed.push_back((unsigned short)(P.dest[P.upd_dest[u] + t] - cb));
                }
            }
            if ((rc = up(&d_ecols, ec))) return rc;
            if ((rc = up(&d_eupd, eu))) return rc;
            if ((rc = up(&d_edest, ed))) return rc;
            h_elevel_ptr = P.elevel_ptr;
        }
        static_assert(sizeof(unsigned) == sizeof(uint32_t), "plan tables are uploaded as they are");
        if ((rc = up(&WD.bentry, P.bentry))) return rc;       // the largest table (11.5 MB on ACTIVSg2000): no staging copy
        if ((rc = up(&WD.wblob, P.wblob))) return rc;
        WD.spine0 = use_wave ? P.spine0 : P.n;
        spine_nd = use_wave ? P.spine_nd : 0;
        if ((rc = up(&d_dense_meta, P.dense_meta))) return rc;
        if ((rc = up(&d_dense_slot, P.dense_slot))) return rc;
        ndmap = (int)P.dense_slot.size();
        ndp = (ndmap + 31) & ~31;
        dense_smem = (size_t)spine_nd * (spine_nd + 4) * sizeof(double) + (size_t)spine_nd * KLU_DENSE_META * sizeof(int);
        if (spine_nd > 0)
            CUDA_TRY(cudaFuncSetAttribute(k_klu_dense_lu, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          (int)dense_smem));
        if (use_wave)
        {
            CUDA_TRY(cudaFuncSetAttribute(k_klu_refactor_wave, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)KLU_WAVE_SMEM));
        }
    }
    // a batch buffer allocated before the tables existed has no storage for the dense trailing block yet
    if (spine_nd > 0 && Bp > 0 && !dD) CUDA_TRY(pool_malloc((void**)&dD, (size_t)ndp * Bp * sizeof(double)));
    refactor_ready = true;
    return ST_OK;
}

// lane-replicated load of one factor computed on the host (the pivot search of klu.numeric): slot v of every matrix of the
// first group = h[v], row scales likewise.  The device then serves solve / extract exactly as after a refactorization.
__global__ void k_klu_load_one(const double* __restrict__ h, long long nslots, const double* __restrict__ hrs, int n, int Bp,
                               double* __restrict__ LU, double* __restrict__ Rs) {
    const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    const long long v = t >> 5;
    const int lane = (int)(t & 31);
    if (v < nslots) LU[v * 32 + lane] = h[v];
    if (v < n) Rs[v * Bp + lane] = hrs[v];
}

int KluDevice::load_host_factor(const double* slots_host, const double* rs_host) {
    CUDA_TRY(cudaSetDevice(device));
    if (n == 0) return ST_OK;
    if (npending > 0) return ST_INVALID;
    int rc = ensure_batch(1);
    if (rc) return rc;
    double* tmp = nullptr;
    CUDA_TRY(pool_malloc((void**)&tmp, (size_t)(nslots + n) * sizeof(double)));
    CUDA_TRY(cudaMemcpyAsync(tmp, slots_host, (size_t)nslots * sizeof(double), cudaMemcpyHostToDevice, stream));
    CUDA_TRY(cudaMemcpyAsync(tmp + nslots, rs_host, (size_t)n * sizeof(double), cudaMemcpyHostToDevice, stream));
    CUDA_TRY(cudaMemsetAsync(d_status, 0, Bp * sizeof(int), stream));
    const long long threads = std::max<long long>(nslots, n) * 32;
    k_klu_load_one<<<(unsigned)((threads + 255) / 256), 256, 0, stream>>>(tmp, nslots, tmp + nslots, n, Bp, dLU, dRs);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaStreamSynchronize(stream));
    pool_free(tmp);
    batch = 1;
    return ST_OK;
}

int KluDevice::ensure_solve_levels(int tr) {
    if (SL[tr].n == n && SL[tr].lvl_ptr) return ST_OK;
    int rc;
    KluSolveLvlH H;
    build_solve_levels(*hS, *hN, *hP, tr != 0, H);
    SL[tr].n = n; SL[tr].nlev = H.nlev;
    if ((rc = up(&SL[tr].lvl_ptr, H.lvl_ptr))) return rc;
    if ((rc = up(&SL[tr].rec, H.rec))) return rc;
    if ((rc = up(&SL[tr].tslot, H.tslot))) return rc;
    if ((rc = up(&SL[tr].tsrc, H.tsrc))) return rc;
    if ((rc = up(&SL[tr].extra, H.extra))) return rc;
    if ((rc = up(&SL[tr].mode, H.mode))) return rc;
    return ST_OK;
}

int KluDevice::ensure_solve_tape(int tr) {
    if (SO[tr].n == n && SO[tr].chunk) return ST_OK;
    int rc;
    KluSolveOneH H;
    build_solve_tape(*hS, *hN, *hP, tr != 0, H);
    SO[tr].n = n; SO[tr].nchunks = (int)H.chunk.size() - 1;
    if ((rc = up(&SO[tr].ops, H.ops))) return rc;
    if ((rc = up(&SO[tr].ent, H.ent))) return rc;
    if ((rc = up(&SO[tr].chunk, H.chunk))) return rc;
    return ST_OK;
}

int KluDevice::ensure_batch(int b) {
    const int bp = (b + 31) & ~31;
    if (bp <= Bp) { batch = b; return ST_OK; }
    pool_free(dAxt); pool_free(dRs); pool_free(dLU); pool_free(d_status); pool_free(dD);
    dAxt = dRs = dLU = dD = nullptr; d_status = nullptr; Bp = 0;
    CUDA_TRY(pool_malloc((void**)&dAxt, std::max<long long>(nnzA, 1) * bp * sizeof(double)));
    CUDA_TRY(pool_malloc((void**)&dRs, std::max<long long>(n, 1) * (long long)bp * sizeof(double)));
    CUDA_TRY(pool_malloc((void**)&dLU, std::max<long long>(nslots, 1) * bp * sizeof(double)));
    if (spine_nd > 0) CUDA_TRY(pool_malloc((void**)&dD, (size_t)ndp * bp * sizeof(double)));
    CUDA_TRY(pool_malloc((void**)&d_status, bp * sizeof(int)));
    Bp = bp; batch = b;
    return ST_OK;
}

// the kernels of one refactorization, enqueued on `stream`; dv = device copy of the caller's [batch][ldv] values
int KluDevice::enqueue_refactor(const double* dv, long long ldv, cudaEvent_t after_transpose) {
    CUDA_TRY(cudaMemsetAsync(d_status, 0, Bp * sizeof(int), stream));
    {
        dim3 grid((unsigned)((nnzA + 31) / 32), (unsigned)(Bp / 32)), block(32, 8);
        k_klu_transpose<<<grid, block, 0, stream>>>(dv, ldv, nnzA, batch, Bp, dAxt);
    }
    if (after_transpose) CUDA_TRY(cudaEventRecord(after_transpose, stream));       // the caller's value buffer is free again
    k_klu_rowscale<<<148 * 8, 256, 0, stream>>>(d_rowptr, d_rowent, n, Bp, dAxt, dRs, use_wave ? 1 : 0);
    if (use_wave) {
        // only the off-diagonal-block entries F need a separate scatter; L/U columns are gathered inside the kernel
        if (nslots > lu_slots)
            k_klu_scatter<<<148 * 8, 256, 0, stream>>>(d_slot_src, d_slot_row, lu_slots, nslots, Bp, dAxt, dRs, dLU, 1, nslots * 32);
        CUDA_TRY(cudaEventRecord(ev[4], stream));
        int early_launches = 0;
        for (size_t l = 0; l + 1 < h_elevel_ptr.size(); l++) {          // early columns, level by level
            const int c0 = h_elevel_ptr[l], c1 = h_elevel_ptr[l + 1];
            if (c1 == c0) continue;
            const dim3 grid((unsigned)((c1 - c0 + 7) / 8), (unsigned)(Bp / 32));
            k_klu_early<KLU_EARLY_MAXLEN, 8><<<grid, 8 * 32, KLU_EARLY_MAXLEN * 8 * 256, stream>>>(d_ecols, c0, c1, d_eupd, d_edest, d_slot_src, nslots * 32, Bp, dAxt, dLU, d_status);
            early_launches++;
        }
        if (WD.nwaves > 0)
            k_klu_refactor_wave<<<Bp / 32, (KLU_WAVE_WARPS + 1) * 32, KLU_WAVE_SMEM, stream>>>(PD, WD, Bp, dAxt, dLU, d_status, ddbg);
        CUDA_TRY(cudaEventRecord(ev[6], stream));
        launches = 2 + early_launches + (WD.nwaves > 0 ? 1 : 0) + (nslots > lu_slots ? 1 : 0) + (spine_nd > 0 ? 3 : 0);
        if (spine_nd > 0) {
            const dim3 tg(ndp / 32, Bp / 32);
            k_klu_dense_pack<<<tg, 256, 0, stream>>>(d_dense_slot, ndmap, ndp, nslots * 32, dLU, dD, 0);
            k_klu_dense_lu<<<batch, KLU_DENSE_THREADS, dense_smem, stream>>>(spine_nd, d_dense_meta, ndp, batch, dD, d_status);
            k_klu_dense_pack<<<tg, 256, 0, stream>>>(d_dense_slot, ndmap, ndp, nslots * 32, dLU, dD, 1);
        }
        CUDA_TRY(cudaEventRecord(ev[5], stream));
    } else {
        k_klu_scatter<<<148 * 16, 256, 0, stream>>>(d_slot_src, d_slot_row, 0, nslots, Bp, dAxt, dRs, dLU, 0, nslots * 32);
        CUDA_TRY(cudaEventRecord(ev[4], stream));
        k_klu_refactor<<<Bp / 32, KLU_WARPS * 32, 0, stream>>>(PD, nslots * 32, dLU, d_status);
        CUDA_TRY(cudaEventRecord(ev[6], stream));
        CUDA_TRY(cudaEventRecord(ev[5], stream));
        launches = 4;
    }
    CUDA_TRY(cudaGetLastError());
    return ST_OK;
}

int KluDevice::refactor(const double* vals, bool on_device, long long batch_, long long ldv, int* status_host) {
    CUDA_TRY(cudaSetDevice(device));
    if (batch_ <= 0 || n == 0) return ST_OK;
    // batches begun with refactor_begin still own the staging buffers and the factor storage: the blocking call may
    // neither grow nor free them (the caching allocator would hand live blocks to another handle)
    if (npending > 0) { set_last_error("refactor_batch: batches begun with refactor_batch_begin are still in flight, call refactor_batch_end first"); return ST_INVALID; }
    if (!refactor_ready) { set_last_error("refactorization tables have not been built"); return ST_INVALID; }
    int rc = ensure_batch((int)batch_);
    if (rc) return rc;
    CUDA_TRY(cudaEventRecord(ev[0], stream));
    const double* dv = vals;
    if (!on_device) {
        const long long need = (batch_ - 1) * ldv + nnzA;
        if (need > capA) {
            pool_free(dA); dA = nullptr; capA = 0;
            CUDA_TRY(pool_malloc((void**)&dA, need * sizeof(double)));
            capA = need;
        }
        CUDA_TRY(cudaMemcpyAsync(dA, vals, need * sizeof(double), cudaMemcpyHostToDevice, stream));
        dv = dA;
    }
    CUDA_TRY(cudaEventRecord(ev[1], stream));
    if ((rc = enqueue_refactor(dv, ldv, nullptr))) return rc;
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaEventRecord(ev[2], stream));
    std::vector<int> st;
    if (status_host) {
        st.resize(Bp);
        CUDA_TRY(cudaMemcpyAsync(st.data(), d_status, Bp * sizeof(int), cudaMemcpyDeviceToHost, stream));
    }
    CUDA_TRY(cudaStreamSynchronize(stream));
    if (status_host) for (long long b = 0; b < batch_; b++) status_host[b] = st[b];
    if (ddbg) {
        long long h[8] = {0};
        cudaMemcpy(h, ddbg, sizeof h, cudaMemcpyDeviceToHost);
        fprintf(stderr, "[klu wave kernel, CTA 0] cycles: gather+prologue %lld  staged-updates %lld  in-wave rounds %lld  (rounds %lld) | warp0 waits for staged data %lld (first batch of a wave %lld)\n", h[0], h[1], h[2], h[3], h[4], h[5]);
        if (const char* f = getenv("B200S_KLU_DEBUG_FILE")) {
            std::vector<long long> hw(8 + 2 * (size_t)WD.nwaves);
            cudaMemcpy(hw.data(), ddbg, hw.size() * sizeof(long long), cudaMemcpyDeviceToHost);
            if (FILE* fp = fopen(f, "w")) {
                for (int w = 0; w < WD.nwaves; w++) fprintf(fp, "%d %d %lld %lld\n", w, h_wave_col0[w], hw[8 + 2 * w], hw[9 + 2 * w]);
                fclose(fp);
            }
        }
    }
    float ms;
    cudaEventElapsedTime(&ms, ev[0], ev[1]); ms_h2d = ms;
    cudaEventElapsedTime(&ms, ev[1], ev[2]); ms_refactor = ms;
    cudaEventElapsedTime(&ms, ev[4], ev[6]); ms_kernel = ms;
    cudaEventElapsedTime(&ms, ev[6], ev[5]); ms_dense = ms;
    return ST_OK;
}

// Pipelined host-buffer refactorization: begin() enqueues the H2D copy on a copy stream and the kernels on the compute
// stream and returns; end() waits for the oldest batch in flight.  With two batches in flight the upload of batch i+1
// overlaps the kernels of batch i (two device staging buffers; a staging buffer is free again as soon as the transpose
// kernel has consumed it).
int KluDevice::refactor_begin(const double* vals, long long batch_, long long ldv) {
    CUDA_TRY(cudaSetDevice(device));
    if (batch_ <= 0 || n == 0) return ST_OK;
    if (npending >= 2) { set_last_error("refactor_batch_begin: two batches already in flight, call refactor_batch_end first"); return ST_INVALID; }
    if (!refactor_ready) { set_last_error("refactorization tables have not been built"); return ST_INVALID; }
    const int bp = ((int)batch_ + 31) & ~31;
    if (npending > 0 && bp > Bp) { set_last_error("refactor_batch_begin: a larger batch needs the batches in flight to finish first"); return ST_INVALID; }
    int rc = ensure_batch((int)batch_);
    if (rc) return rc;
    if (!copy_stream) {
        CUDA_TRY(cudaStreamCreateWithFlags(&copy_stream, cudaStreamNonBlocking));
        for (int q = 0; q < 2; q++) {
            CUDA_TRY(cudaEventCreateWithFlags(&ev_h2d[q], cudaEventDisableTiming));
            CUDA_TRY(cudaEventCreateWithFlags(&ev_free[q], cudaEventDisableTiming));
            CUDA_TRY(cudaEventCreateWithFlags(&ev_done[q], cudaEventDisableTiming));
        }
    }
    const int q = next_buf;
    const long long need = (batch_ - 1) * ldv + nnzA;
    if (need > capAp[q]) {
        if (buf_used[q]) CUDA_TRY(cudaEventSynchronize(ev_free[q]));
        pool_free(dAp[q]); dAp[q] = nullptr; capAp[q] = 0;
        CUDA_TRY(pool_malloc((void**)&dAp[q], need * sizeof(double)));
        capAp[q] = need;
    }
    if ((long long)h_status[q].size() < Bp) {
        if (h_status_pinned[q]) cudaFreeHost(h_status_pinned[q]);
        CUDA_TRY(cudaMallocHost((void**)&h_status_pinned[q], Bp * sizeof(int)));
        h_status[q].resize(Bp);
    }
    if (buf_used[q]) CUDA_TRY(cudaStreamWaitEvent(copy_stream, ev_free[q], 0));
    CUDA_TRY(cudaMemcpyAsync(dAp[q], vals, need * sizeof(double), cudaMemcpyHostToDevice, copy_stream));
    CUDA_TRY(cudaEventRecord(ev_h2d[q], copy_stream));
    CUDA_TRY(cudaStreamWaitEvent(stream, ev_h2d[q], 0));
    if ((rc = enqueue_refactor(dAp[q], ldv, ev_free[q]))) return rc;
    CUDA_TRY(cudaMemcpyAsync(h_status_pinned[q], d_status, Bp * sizeof(int), cudaMemcpyDeviceToHost, stream));
    CUDA_TRY(cudaEventRecord(ev_done[q], stream));
    buf_used[q] = true;
    pend_batch[q] = batch_;
    npending++;
    next_buf ^= 1;
    return ST_OK;
}

int KluDevice::refactor_end(int* status_host) {
    CUDA_TRY(cudaSetDevice(device));
    if (npending == 0) return ST_OK;
    const int q = (npending == 2) ? next_buf : (next_buf ^ 1);     // the oldest batch in flight
    CUDA_TRY(cudaEventSynchronize(ev_done[q]));
    if (status_host) for (long long b = 0; b < pend_batch[q]; b++) status_host[b] = h_status_pinned[q][b];
    npending--;
    return ST_OK;
}

int KluDevice::solve(int trans, double* B, long long nrhs, long long ldB, long long batch_, bool on_device) {
    CUDA_TRY(cudaSetDevice(device));
    if (batch_ <= 0 || n == 0 || nrhs <= 0) return ST_OK;
    if (batch_ > batch) { set_last_error("solve_batch: batch larger than the last refactored batch"); return ST_INVALID; }
    const long long bstride = ldB * nrhs;
    const bool one = batch_ == 1 && solve_one_ok && n <= KLU_ONE_MAXN;
    const long long needX = one ? 0 : 2ll * n * Bp * nrhs;      // V = [Y; Z] per right-hand side
    if (needX > capX) {
        pool_free(dX); dX = nullptr; capX = 0;
        CUDA_TRY(pool_malloc((void**)&dX, needX * sizeof(double)));
        capX = needX;
    }
    CUDA_TRY(cudaEventRecord(ev[0], stream));
    double* db = B;
    const long long totalB = bstride * batch_;
    // the caller's buffer ends with the n entries of the last column: (ldB - n) doubles fewer than batch * nrhs * ldB
    const long long hostB = totalB - (ldB - n);
    if (!on_device) {
        if (totalB > capB) {
            pool_free(dB); dB = nullptr; capB = 0;
            CUDA_TRY(pool_malloc((void**)&dB, totalB * sizeof(double)));
            capB = totalB;
        }
        CUDA_TRY(cudaMemcpyAsync(dB, B, hostB * sizeof(double), cudaMemcpyHostToDevice, stream));
        db = dB;
    }
    if (one) {
        // one matrix: the sequential tape, one CTA per right-hand side, permutations and row scaling inside the kernel
        { const int rct = ensure_solve_tape(trans ? 1 : 0); if (rct) return rct; }
        const size_t sm = klu_solve_one_smem(n);
        if (trans) {
            CUDA_TRY(cudaFuncSetAttribute(k_klu_solve_one<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
            k_klu_solve_one<true><<<(unsigned)nrhs, KLU_ONE_THREADS, sm, stream>>>(SO[1], Bp, dLU, dRs, d_Q, d_Pnum, db, ldB);
        } else {
            CUDA_TRY(cudaFuncSetAttribute(k_klu_solve_one<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
            k_klu_solve_one<false><<<(unsigned)nrhs, KLU_ONE_THREADS, sm, stream>>>(SO[0], Bp, dLU, dRs, d_Pnum, d_Q, db, ldB);
        }
    } else {
        { const int rcl = ensure_solve_levels(trans ? 1 : 0); if (rcl) return rcl; }
        const KluSolveLvlD& L_ = SL[trans ? 1 : 0];
        const dim3 tg((unsigned)((n + 31) / 32), (unsigned)(Bp / 32), (unsigned)nrhs);
        const int nb = (int)batch_;
        // 'N': V = Rs^-1 P b, result scattered through Q;  'T': V = Q' b, result P' Rs^-1 z
        k_klu_solve_load<<<tg, 256, 0, stream>>>(n, nb, Bp, trans ? d_Q : d_Pnum, trans ? nullptr : dRs, db, ldB, bstride, dX);
        const size_t sm = (size_t)KLU_SOLVE_MAXPARTS * 32 * sizeof(double) + (size_t)(2 * L_.nlev + 4) * sizeof(int);
        if (sm > 220 * 1024) { set_last_error("klu solve: too many levels"); return ST_TOO_LARGE; }
        if (sm > 48 * 1024) CUDA_TRY(cudaFuncSetAttribute(k_klu_solve_lvl, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
        k_klu_solve_lvl<<<dim3((unsigned)(Bp / 32), (unsigned)nrhs), KLU_SOLVE_WARPS * 32, sm, stream>>>(L_, Bp, nslots * 32, dLU, dX);
        k_klu_solve_store<<<tg, 256, 0, stream>>>(n, nb, Bp, trans ? d_Pnum : d_Q, trans ? dRs : nullptr, db, ldB, bstride, dX);
    }
    CUDA_TRY(cudaGetLastError());
    if (!on_device) CUDA_TRY(cudaMemcpyAsync(B, dB, hostB * sizeof(double), cudaMemcpyDeviceToHost, stream));
    CUDA_TRY(cudaEventRecord(ev[1], stream));
    CUDA_TRY(cudaStreamSynchronize(stream));
    float ms;
    cudaEventElapsedTime(&ms, ev[0], ev[1]); ms_solve = ms;
    return ST_OK;
}

// ---- entry points used by klu_capi.cu ---------------------------------------------------------------
KluDevice* klu_device_create(const KluPlan& P, const KluNumeric& N, const KluSymbolic& S, int device, int* status) {
    if (device_count() <= 0) { *status = ST_NO_DEVICE; set_last_error("no CUDA device available"); return nullptr; }
    KluDevice* d = new KluDevice();
    d->device = device;
    *status = d->init(P, N, S);
    if (*status != ST_OK) { delete d; return nullptr; }
    return d;
}
void klu_device_destroy(KluDevice* d) { delete d; }
int klu_device_init_refactor(KluDevice* d, const KluPlan& P) { return d->init_refactor(P); }
int klu_device_load_host_factor(KluDevice* d, const double* slots_host, const double* rs_host) { return d->load_host_factor(slots_host, rs_host); }
int klu_device_refactor(KluDevice* d, const double* vals, bool on_device, long long batch, long long ldv, int* status) {
    return d->refactor(vals, on_device, batch, ldv, status);
}
int klu_device_refactor_begin(KluDevice* d, const double* vals, long long batch, long long ldv) { return d->refactor_begin(vals, batch, ldv); }
int klu_device_refactor_end(KluDevice* d, int* status) { return d->refactor_end(status); }
int klu_device_solve(KluDevice* d, int trans, double* B, long long nrhs, long long ldB, long long batch, bool on_device) {
    return d->solve(trans, B, nrhs, ldB, batch, on_device);
}
int klu_device_extract(KluDevice* d, long long b, double* slots_host, double* rs_host) {
    CUDA_TRY(cudaSetDevice(d->device));
    if (b < 0 || b >= d->batch) { set_last_error("extract_batch: matrix index out of range"); return ST_INVALID; }
    double* tmp = nullptr;
    const long long cnt = std::max<long long>(d->nslots, d->n);
    CUDA_TRY(pool_malloc((void**)&tmp, std::max<long long>(cnt, 1) * sizeof(double)));
    k_klu_gather_slots<<<148 * 4, 256, 0, d->stream>>>(d->dLU, d->Bp, (int)b, d->nslots, tmp);
    cudaError_t e = cudaMemcpyAsync(slots_host, tmp, d->nslots * sizeof(double), cudaMemcpyDeviceToHost, d->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(d->stream);
    if (e == cudaSuccess && rs_host) {
        k_klu_gather_slots<<<148, 256, 0, d->stream>>>(d->dRs, d->Bp, (int)b, d->n, tmp);
        e = cudaMemcpyAsync(rs_host, tmp, d->n * sizeof(double), cudaMemcpyDeviceToHost, d->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(d->stream);
    }
    pool_free(tmp);
    CUDA_TRY(e);
    return ST_OK;
}
void klu_device_times(const KluDevice* d, double* h2d, double* refactor, double* solve, double* kernel, double* dense, long long* launches) {
    *h2d = d->ms_h2d; *refactor = d->ms_refactor; *solve = d->ms_solve; *kernel = d->ms_kernel; *dense = d->ms_dense; *launches = d->launches;
}

}  // namespace b200s
