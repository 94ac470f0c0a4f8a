// Dense reduced KKT solver on the device: the B200 counterpart of the reference's `misc.kkt_chol`
// (reference src/python/misc.py:1213-1349, the 'chol' kktsolver of solvers.conelp/coneqp/lp/qp) for componentwise
// inequality cones.  Same algebra as the reference:
//     A' = [Q1 Q2] [R; 0]                       (Householder QR, once per problem: misc.py:1246-1251)
//     K  = [Q1 Q2]' (H + Gs' Gs) [Q1 Q2],  Gs = W^-T G = diag(di) G      (misc.py:1269-1279)
//     K22 = L L'                                (dense Cholesky of order n-p: misc.py:1282)
// and the solve of misc.py:1284-1345.  Q is kept in compact WY form Q = I - V T V' (V n x p, T p x p), so applying it
// costs O(n p) per vector and O(n^2 p) per matrix like the reference's ormqr.  The dense Cholesky of K22 is the
// one-supernode case of the multifrontal engine (chol_gpu.cu: k_panel + k_update DMMA tiles); SYRK and the WY products
// run in k_dgemm (register-tiled FP64).  All device work is on the factor object's stream; no CPU fallback.
#include "../../include/b200sparse.h"
#include "gpu.hpp"
#include "nvtx_range.hpp"
#include "devpool.hpp"
#include <cuda_runtime.h>
#include <algorithm>
#include <cmath>
#include <cstring>
#include <new>
#include <stdexcept>
#include <vector>

using namespace b200s;

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

namespace {

// C (M x N, column-major, ldc) = alpha * op(A) op(B) + beta * C.  op(A)(i,k) = A[i*sai + k*sak], op(B)(k,j) = B[k*sbk + j*sbj]:
// the four transpose combinations are stride choices.  64 x 64 tile per CTA, 256 threads, 4 x 4 register block per thread,
// 16-deep k-tiles staged in shared memory; tiles are loaded with the thread index running along whichever dimension is
// contiguous in memory (coalesced for N and T operands alike).  Products are summed in k order: bit-reproducible.
constexpr int GT = 64, GK = 16;
__global__ void __launch_bounds__(256) k_dgemm(int M, int N, int Kd, double alpha, const double* __restrict__ A, long long sai, long long sak,
                                               const double* __restrict__ B, long long sbk, long long sbj, double beta,
                                               double* __restrict__ Cm, long long ldc) {
    __shared__ double As[GK][GT + 1], Bs[GK][GT + 1];
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int i0 = blockIdx.x * GT, j0 = blockIdx.y * GT;
    double acc[4][4];
#pragma unroll
    for (int a = 0; a < 4; a++)
#pragma unroll
        for (int b = 0; b < 4; b++) acc[a][b] = 0.0;
    const bool a_ifast = sai == 1, b_jfast = sbj == 1;
    for (int k0 = 0; k0 < Kd; k0 += GK) {
#pragma unroll
        for (int q = 0; q < (GT * GK) / 256; q++) {
            const int e = tid + q * 256;
            int i, k;
            if (a_ifast) { i = e & (GT - 1); k = e / GT; } else { k = e & (GK - 1); i = e / GK; }
            As[k][i] = (i0 + i < M && k0 + k < Kd) ? A[(long long)(i0 + i) * sai + (long long)(k0 + k) * sak] : 0.0;
            int j, kk;
            if (b_jfast) { j = e & (GT - 1); kk = e / GT; } else { kk = e & (GK - 1); j = e / GK; }
            Bs[kk][j] = (j0 + j < N && k0 + kk < Kd) ? B[(long long)(k0 + kk) * sbk + (long long)(j0 + j) * sbj] : 0.0;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < GK; k++) {
            double av[4], bv[4];
#pragma unroll
            for (int a = 0; a < 4; a++) av[a] = As[k][tx + 16 * a];
#pragma unroll
            for (int b = 0; b < 4; b++) bv[b] = Bs[k][ty + 16 * b];
#pragma unroll
            for (int a = 0; a < 4; a++)
#pragma unroll
                for (int b = 0; b < 4; b++) acc[a][b] = fma(av[a], bv[b], acc[a][b]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int b = 0; b < 4; b++) {
        const int j = j0 + ty + 16 * b;
        if (j >= N) continue;
#pragma unroll
        for (int a = 0; a < 4; a++) {
            const int i = i0 + tx + 16 * a;
            if (i >= M) continue;
            double* c = Cm + (long long)j * ldc + i;
            *c = beta == 0.0 ? alpha * acc[a][b] : fma(alpha, acc[a][b], beta * *c);
        }
    }
}

// y (rows) = alpha * A x + beta * y, A rows x cols column-major (lda): one thread per row, coalesced over rows
__global__ void k_dgemv_n(long long rows, long long cols, double alpha, const double* __restrict__ A, long long lda,
                          const double* __restrict__ x, double beta, double* __restrict__ y) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= rows) return;
    double s = 0.0;
    for (long long j = 0; j < cols; j++) s = fma(A[j * lda + i], x[j], s);
    y[i] = beta == 0.0 ? alpha * s : fma(alpha, s, beta * y[i]);
}
// y (cols) = alpha * A' x + beta * y: one CTA per column, fixed-order tree reduction (bit-reproducible)
__global__ void __launch_bounds__(256) k_dgemv_t(long long rows, double alpha, const double* __restrict__ A, long long lda,
                                                 const double* __restrict__ x, double beta, double* __restrict__ y) {
    __shared__ double red[256];
    const long long j = blockIdx.x;
    double s = 0.0;
    for (long long i = threadIdx.x; i < rows; i += 256) s = fma(A[j * lda + i], x[i], s);
    red[threadIdx.x] = s;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) y[j] = beta == 0.0 ? alpha * red[0] : fma(alpha, red[0], beta * y[j]);
}
// Gs = diag(di) G   (misc.py:1269-1271 `scale(Gs, W, trans='T', inverse='I')` for the 'l' cone)
__global__ void k_scale_rows(long long m, long long n, const double* __restrict__ di, const double* G, double* Gs) {   // G may alias Gs
    const long long total = m * n;
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x)
        Gs[e] = G[e] * di[e % m];
}
// S(lower) += H(lower), then mirror the lower triangle into the upper one (misc.py:1275-1276: K += H; symm(K, n))
__global__ void k_add_h_symm(long long n, const double* __restrict__ H, double* __restrict__ S) {
    const long long total = n * n;
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        const long long j = e / n, i = e - j * n;
        if (i < j) continue;
        const double v = S[e] + (H ? H[e] : 0.0);
        S[e] = v;
        if (i > j) S[i * n + j] = v;
    }
}
// values of the lower triangle of K22 = K[p:, p:] in CCS order (column by column, rows j..q-1): what the engine's
// analysis of a dense pattern expects
__global__ void k_pack_lower(long long n, long long p, const double* __restrict__ K, double* __restrict__ out) {
    const long long q = n - p, total = q * q;
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        const long long j = e / q, i = e - j * q;
        if (i < j) continue;
        out[j * q - j * (j - 1) / 2 + (i - j)] = K[(p + j) * n + p + i];
    }
}
// x := R^-T x (trans = 1) or R^-1 x (trans = 0), R p x p upper triangular column-major: one warp, substitution with the
// vector in shared memory (p is the number of equality constraints)
__global__ void k_trsv_upper(int p, const double* __restrict__ R, double* __restrict__ x, int trans) {
    extern __shared__ double xs[];
    const int lane = threadIdx.x;
    for (int i = lane; i < p; i += 32) xs[i] = x[i];
    __syncwarp();
    if (trans) {            // R' lower: forward
        for (int j = 0; j < p; j++) {
            double s = 0.0;
            for (int i = lane; i < j; i += 32) s = fma(R[(long long)j * p + i], xs[i], s);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
            if (lane == 0) xs[j] = (xs[j] - s) / R[(long long)j * p + j];
            __syncwarp();
        }
    } else {                // R upper: backward, column oriented
        for (int j = p - 1; j >= 0; j--) {
            if (lane == 0) xs[j] = xs[j] / R[(long long)j * p + j];
            __syncwarp();
            const double xj = xs[j];
            for (int i = lane; i < j; i += 32) xs[i] = fma(-R[(long long)j * p + i], xj, xs[i]);
            __syncwarp();
        }
    }
    for (int i = lane; i < p; i += 32) x[i] = xs[i];
}

int grid1(long long work) { return (int)std::max<long long>(1, std::min<long long>((work + 255) / 256, 148 * 8)); }

}  // namespace

struct b200s_kktd {
    i64 n = 0, ml = 0, p = 0;
    int device = 0;
    std::vector<double> V, T, R;       // compact WY of A' = QR (host copies, kept for b200s_kktd_get_qr)
    std::vector<double> Gh;            // host copy of G until the first factor() uploads it
    bool uploaded = false;
    CholPlan plan;
    CholOpts opts;
    CholDevice* chol = nullptr;
    CholTimes times;
    cudaStream_t own_stream = nullptr; // used when n == p (no Cholesky object)
    double *d_G = nullptr, *d_Gs = nullptr, *d_V = nullptr, *d_T = nullptr, *d_R = nullptr, *d_K = nullptr, *d_H = nullptr,
           *d_W1 = nullptr, *d_W2 = nullptr, *d_di = nullptr, *d_x = nullptr, *d_y = nullptr, *d_z = nullptr, *d_t = nullptr,
           *d_t2 = nullptr, *d_pack = nullptr, *d_yy = nullptr;
    bool factored = false;
    double ms_factor = 0, ms_solve = 0;
    cudaEvent_t ev[2] = {};
    long long launches = 0;
    // the solve as ONE graph: x, y, z through a pinned staging buffer (fixed addresses), every kernel of misc.py:1284-1345 and the
    // sweeps of the Cholesky object recorded once (second call: the first one allocates the workspaces) and replayed -- a 143-variable
    // LP makes ~90 solves of a few microseconds of work each, 15 launches and 6 pageable copies apiece otherwise
    double* h_stage = nullptr;
    cudaGraphExec_t solve_graph = nullptr, factor_graph = nullptr;     // factor(W) with H = None likewise (di and the pivot word staged)
    double* h_fstage = nullptr;
    long long solve_calls = 0, factor_calls = 0;
    bool use_graph = true;
    cudaStream_t stream() const { return chol ? (cudaStream_t)chol_device_stream(chol) : own_stream; }
    ~b200s_kktd() {
        if (uploaded || chol) cudaSetDevice(device);
        if (chol) { cudaStreamSynchronize((cudaStream_t)chol_device_stream(chol)); chol_device_destroy(chol); }
        if (own_stream) { cudaStreamSynchronize(own_stream); cudaStreamDestroy(own_stream); }
        for (double* q : {d_G, d_Gs, d_V, d_T, d_R, d_K, d_H, d_W1, d_W2, d_di, d_x, d_y, d_z, d_t, d_t2, d_pack, d_yy}) pool_free(q);
        for (auto& e : ev) if (e) cudaEventDestroy(e);
        if (solve_graph) cudaGraphExecDestroy(solve_graph);
        if (factor_graph) cudaGraphExecDestroy(factor_graph);
        pinned_free(h_stage); pinned_free(h_fstage);
    }
};

namespace {

// Householder QR of M = A' (n x p, column-major, overwritten by V below the diagonal and R on/above it) with LAPACK's
// dlarfg convention (H_i = I - tau_i v_i v_i', v_i(i) = 1), then the triangular factor T of the compact WY form
// Q = H_1 ... H_p = I - V T V' (forward, columnwise: dlarft).
void householder_qr(i64 n, i64 p, std::vector<double>& M, std::vector<double>& tau) {
    tau.assign((size_t)p, 0.0);
    for (i64 k = 0; k < p; k++) {
        double* c = &M[(size_t)k * n];
        double xnorm = 0.0;
        for (i64 i = k + 1; i < n; i++) xnorm = std::hypot(xnorm, c[i]);
        const double alpha = c[k];
        if (xnorm == 0.0) { tau[k] = 0.0; continue; }
        const double beta = -std::copysign(std::hypot(alpha, xnorm), alpha);
        tau[k] = (beta - alpha) / beta;
        const double sc = 1.0 / (alpha - beta);
        for (i64 i = k + 1; i < n; i++) c[i] *= sc;
        c[k] = beta;
        for (i64 j = k + 1; j < p; j++) {       // apply H_k to the remaining columns
            double* cj = &M[(size_t)j * n];
            double w = cj[k];
            for (i64 i = k + 1; i < n; i++) w += c[i] * cj[i];
            w *= tau[k];
            cj[k] -= w;
            for (i64 i = k + 1; i < n; i++) cj[i] -= w * c[i];
        }
    }
}

int gemm(cudaStream_t st, int M, int N, int Kd, double alpha, const double* A, long long sai, long long sak, const double* B,
         long long sbk, long long sbj, double beta, double* Cm, long long ldc, long long& launches) {
    if (M <= 0 || N <= 0) return ST_OK;
    dim3 g((M + GT - 1) / GT, (N + GT - 1) / GT);
    k_dgemm<<<g, 256, 0, st>>>(M, N, Kd, alpha, A, sai, sak, B, sbk, sbj, beta, Cm, ldc);
    launches++;
    return ST_OK;
}

// x := Q' x (trans = 1) or Q x (trans = 0):  x -= V (op(T) (V' x))
int apply_q(b200s_kktd* K, cudaStream_t st, double* x, int trans) {
    const long long n = K->n, p = K->p;
    if (p == 0) return ST_OK;
    k_dgemv_t<<<(unsigned)p, 256, 0, st>>>(n, 1.0, K->d_V, n, x, 0.0, K->d_t);
    if (trans) k_dgemv_t<<<(unsigned)p, 256, 0, st>>>(p, 1.0, K->d_T, p, K->d_t, 0.0, K->d_t2);
    else k_dgemv_n<<<grid1(p), 256, 0, st>>>(p, p, 1.0, K->d_T, p, K->d_t, 0.0, K->d_t2);
    k_dgemv_n<<<grid1(n), 256, 0, st>>>(n, p, -1.0, K->d_V, n, K->d_t2, 1.0, x);
    K->launches += 3;
    return ST_OK;
}

}  // namespace

extern "C" {

/* misc.kkt_chol(G, dims, A) (src/python/misc.py:1213-1256): G ml x n and A p x n dense column-major (leading dimensions
 * ml and p).  The QR factorization of A' runs on the host once per problem (2 n p^2 flops), everything per iteration
 * runs on the device. */
b200s_status b200s_kktd_create(b200s_int n, b200s_int ml, b200s_int p, const double* G, const double* A, b200s_kktd** out) {
    if (!out) return B200S_INVALID;
    *out = nullptr;
    if (n < 0 || ml < 0 || p < 0 || p > n || (ml > 0 && n > 0 && !G) || (p > 0 && !A)) return B200S_INVALID;
    if ((double)n * (double)n > 2.0e9 || (double)ml * (double)n > 8.0e9) { set_last_error("kkt 'chol': dense system too large"); return B200S_TOO_LARGE; }
    b200s_kktd* K = new (std::nothrow) b200s_kktd();
    if (!K) return B200S_OUT_OF_MEMORY;
    K->n = n; K->ml = ml; K->p = p; K->device = current_device();
    auto fail = [&](int st) { delete K; return (b200s_status)st; };
    try {
        K->Gh.assign(G, G + (size_t)ml * (size_t)n);
        // A' (n x p), QR, T
        K->V.assign((size_t)n * p, 0.0);
        for (i64 j = 0; j < p; j++)
            for (i64 i = 0; i < n; i++) K->V[(size_t)j * n + i] = A[(size_t)i * p + j];
        std::vector<double> tau;
        householder_qr(n, p, K->V, tau);
        K->R.assign((size_t)p * p, 0.0);
        for (i64 j = 0; j < p; j++) {
            for (i64 i = 0; i <= j; i++) K->R[(size_t)j * p + i] = K->V[(size_t)j * n + i];
            for (i64 i = 0; i < j; i++) K->V[(size_t)j * n + i] = 0.0;
            K->V[(size_t)j * n + j] = 1.0;
            if (K->R[(size_t)j * p + j] == 0.0) { set_last_error("kkt 'chol': Rank(A) < p"); return fail(ST_SINGULAR); }
        }
        K->T.assign((size_t)p * p, 0.0);
        std::vector<double> w((size_t)p);
        for (i64 i = 0; i < p; i++) {
            // T(0:i, i) = -tau_i T(0:i, 0:i) V(:, 0:i)' v_i
            for (i64 a = 0; a < i; a++) {
                double s = 0.0;
                for (i64 r = i; r < n; r++) s += K->V[(size_t)a * n + r] * K->V[(size_t)i * n + r];
                w[a] = -tau[i] * s;
            }
            for (i64 a = 0; a < i; a++) {
                double s = 0.0;
                for (i64 b = a; b < i; b++) s += K->T[(size_t)b * p + a] * w[b];
                K->T[(size_t)i * p + a] = s;
            }
            K->T[(size_t)i * p + i] = tau[i];
        }
        // dense lower-triangular pattern of order q = n - p, natural order (one supernode)
        const i64 q = n - p;
        if (q > 0) {
            std::vector<i64> cp((size_t)q + 1), ri((size_t)(q * (q + 1) / 2)), perm((size_t)q);
            i64 e = 0;
            for (i64 j = 0; j < q; j++) { cp[j] = e; for (i64 i = j; i < q; i++) ri[e++] = i; perm[j] = j; }
            cp[q] = e;
            K->opts.nmethods = 1;
            chol_analyze(q, cp.data(), ri.data(), 'L', perm.data(), K->opts, K->plan);
        }
    } catch (const std::bad_alloc&) {
        return fail(ST_OOM);
    } catch (const std::exception& e) {
        set_last_error(e.what());
        return fail(ST_INVALID);
    }
    *out = K;
    return B200S_OK;
}

void b200s_kktd_free(b200s_kktd* K) { delete K; }

static int kktd_ensure_device(b200s_kktd* K) {
    if (K->uploaded) return ST_OK;
    if (device_count() <= 0) { set_last_error("no CUDA device available"); return ST_NO_DEVICE; }
    CUDA_TRY(cudaSetDevice(K->device));
    const i64 n = K->n, ml = K->ml, p = K->p;
    auto up = [&](double** d, const double* h, size_t count) -> int {
        CUDA_TRY(pool_malloc((void**)d, std::max<size_t>(count, 1) * sizeof(double)));
        if (h && count) CUDA_TRY(cudaMemcpy(*d, h, count * sizeof(double), cudaMemcpyHostToDevice));
        return ST_OK;
    };
    int rc;
    const size_t sn = (size_t)n, sm = (size_t)ml, sp = (size_t)p;
    if ((rc = up(&K->d_G, K->Gh.data(), sm * sn)) || (rc = up(&K->d_Gs, nullptr, sm * sn)) || (rc = up(&K->d_V, K->V.data(), sn * sp)) ||
        (rc = up(&K->d_T, K->T.data(), sp * sp)) || (rc = up(&K->d_R, K->R.data(), sp * sp)) || (rc = up(&K->d_K, nullptr, sn * sn)) ||
        (rc = up(&K->d_H, nullptr, sn * sn)) || (rc = up(&K->d_W1, nullptr, sn * sp)) || (rc = up(&K->d_W2, nullptr, sn * sp)) ||
        (rc = up(&K->d_di, nullptr, sm)) || (rc = up(&K->d_x, nullptr, sn)) || (rc = up(&K->d_y, nullptr, sp)) ||
        (rc = up(&K->d_z, nullptr, sm)) || (rc = up(&K->d_t, nullptr, sp)) || (rc = up(&K->d_t2, nullptr, sp)) ||
        (rc = up(&K->d_yy, nullptr, sp)) || (rc = up(&K->d_pack, nullptr, (size_t)((n - p) * (n - p + 1) / 2))))
        return rc;
    if (n - p > 0) {
        int st = ST_OK;
        K->chol = chol_device_create(K->plan, K->opts, K->device, &st);
        if (!K->chol) return st;
    } else CUDA_TRY(cudaStreamCreateWithFlags(&K->own_stream, cudaStreamNonBlocking));
    for (auto& e : K->ev) CUDA_TRY(cudaEventCreate(&e));
    K->Gh.clear(); K->Gh.shrink_to_fit();
    K->uploaded = true;
    return ST_OK;
}

// the device part of factor(): di (and H) -> device, K = Q'(H + Gs'Gs)Q, lower triangle of K22 packed for the Cholesky object
static int kktd_factor_enqueue(b200s_kktd* K, cudaStream_t st, const double* di, const double* H) {
    const long long n = K->n, ml = K->ml, p = K->p, q = n - p;
    K->launches = 0;
    if (ml) CUDA_TRY(cudaMemcpyAsync(K->d_di, di, ml * sizeof(double), cudaMemcpyHostToDevice, st));
    if (H) CUDA_TRY(cudaMemcpyAsync(K->d_H, H, n * n * sizeof(double), cudaMemcpyHostToDevice, st));
    // Gs = diag(di) G;  S = Gs' Gs (+ H), symmetric, in d_K
    if (ml) { k_scale_rows<<<grid1(ml * n), 256, 0, st>>>(ml, n, K->d_di, K->d_G, K->d_Gs); K->launches++; }
    if (ml) gemm(st, (int)n, (int)n, (int)ml, 1.0, K->d_Gs, ml, 1, K->d_Gs, 1, ml, 0.0, K->d_K, n, K->launches);
    else CUDA_TRY(cudaMemsetAsync(K->d_K, 0, n * n * sizeof(double), st));
    k_add_h_symm<<<grid1(n * n), 256, 0, st>>>(n, H ? K->d_H : nullptr, K->d_K); K->launches++;
    if (p) {
        // K := Q' S Q with Q = I - V T V':   W1 = T' (V' S);  S -= V W1;   W2 = (S V) T;  K = S - W2 V'
        gemm(st, (int)p, (int)n, (int)n, 1.0, K->d_V, n, 1, K->d_K, 1, n, 0.0, K->d_W2, p, K->launches);          // W2 (p x n) = V' S
        gemm(st, (int)p, (int)n, (int)p, 1.0, K->d_T, p, 1, K->d_W2, 1, p, 0.0, K->d_W1, p, K->launches);          // W1 = T' W2
        gemm(st, (int)n, (int)n, (int)p, -1.0, K->d_V, 1, n, K->d_W1, 1, p, 1.0, K->d_K, n, K->launches);          // S -= V W1
        gemm(st, (int)n, (int)p, (int)n, 1.0, K->d_K, 1, n, K->d_V, 1, n, 0.0, K->d_W2, n, K->launches);           // W2 (n x p) = S V
        gemm(st, (int)n, (int)p, (int)p, 1.0, K->d_W2, 1, n, K->d_T, 1, p, 0.0, K->d_W1, n, K->launches);          // W1 (n x p) = W2 T
        gemm(st, (int)n, (int)n, (int)p, -1.0, K->d_W1, 1, n, K->d_V, n, 1, 1.0, K->d_K, n, K->launches);          // K = S - W1 V'
    }
    if (q > 0) { k_pack_lower<<<grid1(q * q), 256, 0, st>>>(n, p, K->d_K, K->d_pack); K->launches++; }
    CUDA_TRY(cudaGetLastError());
    return ST_OK;
}

static int kktd_factor_impl(b200s_kktd* K, const double* di, const double* H, b200s_int* minor_out) {
    B200S_NVTX("kktd_factor_impl");
    if (!K || (K->ml > 0 && !di)) return ST_INVALID;
    K->factored = false;
    const long long n = K->n, ml = K->ml, p = K->p, q = n - p;
    if (n == 0) { K->factored = true; return ST_OK; }
    { int rc0 = kktd_ensure_device(K); if (rc0) return rc0; }
    CUDA_TRY(cudaSetDevice(K->device));
    cudaStream_t st = K->stream();
    int rc;
    i64 minor = q;
    // From the second call on (the first one allocates the Cholesky object's panels) factor(W) with H = None is ONE graph: di
    // through pinned staging, the products above, the factorization of K22 enqueued without read-back, the pivot word back
    if (K->factor_calls++ > 0 && K->use_graph && !H && q > 0) {
        if (!K->h_fstage) CUDA_TRY(pinned_malloc((void**)&K->h_fstage, (size_t)(ml + 2) * sizeof(double)));
        double* hdi = K->h_fstage;
        int* hminor = reinterpret_cast<int*>(K->h_fstage + ml);
        if (!K->factor_graph) {
            cudaGraph_t gr = nullptr;
            CUDA_TRY(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
            rc = kktd_factor_enqueue(K, st, hdi, nullptr);
            if (rc == ST_OK) rc = chol_device_factor_enqueue(K->chol, K->d_pack);
            if (rc == ST_OK && cudaMemcpyAsync(hminor, chol_device_minor_ptr(K->chol), sizeof(int), cudaMemcpyDeviceToHost, st) != cudaSuccess) rc = ST_CUDA;
            cudaError_t ce = cudaStreamEndCapture(st, &gr);
            if (rc != ST_OK || ce != cudaSuccess) {      // not capturable here: plain launches from now on
                if (gr) cudaGraphDestroy(gr);
                cudaGetLastError();
                K->use_graph = false;
                return kktd_factor_impl(K, di, H, minor_out);
            }
            cudaError_t ie = cudaGraphInstantiate(&K->factor_graph, gr, 0);
            cudaGraphDestroy(gr);
            CUDA_TRY(ie);
        }
        if (ml) memcpy(hdi, di, ml * sizeof(double));
        chol_device_mark_numeric(K->chol, false);
        CUDA_TRY(cudaEventRecord(K->ev[0], st));
        CUDA_TRY(cudaGraphLaunch(K->factor_graph, st));
        CUDA_TRY(cudaEventRecord(K->ev[1], st));
        CUDA_TRY(cudaStreamSynchronize(st));
        float ms;
        cudaEventElapsedTime(&ms, K->ev[0], K->ev[1]); K->ms_factor = ms;
        if (*hminor != 0x7fffffff) {
            if (minor_out) *minor_out = *hminor;
            return ST_NOT_POSDEF;
        }
        chol_device_mark_numeric(K->chol, true);
        if (minor_out) *minor_out = q;
        K->factored = true;
        return ST_OK;
    }
    CUDA_TRY(cudaEventRecord(K->ev[0], st));
    if ((rc = kktd_factor_enqueue(K, st, di, H))) return rc;
    if (q > 0) {
        rc = chol_device_factorize(K->chol, K->d_pack, true, &minor, &K->times);
        if (minor_out) *minor_out = minor;
        if (rc != ST_OK) return rc;
    }
    CUDA_TRY(cudaEventRecord(K->ev[1], st));
    CUDA_TRY(cudaStreamSynchronize(st));
    float ms;
    cudaEventElapsedTime(&ms, K->ev[0], K->ev[1]); K->ms_factor = ms;
    K->factored = true;
    return ST_OK;
}

/* factor(W, H) of misc.kkt_chol (misc.py:1258-1282): di = W['di'] (ml), H dense n x n column-major (lower triangle
 * referenced) or NULL.  NOT_POSDEF with *minor = failing column of K22 when Q2'(H + Gs'Gs)Q2 is not positive definite
 * (lapack.potrf raising ArithmeticError in the reference). */
b200s_status b200s_kktd_factor(b200s_kktd* K, const double* di, const double* H, b200s_int* minor_out) {
    return (b200s_status)kktd_factor_impl(K, di, H, minor_out);
}

// the device part of solve(): hx, hy, hz -> device, misc.py:1303-1342, device -> hx, hy, hz; everything on stream st
static int kktd_solve_enqueue(b200s_kktd* K, cudaStream_t st, double* x, double* y, double* z) {
    const long long n = K->n, ml = K->ml, p = K->p, q = n - p;
    int rc;
    CUDA_TRY(cudaMemcpyAsync(K->d_x, x, n * sizeof(double), cudaMemcpyHostToDevice, st));
    if (p) CUDA_TRY(cudaMemcpyAsync(K->d_yy, y, p * sizeof(double), cudaMemcpyHostToDevice, st));
    if (ml) {
        CUDA_TRY(cudaMemcpyAsync(K->d_z, z, ml * sizeof(double), cudaMemcpyHostToDevice, st));
        // bzp = di o bz;  x += Gs' bzp                                                   (misc.py:1303-1309)
        k_scale_rows<<<grid1(ml), 256, 0, st>>>(ml, 1, K->d_di, K->d_z, K->d_z);
        k_dgemv_t<<<(unsigned)n, 256, 0, st>>>(ml, 1.0, K->d_Gs, ml, K->d_z, 1.0, K->d_x);
    }
    if ((rc = apply_q(K, st, K->d_x, 1))) return rc;                                      // x := [Q1 Q2]' x   (:1310)
    if (p) {
        CUDA_TRY(cudaMemcpyAsync(K->d_y, K->d_x, p * sizeof(double), cudaMemcpyDeviceToDevice, st));     // y := x[:p]   (:1314-1315)
        CUDA_TRY(cudaMemcpyAsync(K->d_x, K->d_yy, p * sizeof(double), cudaMemcpyDeviceToDevice, st));    // x[:p] := by
        k_trsv_upper<<<1, 32, p * sizeof(double), st>>>((int)p, K->d_R, K->d_x, 1);        // v = R^-T by     (:1318-1319)
        if (q) k_dgemv_n<<<grid1(q), 256, 0, st>>>(q, p, -1.0, K->d_K + p, n, K->d_x, 1.0, K->d_x + p);   // x[p:] -= K21 v (:1323-1324)
    }
    if (q && (rc = chol_device_solve_async(K->chol, 0, K->d_x + p, 1, q))) return rc;     // w = K22^-1 x[p:] (:1325)
    if (p) {
        k_dgemv_t<<<(unsigned)p, 256, 0, st>>>(n, -1.0, K->d_K, n, K->d_x, 1.0, K->d_y);    // y -= [K11 K12] x  (:1329)
        k_trsv_upper<<<1, 32, p * sizeof(double), st>>>((int)p, K->d_R, K->d_y, 0);        // y := R^-1 y       (:1334)
    }
    if ((rc = apply_q(K, st, K->d_x, 0))) return rc;                                      // x := [Q1 Q2] x    (:1337)
    if (ml) k_dgemv_n<<<grid1(ml), 256, 0, st>>>(ml, n, 1.0, K->d_Gs, ml, K->d_x, -1.0, K->d_z);   // z := Gs x - bzp (:1342)
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaMemcpyAsync(x, K->d_x, n * sizeof(double), cudaMemcpyDeviceToHost, st));
    if (p) CUDA_TRY(cudaMemcpyAsync(y, K->d_y, p * sizeof(double), cudaMemcpyDeviceToHost, st));
    if (ml) CUDA_TRY(cudaMemcpyAsync(z, K->d_z, ml * sizeof(double), cudaMemcpyDeviceToHost, st));
    return ST_OK;
}

static int kktd_solve_impl(b200s_kktd* K, double* x, double* y, double* z) {
    B200S_NVTX("kktd_solve_impl");
    if (!K || (K->n > 0 && !x) || (K->p > 0 && !y) || (K->ml > 0 && !z)) return ST_INVALID;
    if (!K->factored) { set_last_error("kkt 'chol': solve called before a successful factor"); return ST_INVALID; }
    const long long n = K->n, ml = K->ml, p = K->p;
    if (n == 0) return ST_OK;
    CUDA_TRY(cudaSetDevice(K->device));
    cudaStream_t st = K->stream();
    int rc;
    if (K->solve_calls++ == 0 || !K->use_graph) {     // first call: plain launches (workspaces of the Cholesky object are allocated here)
        CUDA_TRY(cudaEventRecord(K->ev[0], st));
        if ((rc = kktd_solve_enqueue(K, st, x, y, z))) return rc;
    } else {
        if (!K->h_stage) CUDA_TRY(pinned_malloc((void**)&K->h_stage, (size_t)(n + p + ml) * sizeof(double)));
        double *hx = K->h_stage, *hy = hx + n, *hz = hy + p;
        if (!K->solve_graph) {
            cudaGraph_t gr = nullptr;
            CUDA_TRY(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
            rc = kktd_solve_enqueue(K, st, hx, hy, hz);
            cudaError_t ce = cudaStreamEndCapture(st, &gr);
            if (rc != ST_OK || ce != cudaSuccess) {      // not capturable here: stay with plain launches
                if (gr) cudaGraphDestroy(gr);
                cudaGetLastError();
                K->use_graph = false;
                K->solve_calls = 0;
                return kktd_solve_impl(K, x, y, z);
            }
            cudaError_t ie = cudaGraphInstantiate(&K->solve_graph, gr, 0);
            cudaGraphDestroy(gr);
            CUDA_TRY(ie);
        }
        memcpy(hx, x, n * sizeof(double));
        if (p) memcpy(hy, y, p * sizeof(double));
        if (ml) memcpy(hz, z, ml * sizeof(double));
        CUDA_TRY(cudaEventRecord(K->ev[0], st));
        CUDA_TRY(cudaGraphLaunch(K->solve_graph, st));
        CUDA_TRY(cudaEventRecord(K->ev[1], st));
        CUDA_TRY(cudaStreamSynchronize(st));
        memcpy(x, hx, n * sizeof(double));
        if (p) memcpy(y, hy, p * sizeof(double));
        if (ml) memcpy(z, hz, ml * sizeof(double));
        float ms;
        cudaEventElapsedTime(&ms, K->ev[0], K->ev[1]); K->ms_solve = ms;
        return ST_OK;
    }
    CUDA_TRY(cudaEventRecord(K->ev[1], st));
    CUDA_TRY(cudaStreamSynchronize(st));
    float ms;
    cudaEventElapsedTime(&ms, K->ev[0], K->ev[1]); K->ms_solve = ms;
    return ST_OK;
}

/* solve(x, y, z) of misc.kkt_chol (misc.py:1284-1345), in place on host vectors: bx, by, bz -> ux, uy, W uz. */
b200s_status b200s_kktd_solve(b200s_kktd* K, double* x, double* y, double* z) { return (b200s_status)kktd_solve_impl(K, x, y, z); }

b200s_status b200s_kktd_info(const b200s_kktd* K, b200s_kktd_info_t* info) {
    if (!K || !info) return B200S_INVALID;
    info->n = K->n; info->ml = K->ml; info->p = K->p;
    const double n = (double)K->n, m = (double)K->ml, p = (double)K->p, q = n - p;
    info->flops = 2.0 * m * n * n + 8.0 * n * n * p + q * q * q / 3.0;
    info->ms_factor = K->ms_factor; info->ms_solve = K->ms_solve; info->launches = K->launches;
    return B200S_OK;
}

/* the host QR of A' in compact WY form (tests): V n x p (unit lower trapezoid), T p x p upper, R p x p upper */
b200s_status b200s_kktd_get_qr(const b200s_kktd* K, double* V, double* T, double* R) {
    if (!K) return B200S_INVALID;
    if (V) memcpy(V, K->V.data(), K->V.size() * sizeof(double));
    if (T) memcpy(T, K->T.data(), K->T.size() * sizeof(double));
    if (R) memcpy(R, K->R.data(), K->R.size() * sizeof(double));
    return B200S_OK;
}

}  // extern "C"
