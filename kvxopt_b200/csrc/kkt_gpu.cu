// Device-side reduced KKT solver for LP / QP cones: the GPU counterpart of the reference's
// misc.kkt_chol2 (src/python/misc.py:1352-1567) for problems with only componentwise inequalities (dims['l']).
//
//   factor(W, H):  S = H + G' diag(di)^2 G  (+ A'A when the first S is singular, misc.py:1427-1447)  -- assembled ON THE
//                  DEVICE in a fixed pattern: every stored entry of S is  H_e + sum_t di[k_t]^2 * (G[k_t,i] G[k_t,j]),
//                  the products are constants of the pattern, so one kernel evaluates a precomputed term list
//                  (replaces base.gemm + base.syrk(partial) + the spmatrix add of misc.py:1418-1455, which cost more
//                  host time per interior-point iteration than the GPU factorization of S);
//                  P S P' = L L' by the multifrontal engine (chol_gpu.cu) straight from device values;
//                  Asct = L^-1 P A' (dense n x p on the device), K = Asct' Asct, K = Lk Lk' (one CTA).
//   solve(x,y,z):  the three-step block elimination of misc.py:1489-1565, one upload and one download per call:
//                  z := di.z;  x := L^-1 P (x + G'(di.z) [+ A'y]);  y := K^-1 (Asct'x - y);  x := P' L^-T (x - Asct y);
//                  z := di.(G x) - z.
// Host pointers at the C ABI (b200s_kkt_*, include/b200sparse.h); everything numeric runs on the device.
#include "../../include/b200sparse.h"
#include "gpu.hpp"
#include "nvtx_range.hpp"
#include "devpool.hpp"
#include <cuda_runtime.h>
#include <algorithm>
#include <cmath>
#include <cstdio>
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

// ---- kernels --------------------------------------------------------------------------------------------------
// S_e = H[hsrc[e]] + sum_{t in terms(e)} w(t) * prod[t],  w = di[row]^2 for a G'G term, 1 for an A'A term (row = -1)
__global__ void k_kkt_assemble(long long nnzS, const long long* __restrict__ tptr, const int* __restrict__ trow,
                               const double* __restrict__ tprod, const int* __restrict__ hsrc, const double* __restrict__ Hx,
                               const double* __restrict__ di, double* __restrict__ Sx) {
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < nnzS; e += (long long)gridDim.x * blockDim.x) {
        double v = hsrc[e] >= 0 ? Hx[hsrc[e]] : 0.0;
        for (long long t = tptr[e]; t < tptr[e + 1]; t++) {
            const int k = trow[t];
            const double d = k >= 0 ? di[k] : 1.0;
            v = fma(d * d, tprod[t], v);
        }
        Sx[e] = v;
    }
}
// z := di . z ; w := di . z (new)          (scale(z, W, trans='T', inverse='I') for the 'l' cone, then Gs' z = G'(di . z))
__global__ void k_kkt_scale_z(long long ml, const double* __restrict__ di, double* __restrict__ z, double* __restrict__ w) {
    for (long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x; k < ml; k += (long long)gridDim.x * blockDim.x) {
        const double z1 = z[k] * di[k];
        z[k] = z1;
        w[k] = z1 * di[k];
    }
}
// x_j += sum_k M(k,j) v_k  for a CCS matrix M (rows x n): one thread per column, fixed order, no atomics
__global__ void k_kkt_gemv_t(long long n, const long long* __restrict__ cp, const int* __restrict__ ri, const double* __restrict__ vx,
                             const double* __restrict__ v, double* __restrict__ x) {
    for (long long j = blockIdx.x * (long long)blockDim.x + threadIdx.x; j < n; j += (long long)gridDim.x * blockDim.x) {
        double s = 0.0;
        for (long long q = cp[j]; q < cp[j + 1]; q++) s = fma(vx[q], v[ri[q]], s);
        x[j] += s;
    }
}
// z_k := di_k * sum_j G(k,j) x_j - z_k   (CSR of G: one thread per row)
__global__ void k_kkt_gz(long long ml, const long long* __restrict__ rp, const int* __restrict__ cj, const double* __restrict__ vx,
                         const double* __restrict__ di, const double* __restrict__ x, double* __restrict__ z) {
    for (long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x; k < ml; k += (long long)gridDim.x * blockDim.x) {
        double s = 0.0;
        for (long long q = rp[k]; q < rp[k + 1]; q++) s = fma(vx[q], x[cj[q]], s);
        z[k] = di[k] * s - z[k];
    }
}
// Asct(i, a) = A(a, i)   (A is p x n in CCS; Asct is dense n x p, column-major)
__global__ void k_kkt_fill_at(long long n, const long long* __restrict__ cp, const int* __restrict__ ri, const double* __restrict__ vx,
                              double* __restrict__ At) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        for (long long q = cp[i]; q < cp[i + 1]; q++) At[i + (long long)ri[q] * n] = vx[q];
}
// out[a + b*p] = sum_i M(i,a) N(i,b)  -- one CTA per (a, b), fixed-order tree reduction (deterministic)
__global__ void __launch_bounds__(256) k_kkt_dots(long long n, int p, const double* __restrict__ M, const double* __restrict__ N_, int ldn,
                                                  double* __restrict__ out, int lower_only) {
    const int a = blockIdx.x, b = blockIdx.y;
    if (lower_only && b > a) return;
    __shared__ double red[256];
    double s = 0.0;
    for (long long i = threadIdx.x; i < n; i += 256) s = fma(M[i + (long long)a * n], N_[i + (long long)b * ldn], s);
    red[threadIdx.x] = s;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        out[a + (long long)b * p] = red[0];
        if (lower_only) out[b + (long long)a * p] = red[0];
    }
}
// K = Lk Lk' in place (lower, column-major p x p), one CTA, right-looking; flag = first non-positive pivot + 1
__global__ void __launch_bounds__(256) k_kkt_potrf(int p, double* __restrict__ K, int* __restrict__ flag) {
    __shared__ double piv;
    for (int j = 0; j < p; j++) {
        if (threadIdx.x == 0) {
            const double d = K[j + (long long)j * p];
            if (!(d > 0.0)) { if (*flag == 0) *flag = j + 1; piv = 1.0; }
            else piv = sqrt(d);
            K[j + (long long)j * p] = piv;
        }
        __syncthreads();
        const double r = 1.0 / piv;
        for (int i = j + 1 + threadIdx.x; i < p; i += 256) K[i + (long long)j * p] *= r;
        __syncthreads();
        const int m = p - j - 1;
        for (long long t = threadIdx.x; t < (long long)m * m; t += 256) {
            const int i = j + 1 + (int)(t % m), c = j + 1 + (int)(t / m);
            if (i >= c) K[i + (long long)c * p] -= K[i + (long long)j * p] * K[c + (long long)j * p];
        }
        __syncthreads();
    }
}
// y := Lk^-T Lk^-1 (t - y)   (one warp; p is small)
__global__ void k_kkt_ksolve(int p, const double* __restrict__ K, const double* __restrict__ t, double* __restrict__ y) {
    extern __shared__ double v[];
    const int lane = threadIdx.x;
    for (int i = lane; i < p; i += 32) v[i] = t[i] - y[i];
    __syncwarp();
    for (int j = 0; j < p; j++) {
        if (lane == 0) v[j] /= K[j + (long long)j * p];
        __syncwarp();
        const double vj = v[j];
        for (int i = j + 1 + lane; i < p; i += 32) v[i] -= K[i + (long long)j * p] * vj;
        __syncwarp();
    }
    for (int j = p - 1; j >= 0; j--) {
        double s = 0.0;
        for (int i = j + 1 + lane; i < p; i += 32) s += K[i + (long long)j * p] * v[i];
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == 0) v[j] = (v[j] - s) / K[j + (long long)j * p];
        __syncwarp();
    }
    for (int i = lane; i < p; i += 32) y[i] = v[i];
}
// x_i -= sum_a Asct(i,a) y_a
__global__ void k_kkt_axpy_cols(long long n, int p, const double* __restrict__ At, const double* __restrict__ y, double* __restrict__ x) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        double s = 0.0;
        for (int a = 0; a < p; a++) s = fma(At[i + (long long)a * n], y[a], s);
        x[i] -= s;
    }
}

struct Ccs {
    std::vector<long long> p;
    std::vector<int> i;
    std::vector<double> x;
};

}  // namespace

struct b200s_kkt {
    i64 n = 0, ml = 0, p = 0;
    Ccs G, Gr /*CSR of G*/, A;
    std::vector<i64> Hp, Hi;       // lower-triangle pattern of H as the caller stores it (entries above the diagonal ignored)
    bool hasH = false, withA = false;
    // pattern of S (lower CCS, 64-bit as the engine's analyze takes it) and the assembly plan
    std::vector<i64> Sp, Si;
    std::vector<long long> tptr;
    std::vector<int> trow, hsrc;
    std::vector<double> tprod;
    CholPlan plan;
    CholOpts opts;
    CholDevice* chol = nullptr;
    CholTimes times;
    int device = 0;
    i64 nnzH = 0;
    // device
    long long *d_tptr = nullptr, *d_Gp = nullptr, *d_Grp = nullptr, *d_Ap = nullptr;
    int *d_trow = nullptr, *d_hsrc = nullptr, *d_Gi = nullptr, *d_Gcj = nullptr, *d_Ai = nullptr, *d_flag = nullptr;
    double *d_tprod = nullptr, *d_Gx = nullptr, *d_Grx = nullptr, *d_Ax = nullptr, *d_Hx = nullptr, *d_di = nullptr, *d_Sx = nullptr,
           *d_x = nullptr, *d_y = nullptr, *d_z = nullptr, *d_w = nullptr, *d_At = nullptr, *d_K = nullptr, *d_t = nullptr;
    bool uploaded = false, factored = false;
    double ms_assemble = 0, ms_factor = 0, ms_solve = 0;
    cudaEvent_t ev[4] = {};

    void free_plan_device() {
        pool_free(d_tptr); pool_free(d_trow); pool_free(d_hsrc); pool_free(d_tprod); pool_free(d_Sx);
        d_tptr = nullptr; d_trow = d_hsrc = nullptr; d_tprod = d_Sx = nullptr;
        if (chol) { chol_device_destroy(chol); chol = nullptr; }
    }
    ~b200s_kkt() {
        if (uploaded || chol) cudaSetDevice(device);
        if (chol) cudaStreamSynchronize((cudaStream_t)chol_device_stream(chol));
        free_plan_device();
        pool_free(d_Gp); pool_free(d_Grp); pool_free(d_Ap); pool_free(d_Gi); pool_free(d_Gcj); pool_free(d_Ai); pool_free(d_flag);
        pool_free(d_Gx); pool_free(d_Grx); pool_free(d_Ax); pool_free(d_Hx); pool_free(d_di); pool_free(d_x); pool_free(d_y);
        pool_free(d_z); pool_free(d_w); pool_free(d_At); pool_free(d_K); pool_free(d_t);
        for (auto& e : ev) if (e) cudaEventDestroy(e);
    }
};

namespace {

template <class T> int up(T** dst, const std::vector<T>& src, size_t min_count = 1) {
    CUDA_TRY(pool_malloc((void**)dst, std::max(src.size(), min_count) * sizeof(T)));
    if (!src.empty()) CUDA_TRY(cudaMemcpy(*dst, src.data(), src.size() * sizeof(T), cudaMemcpyHostToDevice));
    return ST_OK;
}

void to_ccs(i64 ncols, const b200s_int* cp, const b200s_int* ri, const double* vx, i64 nrows, Ccs& M, const char* what) {
    M.p.assign(ncols + 1, 0);
    if (ncols == 0) return;
    if (!cp) {
        if (nrows == 0) return;          // 0 x n matrix (no equality constraints): kvxopt passes it on the IPM path
        throw std::invalid_argument(std::string(what) + ": null column pointers");
    }
    const i64 nnz = cp[ncols];
    if (cp[0] != 0 || nnz < 0 || (nnz > 0 && (!ri || !vx))) throw std::invalid_argument(std::string(what) + ": bad CCS arrays");
    M.i.resize(nnz); M.x.resize(nnz);
    for (i64 j = 0; j < ncols; j++) {
        if (cp[j + 1] < cp[j]) throw std::invalid_argument(std::string(what) + ": column pointers not monotone");
        M.p[j + 1] = cp[j + 1];
        for (i64 q = cp[j]; q < cp[j + 1]; q++) {
            if (ri[q] < 0 || ri[q] >= nrows) throw std::invalid_argument(std::string(what) + ": row index out of range");
            if (q > cp[j] && ri[q] <= ri[q - 1]) throw std::invalid_argument(std::string(what) + ": row indices not sorted");
            M.i[q] = (int)ri[q]; M.x[q] = vx[q];
        }
    }
}
void transpose(const Ccs& M, i64 nrows, i64 ncols, Ccs& T) {
    T.p.assign(nrows + 1, 0);
    T.i.resize(M.i.size()); T.x.resize(M.x.size());
    for (int r : M.i) T.p[r + 1]++;
    for (i64 r = 0; r < nrows; r++) T.p[r + 1] += T.p[r];
    std::vector<long long> pos(T.p.begin(), T.p.end() - 1);
    for (i64 j = 0; j < ncols; j++)
        for (long long q = M.p[j]; q < M.p[j + 1]; q++) { const long long t = pos[M.i[q]]++; T.i[t] = (int)j; T.x[t] = M.x[q]; }
}

// pattern of tril(H + G'G [+ A'A]) and, for every stored entry, its constant term list
void build_plan(b200s_kkt& K) {
    const i64 n = K.n;
    Ccs Ar;
    if (K.withA) transpose(K.A, K.p, n, Ar);
    K.Sp.assign(n + 1, 0); K.Si.clear(); K.tptr.assign(1, 0); K.trow.clear(); K.tprod.clear(); K.hsrc.clear();
    std::vector<int> pos(n, -1), rows;
    struct Term { int e, row; double prod; };
    std::vector<Term> terms;
    std::vector<int> hs;
    for (i64 j = 0; j < n; j++) {
        rows.clear(); terms.clear();
        auto touch = [&](int i) { if (pos[i] < 0) { pos[i] = 1; rows.push_back(i); } };
        if (K.hasH)
            for (i64 q = K.Hp[j]; q < K.Hp[j + 1]; q++) if (K.Hi[q] >= j) touch((int)K.Hi[q]);
        for (long long q = K.G.p[j]; q < K.G.p[j + 1]; q++) {
            const int k = K.G.i[q];
            for (long long r = K.Gr.p[k]; r < K.Gr.p[k + 1]; r++) if (K.Gr.i[r] >= j) touch(K.Gr.i[r]);
        }
        if (K.withA)
            for (long long q = K.A.p[j]; q < K.A.p[j + 1]; q++) {
                const int a = K.A.i[q];
                for (long long r = Ar.p[a]; r < Ar.p[a + 1]; r++) if (Ar.i[r] >= j) touch(Ar.i[r]);
            }
        std::sort(rows.begin(), rows.end());
        const i64 e0 = (i64)K.Si.size();
        for (size_t t = 0; t < rows.size(); t++) { pos[rows[t]] = (int)t; K.Si.push_back(rows[t]); }
        hs.assign(rows.size(), -1);
        if (K.hasH)
            for (i64 q = K.Hp[j]; q < K.Hp[j + 1]; q++) if (K.Hi[q] >= j) hs[pos[K.Hi[q]]] = (int)q;
        for (long long q = K.G.p[j]; q < K.G.p[j + 1]; q++) {
            const int k = K.G.i[q];
            const double gkj = K.G.x[q];
            for (long long r = K.Gr.p[k]; r < K.Gr.p[k + 1]; r++)
                if (K.Gr.i[r] >= j) terms.push_back({pos[K.Gr.i[r]], k, K.Gr.x[r] * gkj});
        }
        if (K.withA)
            for (long long q = K.A.p[j]; q < K.A.p[j + 1]; q++) {
                const int a = K.A.i[q];
                for (long long r = Ar.p[a]; r < Ar.p[a + 1]; r++)
                    if (Ar.i[r] >= j) terms.push_back({pos[Ar.i[r]], -1, Ar.x[r] * K.A.x[q]});
            }
        // terms grouped by entry, inside an entry ordered by constraint row (fixed summation order)
        std::stable_sort(terms.begin(), terms.end(), [](const Term& a, const Term& b) { return a.e < b.e; });
        size_t t = 0;
        for (size_t e = 0; e < rows.size(); e++) {
            while (t < terms.size() && terms[t].e == (int)e) { K.trow.push_back(terms[t].row); K.tprod.push_back(terms[t].prod); t++; }
            K.tptr.push_back((long long)K.trow.size());
            K.hsrc.push_back(hs[e]);
        }
        (void)e0;
        for (int i : rows) pos[i] = -1;
        K.Sp[j + 1] = (i64)K.Si.size();
    }
}

int upload_static(b200s_kkt& K) {
    int rc;
    CUDA_TRY(cudaSetDevice(K.device));
    if ((rc = up(&K.d_Gp, K.G.p))) return rc;
    if ((rc = up(&K.d_Gi, K.G.i))) return rc;
    if ((rc = up(&K.d_Gx, K.G.x))) return rc;
    if ((rc = up(&K.d_Grp, K.Gr.p))) return rc;
    if ((rc = up(&K.d_Gcj, K.Gr.i))) return rc;
    if ((rc = up(&K.d_Grx, K.Gr.x))) return rc;
    if ((rc = up(&K.d_Ap, K.A.p))) return rc;
    if ((rc = up(&K.d_Ai, K.A.i))) return rc;
    if ((rc = up(&K.d_Ax, K.A.x))) return rc;
    CUDA_TRY(pool_malloc((void**)&K.d_Hx, std::max<i64>(K.nnzH, 1) * sizeof(double)));
    CUDA_TRY(pool_malloc((void**)&K.d_di, std::max<i64>(K.ml, 1) * sizeof(double)));
    CUDA_TRY(pool_malloc((void**)&K.d_w, std::max<i64>(K.ml, 1) * sizeof(double)));
    CUDA_TRY(pool_malloc((void**)&K.d_z, std::max<i64>(K.ml, 1) * sizeof(double)));
    CUDA_TRY(pool_malloc((void**)&K.d_x, std::max<i64>(K.n, 1) * sizeof(double)));
    CUDA_TRY(pool_malloc((void**)&K.d_y, std::max<i64>(K.p, 1) * sizeof(double)));
    CUDA_TRY(pool_malloc((void**)&K.d_t, std::max<i64>(K.p, 1) * sizeof(double)));
    CUDA_TRY(pool_malloc((void**)&K.d_K, std::max<i64>(K.p * K.p, 1) * sizeof(double)));
    CUDA_TRY(pool_malloc((void**)&K.d_At, std::max<i64>(K.n * K.p, 1) * sizeof(double)));
    CUDA_TRY(pool_malloc((void**)&K.d_flag, sizeof(int)));
    for (auto& e : K.ev) CUDA_TRY(cudaEventCreate(&e));
    K.uploaded = true;
    return ST_OK;
}

int upload_plan(b200s_kkt& K) {
    int rc;
    CUDA_TRY(cudaSetDevice(K.device));
    K.free_plan_device();
    if ((rc = up(&K.d_tptr, K.tptr))) return rc;
    if ((rc = up(&K.d_trow, K.trow))) return rc;
    if ((rc = up(&K.d_tprod, K.tprod))) return rc;
    if ((rc = up(&K.d_hsrc, K.hsrc))) return rc;
    CUDA_TRY(pool_malloc((void**)&K.d_Sx, std::max<size_t>(K.Si.size(), 1) * sizeof(double)));
    int st = ST_OK;
    K.chol = chol_device_create(K.plan, K.opts, K.device, &st);
    if (!K.chol) return st;
    return ST_OK;
}

int grid_for(long long work) { return (int)std::max<long long>(1, std::min<long long>((work + 255) / 256, 148 * 8)); }

b200s_status analyze(b200s_kkt& K) {
    try {
        build_plan(K);
        K.plan = CholPlan();
        static const i64 zero = 0;
        chol_analyze(K.n, K.n > 0 ? K.Sp.data() : &zero, K.Si.data(), 'L', nullptr, K.opts, K.plan);
    } catch (const std::bad_alloc&) {
        return B200S_OUT_OF_MEMORY;
    } catch (const std::exception& e) {
        set_last_error(e.what());
        return B200S_INVALID;
    }
    return B200S_OK;
}

}  // namespace

extern "C" {

b200s_status b200s_kkt_create(b200s_int n, b200s_int ml, b200s_int p, const b200s_int* Gp, const b200s_int* Gi, const double* Gx,
                              const b200s_int* Ap, const b200s_int* Ai, const double* Ax, const b200s_int* Hp, const b200s_int* Hi,
                              b200s_kkt** out) {
    if (!out) return B200S_INVALID;
    *out = nullptr;
    if (n < 0 || ml < 0 || p < 0 || n > 0x7ffffff0 || ml > 0x7ffffff0) return B200S_INVALID;
    if (p > 2048 || (double)p * (double)n > 2e9) { set_last_error("kkt: too many equality constraints for the dense A S^-1 A' block"); return B200S_TOO_LARGE; }
    b200s_kkt* K = new (std::nothrow) b200s_kkt();
    if (!K) return B200S_OUT_OF_MEMORY;
    K->n = n; K->ml = ml; K->p = p; K->device = current_device();
    try {
        to_ccs(n, Gp, Gi, Gx, ml, K->G, "G");
        transpose(K->G, ml, n, K->Gr);
        to_ccs(n, Ap, Ai, Ax, p, K->A, "A");
        if (Hp) {
            K->hasH = true;
            K->Hp.assign(Hp, Hp + n + 1);
            K->nnzH = n > 0 ? Hp[n] : 0;
            if (K->nnzH < 0 || (K->nnzH > 0 && !Hi)) throw std::invalid_argument("H: bad CCS arrays");
            K->Hi.assign(Hi, Hi + K->nnzH);
            for (i64 j = 0; j < n; j++)
                for (i64 q = Hp[j]; q < Hp[j + 1]; q++)
                    if (Hi[q] < 0 || Hi[q] >= n) throw std::invalid_argument("H: row index out of range");
        }
    } catch (const std::bad_alloc&) {
        delete K;
        return B200S_OUT_OF_MEMORY;
    } catch (const std::exception& e) {
        set_last_error(e.what());
        delete K;
        return B200S_INVALID;
    }
    b200s_status st = analyze(*K);
    if (st != B200S_OK) { delete K; return st; }
    *out = K;
    return B200S_OK;
}

void b200s_kkt_free(b200s_kkt* K) { delete K; }

/* misc.py:1427-1447: when the first S is not positive definite the reference switches, for the rest of the solve, to
 * S = H + G'D G + A'A; the pattern changes, so the symbolic analysis is redone once. */
b200s_status b200s_kkt_set_singular(b200s_kkt* K, int on) {
    if (!K) return B200S_INVALID;
    if ((on != 0) == K->withA) return B200S_OK;
    K->withA = on != 0;
    K->factored = false;
    if (K->chol || K->d_tptr) { cudaSetDevice(K->device); K->free_plan_device(); }
    return analyze(*K);
}

static int kkt_factor_impl(b200s_kkt* K, const double* di, const double* Hx, b200s_int* minor_out) {
    B200S_NVTX("kkt_factor_impl");
    if (!K || (K->ml > 0 && !di) || (K->hasH && K->nnzH > 0 && !Hx)) return ST_INVALID;
    K->factored = false;
    if (device_count() <= 0) { set_last_error("no CUDA device available"); return ST_NO_DEVICE; }
    int rc;
    if (!K->uploaded && (rc = upload_static(*K))) return rc;
    if (!K->chol && K->n > 0 && (rc = upload_plan(*K))) return rc;
    if (K->n == 0) { K->factored = true; if (minor_out) *minor_out = 0; return ST_OK; }
    CUDA_TRY(cudaSetDevice(K->device));
    cudaStream_t st = (cudaStream_t)chol_device_stream(K->chol);
    CUDA_TRY(cudaEventRecord(K->ev[0], st));
    if (K->ml) CUDA_TRY(cudaMemcpyAsync(K->d_di, di, K->ml * sizeof(double), cudaMemcpyHostToDevice, st));
    if (K->nnzH) CUDA_TRY(cudaMemcpyAsync(K->d_Hx, Hx, K->nnzH * sizeof(double), cudaMemcpyHostToDevice, st));
    const long long nnzS = (long long)K->Si.size();
    k_kkt_assemble<<<grid_for(nnzS), 256, 0, st>>>(nnzS, K->d_tptr, K->d_trow, K->d_tprod, K->d_hsrc, K->d_Hx, K->d_di, K->d_Sx);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaEventRecord(K->ev[1], st));
    i64 minor = K->n;
    rc = chol_device_factorize(K->chol, K->d_Sx, true, &minor, &K->times);
    if (minor_out) *minor_out = minor;
    if (rc != ST_OK) return rc;
    if (K->p > 0) {
        const int p = (int)K->p;
        CUDA_TRY(cudaMemsetAsync(K->d_At, 0, (size_t)K->n * p * sizeof(double), st));
        k_kkt_fill_at<<<grid_for(K->n), 256, 0, st>>>(K->n, K->d_Ap, K->d_Ai, K->d_Ax, K->d_At);
        if ((rc = chol_device_solve_async(K->chol, 9, K->d_At, p, K->n))) return rc;     // Asct = L^-1 P A'
        k_kkt_dots<<<dim3(p, p), 256, 0, st>>>(K->n, p, K->d_At, K->d_At, (int)K->n, K->d_K, 1);
        CUDA_TRY(cudaMemsetAsync(K->d_flag, 0, sizeof(int), st));
        k_kkt_potrf<<<1, 256, 0, st>>>(p, K->d_K, K->d_flag);
        int flag = 0;
        CUDA_TRY(cudaMemcpyAsync(&flag, K->d_flag, sizeof(int), cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaStreamSynchronize(st));
        if (flag) { if (minor_out) *minor_out = flag - 1; set_last_error("kkt: A S^-1 A' is not positive definite"); return ST_NOT_POSDEF; }
    }
    CUDA_TRY(cudaEventRecord(K->ev[2], st));
    CUDA_TRY(cudaStreamSynchronize(st));
    float ms;
    cudaEventElapsedTime(&ms, K->ev[0], K->ev[1]); K->ms_assemble = ms;
    cudaEventElapsedTime(&ms, K->ev[1], K->ev[2]); K->ms_factor = ms;
    K->factored = true;
    return ST_OK;
}

static int kkt_solve_impl(b200s_kkt* K, double* x, double* y, double* z) {
    B200S_NVTX("kkt_solve_impl");
    if (!K || (K->n > 0 && !x) || (K->p > 0 && !y) || (K->ml > 0 && !z)) return ST_INVALID;
    if (!K->factored) { set_last_error("kkt: solve called before a successful factor"); return ST_INVALID; }
    if (K->n == 0) return ST_OK;
    CUDA_TRY(cudaSetDevice(K->device));
    cudaStream_t st = (cudaStream_t)chol_device_stream(K->chol);
    const long long n = K->n, ml = K->ml;
    const int p = (int)K->p;
    int rc;
    CUDA_TRY(cudaEventRecord(K->ev[0], st));
    CUDA_TRY(cudaMemcpyAsync(K->d_x, x, n * sizeof(double), cudaMemcpyHostToDevice, st));
    if (p) CUDA_TRY(cudaMemcpyAsync(K->d_y, y, p * sizeof(double), cudaMemcpyHostToDevice, st));
    if (ml) {
        CUDA_TRY(cudaMemcpyAsync(K->d_z, z, ml * sizeof(double), cudaMemcpyHostToDevice, st));
        k_kkt_scale_z<<<grid_for(ml), 256, 0, st>>>(ml, K->d_di, K->d_z, K->d_w);
        k_kkt_gemv_t<<<grid_for(n), 256, 0, st>>>(n, K->d_Gp, K->d_Gi, K->d_Gx, K->d_w, K->d_x);
    }
    if (K->withA && p) k_kkt_gemv_t<<<grid_for(n), 256, 0, st>>>(n, K->d_Ap, K->d_Ai, K->d_Ax, K->d_y, K->d_x);
    if ((rc = chol_device_solve_async(K->chol, 9, K->d_x, 1, n))) return rc;
    if (p) {
        k_kkt_dots<<<dim3(p, 1), 256, 0, st>>>(n, p, K->d_At, K->d_x, (int)n, K->d_t, 0);
        k_kkt_ksolve<<<1, 32, p * sizeof(double), st>>>(p, K->d_K, K->d_t, K->d_y);
        k_kkt_axpy_cols<<<grid_for(n), 256, 0, st>>>(n, p, K->d_At, K->d_y, K->d_x);
    }
    if ((rc = chol_device_solve_async(K->chol, 10, K->d_x, 1, n))) return rc;
    if (ml) k_kkt_gz<<<grid_for(ml), 256, 0, st>>>(ml, K->d_Grp, K->d_Gcj, K->d_Grx, K->d_di, K->d_x, K->d_z);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaMemcpyAsync(x, K->d_x, n * sizeof(double), cudaMemcpyDeviceToHost, st));
    if (p) CUDA_TRY(cudaMemcpyAsync(y, K->d_y, p * sizeof(double), cudaMemcpyDeviceToHost, st));
    if (ml) CUDA_TRY(cudaMemcpyAsync(z, K->d_z, ml * sizeof(double), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaEventRecord(K->ev[1], st));
    CUDA_TRY(cudaStreamSynchronize(st));
    float ms;
    cudaEventElapsedTime(&ms, K->ev[0], K->ev[1]); K->ms_solve = ms;
    return ST_OK;
}

b200s_status b200s_kkt_factor(b200s_kkt* K, const double* di, const double* Hx, b200s_int* minor_out) {
    return (b200s_status)kkt_factor_impl(K, di, Hx, minor_out);
}
b200s_status b200s_kkt_solve(b200s_kkt* K, double* x, double* y, double* z) { return (b200s_status)kkt_solve_impl(K, x, y, z); }

/* Pattern of S; b200s_kkt_plan_check_host evaluates the term lists on the host so that the CPU test-suite can verify the
 * host-built plan against H + G'D^2G without a GPU.  It is not a solve path and nothing in the Python mirror calls it. */
b200s_status b200s_kkt_info(const b200s_kkt* K, b200s_kkt_info_t* info) {
    if (!K || !info) return B200S_INVALID;
    info->n = K->n; info->ml = K->ml; info->p = K->p; info->nnz_S = (b200s_int)K->Si.size(); info->nterms = (b200s_int)K->trow.size();
    info->nnz_L = K->plan.nnzL; info->flops = K->plan.flops; info->singular_mode = K->withA ? 1 : 0;
    info->ms_assemble = K->ms_assemble; info->ms_factor = K->ms_factor; info->ms_solve = K->ms_solve;
    info->Sp = K->Sp.data(); info->Si = K->Si.data();
    return B200S_OK;
}
b200s_status b200s_kkt_plan_check_host(const b200s_kkt* K, const double* di, const double* Hx, double* Sx) {
    if (!K || !Sx) return B200S_INVALID;
    for (size_t e = 0; e < K->Si.size(); e++) {
        double v = K->hsrc[e] >= 0 ? Hx[K->hsrc[e]] : 0.0;
        for (long long t = K->tptr[e]; t < K->tptr[e + 1]; t++) {
            const double d = K->trow[t] >= 0 ? di[K->trow[t]] : 1.0;
            v = std::fma(d * d, K->tprod[t], v);
        }
        Sx[e] = v;
    }
    return B200S_OK;
}

}  // extern "C"


/* ---- value assembler: y = M w on the device, M a fixed sparse matrix in CSR (include/b200sparse.h: b200s_spmv_*) ---------------
 * The KKT plug-ins whose matrix is assembled outside this file (kkt.ldl2: K = [[H + G' D^2 G, A'], [A, 0]]) write every stored
 * entry of K as a fixed linear combination of the numbers that change per interior-point iteration (w = [d^2; H values; A
 * values]); the combination is evaluated here, one thread per entry, terms in list order (deterministic sums), and the result
 * stays in device memory for b200s_chol_factorize_dev. */
namespace {
__global__ void k_spmv_rows(long long nrows, const long long* __restrict__ rowptr, const int* __restrict__ colind,
                            const double* __restrict__ val, const double* __restrict__ w, double* __restrict__ y) {
    for (long long r = blockIdx.x * (long long)blockDim.x + threadIdx.x; r < nrows; r += (long long)gridDim.x * blockDim.x) {
        double v = 0.0;
        for (long long t = rowptr[r]; t < rowptr[r + 1]; t++) v = fma(val[t], w[colind[t]], v);
        y[r] = v;
    }
}
}  // namespace

struct b200s_spmv {
    i64 nrows = 0, ncols = 0, nnz = 0;
    int device = 0;
    cudaStream_t stream = nullptr;
    long long* d_rowptr = nullptr;
    int* d_colind = nullptr;
    double *d_val = nullptr, *d_w = nullptr, *d_y = nullptr;
    ~b200s_spmv() {
        cudaSetDevice(device);
        if (stream) { cudaStreamSynchronize(stream); cudaStreamDestroy(stream); }
        pool_free(d_rowptr); pool_free(d_colind); pool_free(d_val); pool_free(d_w); pool_free(d_y);
    }
};

static int spmv_create_impl(b200s_spmv& M, const b200s_int* rowptr, const b200s_int* colind, const double* val) {
    CUDA_TRY(cudaSetDevice(M.device));
    CUDA_TRY(cudaStreamCreateWithFlags(&M.stream, cudaStreamNonBlocking));
    std::vector<long long> rp(rowptr, rowptr + M.nrows + 1);
    std::vector<int> ci((size_t)M.nnz);
    for (i64 t = 0; t < M.nnz; t++) ci[t] = (int)colind[t];
    std::vector<double> vx(val, val + M.nnz);
    int rc;
    if ((rc = up(&M.d_rowptr, rp))) return rc;
    if ((rc = up(&M.d_colind, ci))) return rc;
    if ((rc = up(&M.d_val, vx))) return rc;
    CUDA_TRY(pool_malloc((void**)&M.d_w, std::max<i64>(M.ncols, 1) * sizeof(double)));
    CUDA_TRY(pool_malloc((void**)&M.d_y, std::max<i64>(M.nrows, 1) * sizeof(double)));
    return ST_OK;
}

b200s_status b200s_spmv_create(b200s_int nrows, b200s_int ncols, const b200s_int* rowptr, const b200s_int* colind, const double* val,
                               b200s_spmv** out) {
    if (!out) return B200S_INVALID;
    *out = nullptr;
    if (nrows < 0 || ncols < 0 || nrows > 0x7fffffff - 16 || ncols > 0x7fffffff - 16 || !rowptr || rowptr[0] != 0) return B200S_INVALID;
    const i64 nnz = rowptr[nrows];
    if (nnz < 0 || (nnz > 0 && (!colind || !val))) return B200S_INVALID;
    for (i64 r = 0; r < nrows; r++) if (rowptr[r + 1] < rowptr[r]) { set_last_error("spmv: row pointers must be nondecreasing"); return B200S_INVALID; }
    for (i64 t = 0; t < nnz; t++) if (colind[t] < 0 || colind[t] >= ncols) { set_last_error("spmv: column index out of range"); return B200S_INVALID; }
    if (device_count() <= 0) return B200S_NO_DEVICE;
    b200s_spmv* M = new (std::nothrow) b200s_spmv();
    if (!M) return B200S_OUT_OF_MEMORY;
    M->nrows = nrows; M->ncols = ncols; M->nnz = nnz; M->device = current_device();
    int rc;
    try { rc = spmv_create_impl(*M, rowptr, colind, val); }
    catch (const std::bad_alloc&) { rc = ST_OOM; }
    if (rc != ST_OK) { delete M; return (b200s_status)rc; }
    *out = M;
    return B200S_OK;
}

static int spmv_apply_impl(b200s_spmv* M, const double* w_host, double** y_dev) {
    B200S_NVTX("spmv_apply");
    CUDA_TRY(cudaSetDevice(M->device));
    if (M->ncols) CUDA_TRY(cudaMemcpyAsync(M->d_w, w_host, M->ncols * sizeof(double), cudaMemcpyHostToDevice, M->stream));
    if (M->nrows) {
        const int grid = (int)std::max<i64>(1, std::min<i64>((M->nrows + 255) / 256, 148 * 8));
        k_spmv_rows<<<grid, 256, 0, M->stream>>>(M->nrows, M->d_rowptr, M->d_colind, M->d_val, M->d_w, M->d_y);
        CUDA_TRY(cudaGetLastError());
    }
    CUDA_TRY(cudaStreamSynchronize(M->stream));
    *y_dev = M->d_y;
    return ST_OK;
}
b200s_status b200s_spmv_apply(b200s_spmv* M, const double* w_host, double** y_dev_out) {
    if (!M || !y_dev_out || (M->ncols > 0 && !w_host)) return B200S_INVALID;
    return (b200s_status)spmv_apply_impl(M, w_host, y_dev_out);
}
b200s_status b200s_spmv_get(b200s_spmv* M, double* y_host) {
    if (!M || (M->nrows > 0 && !y_host)) return B200S_INVALID;
    if (cudaSetDevice(M->device) != cudaSuccess) return B200S_CUDA_ERROR;
    if (M->nrows && cudaMemcpy(y_host, M->d_y, M->nrows * sizeof(double), cudaMemcpyDeviceToHost) != cudaSuccess) return B200S_CUDA_ERROR;
    return B200S_OK;
}
void b200s_spmv_free(b200s_spmv* M) { delete M; }
