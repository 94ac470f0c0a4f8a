// FP64 peak microbenchmark for B200 (sm_100a): DMMA (mma.sync f64) shapes, DFMA, HBM copy.
// Output: one JSON object on stdout.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_peak fp64_peak.cu
#include <cstdio>
#include <cuda_runtime.h>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("CUDA error %s at %d\n",cudaGetErrorString(e),__LINE__);return 1;}}while(0)

template<int NACC>
__global__ void k_dmma884(double* out, int iters, double a0, double b0) {
  double a = a0 + threadIdx.x * 1e-9, b = b0;
  double c[NACC][2];
  #pragma unroll
  for (int i = 0; i < NACC; i++) { c[i][0] = 0; c[i][1] = 0; }
  for (int it = 0; it < iters; it++) {
    #pragma unroll
    for (int i = 0; i < NACC; i++)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                   : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
  }
  double s = 0;
  #pragma unroll
  for (int i = 0; i < NACC; i++) s += c[i][0] + c[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template<int NACC>
__global__ void k_dmma16816(double* out, int iters, double a0, double b0) {
  double a[8], b[4];
  #pragma unroll
  for (int i = 0; i < 8; i++) a[i] = a0 + i + threadIdx.x * 1e-9;
  #pragma unroll
  for (int i = 0; i < 4; i++) b[i] = b0 + i;
  double c[NACC][4];
  #pragma unroll
  for (int i = 0; i < NACC; i++) { c[i][0] = 0; c[i][1] = 0; c[i][2] = 0; c[i][3] = 0; }
  for (int it = 0; it < iters; it++) {
    #pragma unroll
    for (int i = 0; i < NACC; i++)
      asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};"
                   : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3])
                   : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]),
                     "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
  }
  double s = 0;
  #pragma unroll
  for (int i = 0; i < NACC; i++) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template<int NACC>
__global__ void k_dmma1684(double* out, int iters, double a0, double b0) {
  double a[2], b[1];
  a[0] = a0 + threadIdx.x * 1e-9; a[1] = a0 + 1; b[0] = b0;
  double c[NACC][4];
  #pragma unroll
  for (int i = 0; i < NACC; i++) { c[i][0] = 0; c[i][1] = 0; c[i][2] = 0; c[i][3] = 0; }
  for (int it = 0; it < iters; it++) {
    #pragma unroll
    for (int i = 0; i < NACC; i++)
      asm volatile("mma.sync.aligned.m16n8k4.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
                   : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3])
                   : "d"(a[0]), "d"(a[1]), "d"(b[0]));
  }
  double s = 0;
  #pragma unroll
  for (int i = 0; i < NACC; i++) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template<int NACC>
__global__ void k_dfma(double* out, int iters, double a0, double b0) {
  double a = a0 + threadIdx.x * 1e-9, b = b0;
  double c[NACC];
  #pragma unroll
  for (int i = 0; i < NACC; i++) c[i] = i;
  for (int it = 0; it < iters; it++) {
    #pragma unroll
    for (int i = 0; i < NACC; i++) c[i] = fma(c[i], a, b);
  }
  double s = 0;
  #pragma unroll
  for (int i = 0; i < NACC; i++) s += c[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_copy(const double4* __restrict__ a, double4* __restrict__ b, size_t n) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x, st = (size_t)gridDim.x * blockDim.x;
  for (; i < n; i += st) b[i] = a[i];
}
template<class F> float timeit(F f, int reps) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  f(); cudaDeviceSynchronize();
  float best = 1e30f;
  for (int r = 0; r < reps; r++) { cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms; }
  return best;
}
int main() {
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  int sm = p.multiProcessorCount;
  double* out; CK(cudaMalloc(&out, sizeof(double) * sm * 8 * 1024));
  int iters = 4096;
  printf("{\"gpu\": \"%s\", \"sms\": %d", p.name, sm);
  for (int wpb : {4, 8, 16}) {
    int thr = wpb * 32, blocks = sm * (wpb <= 8 ? 2 : 1);
    double warps = (double)blocks * wpb;
    float ms;
    ms = timeit([&] { k_dmma884<8><<<blocks, thr>>>(out, iters, 1.0, 1e-3); }, 5);
    printf(", \"dmma_m8n8k4_w%d_tflops\": %.2f", wpb, warps * iters * 8 * (2.0 * 8 * 8 * 4) / (ms * 1e-3) / 1e12);
    ms = timeit([&] { k_dmma1684<8><<<blocks, thr>>>(out, iters, 1.0, 1e-3); }, 5);
    printf(", \"dmma_m16n8k4_w%d_tflops\": %.2f", wpb, warps * iters * 8 * (2.0 * 16 * 8 * 4) / (ms * 1e-3) / 1e12);
    ms = timeit([&] { k_dmma16816<8><<<blocks, thr>>>(out, iters, 1.0, 1e-3); }, 5);
    printf(", \"dmma_m16n8k16_w%d_tflops\": %.2f", wpb, warps * iters * 8 * (2.0 * 16 * 8 * 16) / (ms * 1e-3) / 1e12);
    ms = timeit([&] { k_dfma<16><<<blocks, thr>>>(out, iters * 4, 1.0000001, 1e-3); }, 5);
    printf(", \"dfma_w%d_tflops\": %.2f", wpb, warps * 32 * iters * 4.0 * 16 * 2.0 / (ms * 1e-3) / 1e12);
  }
  CK(cudaGetLastError());
  size_t n = (size_t)1 << 30;  // bytes per buffer
  double4 *a, *b; CK(cudaMalloc(&a, n)); CK(cudaMalloc(&b, n)); CK(cudaMemset(a, 1, n));
  float ms = timeit([&] { k_copy<<<sm * 16, 512>>>(a, b, n / sizeof(double4)); }, 5);
  printf(", \"copy_gbs\": %.1f", 2.0 * n / (ms * 1e-3) / 1e9);
  ms = timeit([&] { cudaMemcpyAsync(b, a, n, cudaMemcpyDeviceToDevice); }, 5);
  printf(", \"memcpy_d2d_gbs\": %.1f", 2.0 * n / (ms * 1e-3) / 1e9);
  CK(cudaDeviceSynchronize());
  printf("}\n");
  return 0;
}
