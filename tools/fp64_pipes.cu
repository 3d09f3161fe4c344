// Microbenchmark: do DFMA and DMMA (mma.sync.m8n8k4.f64) share one execution pipe on sm_100a?
// Runs three kernels with 4 warps per SMSP worth of independent chains: DFMA only, DMMA only, both interleaved.
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma(double (&c)[2], double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c[0]), "+d"(c[1]) : "d"(a), "d"(b));
}
template <int MODE>
__global__ void __launch_bounds__(128) k(double *out, int iters, double x) {
  double f[8], c[4][2];
  for (int i = 0; i < 8; i++) f[i] = threadIdx.x * 1e-9 + i;
  for (int i = 0; i < 4; i++) c[i][0] = c[i][1] = 0.0;
  for (int it = 0; it < iters; it++) {
    if (MODE & 1) {
#pragma unroll
      for (int r = 0; r < 4; r++)
#pragma unroll
        for (int i = 0; i < 8; i++) f[i] = fma(f[i], x, 1e-3);   // 32 DFMA
    }
    if (MODE & 2) {
#pragma unroll
      for (int i = 0; i < 4; i++) dmma(c[i], f[i], x);           // 4 DMMA = 64 pipe cycles if 16 each
    }
  }
  double s = 0;
  for (int i = 0; i < 8; i++) s += f[i];
  for (int i = 0; i < 4; i++) s += c[i][0] + c[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
  double *out; cudaMalloc(&out, 148 * 4 * 128 * 8);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int iters = 20000;
  for (int mode = 1; mode <= 3; mode++) {
    for (int rep = 0; rep < 2; rep++) {
      cudaEventRecord(e0);
      if (mode == 1) k<1><<<148 * 4, 128>>>(out, iters, 0.999);
      if (mode == 2) k<2><<<148 * 4, 128>>>(out, iters, 0.999);
      if (mode == 3) k<3><<<148 * 4, 128>>>(out, iters, 0.999);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      if (rep) printf("mode %d (%s): %.3f ms; per iteration per SMSP: %.1f cycles at 1.965 GHz (4 warps x %d DFMA + %d DMMA)\n", mode,
                      mode == 1 ? "DFMA" : mode == 2 ? "DMMA" : "both", ms, ms * 1e-3 * 1.965e9 / iters, (mode & 1) ? 32 : 0, (mode & 2) ? 4 : 0);
    }
  }
  return 0;
}
