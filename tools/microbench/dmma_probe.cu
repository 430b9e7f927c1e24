// dmma_probe.cu -- is the FP64 tensor path (mma.sync.m8n8k4.f64, SASS DMMA) a second FP64 resource next to the DFMA pipe on B200?
// 8 warps per SM; flop per cycle per SM for: DFMA only, DMMA only, both interleaved in one instruction stream.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dmma_probe dmma_probe.cu
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma(double (&c)[2], double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c[0]), "+d"(c[1]) : "d"(a), "d"(b));
}
template <int MODE>
__global__ void __launch_bounds__(256, 1) probe(double* sink, long long* cyc, int iters, double seed) {
  double f[8], c[4][2];
#pragma unroll
  for (int k = 0; k < 8; k++) f[k] = seed + threadIdx.x + k;
#pragma unroll
  for (int k = 0; k < 4; k++) { c[k][0] = seed * k; c[k][1] = seed + k; }
  const double a = 1.0 + 1e-9 * threadIdx.x, b = 1.0 - 1e-9 * threadIdx.x;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int r = 0; r < 4; r++) {
      if (MODE != 1) {
#pragma unroll
        for (int k = 0; k < 8; k++) f[k] = __fma_rn(f[k], 1.0000001, 1e-9);
      }
      if (MODE != 0) {
#pragma unroll
        for (int k = 0; k < 4; k++) dmma(c[k], a, b);
      }
    }
  }
  const long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int k = 0; k < 8; k++) s += f[k];
#pragma unroll
  for (int k = 0; k < 4; k++) s += c[k][0] + c[k][1];
  if (s == 12345.678) sink[0] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int MODE>
void run(const char* name) {
  double* sink; long long* cyc;
  cudaMalloc(&sink, 8); cudaMalloc(&cyc, 148 * 8);
  const int iters = 2000;
  probe<MODE><<<148, 256>>>(sink, cyc, iters, 0.5);
  probe<MODE><<<148, 256>>>(sink, cyc, iters, 0.5);
  cudaDeviceSynchronize();
  long long h[148];
  cudaMemcpy(h, cyc, sizeof h, cudaMemcpyDeviceToHost);
  // per iteration and warp: 32 DFMA (32 lanes x 2 flop) and / or 16 DMMA (8 x 8 x 4 x 2 = 512 flop)
  const double flop = (double)iters * 8.0 * ((MODE != 1 ? 32.0 * 64.0 : 0.0) + (MODE != 0 ? 16.0 * 512.0 : 0.0));
  printf("%-14s %8.0f cycles   %.1f flop/cycle/SM\n", name, (double)h[0], flop / (double)h[0]);
}
int main() {
  run<0>("DFMA only");
  run<1>("DMMA only");
  run<2>("DFMA + DMMA");
  return 0;
}
