// dfma_operand_probe.cu -- FP64 pipe rate per operand pattern, 2 warps per scheduler (256 threads per SM) like the blind rotation:
//   mode 0  a = fma(a, m, c)      two loop-invariant operands (what fb_measure_fp64_peak runs: operand-reuse friendly)
//   mode 1  a = fma(b, c, a)      three distinct, changing register operands per instruction (the Fourier MAC, twiddle products)
//   mode 2  a = fma(b, K, a)      K from the constant bank (the generated butterflies)
//   mode 3  a = a + b             DADD, two register operands
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dfma_operand_probe dfma_operand_probe.cu
#include <cstdio>
#include <cuda_runtime.h>
__constant__ double kc[8] = {1.0000001, 0.9999999, 1.0000002, 0.9999998, 1.0000003, 0.9999997, 1.0000004, 0.9999996};
template <int MODE>
__global__ void __launch_bounds__(256, 1) probe(double* sink, long long* cyc, int iters, double seed) {
  double a[8], b[8], c[8];
#pragma unroll
  for (int k = 0; k < 8; k++) { a[k] = seed + threadIdx.x + k; b[k] = 1.0 + 1e-9 * (threadIdx.x + k); c[k] = 1e-7 * (k + 1) + 1e-12 * threadIdx.x; }
  const double m = 1.0000001 + seed * 1e-12, cc = 1e-9 + seed * 1e-15;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int r = 0; r < 4; r++) {
#pragma unroll
      for (int k = 0; k < 8; k++) {
        if (MODE == 0) a[k] = __fma_rn(a[k], m, cc);
        if (MODE == 1) a[k] = __fma_rn(b[(k + r) & 7], c[(k + 3 * r + 1) & 7], a[k]);
        if (MODE == 2) a[k] = __fma_rn(b[(k + r) & 7], kc[(k + r) & 7], a[k]);
        if (MODE == 3) a[k] = __dadd_rn(a[k], b[(k + r) & 7]);
      }
    }
    if (MODE == 1) {   // keep b and c changing so that nothing is loop invariant
#pragma unroll
      for (int k = 0; k < 8; k++) { b[k] = __fma_rn(a[k], 1e-30, b[k]); }
    }
  }
  const long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int k = 0; k < 8; k++) s += a[k] + b[k] + c[k];
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
  const double n = (double)iters * (32 + (MODE == 1 ? 8 : 0));      // FP64 instructions per warp
  printf("%-28s %.2f cycles per FP64 instruction per warp (2 warps per scheduler; pipe-bound = 4.00)\n", name, (double)h[0] / n);
}
int main() {
  run<0>("fma(a, m, c) invariant m, c");
  run<1>("fma(b, c, a) three registers");
  run<2>("fma(b, K, a) constant bank");
  run<3>("a + b");
  return 0;
}
