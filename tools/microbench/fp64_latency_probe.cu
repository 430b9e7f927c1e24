// fp64_latency_probe.cu -- dependent-issue latency of FP64 instructions on B200: one warp per scheduler runs K independent chains of
// DFMA (or DADD); cycles per instruction against K tells the latency (K = 1) and how much instruction-level parallelism a lone warp
// needs to keep the FP64 pipe (one warp instruction per 2 cycles per scheduler) busy.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_latency_probe fp64_latency_probe.cu
#include <cstdio>
#include <cuda_runtime.h>
template <int K, bool ADD>
__global__ void __launch_bounds__(128, 1) probe(double* out, long long* cyc, int iters, double m, double c) {
  double a[K];
#pragma unroll
  for (int k = 0; k < K; k++) a[k] = (double)(threadIdx.x + k);
  const long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int r = 0; r < 16; r++) {
#pragma unroll
      for (int k = 0; k < K; k++) a[k] = ADD ? a[k] + c : __fma_rn(a[k], m, c);
    }
  }
  const long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int k = 0; k < K; k++) s += a[k];
  out[blockIdx.x * 128 + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int K, bool ADD>
void run() {
  double* out; long long* cyc;
  cudaMalloc(&out, 148 * 128 * 8); cudaMalloc(&cyc, 148 * 8);
  const int iters = 2000;
  probe<K, ADD><<<148, 128>>>(out, cyc, iters, 1.0000001, 0.5);
  probe<K, ADD><<<148, 128>>>(out, cyc, iters, 1.0000001, 0.5);
  cudaDeviceSynchronize();
  long long h[148];
  cudaMemcpy(h, cyc, sizeof h, cudaMemcpyDeviceToHost);
  printf("%s  %2d independent chains per warp (one warp per scheduler): %.2f cycles per instruction\n", ADD ? "DADD" : "DFMA", K, (double)h[0] / (iters * 16.0 * K));
  cudaFree(out); cudaFree(cyc);
}
int main() {
  run<1, false>(); run<2, false>(); run<4, false>(); run<8, false>(); run<16, false>();
  run<1, true>(); run<2, true>(); run<4, true>(); run<8, true>();
  return 0;
}
