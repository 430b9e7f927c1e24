// icache_probe.cu -- how much does a straight-line loop body larger than the instruction cache cost on B200?
// One CTA of 8 warps per SM runs a loop whose body is KB kilobytes of dependent-free DFMA/IADD instructions.
// Mode 0: all warps in lock-step.  Mode 1: odd warps start half a body later (two instruction streams per SM).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o icache_probe icache_probe.cu && ./icache_probe
#include <cstdio>
#include <cuda_runtime.h>

template <int N>
__device__ __forceinline__ void body(double (&a)[8], const double m, const double c) {
#pragma unroll
  for (int i = 0; i < N; i++) a[i & 7] = __fma_rn(a[i & 7], m, c + (double)(i >> 3) * 1e-9);   // distinct constant -> no code folding
}

template <int N>
__global__ void __launch_bounds__(256, 1) probe_unused(double* sink, int iters, int mode, long long* cycles) {
  double a[8];
  for (int k = 0; k < 8; k++) a[k] = threadIdx.x + k;
  const int warp = threadIdx.x >> 5;
  const bool second_half_first = mode == 1 && (warp & 1);
  __syncthreads();
  const long long t0 = clock64();
  if (second_half_first) body<N / 2>(a, 1.0000001, 1e-9);   // de-phase by half a body
  for (int it = 0; it < iters; it++) body<N>(a, 1.0000001, 1e-9);
  const long long t1 = clock64();
  double s = 0;
  for (int k = 0; k < 8; k++) s += a[k];
  if (s == 1234.5) sink[0] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) cycles[0] = t1 - t0;
}

template <int N>
__global__ void __launch_bounds__(256, 1) probe(double* sink, int iters, int mode, long long* cycles) {
  double a[8];
  for (int k = 0; k < 8; k++) a[k] = threadIdx.x + k;
  const int warp = threadIdx.x >> 5;
  const bool second_half_first = mode == 1 && (warp & 1);
  __syncthreads();
  const long long t0 = clock64();
  if (second_half_first) body<N / 2>(a, 1.0000001, 1e-9);
  for (int it = 0; it < iters; it++) body<N>(a, 1.0000001, 1e-9);
  const long long t1 = clock64();
  double s = 0;
  for (int k = 0; k < 8; k++) s += a[k];
  if (s == 1234.5) sink[0] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) cycles[0] = t1 - t0;
}

// lone warps: 64 threads per CTA = one warp on each of two schedulers (the S = 1 case of the blind rotation)
template <int N>
void run_lone(double* sink, long long* d_cyc) {
  const int iters = 200;
  probe<N><<<148, 64>>>(sink, iters, 0, d_cyc);
  probe<N><<<148, 64>>>(sink, iters, 0, d_cyc);
  cudaDeviceSynchronize();
  long long c = 0;
  cudaMemcpy(&c, d_cyc, sizeof c, cudaMemcpyDeviceToHost);
  printf("lone warps: body %4d instr (~%3d KB)  cycles/instr %.3f (ideal 2.0)\n", N, N * 16 / 1024, (double)c / iters / N);
}

template <int N>
void run(double* sink, long long* d_cyc) {
  const int iters = 200;
  const int grids[3] = {37, 74, 148};   // 148 CTAs: both SMs of every TPC are busy
  for (int g = 0; g < 3; g++)
    for (int mode = 0; mode < 2; mode++) {
      probe<N><<<grids[g], 256>>>(sink, iters, mode, d_cyc);
      probe<N><<<grids[g], 256>>>(sink, iters, mode, d_cyc);
      cudaDeviceSynchronize();
      long long c = 0;
      cudaMemcpy(&c, d_cyc, sizeof c, cudaMemcpyDeviceToHost);
      // every instruction is a DFMA (16 B): 2 warps per scheduler, 2 cycles per DFMA -> ideal 4 cycles per instruction per warp
      printf("body %4d instr (~%3d KB)  CTAs %3d  mode %d  cycles/instr %.3f\n", N, N * 16 / 1024, grids[g], mode, (double)c / iters / N);
    }
}

int main() {
  double* sink; long long* d_cyc;
  cudaMalloc(&sink, 8); cudaMalloc(&d_cyc, 8);
  run_lone<512>(sink, d_cyc); run_lone<1024>(sink, d_cyc); run_lone<2048>(sink, d_cyc); run_lone<3072>(sink, d_cyc);
  run_lone<4096>(sink, d_cyc); run_lone<6144>(sink, d_cyc); run_lone<8192>(sink, d_cyc);
  run<512>(sink, d_cyc);
  run<1024>(sink, d_cyc);
  run<2048>(sink, d_cyc);
  run<3072>(sink, d_cyc);
  run<4096>(sink, d_cyc);
  run<6144>(sink, d_cyc);
  run<8192>(sink, d_cyc);
  cudaError_t e = cudaGetLastError();
  printf("%s\n", cudaGetErrorString(e));
  return 0;
}
