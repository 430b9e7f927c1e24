// ifetch_int_probe.cu -- instruction-supply limit with cheap ALU instructions (1 issue cycle each).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ifetch_int_probe ifetch_int_probe.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int N>
__device__ __forceinline__ void body(unsigned (&a)[8]) {
#pragma unroll
  for (int i = 0; i < N; i++) a[i & 7] = (a[i & 7] ^ (unsigned)(0x9E3779B1u * (i + 1))) + a[(i + 3) & 7];   // LOP3 + IADD3, distinct immediates
}

template <int N>
__global__ void __launch_bounds__(256, 1) probe(unsigned* sink, int iters, long long* cycles, int* ninstr) {
  unsigned a[8];
  for (int k = 0; k < 8; k++) a[k] = threadIdx.x * 7 + k;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; it++) body<N>(a);
  const long long t1 = clock64();
  unsigned s = 0;
  for (int k = 0; k < 8; k++) s += a[k];
  if (s == 12345u) sink[0] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) cycles[0] = t1 - t0;
}

template <int N>
void run(unsigned* sink, long long* d_cyc, int threads) {
  const int iters = 200;
  probe<N><<<148, threads>>>(sink, iters, d_cyc, nullptr);
  probe<N><<<148, threads>>>(sink, iters, d_cyc, nullptr);
  cudaDeviceSynchronize();
  long long c = 0;
  cudaMemcpy(&c, d_cyc, sizeof c, cudaMemcpyDeviceToHost);
  printf("threads %3d  body %5d stmts (2 instr each, ~%3d KB)  cycles per instruction per warp %.3f\n", threads, N, N * 32 / 1024, (double)c / iters / (2.0 * N));
}

int main() {
  unsigned* sink; long long* d_cyc;
  cudaMalloc(&sink, 8); cudaMalloc(&d_cyc, 8);
  for (int threads : {64, 128, 256, 512}) {
    run<256>(sink, d_cyc, threads);
    run<512>(sink, d_cyc, threads);
    run<1024>(sink, d_cyc, threads);
    run<2048>(sink, d_cyc, threads);
    run<3072>(sink, d_cyc, threads);
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
