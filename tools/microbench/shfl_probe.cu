// shfl_probe.cu -- SHFL.BFLY throughput with 8 warps per SM (the Fourier MAC exchanges spectra between lanes l and l ^ 16)
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>
__global__ void __launch_bounds__(256, 1) probe(uint32_t* out, long long* cyc, int iters) {
  uint32_t a[16];
#pragma unroll
  for (int k = 0; k < 16; k++) a[k] = threadIdx.x * 17 + k;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int k = 0; k < 16; k++) a[k] = __shfl_xor_sync(0xffffffffu, a[k], 16) + 1;
  }
  const long long t1 = clock64();
  uint32_t s = 0;
#pragma unroll
  for (int k = 0; k < 16; k++) s ^= a[k];
  if (s == 0x12345678u) out[0] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
int main() {
  uint32_t* out; long long* cyc;
  cudaMalloc(&out, 4); cudaMalloc(&cyc, 148 * 8);
  const int iters = 2000;
  probe<<<148, 256>>>(out, cyc, iters);
  probe<<<148, 256>>>(out, cyc, iters);
  cudaDeviceSynchronize();
  long long h[148];
  cudaMemcpy(h, cyc, sizeof h, cudaMemcpyDeviceToHost);
  printf("SHFL.BFLY: %.2f cycles per warp instruction per SM (8 warps)\n", (double)h[0] / (iters * 16.0 * 8.0));
  return 0;
}
