// lds_cost_probe.cu -- what does a shared-memory load cost the scheduler that issues it, next to FP64 work?
// 148 CTAs x 256 threads (2 warps per scheduler).  Per iteration and warp: 16 DFMA (constant-bank multiplier) whose addends are
// the loaded values, plus one of: 4 x LDS.128 | 8 x LDS.64 | 16 x LDS.32 (64 bytes per thread each), or 8 x LDS.128 | 16 x LDS.64.
// FP64 alone: 64 cycles per iteration.  Shared-memory pipe alone (8 warps): LDS.128 2.0, LDS.64 1.0, LDS.32 0.5 cycles each.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o lds_cost_probe lds_cost_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__constant__ double kc[8] = {1.0000001, 0.9999999, 1.0000002, 0.9999998, 1.0000003, 0.9999997, 1.0000004, 0.9999996};
template <int W, int NL, int ND>   // W = bytes per load (4, 8, 16), NL loads per iteration
__global__ void __launch_bounds__(256, 1) probe(double* sink, long long* cyc, int iters, double seed) {
  extern __shared__ __align__(16) unsigned char sm[];
  for (int i = threadIdx.x; i < 16384; i += 256) reinterpret_cast<float*>(sm)[i] = 1e-30f * i;
  double a[16];
#pragma unroll
  for (int k = 0; k < 16; k++) a[k] = seed + threadIdx.x + k;
  const uint32_t base = (uint32_t)__cvta_generic_to_shared(sm) + threadIdx.x * W;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
    double v[16];
#pragma unroll
    for (int k = 0; k < 16; k++) v[k] = 0.0;
#pragma unroll
    for (int l = 0; l < NL; l++) {
      const uint32_t ad = base + (uint32_t)(l * 256 * W);
      if (W == 16) asm volatile("ld.volatile.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v[(2 * l) & 15]), "=d"(v[(2 * l + 1) & 15]) : "r"(ad));
      if (W == 8) asm volatile("ld.volatile.shared.f64 %0, [%1];" : "=d"(v[l & 15]) : "r"(ad));
      if (W == 4) { float f; asm volatile("ld.volatile.shared.f32 %0, [%1];" : "=f"(f) : "r"(ad)); v[l & 15] = __hiloint2double(__float_as_int(f), 0); }
    }
#pragma unroll
    for (int k = 0; k < ND; k++) a[k] = __fma_rn(a[k], kc[k & 7], v[k]);
    if (ND == 0) {
#pragma unroll
      for (int k = 0; k < 16; k++) a[k] = v[k];
    }
  }
  const long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int k = 0; k < 16; k++) s += a[k];
  if (s == 12345.678) sink[0] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int W, int NL, int ND>
void run() {
  double* sink; long long* cyc;
  cudaMalloc(&sink, 8); cudaMalloc(&cyc, 148 * 8);
  const int iters = 4000;
  cudaFuncSetAttribute(probe<W, NL, ND>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
  probe<W, NL, ND><<<148, 256, 65536>>>(sink, cyc, iters, 0.5);
  probe<W, NL, ND><<<148, 256, 65536>>>(sink, cyc, iters, 0.5);
  cudaDeviceSynchronize();
  long long h[148];
  cudaMemcpy(h, cyc, sizeof h, cudaMemcpyDeviceToHost);
  printf("%2d x LDS.%-3d + %2d DFMA: %6.1f cycles per iteration   (pipe alone: FP64 %d, shared memory %d)\n", NL, W * 8, ND, (double)h[0] / iters, 4 * ND,
         NL * W);
  cudaFree(sink); cudaFree(cyc);
}
int main() {
  run<16, 0, 16>();
  run<16, 4, 16>(); run<8, 8, 16>(); run<4, 16, 16>();
  run<16, 8, 16>(); run<8, 16, 16>();
  // below the shared-memory bound (loads hidden under the FP64 pipe if they cost the scheduler nothing)
  run<16, 1, 16>(); run<16, 2, 16>(); run<8, 2, 16>(); run<8, 4, 16>(); run<4, 4, 16>(); run<4, 8, 16>();
  run<16, 4, 0>(); run<8, 8, 0>(); run<4, 16, 0>(); run<16, 8, 0>();
  return 0;
}
