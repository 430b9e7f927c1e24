// smem_width_probe.cu -- shared-memory throughput per access width with 8 warps per SM, conflict-free patterns:
// bytes per cycle per SM for LDS.32 / LDS.64 / LDS.128 and STS.32 / STS.64 / STS.128.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o smem_width_probe smem_width_probe.cu
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>
template <int W, bool ST>
__global__ void __launch_bounds__(256, 1) probe(uint32_t* out, long long* cyc, int iters) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  unsigned char* base = smem + warp * 8192 + lane * W;   // lane-consecutive: conflict free at every width
  uint32_t acc = threadIdx.x;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int g = 0; g < 16; g++) {
      unsigned char* p = base + g * (32 * W);
      if (W == 4) {
        if (ST) asm volatile("st.shared.u32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(p)), "r"(acc) : "memory");
        else { uint32_t v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"((uint32_t)__cvta_generic_to_shared(p)) : "memory"); acc ^= v; }
      } else if (W == 8) {
        if (ST) asm volatile("st.shared.v2.u32 [%0], {%1, %1};" ::"r"((uint32_t)__cvta_generic_to_shared(p)), "r"(acc) : "memory");
        else { uint32_t v, w2; asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v), "=r"(w2) : "r"((uint32_t)__cvta_generic_to_shared(p)) : "memory"); acc ^= v ^ w2; }
      } else {
        if (ST) asm volatile("st.shared.v4.u32 [%0], {%1, %1, %1, %1};" ::"r"((uint32_t)__cvta_generic_to_shared(p)), "r"(acc) : "memory");
        else { uint32_t a, b, c, d; asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(a), "=r"(b), "=r"(c), "=r"(d) : "r"((uint32_t)__cvta_generic_to_shared(p)) : "memory"); acc ^= a ^ b ^ c ^ d; }
      }
    }
  }
  const long long t1 = clock64();
  if (acc == 0x12345678u) out[0] = acc;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int W, bool ST>
void run(const char* name) {
  uint32_t* out; long long* cyc;
  cudaMalloc(&out, 4); cudaMalloc(&cyc, 148 * 8);
  const int iters = 500;
  cudaFuncSetAttribute(probe<W, ST>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
  probe<W, ST><<<148, 256, 65536>>>(out, cyc, iters);
  probe<W, ST><<<148, 256, 65536>>>(out, cyc, iters);
  cudaDeviceSynchronize();
  long long h[148];
  cudaMemcpy(h, cyc, sizeof h, cudaMemcpyDeviceToHost);
  const double bytes = (double)iters * 16 * 256 * W;
  printf("%-8s %6.1f B/cycle/SM   %.2f cycles per warp instruction\n", name, bytes / (double)h[0], (double)h[0] / (iters * 16.0 * 8.0));
}
int main() {
  run<4, false>("LDS.32"); run<8, false>("LDS.64"); run<16, false>("LDS.128");
  run<4, true>("STS.32"); run<8, true>("STS.64"); run<16, true>("STS.128");
  return 0;
}
