// smem_bw_probe.cu -- shared-memory LOAD bandwidth, measured to completion: the closing clock read takes the XOR of every
// loaded word as an input operand, so it cannot issue before the last load has returned (smem_width_probe.cu read the
// clock right after the last load was ISSUED, which for loads measures the issue rate of a deep queue, not the data path).
// Also mixes loads and stores (a Stockham stage: 8 LDS.128 + 8 STS.128 per thread) to see whether they share one pipe.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o smem_bw_probe smem_bw_probe.cu
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>

__device__ __forceinline__ long long clock_after(uint32_t dep) {
  long long t;
  asm volatile("{ .reg .b32 d; mov.b32 d, %1; mov.u64 %0, %%clock64; }" : "=l"(t) : "r"(dep) : "memory");
  return t;
}

// MODE 0: loads only; 1: stores only; 2: 8 loads then 8 stores per group (both)
template <int W, int MODE>
__global__ void __launch_bounds__(256, 1) probe(uint32_t* out, long long* cyc, int iters) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t acc = threadIdx.x;
  for (int i = threadIdx.x; i < 65536 / 4; i += 256) reinterpret_cast<uint32_t*>(smem)[i] = i;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
    // a different 512-byte-aligned window every iteration (8 warps x 8 KiB)
    const uint32_t base = (uint32_t)__cvta_generic_to_shared(smem) + warp * 8192 + lane * W + ((it & 1) ? 4096 : 0);
#pragma unroll
    for (int g = 0; g < 8; g++) {
      const uint32_t p = base + g * (32 * W);
      if (MODE == 0 || MODE == 2) {
        if (W == 4) { uint32_t v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(p) : "memory"); acc ^= v; }
        else if (W == 8) { uint32_t v, w2; asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v), "=r"(w2) : "r"(p) : "memory"); acc ^= v ^ w2; }
        else { uint32_t a, b, c, d; asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(a), "=r"(b), "=r"(c), "=r"(d) : "r"(p) : "memory"); acc ^= a ^ b ^ c ^ d; }
      }
    }
    if (MODE == 1 || MODE == 2) {
#pragma unroll
      for (int g = 0; g < 8; g++) {
        const uint32_t p = (base ^ 4096u) + g * (32 * W);
        if (W == 4) asm volatile("st.shared.u32 [%0], %1;" ::"r"(p), "r"(acc) : "memory");
        else if (W == 8) asm volatile("st.shared.v2.u32 [%0], {%1, %1};" ::"r"(p), "r"(acc) : "memory");
        else asm volatile("st.shared.v4.u32 [%0], {%1, %1, %1, %1};" ::"r"(p), "r"(acc) : "memory");
      }
    }
  }
  __syncthreads();
  const long long t1 = clock_after(acc);
  if (acc == 0x12345678u) out[0] = acc;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int W, int MODE>
void run(const char* name) {
  uint32_t* out; long long* cyc;
  cudaMalloc(&out, 4); cudaMalloc(&cyc, 148 * 8);
  const int iters = 1000;
  cudaFuncSetAttribute(probe<W, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
  probe<W, MODE><<<148, 256, 65536>>>(out, cyc, iters);
  probe<W, MODE><<<148, 256, 65536>>>(out, cyc, iters);
  cudaDeviceSynchronize();
  long long h[148];
  cudaMemcpy(h, cyc, sizeof h, cudaMemcpyDeviceToHost);
  const double n_instr = (double)iters * 8 * 8 * (MODE == 2 ? 2 : 1);   // warp instructions per SM
  const double bytes = n_instr * 32 * W;
  printf("%-22s %6.1f B/cycle/SM   %.2f cycles per warp instruction\n", name, bytes / (double)h[0], (double)h[0] / n_instr);
  cudaFree(out); cudaFree(cyc);
}
int main() {
  run<4, 0>("LDS.32"); run<8, 0>("LDS.64"); run<16, 0>("LDS.128");
  run<4, 1>("STS.32"); run<8, 1>("STS.64"); run<16, 1>("STS.128");
  run<4, 2>("LDS.32 + STS.32"); run<8, 2>("LDS.64 + STS.64"); run<16, 2>("LDS.128 + STS.128");
  return 0;
}
