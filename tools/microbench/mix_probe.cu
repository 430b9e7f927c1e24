// mix_probe.cu -- do FP64 instructions overlap with shared-memory stores / loads / shuffles / integer work issued by the same warps?
// 148 CTAs x 256 threads (2 warps per scheduler, like the throughput blind rotation), small loop body (instruction-cache resident).
// Per iteration and warp: ND DFMA (constant-bank multiplier) interleaved with the "other" instructions of the mode.
//   cycles per iteration if the two kinds overlap = max(FP64 pipe time, other pipe time); if they serialise = the sum.
// (loads next to FP64 work: lds_cost_probe.cu; instruction classes: coissue_probe.cu)
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mix_probe mix_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__constant__ double kc[8] = {1.0000001, 0.9999999, 1.0000002, 0.9999998, 1.0000003, 0.9999997, 1.0000004, 0.9999996};
// MODE: 0 none, 1 STS.64 x4, 2 STS.128 x2, 3 LDS.128 x4, 4 SHFL x8, 5 LOP3/IADD x16, 6 STS.64 x8, 7 LDS.32 x8 + 8 int (phase-A like)
template <int MODE, int ND>
__global__ void __launch_bounds__(256, 1) probe(double* sink, long long* cyc, int iters, double seed) {
  extern __shared__ __align__(16) unsigned char sm[];
  double a[16];
#pragma unroll
  for (int k = 0; k < 16; k++) a[k] = seed + threadIdx.x + k;
  double* p64 = reinterpret_cast<double*>(sm) + threadIdx.x;          // conflict-free 64-bit column
  double2* p128 = reinterpret_cast<double2*>(sm) + threadIdx.x;       // conflict-free 128-bit
  uint32_t* p32 = reinterpret_cast<uint32_t*>(sm) + threadIdx.x;
  uint32_t u[8];
#pragma unroll
  for (int k = 0; k < 8; k++) u[k] = threadIdx.x * 2654435761u + k;
  double2 acc2 = make_double2(0, 0);
  const uint32_t a64 = (uint32_t)__cvta_generic_to_shared(p64), a128 = (uint32_t)__cvta_generic_to_shared(p128);
  double ld[4] = {0, 0, 0, 0};
  uint32_t sumx = 0;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int k = 0; k < 16; k++) {
      if (k < ND) a[k] = __fma_rn(a[k], kc[k & 7], a[(k + 1) & 15]);
      // memory and integer instructions as volatile asm: nothing is hoisted, merged or promoted to registers
      if ((MODE == 1 && (k & 3) == 3) || (MODE == 6 && (k & 1) == 1))
        asm volatile("st.shared.f64 [%0], %1;" ::"r"(a64 + 2048u * (k >> 1)), "d"(a[k]) : "memory");
      if (MODE == 2 && (k & 7) == 7) asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"(a128 + 4096u * (k >> 3)), "d"(a[k]), "d"(a[k - 1]) : "memory");
      if (MODE == 3 && (k & 3) == 3) {
        double vx, vy;
        asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(vx), "=d"(vy) : "r"(a128 + 4096u * (k >> 2)) : "memory");
        ld[k >> 2] = vx + vy;   // consumed after the loop body's FP64 (one DADD per load)
      }
      if (MODE == 4 && (k & 1) == 1) asm volatile("shfl.sync.bfly.b32 %0, %0, 16, 31, 0xffffffff;" : "+r"(u[k >> 1]));
      if (MODE == 5) { asm volatile("add.u32 %0, %0, %1;" : "+r"(u[k & 7]) : "r"(0x9e3779b9u)); }
      if (MODE == 8) { asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(u[k & 7]) : "r"(0x9e3779b9u), "r"(u[(k + 4) & 7])); asm volatile("add.u32 %0, %0, %1;" : "+r"(u[(k + 2) & 7]) : "r"(0x9e3779b9u)); }
      if (MODE == 7 && (k & 1) == 1) {   // phase-A like: address LOP3, LDS.32, add, and
        uint32_t x, ad;
        asm volatile("lop3.b32 %0, %1, 8188, %2, 0xf8;" : "=r"(ad) : "r"(u[k >> 1]), "r"(a64));
        asm volatile("ld.shared.u32 %0, [%1];" : "=r"(x) : "r"(ad) : "memory");
        asm volatile("add.u32 %0, %0, %1;" : "+r"(x) : "r"(0x80000100u));
        asm volatile("and.b32 %0, %0, 0xFFFFFE00;" : "+r"(x));
        sumx ^= x;
      }
    }
  }
  const long long t1 = clock64();
  double s = acc2.x + ld[0] + ld[1] + ld[2] + ld[3] + sumx;
#pragma unroll
  for (int k = 0; k < 16; k++) s += a[k];
#pragma unroll
  for (int k = 0; k < 8; k++) s += u[k];
  if (s == 12345.678) sink[0] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int MODE, int ND>
void run(const char* name) {
  double* sink; long long* cyc;
  cudaMalloc(&sink, 8); cudaMalloc(&cyc, 148 * 8);
  const int iters = 4000;
  cudaFuncSetAttribute(probe<MODE, ND>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
  probe<MODE, ND><<<148, 256, 65536>>>(sink, cyc, iters, 0.5);
  probe<MODE, ND><<<148, 256, 65536>>>(sink, cyc, iters, 0.5);
  cudaDeviceSynchronize();
  long long h[148];
  cudaMemcpy(h, cyc, sizeof h, cudaMemcpyDeviceToHost);
  printf("%-44s ND=%2d  %.1f cycles per iteration (FP64 pipe alone: %d)\n", name, ND, (double)h[0] / iters, 4 * ND);
  cudaFree(sink); cudaFree(cyc);
}
int main() {
  run<0, 16>("DFMA only");
  run<1, 16>("DFMA + 4 STS.64 (LSU alone 64)");
  run<1, 0>("4 STS.64 alone");
  run<6, 16>("DFMA + 8 STS.64 (LSU alone 128)");
  run<6, 0>("8 STS.64 alone");
  run<2, 16>("DFMA + 2 STS.128 (LSU alone ~74)");
  run<2, 0>("2 STS.128 alone");
  run<4, 16>("DFMA + 8 SHFL (alone 64)");
  run<4, 0>("8 SHFL alone");
  run<8, 16>("DFMA + 16 x (LOP3, IADD)");
  run<8, 0>("16 x (LOP3, IADD) alone");
  run<8, 8>("8 DFMA + 16 x (LOP3, IADD)");
  return 0;
}
