// tmem_ld_probe.cu -- throughput of tcgen05.ld (LDTM) with 8 warps per SM reading thread-private columns, against LDS.128 of the
// same bytes.  Question: can the Fourier GGSW of a CMUX step be served to the MAC from tensor memory instead of shared memory?
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tmem_ld_probe tmem_ld_probe.cu
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int MODE>
__global__ void __launch_bounds__(256, 1) probe(uint32_t* out, long long* cyc, int iters) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tbase = slot + ((uint32_t)((warp & 3) * 32) << 16);
  uint32_t acc = 0;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
    if (MODE == 0) {          // 16 x LDTM.x16 = 256 columns = 1 KiB per thread
#pragma unroll
      for (int g = 0; g < 16; g++) {
        uint32_t v[16];
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                     : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
                       "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                     : "r"(tbase + 16 * g) : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;" : "+r"(v[0]), "+r"(v[5]), "+r"(v[10]), "+r"(v[15]) : : "memory");
        acc += v[0] ^ v[5] ^ v[10] ^ v[15];
      }
    } else if (MODE == 1) {   // 4 x LDTM.x64
#pragma unroll
      for (int g = 0; g < 4; g++) {
        uint32_t v[64];
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x64.b32 "
            "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,"
            "%32,%33,%34,%35,%36,%37,%38,%39,%40,%41,%42,%43,%44,%45,%46,%47,%48,%49,%50,%51,%52,%53,%54,%55,%56,%57,%58,%59,%60,%61,%62,%63}, [%64];"
            : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]),
              "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]),
              "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]),
              "=r"(v[31]), "=r"(v[32]), "=r"(v[33]), "=r"(v[34]), "=r"(v[35]), "=r"(v[36]), "=r"(v[37]), "=r"(v[38]), "=r"(v[39]), "=r"(v[40]),
              "=r"(v[41]), "=r"(v[42]), "=r"(v[43]), "=r"(v[44]), "=r"(v[45]), "=r"(v[46]), "=r"(v[47]), "=r"(v[48]), "=r"(v[49]), "=r"(v[50]),
              "=r"(v[51]), "=r"(v[52]), "=r"(v[53]), "=r"(v[54]), "=r"(v[55]), "=r"(v[56]), "=r"(v[57]), "=r"(v[58]), "=r"(v[59]), "=r"(v[60]),
              "=r"(v[61]), "=r"(v[62]), "=r"(v[63])
            : "r"(tbase + 64 * g) : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;" : "+r"(v[0]), "+r"(v[21]), "+r"(v[42]), "+r"(v[63]) : : "memory");
        acc += v[0] ^ v[21] ^ v[42] ^ v[63];
      }
    } else {                  // 64 x LDS.128 = 1 KiB per thread, conflict-free
      const uint4* p = reinterpret_cast<const uint4*>(smem) + threadIdx.x % 64;
#pragma unroll
      for (int g = 0; g < 64; g++) {
        uint4 v = p[64 * g];
        acc += v.x ^ v.w;
      }
      asm volatile("" ::: "memory");
    }
  }
  const long long t1 = clock64();
  if (acc == 0x12345678u) out[0] = acc;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(slot), "r"(512) : "memory");
}

template <int MODE>
void run(const char* name) {
  uint32_t* out; long long* cyc;
  cudaMalloc(&out, 4); cudaMalloc(&cyc, 148 * 8);
  const int iters = 200;
  cudaFuncSetAttribute(probe<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
  probe<MODE><<<148, 256, 65536>>>(out, cyc, iters);
  probe<MODE><<<148, 256, 65536>>>(out, cyc, iters);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[148];
  cudaMemcpy(h, cyc, sizeof h, cudaMemcpyDeviceToHost);
  // 8 warps x 32 lanes x 1 KiB per iteration = 256 KiB per SM per iteration
  printf("%-12s %s: %.0f cycles per iteration (256 KiB per SM) = %.1f B/cycle/SM\n", name, cudaGetErrorString(e), (double)h[0] / iters,
         262144.0 * iters / (double)h[0]);
}

int main() {
  run<0>("LDTM.x16");
  run<1>("LDTM.x64");
  run<2>("LDS.128");
  return 0;
}
