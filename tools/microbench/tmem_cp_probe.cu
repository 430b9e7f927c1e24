// tmem_cp_probe.cu -- semantics of tcgen05.cp.cta_group::1.64x128b.warpx2::02_13 with a SWIZZLE_NONE shared-memory descriptor:
// which shared-memory bytes land in which TMEM lane / column.  Shared memory is filled with its own word index; after the copy
// every thread reads its TMEM lane back with tcgen05.ld.32x32b.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tmem_cp_probe tmem_cp_probe.cu
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3fff);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3fff) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3fff) << 32;
  d |= (uint64_t)1 << 46;   // descriptor version of sm_100
  return d;                 // layout type 0 = SWIZZLE_NONE, base offset 0
}

__global__ void __launch_bounds__(128, 1) probe(uint32_t* out, uint32_t lbo, uint32_t sbo) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint32_t slot;
  __shared__ __align__(8) uint64_t bar;
  uint32_t* w = reinterpret_cast<uint32_t*>(smem);
  for (int i = threadIdx.x; i < 4096; i += 128) w[i] = i;   // 16 KiB: word i holds i
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(1) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(32) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tbase = slot;
  if (threadIdx.x == 0) {
    // two copies: columns 0-3 from smem byte 0, columns 4-7 from smem byte 4096
    const uint64_t d0 = make_desc(smem_u32(smem), lbo, sbo);
    const uint64_t d1 = make_desc(smem_u32(smem) + 4096, lbo, sbo);
    asm volatile("tcgen05.cp.cta_group::1.64x128b.warpx2::02_13 [%0], %1;" ::"r"(tbase), "l"(d0) : "memory");
    asm volatile("tcgen05.cp.cta_group::1.64x128b.warpx2::02_13 [%0], %1;" ::"r"(tbase + 4), "l"(d1) : "memory");
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
  }
  // everybody waits for the copies
  uint32_t done = 0;
  while (!done) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(smem_u32(&bar)), "r"(0) : "memory");
  }
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  uint32_t v[8];
  const uint32_t taddr = tbase + ((uint32_t)(warp * 32) << 16);
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]) : "r"(taddr) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  for (int c = 0; c < 8; c++) out[threadIdx.x * 8 + c] = v[c];
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tbase), "r"(32) : "memory");
}

int main() {
  uint32_t* d;
  cudaMalloc(&d, 128 * 8 * 4);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384);
  const uint32_t cfg[][2] = {{128, 128}, {16, 128}, {1024, 128}, {128, 256}};
  for (auto& c : cfg) {
    cudaMemset(d, 0xff, 128 * 8 * 4);
    probe<<<1, 128, 16384>>>(d, c[0], c[1]);
    cudaError_t e = cudaDeviceSynchronize();
    uint32_t h[128 * 8];
    cudaMemcpy(h, d, sizeof h, cudaMemcpyDeviceToHost);
    printf("LBO %u SBO %u: %s\n", c[0], c[1], cudaGetErrorString(e));
    if (e != cudaSuccess) return 1;
    for (int t : {0, 1, 7, 8, 9, 31, 32, 33, 63, 64, 65, 96, 127}) {
      printf("  lane %3d:", t);
      for (int k = 0; k < 8; k++) printf(" %5u", h[t * 8 + k]);
      printf("\n");
    }
  }
  return 0;
}
