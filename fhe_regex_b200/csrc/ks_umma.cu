// ks_umma.cu -- K1, the keyswitch contraction on the 5th-generation tensor cores (tcgen05.mma kind::i8, accumulators in
// tensor memory, operands staged by TMA).  Same exact arithmetic as ks_kernels.cu::ks_gemm_kernel (the legacy mma.sync
// path, kept as "ks_variant" 0): D[B x 10240] (s8 digits) x KB[10240 x 5944] (u8 key byte planes) -> s32, byte planes
// recombined with shifts mod 2^64 in the epilogue.  Bit-exact by construction (|sum| < 2^24 in int32).
//
// Replaces tfhe-rs keyswitch_lwe_ciphertext under /root/reference/src/regex/execution.rs:76,93,110,143,173,190.
//
// One persistent CTA per SM, 192 threads, warp-specialised:
//   warp 0     producer: cp.async.bulk.tensor (TMA, SWIZZLE_128B boxes) of the 128 x 128 B digit tile and the 256 x 128 B key
//              tile of a k-block into a 4-stage shared-memory ring, completion on mbarriers
//   warp 1     one thread issues tcgen05.mma.cta_group::1.kind::i8 (M 128, N 256, K 32; four per k-block) from shared-memory
//              matrix descriptors into one of two 256-column accumulators in tensor memory; tcgen05.commit hands the stage
//              back to the producer and, after the last k-block, the accumulator to the epilogue
//   warps 2-5  epilogue: tcgen05.ld 32 columns at a time (lane = ciphertext row), recombine the 8 byte planes of each of the
//              32 output columns, subtract from the body, store; overlaps the main loop of the next tile
// Tiles are walked with the key-column tile fastest, so the 24 tiles that share a digit tile run on neighbouring SMs
// (the 1.3 MB digit tile stays in L2) and the 63 MB of key bytes are re-read from L2, not from HBM.
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include "br_core.cuh"
#include "kernels.h"
#include "ptx_sync.cuh"

namespace fb {

namespace {
constexpr int UM_BM = 128, UM_BN = 256, UM_BK = 128, UM_STAGES = 4;
constexpr int kK = kN * kKsLevels;               // 10240
constexpr int kNPad = 6144;                      // key byte-plane rows padded to a multiple of UM_BN (ks_key_bytes)
constexpr int kNT = kNPad / UM_BN;               // 24 column tiles
constexpr int kKB = kK / UM_BK;                  // 80 k-blocks
constexpr uint32_t kABytes = UM_BM * UM_BK, kBBytes = UM_BN * UM_BK, kStageBytes = kABytes + kBBytes;   // 16 K + 32 K
constexpr size_t kUmmaSmem = (size_t)UM_STAGES * kStageBytes + 1024 /* alignment slack */ + 256 /* barriers */;

// L2 eviction policy of a TMA load.  The key byte planes (61 MB) are re-read by every wave of tiles; the digit matrix (10 KB per
// ciphertext, 291 MB at 28 416) is read by the CTAs that share a row tile and is dead afterwards.  Measured at 28 416 ciphertexts
// (ncu dram__bytes_read per launch, 0.35 GB compulsory): row-major tile order without hints 1.89 GB; key evict_last + digits
// evict_normal 1.73 GB (kept); digits evict_first 2.43 GB (the readers of a digit tile are not in step: an early eviction costs a
// re-read); with the banded tile order below 0.80 GB and 1.22 -> 1.11 ms (profiles/r02_traffic_bench_batch.csv).
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ uint64_t l2_policy_evict_normal() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ void tma_load_2d(uint32_t smem_dst, const CUtensorMap* tm, int c0, int c1, uint64_t* bar, uint64_t policy) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%2, %3}], [%4], %5;" ::"r"(smem_dst),
               "l"(reinterpret_cast<uint64_t>(tm)), "r"(c0), "r"(c1), "r"(smem_u32(bar)), "l"(policy)
               : "memory");
}
// whole-warp-free single-thread wait (the producer and the MMA issuer are single threads)
__device__ __forceinline__ void mbar_wait_thread(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  uint32_t done = 0;
  while (!done) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
  }
}
// K-major SWIZZLE_128B shared-memory matrix descriptor: rows of 128 B, 8-row groups 1024 B apart
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr) {
  return (uint64_t)((saddr >> 4) & 0x3fffu) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
// kind::i8 instruction descriptor: D s32, A signed 8 bit (digits), B unsigned 8 bit (key bytes), both K-major, N 256, M 128
constexpr uint32_t kIdesc = (2u << 4) | (1u << 7) | (0u << 10) | ((uint32_t)(UM_BN >> 3) << 17) | ((uint32_t)(UM_BM >> 4) << 24);

__device__ __forceinline__ void umma_i8(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(da), "l"(db), "r"(kIdesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld32x(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]),
        "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]),
        "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]),
        "=r"(v[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]), "+r"(v[7]), "+r"(v[8]), "+r"(v[9]),
                 "+r"(v[10]), "+r"(v[11]), "+r"(v[12]), "+r"(v[13]), "+r"(v[14]), "+r"(v[15]), "+r"(v[16]), "+r"(v[17]), "+r"(v[18]),
                 "+r"(v[19]), "+r"(v[20]), "+r"(v[21]), "+r"(v[22]), "+r"(v[23]), "+r"(v[24]), "+r"(v[25]), "+r"(v[26]), "+r"(v[27]),
                 "+r"(v[28]), "+r"(v[29]), "+r"(v[30]), "+r"(v[31])
               :
               : "memory");
}

// Tile order.  Output tiles are walked in bands of kBandRows row tiles; inside a band the key column tiles go in groups of
// kBandCols, the row tile next, the column of the group fastest.  A wave of 148 concurrent tiles then works on ~37 digit tiles
// (48 MB, re-used by the 6 groups of the band from L2) and 4 key tiles (10 MB) instead of 6 digit tiles and the whole 61 MB key:
// the key is streamed once per band (6 times at 28 416 ciphertexts) instead of once per wave (36 times).
constexpr int kBandRows = 37, kBandCols = 4;
static_assert(kNT % kBandCols == 0, "key column tiles come in whole groups");
__device__ __forceinline__ void tile_coords(int tile, int m_tiles, int& mt, int& nt) {
  const int band = tile / (kBandRows * kNT);
  const int rows = min(kBandRows, m_tiles - band * kBandRows);          // the last band may be short
  const int u = tile - band * (kBandRows * kNT);
  const int g = u / (rows * kBandCols), v = u - g * (rows * kBandCols);
  mt = band * kBandRows + v / kBandCols;
  nt = g * kBandCols + v % kBandCols;
}

__global__ void __launch_bounds__(192, 1)
ks_umma_kernel(const __grid_constant__ CUtensorMap tm_dig, const __grid_constant__ CUtensorMap tm_key, const uint64_t* __restrict__ in,
               const int32_t* __restrict__ in_rows, uint64_t* __restrict__ out, int count, int m_tiles) {
  extern __shared__ unsigned char smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;          // SWIZZLE_128B tiles want 1024-byte alignment
  unsigned char* smem = smem_raw + (base - smem_u32(smem_raw));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)UM_STAGES * kStageBytes);
  uint64_t* full = bars;                       // [STAGES] bytes of the stage have landed
  uint64_t* empty = bars + UM_STAGES;          // [STAGES] the MMAs that read the stage have completed
  uint64_t* acc_full = bars + 2 * UM_STAGES;   // [2] accumulator complete
  uint64_t* acc_empty = acc_full + 2;          // [2] accumulator drained by the epilogue
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_tiles = m_tiles * kNT;

  if (threadIdx.x == 0) {
    for (int s = 0; s < UM_STAGES; s++) {
      mbar_init(full + s, 1);
      mbar_init(empty + s, 1);
    }
    for (int a = 0; a < 2; a++) {
      mbar_init(acc_full + a, 1);
      mbar_init(acc_empty + a, 4);             // one arrival per epilogue warp
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ---- producer ------------------------------------------------------------------------------------------------------
    if (lane == 0) {
      const uint64_t pol_dig = l2_policy_evict_normal(), pol_key = l2_policy_evict_last();
      uint32_t it = 0;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        int mt, nt;
        tile_coords(tile, m_tiles, mt, nt);
        for (int kb = 0; kb < kKB; kb++, it++) {
          const uint32_t s = it % UM_STAGES, ph = (it / UM_STAGES) & 1u;
          mbar_wait_thread(empty + s, ph ^ 1u);
          mbar_arrive_expect_tx(full + s, kStageBytes);
          const uint32_t dstA = base + s * kStageBytes, dstB = dstA + kABytes;
          tma_load_2d(dstA, &tm_dig, kb * UM_BK, mt * UM_BM, full + s, pol_dig);
          tma_load_2d(dstB, &tm_key, kb * UM_BK, nt * UM_BN, full + s, pol_key);
        }
      }
    }
  } else if (warp == 1) {
    // ---- MMA issuer ----------------------------------------------------------------------------------------------------
    if (lane == 0) {
      uint32_t it = 0, t_local = 0;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, t_local++) {
        const uint32_t a = t_local & 1u, aph = (t_local >> 1) & 1u;
        mbar_wait_thread(acc_empty + a, aph ^ 1u);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t tmem_d = tmem_base + a * (uint32_t)UM_BN;
        for (int kb = 0; kb < kKB; kb++, it++) {
          const uint32_t s = it % UM_STAGES, ph = (it / UM_STAGES) & 1u;
          mbar_wait_thread(full + s, ph);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint64_t da = umma_desc(base + s * kStageBytes), db = umma_desc(base + s * kStageBytes + kABytes);
#pragma unroll
          for (int k = 0; k < UM_BK / 32; k++) umma_i8(tmem_d, da + 2u * k, db + 2u * k, (kb | k) != 0);   // +32 bytes along K
          umma_commit(empty + s);
        }
        umma_commit(acc_full + a);
      }
    }
  } else {
    // ---- epilogue ------------------------------------------------------------------------------------------------------
    const int quarter = warp & 3;               // TMEM lanes 32 * quarter .. + 31 = rows of the tile
    uint32_t t_local = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, t_local++) {
      int mt, nt;
      tile_coords(tile, m_tiles, mt, nt);
      const uint32_t a = t_local & 1u, aph = (t_local >> 1) & 1u;
      mbar_wait(acc_full + a, aph);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const int b = mt * UM_BM + quarter * 32 + lane;
      const bool row_ok = b < count;
      uint64_t body = 0;
      if (row_ok && (nt * (UM_BN / 8) <= kLweN) && (kLweN < (nt + 1) * (UM_BN / 8))) {
        const size_t row = in_rows ? (size_t)in_rows[b] : (size_t)b;
        body = in[row * kBig + kN];
      }
      const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + a * (uint32_t)UM_BN;
#pragma unroll 1
      for (int ch = 0; ch < UM_BN / 32; ch++) {
        uint32_t v[32];
        tmem_ld32x(taddr + 32 * ch, v);
#pragma unroll
        for (int j = 0; j < 4; j++) {
          uint64_t sum = 0;
#pragma unroll
          for (int t = 0; t < 8; t++) sum += (uint64_t)(int64_t)(int32_t)v[8 * j + t] << (8 * t);
          const int c = nt * (UM_BN / 8) + 4 * ch + j;
          if (row_ok && c < kSmall) out[(size_t)b * kSmall + c] = (c == kLweN ? body : 0ull) - sum;
        }
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(acc_empty + a);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_tiled() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// [rows][10240] bytes, box = box_rows x 128 bytes, 128-byte swizzle
bool make_map(CUtensorMap* tm, const void* ptr, uint64_t rows, uint32_t box_rows) {
  EncodeTiledFn fn = encode_tiled();
  if (!fn) return false;
  const cuuint64_t dims[2] = {(cuuint64_t)kK, (cuuint64_t)rows};
  const cuuint64_t strides[1] = {(cuuint64_t)kK};
  const cuuint32_t box[2] = {(cuuint32_t)UM_BK, box_rows};
  const cuuint32_t estr[2] = {1, 1};
  return fn(tm, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}
}  // namespace

// digits already written by ks_decompose_kernel ([m_tiles * 128][10240] s8); kb padded to 6144 rows
cudaError_t launch_keyswitch_umma_gemm(const uint8_t* kb, const int8_t* dig, const uint64_t* in, const int32_t* in_rows, uint64_t* out,
                                       int count, int sms, cudaStream_t st) {
  const int m_tiles = (count + UM_BM - 1) / UM_BM;
  CUtensorMap tm_dig, tm_key;
  if (!make_map(&tm_dig, dig, (uint64_t)m_tiles * UM_BM, UM_BM) || !make_map(&tm_key, kb, kNPad, UM_BN)) return cudaErrorNotSupported;
  static PerDeviceOnce once;
  bool& configured = *once.slot();
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(ks_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kUmmaSmem);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  const int n_tiles = m_tiles * kNT;
  const int grid = n_tiles < sms ? n_tiles : sms;
  ks_umma_kernel<<<grid, 192, kUmmaSmem, st>>>(tm_dig, tm_key, in, in_rows, out, count, m_tiles);
  return cudaGetLastError();
}

}  // namespace fb
