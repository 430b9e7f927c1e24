// ks_kernels.cu -- K1: LWE keyswitch as an exact int8 tensor-core contraction.
//
// Replaces tfhe-rs 0.2.0 keyswitch_lwe_ciphertext under every smart_* call of the reference
// (/root/reference/src/regex/execution.rs:76,93,110,143,173,190).
//
//   out[b][c] = [c == 742] * in[b][2048] - sum_{i<2048} sum_{l<5} d(b,i,l) * KSK[i][l][c]   (mod 2^64)
//
// The digits d are tiny (signed base-8, |d| <= 4) and the key words are 64-bit, so the contraction over
// k = (i,l) (10 240 terms) is done bytewise: KSK[k][c] = sum_t 2^(8t) * kb[k][c][t] with unsigned bytes kb,
// hence sum_k d_k * KSK[k][c] = sum_t 2^(8t) * (sum_k d_k * kb[k][c][t]).  Each inner sum is an s8 x u8 dot
// product of 10 240 terms (|.| <= 10 240 * 4 * 255 < 2^24: exact in the int32 accumulators of the
// tensor cores), i.e. one dense GEMM  D[B x 10240] (s8)  x  KB[10240 x 5944] (u8)  ->  s32, whose epilogue
// recombines the 8 byte planes of a column with shifts mod 2^64.  Bit-exact by construction.
//
//   ksk_bytes_kernel     key load: KSK [10240][743] u64  ->  KB [6016][10240] u8, row n = 8*c + t (k contiguous)
//   ks_decompose_kernel  in [B][2049] u64 -> digits [Bpad][10240] s8 (closest representable on 15 bits,
//                        balanced base-8 digits, level rows most significant first)
//   ks_gemm_kernel       128x128x64 CTA tiles, 4-stage cp.async pipeline, ldmatrix + mma.sync.m16n8k32.s8.u8
#include <cuda_runtime.h>
#include <stdint.h>
#include "br_core.cuh"
#include "kernels.h"

namespace fb {

constexpr int kKsK = kN * kKsLevels;          // 10240 contraction length
constexpr int kKsNReal = kSmall * 8;          // 5944 byte-plane columns
constexpr int KS_BM = 128, KS_BN = 128, KS_BK = 64, KS_STAGES = 4;
constexpr int kKsNTiles = (kKsNReal + KS_BN - 1) / KS_BN;           // 47 column tiles of the mma.sync GEMM
constexpr int kKsNPad = 6144;                                       // rows of the byte-plane matrix: a multiple of the 256-wide tcgen05 tile (ks_umma.cu)

size_t ks_key_bytes() { return (size_t)kKsNPad * kKsK; }
size_t ks_digit_bytes(int count) { return (size_t)((count + KS_BM - 1) / KS_BM) * KS_BM * kKsK; }

// ---- key load: byte-plane transpose ----------------------------------------------------------------
// grid (10240/32, ceil(743/32)), block (32, 8): tile of 32 k x 32 c words through shared memory
__global__ void __launch_bounds__(256)
ksk_bytes_kernel(const uint64_t* __restrict__ ksk, uint8_t* __restrict__ kb) {
  __shared__ uint64_t tile[32][33];
  const int k0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  for (int r = threadIdx.y; r < 32; r += 8) {
    const int c = c0 + threadIdx.x;
    tile[r][threadIdx.x] = (c < kSmall) ? ksk[(size_t)(k0 + r) * kSmall + c] : 0ull;
  }
  __syncthreads();
  // thread (x, y): k = k0 + x; columns c0 + y, y + 8, ...; 8 bytes -> 8 rows of KB
  for (int cc = threadIdx.y; cc < 32; cc += 8) {
    const int c = c0 + cc;
    if (c >= kSmall) continue;
    const uint64_t v = tile[threadIdx.x][cc];
#pragma unroll
    for (int t = 0; t < 8; t++) kb[(size_t)(c * 8 + t) * kKsK + k0 + threadIdx.x] = (uint8_t)(v >> (8 * t));
  }
}

// ---- digits ---------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t ks_pack_digits(uint64_t x) {
  // closest representable on the top 15 bits, then 5 balanced base-8 digits, least significant first;
  // the digit produced at iteration t multiplies KSK level row l = 4 - t (rows are stored most
  // significant level first).  Packed as 5 signed nibbles, nibble l = digit for row l.
  uint32_t state = (uint32_t)((((x >> 48) + 1ull) >> 1) & 0x7FFFull);
  uint32_t packed = 0;
#pragma unroll
  for (int t = 0; t < kKsLevels; t++) {
    uint32_t res = state & 7u;
    state >>= 3;
    uint32_t carry = (((res - 1u) | state) & res) >> 2;
    state += carry;
    uint32_t digit = (res - (carry << 3)) & 0xFu;
    packed |= digit << (4 * (kKsLevels - 1 - t));
  }
  return packed;
}

// one thread per (sample, 4 consecutive mask words): 20 digit bytes = 5 aligned 32-bit stores
__global__ void __launch_bounds__(256)
ks_decompose_kernel(const uint64_t* __restrict__ in, const int32_t* __restrict__ in_rows, int8_t* __restrict__ dig, int count,
                    int count_pad) {
  const int b = blockIdx.x;
  const int i4 = blockIdx.y * 256 + threadIdx.x;  // group of 4 mask words
  if (b >= count_pad || i4 >= kN / 4) return;
  uint32_t w[5] = {0, 0, 0, 0, 0};
  if (b < count) {
    const size_t row = in_rows ? (size_t)in_rows[b] : (size_t)b;
    const uint64_t* src = in + row * kBig + 4 * i4;
    uint8_t bytes[20];
#pragma unroll
    for (int j = 0; j < 4; j++) {
      const uint32_t p = ks_pack_digits(src[j]);
#pragma unroll
      for (int l = 0; l < kKsLevels; l++) bytes[5 * j + l] = (uint8_t)(((int32_t)(p << (28 - 4 * l))) >> 28);
    }
#pragma unroll
    for (int q = 0; q < 5; q++)
      w[q] = (uint32_t)bytes[4 * q] | ((uint32_t)bytes[4 * q + 1] << 8) | ((uint32_t)bytes[4 * q + 2] << 16) | ((uint32_t)bytes[4 * q + 3] << 24);
  }
  uint32_t* dst = reinterpret_cast<uint32_t*>(dig + (size_t)b * kKsK + 20 * (size_t)i4);
#pragma unroll
  for (int q = 0; q < 5; q++) dst[q] = w[q];
}

// ---- GEMM -----------------------------------------------------------------------------------------
__device__ __forceinline__ void cp_async16(uint32_t smem_addr, const void* g) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_addr), "l"(g) : "memory");
}
__device__ __forceinline__ void ldmatrix_x4(uint32_t (&r)[4], uint32_t smem_addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(smem_addr));
}
__device__ __forceinline__ void mma_s8u8(int32_t (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.s8.u8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+r"(c[0]), "+r"(c[1]), "+r"(c[2]), "+r"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// tile row r (64 bytes = 4 chunks of 16 B): chunk j is stored at j ^ ((r >> 1) & 3) -> conflict-free ldmatrix
__device__ __forceinline__ uint32_t tile_off(int r, int chunk) { return (uint32_t)(r * KS_BK + ((chunk ^ ((r >> 1) & 3)) << 4)); }

__global__ void __launch_bounds__(256)
ks_gemm_kernel(const int8_t* __restrict__ dig, const uint8_t* __restrict__ kb, const uint64_t* __restrict__ in,
               const int32_t* __restrict__ in_rows, uint64_t* __restrict__ out, int count) {
  extern __shared__ __align__(128) unsigned char smem[];
  unsigned char* As = smem;                                     // [STAGES][128][64]
  unsigned char* Bs = smem + KS_STAGES * KS_BM * KS_BK;         // [STAGES][128][64]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int warp_m = warp >> 2, warp_n = warp & 3;             // 2 x 4 warps, warp tile 64 x 32
  const int n0 = blockIdx.x * KS_BN, m0 = blockIdx.y * KS_BM;
  const int8_t* gA = dig + (size_t)m0 * kKsK;
  const uint8_t* gB = kb + (size_t)n0 * kKsK;
  const uint32_t sA = (uint32_t)__cvta_generic_to_shared(As), sB = (uint32_t)__cvta_generic_to_shared(Bs);

  auto load_stage = [&](int stage, int kt) {
    const int kofs = kt * KS_BK;
#pragma unroll
    for (int j = 0; j < 2; j++) {
      const int idx = tid + j * 256;          // 512 chunks per operand tile
      const int r = idx >> 2, ch = idx & 3;
      cp_async16(sA + stage * (KS_BM * KS_BK) + tile_off(r, ch), gA + (size_t)r * kKsK + kofs + ch * 16);
      cp_async16(sB + stage * (KS_BN * KS_BK) + tile_off(r, ch), gB + (size_t)r * kKsK + kofs + ch * 16);
    }
  };

  int32_t acc[4][4][4];
#pragma unroll
  for (int mi = 0; mi < 4; mi++)
#pragma unroll
    for (int ni = 0; ni < 4; ni++)
#pragma unroll
      for (int e = 0; e < 4; e++) acc[mi][ni][e] = 0;

  constexpr int KT = kKsK / KS_BK;  // 160
#pragma unroll
  for (int st = 0; st < KS_STAGES - 1; st++) {
    load_stage(st, st);
    asm volatile("cp.async.commit_group;" ::: "memory");
  }
  for (int kt = 0; kt < KT; kt++) {
    asm volatile("cp.async.wait_group %0;" ::"n"(KS_STAGES - 2) : "memory");
    __syncthreads();
    if (kt + KS_STAGES - 1 < KT) load_stage((kt + KS_STAGES - 1) % KS_STAGES, kt + KS_STAGES - 1);
    asm volatile("cp.async.commit_group;" ::: "memory");
    const int stage = kt % KS_STAGES;
    const uint32_t aBase = sA + stage * (KS_BM * KS_BK), bBase = sB + stage * (KS_BN * KS_BK);
#pragma unroll
    for (int kk = 0; kk < 2; kk++) {
      uint32_t a[4][4], b[2][4];
#pragma unroll
      for (int mi = 0; mi < 4; mi++) ldmatrix_x4(a[mi], aBase + tile_off(warp_m * 64 + mi * 16 + (lane & 15), kk * 2 + (lane >> 4)));
#pragma unroll
      for (int nj = 0; nj < 2; nj++)
        ldmatrix_x4(b[nj], bBase + tile_off(warp_n * 32 + nj * 16 + (lane & 7) + ((lane >> 4) << 3), kk * 2 + ((lane >> 3) & 1)));
#pragma unroll
      for (int mi = 0; mi < 4; mi++)
#pragma unroll
        for (int ni = 0; ni < 4; ni++) mma_s8u8(acc[mi][ni], a[mi], b[ni >> 1][(ni & 1) * 2], b[ni >> 1][(ni & 1) * 2 + 1]);
    }
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");

  // epilogue: column n = 8*c + t; this thread holds bytes t = 2*(lane%4) + {0,1} of one c per n8 tile
  const int g = lane >> 2, t0 = (lane & 3) * 2;
#pragma unroll
  for (int mi = 0; mi < 4; mi++) {
#pragma unroll
    for (int ni = 0; ni < 4; ni++) {
      const int c = (n0 + warp_n * 32 + ni * 8) >> 3;
#pragma unroll
      for (int h = 0; h < 2; h++) {  // rows g and g + 8
        uint64_t v = ((uint64_t)(int64_t)acc[mi][ni][2 * h] << (8 * t0)) + ((uint64_t)(int64_t)acc[mi][ni][2 * h + 1] << (8 * t0 + 8));
        v += __shfl_xor_sync(0xffffffffu, v, 1);
        v += __shfl_xor_sync(0xffffffffu, v, 2);
        const int b = m0 + warp_m * 64 + mi * 16 + g + 8 * h;
        if ((lane & 3) == 0 && b < count && c < kSmall) {
          uint64_t body = 0;
          if (c == kLweN) {
            const size_t row = in_rows ? (size_t)in_rows[b] : (size_t)b;
            body = in[row * kBig + kN];
          }
          out[(size_t)b * kSmall + c] = body - v;
        }
      }
    }
  }
}

// ---- launchers -------------------------------------------------------------------------------------
cudaError_t launch_ksk_bytes(const uint64_t* ksk, uint8_t* kb, cudaStream_t st) {
  cudaError_t e = cudaMemsetAsync(kb, 0, ks_key_bytes(), st);  // padding rows 5944..6143
  if (e != cudaSuccess) return e;
  ksk_bytes_kernel<<<dim3(kKsK / 32, (kSmall + 31) / 32), dim3(32, 8), 0, st>>>(ksk, kb);
  return cudaGetLastError();
}

cudaError_t launch_keyswitch_umma_gemm(const uint8_t* kb, const int8_t* dig, const uint64_t* in, const int32_t* in_rows, uint64_t* out,
                                       int count, int sms, cudaStream_t st);   // ks_umma.cu

// variant 0: mma.sync GEMM (this file); 1: tcgen05 GEMM (ks_umma.cu).  Same digits, same key bytes, same result.
cudaError_t launch_keyswitch_mma(const uint8_t* kb, int8_t* dig, const uint64_t* in, const int32_t* in_rows, uint64_t* out, int count,
                                 int variant, int sms, cudaStream_t st) {
  if (count <= 0) return cudaSuccess;
  const int mt = (count + KS_BM - 1) / KS_BM;
  ks_decompose_kernel<<<dim3(mt * KS_BM, 2), 256, 0, st>>>(in, in_rows, dig, count, mt * KS_BM);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  if (variant == 1) return launch_keyswitch_umma_gemm(kb, dig, in, in_rows, out, count, sms, st);
  constexpr int smem = KS_STAGES * (KS_BM + KS_BN) * KS_BK;  // 65536
  static PerDeviceOnce once;
  bool& configured = *once.slot();
  if (!configured) {
    e = cudaFuncSetAttribute(ks_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  ks_gemm_kernel<<<dim3(kKsNTiles, mt), 256, smem, st>>>(dig, kb, in, in_rows, out, count);
  return cudaGetLastError();
}

}  // namespace fb
