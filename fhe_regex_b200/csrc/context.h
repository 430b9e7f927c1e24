// context.h -- fb_ctx: per-GPU state behind the C ABI (include/fhe_b200.h).  Internal.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <memory>
#include <string>
#include <utility>
#include <vector>
#include "../../include/fhe_b200.h"
#include "kernels.h"

namespace fbre { struct Plan; }

struct fb_devbuf {
  void* p = nullptr;
  size_t cap = 0;
};

// a device-resident array of ciphertext rows (2049 words) or accumulator polynomials (2048 words) behind an fb_handle
struct fb_arena {
  uint64_t* p = nullptr;
  size_t rows = 0, row_words = 0;
};

struct fb_event_pair {
  cudaEvent_t a, b;
  int kind;  // 0 ks, 1 br, 2 lin
};

struct fb_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  cudaStream_t h2d_stream = nullptr, d2h_stream = nullptr;   // fb_pbs_batch: copies of neighbouring chunks under the bootstraps
  std::vector<cudaEvent_t> pipe_events;
  std::string err;
  // keys
  uint8_t* d_kb = nullptr;     // byte planes of the KSK for the tensor-core keyswitch
  fb::c2* d_fbsk = nullptr;
  fb::c2* d_fbsk_lm = nullptr; // the same key in the lane-major order of the tensor-memory MAC (br_fused.cu::fbsk_lm_kernel)
  fb::c2* d_tabs = nullptr;
  fb::c2* d_wtab = nullptr;    // table of the latency blind rotation (inside the d_tabs allocation)
  fb::c2* d_dtab = nullptr;    // table of the cluster blind rotation (inside the d_tabs allocation)
  int duo_max = 0;             // batches up to this many PBS take the cluster kernel (off by default; option "cluster_threshold")
  int duo_pairs = 0;           // CTA pairs the device runs at once (cudaOccupancyMaxActiveClusters)
  int quantum = 592;           // SM count x PBS per CTA of the throughput blind rotation
  int wide_max = 296;          // batches up to this many PBS take the latency kernel (option "latency_threshold"; 0 = never)
  int wide_skew = 200;         // latency kernel: cycles one half of the CTA is held back after the MAC (option "wide_skew")
  int wide_prefetch = 3;       // latency kernel: GGSW groups fetched before the pre-MAC barrier (option "wide_prefetch")
  int wide_pair = 1;           // 1: narrow batches wider than the SM count take two PBS per CTA (br_wide2.cu); 2: every narrow batch does (tests); 0: never (option "wide_pair")
  int wide_pair_offset = 0;    // pair kernel: cycles the second sample of a CTA starts late (option "wide_pair_offset")
  int wide_pair_prefetch = 1;  // pair kernel: GGSW groups fetched before the pre-MAC barrier (option "wide_pair_prefetch")
  int br_variant = 2;          // throughput blind rotation at 4 PBS per SM: 0 = phase-by-phase body (kernels.cu), 1 = fused body (br_fused.cu), 2 = fused + digits through I2F
  int ks_variant = 1;          // keyswitch GEMM: 0 = mma.sync (ks_kernels.cu), 1 = tcgen05 (ks_umma.cu) (option "ks_variant")
  int sms = 148;               // multiprocessors of the device
  int br_samples = 4;          // fused throughput kernel: PBS per CTA, 4 or 6 (option "br_samples"); quantum follows
  int br_stagger_groups = 0;   // 1: the skew goes to the odd samples only (option "br_stagger_groups")
  int br_planes = 2;           // 2: a transpose plane per component in the fused kernel at 4 PBS per CTA (option "br_planes")
  int pbs_chunks = 3;          // fb_pbs_batch: 3 = chunks of 4, rest, 4 waves; 5 = 1, 6, rest, 6, 1 for batches of at least 24 waves (option "pbs_chunks")
  int br_sync = 1;             // fused throughput kernel: 1 = the samples of a CTA start their rotation together (option "br_sync")
  int br_resync = 8;           // ... and meet at a CTA-wide barrier every this many CMUX steps, 0 = never (option "br_resync")
  int br_barriers = 0;         // fused throughput kernel: 1 keeps two unneeded barriers per step (option "br_barriers", A/B only)
  int br_stagger = 0;          // fused throughput kernel: start skew between the samples of a CTA, cycles per sample index (option "br_stagger")
  bool plan_absorb = true;     // false: reference-shaped plan (option "plan_reference_shaped" = 1)
  bool plan_timing = false;    // planner phase times on stderr (option "plan_timing")
  bool have_key = false;
  // multi-GPU: this context as a rank of an NCCL communicator (comm.cu); ncclComm_t kept opaque here
  void* comm = nullptr;
  int comm_rank = 0, comm_world = 1;
  int dist_shard_min = 148;    // fb_has_match_dist: levels of at most this many PBS (one wave of SMs) are computed by every rank, not sharded (option "dist_shard_min")
  uint64_t comm_exchanges = 0, comm_bytes = 0;   // level exchanges issued / ciphertext bytes gathered per rank
  // scratch for the batch entry points
  fb_devbuf in, small, out, luts, lut_idx, digits;
  // has_match: arena of ciphertext rows, flattened plan arrays, the fixed accumulator table (uploaded once)
  fb_devbuf arena, plan_i32, plan_i64, plan_u64, plan_u32, regex_luts;
  bool regex_luts_ready = false;
  // op-level boundary (handles.cu): caller-visible arenas and the staging buffers of their index arrays
  std::vector<fb_arena> arenas;
  fb_devbuf op_i32, op_i64, op_u64, op_u32, op_rows;
  // lowered plans of recent (pattern, content length, rank, world) requests: the plan depends on nothing else, so a
  // server matching many contents against one pattern builds it once (most recent first, at most 8)
  std::vector<std::pair<std::string, std::shared_ptr<const fbre::Plan>>> plan_cache;
  // timing
  bool timing = false;
  std::vector<fb_event_pair> pending;
  std::vector<fb_event_pair> pool;
  fb_kernel_stats ks{};
};

int fb_fail(fb_ctx* ctx, int code, const std::string& msg);
int fb_cuda_fail(fb_ctx* ctx, cudaError_t e, const char* what);
int fb_reserve(fb_ctx* ctx, fb_devbuf& b, size_t bytes);
#define FB_CUDA(ctx, call)                                        \
  do {                                                            \
    cudaError_t _e = (call);                                      \
    if (_e != cudaSuccess) return fb_cuda_fail(ctx, _e, #call);   \
  } while (0)

// comm.cu: slice r of `world` contiguous slices of n items; in-place exchange of row slices over the communicator
void fb_comm_slice(size_t n, int r, int world, size_t* lo, size_t* hi);
int fb_comm_exchange_rows(fb_ctx* ctx, uint64_t* d_rows, size_t n, size_t row_words);

// api.cu: both keys already in device memory (standard domain) -> resident forms
int fb_install_server_key_device(fb_ctx* ctx, uint64_t* d_ksk, uint64_t* d_bsk_std);

// timed launches on ctx->stream (timing is a no-op unless enabled)
int fb_run_keyswitch(fb_ctx* ctx, const uint64_t* d_in, const int32_t* d_in_rows, uint64_t* d_small, int count);
int fb_run_blind_rotate(fb_ctx* ctx, const uint64_t* d_small, const uint64_t* d_luts, const uint32_t* d_lut_idx,
                        uint64_t* d_out, const int32_t* d_out_rows, int count);
int fb_run_lincomb(fb_ctx* ctx, uint64_t* d_arena, const int32_t* out_rows, const int32_t* term_off,
                   const int32_t* term_rows, const int64_t* term_coef, const uint64_t* body_const, int n_out);
