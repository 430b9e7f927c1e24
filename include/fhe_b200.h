/*
 * fhe_b200.h -- C ABI of libfhe_b200.so: the B200-native TFHE evaluation backend behind fhe-regex's
 * execution hot path.
 *
 * The reference (RKlompUU/fhe-regex) has NO FFI/plugin boundary: `Execution` holds a concrete
 * `tfhe::integer::ServerKey` (src/regex/execution.rs:38) and calls its methods.  The boundary below is
 * therefore introduced at exactly the tfhe-rs symbols the reference touches; each entry
 * point cites the reference call site it replaces.  A Rust shim binds these with `extern "C"` (see
 * INTEGRATION.md); tests and bench bind them with ctypes.
 *
 * Conventions
 *   - every function returns 0 on success or a negative FB_ERR_* code; nothing throws or aborts
 *     across the boundary; fb_last_error() gives the message of the last failure on that context.
 *   - plain pointers + sizes only.  Pointers named h_* are caller-owned HOST buffers, read/written
 *     only during the call.  Pointers named d_* are DEVICE pointers on the context's GPU.
 *   - one host thread drives one context; one context drives one GPU (one process per GPU).
 *   - there is no CPU fallback: fb_ctx_create fails without an sm_100 device.
 *
 * Ciphertext / key layouts (PARAM_MESSAGE_2_CARRY_2, 4 radix blocks; ciphertext.rs:42-45):
 *   LWE (big key)    2049 x u64 : mask[2048], body          -- a shortint block under the big key
 *   LWE (small key)   743 x u64 : mask[742], body
 *   radix ciphertext 4 x 2049 x u64, block 0 = least significant 2 bits (ciphertext.rs:8-30)
 *   KSK  [2048][5][743] u64   level rows most-significant first (tfhe-rs LweKeyswitchKey)
 *   BSK  [742][1][2][2][2048] u64 standard domain: [lwe bit][level][row][polynomial][coefficient]
 *   LUT  [2048] u64 accumulator body polynomial (tfhe-rs shortint generate_accumulator)
 */
#ifndef FHE_B200_H
#define FHE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FB_LWE_BIG_WORDS 2049
#define FB_LWE_SMALL_WORDS 743
#define FB_POLY_SIZE 2048
#define FB_RADIX_BLOCKS 4
#define FB_KSK_WORDS (2048ull * 5ull * 743ull)
#define FB_BSK_WORDS (742ull * 2ull * 2ull * 2048ull)

#define FB_OK 0
#define FB_ERR_NO_DEVICE (-1)   /* no CUDA device of compute capability 10.x */
#define FB_ERR_CUDA (-2)        /* a CUDA runtime call failed */
#define FB_ERR_ARG (-3)         /* invalid argument */
#define FB_ERR_NO_KEY (-4)      /* server key not loaded */
#define FB_ERR_PARSE (-5)       /* pattern does not parse: anyhow::Error of parse(), parser.rs:146-184 */
#define FB_ERR_PANIC (-6)       /* the reference would panic on this input (engine.rs:189-190, parser.rs:349-351) */
#define FB_ERR_FORMAT (-7)      /* malformed serialized key */

typedef struct fb_ctx fb_ctx;

/* ---- context ---------------------------------------------------------------------------------- */
/* Replaces holding a `ServerKey` by value (engine.rs:20 `Execution::new(sk.clone())`). */
int fb_ctx_create(fb_ctx** out, int device);
void fb_ctx_destroy(fb_ctx* ctx);
const char* fb_last_error(const fb_ctx* ctx);
/* cudaStream_t every kernel of this context is launched on (for external event timing / sync) */
void* fb_ctx_stream(fb_ctx* ctx);
int fb_sync(fb_ctx* ctx);

/* ---- options ---------------------------------------------------------------------------------- */
/* Every knob is per context and set explicitly; the library reads nothing from the environment.
 *   "latency_threshold"      batches up to this many PBS run the one-PBS-per-CTA blind rotation (default 296, 0 = never)
 *   "cluster_threshold"      batches up to this many PBS run the one-PBS-per-SM-pair blind rotation (default 0 = never)
 *   "br_variant"             throughput blind rotation: 0 phase-by-phase body, 1 fused body, 2 fused body with the digits through
 *                            the integer-to-double unit (default), 3 / 4 = 1 / 2 with the Fourier key in tensor memory
 *   "br_planes"              fused body at 4 PBS per CTA, transposes: 1 both components through one plane one after the other,
 *                            2 a plane per component -- the real one inside the accumulator copy (default; one barrier per
 *                            transpose), 3 = 1 with the planes inside the accumulator copies and both full inter-pass twiddle
 *                            tables in shared memory (built once per launch).  Bit-identical outputs.
 *   "br_samples"             fused body: PBS per CTA, 4 (default) or 6 (transpose planes inside the accumulator copies, 12 warps of
 *                            168 registers, 888 PBS per wave on a B200; used for batches wider than 4 per SM).  fb_pbs_batch_quantum
 *                            follows.  Bit-identical outputs.
 *   "br_stagger"             fused throughput kernel: start skew between the samples of a CTA, cycles per sample index (default 0)
 *   "br_stagger_groups"      1: the skew goes to the odd samples only (two scheduler groups, one instruction stream per scheduler)
 *   "pbs_chunks"             fb_pbs_batch pipelines large batches in chunks of whole waves: 3 = 4, rest, 4 waves (default), 5 = 1, 6,
 *                            rest, 6, 1 waves for batches of at least 24 waves (less copy exposed, more launch tails: measured slower)
 *   "plan_timing"            also prints the timeline of every pipelined fb_pbs_batch call
 *   "br_sync"                fused body: 1 = the samples of a CTA start their rotation together, behind one CTA-wide barrier after
 *                            the accumulators are read (default: their read latencies differ, and samples that leave the
 *                            initialisation apart stay apart: 2-3 % with 36 accumulators in random order), 0 = not
 *   "br_resync"              fused body: the samples of a CTA meet at a CTA-wide barrier every this many CMUX steps (default 8:
 *                            -5 % against never; 1 = every step is slower than never; 0 = never).  Bit-identical outputs.
 *   "br_barriers"            1: keep the two per-step barriers round 2 found unnecessary (A/B measurements only)
 *   "ks_variant"             keyswitch GEMM: 0 mma.sync, 1 tcgen05.mma kind::i8 with TMA operands and TMEM accumulators (default 1)
 *   "wide_skew", "wide_prefetch"   tuning of the latency kernel (defaults 200 cycles, 3 groups)
 *   "wide_pair"              two PBS per CTA (twiddles in tensor memory): 1 for batches between one and two waves of SMs
 *                            (149 .. 296 on a B200, default), 2 for every narrow batch, 0 never
 *   "wide_pair_prefetch", "wide_pair_offset"   tuning of the two-PBS-per-CTA kernel (defaults 1 group, 0 cycles)
 *   "dist_shard_min"         fb_has_match_dist: levels of at most this many PBS are bootstrapped by every rank instead of being
 *                            cut into slices and exchanged (default: the SM count, i.e. what one GPU does in one wave)
 *   "plan_reference_shaped"  1: has_match evaluates every variant the reference enumerates (default 0: implied OR
 *                            operands are absorbed; same decrypted result)
 *   "plan_timing"            1: planner phase times on stderr
 * Unknown names and out-of-range values return FB_ERR_ARG. */
int fb_set_option(fb_ctx* ctx, const char* name, int64_t value);
int fb_get_option(fb_ctx* ctx, const char* name, int64_t* value);

/* ---- server key ------------------------------------------------------------------------------- */
/* Replaces `ServerKey::new(&client_key)` / `gen_keys_radix` output being handed to has_match
 * (engine.rs:252, ciphertext.rs:44, mod.rs:16).  Uploads the KSK, converts the BSK to the Fourier
 * domain on the device (tfhe-rs convert_standard_lwe_bootstrap_key_to_fourier). */
int fb_load_server_key_raw(fb_ctx* ctx, const uint64_t* h_ksk, const uint64_t* h_bsk_std);
/* The same hand-over for a key AS THE REFERENCE HOLDS IT: a tfhe-rs 0.2.0 ServerKey keeps only the Fourier-domain
 * bootstrapping key.  h_fbsk[742][1][2][2][1024][2] f64 (re, im): [lwe bit][level][row][polynomial][frequency] in the
 * plan-independent natural frequency order in which tfhe-rs SERIALIZES a FourierLweBootstrapKey (its in-memory order is
 * the FFT plan's machine-dependent permutation; see fhe_regex_b200/csrc/wire.cpp).  That is this library's resident
 * layout, so the load is a copy: no conversion, no standard-domain key needed. */
int fb_load_server_key_fourier(fb_ctx* ctx, const uint64_t* h_ksk, const double* h_fbsk);
/* bincode::serialize(&tfhe::integer::ServerKey) of PARAM_MESSAGE_2_CARRY_2, exactly fb_server_key_bincode_size() bytes:
 * the blob a maintainer gets from the value engine.rs:252 / ciphertext.rs:44 produce.  FB_ERR_FORMAT if any length or
 * parameter differs. */
int fb_load_server_key_bincode(fb_ctx* ctx, const uint8_t* buf, size_t len);
/* `ServerKey::new(&client_key)` / `gen_keys_radix` (engine.rs:252, ciphertext.rs:44) ON the GPU: both keys are generated in
 * device memory from the client's secret key bits (KSK rows and GGSW rows encrypted by CUDA kernels, counter-based PRNG seeded by
 * `seed`, Box-Muller noise with the parameter set's standard deviations) and installed in this context; h_ksk / h_bsk_std, if
 * not NULL, receive copies in the layouts of fb_load_server_key_raw.  Statistically equivalent to, not bit-identical with, a
 * tfhe-rs key.  Milliseconds instead of the seconds of the CPU keygen. */
int fb_keygen_server_gpu(fb_ctx* ctx, const uint64_t* h_big_key, const uint64_t* h_small_key, uint64_t seed, uint64_t* h_ksk,
                         uint64_t* h_bsk_std);
/* read back the Fourier BSK ([742][2][2][1024] complex f64, natural frequency order) -- tests only */
int fb_get_fourier_bsk(fb_ctx* ctx, double* h_out);

/* ---- hot path: batched keyswitch / programmable bootstrap ------------------------------------- */
/* K1 alone, exposed for the bit-exact gate.  Under every smart_* (execution.rs:76-190):
 * tfhe-rs keyswitch_lwe_ciphertext.  h_in[count][2049] -> h_out[count][743] */
int fb_keyswitch_batch(fb_ctx* ctx, const uint64_t* h_in, size_t count, uint64_t* h_out);
/* KS -> blind rotate -> sample extract (shortint keyswitch_programmable_bootstrap), one LUT per input.
 * h_in[count][2049], h_luts[n_luts][2048], h_lut_idx[count] -> h_out[count][2049] */
int fb_pbs_batch(fb_ctx* ctx, const uint64_t* h_in, const uint64_t* h_luts, size_t n_luts, const uint32_t* h_lut_idx,
                 size_t count, uint64_t* h_out);
/* blind rotate + sample extract on already keyswitched inputs h_small[count][743] (stage-wise parity) */
int fb_bootstrap_small_batch(fb_ctx* ctx, const uint64_t* h_small, const uint64_t* h_luts, size_t n_luts,
                             const uint32_t* h_lut_idx, size_t count, uint64_t* h_out);
/* same as fb_pbs_batch on DEVICE buffers, asynchronous on fb_ctx_stream(); the context's scratch is
 * grown as needed.  d_in[count][2049], d_luts[n_luts][2048], d_lut_idx[count] -> d_out[count][2049] */
int fb_pbs_batch_dev(fb_ctx* ctx, const uint64_t* d_in, const uint64_t* d_luts, const uint32_t* d_lut_idx, size_t count,
                     uint64_t* d_out);

/* ---- regex entry point ------------------------------------------------------------------------ */
typedef struct fb_match_stats {
  uint64_t variants;        /* branches produced by build_branches over all start offsets (engine.rs:15-18) */
  uint64_t ct_ops;          /* cache-missing homomorphic ops, what engine.rs:36-40 logs */
  uint64_t cache_hits;      /* engine.rs:36-40 */
  uint64_t ops_eq, ops_gt, ops_le, ops_and, ops_or, ops_not; /* ct_ops by type */
  uint64_t pbs;             /* programmable bootstraps executed by this backend */
  uint64_t levels;          /* dependent PBS levels (kernel launch rounds) */
  uint64_t max_level_width; /* widest PBS batch */
  double gpu_ms;            /* device time of the evaluation (CUDA events on the context stream) */
} fb_match_stats;

/* Replaces `has_match(&ServerKey, &[RadixCiphertext], &str) -> Result<RadixCiphertext>` (engine.rs:8-42).
 * h_content[n_chars][4][2049] (encrypt_str layout, ciphertext.rs:32-40), pattern NUL-terminated.
 * h_out[4][2049]: block 0 encrypts 0/1, blocks 1-3 are trivial zeros, so RadixClientKey::decrypt
 * (mod.rs:17) returns the same 0/1 as the reference.  stats may be NULL. */
int fb_has_match(fb_ctx* ctx, const uint64_t* h_content, size_t n_chars, const char* pattern, uint64_t* h_out,
                 fb_match_stats* stats);
/* The same match for n_contents contents of n_chars characters each against one pattern -- what a server holding many
 * encrypted documents does with engine.rs:8-42.  h_contents[n_contents][n_chars][4][2049], h_out[n_contents][4][2049].
 * The instances share one plan and run level by level in the same launches (every PBS batch n_contents times as wide),
 * so the GPU works at its throughput rate instead of one blind-rotation latency per level and content.
 * stats: the per-content counters; gpu_ms is the total. */
int fb_has_match_many(fb_ctx* ctx, const uint64_t* h_contents, size_t n_contents, size_t n_chars, const char* pattern,
                      uint64_t* h_out, fb_match_stats* stats);
/* rank's share of the match: the rank-th of `world` contiguous slices of the final OR's operands after global
 * absorption (reference-shaped plan, option "plan_reference_shaped": the variants of start offsets i % world == rank).  The OR
 * of all ranks' results is the match result: all-gather them and fold with fb_or_fold.  stats
 * carries the reference's counters of the whole match and pbs / levels of this rank's plan. */
int fb_has_match_shard(fb_ctx* ctx, const uint64_t* h_content, size_t n_chars, const char* pattern, int rank, int world,
                       uint64_t* h_out, fb_match_stats* stats);
/* ---- multi-GPU: one process per GPU, every context a rank of one NCCL communicator (keys replicated) ------------- */
#define FB_COMM_ID_BYTES 128
/* rank 0 makes an id (an ncclUniqueId) and hands it to the other ranks by whatever means the host has (MPI, a store,
 * torch.distributed.broadcast ...); then every rank joins.  NCCL is loaded at run time; FB_ERR_NO_DEVICE if absent. */
int fb_comm_unique_id(uint8_t* id /* [FB_COMM_ID_BYTES] */);
int fb_comm_init(fb_ctx* ctx, const uint8_t* id, int rank, int world);
int fb_comm_destroy(fb_ctx* ctx);
int fb_comm_info(fb_ctx* ctx, int* rank, int* world);
/* has_match (engine.rs:8-42) across the communicator -- COLLECTIVE: every rank calls it with the same content and
 * pattern, every rank gets the result in h_out.  Each PBS level of the plan is cut into `world` contiguous slices, rank r
 * bootstraps slice r, and the slices are exchanged device to device over NVLink (grouped ncclBroadcast, in place in the
 * ciphertext arena, on the context stream): no host round trip between levels, and the final bitor fold of the reference
 * (engine.rs:22-35) is simply the last levels of the same plan.  Rank r reads only slice r of h_content (the content is
 * uploaded over PCIe once in total and gathered over NVLink).  stats: the plan's counters (whole match) and this rank's
 * device time. */
int fb_has_match_dist(fb_ctx* ctx, const uint64_t* h_content, size_t n_chars, const char* pattern, uint64_t* h_out,
                      fb_match_stats* stats);

/* OR of n single-block booleans h_in[n][2049] -> h_out[4][2049] radix (the final fold, engine.rs:30-33) */
int fb_or_fold(fb_ctx* ctx, const uint64_t* h_in, size_t n, uint64_t* h_out);

/* Parser surface kept from the reference: parse() (parser.rs:146).  Writes the `{:?}` rendering of
 * the RegExpr (parser.rs:87-144) into out (NUL-terminated, truncated to cap).  No GPU needed. */
int fb_parse_debug(const char* pattern, char* out, size_t cap);
/* flags of the host-only planner entry points below (a context carries the same choice as option
 * "plan_reference_shaped") */
#define FB_PLAN_REFERENCE_SHAPED 1u /* evaluate every variant the reference enumerates: no absorption of implied OR operands */
/* plaintext dry run of the variant generator + executor bookkeeping (no ciphertexts, no GPU):
 * fills variants / ct_ops / cache_hits / ops_* / pbs / levels / max_level_width. */
int fb_plan_stats(const char* pattern, size_t n_chars, uint32_t flags, fb_match_stats* stats);

/* PBS batch width of every level of the lowered plan (host only): returns the number of levels (>= 0, may
 * exceed cap; only cap entries are written) or a negative error code */
int fb_plan_level_widths(const char* pattern, size_t n_chars, int rank, int world, uint32_t flags, int32_t* widths, size_t cap);

/* plaintext dry run of the lowered circuit on cleartext content bytes (host only, no ciphertexts):
 * result = what decrypt(has_match(..)) would give for this rank's share; used to test the lowering. */
int fb_plan_eval_plain(const char* pattern, const uint8_t* content, size_t n_chars, int rank, int world, uint32_t flags, int* result);

/* ---- op-level boundary: device-resident arenas for a host that keeps its own executor ---------------------------- */
/* For a maintainer who keeps `Execution` (execution.rs:37-223: the six smart_* call sites :76,93,110,143,173,190 and the
 * structural cache with_cache :212-222) and only swaps the arithmetic: ciphertexts live in an arena on the GPU, a level
 * of cache-missing ops is flushed as one fb_lincomb (packing, sums of booleans, NOT) plus one fb_pbs_rows (KS -> BR -> SE
 * through per-row LUTs), and nothing returns to the host until the result is downloaded.  All calls are asynchronous on
 * fb_ctx_stream() except fb_ct_download. */
typedef uint64_t fb_handle;   /* 0 is never a valid handle */
#define FB_REGEX_LUTS 51      /* rows of the accumulator table of fb_has_match (fb_regex_lut_table) */
/* rows x row_words u64 on the device; row_words = 2049 (LWE ciphertexts under the big key) or 2048 (LUT polynomials) */
int fb_ct_alloc(fb_ctx* ctx, size_t rows, size_t row_words, fb_handle* out);
int fb_ct_free(fb_ctx* ctx, fb_handle h);
int fb_ct_upload(fb_ctx* ctx, fb_handle h, size_t first_row, const uint64_t* h_rows, size_t count);
int fb_ct_download(fb_ctx* ctx, fb_handle h, size_t first_row, size_t count, uint64_t* h_rows);   /* synchronizes */
/* arena[out_rows[o]] = sum over t in [term_off[o], term_off[o+1]) of term_coef[t] * arena[term_rows[t]]
 *                      + trivial(body_const[o])         -- bivariate packing, k-ary sums, 1 - x; o < n_out */
int fb_lincomb(fb_ctx* ctx, fb_handle h, const int32_t* h_out_rows, const int32_t* h_term_off, const int32_t* h_term_rows,
               const int64_t* h_term_coef, const uint64_t* h_body_const, size_t n_out);
/* arena[out_row_base + b] = PBS(arena[in_rows[b]], luts[lut_idx[b]]), b < count  (what every smart_* block op is) */
int fb_pbs_rows(fb_ctx* ctx, fb_handle h, const int32_t* h_in_rows, fb_handle luts, const uint32_t* h_lut_idx, size_t count,
                size_t out_row_base);
/* The library's own level-synchronous plan of a match as a stream of int64 (layout: handles.cu), so that a host can
 * drive or check it level by level with the calls above.  out == NULL: *n_words = required length.  Host only. */
int fb_plan_export(const char* pattern, size_t n_chars, uint32_t flags, int64_t* out, size_t cap, size_t* n_words);
/* the accumulator table fb_has_match bootstraps through: h_out[FB_REGEX_LUTS][2048].  Host only. */
int fb_regex_lut_table(uint64_t* h_out);

/* ---- timing ----------------------------------------------------------------------------------- */
typedef struct fb_kernel_stats {
  uint64_t ks_launches, br_launches, lin_launches;
  uint64_t ks_samples, br_samples;
  double ks_ms, br_ms, lin_ms; /* summed CUDA-event durations on the context stream */
} fb_kernel_stats;
/* per-kernel CUDA-event timing of every launch since the last reset (events are resolved at call time,
 * which synchronizes the stream) */
int fb_kernel_stats_reset(fb_ctx* ctx);
int fb_kernel_stats_get(fb_ctx* ctx, fb_kernel_stats* out);
int fb_kernel_timing_enable(fb_ctx* ctx, int on);
/* FP64 FMA-pipe throughput of the device in TFLOP/s (dependent-chain DFMA probe, best of reps): the
 * roofline denominator of the blind rotation */
int fb_measure_fp64_peak(fb_ctx* ctx, int reps, double* tflops);
/* Batches of up to max_count PBS run the latency variant of the blind rotation (one PBS per CTA, br_wide.cu);
 * larger ones the throughput variant (up to 4 PBS per CTA).  Default 296 (two waves of 148 SMs), 0 = never.
 * Returns the previous value (or a negative error code).  Both variants compute the same function. */
int fb_set_latency_threshold(fb_ctx* ctx, int max_count);
/* Batches of up to max_count PBS run the cluster variant (one PBS per pair of SMs, br_duo.cu).  Default 0 = never:
 * measured on B200 it is no faster than the one-PBS-per-CTA kernel.  Returns the previous value. */
int fb_set_cluster_threshold(fb_ctx* ctx, int max_count);
/* PBS batch sizes that fill the GPU evenly are multiples of this (SM count x samples per CTA) */
int fb_pbs_batch_quantum(fb_ctx* ctx);

/* ---- page-locked host buffers ------------------------------------------------------------------- */
/* Every entry point takes plain caller-owned host pointers (the reference hands `&[RadixCiphertext]` to has_match,
 * engine.rs:8-12, out of the `Vec` encrypt_str built, ciphertext.rs:32-40).  A buffer obtained here is page-locked: its
 * uploads run at PCIe speed and truly asynchronously (a 64-character content is 4.2 MB, a 256-character one 16.8 MB --
 * 0.3 / 1.2 ms from pageable memory against 0.1 / 0.4 ms from page-locked memory).  Optional: pageable buffers work everywhere.
 * fb_host_alloc returns FB_ERR_CUDA when no device is present. */
int fb_host_alloc(size_t bytes, void** out);
void fb_host_free(void* p);

/* ---- tfhe-rs 0.2.0 wire formats (host only; layouts in fhe_regex_b200/csrc/wire.cpp) ---------------------------- */
/* integer::ServerKey <-> (keyswitch key container, Fourier bootstrapping key in serialized order) */
size_t fb_server_key_bincode_size(void);
int fb_server_key_from_bincode(const uint8_t* buf, size_t len, uint64_t* h_ksk, double* h_fbsk);
/* out == NULL: *written = required size */
int fb_server_key_to_bincode(const uint64_t* h_ksk, const double* h_fbsk, uint8_t* out, size_t cap, size_t* written);
/* integer::RadixCiphertext (ciphertext.rs:6,29): 4 shortint blocks; degrees[4] may be NULL (written as 3 = fresh) */
size_t fb_radix_bincode_size(void);
int fb_radix_from_bincode(const uint8_t* buf, size_t len, uint64_t* h_ct, uint64_t* degrees);
int fb_radix_to_bincode(const uint64_t* h_ct, const uint64_t* degrees, uint8_t* out, size_t cap, size_t* written);
/* StringCiphertext = Vec<RadixCiphertext> (ciphertext.rs:6, :32-40) <-> h_content[n_chars][4][2049], the layout
 * fb_has_match takes.  h_content == NULL: *n_chars = number of characters in the blob.  Blocks must be fresh
 * (degree <= 3), as encrypt_str produces them. */
int fb_string_ciphertext_from_bincode(const uint8_t* buf, size_t len, uint64_t* h_content, size_t cap_chars, size_t* n_chars);
int fb_string_ciphertext_to_bincode(const uint64_t* h_content, size_t n_chars, uint8_t* out, size_t cap, size_t* written);

/* ---- client-side glue (tests / bench / demo only; not on the server hot path) ------------------ */
/* Deserialize a bincode RadixClientKey like test_data/client_key (engine.rs:248-251).
 * Writes big[2048], small[742] secret key bits. */
int fb_client_key_from_bincode(const uint8_t* buf, size_t len, uint64_t* big_key, uint64_t* small_key);
/* ServerKey::new(&client_key) (engine.rs:252): fresh KSK/BSK from the secret keys, seeded test PRNG */
int fb_client_keygen_server(const uint64_t* big_key, const uint64_t* small_key, uint64_t seed, uint64_t* h_ksk,
                            uint64_t* h_bsk_std);
/* RadixClientKey::encrypt per byte (ciphertext.rs:32-40): h_out[n][4][2049]; fails on non-ASCII */
int fb_client_encrypt_str(const uint64_t* big_key, const uint8_t* bytes, size_t n, uint64_t seed, uint64_t* h_out);
/* create_trivial_radix per byte (ciphertext.rs:8-30), what the reference's tests feed (engine.rs:282-286) */
int fb_client_trivial_str(const uint8_t* bytes, size_t n, uint64_t* h_out);
/* shortint block: encrypt message m (4 bits, message+carry) / phase / decrypt */
int fb_client_encrypt_block(const uint64_t* big_key, uint64_t m, uint64_t seed, uint64_t stream, uint64_t* h_out);
uint64_t fb_client_phase(const uint64_t* key, size_t dim, const uint64_t* ct);
uint64_t fb_client_decrypt_block(const uint64_t* big_key, const uint64_t* ct);
/* RadixClientKey::decrypt (mod.rs:17): sum of blocks * 4^i mod 256 */
uint64_t fb_client_decrypt_radix(const uint64_t* big_key, const uint64_t* ct);
/* shortint generate_accumulator: f16[i] = f(i) for i < 16 -> lut[2048] */
int fb_make_lut(const uint64_t* f16, uint64_t* lut);

#ifdef __cplusplus
}
#endif
#endif /* FHE_B200_H */
