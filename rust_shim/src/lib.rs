//! Rust face of libfhe_b200.so for RKlompUU/fhe-regex (UNCOMPILED HERE: the image has no Rust toolchain).
//!
//! Keeps the reference's surface: `has_match(&ServerKey, &[RadixCiphertext], &str) -> Result<RadixCiphertext>`
//! (src/regex/engine.rs:8-42), with the server key living on the GPU behind `B200ServerKey`
//! (replaces `Execution::new(sk.clone())`, engine.rs:20).  Every `extern "C"` item mirrors include/fhe_b200.h.
use std::ffi::{CStr, CString};
use std::os::raw::{c_char, c_int};

use anyhow::{anyhow, bail, Result};
use log::info;

pub const FB_LWE_BIG_WORDS: usize = 2049; // one shortint block under the big LWE key (mask 2048 + body)
pub const FB_RADIX_BLOCKS: usize = 4; // 8-bit character = 4 blocks of 2 bits (ciphertext.rs:42-45)

#[repr(C)]
pub struct FbCtx {
    _private: [u8; 0],
}

#[repr(C)]
#[derive(Default, Debug, Clone, Copy)]
pub struct FbMatchStats {
    pub variants: u64,
    pub ct_ops: u64,
    pub cache_hits: u64,
    pub ops_eq: u64,
    pub ops_gt: u64,
    pub ops_le: u64,
    pub ops_and: u64,
    pub ops_or: u64,
    pub ops_not: u64,
    pub pbs: u64,
    pub levels: u64,
    pub max_level_width: u64,
    pub gpu_ms: f64,
}

extern "C" {
    fn fb_ctx_create(out: *mut *mut FbCtx, device: c_int) -> c_int;
    fn fb_ctx_destroy(ctx: *mut FbCtx);
    fn fb_last_error(ctx: *const FbCtx) -> *const c_char;
    fn fb_load_server_key_raw(ctx: *mut FbCtx, h_ksk: *const u64, h_bsk_std: *const u64) -> c_int;
    fn fb_load_server_key_bincode(ctx: *mut FbCtx, buf: *const u8, len: usize) -> c_int;
    fn fb_comm_unique_id(id: *mut u8) -> c_int;
    fn fb_comm_init(ctx: *mut FbCtx, id: *const u8, rank: c_int, world: c_int) -> c_int;
    fn fb_has_match_dist(ctx: *mut FbCtx, h_content: *const u64, n_chars: usize, pattern: *const c_char, h_out: *mut u64,
                         stats: *mut FbMatchStats) -> c_int;
    fn fb_has_match(ctx: *mut FbCtx, h_content: *const u64, n_chars: usize, pattern: *const c_char, h_out: *mut u64,
                    stats: *mut FbMatchStats) -> c_int;
    fn fb_has_match_shard(ctx: *mut FbCtx, h_content: *const u64, n_chars: usize, pattern: *const c_char, rank: c_int,
                          world: c_int, h_out: *mut u64, stats: *mut FbMatchStats) -> c_int;
    fn fb_or_fold(ctx: *mut FbCtx, h_in: *const u64, n: usize, h_out: *mut u64) -> c_int;
    fn fb_set_option(ctx: *mut FbCtx, name: *const c_char, value: i64) -> c_int;
    fn fb_host_alloc(bytes: usize, out: *mut *mut std::ffi::c_void) -> c_int;
    fn fb_host_free(p: *mut std::ffi::c_void);
}

/// A page-locked staging buffer of the library (`fb_host_alloc`): content flattened into it uploads at PCIe speed
/// (a 256-character content is 16.8 MB: 0.4 ms instead of 1.2 ms from a `Vec`).  Optional -- every entry point takes any host slice.
pub struct PinnedWords {
    ptr: *mut u64,
    len: usize,
}

impl PinnedWords {
    pub fn new(len: usize) -> Result<Self> {
        let mut p: *mut std::ffi::c_void = std::ptr::null_mut();
        if unsafe { fb_host_alloc(len * 8, &mut p) } != 0 || p.is_null() {
            bail!("fb_host_alloc failed (no device?)");
        }
        Ok(Self { ptr: p as *mut u64, len })
    }
    pub fn as_mut_slice(&mut self) -> &mut [u64] {
        unsafe { std::slice::from_raw_parts_mut(self.ptr, self.len) }
    }
    pub fn as_slice(&self) -> &[u64] {
        unsafe { std::slice::from_raw_parts(self.ptr, self.len) }
    }
}

impl Drop for PinnedWords {
    fn drop(&mut self) {
        unsafe { fb_host_free(self.ptr as *mut std::ffi::c_void) }
    }
}

const FB_ERR_PARSE: c_int = -5; // Err of parse() (parser.rs:146-184)
const FB_ERR_PANIC: c_int = -6; // inputs on which the reference panics (engine.rs:189-190, parser.rs:349-351)

/// The server key resident on one B200: keyswitch key as byte planes, bootstrapping key in the Fourier domain.
pub struct B200ServerKey {
    ctx: *mut FbCtx,
}

impl B200ServerKey {
    /// `ksk`: LweKeyswitchKey container [2048][5][743] u64 (level rows most significant first);
    /// `bsk_std`: standard-domain LweBootstrapKey container [742][1][2][2][2048] u64 -- the two containers
    /// `ServerKey::new(&client_key)` builds (engine.rs:252) before converting the latter to Fourier.
    pub fn from_raw(device: i32, ksk: &[u64], bsk_std: &[u64]) -> Result<Self> {
        assert_eq!(ksk.len(), 2048 * 5 * 743);
        assert_eq!(bsk_std.len(), 742 * 2 * 2 * 2048);
        let mut ctx = std::ptr::null_mut();
        let rc = unsafe { fb_ctx_create(&mut ctx, device) };
        if rc != 0 {
            bail!("libfhe_b200: {} (no CPU fallback)", last_error(std::ptr::null()));
        }
        let rc = unsafe { fb_load_server_key_raw(ctx, ksk.as_ptr(), bsk_std.as_ptr()) };
        if rc != 0 {
            let msg = last_error(ctx);
            unsafe { fb_ctx_destroy(ctx) };
            bail!("libfhe_b200: {msg}");
        }
        Ok(Self { ctx })
    }
}

impl B200ServerKey {
    /// From the reference's own keygen output (`gen_keys()` ciphertext.rs:42-45, `ServerKey::new(&client_key)` engine.rs:252):
    /// tfhe-rs 0.2.0 keeps the bootstrapping key in the Fourier domain only and serializes it in a plan-independent natural
    /// frequency order, which is the library's resident layout (fhe_regex_b200/csrc/wire.cpp).
    pub fn new(device: i32, sk: &tfhe::integer::ServerKey) -> Result<Self> {
        let blob = bincode::serialize(sk)?;
        let mut ctx = std::ptr::null_mut();
        if unsafe { fb_ctx_create(&mut ctx, device) } != 0 {
            bail!("libfhe_b200: {} (no CPU fallback)", last_error(std::ptr::null()));
        }
        if unsafe { fb_load_server_key_bincode(ctx, blob.as_ptr(), blob.len()) } != 0 {
            let msg = last_error(ctx);
            unsafe { fb_ctx_destroy(ctx) };
            bail!("libfhe_b200: {msg}");
        }
        Ok(Self { ctx })
    }

    /// Join the NCCL communicator of `id` (128 bytes from `comm_unique_id()` on rank 0, handed over by the host's own means).
    /// per-context knob (include/fhe_b200.h lists them, e.g. "plan_reference_shaped", "latency_threshold")
    pub fn set_option(&self, name: &str, value: i64) -> Result<()> {
        let n = CString::new(name)?;
        check(self.ctx, unsafe { fb_set_option(self.ctx, n.as_ptr(), value) })
    }
    pub fn comm_init(&self, id: &[u8; 128], rank: i32, world: i32) -> Result<()> {
        check(self.ctx, unsafe { fb_comm_init(self.ctx, id.as_ptr(), rank, world) })
    }
}

pub fn comm_unique_id() -> Result<[u8; 128]> {
    let mut id = [0u8; 128];
    if unsafe { fb_comm_unique_id(id.as_mut_ptr()) } != 0 {
        bail!("libfhe_b200: NCCL unavailable");
    }
    Ok(id)
}

/// has_match across the communicator (collective: every rank calls it with the same content and pattern, every rank gets
/// the result): every PBS level of the plan is cut into `world` slices exchanged over NVLink inside the library.
pub fn has_match_dist_flat(sk: &B200ServerKey, content: &[u64], n_chars: usize, pattern: &str) -> Result<Vec<u64>> {
    let mut out = vec![0u64; FB_RADIX_BLOCKS * FB_LWE_BIG_WORDS];
    let pat = CString::new(pattern)?;
    let rc = unsafe { fb_has_match_dist(sk.ctx, content.as_ptr(), n_chars, pat.as_ptr(), out.as_mut_ptr(), std::ptr::null_mut()) };
    check(sk.ctx, rc)?;
    Ok(out)
}

impl Drop for B200ServerKey {
    fn drop(&mut self) {
        unsafe { fb_ctx_destroy(self.ctx) }
    }
}

fn last_error(ctx: *const FbCtx) -> String {
    unsafe { CStr::from_ptr(fb_last_error(ctx)) }.to_string_lossy().into_owned()
}

/// `[n_chars][4][2049]` u64, block 0 least significant: the layout of `encrypt_str` (ciphertext.rs:32-40).
pub type FlatContent = Vec<u64>;

fn check(ctx: *const FbCtx, rc: c_int) -> Result<()> {
    match rc {
        0 => Ok(()),
        FB_ERR_PARSE => Err(anyhow!("{}", last_error(ctx))), // `let re = parse(pattern)?;` (engine.rs:13)
        FB_ERR_PANIC => panic!("{}", last_error(ctx)),       // the reference panics on these inputs too
        _ => Err(anyhow!("libfhe_b200: {}", last_error(ctx))),
    }
}

/// has_match (engine.rs:8-42) on flattened ciphertexts.  Returns the 4 blocks of the result radix ciphertext:
/// block 0 encrypts 0/1, blocks 1-3 are trivial zeros, so `RadixClientKey::decrypt` (mod.rs:17) gives 0/1.
pub fn has_match_flat(sk: &B200ServerKey, content: &[u64], n_chars: usize, pattern: &str) -> Result<(Vec<u64>, FbMatchStats)> {
    assert_eq!(content.len(), n_chars * FB_RADIX_BLOCKS * FB_LWE_BIG_WORDS);
    let mut out = vec![0u64; FB_RADIX_BLOCKS * FB_LWE_BIG_WORDS];
    let mut stats = FbMatchStats::default();
    let pat = CString::new(pattern)?;
    let rc = unsafe { fb_has_match(sk.ctx, content.as_ptr(), n_chars, pat.as_ptr(), out.as_mut_ptr(), &mut stats) };
    check(sk.ctx, rc)?;
    info!("{} ciphertext operations, {} cache hits", stats.ct_ops, stats.cache_hits); // engine.rs:36-40
    Ok((out, stats))
}

/// One rank's share of a match sharded over `world` GPUs without NCCL: the rank-th contiguous slice (operands sorted by
/// content position, `[n r / world, n (r + 1) / world)`) of the final OR's operands, after global absorption in the default plan
/// and of every enumerated variant in the reference-shaped one.  Gather the first 2049 words of every rank's result and
/// finish with `or_fold`.
pub fn has_match_shard_flat(sk: &B200ServerKey, content: &[u64], n_chars: usize, pattern: &str, rank: i32, world: i32) -> Result<Vec<u64>> {
    let mut out = vec![0u64; FB_RADIX_BLOCKS * FB_LWE_BIG_WORDS];
    let pat = CString::new(pattern)?;
    let rc = unsafe {
        fb_has_match_shard(sk.ctx, content.as_ptr(), n_chars, pat.as_ptr(), rank, world, out.as_mut_ptr(), std::ptr::null_mut())
    };
    check(sk.ctx, rc)?;
    Ok(out)
}

/// The final bitor fold (engine.rs:30-33) over the gathered per-rank booleans (`parts`: n x 2049 u64).
pub fn or_fold(sk: &B200ServerKey, parts: &[u64], n: usize) -> Result<Vec<u64>> {
    let mut out = vec![0u64; FB_RADIX_BLOCKS * FB_LWE_BIG_WORDS];
    check(sk.ctx, unsafe { fb_or_fold(sk.ctx, parts.as_ptr(), n, out.as_mut_ptr()) })?;
    Ok(out)
}

/// The reference's signature.  Flattening a `RadixCiphertext` into its blocks' LWE containers (and back) uses
/// tfhe-rs 0.2.0 accessors (`blocks()`, `ct.as_ref()`, `RadixCiphertext::from(Vec<shortint::Ciphertext>)`).
pub fn has_match(sk: &B200ServerKey, content: &[tfhe::integer::RadixCiphertext], pattern: &str) -> Result<tfhe::integer::RadixCiphertext> {
    let mut flat: FlatContent = Vec::with_capacity(content.len() * FB_RADIX_BLOCKS * FB_LWE_BIG_WORDS);
    for ch in content {
        for block in ch.blocks() {
            flat.extend_from_slice(block.ct.as_ref());
        }
    }
    let (out, _stats) = has_match_flat(sk, &flat, content.len(), pattern)?;
    Ok(radix_from_flat(&out))
}

fn radix_from_flat(words: &[u64]) -> tfhe::integer::RadixCiphertext {
    use tfhe::core_crypto::prelude::LweCiphertextOwned;
    use tfhe::shortint::{ciphertext::Degree, CarryModulus, Ciphertext, MessageModulus};
    let blocks: Vec<Ciphertext> = words
        .chunks_exact(FB_LWE_BIG_WORDS)
        .enumerate()
        .map(|(i, w)| Ciphertext {
            ct: LweCiphertextOwned::from_container(w.to_vec()),
            degree: Degree(if i == 0 { 1 } else { 0 }),
            message_modulus: MessageModulus(4),
            carry_modulus: CarryModulus(4),
        })
        .collect();
    tfhe::integer::RadixCiphertext::from(blocks)
}
