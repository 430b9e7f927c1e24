"""fhe_regex_b200 -- B200-native TFHE evaluation backend for fhe-regex's execution hot path.

Thin ctypes binding of the C ABI in include/fhe_b200.h (libfhe_b200.so: hand-written sm_100a CUDA
kernels + C++ host).  Names follow the reference: `has_match(server_key, content, pattern)`
(src/regex/engine.rs:8), `encrypt_str` / `gen_keys` (src/regex/ciphertext.rs:32-45), `parse`
(src/regex/parser.rs:146).  There is no CPU fallback: compute entry points raise without a B200.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libfhe_b200.so")

BIG, SMALL, POLY = 2049, 743, 2048
KSK_WORDS = 2048 * 5 * 743
BSK_WORDS = 742 * 2 * 2 * 2048

FB_OK, FB_ERR_NO_DEVICE, FB_ERR_CUDA, FB_ERR_ARG, FB_ERR_NO_KEY, FB_ERR_PARSE, FB_ERR_PANIC, FB_ERR_FORMAT = 0, -1, -2, -3, -4, -5, -6, -7


class FbError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("libfhe_b200 error %d: %s" % (code, msg))
        self.code = code


class ParseError(FbError):
    """anyhow::Error of parse() (parser.rs:146-184)."""


class ReferencePanic(FbError):
    """Input on which the reference panics (engine.rs:189-190, parser.rs:349-351)."""


class MatchStats(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in ("variants", "ct_ops", "cache_hits", "ops_eq", "ops_gt", "ops_le", "ops_and",
                                          "ops_or", "ops_not", "pbs", "levels", "max_level_width")] + [("gpu_ms", C.c_double)]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


class KernelStats(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in ("ks_launches", "br_launches", "lin_launches", "ks_samples", "br_samples")] + \
               [(n, C.c_double) for n in ("ks_ms", "br_ms", "lin_ms")]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


_lib = None


def lib():
    """Load libfhe_b200.so (building it first if the sources are newer).  Fails loudly if missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        from . import build as _b
        _b.build()
    L = C.CDLL(LIB_PATH)
    vp, u64p, u32p, u8p = C.c_void_p, C.POINTER(C.c_uint64), C.POINTER(C.c_uint32), C.POINTER(C.c_uint8)
    sz = C.c_size_t
    L.fb_ctx_create.argtypes = [C.POINTER(vp), C.c_int]
    L.fb_ctx_destroy.argtypes = [vp]
    L.fb_ctx_destroy.restype = None
    L.fb_last_error.argtypes = [vp]
    L.fb_last_error.restype = C.c_char_p
    L.fb_ctx_stream.argtypes = [vp]
    L.fb_ctx_stream.restype = vp
    L.fb_sync.argtypes = [vp]
    L.fb_load_server_key_raw.argtypes = [vp, vp, vp]
    L.fb_get_fourier_bsk.argtypes = [vp, vp]
    L.fb_keyswitch_batch.argtypes = [vp, vp, sz, vp]
    L.fb_pbs_batch.argtypes = [vp, vp, vp, sz, vp, sz, vp]
    L.fb_bootstrap_small_batch.argtypes = [vp, vp, vp, sz, vp, sz, vp]
    L.fb_pbs_batch_dev.argtypes = [vp, vp, vp, vp, sz, vp]
    L.fb_has_match.argtypes = [vp, vp, sz, C.c_char_p, vp, C.POINTER(MatchStats)]
    L.fb_has_match_shard.argtypes = [vp, vp, sz, C.c_char_p, C.c_int, C.c_int, vp, C.POINTER(MatchStats)]
    L.fb_has_match_many.argtypes = [vp, vp, sz, sz, C.c_char_p, vp, C.POINTER(MatchStats)]
    L.fb_or_fold.argtypes = [vp, vp, sz, vp]
    L.fb_parse_debug.argtypes = [C.c_char_p, C.c_char_p, sz]
    L.fb_plan_stats.argtypes = [C.c_char_p, sz, C.c_uint32, C.POINTER(MatchStats)]
    L.fb_plan_level_widths.argtypes = [C.c_char_p, sz, C.c_int, C.c_int, C.c_uint32, vp, sz]
    L.fb_plan_eval_plain.argtypes = [C.c_char_p, vp, sz, C.c_int, C.c_int, C.c_uint32, C.POINTER(C.c_int)]
    L.fb_keygen_server_gpu.argtypes = [vp, vp, vp, C.c_uint64, vp, vp]
    L.fb_load_server_key_fourier.argtypes = [vp, vp, vp]
    L.fb_load_server_key_bincode.argtypes = [vp, vp, sz]
    L.fb_server_key_bincode_size.restype = sz
    L.fb_server_key_bincode_size.argtypes = []
    L.fb_server_key_from_bincode.argtypes = [vp, sz, vp, vp]
    L.fb_server_key_to_bincode.argtypes = [vp, vp, vp, sz, C.POINTER(sz)]
    L.fb_radix_bincode_size.restype = sz
    L.fb_radix_bincode_size.argtypes = []
    L.fb_radix_from_bincode.argtypes = [vp, sz, vp, vp]
    L.fb_radix_to_bincode.argtypes = [vp, vp, vp, sz, C.POINTER(sz)]
    L.fb_string_ciphertext_from_bincode.argtypes = [vp, sz, vp, sz, C.POINTER(sz)]
    L.fb_string_ciphertext_to_bincode.argtypes = [vp, sz, vp, sz, C.POINTER(sz)]
    L.fb_ct_alloc.argtypes = [vp, sz, sz, C.POINTER(C.c_uint64)]
    L.fb_ct_free.argtypes = [vp, C.c_uint64]
    L.fb_ct_upload.argtypes = [vp, C.c_uint64, sz, vp, sz]
    L.fb_ct_download.argtypes = [vp, C.c_uint64, sz, sz, vp]
    L.fb_lincomb.argtypes = [vp, C.c_uint64, vp, vp, vp, vp, vp, sz]
    L.fb_pbs_rows.argtypes = [vp, C.c_uint64, vp, C.c_uint64, vp, sz, sz]
    L.fb_plan_export.argtypes = [C.c_char_p, sz, C.c_uint32, vp, sz, C.POINTER(sz)]
    L.fb_regex_lut_table.argtypes = [vp]
    L.fb_comm_unique_id.argtypes = [vp]
    L.fb_comm_init.argtypes = [vp, vp, C.c_int, C.c_int]
    L.fb_comm_destroy.argtypes = [vp]
    L.fb_comm_info.argtypes = [vp, C.POINTER(C.c_int), C.POINTER(C.c_int)]
    L.fb_has_match_dist.argtypes = [vp, vp, sz, C.c_char_p, vp, C.POINTER(MatchStats)]
    L.fb_set_option.argtypes = [vp, C.c_char_p, C.c_int64]
    L.fb_get_option.argtypes = [vp, C.c_char_p, C.POINTER(C.c_int64)]
    L.fb_kernel_stats_reset.argtypes = [vp]
    L.fb_set_latency_threshold.argtypes = [vp, C.c_int]
    L.fb_set_cluster_threshold.argtypes = [vp, C.c_int]
    L.fb_kernel_stats_get.argtypes = [vp, C.POINTER(KernelStats)]
    L.fb_kernel_timing_enable.argtypes = [vp, C.c_int]
    L.fb_measure_fp64_peak.argtypes = [vp, C.c_int, C.POINTER(C.c_double)]
    L.fb_pbs_batch_quantum.argtypes = [vp]
    L.fb_host_alloc.argtypes = [sz, C.POINTER(vp)]
    L.fb_host_free.argtypes = [vp]
    L.fb_host_free.restype = None
    L.fb_client_key_from_bincode.argtypes = [vp, sz, vp, vp]
    L.fb_client_keygen_server.argtypes = [vp, vp, C.c_uint64, vp, vp]
    L.fb_client_encrypt_str.argtypes = [vp, vp, sz, C.c_uint64, vp]
    L.fb_client_trivial_str.argtypes = [vp, sz, vp]
    L.fb_client_encrypt_block.argtypes = [vp, C.c_uint64, C.c_uint64, C.c_uint64, vp]
    L.fb_client_phase.argtypes = [vp, sz, vp]
    L.fb_client_phase.restype = C.c_uint64
    L.fb_client_decrypt_block.argtypes = [vp, vp]
    L.fb_client_decrypt_block.restype = C.c_uint64
    L.fb_client_decrypt_radix.argtypes = [vp, vp]
    L.fb_client_decrypt_radix.restype = C.c_uint64
    L.fb_make_lut.argtypes = [vp, vp]
    _lib = L
    return L


def _p(a: np.ndarray):
    assert a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(C.c_void_p)


def _raise(code, msg):
    if code == FB_ERR_PARSE:
        raise ParseError(code, msg)
    if code == FB_ERR_PANIC:
        raise ReferencePanic(code, msg)
    raise FbError(code, msg)


# -------------------------------------------------------------------------------------------------
# host-only surface (no GPU needed)
# -------------------------------------------------------------------------------------------------
def parse(pattern: str) -> str:
    """parse() of the reference (parser.rs:146); returns the `{:?}` rendering of the RegExpr."""
    buf = C.create_string_buffer(1 << 16)
    rc = lib().fb_parse_debug(pattern.encode("latin-1"), buf, len(buf))
    if rc != FB_OK:
        _raise(rc, buf.value.decode("latin-1"))
    return buf.value.decode("latin-1")


PLAN_REFERENCE_SHAPED = 1   # FB_PLAN_REFERENCE_SHAPED: evaluate every variant the reference enumerates


def plan_stats(pattern: str, n_chars: int, reference_shaped: bool = False) -> dict:
    st = MatchStats()
    rc = lib().fb_plan_stats(pattern.encode("latin-1"), n_chars, PLAN_REFERENCE_SHAPED if reference_shaped else 0, C.byref(st))
    if rc != FB_OK:
        _raise(rc, "plan failed for %r" % pattern)
    return st.as_dict()


def plan_level_widths(pattern: str, n_chars: int, rank: int = 0, world: int = 1, reference_shaped: bool = False) -> list:
    """PBS batch width of every level of the lowered plan (host only)"""
    buf = np.zeros(256, dtype=np.int32)
    n = lib().fb_plan_level_widths(pattern.encode(), n_chars, rank, world, PLAN_REFERENCE_SHAPED if reference_shaped else 0, _p(buf), buf.size)
    if n < 0:
        _raise(n, "plan_level_widths")
    return [int(x) for x in buf[:min(n, buf.size)]]


def plan_eval_plain(pattern: str, content: str, rank: int = 0, world: int = 1, reference_shaped: bool = False) -> int:
    """Dry run of the lowered PBS circuit on cleartext bytes (host only): the 0/1 decrypt would give."""
    b = content.encode("latin-1")
    raw = np.frombuffer(b, dtype=np.uint8) if len(b) else np.zeros(1, dtype=np.uint8)
    res = C.c_int(-1)
    rc = lib().fb_plan_eval_plain(pattern.encode("latin-1"), _p(np.ascontiguousarray(raw)), len(b), rank, world,
                                  PLAN_REFERENCE_SHAPED if reference_shaped else 0, C.byref(res))
    if rc != FB_OK:
        _raise(rc, "plan dry run failed for %r" % pattern)
    return res.value


# ---- tfhe-rs 0.2.0 wire formats (fhe_regex_b200/csrc/wire.cpp) ---------------------------------------------------
def server_key_to_bincode(ksk: np.ndarray, fourier_bsk: np.ndarray) -> bytes:
    """bincode::serialize(&integer::ServerKey) from the keyswitch container and the Fourier key in serialized order"""
    ksk = np.ascontiguousarray(ksk, dtype=np.uint64)
    f = np.ascontiguousarray(fourier_bsk, dtype=np.float64)
    assert ksk.size == KSK_WORDS and f.size == 742 * 4 * 1024 * 2
    out = np.empty(lib().fb_server_key_bincode_size(), dtype=np.uint8)
    n = C.c_size_t(0)
    rc = lib().fb_server_key_to_bincode(_p(ksk), _p(f), _p(out), out.size, C.byref(n))
    if rc != FB_OK or n.value != out.size:
        _raise(rc, "server key serialization failed")
    return out.tobytes()


def server_key_from_bincode(blob: bytes):
    buf = np.frombuffer(blob, dtype=np.uint8)
    ksk = np.empty(KSK_WORDS, dtype=np.uint64)
    f = np.empty((742, 2, 2, 1024, 2), dtype=np.float64)
    rc = lib().fb_server_key_from_bincode(_p(buf), buf.size, _p(ksk), _p(f))
    if rc != FB_OK:
        _raise(rc, "not a bincode tfhe::integer::ServerKey of PARAM_MESSAGE_2_CARRY_2")
    return ksk.reshape(2048, 5, 743), f


def string_ciphertext_to_bincode(content: np.ndarray) -> bytes:
    """bincode::serialize(&Vec<RadixCiphertext>) = StringCiphertext (ciphertext.rs:6) from [n][4][2049] words"""
    content = np.ascontiguousarray(content, dtype=np.uint64).reshape(-1, 4, BIG)
    n = C.c_size_t(0)
    lib().fb_string_ciphertext_to_bincode(_p(content), content.shape[0], None, 0, C.byref(n))
    out = np.empty(n.value, dtype=np.uint8)
    rc = lib().fb_string_ciphertext_to_bincode(_p(content), content.shape[0], _p(out), out.size, C.byref(n))
    if rc != FB_OK:
        _raise(rc, "ciphertext serialization failed")
    return out.tobytes()


def string_ciphertext_from_bincode(blob: bytes) -> np.ndarray:
    buf = np.frombuffer(blob, dtype=np.uint8)
    n = C.c_size_t(0)
    rc = lib().fb_string_ciphertext_from_bincode(_p(buf), buf.size, None, 0, C.byref(n))
    if rc != FB_OK:
        _raise(rc, "not a bincode Vec<RadixCiphertext>")
    out = np.empty((n.value, 4, BIG), dtype=np.uint64)
    rc = lib().fb_string_ciphertext_from_bincode(_p(buf), buf.size, _p(out), n.value, C.byref(n))
    if rc != FB_OK:
        _raise(rc, "not a bincode Vec<RadixCiphertext>")
    return out


def radix_to_bincode(ct: np.ndarray, degrees=None) -> bytes:
    ct = np.ascontiguousarray(ct, dtype=np.uint64).reshape(4, BIG)
    out = np.empty(lib().fb_radix_bincode_size(), dtype=np.uint8)
    n = C.c_size_t(0)
    deg = None if degrees is None else np.ascontiguousarray(degrees, dtype=np.uint64)
    rc = lib().fb_radix_to_bincode(_p(ct), None if deg is None else _p(deg), _p(out), out.size, C.byref(n))
    if rc != FB_OK:
        _raise(rc, "ciphertext serialization failed")
    return out.tobytes()


def radix_from_bincode(blob: bytes):
    buf = np.frombuffer(blob, dtype=np.uint8)
    ct = np.empty((4, BIG), dtype=np.uint64)
    deg = np.empty(4, dtype=np.uint64)
    rc = lib().fb_radix_from_bincode(_p(buf), buf.size, _p(ct), _p(deg))
    if rc != FB_OK:
        _raise(rc, "not a bincode RadixCiphertext")
    return ct, deg


def make_lut(f) -> np.ndarray:
    table = np.array([int(f(i)) & 15 for i in range(16)], dtype=np.uint64)
    out = np.empty(POLY, dtype=np.uint64)
    lib().fb_make_lut(_p(table), _p(out))
    return out


class ClientKey:
    """RadixClientKey: secret keys as serialized in test_data/client_key (engine.rs:238-254)."""

    def __init__(self, big: np.ndarray, small: np.ndarray):
        self.big, self.small = np.ascontiguousarray(big, dtype=np.uint64), np.ascontiguousarray(small, dtype=np.uint64)

    @staticmethod
    def from_bincode(buf: bytes) -> "ClientKey":
        big, small = np.empty(2048, dtype=np.uint64), np.empty(742, dtype=np.uint64)
        raw = np.frombuffer(buf, dtype=np.uint8)
        rc = lib().fb_client_key_from_bincode(_p(np.ascontiguousarray(raw)), len(buf), _p(big), _p(small))
        if rc != FB_OK:
            _raise(rc, "malformed client key")
        return ClientKey(big, small)

    @staticmethod
    def load(path: str) -> "ClientKey":
        with open(path, "rb") as f:
            return ClientKey.from_bincode(f.read())

    def encrypt_block(self, m: int, seed: int = 1, stream: int = 0) -> np.ndarray:
        out = np.empty(BIG, dtype=np.uint64)
        lib().fb_client_encrypt_block(_p(self.big), m, seed, stream, _p(out))
        return out

    def encrypt_blocks(self, msgs, seed: int = 1, stream0: int = 0) -> np.ndarray:
        out = np.empty((len(msgs), BIG), dtype=np.uint64)
        for i, m in enumerate(msgs):
            lib().fb_client_encrypt_block(_p(self.big), int(m), seed, stream0 + i, _p(out[i]))
        return out

    def decrypt_block(self, ct: np.ndarray) -> int:
        return int(lib().fb_client_decrypt_block(_p(self.big), _p(np.ascontiguousarray(ct))))

    def phase(self, ct: np.ndarray) -> int:
        return int(lib().fb_client_phase(_p(self.big), 2048, _p(np.ascontiguousarray(ct))))

    def decrypt(self, radix_ct: np.ndarray) -> int:
        """RadixClientKey::decrypt (mod.rs:17)."""
        return int(lib().fb_client_decrypt_radix(_p(self.big), _p(np.ascontiguousarray(radix_ct))))


def pinned_empty(shape, dtype=np.uint64) -> np.ndarray:
    """numpy array over a page-locked buffer of the library (fb_host_alloc): uploads from it run at PCIe speed.  Falls back to
    an ordinary array when no device is present (this is host memory only -- the compute entry points never fall back)."""
    import weakref
    dtype = np.dtype(dtype)
    n = int(np.prod(shape, dtype=np.int64)) * dtype.itemsize
    ptr = C.c_void_p()
    if n == 0 or lib().fb_host_alloc(n, C.byref(ptr)) != FB_OK or not ptr.value:
        return np.empty(shape, dtype=dtype)
    buf = (C.c_ubyte * n).from_address(ptr.value)
    weakref.finalize(buf, lib().fb_host_free, ptr.value)   # the array (and its views) keep buf alive through .base
    return np.frombuffer(buf, dtype=dtype).reshape(shape)


def encrypt_str(client_key: ClientKey, s: str, seed: int = 1) -> np.ndarray:
    """encrypt_str (ciphertext.rs:32-40): [len, 4, 2049] u64; ValueError on non-ASCII."""
    b = s.encode("latin-1", errors="replace") if s.isascii() else None
    if b is None:
        raise ValueError("content contains non-ascii characters")
    out = pinned_empty((len(b), 4, BIG), np.uint64)   # what a host hands to has_match: page-locked when a device is there
    raw = np.frombuffer(b, dtype=np.uint8) if len(b) else np.zeros(0, dtype=np.uint8)
    rc = lib().fb_client_encrypt_str(_p(client_key.big), _p(np.ascontiguousarray(raw)), len(b), seed, _p(out))
    if rc != FB_OK:
        raise ValueError("content contains non-ascii characters")
    return out


def trivial_str(s: str) -> np.ndarray:
    """create_trivial_radix per byte (ciphertext.rs:8-30), as the reference's tests do (engine.rs:282-286)."""
    b = s.encode("latin-1")
    out = np.empty((len(b), 4, BIG), dtype=np.uint64)
    raw = np.frombuffer(b, dtype=np.uint8) if len(b) else np.zeros(0, dtype=np.uint8)
    lib().fb_client_trivial_str(_p(np.ascontiguousarray(raw)), len(b), _p(out))
    return out


def keygen_server_raw(client_key: ClientKey, seed: int = 0):
    """ServerKey::new(&client_key) (engine.rs:252) -> (ksk [2048,5,743], bsk_std [742,1,2,2,2048])."""
    ksk = np.empty((2048, 5, 743), dtype=np.uint64)
    bsk = np.empty((742, 1, 2, 2, 2048), dtype=np.uint64)
    rc = lib().fb_client_keygen_server(_p(client_key.big), _p(client_key.small), seed, _p(ksk), _p(bsk))
    if rc != FB_OK:
        _raise(rc, "keygen failed")
    return ksk, bsk


# -------------------------------------------------------------------------------------------------
# GPU surface
# -------------------------------------------------------------------------------------------------
class ServerKey:
    """The server key resident on one B200 (KSK + Fourier BSK in HBM) plus the evaluation context."""

    def __init__(self, ksk: np.ndarray = None, bsk_std: np.ndarray = None, device: int = 0, *, fourier_bsk: np.ndarray = None,
                 bincode: bytes = None, keygen_from: "ClientKey" = None, seed: int = 0, keep_generated: bool = False):
        """ksk + bsk_std: the two tfhe-rs containers in the standard domain (fb_load_server_key_raw);
        ksk + fourier_bsk: the Fourier key in tfhe-rs's serialized order (fb_load_server_key_fourier);
        bincode: a serialized tfhe::integer::ServerKey (fb_load_server_key_bincode)."""
        L = lib()
        self._h = C.c_void_p()
        rc = L.fb_ctx_create(C.byref(self._h), device)
        if rc != FB_OK:
            msg = L.fb_last_error(None).decode()
            self._h = None
            raise FbError(rc, msg)
        if keygen_from is not None:   # ServerKey::new(&client_key) on the GPU (fb_keygen_server_gpu)
            self.generated = None
            if keep_generated:
                self.generated = (np.empty((2048, 5, 743), dtype=np.uint64), np.empty((742, 1, 2, 2, 2048), dtype=np.uint64))
            self._check(L.fb_keygen_server_gpu(self._h, _p(keygen_from.big), _p(keygen_from.small), seed,
                                               _p(self.generated[0]) if keep_generated else None,
                                               _p(self.generated[1]) if keep_generated else None))
            return
        if bincode is not None:
            buf = np.frombuffer(bincode, dtype=np.uint8)
            self._check(L.fb_load_server_key_bincode(self._h, _p(buf), buf.size))
            return
        ksk = np.ascontiguousarray(ksk, dtype=np.uint64)
        assert ksk.size == KSK_WORDS
        if fourier_bsk is not None:
            f = np.ascontiguousarray(fourier_bsk, dtype=np.float64)
            assert f.size == 742 * 4 * 1024 * 2
            self._check(L.fb_load_server_key_fourier(self._h, _p(ksk), _p(f)))
            return
        bsk_std = np.ascontiguousarray(bsk_std, dtype=np.uint64)
        assert bsk_std.size == BSK_WORDS
        self._check(L.fb_load_server_key_raw(self._h, _p(ksk), _p(bsk_std)))

    def _check(self, rc):
        if rc != FB_OK:
            _raise(rc, lib().fb_last_error(self._h).decode())

    def close(self):
        if getattr(self, "_h", None):
            lib().fb_ctx_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def stream(self) -> int:
        return int(lib().fb_ctx_stream(self._h) or 0)

    def sync(self):
        self._check(lib().fb_sync(self._h))

    def fourier_bsk(self) -> np.ndarray:
        out = np.empty((742, 2, 2, 1024, 2), dtype=np.float64)
        self._check(lib().fb_get_fourier_bsk(self._h, _p(out)))
        return out

    def keyswitch(self, cts: np.ndarray) -> np.ndarray:
        cts = np.ascontiguousarray(cts.reshape(-1, BIG), dtype=np.uint64)
        out = np.empty((cts.shape[0], SMALL), dtype=np.uint64)
        self._check(lib().fb_keyswitch_batch(self._h, _p(cts), cts.shape[0], _p(out)))
        return out

    def pbs(self, cts: np.ndarray, luts: np.ndarray, lut_idx) -> np.ndarray:
        cts = np.ascontiguousarray(cts.reshape(-1, BIG), dtype=np.uint64)
        luts = np.ascontiguousarray(luts.reshape(-1, POLY), dtype=np.uint64)
        idx = np.ascontiguousarray(np.asarray(lut_idx, dtype=np.uint32))
        assert idx.shape[0] == cts.shape[0]
        out = np.empty_like(cts)
        self._check(lib().fb_pbs_batch(self._h, _p(cts), _p(luts), luts.shape[0], _p(idx), cts.shape[0], _p(out)))
        return out

    def bootstrap_small(self, small: np.ndarray, luts: np.ndarray, lut_idx) -> np.ndarray:
        small = np.ascontiguousarray(small.reshape(-1, SMALL), dtype=np.uint64)
        luts = np.ascontiguousarray(luts.reshape(-1, POLY), dtype=np.uint64)
        idx = np.ascontiguousarray(np.asarray(lut_idx, dtype=np.uint32))
        out = np.empty((small.shape[0], BIG), dtype=np.uint64)
        self._check(lib().fb_bootstrap_small_batch(self._h, _p(small), _p(luts), luts.shape[0], _p(idx), small.shape[0], _p(out)))
        return out

    def pbs_dev(self, d_in: int, d_luts: int, d_lut_idx: int, count: int, d_out: int):
        """device pointers (e.g. torch tensor .data_ptr()); asynchronous on self.stream"""
        self._check(lib().fb_pbs_batch_dev(self._h, d_in, d_luts, d_lut_idx, count, d_out))

    def timing(self, on: bool):
        self._check(lib().fb_kernel_timing_enable(self._h, int(on)))

    def kernel_stats(self, reset: bool = False) -> dict:
        st = KernelStats()
        self._check(lib().fb_kernel_stats_get(self._h, C.byref(st)))
        if reset:
            self._check(lib().fb_kernel_stats_reset(self._h))
        return st.as_dict()

    def fp64_peak_tflops(self, reps: int = 5) -> float:
        v = C.c_double(0)
        self._check(lib().fb_measure_fp64_peak(self._h, reps, C.byref(v)))
        return v.value

    # ---- op-level boundary: device arenas (include/fhe_b200.h, handles.cu) ----
    def ct_alloc(self, rows: int, row_words: int = BIG) -> int:
        h = C.c_uint64(0)
        self._check(lib().fb_ct_alloc(self._h, rows, row_words, C.byref(h)))
        return int(h.value)

    def ct_free(self, h: int):
        self._check(lib().fb_ct_free(self._h, h))

    def ct_upload(self, h: int, first_row: int, rows: np.ndarray):
        rows = np.ascontiguousarray(rows, dtype=np.uint64)
        rows = rows.reshape(-1, rows.shape[-1])
        self._check(lib().fb_ct_upload(self._h, h, first_row, _p(rows), rows.shape[0]))

    def ct_download(self, h: int, first_row: int, count: int, row_words: int = BIG) -> np.ndarray:
        out = np.empty((count, row_words), dtype=np.uint64)
        self._check(lib().fb_ct_download(self._h, h, first_row, count, _p(out)))
        return out

    def lincomb(self, h: int, out_rows, term_off, term_rows, term_coef, body_const):
        a = [np.ascontiguousarray(out_rows, dtype=np.int32), np.ascontiguousarray(term_off, dtype=np.int32),
             np.ascontiguousarray(term_rows, dtype=np.int32), np.ascontiguousarray(term_coef, dtype=np.int64),
             np.ascontiguousarray(body_const, dtype=np.uint64)]
        if a[0].size == 0:
            return
        assert a[1].size == a[0].size + 1 and a[4].size == a[0].size and a[2].size == a[3].size == a[1][-1]
        ptr = [_p(x) if x.size else None for x in a]
        self._check(lib().fb_lincomb(self._h, h, ptr[0], ptr[1], ptr[2], ptr[3], ptr[4], a[0].size))

    def pbs_rows(self, h: int, in_rows, luts_h: int, lut_idx, out_row_base: int):
        rows = np.ascontiguousarray(in_rows, dtype=np.int32)
        idx = np.ascontiguousarray(lut_idx, dtype=np.uint32)
        assert rows.size == idx.size
        if rows.size:
            self._check(lib().fb_pbs_rows(self._h, h, _p(rows), luts_h, _p(idx), rows.size, out_row_base))

    def comm_init(self, comm_id: bytes, rank: int, world: int):
        """join the NCCL communicator of comm_id (fb.comm_unique_id() on rank 0, handed to the other ranks by the host)"""
        buf = np.frombuffer(comm_id, dtype=np.uint8)
        assert buf.size == 128
        self._check(lib().fb_comm_init(self._h, _p(np.ascontiguousarray(buf)), rank, world))

    def set_option(self, name: str, value: int) -> int:
        """fb_set_option: per-context knob (include/fhe_b200.h lists them); returns the previous value"""
        prev = C.c_int64(0)
        self._check(lib().fb_get_option(self._h, name.encode(), C.byref(prev)))
        self._check(lib().fb_set_option(self._h, name.encode(), int(value)))
        return int(prev.value)

    def get_option(self, name: str) -> int:
        v = C.c_int64(0)
        self._check(lib().fb_get_option(self._h, name.encode(), C.byref(v)))
        return int(v.value)

    def set_latency_threshold(self, max_count: int) -> int:
        """batches of up to max_count PBS use the one-PBS-per-CTA blind rotation; returns the previous value"""
        prev = lib().fb_set_latency_threshold(self._h, int(max_count))
        if prev < 0:
            self._check(prev)
        return prev

    def set_cluster_threshold(self, max_count: int) -> int:
        """batches of up to max_count PBS use the one-PBS-per-SM-pair blind rotation; returns the previous value"""
        prev = lib().fb_set_cluster_threshold(self._h, int(max_count))
        if prev < 0:
            self._check(prev)
        return prev

    def pbs_quantum(self) -> int:
        return int(lib().fb_pbs_batch_quantum(self._h))

    def or_fold(self, booleans: np.ndarray) -> np.ndarray:
        booleans = np.ascontiguousarray(booleans.reshape(-1, BIG), dtype=np.uint64)
        out = np.empty((4, BIG), dtype=np.uint64)
        self._check(lib().fb_or_fold(self._h, _p(booleans), booleans.shape[0], _p(out)))
        return out


def has_match(server_key: ServerKey, content: np.ndarray, pattern: str, return_stats: bool = False, rank: int = 0, world: int = 1):
    """has_match(&ServerKey, &[RadixCiphertext], &str) -> Result<RadixCiphertext> (engine.rs:8-42).

    content: [n, 4, 2049] u64 (encrypt_str layout).  Returns a radix ciphertext [4, 2049] whose
    decryption (ClientKey.decrypt) is 0/1.  Raises ParseError like the reference's Err, ReferencePanic
    where the reference would panic."""
    content = np.ascontiguousarray(content, dtype=np.uint64)
    n = content.shape[0] if content.ndim == 3 else 0
    out = np.empty((4, BIG), dtype=np.uint64)
    st = MatchStats()
    rc = lib().fb_has_match_shard(server_key._h, _p(content) if n else None, n, pattern.encode("latin-1"), rank, world, _p(out), C.byref(st))
    server_key._check(rc)
    return (out, st.as_dict()) if return_stats else out


def regex_lut_table() -> np.ndarray:
    out = np.empty((51, POLY), dtype=np.uint64)
    rc = lib().fb_regex_lut_table(_p(out))
    if rc != FB_OK:
        _raise(rc, "fb_regex_lut_table")
    return out


def plan_export(pattern: str, n_chars: int, reference_shaped: bool = False) -> dict:
    """the library's level-synchronous plan of a match (fb_plan_export), parsed into python lists"""
    flags = PLAN_REFERENCE_SHAPED if reference_shaped else 0
    n = C.c_size_t(0)
    rc = lib().fb_plan_export(pattern.encode("latin-1"), n_chars, flags, None, 0, C.byref(n))
    if rc != FB_OK:
        _raise(rc, "plan failed for %r" % pattern)
    w = np.empty(n.value, dtype=np.int64)
    rc = lib().fb_plan_export(pattern.encode("latin-1"), n_chars, flags, _p(w), w.size, C.byref(n))
    if rc != FB_OK:
        _raise(rc, "plan export failed")
    pos = [0]

    def take(k):
        v = w[pos[0]:pos[0] + k]
        pos[0] += k
        return v

    n_rows, kind, res_row, n_levels = (int(x) for x in take(4))
    levels = []
    for _ in range(n_levels):
        n_lin, n_terms, n_pbs, base = (int(x) for x in take(4))
        lv = {"out_row_base": base, "lin_out_rows": take(n_lin).astype(np.int32)}
        lv["lin_term_off"] = take(n_lin + 1 if n_lin else 1).astype(np.int32)
        lv["lin_term_rows"] = take(n_terms).astype(np.int32)
        lv["lin_coef"] = take(n_terms).copy()
        lv["lin_const"] = take(n_lin).view(np.uint64).copy()
        lv["in_rows"] = take(n_pbs).astype(np.int32)
        lv["lut_idx"] = take(n_pbs).astype(np.uint32)
        levels.append(lv)
    assert pos[0] == w.size
    return {"n_rows": n_rows, "result_kind": kind, "result_row": res_row, "levels": levels}


def comm_unique_id() -> bytes:
    buf = np.zeros(128, dtype=np.uint8)
    rc = lib().fb_comm_unique_id(_p(buf))
    if rc != FB_OK:
        _raise(rc, "NCCL unavailable")
    return buf.tobytes()


def has_match_dist(server_key: ServerKey, content: np.ndarray, pattern: str, return_stats: bool = False):
    """collective has_match over the communicator of server_key.comm_init (fb_has_match_dist)"""
    content = np.ascontiguousarray(content, dtype=np.uint64).reshape(-1, 4, BIG)
    out = np.empty((4, BIG), dtype=np.uint64)
    st = MatchStats()
    server_key._check(lib().fb_has_match_dist(server_key._h, _p(content), content.shape[0], pattern.encode("latin-1"), _p(out), C.byref(st)))
    return (out, st.as_dict()) if return_stats else out


def has_match_many(server_key: ServerKey, contents: np.ndarray, pattern: str, return_stats: bool = False):
    """The same match for many contents of one length against one pattern, level by level in shared launches.

    contents: [m, n, 4, 2049] u64 (m stacked encrypt_str results).  Returns [m, 4, 2049]: one radix ciphertext per
    content, each decrypting to what has_match gives for that content alone."""
    contents = np.ascontiguousarray(contents, dtype=np.uint64)
    if contents.ndim != 4:
        raise ValueError("contents must be [m, n_chars, 4, 2049]")
    m, n = contents.shape[0], contents.shape[1]
    out = np.empty((m, 4, BIG), dtype=np.uint64)
    st = MatchStats()
    rc = lib().fb_has_match_many(server_key._h, _p(contents) if m * n else None, m, n, pattern.encode("latin-1"), _p(out) if m else None, C.byref(st))
    server_key._check(rc)
    return (out, st.as_dict()) if return_stats else out
