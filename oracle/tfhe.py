"""ctypes/numpy harness around oracle/libtfhe_oracle.so (CPU restatement of tfhe-rs 0.2.0 arithmetic).

TEST INFRASTRUCTURE ONLY -- imported by tests/, __graft_entry__.smoke() and bench.py's CPU-baseline
legs.  The product package (fhe_regex_b200/) never imports this module.

Parity status: ciphertext-level parity is UNPINNED (the reference holds no known-answer vectors for
keyswitch/bootstrap); see the header of tfhe_oracle.c.
"""
from __future__ import annotations

import ctypes as C
import os
import struct
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libtfhe_oracle.so")

LWE_N, GLWE_N, KS_LEVELS = 742, 2048, 5
BIG, SMALL = GLWE_N + 1, LWE_N + 1
DELTA_LOG = 59
SIGMA_LWE = 7.069849454709433e-06
SIGMA_GLWE = 2.9403601535432533e-16


def build(force: bool = False) -> str:
    src = os.path.join(HERE, "tfhe_oracle.c")
    if force or not os.path.exists(LIB_PATH) or os.path.getmtime(LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", HERE, "-s"])
    return LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(LIB_PATH)
        u64p = C.POINTER(C.c_uint64)
        dblp = C.POINTER(C.c_double)
        L.orc_decompose.argtypes = [C.c_uint64, C.c_int, C.c_int, C.POINTER(C.c_int64)]
        L.orc_closest_representable.argtypes = [C.c_uint64, C.c_int, C.c_int]
        L.orc_closest_representable.restype = C.c_uint64
        L.orc_lwe_encrypt.argtypes = [u64p, C.c_int, C.c_uint64, C.c_double, C.c_uint64, C.c_uint64, u64p]
        L.orc_lwe_phase.argtypes = [u64p, C.c_int, u64p]
        L.orc_lwe_phase.restype = C.c_uint64
        L.orc_decode.argtypes = [C.c_uint64]
        L.orc_decode.restype = C.c_uint64
        L.orc_keygen_ksk.argtypes = [u64p, u64p, C.c_double, C.c_uint64, u64p]
        L.orc_keygen_bsk.argtypes = [u64p, u64p, C.c_double, C.c_uint64, u64p]
        L.orc_keyswitch.argtypes = [u64p, u64p, u64p]
        L.orc_keyswitch_batch.argtypes = [u64p, u64p, u64p, C.c_int, C.c_int]
        L.orc_modswitch.argtypes = [C.c_uint64]
        L.orc_modswitch.restype = C.c_uint64
        L.orc_make_lut.argtypes = [u64p, u64p]
        L.orc_bsk_to_fourier.argtypes = [u64p, dblp]
        L.orc_blind_rotate_fft.argtypes = [dblp, u64p, u64p, u64p]
        L.orc_blind_rotate_exact.argtypes = [u64p, u64p, u64p, u64p]
        L.orc_sample_extract.argtypes = [u64p, u64p]
        L.orc_pbs.argtypes = [u64p, dblp, u64p, u64p, u64p]
        L.orc_pbs_batch.argtypes = [u64p, dblp, u64p, u64p, C.POINTER(C.c_uint32), u64p, C.c_int, C.c_int]
        L.orc_bootstrap_small.argtypes = [dblp, u64p, u64p, u64p]
        L.orc_max_threads.restype = C.c_int
        _lib = L
    return _lib


def _u64(a):
    assert a.dtype == np.uint64 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(C.POINTER(C.c_uint64))


def _dbl(a):
    assert a.dtype == np.float64 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(C.POINTER(C.c_double))


# ---------------------------------------------------------------------------------------------
# fixture: bincode-serialised RadixClientKey, layout measured in SURVEY.md 8c
# ---------------------------------------------------------------------------------------------
class ClientKey:
    """Secret keys + parameters as stored in the reference fixture test_data/client_key
    (written by engine.rs:232-246, read by engine.rs:248-254)."""

    def __init__(self, big, glwe, small, params, num_blocks):
        self.big, self.glwe, self.small, self.params, self.num_blocks = big, glwe, small, params, num_blocks

    @staticmethod
    def from_bincode(buf: bytes) -> "ClientKey":
        o = 0

        def vec():
            nonlocal o
            (n,) = struct.unpack_from("<Q", buf, o)
            o += 8
            a = np.frombuffer(buf, dtype="<u8", count=n, offset=o).astype(np.uint64)
            o += 8 * n
            return a

        big = vec()
        glwe = vec()
        (poly,) = struct.unpack_from("<Q", buf, o)
        o += 8
        small = vec()
        names = ["lwe_dimension", "glwe_dimension", "polynomial_size", "lwe_modular_std_dev",
                 "glwe_modular_std_dev", "pbs_base_log", "pbs_level", "ks_base_log", "ks_level",
                 "pfks_level", "pfks_base_log", "pfks_modular_std_dev", "cbs_level", "cbs_base_log",
                 "message_modulus", "carry_modulus"]
        params = {}
        for nm in names:
            if nm.endswith("std_dev"):
                (v,) = struct.unpack_from("<d", buf, o)
            else:
                (v,) = struct.unpack_from("<Q", buf, o)
            params[nm] = v
            o += 8
        params["glwe_key_polynomial_size"] = poly
        (nb,) = struct.unpack_from("<Q", buf, o)
        o += 8
        assert o == len(buf), (o, len(buf))
        return ClientKey(big, glwe, small, params, nb)

    @staticmethod
    def load(path: str) -> "ClientKey":
        with open(path, "rb") as f:
            return ClientKey.from_bincode(f.read())


class ServerKey:
    def __init__(self, ksk, bsk):
        self.ksk, self.bsk = ksk, bsk  # [2048,5,743] u64 ; [742,1,2,2,2048] u64
        self._fbsk = None

    @property
    def fbsk(self):
        if self._fbsk is None:
            self._fbsk = np.empty(LWE_N * 4 * GLWE_N, dtype=np.float64)
            lib().orc_bsk_to_fourier(_u64(self.bsk.reshape(-1)), _dbl(self._fbsk))
        return self._fbsk


def keygen_server(ck: ClientKey, seed: int = 0) -> ServerKey:
    """ServerKey::new(&client_key) (engine.rs:252): fresh random BSK/KSK from the secret keys."""
    ksk = np.empty((GLWE_N, KS_LEVELS, SMALL), dtype=np.uint64)
    bsk = np.empty((LWE_N, 1, 2, 2, GLWE_N), dtype=np.uint64)
    lib().orc_keygen_ksk(_u64(ck.big), _u64(ck.small), SIGMA_LWE, seed, _u64(ksk.reshape(-1)))
    lib().orc_keygen_bsk(_u64(ck.small), _u64(ck.glwe), SIGMA_GLWE, seed, _u64(bsk.reshape(-1)))
    return ServerKey(ksk, bsk)


# ---------------------------------------------------------------------------------------------
# ciphertext helpers
# ---------------------------------------------------------------------------------------------
def encrypt_shortint(ck: ClientKey, m: int, seed: int, stream: int) -> np.ndarray:
    """shortint ClientKey::encrypt: big-key LWE, glwe noise, pt = m * 2^59."""
    out = np.empty(BIG, dtype=np.uint64)
    lib().orc_lwe_encrypt(_u64(ck.big), GLWE_N, (m & 15) << DELTA_LOG, SIGMA_GLWE, seed, stream, _u64(out))
    return out


def encrypt_batch(ck: ClientKey, msgs, seed: int = 1, stream0: int = 0) -> np.ndarray:
    out = np.empty((len(msgs), BIG), dtype=np.uint64)
    for i, m in enumerate(msgs):
        out[i] = encrypt_shortint(ck, int(m), seed, stream0 + i)
    return out


def trivial_shortint(m: int) -> np.ndarray:
    out = np.zeros(BIG, dtype=np.uint64)
    out[GLWE_N] = np.uint64((m & 15) << DELTA_LOG)
    return out


def phase_big(ck: ClientKey, ct: np.ndarray) -> int:
    return int(lib().orc_lwe_phase(_u64(ck.big), GLWE_N, _u64(np.ascontiguousarray(ct))))


def phase_small(ck: ClientKey, ct: np.ndarray) -> int:
    return int(lib().orc_lwe_phase(_u64(ck.small), LWE_N, _u64(np.ascontiguousarray(ct))))


def decode(phase: int) -> int:
    return ((phase + (1 << 58)) >> 59) & 15


def decrypt_shortint(ck: ClientKey, ct: np.ndarray) -> int:
    return decode(phase_big(ck, ct))


def phase_batch(key: np.ndarray, cts: np.ndarray) -> np.ndarray:
    """vectorised phases for [B, dim+1] ciphertexts (wrapping u64)."""
    dim = key.shape[0]
    with np.errstate(over="ignore"):
        acc = (cts[:, :dim] * key[None, :]).sum(axis=1, dtype=np.uint64)
        return cts[:, dim] - acc


def torus_err(phase: np.ndarray, expected_pt: np.ndarray) -> np.ndarray:
    """signed torus distance (fraction of the torus) between phases and expected plaintexts."""
    with np.errstate(over="ignore"):
        d = (phase - expected_pt).astype(np.uint64).view(np.int64)
    return d.astype(np.float64) / 2.0 ** 64


def encrypt_radix(ck: ClientKey, value: int, seed: int, stream: int) -> np.ndarray:
    """RadixClientKey::encrypt (ciphertext.rs:38): 4 blocks of 2 bits, little endian -> [4, 2049]."""
    return np.stack([encrypt_shortint(ck, (value >> (2 * b)) & 3, seed, stream * 4 + b) for b in range(4)])


def trivial_radix(value: int) -> np.ndarray:
    """create_trivial_radix (ciphertext.rs:8-30)."""
    return np.stack([trivial_shortint((value >> (2 * b)) & 3) for b in range(4)])


def decrypt_radix(ck: ClientKey, ct: np.ndarray) -> int:
    """RadixClientKey::decrypt (mod.rs:17): sum of block message+carry * 4^i, mod 256."""
    v = 0
    for b in range(ct.shape[0]):
        v += decrypt_shortint(ck, ct[b]) * (4 ** b)
    return v % 256


def encrypt_str(ck: ClientKey, s: str, seed: int = 1) -> np.ndarray:
    """encrypt_str (ciphertext.rs:32-40): ASCII only, one radix ciphertext per byte -> [n, 4, 2049]."""
    if not s.isascii():
        raise ValueError("content contains non-ascii characters")
    if len(s) == 0:
        return np.empty((0, 4, BIG), dtype=np.uint64)
    return np.stack([encrypt_radix(ck, b, seed, i) for i, b in enumerate(s.encode())])


# ---------------------------------------------------------------------------------------------
# arithmetic
# ---------------------------------------------------------------------------------------------
def make_lut(f) -> np.ndarray:
    table = np.array([int(f(i)) & 15 for i in range(16)], dtype=np.uint64)
    out = np.empty(GLWE_N, dtype=np.uint64)
    lib().orc_make_lut(_u64(table), _u64(out))
    return out


def keyswitch(sk: ServerKey, cts: np.ndarray, nthreads: int = 0) -> np.ndarray:
    cts = np.ascontiguousarray(cts.reshape(-1, BIG))
    out = np.empty((cts.shape[0], SMALL), dtype=np.uint64)
    lib().orc_keyswitch_batch(_u64(sk.ksk.reshape(-1)), _u64(cts.reshape(-1)), _u64(out.reshape(-1)), cts.shape[0], nthreads)
    return out


def pbs(sk: ServerKey, cts: np.ndarray, luts: np.ndarray, lut_idx, nthreads: int = 0) -> np.ndarray:
    cts = np.ascontiguousarray(cts.reshape(-1, BIG))
    luts = np.ascontiguousarray(luts.reshape(-1, GLWE_N))
    idx = np.ascontiguousarray(np.asarray(lut_idx, dtype=np.uint32))
    assert idx.shape[0] == cts.shape[0]
    out = np.empty_like(cts)
    lib().orc_pbs_batch(_u64(sk.ksk.reshape(-1)), _dbl(sk.fbsk), _u64(cts.reshape(-1)), _u64(luts.reshape(-1)),
                        idx.ctypes.data_as(C.POINTER(C.c_uint32)), _u64(out.reshape(-1)), cts.shape[0], nthreads)
    return out


def bootstrap_small(sk: ServerKey, small: np.ndarray, lut: np.ndarray) -> np.ndarray:
    out = np.empty(BIG, dtype=np.uint64)
    lib().orc_bootstrap_small(_dbl(sk.fbsk), _u64(np.ascontiguousarray(small)), _u64(np.ascontiguousarray(lut)), _u64(out))
    return out


def blind_rotate(sk: ServerKey, small: np.ndarray, lut: np.ndarray, exact: bool = False) -> np.ndarray:
    acc = np.empty((2, GLWE_N), dtype=np.uint64)
    if exact:
        lib().orc_blind_rotate_exact(_u64(sk.bsk.reshape(-1)), _u64(np.ascontiguousarray(small)), _u64(np.ascontiguousarray(lut)), _u64(acc.reshape(-1)))
    else:
        lib().orc_blind_rotate_fft(_dbl(sk.fbsk), _u64(np.ascontiguousarray(small)), _u64(np.ascontiguousarray(lut)), _u64(acc.reshape(-1)))
    return acc


def sample_extract(acc: np.ndarray) -> np.ndarray:
    out = np.empty(BIG, dtype=np.uint64)
    lib().orc_sample_extract(_u64(np.ascontiguousarray(acc.reshape(-1))), _u64(out))
    return out


def decompose(x: int, base_log: int, levels: int):
    d = (C.c_int64 * levels)()
    lib().orc_decompose(x, base_log, levels, d)
    return list(d)


def modswitch(x: int) -> int:
    return int(lib().orc_modswitch(x))


def max_threads() -> int:
    return int(lib().orc_max_threads())
