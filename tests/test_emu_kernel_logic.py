"""CPU lane-by-lane emulation of the blind-rotation kernel's data flow (tests/emu/emu_br.cpp, built from
the same __host__ __device__ phase functions the CUDA kernel uses) against the oracle.  Lets the index /
twiddle / swizzle / accumulator-layout logic be checked in the build container, which has no GPU."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from oracle import tfhe

EMU_DIR = os.path.join(os.path.dirname(__file__), "emu")
CSRC = os.path.join(os.path.dirname(EMU_DIR), "..", "fhe_regex_b200", "csrc")


@pytest.fixture(scope="module")
def emu():
    so = os.path.join(EMU_DIR, "libemu_br.so")
    srcs = [os.path.join(EMU_DIR, "emu_br.cpp"), os.path.join(CSRC, "br_core.cuh"), os.path.join(CSRC, "fft32_gen.h")]
    if not os.path.exists(so) or any(os.path.getmtime(so) < os.path.getmtime(s) for s in srcs):
        subprocess.check_call(["/usr/bin/g++", "-O2", "-march=x86-64-v3", "-fPIC", "-shared", "-o", so, srcs[0]])
    return ctypes.CDLL(so)


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


@pytest.fixture(scope="module")
def emu_fbsk(emu, server_key):
    out = np.zeros((742, 2, 2, 1024, 2), dtype=np.float64)
    emu.emu_bsk_to_fourier(_p(server_key.bsk), _p(out))
    return out


def test_emulated_negacyclic_product_matches_exact(emu):
    rng = np.random.default_rng(3)
    a = rng.integers(-(1 << 22), 1 << 22, size=2048, dtype=np.int64)
    b = rng.integers(0, 1 << 64, size=2048, dtype=np.uint64)
    got = np.zeros(2048, dtype=np.uint64)
    emu.emu_negacyclic_mul(_p(a), _p(b), _p(got))
    # exact negacyclic product mod 2^64 with python ints on a few coefficients
    ai, bi = [int(x) for x in a], [int(x) for x in b]
    for j in (0, 1, 7, 1023, 1024, 2047):
        s = 0
        for t in range(2048):
            u = j - t
            s += ai[t] * bi[u] if u >= 0 else -ai[t] * bi[u + 2048]
        err = tfhe.torus_err(np.array([got[j]], dtype=np.uint64), np.array([s % (1 << 64)], dtype=np.uint64))[0]
        assert abs(err) < 2 ** -20, (j, err)


def test_emulated_blind_rotate_decrypts_like_the_oracle(emu, emu_fbsk, client_key, server_key):
    msgs = np.array([3, 12], dtype=np.int64)
    cts = tfhe.encrypt_batch(client_key, msgs, seed=77)
    small = tfhe.keyswitch(server_key, cts)
    lut = tfhe.make_lut(lambda x: (x * 3 + 1) % 16)
    for b in range(len(msgs)):
        acc = np.zeros(2 * 2048, dtype=np.uint64)
        emu.emu_blind_rotate(_p(emu_fbsk), _p(small[b]), _p(lut), _p(acc), -1)
        out = tfhe.sample_extract(acc)
        exp = (int(msgs[b]) * 3 + 1) % 16
        assert tfhe.decrypt_shortint(client_key, out) == exp
        ref = tfhe.bootstrap_small(server_key, small[b], lut)
        assert tfhe.decrypt_shortint(client_key, ref) == exp
        ph = tfhe.phase_batch(client_key.big, np.stack([out, ref]))
        err = tfhe.torus_err(ph, np.array([exp << 59] * 2, dtype=np.uint64))
        assert np.abs(err).max() < 4e-4, err


def test_emulated_fused_body_matches_phase_by_phase_body(emu, emu_fbsk, client_key, server_key):
    """br_fused.cu re-orders the per-thread program (digits interleaved with pass 1, MAC block by block between the
    transforms, untwist folded into the last butterflies, three-FMA torus rounding): same function.  After ONE CMUX
    step the accumulators agree to f64 rounding (values of ~2^27 carry ~2^-25 of absolute error = ~2^7 units of 2^-32);
    later steps are not comparable bit by bit (a digit that rounds the other way adds a whole key coefficient -- noise
    to the decryption, not to the ciphertext bits), so the full rotation is compared by decryption and error bound."""
    cts = tfhe.encrypt_batch(client_key, np.array([7], dtype=np.int64), seed=78)
    small = tfhe.keyswitch(server_key, cts)[0]
    lut = tfhe.make_lut(lambda x: (x * 7 + 2) % 16)
    accs = []
    for fused in (0, 1):
        emu.emu_set_fused(fused)
        acc = np.zeros(2 * 2048, dtype=np.uint64)
        emu.emu_blind_rotate(_p(emu_fbsk), _p(small), _p(lut), _p(acc), 1)
        accs.append(acc)
    emu.emu_set_fused(0)
    d = (accs[0] >> np.uint64(32)).astype(np.int64) - (accs[1] >> np.uint64(32)).astype(np.int64)
    d = (d + (1 << 31)) % (1 << 32) - (1 << 31)
    assert 0 < np.abs(accs[0]).max() and np.abs(d).max() <= 512, np.abs(d).max()
    # and the full rotation decrypts correctly through the fused body
    emu.emu_set_fused(1)
    acc = np.zeros(2 * 2048, dtype=np.uint64)
    emu.emu_blind_rotate(_p(emu_fbsk), _p(small), _p(lut), _p(acc), -1)
    emu.emu_set_fused(0)
    out = tfhe.sample_extract(acc)
    assert tfhe.decrypt_shortint(client_key, out) == (7 * 7 + 2) % 16
    err = tfhe.torus_err(tfhe.phase_batch(client_key.big, out[None]), np.array([((7 * 7 + 2) % 16) << 59], dtype=np.uint64))
    assert np.abs(err).max() < 4e-4, err


def test_emulated_trivial_input_is_exact(emu, emu_fbsk, server_key):
    lut = tfhe.make_lut(lambda x: (5 * x) % 16)
    small = np.zeros(743, dtype=np.uint64)
    small[742] = np.uint64(9 << 59)
    acc = np.zeros(2 * 2048, dtype=np.uint64)
    emu.emu_blind_rotate(_p(emu_fbsk), _p(small), _p(lut), _p(acc), -1)
    ref = tfhe.bootstrap_small(server_key, small, lut)
    assert (tfhe.sample_extract(acc) == ref).all()


# ---- latency variant (one sample per CTA, fhe_regex_b200/csrc/br_wide.cuh) ----------------------------------

@pytest.fixture(scope="module")
def emu_wide():
    so = os.path.join(EMU_DIR, "libemu_wide.so")
    srcs = [os.path.join(EMU_DIR, "emu_wide.cpp"), os.path.join(CSRC, "br_wide.cuh"), os.path.join(CSRC, "br_core.cuh")]
    if not os.path.exists(so) or any(os.path.getmtime(so) < os.path.getmtime(s) for s in srcs):
        subprocess.check_call(["/usr/bin/g++", "-O2", "-march=x86-64-v3", "-fPIC", "-shared", "-o", so, srcs[0]])
    return ctypes.CDLL(so)


def test_wide_spectrum_equals_key_conversion_order(emu_wide, emu_fbsk, server_key):
    """Both kernels share one Fourier key: the Stockham stages must produce the 32x32 kernel's natural order."""
    bsk = server_key.bsk.reshape(742, 2, 2, 2048)
    for (i, r) in ((0, 0), (5, 1), (741, 0)):
        spec = np.zeros((2, 1024, 2), dtype=np.float64)
        emu_wide.emu_wide_forward_torus(_p(np.ascontiguousarray(bsk[i, r])), _p(spec))
        ref = emu_fbsk[i, r]
        scale = np.abs(ref).max()
        assert np.abs(spec - ref).max() < 1e-12 * scale, (i, r, np.abs(spec - ref).max(), scale)


def test_wide_negacyclic_product_matches_exact(emu_wide):
    rng = np.random.default_rng(4)
    a = rng.integers(-(1 << 22), 1 << 22, size=2048, dtype=np.int64)
    b = rng.integers(0, 1 << 64, size=2048, dtype=np.uint64)
    got = np.zeros(2048, dtype=np.uint64)
    emu_wide.emu_wide_negacyclic_mul(_p(a), _p(b), _p(got))
    ai, bi = [int(x) for x in a], [int(x) for x in b]
    for j in (0, 1, 7, 127, 128, 1023, 1024, 1500, 2047):
        s = 0
        for t in range(2048):
            u = j - t
            s += ai[t] * bi[u] if u >= 0 else -ai[t] * bi[u + 2048]
        err = tfhe.torus_err(np.array([got[j]], dtype=np.uint64), np.array([s % (1 << 64)], dtype=np.uint64))[0]
        assert abs(err) < 2 ** -20, (j, err)


def test_wide_blind_rotate_decrypts_like_the_oracle(emu_wide, emu_fbsk, client_key, server_key):
    msgs = np.array([5, 14], dtype=np.int64)
    cts = tfhe.encrypt_batch(client_key, msgs, seed=78)
    small = tfhe.keyswitch(server_key, cts)
    lut = tfhe.make_lut(lambda x: (x * 7 + 2) % 16)
    for b in range(len(msgs)):
        acc = np.zeros(2 * 2048, dtype=np.uint64)
        emu_wide.emu_wide_blind_rotate(_p(emu_fbsk), _p(small[b]), _p(lut), _p(acc), -1)
        out = tfhe.sample_extract(acc)
        exp = (int(msgs[b]) * 7 + 2) % 16
        assert tfhe.decrypt_shortint(client_key, out) == exp
        ph = tfhe.phase_batch(client_key.big, out[None])
        err = tfhe.torus_err(ph, np.array([exp << 59], dtype=np.uint64))
        assert np.abs(err).max() < 4e-4, err


def test_wide_trivial_input_is_exact(emu_wide, emu_fbsk, server_key):
    lut = tfhe.make_lut(lambda x: (3 * x + 1) % 16)
    small = np.zeros(743, dtype=np.uint64)
    small[742] = np.uint64(11 << 59)
    acc = np.zeros(2 * 2048, dtype=np.uint64)
    emu_wide.emu_wide_blind_rotate(_p(emu_fbsk), _p(small), _p(lut), _p(acc), -1)
    ref = tfhe.bootstrap_small(server_key, small, lut)
    assert (tfhe.sample_extract(acc) == ref).all()


# ---- cluster variant (one sample per pair of CTAs, fhe_regex_b200/csrc/br_duo.cuh) ---------------------------

@pytest.fixture(scope="module")
def emu_duo():
    so = os.path.join(EMU_DIR, "libemu_duo.so")
    srcs = [os.path.join(EMU_DIR, "emu_duo.cpp"), os.path.join(CSRC, "br_duo.cuh"), os.path.join(CSRC, "br_core.cuh")]
    if not os.path.exists(so) or any(os.path.getmtime(so) < os.path.getmtime(s) for s in srcs):
        subprocess.check_call(["/usr/bin/g++", "-O2", "-march=x86-64-v3", "-fPIC", "-shared", "-o", so, srcs[0]])
    return ctypes.CDLL(so)


def test_duo_spectrum_equals_key_conversion_order(emu_duo, emu_fbsk, server_key):
    bsk = server_key.bsk.reshape(742, 2, 2, 2048)
    for (i, r, c) in ((0, 0, 0), (5, 1, 1), (741, 0, 1)):
        spec = np.zeros((1024, 2), dtype=np.float64)
        emu_duo.emu_duo_forward_torus(_p(np.ascontiguousarray(bsk[i, r, c])), _p(spec))
        ref = emu_fbsk[i, r, c]
        scale = np.abs(ref).max()
        assert np.abs(spec - ref).max() < 1e-12 * scale, (i, r, c, np.abs(spec - ref).max(), scale)


def test_duo_negacyclic_product_matches_exact(emu_duo):
    rng = np.random.default_rng(5)
    a = rng.integers(-(1 << 22), 1 << 22, size=2048, dtype=np.int64)
    b = rng.integers(0, 1 << 64, size=2048, dtype=np.uint64)
    got = np.zeros(2048, dtype=np.uint64)
    emu_duo.emu_duo_negacyclic_mul(_p(a), _p(b), _p(got))
    ai, bi = [int(x) for x in a], [int(x) for x in b]
    for j in (0, 1, 7, 255, 256, 1023, 1024, 1500, 2047):
        s = 0
        for t in range(2048):
            u = j - t
            s += ai[t] * bi[u] if u >= 0 else -ai[t] * bi[u + 2048]
        err = tfhe.torus_err(np.array([got[j]], dtype=np.uint64), np.array([s % (1 << 64)], dtype=np.uint64))[0]
        assert abs(err) < 2 ** -20, (j, err)


def test_duo_blind_rotate_decrypts_like_the_oracle(emu_duo, emu_fbsk, client_key, server_key):
    msgs = np.array([6, 9], dtype=np.int64)
    cts = tfhe.encrypt_batch(client_key, msgs, seed=79)
    small = tfhe.keyswitch(server_key, cts)
    lut = tfhe.make_lut(lambda x: (x * 5 + 3) % 16)
    for b in range(len(msgs)):
        acc = np.zeros(2 * 2048, dtype=np.uint64)
        emu_duo.emu_duo_blind_rotate(_p(emu_fbsk), _p(small[b]), _p(lut), _p(acc), -1)
        out = tfhe.sample_extract(acc)
        exp = (int(msgs[b]) * 5 + 3) % 16
        assert tfhe.decrypt_shortint(client_key, out) == exp
        ph = tfhe.phase_batch(client_key.big, out[None])
        err = tfhe.torus_err(ph, np.array([exp << 59], dtype=np.uint64))
        assert np.abs(err).max() < 4e-4, err


def test_duo_trivial_input_is_exact(emu_duo, emu_fbsk, server_key):
    lut = tfhe.make_lut(lambda x: (7 * x + 1) % 16)
    small = np.zeros(743, dtype=np.uint64)
    small[742] = np.uint64(13 << 59)
    acc = np.zeros(2 * 2048, dtype=np.uint64)
    emu_duo.emu_duo_blind_rotate(_p(emu_fbsk), _p(small), _p(lut), _p(acc), -1)
    ref = tfhe.bootstrap_small(server_key, small, lut)
    assert (tfhe.sample_extract(acc) == ref).all()


@pytest.mark.parametrize("layout", [1, 2, 3])
def test_emulated_plane_layouts_are_bit_identical(emu, layout):
    """round 2: the transposes of the fused throughput kernel through (1) planes inside the accumulator copies + overflow
    blocks, (2) a plane per component with the twiddles fused into the stores, (3) layout 1 with the full twiddle tables --
    every register bit-identical to the contiguous split planes, and nothing written outside the blocks"""
    for seed in (1, 2, 3, 4):
        assert emu.emu_plane_layouts_check(layout, seed) == 0
