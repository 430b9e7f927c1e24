"""CPU emulation tests of the radix-16 stages (run from the repository root: pytest tools/experiments/w16/test_emu_w16.py).
The fixtures (emu_fbsk, server_key, client_key, _p, tfhe) are those of tests/test_emu_kernel_logic.py."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "..", "..", "tests"))
from test_emu_kernel_logic import *  # noqa: F401,F403
EMU16_DIR = os.path.dirname(os.path.abspath(__file__))

# ---- radix-16 latency variant (one sample per CTA of 128 threads, fhe_regex_b200/csrc/br_w16.cuh) -----------------

@pytest.fixture(scope="module")
def emu_w16():
    so = os.path.join(EMU16_DIR, "libemu_w16.so")
    srcs = [os.path.join(EMU16_DIR, "emu_w16.cpp"), os.path.join(EMU16_DIR, "br_w16.cuh"), os.path.join(CSRC, "br_core.cuh")]
    if not os.path.exists(so) or any(os.path.getmtime(so) < os.path.getmtime(s) for s in srcs):
        subprocess.check_call(["/usr/bin/g++", "-O2", "-march=x86-64-v3", "-fPIC", "-shared", "-o", so, srcs[0]])
    return ctypes.CDLL(so)


def test_w16_spectrum_equals_key_conversion_order(emu_w16, emu_fbsk, server_key):
    """All blind rotations share one Fourier key: the 16 x 16 x 4 stages must produce the 32x32 kernel's natural order."""
    bsk = server_key.bsk.reshape(742, 2, 2, 2048)
    for (i, r) in ((0, 0), (5, 1), (741, 0)):
        spec = np.zeros((2, 1024, 2), dtype=np.float64)
        emu_w16.emu_w16_forward_torus(_p(np.ascontiguousarray(bsk[i, r])), _p(spec))
        ref = emu_fbsk[i, r]
        scale = np.abs(ref).max()
        assert np.abs(spec - ref).max() < 1e-12 * scale, (i, r, np.abs(spec - ref).max(), scale)


def test_w16_negacyclic_product_matches_exact(emu_w16):
    rng = np.random.default_rng(14)
    a = rng.integers(-(1 << 22), 1 << 22, size=2048, dtype=np.int64)
    b = rng.integers(0, 1 << 64, size=2048, dtype=np.uint64)
    got = np.zeros(2048, dtype=np.uint64)
    emu_w16.emu_w16_negacyclic_mul(_p(a), _p(b), _p(got))
    ai, bi = [int(x) for x in a], [int(x) for x in b]
    for j in (0, 1, 7, 63, 64, 127, 128, 1023, 1024, 1500, 2047):
        s = 0
        for t in range(2048):
            u = j - t
            s += ai[t] * bi[u] if u >= 0 else -ai[t] * bi[u + 2048]
        err = tfhe.torus_err(np.array([got[j]], dtype=np.uint64), np.array([s % (1 << 64)], dtype=np.uint64))[0]
        assert abs(err) < 2 ** -20, (j, err)


def test_w16_blind_rotate_decrypts_like_the_oracle(emu_w16, emu_fbsk, client_key, server_key):
    msgs = np.array([5, 14], dtype=np.int64)
    cts = tfhe.encrypt_batch(client_key, msgs, seed=78)
    small = tfhe.keyswitch(server_key, cts)
    lut = tfhe.make_lut(lambda x: (x * 7 + 2) % 16)
    for b in range(len(msgs)):
        acc = np.zeros(2 * 2048, dtype=np.uint64)
        emu_w16.emu_w16_blind_rotate(_p(emu_fbsk), _p(small[b]), _p(lut), _p(acc), -1)
        out = tfhe.sample_extract(acc)
        exp = (int(msgs[b]) * 7 + 2) % 16
        assert tfhe.decrypt_shortint(client_key, out) == exp
        ph = tfhe.phase_batch(client_key.big, out[None])
        err = tfhe.torus_err(ph, np.array([exp << 59], dtype=np.uint64))
        assert np.abs(err).max() < 4e-4, err


def test_w16_step_agrees_with_the_radix_8_kernel(emu_w16, emu_wide, emu_fbsk, client_key, server_key):
    """the first CMUX step through both latency formulations: the accumulators differ by f64 rounding only (after that a
    last-bit difference may flip a digit, and the mask words are no longer comparable -- only the phases are)"""
    cts = tfhe.encrypt_batch(client_key, np.array([9]), seed=79)
    small = tfhe.keyswitch(server_key, cts)[0]
    lut = tfhe.make_lut(lambda x: x)
    a16 = np.zeros(2 * 2048, dtype=np.uint64)
    a8 = np.zeros(2 * 2048, dtype=np.uint64)
    emu_w16.emu_w16_blind_rotate(_p(emu_fbsk), _p(small), _p(lut), _p(a16), 1)
    emu_wide.emu_wide_blind_rotate(_p(emu_fbsk), _p(small), _p(lut), _p(a8), 1)
    d = ((a16 >> np.uint64(32)).astype(np.int64) - (a8 >> np.uint64(32)).astype(np.int64) + (1 << 31)) % (1 << 32) - (1 << 31)
    assert np.abs(d).max() <= 1024, np.abs(d).max()   # f64 granularity at the increments' magnitude (~2^27) is 2^-26 of the torus = 64 units of 2^-32


def test_w16_trivial_input_is_exact(emu_w16, emu_fbsk, server_key):
    lut = tfhe.make_lut(lambda x: (3 * x + 1) % 16)
    small = np.zeros(743, dtype=np.uint64)
    small[742] = np.uint64(11 << 59)
    acc = np.zeros(2 * 2048, dtype=np.uint64)
    emu_w16.emu_w16_blind_rotate(_p(emu_fbsk), _p(small), _p(lut), _p(acc), -1)
    ref = tfhe.bootstrap_small(server_key, small, lut)
    assert (tfhe.sample_extract(acc) == ref).all()


