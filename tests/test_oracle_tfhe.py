"""Oracle (TFHE arithmetic restatement) self-consistency + fixture checks.  CPU only."""
import numpy as np
import pytest

from oracle import tfhe


def test_fixture_layout(client_key):
    # offsets/values measured on test_data/client_key
    ck = client_key
    p = ck.params
    assert (p["lwe_dimension"], p["glwe_dimension"], p["polynomial_size"]) == (742, 1, 2048)
    assert (p["pbs_base_log"], p["pbs_level"], p["ks_base_log"], p["ks_level"]) == (23, 1, 3, 5)
    assert (p["message_modulus"], p["carry_modulus"], ck.num_blocks) == (4, 4, 4)
    assert p["lwe_modular_std_dev"] == 7.069849454709433e-06
    assert p["glwe_modular_std_dev"] == 2.9403601535432533e-16
    assert int(ck.big.sum()) == 1031 and int(ck.small.sum()) == 395
    assert (ck.big == ck.glwe).all() and set(np.unique(ck.big)) == {0, 1}


def test_decomposer_known_answers():
    # hand-derived from the tfhe-rs rule restated in tfhe_oracle.c (closest_representable + decompose_one_level)
    assert tfhe.decompose(1 << 63, 3, 5) == [0, 0, 0, 0, 4]
    assert tfhe.decompose((4 << 49) + (4 << 52), 3, 5) == [-4, -3, 1, 0, 0]
    assert tfhe.decompose((1 << 48), 3, 5) == [1, 0, 0, 0, 0]          # rounds up on bit 48
    assert tfhe.decompose((1 << 48) - 1, 3, 5) == [0, 0, 0, 0, 0]
    assert tfhe.decompose((1 << 64) - 1, 3, 5) == [0, 0, 0, 0, 0]      # wraps to 0
    assert tfhe.decompose(1 << 63, 23, 1) == [1 << 22]                   # tie keeps +B/2
    assert tfhe.decompose((1 << 63) + (1 << 41), 23, 1) == [-(1 << 22) + 1]
    assert tfhe.decompose((1 << 40), 23, 1) == [1]
    rng = np.random.default_rng(0)
    for x in rng.integers(0, 2 ** 64, size=200, dtype=np.uint64):
        x = int(x)
        d = tfhe.decompose(x, 3, 5)
        assert all(-4 <= v <= 4 for v in d)
        recomposed = sum(v << (49 + 3 * l) for l, v in enumerate(d)) % 2 ** 64
        assert recomposed == int(tfhe.lib().orc_closest_representable(x, 3, 5))


def test_modswitch():
    assert tfhe.modswitch(0) == 0
    assert tfhe.modswitch((1 << 52) - 1) == 1
    assert tfhe.modswitch(1 << 51) == 1
    assert tfhe.modswitch((1 << 51) - 1) == 0
    assert tfhe.modswitch((1 << 64) - 1) == 4096


def test_lut_layout():
    lut = tfhe.make_lut(lambda x: x)
    assert lut[0] == 0 and lut[63] == 0 and lut[64] == 1 << 59
    assert lut[2048 - 64] == (-(0 << 59)) % 2 ** 64
    lut = tfhe.make_lut(lambda x: 15 - x)
    assert lut[0] == 15 << 59 and lut[2047] == (-(15 << 59)) % 2 ** 64 and lut[2048 - 65] == 0


def test_encrypt_decrypt(client_key):
    cts = tfhe.encrypt_batch(client_key, range(16))
    assert [tfhe.decrypt_shortint(client_key, c) for c in cts] == list(range(16))
    for v in (0, 1, 97, 255):
        assert tfhe.decrypt_radix(client_key, tfhe.encrypt_radix(client_key, v, 3, v)) == v
        assert tfhe.decrypt_radix(client_key, tfhe.trivial_radix(v)) == v


def test_keyswitch_noise(client_key, server_key):
    msgs = np.arange(64) % 16
    cts = tfhe.encrypt_batch(client_key, msgs, seed=5)
    small = tfhe.keyswitch(server_key, cts)
    err = tfhe.torus_err(tfhe.phase_batch(client_key.small, small), msgs.astype(np.uint64) << np.uint64(59))
    # expected variance: 2048*5 * E[d^2]*sigma_lwe^2 + rounding(2^-15 steps over 1031 key bits) ~ 2.9e-6
    assert 1.0e-3 < err.std() < 2.6e-3
    assert np.abs(err).max() < 1 / 64


def test_keyswitch_trivial_is_trivial(server_key):
    small = tfhe.keyswitch(server_key, tfhe.trivial_shortint(7))[0]
    assert (small[:742] == 0).all() and small[742] == 7 << 59


def test_pbs_all_messages(client_key, server_key):
    msgs = np.arange(16)
    cts = tfhe.encrypt_batch(client_key, msgs, seed=7)
    fs = [lambda x: x, lambda x: (x * x) % 16, lambda x: int(x == 5), lambda x: int(x >= 1)]
    luts = np.stack([tfhe.make_lut(f) for f in fs])
    for li, f in enumerate(fs):
        out = tfhe.pbs(server_key, cts, luts, [li] * 16)
        assert [tfhe.decrypt_shortint(client_key, c) for c in out] == [f(int(m)) & 15 for m in msgs]
        exp = np.array([f(int(m)) & 15 for m in msgs], dtype=np.uint64) << np.uint64(59)
        err = tfhe.torus_err(tfhe.phase_batch(client_key.big, out), exp)
        assert np.abs(err).max() < 4e-4  # sigma_out ~ 3e-5


def test_pbs_trivial_input(client_key, server_key):
    # reference tests run on trivial ciphertexts (engine.rs:282-286): every CMUX is skipped
    lut = tfhe.make_lut(lambda x: (3 * x) % 16)
    out = tfhe.pbs(server_key, np.stack([tfhe.trivial_shortint(m) for m in range(16)]), lut[None], [0] * 16)
    assert [tfhe.decrypt_shortint(client_key, c) for c in out] == [(3 * m) % 16 for m in range(16)]
    assert (out[:, :2048] == 0).all()


def test_blind_rotate_fft_vs_exact(client_key, server_key):
    ct = tfhe.encrypt_shortint(client_key, 9, 11, 0)
    small = tfhe.keyswitch(server_key, ct)[0]
    lut = tfhe.make_lut(lambda x: x)
    o_fft = tfhe.sample_extract(tfhe.blind_rotate(server_key, small, lut))
    o_exact = tfhe.sample_extract(tfhe.blind_rotate(server_key, small, lut, exact=True))
    exp = np.array([9 << 59], dtype=np.uint64)
    for o in (o_fft, o_exact):
        assert tfhe.decrypt_shortint(client_key, o) == 9
        err = tfhe.torus_err(np.array([tfhe.phase_big(client_key, o)], dtype=np.uint64), exp)
        assert abs(err[0]) < 3e-4
