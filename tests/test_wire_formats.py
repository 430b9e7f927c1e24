"""tfhe-rs 0.2.0 wire formats at the boundary (fhe_regex_b200/csrc/wire.cpp): bincode of integer::ServerKey,
RadixCiphertext and StringCiphertext = Vec<RadixCiphertext> -- what the reference's keygen / encrypt_str produce and
has_match consumes (/root/reference/src/regex/engine.rs:8-12, :248-254; ciphertext.rs:6, :29-45).  Host only."""
import struct

import numpy as np
import pytest

import fhe_regex_b200 as fb
from oracle import tfhe


def brev10(k):
    return int("{:010b}".format(k)[::-1], 2)


def natural_fourier_key(osk) -> np.ndarray:
    """the oracle's Fourier key (bit-reversed frequency order, re block then im block per polynomial) re-ordered into the
    natural order of the serialized tfhe-rs key: [742][2][2][1024][re, im]"""
    f = np.asarray(osk.fbsk).reshape(742 * 4, 2, 1024)
    perm = np.array([brev10(k) for k in range(1024)])
    out = np.empty((742 * 4, 1024, 2), dtype=np.float64)
    out[:, :, 0] = f[:, 0, perm]
    out[:, :, 1] = f[:, 1, perm]
    return out.reshape(742, 2, 2, 1024, 2)


def test_server_key_bincode_layout_and_round_trip(server_key):
    fk = natural_fourier_key(server_key)
    blob = fb.server_key_to_bincode(server_key.ksk, fk)
    assert len(blob) == fb.lib().fb_server_key_bincode_size() == 8 + 2048 * 5 * 743 * 8 + 24 + 24 + 2968 * (8 + 16384) + 32 + 24
    u = lambda off: struct.unpack_from("<Q", blob, off)[0]
    # LweKeyswitchKey { data, decomp_base_log, decomp_level_count, output_lwe_size }
    assert u(0) == 2048 * 5 * 743
    o = 8 + 2048 * 5 * 743 * 8
    assert (u(o), u(o + 8), u(o + 16)) == (3, 5, 743)
    # FourierPolynomialList: seq of 2 + count, polynomial_size, count, then per polynomial len + c64s
    assert (u(o + 24), u(o + 32), u(o + 40), u(o + 48)) == (2 + 2968, 2048, 2968, 1024)
    assert struct.unpack_from("<dd", blob, o + 56) == (fk.reshape(-1)[0], fk.reshape(-1)[1])
    tail = len(blob) - 56
    assert [u(tail + 8 * i) for i in range(7)] == [742, 2, 23, 1, 4, 4, 15]
    ksk, fk2 = fb.server_key_from_bincode(blob)
    assert (ksk == server_key.ksk).all() and (fk2 == fk).all()


def test_server_key_bincode_rejects_anything_else(server_key):
    fk = natural_fourier_key(server_key)
    blob = bytearray(fb.server_key_to_bincode(server_key.ksk, fk))
    with pytest.raises(fb.FbError):
        fb.server_key_from_bincode(bytes(blob[:-8]))                      # truncated
    bad = bytearray(blob)
    struct.pack_into("<Q", bad, 8 + 2048 * 5 * 743 * 8, 4)              # keyswitch base log 4: another parameter set
    with pytest.raises(fb.FbError):
        fb.server_key_from_bincode(bytes(bad))
    bad = bytearray(blob)
    struct.pack_into("<d", bad, 8 + 2048 * 5 * 743 * 8 + 56, float("nan"))
    with pytest.raises(fb.FbError):
        fb.server_key_from_bincode(bytes(bad))
    with pytest.raises(fb.FbError):
        fb.server_key_from_bincode(open(__file__, "rb").read())


def test_string_ciphertext_bincode_round_trip(client_key):
    fck = fb.ClientKey.load(tfhe_fixture())
    ct = fb.encrypt_str(fck, "aBc.", seed=5)
    blob = fb.string_ciphertext_to_bincode(ct)
    per = fb.lib().fb_radix_bincode_size()
    assert per == 8 + 4 * (8 + 2049 * 8 + 24) and len(blob) == 8 + 4 * per
    assert struct.unpack_from("<Q", blob, 0)[0] == 4 and struct.unpack_from("<Q", blob, 8)[0] == 4
    assert struct.unpack_from("<Q", blob, 16)[0] == 2049
    # block 0 of character 0: words, then degree 3 (fresh), message modulus 4, carry modulus 4
    assert struct.unpack_from("<QQQ", blob, 24 + 2049 * 8) == (3, 4, 4)
    back = fb.string_ciphertext_from_bincode(blob)
    assert back.shape == ct.shape and (back == ct).all()
    assert [fck.decrypt(c) for c in back] == [ord(c) for c in "aBc."]
    # the empty string (engine.rs:59-65 handles n = 0) and malformed blobs
    assert fb.string_ciphertext_from_bincode(fb.string_ciphertext_to_bincode(ct[:0])).shape == (0, 4, 2049)
    with pytest.raises(fb.FbError):
        fb.string_ciphertext_from_bincode(blob[:-1])
    bad = bytearray(blob)
    struct.pack_into("<Q", bad, 24 + 2049 * 8, 7)      # a block with a used carry: not what encrypt_str produces
    with pytest.raises(fb.FbError):
        fb.string_ciphertext_from_bincode(bytes(bad))


def test_result_radix_bincode(client_key):
    """the RadixCiphertext has_match returns (block 0 the boolean, blocks 1-3 trivial zero) serializes like
    create_trivial_radix's output would (ciphertext.rs:8-30) and decrypts through the oracle"""
    fck = fb.ClientKey.load(tfhe_fixture())
    res = fb.trivial_str("\x01")[0]
    blob = fb.radix_to_bincode(res, degrees=[1, 0, 0, 0])
    ct, deg = fb.radix_from_bincode(blob)
    assert (ct == res).all() and list(deg) == [1, 0, 0, 0] and fck.decrypt(ct) == 1


def tfhe_fixture():
    import os
    return os.path.join(os.path.dirname(__file__), "golden", "client_key")
