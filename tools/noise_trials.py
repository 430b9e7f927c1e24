#!/usr/bin/env python3
"""Bootstrap correctness / noise over many trials on the GPU (north-star criterion: no decryption failure
across >= 10^6 trials, output-noise variance within the parameter set's bound).

    python tools/noise_trials.py [trials] [--variant throughput|latency|cluster] [--worst-case] [--json out.json]

Every trial is a fresh encryption (own mask, own Gaussian noise) of a uniform 4-bit message, bootstrapped
through a LUT drawn from the has_match table plus identity / affine LUTs.  Outputs are decrypted on the CPU
with the fixture secret key; the phase error against the expected plaintext is accumulated."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import fhe_regex_b200 as fb  # noqa: E402


def run(trials: int, chunk: int = 28416, seed: int = 2026, variant: str = "throughput"):
    """variant: which blind rotation every batch goes through -- throughput (up to 4 PBS per CTA; what wide batches
    get), latency (one PBS per CTA; what the narrow levels of a match get) or cluster (one PBS per pair of SMs)"""
    ck = fb.ClientKey.load(os.path.join(ROOT, "tests", "golden", "client_key"))
    ksk, bsk = fb.keygen_server_raw(ck, seed=0)
    sk = fb.ServerKey(ksk, bsk)
    sk.set_cluster_threshold((1 << 30) if variant == "cluster" else 0)
    sk.set_latency_threshold((1 << 30) if variant == "latency" else 0)
    fs = [lambda x: x, lambda x: (5 * x + 3) % 16, lambda x: int(x == 7), lambda x: int(x >= 1), lambda x: int(x > 9), lambda x: 15 - x,
          lambda x: int(x == 2), lambda x: int(x < 2)]
    luts = np.stack([fb.make_lut(f) for f in fs])
    tab = np.array([[f(m) & 15 for m in range(16)] for f in fs], dtype=np.uint64)
    key = ck.big.astype(np.uint64)
    rng = np.random.default_rng(seed)
    done = fails = 0
    s1 = s2 = 0.0
    amax = 0.0
    t0 = time.time()
    gpu_s = 0.0
    while done < trials:
        n = min(chunk, trials - done)
        msgs = rng.integers(0, 16, size=n)
        idx = rng.integers(0, len(fs), size=n).astype(np.uint32)
        cts = ck.encrypt_blocks(msgs, seed=seed + 1, stream0=done)     # fresh mask + noise per trial
        tg = time.time()
        out = sk.pbs(cts, luts, idx)
        gpu_s += time.time() - tg
        exp = tab[idx, msgs]
        with np.errstate(over="ignore"):
            phase = out[:, 2048] - (out[:, :2048] * key[None, :]).sum(axis=1, dtype=np.uint64)
            dec = ((phase + np.uint64(1 << 58)) >> np.uint64(59)) & np.uint64(15)
            err = (phase - (exp << np.uint64(59))).view(np.int64).astype(np.float64) / 2.0 ** 64
        fails += int((dec != exp).sum())
        s1 += float(err.sum())
        s2 += float((err * err).sum())
        amax = max(amax, float(np.abs(err).max()))
        done += n
    sk.close()
    mean = s1 / done
    var = s2 / done - mean * mean
    return {"variant": variant, "trials": done, "decryption_failures": fails, "err_mean": mean, "err_std": var ** 0.5, "err_var": var, "err_abs_max": amax,
            "expected_std_bound": 3.7e-5, "half_box": 1.0 / 64, "wall_s": time.time() - t0, "pbs_call_s": gpu_s}


def run_worst_case(trials: int, chunk: int = 28416, pool: int = 30 * 28416, seed: int = 2027, variant: str = "throughput"):
    """The noisiest input the match path produces: a sum of 15 BOOTSTRAPPED booleans (norm2 = 15, the widest k-ary
    and/or of the lowering, regex_host.cpp) -> keyswitch -> PBS through the LUTs that consume such sums (x == k, x >= 1).
    A pool of bootstrapped booleans is made first (fresh encryptions of random bits through an identity LUT, so every
    member carries real bootstrap output noise); every trial sums 15 distinct pool members on the device (torch, wrapping
    int64 adds -- linear glue, not part of the measured path) and bootstraps the sum through fb_pbs_batch_dev."""
    import torch
    ck = fb.ClientKey.load(os.path.join(ROOT, "tests", "golden", "client_key"))
    ksk, bsk = fb.keygen_server_raw(ck, seed=0)
    sk = fb.ServerKey(ksk, bsk)
    sk.set_cluster_threshold((1 << 30) if variant == "cluster" else 0)
    sk.set_latency_threshold((1 << 30) if variant == "latency" else 0)
    dev = torch.device("cuda", 0)
    fs = [(lambda k: (lambda x: int(x == k)))(k) for k in range(16)] + [lambda x: int(x >= 1)]
    luts = np.stack([fb.make_lut(f) for f in fs])
    d_luts = torch.from_numpy(luts.view(np.int64)).to(dev)
    ident = torch.from_numpy(fb.make_lut(lambda x: x & 1).view(np.int64)).to(dev).reshape(1, -1)
    key = torch.from_numpy(ck.big.astype(np.uint64).view(np.int64)).to(dev)
    rng = np.random.default_rng(seed)
    t0 = time.time()
    # pool of bootstrapped booleans, kept on the device
    bits = rng.integers(0, 2, size=pool)
    d_pool = torch.empty((pool, 2049), dtype=torch.int64, device=dev)
    zeros = torch.zeros(chunk, dtype=torch.int32, device=dev)
    for o in range(0, pool, chunk):
        n = min(chunk, pool - o)
        d_in = torch.from_numpy(ck.encrypt_blocks(bits[o:o + n], seed=seed + 1, stream0=o).view(np.int64)).to(dev)
        sk.pbs_dev(d_in.data_ptr(), ident.data_ptr(), zeros.data_ptr(), n, d_pool[o:o + n].data_ptr())
        sk.sync()
    d_bits = torch.from_numpy(bits).to(dev)
    done = fails = 0
    s1 = s2 = 0.0
    amax = 0.0
    hist = np.zeros(16, dtype=np.int64)
    tab = torch.tensor([[f(m) & 15 for m in range(16)] for f in fs], dtype=torch.int64, device=dev)
    while done < trials:
        n = min(chunk, trials - done)
        # 15 distinct members per trial: a random start and 15 strides through the pool
        start = torch.from_numpy(rng.integers(0, pool, size=n)).to(dev)
        step = torch.from_numpy(rng.integers(1, pool // 16, size=n)).to(dev)
        members = (start[:, None] + step[:, None] * torch.arange(15, device=dev)[None, :]) % pool
        d_sum = torch.zeros((n, 2049), dtype=torch.int64, device=dev)
        for t in range(15):
            d_sum += d_pool[members[:, t]]
        sums = d_bits[members].sum(dim=1)
        idx = torch.from_numpy(rng.integers(0, len(fs), size=n).astype(np.int32)).to(dev)
        d_out = torch.empty_like(d_sum)
        torch.cuda.synchronize()
        sk.pbs_dev(d_sum.data_ptr(), d_luts.data_ptr(), idx.data_ptr(), n, d_out.data_ptr())
        sk.sync()
        exp = tab[idx.long(), sums]
        phase = d_out[:, 2048] - (d_out[:, :2048] * key[None, :]).sum(dim=1)
        dec = ((phase + (1 << 58)) >> 59) & 15
        err = (phase - (exp << 59)).double() / 2.0 ** 64
        fails += int((dec != exp).sum())
        s1 += float(err.sum())
        s2 += float((err * err).sum())
        amax = max(amax, float(err.abs().max()))
        hist += np.bincount(sums.cpu().numpy(), minlength=16)[:16]
        done += n
    sk.close()
    mean = s1 / done
    var = s2 / done - mean * mean
    return {"variant": variant, "input": "sum of 15 bootstrapped booleans -> KS -> PBS (x == k, x >= 1)", "trials": done, "pool": pool,
            "decryption_failures": fails, "err_mean": mean, "err_std": var ** 0.5, "err_var": var, "err_abs_max": amax,
            "expected_std_bound": 3.7e-5, "half_box": 1.0 / 64, "input_sum_histogram": hist.tolist(), "wall_s": time.time() - t0}


if __name__ == "__main__":
    trials = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 1000000
    variant = sys.argv[sys.argv.index("--variant") + 1] if "--variant" in sys.argv else "throughput"
    res = run_worst_case(trials, variant=variant) if "--worst-case" in sys.argv else run(trials, variant=variant)
    print(json.dumps(res))
    if "--json" in sys.argv:
        with open(sys.argv[sys.argv.index("--json") + 1], "w") as f:
            json.dump(res, f, indent=1)
    sys.exit(1 if res["decryption_failures"] else 0)
