#!/usr/bin/env python3
"""bench.py -- batched programmable bootstraps per second on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--batch B]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W

A "step" is one pass of the hot path (keyswitch -> blind rotate -> sample extract, include/fhe_b200.h
fb_pbs_batch_dev) over one batch of B synthetic LWE ciphertexts per GPU: real encryptions of uniform
messages under the fixture secret key, LUT ids uniform over the table has_match uses.  The batch has the
shape of one DAG level of the 256-char contains-match /a+b?c/ (level widths up to 21 550),
rounded to a multiple of the SM-count quantum.  Every rank holds a replica of the server key and its own
batch (weak scaling, no data-path collective: every PBS of a level is independent).

  value     PBS/s, whole job, inputs resident in HBM, CUDA events on the context stream, max over ranks
  e2e       the same through the host-buffer C-ABI call fb_pbs_batch (pinned host -> device copy of the
            ciphertexts, device -> host copy of the results inside the timed region)
  roofline  blind-rotate kernel: algorithmic FP64 flop / measured launch duration vs the FP64 FMA-pipe
            peak measured live by fb_measure_fp64_peak (MEASURED_PEAKS.json has no FP64 figure)
  cpu_baseline  the CPU oracle (oracle/tfhe_oracle.c, a restatement of tfhe-rs 0.2.0's algorithm -- the
            reference itself is Rust and cannot be built here) timed on a bounded sample
  match     ms per regex match through fb_has_match (64-char content, BASELINE configs 3-5)

--impl reference times only the CPU oracle, all host threads.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

FLOP_PER_PBS = 742 * 262144            # 4 transforms x (5*1024*10 + 6*1024) + 4*1024*8 per CMUX
# DRAM bytes of one blind_rotate_fused launch at the default batch, from the ncu pass recorded in
# profiles/r02_traffic_bench_batch.csv (dram__bytes_read.sum + dram__bytes_write.sum, B = 28 416)
BR_TRAFFIC_MEASURED = {28416: 317583360 + 456988160}
BSK_BYTES = 742 * 4 * 1024 * 16
KS_MAC_PER_PBS = 2048 * 5 * 743
KS_BYTES_PER_LAUNCH_KEY = 2048 * 5 * 743 * 8
BIG, SMALL, POLY = 2049, 743, 2048
# what the path computes in: f64 negacyclic FFT; the blind-rotation accumulator lives on the top 32 torus bits (u32,
# narrower than the reference's u64 accumulator: DESIGN.md section 3 and tests/test_gpu_parity.py for the noise argument);
# the keyswitch is exact mod 2^64
DTYPE = "f64 FFT + u32 torus accumulator + u64 keyswitch (exact)"
DTYPE_CPU = "f64 FFT + u64 torus accumulator + u64 keyswitch"
METRIC = "bootstraps_per_sec"
UNIT = "PBS/s"
CK_PATH = os.path.join(ROOT, "tests", "golden", "client_key")


def env_int(name, default):
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md)."""

    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index: int):
        self.index, self.rows, self.proc, self.th = index, [], None, None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None
            return
        self.th = threading.Thread(target=self._read, daemon=True)
        self.th.start()

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [x.strip() for x in line.split(",")]))

    def stop(self, t0: float, t1: float) -> dict:
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        rows = [r for (t, r) in self.rows if t0 <= t <= t1 and len(r) >= 7] or [r for (_, r) in self.rows if len(r) >= 7]
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        sm = sorted(float(r[0]) for r in rows)
        reasons = []
        for i, name in ((3, "hw_slowdown"), (4, "hw_thermal_slowdown"), (5, "sw_thermal_slowdown"), (6, "sw_power_cap")):
            if any(r[i].lower().startswith("active") for r in rows):
                reasons.append(name)
        pw = [float(r[2]) for r in rows if r[2].replace(".", "", 1).isdigit()]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(rows[0][1]), "power_w_max": max(pw) if pw else None,
                "samples": len(rows), "reasons": reasons}


def lut_table():
    """the accumulator table fb_has_match uploads (regex_host.h LutId): x==v, x>v, sum==k, >=1, >=2, <2"""
    import fhe_regex_b200 as fb
    fs = [(lambda v: (lambda x: int(x == v)))(v) for v in range(16)] + [(lambda v: (lambda x: int(x > v)))(v) for v in range(16)]
    fs += [lambda x: int(x >= 1), lambda x: int(x >= 2), lambda x: int(x < 2), lambda x: x]
    return np.stack([fb.make_lut(f) for f in fs]), fs


def make_inputs(ck, count: int, seed: int):
    """`count` LWE ciphertexts: 512 distinct real encryptions tiled (the kernels' work does not depend on the
    plaintext), messages uniform in [0,16)"""
    rng = np.random.default_rng(seed)
    base = min(count, 512)
    msgs = rng.integers(0, 16, size=base)
    cts = ck.encrypt_blocks(msgs, seed=seed + 1)
    reps = (count + base - 1) // base
    return np.ascontiguousarray(np.tile(cts, (reps, 1))[:count]), np.tile(msgs, reps)[:count]


def host_threads() -> int:
    """host cores this process may use -- NOT omp_get_max_threads(): torchrun exports OMP_NUM_THREADS=1 to its workers"""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def cpu_pbs_rate(n_samples: int, threads: int, seed: int = 0):
    """time the CPU oracle on n_samples PBS with `threads` OpenMP threads"""
    from oracle import tfhe
    ock = tfhe.ClientKey.load(CK_PATH)
    osk = tfhe.keygen_server(ock, seed=0)
    osk.fbsk  # Fourier conversion outside the timed region (the reference converts at keygen too)
    msgs = np.arange(n_samples) % 16
    cts = tfhe.encrypt_batch(ock, msgs, seed=seed + 1)
    lut = tfhe.make_lut(lambda x: (x + 1) % 16)
    tfhe.pbs(osk, cts[: max(1, threads)], lut[None], np.zeros(max(1, threads), dtype=np.uint32), nthreads=threads)  # warm
    t = time.perf_counter()
    out = tfhe.pbs(osk, cts, lut[None], np.zeros(n_samples, dtype=np.uint32), nthreads=threads)
    dt = time.perf_counter() - t
    ok = all(tfhe.decrypt_shortint(ock, out[i]) == (int(msgs[i]) + 1) % 16 for i in range(0, n_samples, max(1, n_samples // 8)))
    assert ok, "CPU oracle produced a wrong decryption"
    return n_samples / dt, dt


def run_reference(args, rank: int):
    """--impl reference: the reference's CPU algorithm (oracle port; tfhe-rs is Rust and cannot be built here)
    on all host threads, bounded sample per step."""
    if rank != 0:
        return
    from oracle import tfhe
    threads = host_threads()
    per_step = max(threads * 32, 64)   # ~1-2 s of host work per step
    ock = tfhe.ClientKey.load(CK_PATH)
    osk = tfhe.keygen_server(ock, seed=0)
    osk.fbsk
    msgs = np.arange(per_step) % 16
    cts = tfhe.encrypt_batch(ock, msgs, seed=2)
    lut = tfhe.make_lut(lambda x: x)
    idx = np.zeros(per_step, dtype=np.uint32)
    for _ in range(min(args.warmup, 1)):
        tfhe.pbs(osk, cts, lut[None], idx, nthreads=threads)
    t = time.perf_counter()
    for _ in range(args.steps):
        tfhe.pbs(osk, cts, lut[None], idx, nthreads=threads)
    dt = time.perf_counter() - t
    v = per_step * args.steps / dt
    sample = "%d PBS per step x %d steps, OpenMP over the batch" % (per_step, args.steps)
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": DTYPE_CPU, "data": "synthetic",
        "config": {"workload": "batched PBS (KS->BR->SE), PARAM_MESSAGE_2_CARRY_2, CPU restatement of tfhe-rs 0.2.0 (oracle/tfhe_oracle.c)",
                   "batch_per_step": per_step},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }), file=_JSON_OUT, flush=True)


_JSON_OUT = sys.stdout


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=0, help="PBS per GPU per step (default: 48 x quantum ~ 21k)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-match", action="store_true")
    ap.add_argument("--cpu-samples", type=int, default=0)
    ap.add_argument("--option", action="append", default=[], metavar="NAME=VALUE", help="fb_set_option on the context (profiling / A-B runs)")
    args = ap.parse_args()

    # exactly ONE line on stdout: libraries (NCCL prints its version banner to stdout) get stderr instead
    global _JSON_OUT
    sys.stdout.flush()
    _JSON_OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)

    rank, world, local_rank = env_int("RANK", 0), env_int("WORLD_SIZE", 1), env_int("LOCAL_RANK", 0)
    if args.impl == "reference":
        run_reference(args, rank)
        return

    import torch
    import torch.distributed as dist
    import fhe_regex_b200 as fb

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: libfhe_b200 has no CPU fallback (use --impl reference for the CPU oracle)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    L = fb.lib()
    ck = fb.ClientKey.load(CK_PATH)
    ksk, bsk = fb.keygen_server_raw(ck, seed=0)       # ServerKey::new analogue, replica on every rank
    sk = fb.ServerKey(ksk, bsk, device=local_rank)
    del ksk, bsk
    for ov in args.option:
        name, _, val = ov.partition("=")
        sk.set_option(name, int(val))
    q = sk.pbs_quantum()
    B = args.batch if args.batch > 0 else 48 * q
    luts_np, fs = lut_table()
    cts_np, msgs = make_inputs(ck, B, seed=100 + rank)
    idx_np = np.random.default_rng(7 + rank).integers(0, luts_np.shape[0], size=B).astype(np.uint32)

    dev = torch.device("cuda", local_rank)
    h_in = torch.from_numpy(cts_np.view(np.int64)).pin_memory()
    h_out = torch.empty_like(h_in).pin_memory()
    d_in = h_in.to(dev)
    d_out = torch.empty_like(d_in)
    d_luts = torch.from_numpy(luts_np.view(np.int64)).to(dev)
    d_idx = torch.from_numpy(idx_np.view(np.int32)).to(dev)
    torch.cuda.synchronize()
    ext = torch.cuda.ExternalStream(sk.stream, device=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        sk.sync()

    def step_dev():
        sk.pbs_dev(d_in.data_ptr(), d_luts.data_ptr(), d_idx.data_ptr(), B, d_out.data_ptr())

    def step_host():
        sk._check(L.fb_pbs_batch(sk._h, h_in.data_ptr(), luts_np.ctypes.data, luts_np.shape[0], idx_np.ctypes.data, B, h_out.data_ptr()))

    # ---- device-resident leg -----------------------------------------------------------------------
    for _ in range(max(args.warmup, 3)):
        step_dev()
    sk.sync()
    sk.timing(True)
    sk.kernel_stats(reset=True)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.time()
    ev0.record(ext)
    for _ in range(args.steps):
        step_dev()
    ev1.record(ext)
    barrier()
    t1 = time.time()
    ms = ev0.elapsed_time(ev1)
    kst = sk.kernel_stats(reset=True)
    sk.timing(False)
    clocks = sampler.stop(t0, t1) if rank == 0 else None

    # correctness of what was timed: decrypt a sample of the outputs
    out_np = d_out.cpu().numpy().view(np.uint64)
    chk = np.linspace(0, B - 1, num=min(B, 64), dtype=np.int64)
    for i in chk:
        exp = fs[int(idx_np[i])](int(msgs[i])) & 15
        got = ck.decrypt_block(out_np[i])
        assert got == exp, "PBS output %d decrypts to %d, expected %d" % (i, got, exp)

    # ---- host-buffer (e2e) leg ---------------------------------------------------------------------
    for _ in range(2):
        step_host()
    sampler_h = ClockSampler(local_rank)
    if rank == 0:
        sampler_h.start()
        time.sleep(0.3)
    sk.timing(True)
    sk.kernel_stats(reset=True)
    barrier()
    th0 = time.time()
    te = time.perf_counter()
    for _ in range(args.steps):
        step_host()
    barrier()
    e2e_s = time.perf_counter() - te
    th1 = time.time()
    kst_h = sk.kernel_stats(reset=True)      # the same kernels, timed inside the host-buffer calls
    sk.timing(False)
    clocks_h = sampler_h.stop(th0, th1) if rank == 0 else None
    # correctness of what the host-buffer leg produced: decrypt the same sample of ITS outputs
    h_np = h_out.numpy().view(np.uint64)
    for i in chk:
        exp = fs[int(idx_np[i])](int(msgs[i])) & 15
        got = ck.decrypt_block(h_np[i])
        assert got == exp, "fb_pbs_batch output %d decrypts to %d, expected %d" % (i, got, exp)

    if world > 1:
        t = torch.tensor([ms, e2e_s * 1e3], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, e2e_ms = float(t[0]), float(t[1])
    else:
        e2e_ms = e2e_s * 1e3

    value = world * B * args.steps / (ms * 1e-3)
    e2e_value = world * B * args.steps / (e2e_ms * 1e-3)

    line = None
    if rank == 0:
        fp64_peak = sk.fp64_peak_tflops(5)
        br_ms = kst["br_ms"] / max(1, kst["br_launches"])
        ks_ms = kst["ks_ms"] / max(1, kst["ks_launches"])
        achieved = B * FLOP_PER_PBS / (br_ms * 1e-3) / 1e12
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except (OSError, ValueError):
            pass
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        ks_bytes = KS_BYTES_PER_LAUNCH_KEY + B * (BIG + SMALL) * 8
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": DTYPE, "data": "synthetic",
            "config": {"workload": "batched PBS (KS->BR->SE), PARAM_MESSAGE_2_CARRY_2 (n=742,N=2048,k=1), one DAG-level-sized batch of the 256-char /a+b?c/ match",
                       "batch_per_gpu": B, "global_batch": B * world, "luts": int(luts_np.shape[0]),
                       "l2": "inputs+outputs %.0f MB per step > 126 MB L2; keys (109 MB) are re-read by design" % (2 * B * BIG * 8 / 1e6),
                       "parallelism": "replicated keys, batch sharded x%d" % world},
            "roofline": {"kernel": "blind_rotate_fused_kernel<%d, %d> (br_fused.cu)" % (
                             sk.get_option("br_samples"),
                             (sk.get_option("br_variant") - 1) | (8 if sk.get_option("br_planes") == 2 and sk.get_option("br_samples") == 4 else 0)
                             | (64 if sk.get_option("br_planes") == 3 and sk.get_option("br_samples") == 4 else 0)), "bound": "fp64", "achieved": achieved, "peak": fp64_peak, "unit": "TFLOP/s",
                         "frac": achieved / fp64_peak if fp64_peak else None, "traffic": BR_TRAFFIC_MEASURED.get(B),
                         "traffic_unit": "B/launch (ncu dram read+write, profiles/r02_traffic_bench_batch.csv)",
                         "algorithmic_bytes": BSK_BYTES + B * (SMALL + BIG) * 8 + int(luts_np.nbytes),
                         "peak_source": "measured live: fb_measure_fp64_peak (DFMA chains); nominal 148 SM x 64 FMA/clk x 2 x 1.965 GHz = 37.2",
                         "flop_per_pbs": FLOP_PER_PBS, "avg_launch_ms": br_ms, "share_of_step": kst["br_ms"] / (kst["br_ms"] + kst["ks_ms"] + kst["lin_ms"]),
                         "keyswitch": {"bound": "int/hbm", "avg_launch_ms": ks_ms, "gmac_per_s": B * KS_MAC_PER_PBS / (ks_ms * 1e-3) / 1e9,
                                       "algorithmic_gbs": ks_bytes / (ks_ms * 1e-3) / 1e9, "hbm_peak_gbs": hbm_peak,
                                       "hbm_peak_source": "MEASURED_PEAKS.json" if peaks else "fallback"}},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": B * BIG * 8 + luts_np.nbytes + idx_np.nbytes, "d2h_bytes_per_step": B * BIG * 8,
                    "ms_per_step": e2e_ms / args.steps, "api": "fb_pbs_batch (host buffers, pinned)",
                    "kernel_ms_per_step": (kst_h["br_ms"] + kst_h["ks_ms"]) / args.steps, "clocks": clocks_h},
            # kernels of this repository launched inside the timed region: per step ks_decompose + ks_gemm + blind_rotate
            "gpu_launches": int(2 * kst["ks_launches"] + kst["br_launches"] + kst["lin_launches"]),
            "clocks": clocks,
        }

    # ---- regex matches through the reference-facing entry point -------------------------------------
    if not args.no_match:
        from oracle import regex_plain as rp   # checker only
        rng = np.random.default_rng(5)
        c64 = "".join(rng.choice(list("abcx"), size=64))
        matches = []
        c256 = "".join(np.random.default_rng(6).choice(list("abx"), size=256))   # no 'c': every variant must be evaluated
        # (content, pattern, reference_shaped): the default plan absorbs OR operands implied by another operand
        # (decrypt-identical, O(n) instead of O(n^2) PBS for /a+.../); option plan_reference_shaped evaluates every variant
        # the reference enumerates -- the large sharded PBS batch BASELINE.json's config 5 describes.
        _exp_memo = {}

        def expected(content, pattern):   # the Python oracle takes ~9 s on 256 characters: once per (content, pattern)
            if (content, pattern) not in _exp_memo:
                _exp_memo[(content, pattern)] = rp.has_match(content, pattern)
            return _exp_memo[(content, pattern)]

        cases = [(c64, "/a+b?c/", False), (c64, "/ab{2,4}c/", False), (c64, r"/[a-d][^x-z]\./", False), (c256, "/a+b?c/", False),
                 (c64, "/a+b?c/", True), (c256, "/a+b?c/", True)]
        # N > 1: the contexts form an NCCL communicator inside the library (fb_comm_init) and the match is the collective
        # fb_has_match_dist: every PBS level of the plan cut into N slices, slices exchanged over NVLink in the arena --
        # no host hop, the bitor fold is the plan's last levels.  N = 1: fb_has_match.
        if world > 1:
            idt = torch.zeros(128, dtype=torch.uint8, device=dev)
            if rank == 0:
                idt = torch.from_numpy(np.frombuffer(fb.comm_unique_id(), dtype=np.uint8).copy()).to(dev)
            dist.broadcast(idt, 0)
            sk.comm_init(idt.cpu().numpy().tobytes(), rank, world)
        for content, pattern, ref_shaped in cases:
            sk.set_option("plan_reference_shaped", 1 if ref_shaped else 0)
            ct = fb.encrypt_str(ck, content, seed=9)
            run = (lambda: fb.has_match_dist(sk, ct, pattern, return_stats=True)) if world > 1 else \
                  (lambda: fb.has_match(sk, ct, pattern, return_stats=True))
            barrier()
            tc = time.perf_counter()
            run()                                                      # cold: parses, enumerates variants, lowers to a PBS plan
            cold = (time.perf_counter() - tc) * 1e3
            walls = []
            for _rep in range(3):                                      # warm: plan cached in the context (same pattern, same length)
                barrier()
                tm = time.perf_counter()
                part, st = run()
                barrier()
                walls.append((time.perf_counter() - tm) * 1e3)
            wall = sorted(walls)[1]
            if world > 1:
                tw = torch.tensor([wall], dtype=torch.float64, device=dev)
                dist.all_reduce(tw, op=dist.ReduceOp.MAX)
                wall = float(tw[0])
            if rank == 0:
                res = ck.decrypt(part)
                exp = expected(content, pattern)
                assert res == exp, (pattern, res, exp)
                matches.append({"pattern": pattern, "n_chars": len(content), "plan": "reference-shaped" if ref_shaped else "absorbed",
                                "ms": wall, "ms_cold_rank0": cold, "gpu_ms_rank0": st["gpu_ms"], "pbs": st["pbs"], "n_gpus": world,
                                "levels": st["levels"], "level_widths": fb.plan_level_widths(pattern, len(content), reference_shaped=ref_shaped),
                                "api": "fb_has_match_dist (level-sharded, NCCL in the library)" if world > 1 else "fb_has_match",
                                "ref_ct_ops": st["ct_ops"], "result": res})
        sk.set_option("plan_reference_shaped", 0)
        # many contents against one pattern in shared launches (fb_has_match_many): the levels are wide enough for
        # the throughput kernel, a match costs its PBS at the throughput rate instead of one latency per level
        many = []
        # every rank matches its own m contents (weak scaling: documents are independent, no collective on the data path)
        for n_chars, m, pattern in ((64, 64, "/a+b?c/"), (256, 16, "/a+b?c/")):
            rng2 = np.random.default_rng(11)
            # distinct contents: 4 at 64 characters; at 256 the no-match content of the single-match record and a
            # copy of it with a match planted at the end (2 oracle runs of ~9 s are enough)
            distinct = (["".join(rng2.choice(list("abcx"), size=n_chars)) for _ in range(4)] if n_chars == 64
                        else [c256, c256[:-3] + "abc"])
            base = [fb.encrypt_str(ck, t_, seed=20 + i) for i, t_ in enumerate(distinct)]
            texts = [distinct[i % len(distinct)] for i in range(m)]
            cts = fb.pinned_empty((m,) + base[0].shape, np.uint64)      # what a serving host stages its documents in (fb_host_alloc)
            for i in range(m):
                cts[i] = base[i % len(distinct)]
            fb.has_match_many(sk, cts, pattern)
            barrier()
            tm = time.perf_counter()
            outs, st = fb.has_match_many(sk, cts, pattern, return_stats=True)
            barrier()
            wall = (time.perf_counter() - tm) * 1e3
            got = [int(ck.decrypt(o)) for o in outs]
            if world > 1:     # the ranks hold the same contents: rank 0 checks everybody's decryptions against the oracle
                mine = torch.tensor(got, dtype=torch.int64, device=dev)
                allgot = torch.empty((world, m), dtype=torch.int64, device=dev)
                dist.all_gather_into_tensor(allgot, mine)
                tw = torch.tensor([wall], dtype=torch.float64, device=dev)
                dist.all_reduce(tw, op=dist.ReduceOp.MAX)
                wall = float(tw[0])
                got_all = allgot.cpu().tolist()
            else:
                got_all = [got]
            if rank == 0:
                exp_list = [expected(t_, pattern) for t_ in texts]
                assert all(g == exp_list for g in got_all), (pattern, got_all)
                many.append({"pattern": pattern, "n_chars": n_chars, "contents_per_gpu": m, "contents": m * world, "ms_total": wall,
                             "ms_per_match": wall / (m * world), "gpu_ms_total_rank0": st["gpu_ms"], "pbs_per_match": st["pbs"],
                             "matches_per_s": m * world / (wall * 1e-3)})
        if rank == 0:
            line["match_many"] = many
            line["match"] = matches
            line["ms_per_match_64"] = matches[0]["ms"]
            line["ms_per_match_256"] = matches[3]["ms"]
            line["ms_per_match_64_reference_shaped"] = matches[4]["ms"]
            line["ms_per_match_256_reference_shaped"] = matches[5]["ms"]

    sk.close()
    if rank == 0:
        if not args.no_cpu_baseline:
            threads = host_threads()
            n = args.cpu_samples or max(256, threads * 256)   # ~10-15 s on the box's host cores
            v, dt = cpu_pbs_rate(n, threads)
            v1, dt1 = cpu_pbs_rate(64, 1)                     # ~3 s: the reference itself is single-threaded (execution.rs:76-190)
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": threads, "kind": "port",
                                    "sample": "%d PBS, OpenMP over the batch, %.1f s; single thread: %.1f PBS/s (%.1f ms/PBS)" % (n, dt, v1, 1e3 / v1),
                                    "single_thread_value": v1}
        else:
            line["cpu_baseline"] = None
        print(json.dumps(line), file=_JSON_OUT, flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
