#!/usr/bin/env python3
"""Level-sharded has_match over an NCCL communicator inside the library (fb_comm_init / fb_has_match_dist), N ranks:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29517 tools/dist_match_check.py

Checks: every rank returns the same ciphertext, it decrypts to the oracle's result, and equals (as a decryption) the
single-GPU fb_has_match; prints the wall time per match (max over ranks) next to the single-GPU time of rank 0."""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist
import fhe_regex_b200 as fb
from oracle import regex_plain as rp   # checker only


def main():
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    ck = fb.ClientKey.load(os.path.join(ROOT, "tests", "golden", "client_key"))
    ksk, bsk = fb.keygen_server_raw(ck, seed=0)
    sk = fb.ServerKey(ksk, bsk, device=local)
    idt = torch.zeros(128, dtype=torch.uint8, device=dev)
    if rank == 0:
        idt = torch.from_numpy(np.frombuffer(fb.comm_unique_id(), dtype=np.uint8).copy()).to(dev)
    if world > 1:
        dist.broadcast(idt, 0)
    sk.comm_init(idt.cpu().numpy().tobytes(), rank, world)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    rng = np.random.default_rng(5)
    c64 = "".join(rng.choice(list("abcx"), size=64))
    c256 = "".join(np.random.default_rng(6).choice(list("abx"), size=256))
    cases = [("xabbcx", "/ab{2,4}c/", 0), ("aq.", r"/^[a-d][^x-z]\.$/", 0), (c64, "/a+b?c/", 0), (c64, r"/[a-d][^x-z]\./", 0),
             (c256, "/a+b?c/", 0), (c256[:-3] + "abc", "/a+b?c/", 0), (c64, "/a+b?c/", 1), (c256, "/a+b?c/", 1)]
    out = []
    for content, pattern, ref_shaped in cases:
        sk.set_option("plan_reference_shaped", ref_shaped)
        ct = fb.encrypt_str(ck, content, seed=9)
        fb.has_match_dist(sk, ct, pattern)                 # cold: plan
        walls = []
        for _ in range(3):
            barrier()
            t = time.perf_counter()
            res, st = fb.has_match_dist(sk, ct, pattern, return_stats=True)
            barrier()
            walls.append((time.perf_counter() - t) * 1e3)
        wall = sorted(walls)[1]
        if world > 1:
            g = torch.from_numpy(res.view(np.int64)).to(dev)
            allg = torch.empty((world,) + tuple(g.shape), dtype=torch.int64, device=dev)
            dist.all_gather_into_tensor(allg, g)
            assert bool((allg == allg[0:1]).all()), "ranks disagree on the result ciphertext"
            tw = torch.tensor([wall], dtype=torch.float64, device=dev)
            dist.all_reduce(tw, op=dist.ReduceOp.MAX)
            wall = float(tw[0])
        if rank == 0:
            exp = rp.has_match(content, pattern) if len(content) <= 64 else int("c" in content)   # /a+b?c/ on {a,b,x}* + planted abc
            got = ck.decrypt(res)
            assert got == exp, (pattern, got, exp)
            t = time.perf_counter()
            solo, st1 = fb.has_match(sk, ct, pattern, return_stats=True)
            solo_ms = (time.perf_counter() - t) * 1e3
            t = time.perf_counter()
            solo, st1 = fb.has_match(sk, ct, pattern, return_stats=True)
            solo_ms = min(solo_ms, (time.perf_counter() - t) * 1e3)
            assert ck.decrypt(solo) == exp
            out.append({"pattern": pattern, "n_chars": len(content), "plan": "reference-shaped" if ref_shaped else "absorbed", "world": world,
                        "ms": wall, "gpu_ms_rank0": st["gpu_ms"], "single_gpu_ms": solo_ms, "pbs": st["pbs"], "levels": st["levels"], "result": got})
            print(json.dumps(out[-1]), flush=True)
        barrier()
    sk.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
