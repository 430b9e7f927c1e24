"""N>1 path on CPU: world_size-2 gloo run of the sharding logic the multi-GPU bench/has_match uses
(SURVEY.md 8e): every rank evaluates the variants of its start offsets (i % world == rank), the partial
booleans are all-gathered and OR-folded.  Ciphertexts are replaced by the plaintext dry run of the lowered
PBS plan (fb_plan_eval_plain), so no GPU is needed; the collective and the partition are the real ones."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

CASES = [("xxabbcxxxxaacxxx", "/a+b?c/"), ("xxabbcxxxxaacxxx", "/^zz/"), ("abcabcabc", "/ab{2,4}c/"), ("zzzzzzab", "/ab$/"),
         ("bq.", r"/^[a-d][^x-z]\.$/"), ("aaaaaaaaaaaaaaaa", "/a+b?c/")]


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_path):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import fhe_regex_b200 as fb
    from oracle import regex_plain as rp
    results = []
    for content, pattern in CASES:
        part = fb.plan_eval_plain(pattern, content, rank=rank, world=world)       # this rank's share of the variants
        t = torch.tensor([part], dtype=torch.int64)
        gathered = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
        dist.all_gather(gathered, t)                                               # the partial-result gather
        folded = int(any(int(g.item()) for g in gathered))                         # the final bitor fold (engine.rs:30-33)
        results.append((folded, rp.has_match(content, pattern)))
    # the batch split of the raw-PBS bench: contiguous, disjoint, covering
    B = 1000
    lo, hi = rank * B // world, (rank + 1) * B // world
    sizes = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(sizes, torch.tensor([hi - lo], dtype=torch.int64))
    total = sum(int(x.item()) for x in sizes)
    dist.barrier()
    if rank == 0:
        with open(out_path, "w") as f:
            f.write(repr((results, total)))
    dist.destroy_process_group()


def test_sharded_match_world2_gloo(tmp_path):
    out = str(tmp_path / "res.txt")
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    results, total = eval(open(out).read())
    assert total == 1000
    for (got, exp), (content, pattern) in zip(results, CASES):
        assert got == exp, (content, pattern)
