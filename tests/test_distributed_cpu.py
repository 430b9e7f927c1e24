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


# ---- the level-sharded collective match (fb_has_match_dist, regex_api.cu run_plan dist) on CPU -----------------------
# Same partition (slice r of every level = [n r / W, n (r + 1) / W), comm.cu::fb_comm_slice) and the same exchange pattern
# (every rank holds the whole arena; after a level each rank contributes its slice of the level's output rows), with rows
# holding plaintext messages instead of ciphertexts and gloo instead of NCCL.  The plan is the library's own
# (fb_plan_export); a PBS is its LUT (decoded from fb_regex_lut_table), a linear combination is integer arithmetic.
DIST_CASES = CASES + [("xabbcxabbbbcxx", "/ab{2,4}c/"), ("x" * 40 + "aabc" + "x" * 20, "/a+b?c/"), ("abab", "/^(ab|c)+$/"), ("", "/^$/")]


def _dist_level_worker(rank, world, port, out_path, shard_min=0):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import numpy as np
    import fhe_regex_b200 as fb
    from oracle import regex_plain as rp
    luts = fb.regex_lut_table()
    # f(x) of LUT id: the accumulator holds f(x) << 59 on the box of x (128 coefficients, centred on 128 x)
    lut_f = np.array([[int(luts[i][128 * x] >> np.uint64(59)) for x in range(16)] for i in range(luts.shape[0])], dtype=np.int64)
    results = []
    for content, pattern in DIST_CASES:
        for ref_shaped in (False, True):
            plan = fb.plan_export(pattern, len(content), reference_shaped=ref_shaped)
            if plan["result_kind"] < 2:
                results.append((plan["result_kind"], rp.has_match(content, pattern)))
                continue
            arena = torch.zeros(plan["n_rows"], dtype=torch.int64)
            blocks = torch.tensor([(ord(c) >> (2 * b)) & 3 for c in content for b in range(4)], dtype=torch.int64)
            # content: rank r "uploads" slice r, the rest arrives through the exchange
            n_in = blocks.numel()
            lo, hi = n_in * rank // world, n_in * (rank + 1) // world
            arena[lo:hi] = blocks[lo:hi]
            _exchange(arena, 0, n_in, rank, world)
            for lv in plan["levels"]:
                for o, row in enumerate(lv["lin_out_rows"]):
                    t0, t1 = int(lv["lin_term_off"][o]), int(lv["lin_term_off"][o + 1])
                    v = int(lv["lin_const"][o] >> np.uint64(59))
                    for t in range(t0, t1):
                        v += int(lv["lin_coef"][t]) * int(arena[int(lv["lin_term_rows"][t])])
                    arena[int(row)] = v
                n = len(lv["in_rows"])
                lo, hi = n * rank // world, n * (rank + 1) // world
                if n <= shard_min:                    # narrow level: every rank computes all of it, no exchange (option dist_shard_min)
                    lo, hi = 0, n
                base = lv["out_row_base"]
                outs = []
                for b in range(lo, hi):
                    x = int(arena[int(lv["in_rows"][b])])
                    assert 0 <= x <= 15, "PBS input outside the message space"
                    outs.append(int(lut_f[int(lv["lut_idx"][b]), x]))
                arena[base + lo: base + hi] = torch.tensor(outs, dtype=torch.int64)
                if n > shard_min:
                    arena[base: base + lo] = -99      # rows this rank did not compute: must come from the exchange
                    arena[base + hi: base + n] = -99
                    _exchange(arena, base, n, rank, world)
            results.append((int(arena[plan["result_row"]]), rp.has_match(content, pattern)))
    # every rank ends with the same result
    mine = torch.tensor([r[0] for r in results], dtype=torch.int64)
    allr = [torch.zeros_like(mine) for _ in range(world)]
    dist.all_gather(allr, mine)
    same = all(bool((a == mine).all()) for a in allr)
    dist.barrier()
    if rank == 0:
        with open(out_path, "w") as f:
            f.write(repr((results, same)))
    dist.destroy_process_group()


def _exchange(arena, base, n, rank, world):
    """in-place exchange of row slices: one broadcast per non-empty slice (comm.cu::fb_comm_exchange_rows)"""
    for root in range(world):
        lo, hi = n * root // world, n * (root + 1) // world
        if hi > lo:
            buf = arena[base + lo: base + hi].clone()
            dist.broadcast(buf, root)
            arena[base + lo: base + hi] = buf


@pytest.mark.parametrize("world,shard_min", [(2, 0), (3, 0), (2, 6)])
def test_level_sharded_match_gloo(tmp_path, world, shard_min):
    out = str(tmp_path / "res.txt")
    mp.spawn(_dist_level_worker, args=(world, _free_port(), out, shard_min), nprocs=world, join=True)
    results, same = eval(open(out).read())
    assert same
    assert len(results) == 2 * len(DIST_CASES)
    for k, (got, exp) in enumerate(results):
        assert got == exp, DIST_CASES[k // 2]
