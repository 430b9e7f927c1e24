#!/usr/bin/env python3
"""Raw-PBS microbenchmark of SURVEY.md 8(d): B independent LWE inputs (messages uniform in [0,16), LUT ids uniform over the
has_match table), KS + PBS per second at B in {1, 64, 1 024, 16 384, 131 072}; device-resident (fb_pbs_batch_dev, CUDA
events through the library's kernel timers) and through host buffers (fb_pbs_batch, wall clock).
usage: tools/batch_sweep.py [out.json] [sizes...]"""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import fhe_regex_b200 as fb
from bench import lut_table, make_inputs, CK_PATH

out_path = sys.argv[1] if len(sys.argv) > 1 and not sys.argv[1].isdigit() else None
sizes = [int(a) for a in sys.argv[1:] if a.isdigit()] or [1, 64, 1024, 16384, 131072]
ck = fb.ClientKey.load(CK_PATH)
ksk, bsk = fb.keygen_server_raw(ck, seed=0)
sk = fb.ServerKey(ksk, bsk)
L = fb.lib()
luts_np, fs = lut_table()
dev = torch.device("cuda", 0)
d_luts = torch.from_numpy(luts_np.view(np.int64)).to(dev)
rows = []
for B in sizes:
    cts, msgs = make_inputs(ck, B, seed=300)
    idx = np.random.default_rng(3).integers(0, luts_np.shape[0], size=B).astype(np.uint32)
    h_in = torch.from_numpy(cts.view(np.int64)).pin_memory()
    h_out = torch.empty_like(h_in).pin_memory()
    d_in, d_idx = h_in.to(dev), torch.from_numpy(idx.view(np.int32)).to(dev)
    d_out = torch.empty_like(d_in)
    reps = 3 if B >= 16384 else 10
    for _ in range(2):
        sk.pbs_dev(d_in.data_ptr(), d_luts.data_ptr(), d_idx.data_ptr(), B, d_out.data_ptr())
    sk.sync()
    sk.timing(True)
    sk.kernel_stats(reset=True)
    t = time.perf_counter()
    for _ in range(reps):
        sk.pbs_dev(d_in.data_ptr(), d_luts.data_ptr(), d_idx.data_ptr(), B, d_out.data_ptr())
    sk.sync()
    wall_dev = (time.perf_counter() - t) / reps
    st = sk.kernel_stats(reset=True)
    sk.timing(False)
    out = d_out.cpu().numpy().view(np.uint64)
    for i in np.linspace(0, B - 1, num=min(B, 16), dtype=np.int64):
        assert ck.decrypt_block(out[i]) == fs[int(idx[i])](int(msgs[i])) & 15
    sk._check(L.fb_pbs_batch(sk._h, h_in.data_ptr(), luts_np.ctypes.data, luts_np.shape[0], idx.ctypes.data, B, h_out.data_ptr()))
    t = time.perf_counter()
    for _ in range(reps):
        sk._check(L.fb_pbs_batch(sk._h, h_in.data_ptr(), luts_np.ctypes.data, luts_np.shape[0], idx.ctypes.data, B, h_out.data_ptr()))
    wall_host = (time.perf_counter() - t) / reps
    gpu_ms = (st["br_ms"] + st["ks_ms"]) / reps
    rows.append({"batch": B, "gpu_ms": gpu_ms, "br_ms": st["br_ms"] / reps, "ks_ms": st["ks_ms"] / reps, "pbs_per_s_device": B / (gpu_ms * 1e-3),
                 "wall_ms_device_call": wall_dev * 1e3, "wall_ms_host_buffers": wall_host * 1e3, "pbs_per_s_host_buffers": B / wall_host})
    print(json.dumps(rows[-1]), flush=True)
    del h_in, h_out, d_in, d_out
sk.close()
if out_path:
    json.dump(rows, open(out_path, "w"), indent=1)
