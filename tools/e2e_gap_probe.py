#!/usr/bin/env python3
"""Where does the time of fb_pbs_batch (host buffers) go beyond the kernels?  Page-locked buffers from fb_host_alloc, 28 416
PBS, kernel timing on: wall time of the call against the keyswitch + blind-rotation time measured inside it, per chunk
schedule (option pbs_chunks), next to the same batch device-resident."""
import ctypes as C, json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import fhe_regex_b200 as fb  # noqa: E402

def main():
    ck = fb.ClientKey.load(os.path.join(ROOT, "tests", "golden", "client_key"))
    ksk, bsk = fb.keygen_server_raw(ck, seed=0)
    sk = fb.ServerKey(ksk, bsk)
    B = 28416
    base = ck.encrypt_blocks(np.arange(256) % 16, seed=3)
    h_in = fb.pinned_empty((B, fb.BIG)); h_out = fb.pinned_empty((B, fb.BIG))
    h_in[:] = np.tile(base, (B // 256, 1))
    n_luts = int(os.environ.get("PROBE_LUTS", "1"))        # bench.py uses 36 accumulators and uniform indices
    lut = np.ascontiguousarray(np.stack([fb.make_lut(lambda x, k=k: (3 * x + 1 + (0 if os.environ.get("PROBE_SAME") else k)) % 16) for k in range(n_luts)]))
    idx = np.random.default_rng(7).integers(0, n_luts, size=B).astype(np.uint32)
    if os.environ.get("PROBE_SORTED"):
        idx = np.sort(idx)
    L = fb.lib()
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    def call():
        rc = L.fb_pbs_batch(sk._h, p(h_in), p(lut), n_luts, p(idx), B, p(h_out)); assert rc == 0
    sk.timing(True)
    sk.set_option("br_sync", int(os.environ.get("PROBE_SYNC", "1")))
    sk.set_option("br_resync", int(os.environ.get("PROBE_RESYNC", "8")))
    for chunks in (3,):
        sk.set_option("pbs_chunks", chunks)
        call(); call()
        sk.kernel_stats(reset=True)
        t = time.perf_counter()
        for _ in range(3): call()
        wall = (time.perf_counter() - t) / 3 * 1e3
        st = sk.kernel_stats(reset=True)
        sk.set_option("plan_timing", 1); call(); sk.set_option("plan_timing", 0)   # timeline of one call on stderr
        print(json.dumps({"pbs_chunks": chunks, "wall_ms_per_call": wall, "br_ms_per_call": st["br_ms"] / 3, "ks_ms_per_call": st["ks_ms"] / 3,
                          "br_launches_per_call": st["br_launches"] / 3, "outside_kernels_ms": wall - (st["br_ms"] + st["ks_ms"]) / 3}), flush=True)
    # the same batch device-resident (fb_pbs_batch_dev on torch tensors), same accumulators and indices
    try:
        import torch
        d_in = torch.from_numpy(h_in.view(np.int64).copy()).cuda(); d_out = torch.empty_like(d_in)
        d_l = torch.from_numpy(lut.view(np.int64)).cuda(); d_i = torch.from_numpy(idx.view(np.int32)).cuda()
        for _ in range(2): sk.pbs_dev(d_in.data_ptr(), d_l.data_ptr(), d_i.data_ptr(), B, d_out.data_ptr())
        sk.sync(); sk.kernel_stats(reset=True)
        for _ in range(3): sk.pbs_dev(d_in.data_ptr(), d_l.data_ptr(), d_i.data_ptr(), B, d_out.data_ptr())
        sk.sync(); st = sk.kernel_stats(reset=True)
        print(json.dumps({"device_resident": True, "br_ms_per_call": st["br_ms"] / 3, "ks_ms_per_call": st["ks_ms"] / 3}), flush=True)
    except ImportError:
        pass
    assert ck.decrypt_block(h_out[5]) == (3 * 5 + 1 + (0 if os.environ.get("PROBE_SAME") else int(idx[5]))) % 16
    sk.close()

if __name__ == "__main__":
    main()
