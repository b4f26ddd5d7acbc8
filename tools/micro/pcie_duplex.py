#!/usr/bin/env python
"""Ceiling of the host-buffer lookup pipeline: 800 MB host->device and 800 MB device->host moved
concurrently in 32 MB pieces on two streams with no kernels at all (pinned memory), next to the
library's e2e lookup of 1e8 k-mers (same byte counts) for several pipeline chunk sizes."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


def main():
    import torch
    import sshash_b200
    n = 100_000_000
    dev = torch.device("cuda", 0)
    h_in = torch.empty(n, dtype=torch.int64, pin_memory=True)
    h_out = torch.empty(n, dtype=torch.int64, pin_memory=True)
    d_in = torch.empty(n, dtype=torch.int64, device=dev)
    d_out = torch.empty(n, dtype=torch.int64, device=dev)
    res = {}
    piece = (32 << 20) // 8
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    for mode in ("h2d_only", "d2h_only", "duplex"):
        for it in range(4):
            if it == 1:
                torch.cuda.synchronize(); t0 = time.perf_counter()
            for lo in range(0, n, piece):
                hi = min(n, lo + piece)
                if mode != "d2h_only":
                    with torch.cuda.stream(s1):
                        d_in[lo:hi].copy_(h_in[lo:hi], non_blocking=True)
                if mode != "h2d_only":
                    with torch.cuda.stream(s2):
                        h_out[lo:hi].copy_(d_out[lo:hi], non_blocking=True)
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / 3
        res[mode] = {"ms": dt * 1e3, "GB_per_s_per_direction": n * 8 / dt / 1e9}
    res["duplex"]["lookups_per_s_ceiling"] = n / (res["duplex"]["ms"] * 1e-3)
    idx = os.path.join(ROOT, "tests", "golden", "se_k31_m13.sshash")
    for chunk in (8 << 20, 16 << 20, 32 << 20, 64 << 20):
        os.environ["SSHASH_GPU_BATCH_CHUNK"] = str(chunk)
        d = sshash_b200.Dictionary(idx)
        ids = torch.randint(0, d.num_kmers(), (n,), device=dev, dtype=torch.int64)
        h_in.copy_(d.access_batch(ids))
        a, b = h_in.numpy().view(np.uint64), h_out.numpy().view(np.uint64)
        d.lookup_batch(a, out=b)
        t0 = time.perf_counter()
        for _ in range(3):
            d.lookup_batch(a, out=b)
        dt = (time.perf_counter() - t0) / 3
        assert torch.equal(h_out, ids.cpu())
        res["e2e_chunk_%dMB" % (chunk >> 20)] = {"ms": dt * 1e3, "lookups_per_s": n / dt}
        d.close()
    print(json.dumps(res))


if __name__ == "__main__":
    main()
