#!/usr/bin/env python
"""Single-GPU replay of the per-rank query streams of tools/cfg5_sharded.py (seed 1000 + rank):
positives must return the sampled ids; mismatches are printed and re-run on the reference."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))


def main():
    import torch
    import sshash_b200
    from bench import rc_packed_torch
    from bench_configs import build_index
    from oracle import ref
    wd = sys.argv[1] if len(sys.argv) > 1 else "/tmp/w"
    os.makedirs(wd, exist_ok=True)
    idx, _ = build_index(wd, 2500000, 1030, 31, 21)
    d = sshash_b200.Dictionary(idx)
    dev = torch.device("cuda", 0)
    B = 125_000_000
    k, nk = d.k(), d.num_kmers()
    bad_total = 0
    for rank in range(8):
        gen = torch.Generator(device=dev).manual_seed(1000 + rank)
        ids = torch.randint(0, nk, (B // 2,), generator=gen, device=dev, dtype=torch.int64)
        pos = d.access_batch(ids)
        pos[1::2] = rc_packed_torch(pos[1::2], k)
        out = d.lookup_batch(pos)
        bad = (out != ids).nonzero().flatten()
        print("rank-stream", rank, "mismatches", bad.numel(), flush=True)
        if bad.numel():
            bad_total += bad.numel()
            b = bad[:8]
            km = pos[b].cpu().numpy().view(np.uint64)
            rd = ref.RefDictionary(idx, max_k=31)
            want = rd.lookup(km)
            rd.close()
            print(json.dumps({"positions": b.tolist(), "kmers": [hex(int(x)) for x in km], "sampled_ids": ids[b].tolist(),
                              "gpu": out[b].tolist(), "reference": [int(x) for x in want.view(np.int64)]}), flush=True)
            full = d.lookup_batch(pos[b], full=True)
            print("gpu full records:", full.cpu().tolist(), flush=True)
    print("total mismatches", bad_total)
    d.close()


if __name__ == "__main__":
    main()
