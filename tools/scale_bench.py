#!/usr/bin/env python
"""HBM-resident scale test of the lookup kernel: a synthetic ~5e8-k-mer index
built on the box by the reference builder, 1e8 positive / negative queries, device-resident timing.

    python tools/scale_bench.py [--strings 500000] [--length 1030] [-k 31] [-m 17] [--queries 100000000]
"""
import argparse
import json
import os
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))


def time_lookup(d, kmers, out, steps=5, warmup=3):
    import torch
    for _ in range(warmup):
        d.lookup_batch(kmers, out=out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        d.lookup_batch(kmers, out=out)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def run(strings, length, k, m, canonical, queries, workdir, keep=False, index=None):
    import torch
    import sshash_b200
    from bench import rc_packed_torch, rc_packed_torch2
    import make_synth_index as msi
    from oracle import ref
    idx = os.path.join(workdir, "synth_%d_%d_k%d_m%d%s.sshash" % (strings, length, k, m, "_c" if canonical else ""))
    if index:
        idx, keep = index, True
    t0 = time.time()
    if not os.path.exists(idx):
        fa = idx + ".fa"
        msi.write_fasta(fa, strings, length, 42)
        ref.build(fa, k, m, idx, canonical=canonical, threads=len(os.sched_getaffinity(0)), tmp_dir=workdir,
                  max_k=31 if k <= 31 else 63)
        os.remove(fa)
    build_s = time.time() - t0
    t0 = time.time()
    d = sshash_b200.Dictionary(idx)
    open_s = time.time() - t0
    k = d.k()
    dev = torch.device("cuda", 0)
    gen = torch.Generator(device=dev).manual_seed(7)
    n = queries
    ids = torch.randint(0, d.num_kmers(), (n,), generator=gen, device=dev, dtype=torch.int64)
    fwd = d.access_batch(ids)
    out = torch.empty(n, dtype=torch.int64, device=dev)
    res = {"index": os.path.basename(idx), "num_kmers": d.num_kmers(), "index_bytes": d.info["index_file_bytes"],
           "device_bytes": d.info["device_bytes"], "mphf_partitions": d.info["mphf_partitions"],
           "num_minimizers": d.info["num_minimizers"], "build_s": build_s, "open_s": open_s, "queries": n}
    ms = time_lookup(d, fwd, out)
    assert torch.equal(out, ids)
    res["positive_forward"] = {"ms": ms, "lookups_per_s": n / ms * 1e3}
    mix = fwd.clone()
    if d.words == 1:
        mix[1::2] = rc_packed_torch(mix[1::2], k)
    else:
        lo, hi = rc_packed_torch2(mix[1::2, 0], mix[1::2, 1], k)
        mix[1::2, 0] = lo
        mix[1::2, 1] = hi
    ms = time_lookup(d, mix, out)
    assert torch.equal(out, ids)
    res["positive_50rc"] = {"ms": ms, "lookups_per_s": n / ms * 1e3}
    if d.words == 1:
        neg = torch.randint(0, 2 ** (2 * k), (n,), generator=gen, device=dev, dtype=torch.int64)
    else:
        neg = torch.randint(0, 2 ** 62, (n, 2), generator=gen, device=dev, dtype=torch.int64)
        neg[:, 1] &= (1 << (2 * k - 64)) - 1
    ms = time_lookup(d, neg, out)
    res["negative"] = {"ms": ms, "lookups_per_s": n / ms * 1e3, "found": int((out != -1).sum())}
    d.close()
    if not keep:
        os.remove(idx)
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--strings", type=int, default=500000)
    ap.add_argument("--length", type=int, default=1030)
    ap.add_argument("-k", type=int, default=31)
    ap.add_argument("-m", type=int, default=17)
    ap.add_argument("--canonical", action="store_true")
    ap.add_argument("--queries", type=int, default=100_000_000)
    ap.add_argument("--workdir", default=None)
    ap.add_argument("--keep", action="store_true")
    ap.add_argument("--index", default=None, help="use this existing index instead of building a synthetic one")
    args = ap.parse_args()
    wd = args.workdir or tempfile.mkdtemp(prefix="sshash_scale_")
    os.makedirs(wd, exist_ok=True)
    print(json.dumps(run(args.strings, args.length, args.k, args.m, args.canonical, args.queries, wd, args.keep, args.index)))


if __name__ == "__main__":
    main()
