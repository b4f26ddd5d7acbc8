#!/usr/bin/env python
"""A short lookup run to put under ncu: opens an index, makes 1e8 device-resident queries of one
kind and launches the lookup path three times.

    ncu --set full --clock-control none --import-source on -k regex:lookup -s 2 -c 1 -o gpurun_out/x \
        python tools/ncu_target.py --index /tmp/ix/synth_....sshash --mode mix [--sorted]
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--index", required=True)
    ap.add_argument("--mode", default="mix", choices=["fwd", "mix", "neg"])
    ap.add_argument("--queries", type=int, default=100_000_000)
    ap.add_argument("--sorted", action="store_true")
    ap.add_argument("--max-k", type=int, default=0)
    ap.add_argument("--launches", type=int, default=3)
    a = ap.parse_args()
    import torch
    import sshash_b200
    from bench import rc_packed_torch, rc_packed_torch2
    d = sshash_b200.Dictionary(a.index, max_k=a.max_k)
    dev = torch.device("cuda", 0)
    n, k = a.queries, d.k()
    gen = torch.Generator(device=dev).manual_seed(7)
    ids = torch.randint(0, d.num_kmers(), (n,), generator=gen, device=dev, dtype=torch.int64)
    q = d.access_batch(ids)
    if a.mode == "mix":
        if d.words == 1:
            q[1::2] = rc_packed_torch(q[1::2], k)
        else:
            lo, hi = rc_packed_torch2(q[1::2, 0], q[1::2, 1], k)
            q[1::2, 0], q[1::2, 1] = lo, hi
    elif a.mode == "neg":
        if d.words == 1:
            q = torch.randint(0, 2 ** (2 * k), (n,), generator=gen, device=dev, dtype=torch.int64)
        else:
            q = torch.randint(0, 2 ** 62, (n, 2), generator=gen, device=dev, dtype=torch.int64)
            q[:, 1] &= (1 << (2 * k - 64)) - 1
    if a.sorted:
        parts = torch.empty(n, dtype=torch.int32, device=dev)
        sshash_b200._lib.check(d._lib.sshash_gpu_minimizer_partition_batch(d._h, q.data_ptr(), n, parts.data_ptr(), None))
        torch.cuda.synchronize()
        q = q[torch.sort(parts, stable=True)[1]].contiguous()
    out = torch.empty(n, dtype=torch.int64, device=dev)
    for _ in range(a.launches):
        d.lookup_batch(q, out=out)
    torch.cuda.synchronize()
    d.close()


if __name__ == "__main__":
    main()
