#!/usr/bin/env python
"""BASELINE.json configs[4]: k=31 human-scale index, a large query batch sharded across the GPUs of
one box, NCCL gather of the ids to rank 0 (SURVEY.md 8d cfg 5 / 8e).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port 29517 tools/cfg5_sharded.py [--queries-per-rank 1250000000] [--batch 125000000]

Every rank holds a replica of the index (synthetic 2.5e9-k-mer k=31 m=21 index, built once by rank 0
with the unmodified reference builder: index construction is out of scope) and its own contiguous
shard of the global batch, generated on the device batch by batch: even positions = positive
k-mers (access(uniform id), every other one reverse-complemented), odd positions = uniform random
k-mers (negative with probability 1 - 1e-9).  Each batch goes through sshash_b200.sharded.ShardedLookup:
in "peer" mode (default on NCCL) every rank's lookup kernel stores its ids straight into rank 0's
gathered vector through NVLink peer stores (symmetric memory); in "p2p" mode ids are written locally
and gathered with chunked NCCL send/recv.  Checks: positives return the sampled ids; rank 0 verifies the gathered vector
against per-rank checksums (global query order = rank order); the first 1e6 queries of rank 0 are
compared with the reference CPU dictionary (oracle/_ref) when it is present.
Timing: CUDA events around every ShardedLookup.lookup call, summed, max over ranks.
"""
import argparse
import datetime
import json
import os
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--strings", type=int, default=2500000)
    ap.add_argument("--length", type=int, default=1030)
    ap.add_argument("-k", type=int, default=31)
    ap.add_argument("-m", type=int, default=21)
    ap.add_argument("--index", default=None)
    ap.add_argument("--workdir", default=None)
    ap.add_argument("--queries-per-rank", type=int, default=1_250_000_000)
    ap.add_argument("--batch", type=int, default=125_000_000)
    ap.add_argument("--oracle-sample", type=int, default=1_000_000)
    ap.add_argument("--chunk", type=int, default=1 << 23, help="p2p mode: queries per lookup launch / per send of the gather")
    ap.add_argument("--mode", default="auto", choices=["auto", "peer", "copy", "staged", "p2p"],
                    help="gather: ids stored straight into rank 0's vector over NVLink (peer) or NCCL send/recv (p2p)")
    ap.add_argument("--ids32", action="store_true", help="gather 32-bit ids (sshash_gpu_lookup_batch_u32): half the bytes into rank 0")
    a = ap.parse_args()

    import torch
    import torch.distributed as dist
    import sshash_b200
    from bench import rc_packed_torch
    from sshash_b200.sharded import ShardedLookup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(minutes=30))

    wd = a.workdir or os.path.join(tempfile.gettempdir(), "sshash_cfg5")
    idx = a.index or os.path.join(wd, "synth_%d_%d_k%d_m%d.sshash" % (a.strings, a.length, a.k, a.m))
    build_s = 0.0
    if rank == 0 and not os.path.exists(idx):
        import make_synth_index as msi
        from oracle import ref
        os.makedirs(wd, exist_ok=True)
        t0 = time.time()
        fa = idx + ".fa"
        msi.write_fasta(fa, a.strings, a.length, 42)
        ref.build(fa, a.k, a.m, idx + ".tmp", threads=len(os.sched_getaffinity(0)), tmp_dir=wd, max_k=31)
        os.remove(fa)
        os.rename(idx + ".tmp", idx)
        build_s = time.time() - t0
    if world > 1:
        dist.barrier()
    t0 = time.time()
    d = sshash_b200.Dictionary(idx, device=local)
    open_s = time.time() - t0
    k, nk = d.k(), d.num_kmers()
    assert d.words == 1, "cfg5 is a k <= 31 configuration"

    n_batches = max(1, a.queries_per_rank // a.batch)
    B = a.batch
    sl = ShardedLookup.for_dictionary(d, chunk_queries=a.chunk, mode=a.mode, ids32=a.ids32) if world > 1 else None
    scratch = torch.empty(B, dtype=torch.int64, device=dev)
    gen = torch.Generator(device=dev).manual_seed(1000 + rank)
    total_ms = 0.0
    lookup_only_ms = 0.0
    per_batch_ms = []
    found_neg = 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    oracle_checked = 0
    rc_twins = 0
    for b in range(n_batches + 1):                      # batch 0 is the warm-up
        ids = torch.randint(0, nk, (B // 2,), generator=gen, device=dev, dtype=torch.int64)
        q = torch.empty(B, dtype=torch.int64, device=dev)
        pos = d.access_batch(ids)
        pos[1::2] = rc_packed_torch(pos[1::2], k)
        q[0::2] = pos
        q[1::2] = torch.randint(0, 2 ** (2 * k), (B - B // 2,), generator=gen, device=dev, dtype=torch.int64)
        del pos
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        e0.record()
        if sl is not None:
            out, gathered = sl.lookup(q, dst=0, sizes=[B] * world)
        else:
            out, gathered = d.lookup_batch(q), None
        e1.record()
        torch.cuda.synchronize()
        if out.dtype != torch.int64:                      # 32-bit ids: widen for the checks (UINT32_MAX -> -1)
            widen = lambda t: torch.where(t == -1, torch.full((), -1, dtype=torch.int64, device=t.device), t.to(torch.int64) & 0xFFFFFFFF)
            out = widen(out)
            if gathered is not None:
                gathered = widen(gathered)
        # positives return the sampled ids -- except where the (regular) index holds a k-mer AND its
        # reverse complement as two entries (a 32-base reverse palindrome in the text: ~0.6 expected
        # in 2.5e9 random bases); those few must then agree with the reference CPU dictionary
        bad = (out[0::2] != ids).nonzero().flatten()
        if bad.numel():
            from oracle import ref
            assert bad.numel() <= 16 and ref.available(31), "positive queries must return the sampled ids"
            rd = ref.RefDictionary(idx, max_k=31)
            want = rd.lookup(q[0::2][bad].cpu().numpy().view(np.uint64))
            rd.close()
            assert (out[0::2][bad].cpu().numpy().view(np.uint64) == want).all(), "ids differ from the reference CPU dictionary"
            rc_twins += int(bad.numel())
        if b == 0:
            if rank == 0 and a.oracle_sample:
                from oracle import ref
                if ref.available(31):
                    m = min(a.oracle_sample, B)
                    rd = ref.RefDictionary(idx, max_k=31)
                    want = rd.lookup(q[:m].cpu().numpy().view(np.uint64), threads=len(os.sched_getaffinity(0)))
                    rd.close()
                    assert (out[:m].cpu().numpy().view(np.uint64) == want).all(), "ids differ from the reference CPU dictionary"
                    oracle_checked = m
            continue
        total_ms += e0.elapsed_time(e1)
        per_batch_ms.append(round(e0.elapsed_time(e1), 3))
        found_neg += int((out[1::2] != -1).sum())
        # gathered order: slice r of the gathered vector must be rank r's ids (compare checksums)
        if world > 1:
            mine = torch.stack([out.sum(), (out * torch.arange(1, B + 1, device=dev)).sum()])
            sums = [torch.zeros(2, dtype=torch.int64, device=dev) for _ in range(world)]
            dist.all_gather(sums, mine)
            if rank == 0:
                w = torch.arange(1, B + 1, device=dev)
                for r in range(world):
                    sl_r = gathered[r * B:(r + 1) * B]
                    assert torch.equal(torch.stack([sl_r.sum(), (sl_r * w).sum()]), sums[r]), "gathered ids out of order"
        # the same batch without the gather, for reference
        e0.record()
        d.lookup_batch(q, out=scratch)
        e1.record()
        torch.cuda.synchronize()
        lookup_only_ms += e0.elapsed_time(e1)
        del gathered
    t = torch.tensor([total_ms, lookup_only_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms, lookup_only_ms = float(t[0]), float(t[1])
    nq = world * n_batches * B
    if rank == 0:
        print(json.dumps({
            "config": "cfg5: k=%d m=%d synthetic human-scale index, query batch sharded across %d GPU(s), NCCL gather of ids to rank 0"
                      % (k, a.m, world),
            "index": os.path.basename(idx), "num_kmers": nk, "device_bytes_per_gpu": d.info["device_bytes"],
            "n_gpus": world, "queries_total": nq, "queries_per_rank": n_batches * B, "batch_per_rank": B, "gather_mode": (sl.mode if sl is not None else None), "ids_bits": 32 if a.ids32 else 64,
            "mix": "50 % positive (half of them reverse-complemented), 50 % uniform random (negative)",
            "lookup_plus_gather": {"ms": total_ms, "lookups_per_s": nq / total_ms * 1e3,
                                   "ids_bytes_to_rank0": (world - 1) * n_batches * B * (4 if a.ids32 else 8),
                                   "rank0_ingress_GBps": (world - 1) * n_batches * B * (4 if a.ids32 else 8) / total_ms / 1e6},
            "per_batch_ms_rank0": per_batch_ms,
            "lookup_only": {"ms": lookup_only_ms, "lookups_per_s": nq / lookup_only_ms * 1e3},
            "random_kmers_found": found_neg, "reverse_palindrome_twins_rank0": rc_twins, "checked_vs_reference": oracle_checked,
            "build_s": build_s, "open_s": open_s}), flush=True)
    d.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
