"""CPU tests of the N>1 path: world_size-2 (and 3) gloo process groups exercise the query sharding
and the chunked gather of ids with an injected lookup function (the product's lookup itself only
runs on a GPU)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from sshash_b200.sharded import ShardedLookup, shard_range, shard_sizes


def test_shard_range_properties():
    for n in (0, 1, 7, 100, 12345):
        for world in (1, 2, 3, 8):
            sizes = shard_sizes(n, world)
            assert sum(sizes) == n and max(sizes) - min(sizes) <= 1
            pos = 0
            for r in range(world):
                lo, hi = shard_range(n, r, world)
                assert lo == pos and hi - lo == sizes[r]
                pos = hi


def test_modes_that_write_into_the_root_vector_need_lookup_into():
    for mode in ("peer", "copy", "staged"):
        with pytest.raises(ValueError, match="needs lookup_into"):
            ShardedLookup(lambda k: k, mode=mode)
    with pytest.raises(ValueError, match="mode must be"):
        ShardedLookup(lambda k: k, mode="nccl")


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n, chunk, words, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(5)
        kmers = torch.from_numpy(rng.integers(0, 2**62, n * words).astype(np.int64))
        lo, hi = shard_range(n, rank, world)
        local = kmers[lo * words:hi * words].clone()

        def fake_lookup(chunk_kmers):      # deterministic stand-in for Dictionary.lookup_batch
            return chunk_kmers.view(-1, words)[:, 0] * 3 + 1

        sl = ShardedLookup(fake_lookup, words=words, chunk_queries=chunk)
        local_ids, gathered = sl.lookup(local, dst=0)
        expect = kmers.view(-1, words)[:, 0] * 3 + 1
        ok = torch.equal(local_ids, expect[lo:hi])
        if rank == 0:
            ok = ok and gathered is not None and torch.equal(gathered, expect)
        else:
            ok = ok and gathered is None
        _, none = sl.lookup(local, dst=None)
        ok = ok and none is None
        # caller-provided shard sizes skip the all_gather; 32-bit ids travel as int32
        sl32 = ShardedLookup(lambda c: (c.view(-1, words)[:, 0] % 1000).to(torch.int32), words=words, chunk_queries=chunk,
                             ids_dtype=torch.int32)
        ids32, g32 = sl32.lookup(local, dst=0, sizes=shard_sizes(n, world))
        exp32 = (kmers.view(-1, words)[:, 0] % 1000).to(torch.int32)
        ok = ok and ids32.dtype == torch.int32 and torch.equal(ids32, exp32[lo:hi])
        if rank == 0:
            ok = ok and torch.equal(g32, exp32)
        try:
            sl32.lookup(local, dst=0, sizes=[n] * (world + 1))
            ok = False
        except ValueError:
            pass
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n,chunk,words", [(2, 1000, 128, 1), (2, 1001, 1 << 20, 2), (3, 50, 7, 1), (2, 1, 4, 1)])
def test_sharded_lookup_gloo(world, n, chunk, words):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, chunk, words, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=120)
    res = sorted(q.get(timeout=5) for _ in range(world))
    assert res == [(r, True) for r in range(world)]
