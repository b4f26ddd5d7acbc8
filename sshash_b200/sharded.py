"""Query-sharded lookups over the GPUs of one node.

Lookups are independent and read-only against a static index that fits one GPU, so the path shards
by QUERY: every rank (one process per GPU, torch.distributed) holds a replica of the index, looks
up its own contiguous slice of the batch, and the only communication is the gather of the
resulting ids -- there is no exchange step inside the algorithm.  The gather is chunked and
overlapped with the lookup kernel of the next chunk (NCCL over NVLink on GPUs; gloo in the CPU
tests, where the lookup function is injected).
"""
from __future__ import annotations

from typing import Callable, List, Optional, Tuple


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous slice [lo, hi) of n queries owned by `rank`: sizes differ by at most one and
    concatenating the slices in rank order restores the global query order."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_sizes(n: int, world: int) -> List[int]:
    return [shard_range(n, r, world)[1] - shard_range(n, r, world)[0] for r in range(world)]


class ShardedLookup:
    """lookup_fn(kmers_chunk) -> ids_chunk runs on this rank's device (Dictionary.lookup_batch)."""

    def __init__(self, lookup_fn: Callable, words: int = 1, group=None, chunk_queries: int = 1 << 25):
        import torch.distributed as dist
        self.dist = dist
        self.lookup_fn = lookup_fn
        self.words = words
        self.group = group
        self.chunk = int(chunk_queries)
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)

    @classmethod
    def for_dictionary(cls, dictionary, group=None, chunk_queries: int = 1 << 25):
        return cls(lambda k: dictionary.lookup_batch(k), words=dictionary.words, group=group,
                   chunk_queries=chunk_queries)

    def lookup(self, local_kmers, dst: Optional[int] = 0):
        """Look up this rank's shard.  Returns (local_ids, gathered) where `gathered` is, on rank
        `dst`, the ids of ALL ranks concatenated in rank order (= global query order for shards made
        with shard_range); None elsewhere, or everywhere when dst is None."""
        import torch
        dist = self.dist
        n_local = local_kmers.numel() // self.words
        local_ids = torch.empty(n_local, dtype=torch.int64, device=local_kmers.device)
        if dst is None or self.world == 1:
            for lo in range(0, n_local, self.chunk):
                hi = min(n_local, lo + self.chunk)
                local_ids[lo:hi] = self.lookup_fn(local_kmers[lo * self.words:hi * self.words])
            return local_ids, (local_ids if (dst is not None and self.world == 1) else None)
        # sizes of all shards (ragged shards allowed)
        sizes_t = [torch.zeros(1, dtype=torch.int64, device=local_kmers.device) for _ in range(self.world)]
        dist.all_gather(sizes_t, torch.tensor([n_local], dtype=torch.int64, device=local_kmers.device),
                        group=self.group)
        sizes = [int(s.item()) for s in sizes_t]
        starts = [sum(sizes[:r]) for r in range(self.world)]
        gathered = None
        pending = []
        if self.rank == dst:
            gathered = torch.empty(sum(sizes), dtype=torch.int64, device=local_kmers.device)
            # post the receives for every remote chunk up front; they complete as the senders progress
            for r in range(self.world):
                if r == dst:
                    continue
                for lo in range(0, sizes[r], self.chunk):
                    hi = min(sizes[r], lo + self.chunk)
                    pending.append(dist.irecv(gathered[starts[r] + lo:starts[r] + hi], src=r, group=self.group))
        for lo in range(0, n_local, self.chunk):
            hi = min(n_local, lo + self.chunk)
            ids = self.lookup_fn(local_kmers[lo * self.words:hi * self.words])
            local_ids[lo:hi] = ids
            if self.rank != dst:
                # the send of chunk c overlaps the lookup kernel of chunk c+1 (NCCL runs on its own stream)
                pending.append(dist.isend(local_ids[lo:hi], dst=dst, group=self.group))
        if self.rank == dst:
            gathered[starts[dst]:starts[dst] + n_local] = local_ids
        for p in pending:
            p.wait()
        return local_ids, gathered
