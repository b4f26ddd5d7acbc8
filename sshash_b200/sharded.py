"""Query-sharded lookups over the GPUs of one node.

Lookups are independent and read-only against a static index that fits one GPU, so the path shards
by QUERY: every rank (one process per GPU, torch.distributed) holds a replica of the index, looks
up its own contiguous slice of the batch, and the only communication is the gather of the
resulting ids -- there is no exchange step inside the algorithm.  Two gather modes:

  "peer"  (GPUs of one NVLink/NVSwitch box) the gather is FUSED into the lookup kernel: rank dst owns
          the gathered vector in symmetric memory (torch.distributed._symmetric_memory), every rank
          maps it, and each rank's lookup kernel stores its ids straight into its slice of that
          vector through NVLink peer stores (the C ABI takes any device pointer as `kmer_ids`).  No
          gather kernel, no staging copy, no SMs taken from the lookups; a device-side barrier on
          the stream publishes the result.  Measured at N=2 on cfg2: 4.72 ms per 2x1e8 lookups with
          the ids gathered vs 4.72 ms without (tools/micro/peer_gather.py).
  "copy"  ids written locally chunk by chunk, each finished chunk pushed into rank dst's symmetric
          vector by a device-to-device memcpy on a side stream (copy engines over NVLink) while the
          next chunk is looked up.  For many senders: the 8-byte stores of "peer" mode that come out
          of the reverse-complement queue are scattered, and 7 ranks of them fill rank 0's NVLink
          ingress at ~300 GB/s, whereas the copy engines move full-size packets.
  "staged" the same data movement as "copy", but driven inside the library: rank r hands the C ABI its
          slice of rank dst's vector as the output pointer with peer-inplace OFF, and the library's
          own chunk pipeline (api.cu run_batched: kernel of chunk c into a staging buffer, copy-engine
          push of chunk c-1, three streams) delivers it -- no per-chunk Python / event overhead.
  "p2p"   chunked isend/irecv of locally written ids, overlapped with the next chunk's lookup
          kernel (NCCL on GPUs; gloo in the CPU tests, where the lookup function is injected).  An
          NCCL send kernel needs SMs the persistent lookup CTAs hold, so this mode costs the full
          transfer time on top of the lookups (+2.1 ms per GB at N=2).
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

    def __init__(self, lookup_fn: Callable, words: int = 1, group=None, chunk_queries: int = 1 << 25,
                 lookup_into: Optional[Callable] = None, mode: str = "p2p", ids_dtype=None):
        import torch.distributed as dist
        if mode not in ("p2p", "peer", "copy", "staged"):
            raise ValueError("mode must be 'p2p', 'peer', 'copy' or 'staged'")
        if mode in ("peer", "copy", "staged") and lookup_into is None:
            raise ValueError("mode '%s' needs lookup_into" % mode)
        self.dist = dist
        self.lookup_fn = lookup_fn
        self.lookup_into = lookup_into       # lookup_into(kmers, out): ids written into `out` (any device pointer)
        self.words = words
        self.group = group
        self.chunk = int(chunk_queries)
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        self.mode = mode
        if ids_dtype is None:
            import torch
            ids_dtype = torch.int64
        self.ids_dtype = ids_dtype           # torch.int64 (reference-width ids) or torch.int32 (sshash_gpu_lookup_batch_u32)
        self._symm = None                    # (buffer, handle, capacity, dst)
        self._side = None                    # copy stream of mode "copy"

    @classmethod
    def for_dictionary(cls, dictionary, group=None, chunk_queries: int = 1 << 25, mode: str = "auto", ids32: bool = False):
        """mode "auto": on NCCL (GPUs of one box) peer stores for 2 ranks, the library-staged copy-engine
        pipeline beyond (rank dst's NVLink ingress is the limit there and the scattered stores that leave the
        reverse-complement queue use it badly); p2p on other backends.  Measured on cfg2, 1e8 queries per
        rank, 32-bit ids: N=2 4.90 / 5.09 / 5.83 ms for peer / staged / copy (the Python-driven "copy" pays
        ~60 us of host time per piece); N=8 9.5 / - / 5.9 ms."""
        import torch.distributed as dist
        if mode == "auto":
            if dist.get_backend(group) == "nccl" and dist.get_world_size(group) > 1:
                # beyond two ranks: the library-staged pipeline for light kernels (small indexes: the Python-driven
                # "copy" is host-bound there), "copy" with its larger pieces for heavy ones (2.5e9-k-mer index at
                # N=8: 72.4 vs 68.1 G lookups/s gathered)
                big = getattr(dictionary, "info", {}).get("device_bytes", 0) > (1 << 30)
                mode = "peer" if dist.get_world_size(group) <= 2 else ("copy" if big else "staged")
            else:
                mode = "p2p"
        import torch
        if hasattr(dictionary, "set_peer_inplace"):   # "peer": the lookup kernels store straight into rank dst's peer-mapped vector
            dictionary.set_peer_inplace(mode == "peer")
        if ids32:   # 32-bit ids (dictionaries with < 2^32 - 1 k-mers): half the bytes into rank dst's NVLink ingress
            return cls(lambda k: dictionary.lookup_batch_u32(k), words=dictionary.words, group=group, chunk_queries=chunk_queries,
                       lookup_into=lambda k, out: dictionary.lookup_batch_u32(k, out=out), mode=mode, ids_dtype=torch.int32)
        return cls(lambda k: dictionary.lookup_batch(k), words=dictionary.words, group=group,
                   chunk_queries=chunk_queries, lookup_into=lambda k, out: dictionary.lookup_batch(k, out=out), mode=mode)

    def _gathered_buffer(self, total: int, dst: int, device):
        """Symmetric buffer of >= total ids: (this rank's allocation, view of rank dst's allocation)."""
        import torch
        import torch.distributed._symmetric_memory as symm_mem
        if self._symm is None or self._symm[2] < total or self._symm[3] != dst:
            cap = int(total)
            buf = symm_mem.empty(cap, dtype=self.ids_dtype, device=device)
            hdl = symm_mem.rendezvous(buf, self.group if self.group is not None else self.dist.group.WORLD)
            self._symm = (buf, hdl, cap, dst)
        buf, hdl, cap, _ = self._symm
        return buf, hdl, hdl.get_buffer(dst, (cap,), self.ids_dtype)

    def _lookup_peer(self, local_kmers, dst: int, sizes, starts):
        n_local = sizes[self.rank]
        buf, hdl, remote = self._gathered_buffer(sum(sizes), dst, local_kmers.device)
        mine = remote[starts[self.rank]:starts[self.rank] + n_local]
        hdl.barrier()                                    # rank dst has consumed the previous result (stream order)
        if self.mode == "staged":
            import torch
            torch.cuda.current_stream(local_kmers.device).synchronize()   # the library pipelines on its own streams
        if n_local:
            self.lookup_into(local_kmers, mine)          # ids go over NVLink into rank dst's vector
        hdl.barrier()                                    # on the current stream: every rank's stores are done
        return mine, (buf[:sum(sizes)] if self.rank == dst else None)

    def _lookup_copy(self, local_kmers, dst: int, sizes, starts):
        """ids written locally chunk by chunk; each finished chunk is pushed into rank dst's vector by a
        device-to-device memcpy on a side stream (copy engines over NVLink: no SMs, full-size packets)
        while the next chunk's lookup kernel runs."""
        import torch
        n_local = sizes[self.rank]
        buf, hdl, remote = self._gathered_buffer(sum(sizes), dst, local_kmers.device)
        mine = remote[starts[self.rank]:starts[self.rank] + n_local]
        hdl.barrier()
        if self.rank == dst:                             # the root's slice is local memory: look up in place
            if n_local:
                self.lookup_into(local_kmers, mine)
            local_ids = mine
        else:
            local_ids = torch.empty(n_local, dtype=self.ids_dtype, device=local_kmers.device)
            if self._side is None:
                self._side = torch.cuda.Stream(device=local_kmers.device)
            main = torch.cuda.current_stream(local_kmers.device)
            for lo in range(0, n_local, self.chunk):
                hi = min(n_local, lo + self.chunk)
                self.lookup_into(local_kmers[lo * self.words:hi * self.words], local_ids[lo:hi])
                ev = torch.cuda.Event()
                ev.record(main)
                self._side.wait_event(ev)
                with torch.cuda.stream(self._side):
                    mine[lo:hi].copy_(local_ids[lo:hi], non_blocking=True)
            main.wait_stream(self._side)
        hdl.barrier()
        return local_ids, (buf[:sum(sizes)] if self.rank == dst else None)

    def lookup(self, local_kmers, dst: Optional[int] = 0, sizes: Optional[List[int]] = None):
        """Look up this rank's shard.  `sizes`: the number of queries of EVERY rank when the caller knows them
        (e.g. equal shards) -- saves the all_gather + host synchronisation that finds them out.  Returns (local_ids, gathered) where `gathered` is, on rank
        `dst`, the ids of ALL ranks concatenated in rank order (= global query order for shards made
        with shard_range); None elsewhere, or everywhere when dst is None.  In "peer" mode local_ids
        is this rank's slice of rank dst's vector (peer-mapped memory) and both results are valid
        until the next call."""
        import torch
        dist = self.dist
        n_local = local_kmers.numel() // self.words
        if dst is None or self.world == 1 or self.mode == "p2p":
            local_ids = torch.empty(n_local, dtype=self.ids_dtype, device=local_kmers.device)
        if dst is None or self.world == 1:
            for lo in range(0, n_local, self.chunk):
                hi = min(n_local, lo + self.chunk)
                local_ids[lo:hi] = self.lookup_fn(local_kmers[lo * self.words:hi * self.words])
            return local_ids, (local_ids if (dst is not None and self.world == 1) else None)
        # sizes of all shards (ragged shards allowed)
        if sizes is None:
            sizes_t = [torch.zeros(1, dtype=torch.int64, device=local_kmers.device) for _ in range(self.world)]
            dist.all_gather(sizes_t, torch.tensor([n_local], dtype=torch.int64, device=local_kmers.device),
                            group=self.group)
            sizes = [int(s.item()) for s in sizes_t]
        elif len(sizes) != self.world or sizes[self.rank] != n_local:
            raise ValueError("sizes must list every rank's shard size")
        starts = [sum(sizes[:r]) for r in range(self.world)]
        if self.mode in ("peer", "staged"):      # same call; the dictionary's peer-inplace switch decides who moves the ids
            return self._lookup_peer(local_kmers, dst, sizes, starts)
        if self.mode == "copy":
            return self._lookup_copy(local_kmers, dst, sizes, starts)
        gathered = None
        pending = []
        if self.rank == dst:
            gathered = torch.empty(sum(sizes), dtype=self.ids_dtype, device=local_kmers.device)
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
