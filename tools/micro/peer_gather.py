#!/usr/bin/env python
"""Experiment: lookup kernels writing their ids straight into rank 0's gathered buffer through
NVLink peer stores (torch symmetric memory), vs lookup + NCCL send/recv gather.

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29519 tools/micro/peer_gather.py
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    import torch.distributed._symmetric_memory as symm_mem
    import sshash_b200
    from bench import rc_packed_torch
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    idx = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "tests", "golden", "se_k31_m13.sshash")
    B = int(sys.argv[2]) if len(sys.argv) > 2 else 100_000_000
    d = sshash_b200.Dictionary(idx, device=local)
    gen = torch.Generator(device=dev).manual_seed(5 + rank)
    ids = torch.randint(0, d.num_kmers(), (B,), generator=gen, device=dev, dtype=torch.int64)
    q = d.access_batch(ids)
    q[1::2] = rc_packed_torch(q[1::2], d.k())
    buf = symm_mem.empty(world * B, dtype=torch.int64, device=dev)
    hdl = symm_mem.rendezvous(buf, dist.group.WORLD)
    remote = hdl.get_buffer(0, (world * B,), torch.int64)
    mine = remote[rank * B:(rank + 1) * B]
    local_out = torch.empty(B, dtype=torch.int64, device=dev)
    res = {}
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    for name in ("local", "peer"):  # see sshash_b200.sharded for the copy-engine variant
        for it in range(6):
            if it == 2:
                torch.cuda.synchronize(); dist.barrier(); ev[0].record()
            if name == "local":
                d.lookup_batch(q, out=local_out)
            else:
                d.lookup_batch(q, out=mine)
                hdl.barrier()
        ev[1].record(); torch.cuda.synchronize()
        t = torch.tensor([ev[0].elapsed_time(ev[1]) / 4], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        res[name + "_ms"] = float(t)
    assert torch.equal(local_out, ids)
    dist.barrier()
    # rank 0 checks every slice against the owner's ids
    all_ids = [torch.empty(B, dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(all_ids, ids)
    if rank == 0:
        for r in range(world):
            assert torch.equal(buf[r * B:(r + 1) * B], all_ids[r]), "peer-written ids differ (rank %d)" % r
        res.update({"world": world, "B": B, "index": os.path.basename(idx),
                    "lookups_per_s_local": world * B / res["local_ms"] * 1e3,
                    "lookups_per_s_gathered_by_peer_stores": world * B / res["peer_ms"] * 1e3})
        print(json.dumps(res), flush=True)
    d.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
