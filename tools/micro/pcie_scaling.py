#!/usr/bin/env python
"""How does host<->device traffic scale when 1, 2, 4, 8 GPUs of the box move data at the same time?

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 \
        tools/micro/pcie_scaling.py [--mb 800]

One process per GPU (like bench.py).  For every active set {first 1, 2, 4, ... ranks} the active
ranks move `mb` MB host->device and `mb` MB device->host CONCURRENTLY (pinned memory, two streams,
32 MB pieces, no kernels), three timed repetitions bracketed by barriers; the others idle.  Reported:
per-direction GB/s per GPU (min over the active ranks) and the aggregate.  Run twice: as launched, and
with every process bound to the CPUs next to its GPU (NVML affinity) before the pinned buffers are
allocated -- the A/B for "is it the NUMA placement of the pinned buffers".  This is the ceiling of the
e2e (host-buffer) leg of bench.py at N GPUs: 16 bytes cross PCIe per lookup.
"""
import argparse
import json
import os
import time


def bind_to_gpu_numa(local_rank):
    try:
        import pynvml
        import torch
        pynvml.nvmlInit()
        pr = torch.cuda.get_device_properties(local_rank)
        bus = "%08x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
        pynvml.nvmlDeviceSetCpuAffinity(pynvml.nvmlDeviceGetHandleByPciBusId(bus.encode()))
        return sorted(os.sched_getaffinity(0))
    except Exception as e:   # noqa: BLE001
        return "failed: %s" % e


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mb", type=int, default=800)
    a = ap.parse_args()
    import torch
    import torch.distributed as dist
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    n = a.mb * (1 << 20) // 8
    piece = (32 << 20) // 8
    d_in = torch.empty(n, dtype=torch.int64, device=dev)
    d_out = torch.zeros(n, dtype=torch.int64, device=dev)
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    out = {"world": world, "mb_per_direction": a.mb, "cpus_visible": len(os.sched_getaffinity(0))}
    for placement in ("as_launched", "bound_to_gpu_numa_node"):
        aff = None
        if placement == "bound_to_gpu_numa_node":
            aff = bind_to_gpu_numa(local)
        h_in = torch.empty(n, dtype=torch.int64, pin_memory=True)
        h_in.fill_(1)
        h_out = torch.empty(n, dtype=torch.int64, pin_memory=True)
        h_out.fill_(0)
        res = {}
        active = 1
        while active <= world:
            for mode in ("h2d", "d2h", "duplex"):
                secs = 0.0
                for it in range(4):
                    if world > 1:
                        dist.barrier()
                    torch.cuda.synchronize()
                    t0 = time.perf_counter()
                    if rank < active:
                        for lo in range(0, n, piece):
                            hi = min(n, lo + piece)
                            if mode != "d2h":
                                with torch.cuda.stream(s1):
                                    d_in[lo:hi].copy_(h_in[lo:hi], non_blocking=True)
                            if mode != "h2d":
                                with torch.cuda.stream(s2):
                                    h_out[lo:hi].copy_(d_out[lo:hi], non_blocking=True)
                        torch.cuda.synchronize()
                    dt = time.perf_counter() - t0
                    if it:
                        secs += dt / 3
                t = torch.tensor([secs if rank < active else 0.0], device=dev, dtype=torch.float64)
                if world > 1:
                    dist.all_reduce(t, op=dist.ReduceOp.MAX)
                slowest = float(t.item())
                res["%d_gpus_%s" % (active, mode)] = {"GBps_per_direction_per_gpu": a.mb * (1 << 20) / slowest / 1e9,
                                                      "aggregate_GBps_per_direction": active * a.mb * (1 << 20) / slowest / 1e9}
                if mode == "duplex":
                    res["%d_gpus_%s" % (active, mode)]["lookups_per_s_ceiling_at_16B"] = active * n / slowest
            active *= 2
        out[placement] = res
        if aff is not None:
            out["affinity_rank%d" % rank] = aff if isinstance(aff, str) else "%d cpus: %s.." % (len(aff), aff[:4])
        del h_in, h_out
    if rank == 0:
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
