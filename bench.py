#!/usr/bin/env python
"""bench.py -- k-mer lookups/s of the batched Lookup hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Workload (BASELINE.json configs[1]): the bundled S. enterica k=31 m=13 index
(tests/golden/se_k31_m13.sshash, built by the reference builder), 1e8 uniform-random POSITIVE
k-mers per GPU, every other one reverse-complemented (the reference's own bench protocol,
tools/perf.hpp:41-51).  One step = one pass of the lookup path over the whole 1e8-query batch.

  value : whole-job lookups/s with queries and ids resident in HBM (CUDA events, max over ranks)
  e2e   : the same batch through the C-ABI call with HOST (pinned) buffers: H2D of the packed
          k-mers and D2H of the ids inside the timed region
  roofline : algorithmic bytes per lookup (256 B for a regular index at 50% RC)
             x lookups/s of the lookup kernel vs the measured HBM peak
  cpu_baseline : the reference's own dictionary::lookup on the box's host cores, bounded sample

--impl reference times the reference's CPU implementation (oracle/_ref when built, else the C
oracle port) with all host threads on a bounded sample of the same workload.
With N > 1 (torchrun) every rank holds a replica of the index and its own 1e8-query shard (weak
scaling, no data-path collective); the gather of the ids to rank 0 is timed separately (`gather`):
fused into the lookup kernels as NVLink peer stores, and as NCCL send/recv for comparison.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

INDEX = os.path.join(ROOT, "tests", "golden", "se_k31_m13.sshash")
WORKLOAD = "cfg2: S.enterica k31 m13 index (4.79e6 k-mers), 1e8 uniform-random positive k-mers/GPU, 50% reverse-complemented"
QUERIES_PER_GPU = 100_000_000
B_ALG = 256.0          # k<=31 positive, regular index, 50 % RC mix
CPU_SAMPLE = 20_000_000


def s64(x: int) -> int:
    x &= (1 << 64) - 1
    return x - (1 << 64) if x >> 63 else x


def rc_packed_torch(x, k: int):
    """reverse complement of packed k<=31-mers held in int64 tensors (kmer.hpp:141-165)."""
    c = x ^ s64(0xAAAAAAAAAAAAAAAA)
    for sh, m in ((2, 0x3333333333333333), (4, 0x0F0F0F0F0F0F0F0F), (8, 0x00FF00FF00FF00FF),
                  (16, 0x0000FFFF0000FFFF), (32, 0x00000000FFFFFFFF)):
        c = ((c >> sh) & s64(m)) | ((c & s64(m)) << sh)
    s = 64 - 2 * k
    return (c >> s) & s64((1 << (64 - s)) - 1)


def _rev2bit64(c):
    for sh, m in ((2, 0x3333333333333333), (4, 0x0F0F0F0F0F0F0F0F), (8, 0x00FF00FF00FF00FF),
                  (16, 0x0000FFFF0000FFFF), (32, 0x00000000FFFFFFFF)):
        c = ((c >> sh) & s64(m)) | ((c & s64(m)) << sh)
    return c


def rc_packed_torch2(lo, hi, k: int):
    """reverse complement of packed k<=63-mers held as two int64 words (kmer.hpp:159-165)."""
    a = _rev2bit64(lo ^ s64(0xAAAAAAAAAAAAAAAA))     # becomes the HIGH word before the final shift
    b = _rev2bit64(hi ^ s64(0xAAAAAAAAAAAAAAAA))     # becomes the LOW word
    s = 128 - 2 * k
    if s >= 64:
        t = s - 64
        return ((a >> t) & s64((1 << (64 - t)) - 1)) if t else a, a * 0
    lsr = lambda v, n: (v >> n) & s64((1 << (64 - n)) - 1)
    return lsr(b, s) | (a << (64 - s)), lsr(a, s)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.rows = []
        self.proc = None
        self.gpu = gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) < 9:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def recorded_traffic(n_queries: int):
    """dram bytes per launch of the lookup kernel, from the committed ncu --set full capture
    (profiles/traffic.json holds the per-query figure; one launch = n_queries queries)."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(p):
        try:
            return json.load(open(p))["cfg2_lookup_kernel_dram_bytes_per_query"] * n_queries
        except Exception:
            return None
    return None


def scale_workload():
    """Secondary, HBM-resident workload: a synthetic 5e8-k-mer k=31 m=17 index
    built on this box by the unmodified reference builder (oracle/_ref; index construction is out of
    scope), 1e8 positive / negative queries, device-resident.  Reported beside the headline line."""
    try:
        import tempfile
        sys.path.insert(0, os.path.join(ROOT, "tools"))
        import scale_bench
        wd = tempfile.mkdtemp(prefix="sshash_scale_")
        r = scale_bench.run(500000, 1030, 31, 17, False, 100_000_000, wd)
        peak, _ = measured_peak()
        for key, b_alg in (("positive_forward", 208.0), ("positive_50rc", 256.0), ("negative", 208.0)):
            r[key]["roofline_frac"] = b_alg * r[key]["lookups_per_s"] / 1e9 / peak
            r[key]["algorithmic_bytes_per_lookup"] = b_alg
        r["workload"] = "T5e8: synthetic 5e5 strings x 1030 bases, k=31 m=17 (5e8 k-mers, HBM-resident), 1e8 queries"
        return r
    except Exception as e:  # no reference builder on this box, out of disk, ...
        return {"unavailable": "%s: %s" % (type(e).__name__, e)}


def streaming_workload():
    """BASELINE.json configs[2]: streaming membership over 1e6 synthetic 150-bp reads (50 % hit) on
    the cfg-1 index; windows/s device-resident, through the C ABI with host buffers, and through
    streaming_query_from_file on a 321 MB FASTQ file (records parsed on the GPU), checked against
    the C oracle on the first 20000 reads and against the unmodified reference on a 1e5-read file."""
    try:
        sys.path.insert(0, os.path.join(ROOT, "tools"))
        import stream_bench
        r = stream_bench.run(INDEX, 1_000_000, files=True)
        r["workload"] = "cfg3: 1e6 synthetic 150-bp reads (1.2e8 windows), 50 % of reads from the index, cfg-1 index"
        return r
    except Exception as e:
        return {"unavailable": "%s: %s" % (type(e).__name__, e)}


def host_threads() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def cpu_reference(kmers_host: np.ndarray, threads: int):
    """(lookups/s, kind) of the reference's CPU path over `kmers_host` with `threads` threads."""
    from oracle import ref
    if ref.available(31):
        d = ref.RefDictionary(INDEX, max_k=31)
        d.time_lookup(kmers_host[:200000], threads=threads)  # warm the index
        secs = d.time_lookup(kmers_host, threads=threads)
        return kmers_host.size / secs, "reference"
    from oracle import port  # scalar C port: single thread
    o = port.OracleDictionary(INDEX)
    t0 = time.perf_counter()
    o.lookup(kmers_host)
    return kmers_host.size / (time.perf_counter() - t0), "port"


def make_queries_numpy(n: int, seed: int) -> np.ndarray:
    """The same workload without a GPU (reference arm): positives via the oracle's access()."""
    from oracle import port
    o = port.OracleDictionary(INDEX)
    rng = np.random.default_rng(seed)
    ids = rng.integers(0, o.num_kmers, n).astype(np.uint64)
    k = o.access(ids)
    import torch
    t = torch.from_numpy(k.view(np.int64))
    t[1::2] = rc_packed_torch(t[1::2], o.k)
    return k


def run_reference(args, rank: int, world: int):
    if rank != 0:
        return
    threads = host_threads()
    q = make_queries_numpy(CPU_SAMPLE, 42)
    vals = []
    for _ in range(args.warmup):
        cpu_reference(q[: CPU_SAMPLE // 10], threads)
    kind = "reference"
    t0 = time.perf_counter()
    for _ in range(args.steps):
        v, kind = cpu_reference(q, threads)
        vals.append(v)
    wall = time.perf_counter() - t0
    value = float(np.mean(vals))
    line = {
        "impl": "reference", "metric": "k-mer lookups/sec", "value": value, "unit": "lookups/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * wall / max(args.steps, 1),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "sample": "%d queries per step" % CPU_SAMPLE},
        "cpu_baseline": {"value": value, "unit": "lookups/s", "cores": threads, "kind": kind,
                         "sample": "%d of the 1e8 queries per step, %d host threads" % (CPU_SAMPLE, threads)},
        "e2e": {"value": value, "unit": "lookups/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--queries", type=int, default=QUERIES_PER_GPU, help="queries per GPU per step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-scale", action="store_true", help="skip the HBM-resident 5e8-k-mer secondary workload")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    import sshash_b200

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (the lookup path has no CPU fallback)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        # one process per GPU: run on the CPUs next to this GPU so that the pinned host buffers of the
        # e2e leg are allocated on its NUMA node (at N=1 the process keeps all cores for the CPU baseline)
        try:
            import pynvml
            pynvml.nvmlInit()
            pr = torch.cuda.get_device_properties(local_rank)
            bus = "%08x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)   # CUDA order may differ from NVML's
            pynvml.nvmlDeviceSetCpuAffinity(pynvml.nvmlDeviceGetHandleByPciBusId(bus.encode()))
        except Exception:
            pass
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("TORCH_NCCL_SHOW_EAGER_INIT_P2P_SERIALIZATION_WARNING", "false")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)
    n = args.queries

    d = sshash_b200.Dictionary(INDEX, device=local_rank)
    k = d.k()
    # ---- synthetic queries, generated on the device ---------------------------------------------
    gen = torch.Generator(device=dev).manual_seed(42 + rank)
    ids = torch.randint(0, d.num_kmers(), (n,), generator=gen, device=dev, dtype=torch.int64)
    kmers = d.access_batch(ids)
    kmers[1::2] = rc_packed_torch(kmers[1::2], k)
    out = torch.empty(n, dtype=torch.int64, device=dev)
    stream = torch.cuda.current_stream().cuda_stream
    torch.cuda.synchronize()

    def step_device():
        d.lookup_batch(kmers, out=out, stream=stream)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing -------------------------------------------------------------------
    for _ in range(args.warmup):
        step_device()
    torch.cuda.synchronize()
    assert torch.equal(out, ids), "lookup ids differ from the sampled ids"   # positives are self-checking
    launches0 = sshash_b200.launch_count()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    barrier()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    ev[0].record()
    for i in range(args.steps):
        step_device()
        ev[i + 1].record()
    barrier()
    launches = sshash_b200.launch_count() - launches0
    total_ms = ev[0].elapsed_time(ev[-1])
    kernel_ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(args.steps)]
    t = torch.tensor([total_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = float(t.item())
    value = world * n * args.steps / (total_ms * 1e-3)

    # ---- end-to-end through the C ABI with host (pinned) buffers ----------------------------------
    h_in = torch.empty(n, dtype=torch.int64, pin_memory=True)
    h_in.copy_(kmers)
    h_out = torch.empty(n, dtype=torch.int64, pin_memory=True)
    h_in_np, h_out_np = h_in.numpy().view(np.uint64), h_out.numpy().view(np.uint64)
    torch.cuda.synchronize()
    for _ in range(2):
        d.lookup_batch(h_in_np, out=h_out_np)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        d.lookup_batch(h_in_np, out=h_out_np)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_s = float(t.item())
    assert torch.equal(h_out, ids.cpu()), "e2e ids differ from the sampled ids"
    e2e_value = world * n * args.steps / e2e_s
    clocks = sampler.stop() if rank == 0 else None

    # ---- gather of the ids to rank 0 (the only communication of the sharded design), timed apart.
    # sshash_b200.sharded.ShardedLookup: "peer" = every rank's lookup kernel stores its ids straight into
    # rank 0's vector through NVLink peer stores (the gather is fused into the lookup kernel);
    # "copy" = ids written locally, finished chunks pushed by copy engines; "p2p" = ids written locally +
    # chunked NCCL send/recv, kept for comparison.
    gather = None
    if world > 1:
        from sshash_b200.sharded import ShardedLookup
        gather = {"bytes_to_rank0": (world - 1) * n * 8}
        for mode in ("peer", "copy", "p2p"):
            sl = ShardedLookup.for_dictionary(d, chunk_queries=1 << 24, mode=mode)
            for _ in range(3):
                sl.lookup(kmers, dst=0)
            barrier()
            g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            g0.record()
            for _ in range(3):
                local_ids, gathered = sl.lookup(kmers, dst=0)
            g1.record()
            barrier()
            assert torch.equal(local_ids, ids)
            if rank == 0:
                assert torch.equal(gathered[:n], ids)
            gt = torch.tensor([g0.elapsed_time(g1) / 3], device=dev, dtype=torch.float64)
            dist.all_reduce(gt, op=dist.ReduceOp.MAX)
            gather[mode] = {"lookup_plus_gather_ms": float(gt.item()),
                            "lookups_per_s_with_gather": world * n / (float(gt.item()) * 1e-3)}
            del sl, gathered, local_ids
        best = min(("peer", "copy"), key=lambda m: gather[m]["lookup_plus_gather_ms"])
        gather["mode"] = best + (": ids stored by the lookup kernels straight into rank 0's vector over NVLink (symmetric memory)"
                                 if best == "peer" else ": finished chunks pushed into rank 0's symmetric vector by copy engines over NVLink")
        gather["lookup_plus_gather_ms"] = gather[best]["lookup_plus_gather_ms"]
        gather["lookups_per_s_with_gather"] = gather[best]["lookups_per_s_with_gather"]

    if rank == 0:
        peak, peak_src = measured_peak()
        per_launch_ms = float(np.mean(kernel_ms))
        achieved = B_ALG * n / (per_launch_ms * 1e-3) / 1e9
        line = {
            "metric": "k-mer lookups/sec", "value": value, "unit": "lookups/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": total_ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "queries_per_gpu_per_step": n, "index": "tests/golden/se_k31_m13.sshash",
                       "l2": "query+id streams are 1.6 GB per step (> 126 MB L2); the 2.9 MB index is L2-resident by design of cfg2",
                       "parallelism": "index replicated, queries sharded x%d" % world},
            "e2e": {"value": e2e_value, "unit": "lookups/s", "h2d_bytes_per_step": n * 8, "d2h_bytes_per_step": n * 8},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": recorded_traffic(n), "peak_source": peak_src,
                         "algorithmic_bytes_per_lookup": B_ALG, "kernel": "lookup_kernel<1,0,false>",
                         "kernel_ms_per_launch": per_launch_ms},
        }
        if gather:
            line["gather"] = gather
        if world == 1 and not args.no_scale:
            line["scale"] = scale_workload()
            line["streaming"] = streaming_workload()
        if world == 1 and not args.no_cpu_baseline:
            threads = host_threads()
            sample = h_in_np[:CPU_SAMPLE]
            v, kind = cpu_reference(sample, threads)
            v1, _ = cpu_reference(sample[: CPU_SAMPLE // 8], 1)
            line["cpu_baseline"] = {"value": v, "unit": "lookups/s", "cores": threads, "kind": kind,
                                    "sample": "first %d of the 1e8 queries, %d host threads" % (sample.size, threads),
                                    "single_thread_value": v1}
        print(json.dumps(line), flush=True)
    d.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
