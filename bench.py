#!/usr/bin/env python
"""bench.py -- k-mer lookups/s of the batched Lookup hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Workload (BASELINE.json configs[1]): the bundled S. enterica k=31 m=13 index
(tests/golden/se_k31_m13.sshash, built by the reference builder), 1e8 uniform-random POSITIVE
k-mers per GPU, every other one reverse-complemented (the reference's own bench protocol,
tools/perf.hpp:41-51).  One step = one pass of the lookup path over the whole 1e8-query batch.

  value : N = 1: lookups/s with queries and ids resident in HBM (CUDA events).
          N > 1: every rank looks up its own 1e8-query shard AND the ids of all ranks are gathered on
          rank 0 inside the timed region (BASELINE configs[4]'s definition; 32-bit ids, see `config.ids`);
          max over ranks.  The replicas-only figure (no gather) is `roofline.replicas_only_lookups_per_s`.
  e2e   : the same batch through the C-ABI call with HOST (pinned) buffers: H2D of the packed
          k-mers and D2H of the ids inside the timed region (+ the 32-bit-id and membership forms)
  roofline : algorithmic bytes per lookup (256 B for a regular index at 50% RC)
             x lookups/s of the lookup kernel vs the measured HBM peak; `traffic` = DRAM bytes of one
             launch MEASURED on this box by an ncu pass over the same kernel and inputs after the timed
             region (null when ncu is not usable).  cfg2's index is L2-resident by definition, so the
             HBM-resident evidence is in roofline.hbm_resident: a 5e8-k-mer index built on the box.
  cpu_baseline : the reference's own dictionary::lookup on the box's host cores, bounded sample

--impl reference times the reference's CPU implementation (oracle/_ref when built, else the C
oracle port) with all host threads on the same workload and `config`.
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
NVLINK_PEER_GBS = 770.0   # measured peer-copy bandwidth per direction per GPU on this pool (B200_PROFILING.md)


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


def measured_traffic(index: str, mode: str, n_queries: int, env=None, kernel: str = "lookup_kernel"):
    """DRAM bytes (read + write) of ONE launch of `kernel` over n_queries queries of kind `mode`, measured on
    this box by an ncu pass over tools/ncu_target.py (same library, same index, same query generator
    and seed as the timed legs); None when ncu cannot run here."""
    import shutil
    ncu = shutil.which("ncu") or "/usr/local/cuda/bin/ncu"
    if not os.path.exists(ncu):
        return None
    cmd = [ncu, "--metrics", "dram__bytes_read.sum,dram__bytes_write.sum", "--clock-control", "none", "-k",
           "regex:" + kernel, "-s", "2", "-c", "1", "--csv", sys.executable, os.path.join(ROOT, "tools", "ncu_target.py"),
           "--index", index, "--mode", mode, "--queries", str(n_queries)]
    try:
        e = dict(os.environ)
        e.update(env or {})
        out = subprocess.run(cmd, capture_output=True, text=True, timeout=300, env=e).stdout
        import csv
        rows = [r for r in csv.reader(out.splitlines()) if len(r) > 3]
        hdr = next(r for r in rows if "Metric Name" in r)
        ni, ui, vi = hdr.index("Metric Name"), hdr.index("Metric Unit"), hdr.index("Metric Value")
        scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
        total = 0.0
        for r in rows:
            if len(r) > vi and r[ni] in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                total += float(r[vi].replace(",", "")) * scale[r[ui]]
        return total or None
    except Exception:
        return None


def make_config(n: int, world: int) -> dict:
    """`config` of the JSON line: identical for the B200 arm and the reference arm."""
    return {"workload": WORKLOAD, "queries_per_gpu_per_step": n, "index": "tests/golden/se_k31_m13.sshash",
            "l2": "query+id streams are 1.6 GB per step (> 126 MB L2); the 2.9 MB index is L2-resident by design of cfg2",
            "parallelism": "index replicated, queries sharded x%d" % world,
            "ids": "u64 at N=1; at N>1 the timed value gathers 32-bit ids on rank 0 (num_kmers < 2^32; the u64 gather is in roofline.gather)"}


def hbm_resident_workload():
    """Secondary, HBM-resident workload: a synthetic 5e8-k-mer k=31 m=17 index
    built on this box by the unmodified reference builder (oracle/_ref; index construction is out of
    scope), 1e8 positive / 50 % RC / negative queries, device-resident, + the DRAM bytes of one launch of
    the 50 % RC leg measured by ncu.  Goes into roofline.hbm_resident of the headline line."""
    try:
        import tempfile
        sys.path.insert(0, os.path.join(ROOT, "tools"))
        import scale_bench
        wd = tempfile.mkdtemp(prefix="sshash_scale_")
        r = scale_bench.run(500000, 1030, 31, 17, False, 100_000_000, wd, keep=True)
        peak, _ = measured_peak()
        out = {"workload": "T5e8: synthetic 5e5 strings x 1030 bases, k=31 m=17 (5e8 k-mers, 395 MB index, HBM-resident), 1e8 queries",
               "bound": "hbm random-access rate", "peak": peak, "unit": "GB/s"}
        for key, b_alg in (("positive_forward", 208.0), ("positive_50rc", 256.0), ("negative", 208.0)):
            v = r[key]["lookups_per_s"]
            out[key] = {"lookups_per_s": v, "algorithmic_bytes_per_lookup": b_alg, "achieved": b_alg * v / 1e9,
                        "frac": b_alg * v / 1e9 / peak, "kernel_ms_per_launch": r[key]["ms"]}
        idx = os.path.join(wd, r["index"])
        t = measured_traffic(idx, "mix", 100_000_000)
        out["positive_50rc"]["traffic"] = t
        if t:
            out["positive_50rc"]["dram_bytes_per_lookup"] = t / 1e8
            out["positive_50rc"]["dram_GBps"] = t / (r["positive_50rc"]["ms"] * 1e-3) / 1e9
        try:
            os.remove(idx)
        except OSError:
            pass
        return out
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


def make_queries_numpy(n: int, seed: int, threads: int) -> np.ndarray:
    """The same workload without a GPU (reference arm): positives via access(), every other one
    reverse-complemented; generated on `threads` host threads."""
    import torch
    from concurrent.futures import ThreadPoolExecutor
    from oracle import port, ref
    d = ref.RefDictionary(INDEX, max_k=31) if ref.available(31) else port.OracleDictionary(INDEX)
    nk, k = int(d.num_kmers), int(d.k)
    out = np.empty(n, dtype=np.uint64)
    step = max(1, (n + threads - 1) // threads)

    def work(t):
        lo, hi = t * step, min(n, (t + 1) * step)
        if lo >= hi:
            return
        ids = np.random.default_rng(seed + t).integers(0, nk, hi - lo).astype(np.uint64)
        out[lo:hi] = d.access(ids).reshape(-1)

    with ThreadPoolExecutor(threads) as ex:
        list(ex.map(work, range(threads)))
    t = torch.from_numpy(out.view(np.int64))
    t[1::2] = rc_packed_torch(t[1::2], k)
    return out


def run_reference(args, rank: int, world: int):
    if rank != 0:
        return
    threads = host_threads()
    n = args.queries
    q = make_queries_numpy(n, 42, threads)
    for _ in range(args.warmup):
        cpu_reference(q[: max(n // 10, 1)], threads)
    kind = "reference"
    vals = []
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
        "config": make_config(n, world),
        "cpu_baseline": {"value": value, "unit": "lookups/s", "cores": threads, "kind": kind,
                         "sample": "the whole %d-query batch per step, %d host threads" % (n, threads)},
        "e2e": {"value": value, "unit": "lookups/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def widen_ids(t):
    """ids as int64 bit patterns: 32-bit ids (int32 tensors, UINT32_MAX = not found) are zero-extended, not found -> -1"""
    import torch
    if t.dtype == torch.int64:
        return t
    return torch.where(t == -1, torch.full((), -1, dtype=torch.int64, device=t.device), t.to(torch.int64) & 0xFFFFFFFF)


def time_e2e(fn, steps: int, barrier, dev, world: int):
    """wall time of `steps` calls of a host-buffer entry point, max over ranks"""
    import torch
    import torch.distributed as dist
    for _ in range(2):
        fn()
    barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        fn()
    torch.cuda.synchronize()
    secs = time.perf_counter() - t0
    t = torch.tensor([secs], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--queries", type=int, default=QUERIES_PER_GPU, help="queries per GPU per step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-scale", action="store_true", help="skip the HBM-resident 5e8-k-mer and the streaming secondary workloads")
    ap.add_argument("--gather-chunk", type=int, default=1 << 22, help="N > 1: queries per lookup launch / per pushed piece of the gather")
    ap.add_argument("--no-ncu", action="store_true", help="do not measure roofline.traffic with an ncu pass after the timed region")
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
    torch.cuda.synchronize()

    def step_device():
        d.lookup_batch(kmers, out=out)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        """CUDA-event time (ms) of `steps` calls of fn on the current stream, max over ranks; per-step times of this rank"""
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
        barrier()
        ev[0].record()
        for i in range(steps):
            fn()
            ev[i + 1].record()
        barrier()
        total = ev[0].elapsed_time(ev[-1])
        t = torch.tensor([total], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()), [ev[i].elapsed_time(ev[i + 1]) for i in range(steps)]

    # ---- device-resident lookups, no communication (N = 1: the value; N > 1: the replicas-only side figure) ----
    for _ in range(args.warmup):
        step_device()
    torch.cuda.synchronize()
    assert torch.equal(out, ids), "lookup ids differ from the sampled ids"   # positives are self-checking
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = sshash_b200.launch_count()
    replica_ms, kernel_ms = timed(step_device, args.steps)
    launches = sshash_b200.launch_count() - launches0
    replica_value = world * n * args.steps / (replica_ms * 1e-3)
    total_ms, value = replica_ms, replica_value

    # ---- N > 1: the timed value = lookups + gather of ALL ids on rank 0 (BASELINE configs[4]) -----------------
    gather = None
    if world > 1:
        from sshash_b200.sharded import ShardedLookup
        gather = {}
        # every rank's checksums of its own ids: rank 0 verifies EVERY slice of the gathered vector against them
        w_local = torch.arange(1, n + 1, device=dev, dtype=torch.int64)
        mine = torch.stack([ids.sum(), (ids * w_local).sum()])
        sums = [torch.zeros(2, dtype=torch.int64, device=dev) for _ in range(world)]
        dist.all_gather(sums, mine)

        def verify(gathered, what):
            if rank != 0:
                return
            for r in range(world):
                sl = widen_ids(gathered[r * n:(r + 1) * n])
                got = torch.stack([sl.sum(), (sl * w_local).sum()])
                assert torch.equal(got, sums[r]), "gathered ids of rank %d are wrong (%s)" % (r, what)

        for ids32, modes in ((True, ("copy", "staged", "peer")), (False, ("copy", "staged", "peer", "p2p"))):
            for mode in modes:
                key = ("u32_" if ids32 else "u64_") + mode
                sl = ShardedLookup.for_dictionary(d, chunk_queries=args.gather_chunk, mode=mode, ids32=ids32)
                state = {}

                def step_gather():
                    state["r"] = sl.lookup(kmers, dst=0, sizes=[n] * world)

                for _ in range(3):
                    step_gather()
                torch.cuda.synchronize()
                l0 = sshash_b200.launch_count()
                g_ms, _ = timed(step_gather, args.steps)
                g_launches = sshash_b200.launch_count() - l0
                local_ids, gathered = state["r"]
                assert torch.equal(widen_ids(local_ids), ids), "local ids differ from the sampled ids"
                verify(gathered, "%s ids, mode %s" % ("u32" if ids32 else "u64", mode))
                bytes_in = (world - 1) * n * (4 if ids32 else 8)
                gather[key] = {"ms_per_step": g_ms / args.steps, "lookups_per_s": world * n * args.steps / (g_ms * 1e-3),
                               "rank0_ingress_GBps": bytes_in * args.steps / (g_ms * 1e-3) / 1e9, "launches": int(g_launches)}
                del sl, gathered, local_ids, state
        best = max((k_ for k_ in gather if k_.startswith("u32_")), key=lambda k_: gather[k_]["lookups_per_s"])
        value = gather[best]["lookups_per_s"]
        total_ms = gather[best]["ms_per_step"] * args.steps
        launches = gather[best]["launches"]
        gather["timed_mode"] = best
        gather["rank0_ingress_ceiling_GBps"] = NVLINK_PEER_GBS
        gather["every_slice_verified_on_rank0"] = True

    # ---- end-to-end through the C ABI with host (pinned) buffers ----------------------------------
    h_in = torch.empty(n, dtype=torch.int64, pin_memory=True)
    h_in.copy_(kmers)
    h_out = torch.empty(n, dtype=torch.int64, pin_memory=True)
    h_in_np, h_out_np = h_in.numpy().view(np.uint64), h_out.numpy().view(np.uint64)
    torch.cuda.synchronize()
    e2e_s = time_e2e(lambda: d.lookup_batch(h_in_np, out=h_out_np), args.steps, barrier, dev, world)
    assert torch.equal(h_out, ids.cpu()), "e2e ids differ from the sampled ids"
    e2e_value = world * n * args.steps / e2e_s
    e2e = {"value": e2e_value, "unit": "lookups/s", "h2d_bytes_per_step": n * 8, "d2h_bytes_per_step": n * 8}
    # the same call with 32-bit ids and with membership bytes (4 / 1 bytes back per lookup instead of 8)
    h32 = torch.empty(n, dtype=torch.int32, pin_memory=True)
    h32_np = h32.numpy().view(np.uint32)
    s32 = time_e2e(lambda: d.lookup_batch_u32(h_in_np, out=h32_np), max(2, args.steps // 2), barrier, dev, world)
    assert torch.equal(widen_ids(h32), ids.cpu()), "e2e u32 ids differ from the sampled ids"
    e2e["u32_ids"] = {"value": world * n * max(2, args.steps // 2) / s32, "h2d_bytes_per_step": n * 8, "d2h_bytes_per_step": n * 4}
    hmem = torch.empty(n, dtype=torch.uint8, pin_memory=True)
    hmem_np = hmem.numpy()
    smem_ = time_e2e(lambda: d.is_member_batch(h_in_np, out=hmem_np), max(2, args.steps // 2), barrier, dev, world)
    assert bool(hmem.all()), "e2e membership: every query is a positive"
    e2e["is_member"] = {"value": world * n * max(2, args.steps // 2) / smem_, "h2d_bytes_per_step": n * 8, "d2h_bytes_per_step": n}
    # ceiling of the host-buffer leg, measured here: the same 8 B in + 8 B out per lookup moved concurrently by
    # all ranks with no kernels at all (pinned memory, two streams, 32 MB pieces)
    h_out = torch.empty(n, dtype=torch.int64, pin_memory=True)
    s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()
    piece = (32 << 20) // 8

    def duplex():
        for lo in range(0, n, piece):
            hi = min(n, lo + piece)
            with torch.cuda.stream(s_in):
                kmers[lo:hi].copy_(h_in[lo:hi], non_blocking=True)
            with torch.cuda.stream(s_out):
                h_out[lo:hi].copy_(out[lo:hi], non_blocking=True)
        s_in.synchronize()
        s_out.synchronize()

    dsecs = time_e2e(duplex, 3, barrier, dev, world)
    e2e["pcie_duplex_ceiling"] = {"value": world * n * 3 / dsecs, "unit": "lookups/s",
                                  "GBps_per_direction_all_gpus": world * n * 8 * 3 / dsecs / 1e9,
                                  "what": "16 B per lookup over PCIe, all %d rank(s) at once, no kernels" % world}
    e2e["frac_of_pcie_ceiling"] = e2e_value / e2e["pcie_duplex_ceiling"]["value"]
    del h32, h_out, hmem
    clocks = sampler.stop() if rank == 0 else None

    # ---- N > 1: ONE process, ONE C-ABI handle over all N GPUs (sshash_gpu_multi_*), rank 0 only ------------
    multi = None
    if world > 1:
        store = dist.distributed_c10d._get_default_store()
        barrier()
        if rank == 0:
            try:
                m = sshash_b200.MultiDictionary(INDEX, devices=list(range(world)))
                nq = world * min(n, 50_000_000)
                hq = torch.empty(nq, dtype=torch.int64, pin_memory=True)
                hq.view(world, -1)[:] = h_in[: nq // world]
                ho = torch.empty(nq, dtype=torch.int64, pin_memory=True)
                hq_np, ho_np = hq.numpy().view(np.uint64), ho.numpy().view(np.uint64)
                m.lookup_batch(hq_np, out=ho_np)
                t0 = time.perf_counter()
                for _ in range(3):
                    m.lookup_batch(hq_np, out=ho_np)
                dt = (time.perf_counter() - t0) / 3
                assert torch.equal(ho.view(world, -1)[world - 1], ids.cpu()[: nq // world])
                multi = {"what": "sshash_gpu_multi_lookup_batch: one process, one handle, host buffers sharded over %d GPUs" % world,
                         "queries": nq, "lookups_per_s": nq / dt}
                dq = hq.to(dev)
                do = torch.empty(nq, dtype=torch.int64, device=dev)
                m.lookup_batch(dq, out=do)
                t0 = time.perf_counter()
                for _ in range(3):
                    m.lookup_batch(dq, out=do)
                dt = (time.perf_counter() - t0) / 3
                assert torch.equal(do.view(world, -1)[world - 1].cpu(), ids.cpu()[: nq // world])
                multi["device_buffers_on_gpu0_lookups_per_s"] = nq / dt
                m.close()
                del hq, ho, dq, do
            except Exception as e:   # noqa: BLE001
                multi = {"unavailable": "%s: %s" % (type(e).__name__, e)}
            store.set("multi_leg_done", "1")
        else:
            store.wait(["multi_leg_done"])     # host-side wait: the other ranks keep their GPUs idle meanwhile
        barrier()

    if rank == 0:
        peak, peak_src = measured_peak()
        per_launch_ms = float(np.mean(kernel_ms))
        achieved = B_ALG * n / (per_launch_ms * 1e-3) / 1e9
        traffic = None if args.no_ncu else measured_traffic(INDEX, "mix", n)
        roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                    "traffic": traffic, "peak_source": peak_src, "algorithmic_bytes_per_lookup": B_ALG,
                    "kernel": "lookup_kernel<1,0,false>", "kernel_ms_per_launch": per_launch_ms,
                    "traffic_source": "ncu pass over the same kernel on this box after the timed region" if traffic else None,
                    "binding_limit": "sm issue/ALU -- cfg2's 2.9 MB index is L2-resident, so `achieved` is algorithmic bytes, "
                                     "not DRAM traffic; see hbm_resident for an index that lives in HBM"}
        if world > 1:
            roofline["replicas_only_lookups_per_s"] = replica_value
            roofline["gather"] = gather
            if multi:
                roofline["one_process_multi_gpu_handle"] = multi
        if world == 1 and not args.no_scale:
            roofline["hbm_resident"] = hbm_resident_workload()
        line = {
            "metric": "k-mer lookups/sec", "value": value, "unit": "lookups/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": total_ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
            "config": make_config(n, world),
            "e2e": e2e,
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": roofline,
        }
        if world == 1 and not args.no_cpu_baseline:
            threads = host_threads()
            sample = h_in_np[:CPU_SAMPLE]
            v, kind = cpu_reference(sample, threads)
            v1, _ = cpu_reference(sample[: CPU_SAMPLE // 8], 1)
            line["cpu_baseline"] = {"value": v, "unit": "lookups/s", "cores": threads, "kind": kind,
                                    "sample": "first %d of the 1e8 queries, %d host threads" % (sample.size, threads),
                                    "single_thread_value": v1}
        if world == 1 and not args.no_scale:
            line["streaming"] = streaming_workload()
        print(json.dumps(line), flush=True)
    d.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
