#!/usr/bin/env python
"""Run every BASELINE.json configuration once on one GPU and print one JSON line per config
(device-resident GPU throughput, roofline fraction by SURVEY.md 8d, reference CPU baseline on the
host cores of this box).  Results are committed under profiles/.

    python tools/bench_configs.py [--configs cfg2,t5e8,k63,stream,stream_t5e8,human] [--workdir DIR]

Indexes other than the bundled cfg-1/2 one are synthetic and are built here by the unmodified
reference builder (oracle/_ref); index construction is out of scope of the GPU path.
"""
import argparse
import json
import os
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))

BUNDLED = os.path.join(ROOT, "tests", "golden", "se_k31_m13.sshash")


def threads():
    return len(os.sched_getaffinity(0))


def build_index(workdir, strings, length, k, m, canonical=False):
    import make_synth_index as msi
    from oracle import ref
    idx = os.path.join(workdir, "synth_%d_%d_k%d_m%d%s.sshash" % (strings, length, k, m, "_c" if canonical else ""))
    t0 = time.time()
    if not os.path.exists(idx):
        fa = idx + ".fa"
        msi.write_fasta(fa, strings, length, 42)
        ref.build(fa, k, m, idx, canonical=canonical, threads=threads(), tmp_dir=workdir, max_k=31 if k <= 31 else 63)
        os.remove(fa)
    return idx, time.time() - t0


def lookup_config(name, idx, k, b_alg_fwd, b_alg_mix, queries=100_000_000, cpu_sample=20_000_000):
    import torch
    import sshash_b200
    from bench import measured_peak, rc_packed_torch, rc_packed_torch2
    from scale_bench import time_lookup
    from oracle import ref
    max_k = 31 if k <= 31 else 63
    d = sshash_b200.Dictionary(idx, max_k=max_k)
    dev = torch.device("cuda", 0)
    gen = torch.Generator(device=dev).manual_seed(7)
    n = queries
    ids = torch.randint(0, d.num_kmers(), (n,), generator=gen, device=dev, dtype=torch.int64)
    fwd = d.access_batch(ids)
    out = torch.empty(n, dtype=torch.int64, device=dev)
    peak, _ = measured_peak()
    res = {"config": name, "index": os.path.basename(idx), "k": k, "m": d.m(), "canonical": d.canonical(),
           "num_kmers": d.num_kmers(), "index_bytes": d.info["index_file_bytes"], "queries": n, "gpu": {}}
    ms = time_lookup(d, fwd, out)
    assert torch.equal(out, ids)
    res["gpu"]["positive_forward"] = {"lookups_per_s": n / ms * 1e3, "roofline_frac": b_alg_fwd * n / ms * 1e3 / 1e9 / peak}
    mix = fwd.clone()
    if d.words == 1:
        mix[1::2] = rc_packed_torch(mix[1::2], k)
    else:
        lo, hi = rc_packed_torch2(mix[1::2, 0], mix[1::2, 1], k)
        mix[1::2, 0], mix[1::2, 1] = lo, hi
    ms = time_lookup(d, mix, out)
    assert torch.equal(out, ids)
    res["gpu"]["positive_50rc"] = {"lookups_per_s": n / ms * 1e3, "roofline_frac": b_alg_mix * n / ms * 1e3 / 1e9 / peak}
    if d.words == 1:
        neg = torch.randint(0, 2 ** (2 * k), (n,), generator=gen, device=dev, dtype=torch.int64)
    else:
        neg = torch.randint(0, 2 ** 62, (n, 2), generator=gen, device=dev, dtype=torch.int64)
        neg[:, 1] &= (1 << (2 * k - 64)) - 1
    ms = time_lookup(d, neg, out)
    res["gpu"]["negative"] = {"lookups_per_s": n / ms * 1e3, "found": int((out != -1).sum())}
    # reference CPU: same queries (50 % RC positives), all host threads and one thread
    if ref.available(max_k):
        rd = ref.RefDictionary(idx, max_k=max_k)
        sample = mix[: cpu_sample].cpu().numpy().view(np.uint64).reshape(-1)
        chk = min(1_000_000, cpu_sample)          # GPU ids == the reference CPU dictionary's ids on the first 1e6 queries
        got = rd.lookup(sample[: chk * d.words], threads=threads())
        assert (got.view(np.int64) == ids[:chk].cpu().numpy()).all()
        res["checked_vs_reference"] = chk
        t = threads()
        secs = rd.time_lookup(sample, threads=t)
        one = rd.time_lookup(sample[: (cpu_sample // 8) * d.words], threads=1)
        res["cpu_reference"] = {"positive_50rc_lookups_per_s": cpu_sample / secs, "threads": t,
                                "single_thread_lookups_per_s": (cpu_sample // 8) / one, "sample": cpu_sample}
        rd.close()
    d.close()
    return res


def stream_config(name, idx, k, reads=10_000_000, cpu_reads=200_000):
    import torch
    import sshash_b200
    from bench import measured_peak
    from stream_bench import make_reads
    from oracle import ref
    max_k = 31 if k <= 31 else 63
    d = sshash_b200.Dictionary(idx, max_k=max_k)
    bases, offs = make_reads(d, reads)
    nwin = reads * (150 - k + 1)
    peak, _ = measured_peak()
    res = {"config": name, "index": os.path.basename(idx), "reads": reads, "windows": nwin, "gpu": {}}
    for want in (False, True):
        d.streaming_batch(bases, offs, want_ids=want)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(3):
            ids, rep = d.streaming_batch(bases, offs, want_ids=want)
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / 3
        res["gpu"]["device_ids" if want else "device_report_only"] = {
            "windows_per_s": nwin / dt, "roofline_frac": 106.0 * nwin / dt / 1e9 / peak}
    del ids
    res["report"] = rep
    hb = torch.empty(bases.numel(), dtype=torch.uint8, pin_memory=True)
    hb.copy_(bases)
    ho = offs.cpu().numpy().view(np.uint64)
    hbn = hb.numpy()
    d.streaming_batch(hbn, ho, want_ids=False)
    t0 = time.perf_counter()
    _, rep_h = d.streaming_batch(hbn, ho, want_ids=False)
    dt = time.perf_counter() - t0
    assert rep_h == rep
    res["gpu"]["host_report_only"] = {"windows_per_s": nwin / dt, "h2d_bytes": int(bases.numel() + ho.nbytes)}
    if ref.available(max_k):
        rd = ref.RefDictionary(idx, max_k=max_k)
        m = cpu_reads
        sub = hbn[: m * 150].tobytes()
        oids, _, orep, secs = rd.streaming_reads(sub, ho[: m + 1])
        gids, grep = d.streaming_batch(hbn[: m * 150], ho[: m + 1])
        assert (gids == oids).all() and grep == orep, "GPU streaming differs from the reference"
        res["cpu_reference"] = {"single_thread_windows_per_s": oids.size / secs, "reads": m}
        # all host threads: one reference streaming_query per chunk of reads
        t = threads()
        per = m // t
        outs = [None] * t

        def work(i):
            lo, hi = i * per, (i + 1) * per
            o = (ho[lo: hi + 1] - ho[lo]).copy()
            outs[i] = rd.streaming_reads(hbn[lo * 150: hi * 150].tobytes(), o)[3]
        th = [threading.Thread(target=work, args=(i,)) for i in range(t)]
        t0 = time.perf_counter()
        [x.start() for x in th]
        [x.join() for x in th]
        wall = time.perf_counter() - t0
        res["cpu_reference"].update({"threads": t, "all_threads_windows_per_s": per * t * (150 - k + 1) / wall})
        rd.close()
    d.close()
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--configs", default="cfg2,t5e8,t5e8_canonical,k63,stream,stream_t5e8,human")
    ap.add_argument("--workdir", default=None)
    ap.add_argument("--keep", action="store_true", help="keep the synthetic indexes (for a following ncu pass)")
    a = ap.parse_args()
    wd = a.workdir or tempfile.mkdtemp(prefix="sshash_cfg_")
    os.makedirs(wd, exist_ok=True)
    for c in a.configs.split(","):
        t0 = time.time()
        if c == "cfg2":
            r = lookup_config("cfg2: bundled S.enterica k31 m13, 1e8 queries", BUNDLED, 31, 208.0, 256.0)
        elif c == "t5e8":
            idx, bs = build_index(wd, 500000, 1030, 31, 17)
            r = lookup_config("T5e8: synthetic 5e8 k-mers k31 m17 regular", idx, 31, 208.0, 256.0)
            r["build_s"] = bs
        elif c == "t5e8_canonical":
            idx, bs = build_index(wd, 500000, 1030, 31, 17, canonical=True)
            r = lookup_config("T5e8 canonical: synthetic 5e8 k-mers k31 m17 canonical", idx, 31, 208.0, 208.0)
            r["build_s"] = bs
            os.remove(idx)
        elif c == "k63":
            idx, bs = build_index(wd, 500000, 1062, 63, 25)
            r = lookup_config("cfg4 (scaled): synthetic 5e8 k-mers k63 m25", idx, 63, 248.0, 296.0)
            r["build_s"] = bs
            os.remove(idx)
        elif c == "k63_3e9":
            # BASELINE configs[3] at its stated scale: k=63 m=25 (script/build.py:28-30), ~3e9 k-mers, HBM-resident
            idx, bs = build_index(wd, 3000000, 1062, 63, 25)
            r = lookup_config("cfg4: synthetic 3e9 k-mers k63 m25 (3e6 strings x 1062 bases), HBM-resident", idx, 63, 248.0, 296.0)
            r["build_s"] = bs
            if not a.keep:
                os.remove(idx)
        elif c == "k63_canonical":
            idx, bs = build_index(wd, 500000, 1062, 63, 25, canonical=True)
            r = lookup_config("synthetic 5e8 k-mers k63 m25 canonical", idx, 63, 248.0, 248.0)
            r["build_s"] = bs
            os.remove(idx)
        elif c == "human":
            idx, bs = build_index(wd, 2500000, 1030, 31, 21)
            r = lookup_config("cfg5 index: synthetic human-scale 2.5e9 k-mers k31 m21", idx, 31, 208.0, 256.0)
            r["build_s"] = bs
            if not a.keep:
                os.remove(idx)
        elif c == "stream":
            r = stream_config("cfg3 on the cfg-1 index: 1e7 synthetic 150-bp reads, 50 % hit", BUNDLED, 31)
        elif c == "stream_t5e8":
            idx, bs = build_index(wd, 500000, 1030, 31, 17)
            r = stream_config("cfg3 on T5e8: 1e7 synthetic 150-bp reads, 50 % hit", idx, 31)
        else:
            continue
        r["wall_s"] = time.time() - t0
        print(json.dumps(r), flush=True)


if __name__ == "__main__":
    main()
