#!/usr/bin/env python
"""Round-2 experiment: what bounds the lookup kernel on an index whose pilots do not fit L2?

For one index (synthetic, built here by the unmodified reference builder) time the lookup kernel on
1e8 device-resident queries (forward positives / 50 % RC mix / uniform negatives) under a set of
open-time variants (environment switches of the library), and on the SAME queries sorted by the
MPHF partition of their forward minimizer (sshash_gpu_minimizer_partition_batch + torch.sort), which
is what an on-device partition binning pass would hand to the kernel.

    python tools/exp_locality.py --strings 2500000 --length 1030 -k 31 -m 21 --workdir /tmp/ix
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))

VARIANTS = {
    "r1_layout": {"SSHASH_GPU_LOCATE": "legacy", "SSHASH_GPU_BINNED": "0"},
    "direct": {"SSHASH_GPU_BINNED": "0"},
    # (the pilot-policy variants "direct+pilots_cold" / "+pilots_hot64" of profiles/r2_exp_locality_v1.jsonl needed
    #  switches that were removed from the library after they lost)
    "direct+prefix_window": {"SSHASH_GPU_BINNED": "0", "SSHASH_GPU_L2_WINDOW": "prefix"},
    "binned": {"SSHASH_GPU_BINNED": "1"},
    "binned_noprefetch": {"SSHASH_GPU_BINNED": "1", "SSHASH_GPU_BIN_PREFETCH": "0"},
    "binned_lookahead2": {"SSHASH_GPU_BINNED": "1", "SSHASH_GPU_BIN_LOOKAHEAD": "2"},
    "default": {},
}
ENV_KEYS = ("SSHASH_GPU_L2_WINDOW", "SSHASH_GPU_LOCATE", "SSHASH_GPU_PILOTS_COLD", "SSHASH_GPU_BINNED", "SSHASH_GPU_BIN_PREFETCH", "SSHASH_GPU_BIN_LOOKAHEAD")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--strings", type=int, default=2500000)
    ap.add_argument("--length", type=int, default=1030)
    ap.add_argument("-k", type=int, default=31)
    ap.add_argument("-m", type=int, default=21)
    ap.add_argument("--queries", type=int, default=100_000_000)
    ap.add_argument("--workdir", default="/tmp/ix")
    ap.add_argument("--variants", default="direct,binned,binned_noprefetch,binned_lookahead2")
    ap.add_argument("--no-sorted", action="store_true", help="skip the runs on queries pre-sorted by partition")
    a = ap.parse_args()
    import torch
    import sshash_b200
    from bench import measured_peak, rc_packed_torch
    from bench_configs import build_index
    from scale_bench import time_lookup
    import ctypes as C
    os.makedirs(a.workdir, exist_ok=True)
    idx, build_s = build_index(a.workdir, a.strings, a.length, a.k, a.m)
    peak, _ = measured_peak()
    dev = torch.device("cuda", 0)
    try:
        from cuda.bindings import runtime as rt
        _, persist = rt.cudaDeviceGetAttribute(rt.cudaDeviceAttr.cudaDevAttrMaxPersistingL2CacheSize, 0)
        _, window = rt.cudaDeviceGetAttribute(rt.cudaDeviceAttr.cudaDevAttrMaxAccessPolicyWindowSize, 0)
        _, l2 = rt.cudaDeviceGetAttribute(rt.cudaDeviceAttr.cudaDevAttrL2CacheSize, 0)
        print(json.dumps({"l2_bytes": l2, "max_persisting_l2": persist, "max_access_policy_window": window}), flush=True)
    except Exception as e:
        print(json.dumps({"l2_query_failed": str(e)}), flush=True)
    n = a.queries
    queries = None
    for name in a.variants.split(","):
        env = VARIANTS[name]
        saved = {k: os.environ.get(k) for k in ENV_KEYS}
        for k in saved:
            os.environ.pop(k, None)
        os.environ.update(env)
        d = sshash_b200.Dictionary(idx)
        for k, v in saved.items():
            os.environ.pop(k, None)
            if v is not None:
                os.environ[k] = v
        k = d.k()
        if queries is None:
            gen = torch.Generator(device=dev).manual_seed(7)
            ids = torch.randint(0, d.num_kmers(), (n,), generator=gen, device=dev, dtype=torch.int64)
            fwd = d.access_batch(ids)
            mix = fwd.clone()
            mix[1::2] = rc_packed_torch(mix[1::2], k)
            neg = torch.randint(0, 2 ** (2 * k), (n,), generator=gen, device=dev, dtype=torch.int64)
            parts = torch.empty(n, dtype=torch.int32, device=dev)
            sets = {}
            for qname, q in (("fwd", fwd), ("mix", mix), ("neg", neg)):
                sshash_b200._lib.check(d._lib.sshash_gpu_minimizer_partition_batch(d._h, q.data_ptr(), n, parts.data_ptr(), None))
                torch.cuda.synchronize()
                order = torch.sort(parts.to(torch.int16) if d.info["mphf_partitions"] < 32768 else parts, stable=True)[1]
                sets[qname] = (q, q[order].contiguous(), order)
            del parts
            queries = sets
        out = torch.empty(n, dtype=torch.int64, device=dev)
        res = {"variant": name, "env": env, "index": os.path.basename(idx), "num_kmers": d.num_kmers(),
               "mphf_partitions": d.info["mphf_partitions"], "num_minimizers": d.info["num_minimizers"],
               "device_bytes": d.info["device_bytes"], "build_s": build_s}
        for qname, b_alg in (("fwd", 208.0), ("mix", 256.0), ("neg", 208.0)):
            q, qs, order = queries[qname]
            ms = time_lookup(d, q, out)
            if qname != "neg":
                assert torch.equal(out, ids)
            else:
                found = int((out != -1).sum())
            res[qname] = {"ms": ms, "G_lookups_per_s": n / ms / 1e6, "roofline_frac": b_alg * n / ms * 1e3 / 1e9 / peak}
            if qname == "neg":
                res[qname]["found"] = found
            if not a.no_sorted and "BINNED" not in str(env.get("SSHASH_GPU_BINNED", "0")) and env.get("SSHASH_GPU_BINNED") != "1":
                ms_sorted = time_lookup(d, qs, out)
                if qname != "neg":
                    assert torch.equal(out, ids[order])
                res[qname].update({"sorted_by_partition_G_lookups_per_s": n / ms_sorted / 1e6,
                                   "sorted_roofline_frac": b_alg * n / ms_sorted * 1e3 / 1e9 / peak})
        print(json.dumps(res), flush=True)
        d.close()
        del out


if __name__ == "__main__":
    main()
