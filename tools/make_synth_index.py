#!/usr/bin/env python
"""Build a synthetic SSHash index of a given scale with the UNMODIFIED reference builder.

Index construction is out of scope for the GPU path: the reference's own
builder (compiled into oracle/_ref by oracle/Makefile) produces the .sshash file, exactly as
`sshash build -i <fasta> -k K -m M [--canonical]` would; the GPU library only reads the file.

    python tools/make_synth_index.py --strings 500000 --length 1030 -k 31 -m 13 -o /tmp/t5e8.sshash

Strings are i.i.d. uniform ACGT (seeded): each contributes length-k+1 k-mers, all distinct with
overwhelming probability at these scales (duplicates are tolerated by the format anyway).
"""
import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def write_fasta(path, n_strings, length, seed, chunk=20000):
    rng = np.random.default_rng(seed)
    lut = np.frombuffer(b"ACGT", dtype=np.uint8)
    with open(path, "wb") as f:
        for s0 in range(0, n_strings, chunk):
            n = min(chunk, n_strings - s0)
            seq = lut[rng.integers(0, 4, size=(n, length), dtype=np.uint8)]
            rec = np.empty((n, length + 3), dtype=np.uint8)
            rec[:, 0] = ord(">")
            rec[:, 1] = ord("\n")
            rec[:, 2:-1] = seq
            rec[:, -1] = ord("\n")
            f.write(rec.tobytes())


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--strings", type=int, default=500000)
    ap.add_argument("--length", type=int, default=1030)
    ap.add_argument("-k", type=int, default=31)
    ap.add_argument("-m", type=int, default=13)
    ap.add_argument("--canonical", action="store_true")
    ap.add_argument("--seed", type=int, default=42)
    ap.add_argument("--threads", type=int, default=0)
    ap.add_argument("-o", "--output", required=True)
    ap.add_argument("--tmp", default=None)
    args = ap.parse_args()
    from oracle import ref
    max_k = 31 if args.k <= 31 else 63
    if not ref.available(max_k):
        raise SystemExit("oracle/_ref is not built (needs /root/reference: run `make -C oracle ref` in the build container)")
    tmp = args.tmp or os.path.dirname(os.path.abspath(args.output))
    os.makedirs(tmp, exist_ok=True)
    fa = os.path.join(tmp, os.path.basename(args.output) + ".fa")
    t0 = time.time()
    write_fasta(fa, args.strings, args.length, args.seed)
    t1 = time.time()
    threads = args.threads or len(os.sched_getaffinity(0))
    ref.build(fa, args.k, args.m, args.output, canonical=args.canonical, threads=threads, tmp_dir=tmp, max_k=max_k)
    t2 = time.time()
    os.remove(fa)
    print("synthetic index: %d strings x %d bases, k=%d m=%d -> %s (%.1f MB); fasta %.1fs, reference build %.1fs (%d threads)"
          % (args.strings, args.length, args.k, args.m, args.output, os.path.getsize(args.output) / 1e6, t1 - t0, t2 - t1, threads),
          file=sys.stderr)


if __name__ == "__main__":
    main()
