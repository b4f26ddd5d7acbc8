#!/usr/bin/env python
"""Regenerate tests/golden/*: index fixtures + golden query/answer vectors.

Runs ONLY in the build container, where /root/reference exists: every index is built by the
unmodified reference builder and every expected answer is produced by the unmodified reference
dictionary (oracle/_ref/libsshash_ref{31,63}.so, see oracle/Makefile).  The outputs are committed
so that the GPU box (which has no /root/reference) can check both the C oracle and the CUDA path
against the reference's own answers.

    python tests/golden/make_golden.py

Fixtures (name: source file, k, m, mode, #sequences taken from the head of the file):
"""
import gzip
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import ref  # noqa: E402

DATA = "/root/reference/data/unitigs_stitched/"
TMP = "/tmp/sshash_golden_tmp"

# name, source, k, m, canonical, nseq (None = whole file), build lib (31/63) [, weighted]
FIXTURES = [
    ("se_k31_m13", "salmonella_enterica_k31_ust.fa.gz", 31, 13, False, None, 31),   # BASELINE cfg 1/2
    ("sal100_k31_m7_reg", "salmonella_100_k31_ust.fa.gz", 31, 7, False, 3000, 31),  # heavy buckets, 2 skew partitions
    ("sal100_k31_m7_canon", "salmonella_100_k31_ust.fa.gz", 31, 7, True, 3000, 31),  # canonical + heavy
    ("sal100_k31_m11_canon", "salmonella_100_k31_ust.fa.gz", 31, 11, True, 5000, 31),  # canonical, singleton+midload
    ("sal100_k31_m7_reg_b63", "salmonella_100_k31_ust.fa.gz", 31, 7, False, 3000, 63),  # k<=31 written by the 63 build
    ("se_k63_m21", "se.ust.k63.fa.gz", 63, 21, False, 60, 63),                        # 128-bit k-mers
    ("se_k63_m7_reg", "se.ust.k63.fa.gz", 63, 7, False, 30, 63),                      # 128-bit + heavy (16-byte MPHF keys)
    ("se_k63_m8_canon", "se.ust.k63.fa.gz", 63, 8, True, 30, 63),                     # 128-bit canonical + heavy
    ("se_k47_m8", "se.ust.k47.fa.gz", 47, 8, False, 30, 63),                          # 32 < k < 63
    # weighted dictionaries (README "--weighted"): abundances in the ab:Z: header field
    ("sakai_k31_m13_weighted", "with_weights/ecoli_sakai.ust.k31.fa.gz", 31, 13, False, 400, 31, True),
    ("ecoli_k31_m11_canon_weighted", "with_weights/ecoli.ust.k31.fa.gz", 31, 11, True, 60, 31, True),
]
NQ = 12000  # positives per fixture (+ as many negatives)

# Synthetic inputs that break SSHash's "distinct k-mers" input contract (README) -- the
# reference still builds and answers them, and the CUDA path must give the same answers:
#   name, k, m, canonical, build lib
SYNTH = [
    ("twins_k31_m13_reg", 31, 13, False, 31),     # k-mers present together with their reverse complements + duplicated k-mers
    ("twins_k31_m13_canon", 31, 13, True, 31),    # duplicated k-mers (both strands) in a canonical index
]


def synth_twins_fasta(path, rng):
    """400 random strings of 300 bases; 60 segments of 70 bases are copied into another string, every
    other one reverse-complemented: a regular index then holds 40-k-mer runs twice (duplicates) or in
    both orientations (rc twins, as separate entries)."""
    comp = {"A": "T", "C": "G", "G": "C", "T": "A"}
    strings = ["".join("ACGT"[i] for i in rng.integers(0, 4, 300)) for _ in range(400)]
    for j in range(60):
        a, b = int(rng.integers(0, 400)), int(rng.integers(0, 400))
        if a == b:
            continue
        pa, pb = int(rng.integers(0, 230)), int(rng.integers(0, 230))
        seg = strings[a][pa:pa + 70]
        if j % 2:
            seg = "".join(comp[c] for c in reversed(seg))
        strings[b] = strings[b][:pb] + seg + strings[b][pb + 70:]
    with open(path, "w") as f:
        for i, s in enumerate(strings):
            f.write(">%d\n%s\n" % (i, s))
    return strings


def synth_twin_reads(strings, rng, nreads, k):
    """reads = substrings of the input strings (both strands), so that they run through the copied
    segments in every combination of orientations; plus a few random reads"""
    comp = {"A": "T", "C": "G", "G": "C", "T": "A"}
    reads = []
    for r in range(nreads):
        if r % 8 == 7:
            reads.append("".join("ACGT"[i] for i in rng.integers(0, 4, 150)))
            continue
        s = strings[int(rng.integers(0, len(strings)))]
        p = int(rng.integers(0, len(s) - 150))
        t = s[p:p + 150]
        if rng.integers(0, 2):
            t = "".join(comp[c] for c in reversed(t))
        reads.append(t)
    return reads


def subset(src, nseq, dst):
    with gzip.open(src, "rt") as f, open(dst, "w") as g:
        c = 0
        for line in f:
            g.write(line)
            if not line.startswith(">"):
                c += 1
                if nseq is not None and c >= nseq:
                    break


def rc_packed(x_lo, x_hi, k):
    """reverse complement of packed k-mers (A0 C1 T2 G3; complement = xor 2) on python ints."""
    out = []
    for lo, hi in zip(x_lo.tolist(), x_hi.tolist()):
        x = lo | (hi << 64)
        y = 0
        for i in range(k):
            y |= (((x >> (2 * i)) & 3) ^ 2) << (2 * (k - 1 - i))
        out.append(y)
    return out


def make_queries(d, rng, words):
    n = d.num_kmers
    k = d.k
    pid = rng.integers(0, n, NQ).astype(np.uint64)
    pid[:4] = [0, 1, n - 1, n - 2]
    pos = d.access(pid).reshape(NQ, words)
    lo = pos[:, 0].copy()
    hi = pos[:, 1].copy() if words == 2 else np.zeros(NQ, dtype=np.uint64)
    # odd-indexed positives are reverse-complemented (tools/perf.hpp:41-51)
    rcs = rc_packed(lo[1::2], hi[1::2], k)
    lo[1::2] = [y & (2**64 - 1) for y in rcs]
    hi[1::2] = [y >> 64 for y in rcs]
    # negatives: 1/2 uniform random k-mers, 1/2 single-base mutations of positives (same minimizer,
    # absent k-mer -> exercises the bucket scan / heavy-bucket out-of-range path)
    nneg = NQ
    mask = (1 << (2 * k)) - 1
    rnd = [int(a) | (int(b) << 64) for a, b in zip(rng.integers(0, 2**63, nneg // 2), rng.integers(0, 2**63, nneg // 2))]
    rnd = [r & mask for r in rnd]
    mut = []
    for j in range(nneg - nneg // 2):
        x = int(lo[j]) | (int(hi[j]) << 64)
        p = int(rng.integers(0, k))
        x ^= int(rng.integers(1, 4)) << (2 * p)
        mut.append(x)
    neg = rnd + mut
    nlo = np.array([x & (2**64 - 1) for x in neg], dtype=np.uint64)
    nhi = np.array([x >> 64 for x in neg], dtype=np.uint64)
    lo = np.concatenate([lo, nlo])
    hi = np.concatenate([hi, nhi])
    if words == 1:
        return pid, lo
    return pid, np.stack([lo, hi], axis=1).reshape(-1)


def synth_reads(d, rng, nreads, read_len=150):
    """BASELINE cfg 3 shape: half substrings of indexed strings (random strand), half iid ACGT,
    a few reads with an N, a few too short."""
    k = d.k
    comp = {"A": "T", "C": "G", "G": "C", "T": "A"}
    words = d.words
    reads = []
    for r in range(nreads):
        if r % 2 == 0:
            # walk consecutive k-mer ids from a random start: a substring of one indexed string
            start = int(rng.integers(0, d.num_kmers))
            ids = np.arange(start, min(d.num_kmers, start + read_len - k + 1), dtype=np.uint64)
            km = d.access(ids).reshape(len(ids), words)
            def tostr(row):
                x = int(row[0]) | ((int(row[1]) << 64) if words == 2 else 0)
                return "".join("ACTG"[(x >> (2 * i)) & 3] for i in range(k))
            s = tostr(km[0])
            prev = km[0]
            for row in km[1:]:
                t = tostr(row)
                if t[:-1] != s[-(k - 1):]:
                    break  # crossed into the next string
                s += t[-1]
            if rng.integers(0, 2):
                s = "".join(comp[c] for c in reversed(s))
            if rng.integers(0, 4) == 0 and len(s) > 40:  # mutate one base in the middle
                p = int(rng.integers(10, len(s) - 10))
                s = s[:p] + comp[s[p]] + s[p + 1:]
        else:
            s = "".join("ACGT"[i] for i in rng.integers(0, 4, read_len))
        if r % 97 == 5:
            p = int(rng.integers(0, len(s)))
            s = s[:p] + "N" + s[p + 1:]
        if r % 131 == 7:
            s = s[: int(rng.integers(1, k))]
        if r % 53 == 3:
            s = s.lower()
        reads.append(s)
    return reads


def make_nav(name, lib):
    """navigational-query goldens (src/dictionary.cpp:112-201): kmer_neighbours for the first 400
    golden queries (positives, fwd and RC, + negatives at the tail) and string_neighbours of the
    first 40 strings, complete lookup_result records from the reference."""
    idx = os.path.join(HERE, name + ".sshash")
    d = ref.RefDictionary(idx, max_k=lib)
    z = np.load(os.path.join(HERE, name + ".npz"))
    q = z["queries"].reshape(-1, d.words)
    sel = np.concatenate([q[:300], q[-100:]]).reshape(-1)
    sid = np.arange(min(40, d.num_strings), dtype=np.uint64)
    np.savez_compressed(os.path.join(HERE, name + ".nav.npz"), kmers=sel,
                        both=d.kmer_neighbours(sel, which=3), forward=d.kmer_neighbours(sel, which=1),
                        backward=d.kmer_neighbours(sel, which=2), both_norc=d.kmer_neighbours(sel, check_rc=False),
                        string_ids=sid, strings=d.string_neighbours(sid))
    d.close()


def main():
    os.makedirs(TMP, exist_ok=True)
    if "--nav-only" in sys.argv:
        for fx in FIXTURES:
            make_nav(fx[0], fx[6])
        return
    manifest = {}
    only = [a.split("=", 1)[1] for a in sys.argv if a.startswith("--only=")]
    if only:   # regenerate just these fixtures, keep the other manifest entries
        with open(os.path.join(HERE, "manifest.json")) as f:
            manifest = json.load(f)
    for fx in FIXTURES:
        name, src, k, m, canon, nseq, lib = fx[:7]
        weighted = len(fx) > 7 and fx[7]
        if only and name not in only:
            continue
        fa = os.path.join(TMP, name + ".fa")
        subset(DATA + src, nseq, fa)
        idx = os.path.join(HERE, name + ".sshash")
        ref.build(fa, k, m, idx, canonical=canon, tmp_dir=TMP, max_k=lib, weighted=weighted)
        d = ref.RefDictionary(idx, max_k=lib)
        rng = np.random.default_rng(abs(hash(name)) % (2**32) if False else sum(map(ord, name)))
        pid, q = make_queries(d, rng, d.words)
        ids, full = d.lookup(q, check_rc=True, full=True)
        ids_norc = d.lookup(q, check_rc=False)
        assert (ids[:NQ] == pid).all(), name
        reads = synth_reads(d, rng, 300 if k > 31 else 600)
        bases = "".join(reads).encode()
        offs = np.concatenate([[0], np.cumsum([len(r) for r in reads])]).astype(np.uint64)
        sids, sfull, rep, _ = d.streaming_reads(bases, offs, full=True)
        extra = {}
        if weighted:   # dictionary::weight of the golden positives, of the first ids and of the last ones
            wid = np.concatenate([pid, np.arange(0, 3000, dtype=np.uint64),
                                  np.arange(d.num_kmers - 3000, d.num_kmers, dtype=np.uint64)])
            extra = dict(weight_ids=wid, weights=d.weight(wid))
        np.savez_compressed(
            os.path.join(HERE, name + ".npz"),
            queries=q, positive_ids=pid, ids=ids, ids_norc=ids_norc, full=full,
            read_bases=np.frombuffer(bases, dtype=np.uint8), read_offsets=offs,
            stream_ids=sids, stream_full=sfull,
            stream_report=np.array([rep[n] for n in ("num_kmers", "num_positive_kmers", "num_negative_kmers",
                                                     "num_invalid_kmers", "num_searches", "num_extensions")],
                                   dtype=np.uint64),
            **extra,
        )
        manifest[name] = dict(source=src, k=k, m=m, canonical=canon, nseq=nseq, max_k=lib, weighted=bool(weighted),
                              num_kmers=d.num_kmers, num_strings=d.num_strings,
                              index_bytes=os.path.getsize(idx), stream_report=rep,
                              num_found=int((ids != 2**64 - 1).sum()))
        print(name, manifest[name])
        d.close()
        make_nav(name, lib)
    for name, k, m, canon, lib in SYNTH:
        if only and name not in only:
            continue
        rng = np.random.default_rng(sum(map(ord, name)))
        fa = os.path.join(TMP, name + ".fa")
        strings = synth_twins_fasta(fa, rng)
        idx = os.path.join(HERE, name + ".sshash")
        ref.build(fa, k, m, idx, canonical=canon, tmp_dir=TMP, max_k=lib)
        d = ref.RefDictionary(idx, max_k=lib)
        # queries: every k-mer whose lookup does NOT return its own id (a duplicate or an rc twin answers
        # first) + 6000 others; odd positions reverse-complemented as everywhere; + random negatives
        allid = np.arange(d.num_kmers, dtype=np.uint64)
        back = d.lookup(d.access(allid), check_rc=True)
        odd = allid[back != allid]
        rest = rng.choice(allid[back == allid], 6000, replace=False).astype(np.uint64)
        pid = np.concatenate([odd, rest])
        rng.shuffle(pid)
        lo = d.access(pid)
        rcs = rc_packed(lo[1::2], np.zeros(lo[1::2].size, dtype=np.uint64), k)
        lo[1::2] = np.array(rcs, dtype=np.uint64)
        neg = rng.integers(0, 2**62, 2000).astype(np.uint64)
        q = np.concatenate([lo, neg])
        ids, full = d.lookup(q, check_rc=True, full=True)
        ids_norc = d.lookup(q, check_rc=False)
        reads = synth_twin_reads(strings, rng, 600, k)
        bases = "".join(reads).encode()
        offs = np.concatenate([[0], np.cumsum([len(r) for r in reads])]).astype(np.uint64)
        sids, sfull, rep, _ = d.streaming_reads(bases, offs, full=True)
        np.savez_compressed(
            os.path.join(HERE, name + ".npz"),
            queries=q, positive_ids=pid, ids=ids, ids_norc=ids_norc, full=full,
            read_bases=np.frombuffer(bases, dtype=np.uint8), read_offsets=offs,
            stream_ids=sids, stream_full=sfull,
            stream_report=np.array([rep[n] for n in ("num_kmers", "num_positive_kmers", "num_negative_kmers",
                                                     "num_invalid_kmers", "num_searches", "num_extensions")],
                                   dtype=np.uint64))
        manifest[name] = dict(source="synthetic (synth_twins_fasta)", k=k, m=m, canonical=canon, nseq=400, max_k=lib,
                              weighted=False, num_kmers=d.num_kmers, num_strings=d.num_strings,
                              index_bytes=os.path.getsize(idx), stream_report=rep,
                              num_found=int((ids != 2**64 - 1).sum()), distinct_kmers=False,
                              ids_not_identity=int((ids[:pid.size] != pid).sum()))
        print(name, manifest[name])
        d.close()
        make_nav(name, lib)
    with open(os.path.join(HERE, "manifest.json"), "w") as f:
        json.dump(manifest, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
