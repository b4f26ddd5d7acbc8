"""CPU tests: the C oracle (oracle/sshash_oracle.c) against the reference's own answers.

Pins the oracle against (a) the literal known-answer vectors of SURVEY.md 8c, (b) golden vectors
produced by the unmodified reference (tests/golden/make_golden.py), (c) the size-independent
properties the reference's checkers use (test/check.hpp:29-49: lookup(access(id)) == id), and,
when /root/reference is present, (d) the README Example 2/3 report and a multi-partition index.
"""
import os

import numpy as np
import pytest

from conftest import FIXTURES, NON_DISTINCT, REPORT_KEYS, WEIGHTED, golden
from oracle import port, ref

INVALID = np.uint64(2**64 - 1)


@pytest.fixture(scope="module")
def oracles():
    cache = {}

    def get(name):
        if name not in cache:
            g = golden(name)
            cache[name] = port.OracleDictionary(g.index, max_k=g.max_k)
        return cache[name]

    return get


def test_known_answer_vectors(oracles):
    """bundled S. enterica k=31 m=13."""
    o = oracles("se_k31_m13")
    assert (o.num_kmers, o.num_strings, o.k, o.m, o.canonical) == (4787534, 647, 31, 13, 0)
    assert o.magic == 0x9F29CB17A2A49995 and o.mphf_seed == 1234567890
    table = [
        ("ACCGTATGTCCCTTTTGCCTTGCTGTCGCGC", 0x1DDB9E97AA56E2D4, 0, 1),
        ("GCGCGACAGCAAGGCAAAAGGGACATACGGT", 0x2F484FC01F071377, 0, -1),
        ("TCGGCCACGTTGCTGATCGCCCATACCCATT", 0x2854857639EB45F6, 87, 1),
        ("GCACTACCAGGAACAACTGGAGCAGCTTAAA", 0x00A71CF9043C5247, 88, 1),
        ("CGCGTCGCGGGCGCTGGATAACTTTCTGGCG", 0x37E6A423E77F76DD, 265, 1),
        ("CCGCCTCGTCATCAGCATCGGAGGCATCCAC", 0x1161F3D87186D975, 4787527, 1),
        ("CGTCATCAGCATCGGAGGCATCCACCCACGC", 0x1D15161F3D87186D, 4787533, 1),
    ]
    packed = np.array([t[1] for t in table] + [0x34B4B4B4B4B4B4B4], dtype=np.uint64)
    ids, full = o.lookup(packed, full=True)
    assert ids.tolist() == [t[2] for t in table] + [int(INVALID)]
    assert full["kmer_orientation"][:7].tolist() == [t[3] for t in table]
    ascii_ids = o.lookup_ascii("".join(t[0] for t in table).encode() + b"ACGTACGTACGTACGTACGTACGTACGTACG")
    assert (ascii_ids == ids).all()
    # lower-case input gives the same (include/kmer.hpp:194)
    assert (o.lookup_ascii(table[5][0].lower().encode()) == [4787527]).all()


@pytest.mark.parametrize("name", FIXTURES)
def test_lookup_matches_reference_golden(oracles, name):
    g, o = golden(name), oracles(name)
    ids, full = o.lookup(g.z["queries"], full=True)
    assert (ids == g.z["ids"]).all()
    for f in full.dtype.names:
        assert (full[f] == g.z["full"][f]).all(), f
    assert (o.lookup(g.z["queries"], check_rc=False) == g.z["ids_norc"]).all()
    npos = g.z["positive_ids"].size
    if g.meta.get("distinct_kmers", True):
        assert (ids[:npos] == g.z["positive_ids"]).all()
    else:   # duplicates / rc twins: another occurrence may answer, but it holds the same k-mer
        assert (ids[:npos] != np.uint64(2**64 - 1)).all() and (ids[:npos] != g.z["positive_ids"]).any()


@pytest.mark.parametrize("name", FIXTURES)
def test_streaming_matches_reference_golden(oracles, name):
    g, o = golden(name), oracles(name)
    ids, full, rep = o.streaming_reads(g.z["read_bases"].tobytes(), g.z["read_offsets"], full=True)
    assert (ids == g.z["stream_ids"]).all()
    for f in full.dtype.names:
        assert (full[f] == g.z["stream_full"][f]).all(), f
    assert [rep[k] for k in REPORT_KEYS] == g.z["stream_report"].tolist()


@pytest.mark.parametrize("name", [n for n in FIXTURES if n not in NON_DISTINCT])
def test_lookup_access_roundtrip(oracles, name):
    """test/check.hpp:29-49: lookup(access(id)).kmer_id == id, forward orientation."""
    o = oracles(name)
    rng = np.random.default_rng(7)
    ids = rng.integers(0, o.num_kmers, 20000).astype(np.uint64)
    got, full = o.lookup(o.access(ids), full=True)
    assert (got == ids).all()
    assert (full["kmer_orientation"] == 1).all()
    assert (full["kmer_id_in_string"] == full["kmer_offset"] - full["string_begin"]).all()


@pytest.mark.parametrize("name", FIXTURES)
def test_navigational_queries_match_reference_golden(oracles, name):
    """kmer_neighbours / forward / backward / string_neighbours (src/dictionary.cpp:112-201)."""
    import os
    from conftest import GOLDEN
    o = oracles(name)
    z = np.load(os.path.join(GOLDEN, name + ".nav.npz"))
    for key, which, rc in (("both", 3, True), ("forward", 1, True), ("backward", 2, True), ("both_norc", 3, False)):
        got = o.kmer_neighbours(z["kmers"], check_rc=rc, which=which)
        for f in got.dtype.names:
            assert (got[f] == z[key][f]).all(), (key, f)
    got = o.string_neighbours(z["string_ids"])
    for f in got.dtype.names:
        assert (got[f] == z["strings"][f]).all(), f


@pytest.mark.parametrize("name", WEIGHTED)
def test_weights_match_reference_golden(oracles, name):
    """dictionary::weight (src/dictionary.cpp:96-100): golden weights come from the reference."""
    g, o = golden(name), oracles(name)
    assert o.weighted()
    assert (o.weight(g.z["weight_ids"]) == g.z["weights"]).all()
    assert np.unique(g.z["weights"]).size > 3


def test_unweighted_index_reports_it(oracles):
    assert not oracles("se_k31_m13").weighted()


@pytest.mark.parametrize("name", ["se_k31_m13", "sal100_k31_m11_canon", "se_k63_m21"])
def test_multiline_fasta_is_one_read_per_run(oracles, name, tmp_path):
    """Restatement of streaming_query_from_fasta_file_multiline (src/query.cpp:9-51) used by the GPU
    file driver: every run of non-empty lines (headers included) is ONE read.  Checked against the
    unmodified reference on the same file."""
    g, o = golden(name), oracles(name)
    if not ref.available(g.max_k):
        pytest.skip("oracle/_ref not built")
    raw = g.z["read_bases"].tobytes().decode()
    off = g.z["read_offsets"].astype(np.int64)
    reads = [raw[off[i]:off[i + 1]] for i in range(len(off) - 1)]
    lines, runs, cur = [], [], ""
    for i, r in enumerate(reads):
        if len(r) < 80:
            continue
        rec = [">read%d some ACGT text" % i] + [r[j:j + 60] for j in range(0, len(r), 60)]
        lines += rec
        cur += "".join(rec)
        if i % 7 == 3:
            lines.append("")
            runs.append(cur)
            cur = ""
    if cur:
        runs.append(cur)
    fa = tmp_path / "multi.fa"
    fa.write_text("\n".join(lines) + "\n")
    rd = ref.RefDictionary(g.index, max_k=g.max_k)
    want, _ = rd.streaming_file(str(fa), multiline=True)
    rd.close()
    offs = np.concatenate([[0], np.cumsum([len(r) for r in runs])]).astype(np.uint64)
    _, _, rep = o.streaming_reads("".join(runs).encode(), offs)
    assert [rep[k] for k in REPORT_KEYS] == [want[k] for k in REPORT_KEYS]


def test_bad_files(tmp_path):
    p = tmp_path / "bad.sshash"
    p.write_bytes(b"\x04\x01\x01" + b"\0" * 100)
    with pytest.raises(RuntimeError, match="MAJOR index version mismatch"):
        port.OracleDictionary(str(p))
    good = open(golden("se_k47_m8").index, "rb").read()
    p.write_bytes(good[: len(good) // 2])
    with pytest.raises(RuntimeError, match="malformed"):
        port.OracleDictionary(str(p), max_k=63)
    with pytest.raises(RuntimeError):
        port.OracleDictionary(str(tmp_path / "missing.sshash"))


needs_ref = pytest.mark.skipif(not (ref.available(31) and os.path.isdir("/root/reference/data")),
                               reason="needs oracle/_ref and /root/reference (build container only)")


@needs_ref
def test_readme_example_report(tmp_path):
    """README.md:216-226: salmonella_100 m=15 regular and m=13 canonical vs SRR5833294.10K.fastq.gz."""
    data = "/root/reference/data/"
    fq = data + "queries/SRR5833294.10K.fastq.gz"
    import gzip
    lines = gzip.open(fq, "rt").read().split("\n")
    reads = lines[1::4]
    bases = "".join(reads).encode()
    offs = np.concatenate([[0], np.cumsum([len(r) for r in reads])]).astype(np.uint64)
    for m, canon in ((15, False), (13, True)):
        idx = str(tmp_path / ("s100_%d.sshash" % m))
        ref.build(data + "unitigs_stitched/salmonella_100_k31_ust.fa.gz", 31, m, idx, canonical=canon,
                  threads=4, tmp_dir=str(tmp_path))
        o = port.OracleDictionary(idx)
        _, _, rep = o.streaming_reads(bases, offs)
        assert (rep["num_kmers"], rep["num_positive_kmers"], rep["num_searches"], rep["num_extensions"]) == \
            (460000, 46, 42, 4)
        assert (rep["num_negative_kmers"], rep["num_invalid_kmers"]) == (459097, 857)
        rd = ref.RefDictionary(idx)
        rrep, _ = rd.streaming_file(fq)
        assert rrep == rep


@needs_ref
def test_multi_partition_mphf_against_reference(tmp_path):
    """> 3e6 minimizers => several PTHash partitions (constants.hpp:10-11)."""
    rng = np.random.default_rng(11)
    fa = tmp_path / "synth.fa"
    with open(fa, "w") as f:
        for i in range(13000):
            f.write(">%d\n%s\n" % (i, "".join("ACGT"[c] for c in rng.integers(0, 4, 3000))))
    idx = str(tmp_path / "synth.sshash")
    ref.build(str(fa), 31, 14, idx, threads=8, tmp_dir=str(tmp_path))
    o = port.OracleDictionary(idx)
    assert o.mphf_partitions >= 2
    rd = ref.RefDictionary(idx)
    ids = rng.integers(0, o.num_kmers, 30000).astype(np.uint64)
    q = np.concatenate([rd.access(ids), rng.integers(0, 2**62, 30000).astype(np.uint64)])
    a, fa_ = rd.lookup(q, full=True)
    b, fb = o.lookup(q, full=True)
    assert (a == b).all() and (a[:30000] == ids).all()
    for f in fa_.dtype.names:
        assert (fa_[f] == fb[f]).all(), f
