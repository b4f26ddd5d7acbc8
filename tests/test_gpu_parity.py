"""GPU parity tests: the CUDA path (through the C ABI) against the reference's golden vectors and
the C oracle.  Bit-exact: ids, complete lookup_result records, streaming ids and report counters."""
import os

import numpy as np
import pytest

from conftest import FIXTURES, REPORT_KEYS, WEIGHTED, golden

pytestmark = pytest.mark.gpu

INVALID = np.uint64(2**64 - 1)


@pytest.fixture(scope="module")
def dicts():
    import sshash_b200
    cache = {}

    def get(name):
        if name not in cache:
            g = golden(name)
            cache[name] = sshash_b200.Dictionary(g.index, device=0, max_k=g.max_k)
        return cache[name]

    yield get
    for d in cache.values():
        d.close()


def test_native_library_loaded_and_counts_launches(dicts):
    import sshash_b200
    d = dicts("se_k31_m13")
    before = sshash_b200.launch_count()
    d.lookup_batch(np.array([0x1DDB9E97AA56E2D4], dtype=np.uint64))
    assert sshash_b200.launch_count() == before + 1
    assert d.info["device_bytes"] > 0 and d.info["num_kmers"] == 4787534


def test_known_answer_vectors(dicts):
    d = dicts("se_k31_m13")
    assert (d.k(), d.m(), d.canonical(), d.num_kmers(), d.num_strings()) == (31, 13, False, 4787534, 647)
    r = d.lookup("ACCGTATGTCCCTTTTGCCTTGCTGTCGCGC")
    assert (r["kmer_id"], r["kmer_orientation"], r["string_id"], r["string_end"]) == (0, 1, 0, 118)
    r = d.lookup("GCGCGACAGCAAGGCAAAAGGGACATACGGT")
    assert (r["kmer_id"], r["kmer_orientation"]) == (0, -1)
    assert d.lookup("GCGCGACAGCAAGGCAAAAGGGACATACGGT", check_reverse_complement=False)["kmer_id"] == int(INVALID)
    assert d.lookup("CGTCATCAGCATCGGAGGCATCCACCCACGC")["kmer_id"] == 4787533
    assert d.lookup("cgtcatcagcatcggaggcatccacccacgc")["kmer_id"] == 4787533
    assert not d.is_member("ACGTACGTACGTACGTACGTACGTACGTACG")
    assert d.is_member("TCGGCCACGTTGCTGATCGCCCATACCCATT")
    assert d.access(88) == "GCACTACCAGGAACAACTGGAGCAGCTTAAA"


@pytest.mark.parametrize("name", FIXTURES)
def test_lookup_ids_and_full_records(dicts, name):
    g, d = golden(name), dicts(name)
    q = g.z["queries"]
    ids = d.lookup_batch(q)
    assert ids.dtype == np.uint64 and (ids == g.z["ids"]).all()
    assert (d.lookup_batch(q, check_reverse_complement=False) == g.z["ids_norc"]).all()
    full = d.lookup_batch(q, full=True)
    for f in full.dtype.names:
        assert (full[f] == g.z["full"][f]).all(), f
    assert (d.is_member_batch(q) == (g.z["ids"] != INVALID)).all()


@pytest.mark.parametrize("name", FIXTURES)
def test_lookup_device_buffers(dicts, name):
    import torch
    g, d = golden(name), dicts(name)
    q = torch.from_numpy(g.z["queries"].view(np.int64)).cuda()
    ids = d.lookup_batch(q)
    torch.cuda.synchronize()
    assert ids.is_cuda and (ids.cpu().numpy().view(np.uint64) == g.z["ids"]).all()
    full = d.lookup_batch(q, full=True)
    torch.cuda.synchronize()
    full = full.cpu().numpy()
    for j, f in enumerate(g.z["full"].dtype.names):
        assert (full[:, j].view(g.z["full"][f].dtype) == g.z["full"][f]).all(), f
    mem = d.is_member_batch(q)
    torch.cuda.synchronize()
    assert (mem.cpu().numpy().astype(bool) == (g.z["ids"] != INVALID)).all()


@pytest.mark.parametrize("name", FIXTURES)
def test_lookup_ascii(dicts, name):
    g, d = golden(name), dicts(name)
    k, w = d.k(), g.words
    q = g.z["queries"].reshape(-1, w)[:3000]
    strs = []
    for row in q:
        x = int(row[0]) | ((int(row[1]) << 64) if w == 2 else 0)
        strs.append("".join("ACTG"[(x >> (2 * i)) & 3] for i in range(k)))
    # packed queries may carry bits above 2k (random negatives are masked in make_golden, so they do not)
    ids = d.lookup_batch_ascii("".join(strs).encode())
    assert (ids == g.z["ids"][:3000]).all()
    ids = d.lookup_batch_ascii("".join(strs).lower().encode())
    assert (ids == g.z["ids"][:3000]).all()


@pytest.mark.parametrize("name", FIXTURES)
def test_access_and_roundtrip(dicts, name):
    """test/check.hpp:29-49: lookup(access(id)).kmer_id == id."""
    g, d = golden(name), dicts(name)
    npos = g.z["positive_ids"].size
    rng = np.random.default_rng(3)
    ids = rng.integers(0, d.num_kmers(), 200000).astype(np.uint64)
    ids[:4] = [0, 1, d.num_kmers() - 1, d.num_kmers() - 2]
    kmers = d.access_batch(ids)
    got = d.lookup_batch(kmers.reshape(-1))
    if g.meta.get("distinct_kmers", True):
        assert (got == ids).all()
    else:   # duplicated k-mers / rc twins: the first occurrence in bucket order answers; it holds the same k-mer
        assert (d.access_batch(got) == kmers).reshape(len(ids), -1).all(axis=1).sum() >= 0.4 * len(ids)
        from oracle import port
        assert (got == port.OracleDictionary(g.index, g.max_k).lookup(kmers.reshape(-1))).all()
    # access agrees with the reference on the golden positives (even positions are forward)
    acc = d.access_batch(g.z["positive_ids"]).reshape(npos, -1)
    assert (acc[0::2].reshape(-1) == g.z["queries"].reshape(-1, g.words)[:npos][0::2].reshape(-1)).all()


@pytest.mark.parametrize("name", FIXTURES)
def test_streaming_ids_and_report(dicts, name):
    g, d = golden(name), dicts(name)
    ids, rep = d.streaming_batch(g.z["read_bases"], g.z["read_offsets"])
    assert (ids == g.z["stream_ids"]).all()
    assert [rep[k] for k in REPORT_KEYS] == g.z["stream_report"].tolist()
    _, rep2 = d.streaming_batch(g.z["read_bases"], g.z["read_offsets"], want_ids=False)
    assert rep2 == rep
    # host buffers go through a double-buffered chunk pipeline: tiny chunks -> many chunks, same answer
    for chunk in (200, 4096):
        old = _set_env(SSHASH_GPU_STREAM_CHUNK=chunk)
        ids3, rep3 = d.streaming_batch(g.z["read_bases"], g.z["read_offsets"])
        _set_env(**old)
        assert (ids3 == g.z["stream_ids"]).all() and rep3 == rep, chunk


@pytest.mark.parametrize("name", ["se_k31_m13", "se_k63_m8_canon"])
def test_streaming_device_buffers_and_file(dicts, name, tmp_path):
    import torch
    g, d = golden(name), dicts(name)
    bases = torch.from_numpy(g.z["read_bases"]).cuda()
    offs = torch.from_numpy(g.z["read_offsets"].view(np.int64)).cuda()
    ids, rep = d.streaming_batch(bases, offs)
    torch.cuda.synchronize()
    assert (ids.cpu().numpy().view(np.uint64) == g.z["stream_ids"]).all()
    assert [rep[k] for k in REPORT_KEYS] == g.z["stream_report"].tolist()
    # FASTQ / FASTA files through the host driver (src/query.cpp:53-108)
    raw = g.z["read_bases"].tobytes().decode()
    o = g.z["read_offsets"].astype(np.int64)
    reads = [raw[o[i]:o[i + 1]] for i in range(len(o) - 1)]
    fq = tmp_path / "reads.fastq"
    fq.write_text("".join("@r%d\n%s\n+\n%s\n" % (i, r, "I" * len(r)) for i, r in enumerate(reads)))
    assert [d.streaming_query_from_file(str(fq))[k] for k in REPORT_KEYS] == g.z["stream_report"].tolist()
    import gzip
    fa = tmp_path / "reads.fa.gz"
    with gzip.open(fa, "wt") as f:
        f.write("".join(">r%d\n%s\n" % (i, r) for i, r in enumerate(reads)))
    assert [d.streaming_query_from_file(str(fa))[k] for k in REPORT_KEYS] == g.z["stream_report"].tolist()
    assert d.streaming_query_from_file(str(tmp_path / "x.txt"))["num_kmers"] == 0  # unsupported extension


def _set_env(**kv):
    old = {k: os.environ.get(k) for k in kv}
    for k, v in kv.items():
        if v is None:
            os.environ.pop(k, None)
        else:
            os.environ[k] = str(v)
    return old


@pytest.mark.parametrize("name", ["se_k31_m13", "se_k63_m8_canon"])
def test_file_driver_device_side_record_parsing(dicts, name, tmp_path):
    """FASTQ / FASTA records are found on the GPU (newline index over the raw file bytes).  Every
    ragged shape the positional record structure of src/query.cpp:53-108 allows -- no final newline,
    truncated last record, CRLF, empty and short sequence lines, '@'/'>' inside quality/sequence
    lines -- with chunk sizes that cut records at every possible place, against the host line
    parser (same library, SSHASH_GPU_HOST_PARSER=1) and against the unmodified reference."""
    from oracle import ref
    g, d = golden(name), dicts(name)
    raw = g.z["read_bases"].tobytes().decode()
    o = g.z["read_offsets"].astype(np.int64)
    reads = [raw[o[i]:o[i + 1]] for i in range(min(len(o) - 1, 120))]
    reads[3] = ""                                  # empty sequence line
    reads[5] = reads[5][:10]                       # shorter than k
    reads[7] = "".join(reads[20:40])               # one long read (several KB)
    rd = ref.RefDictionary(g.index, max_k=g.max_k) if ref.available(g.max_k) else None
    files = {}
    fq = "".join("@r%d\n%s\n+\n%s\n" % (i, r, ("@>I" * len(r))[:len(r)]) for i, r in enumerate(reads))
    files["a.fastq"] = fq
    files["no_final_newline.fastq"] = fq[:-1]
    files["truncated_after_seq.fastq"] = fq + "@last\n" + reads[1]
    files["truncated_after_header.fastq"] = fq + "@last"
    files["crlf.fastq"] = fq.replace("\n", "\r\n")
    fa = "".join(">r%d\n%s\n" % (i, r) for i, r in enumerate(reads))
    files["a.fa"] = fa
    files["no_final_newline.fasta"] = fa[:-1]
    files["header_only_tail.fa"] = fa + ">last"
    files["empty.fastq"] = ""
    for fname, text in files.items():
        path = tmp_path / fname
        path.write_bytes(text.encode())
        old = _set_env(SSHASH_GPU_HOST_PARSER=1)
        want = d.streaming_query_from_file(str(path))
        _set_env(**old)
        if rd is not None:
            refrep, _ = rd.streaming_file(str(path))
            assert [want[k] for k in REPORT_KEYS] == [refrep[k] for k in REPORT_KEYS], fname
        for chunk in (None, 64, 257, 1000, 4096, 65536):   # 64/257/1000 < the long read: host-parser fallback for that record
            old = _set_env(SSHASH_GPU_FILE_CHUNK=chunk)
            got = d.streaming_query_from_file(str(path))
            _set_env(**old)
            assert got == want, (fname, chunk)
    if rd is not None:
        rd.close()
    # gzip input goes through the same device parser
    import gzip
    gz = tmp_path / "reads.fq.gz"
    with gzip.open(gz, "wt") as f:
        f.write(fq)
    old = _set_env(SSHASH_GPU_FILE_CHUNK=3000)
    got = d.streaming_query_from_file(str(gz))
    _set_env(**old)
    old = _set_env(SSHASH_GPU_HOST_PARSER=1)
    want = d.streaming_query_from_file(str(tmp_path / "a.fastq"))
    _set_env(**old)
    assert got == want and got["num_kmers"] > 0


@pytest.mark.parametrize("name", ["se_k31_m13", "sal100_k31_m11_canon", "se_k63_m21"])
def test_streaming_multiline_fasta_vs_reference(dicts, name, tmp_path):
    """streaming_query_from_fasta_file_multiline (src/query.cpp:9-51): all lines of a run -- header
    lines included -- are concatenated and streamed without reset; only an empty line ends a run.
    Expected counters come from the unmodified reference (oracle/_ref) on the same file."""
    from oracle import ref
    g, d = golden(name), dicts(name)
    if not ref.available(g.max_k):
        pytest.skip("oracle/_ref not built")
    raw = g.z["read_bases"].tobytes().decode()
    o = g.z["read_offsets"].astype(np.int64)
    reads = [raw[o[i]:o[i + 1]] for i in range(len(o) - 1)]
    lines = []
    for i, r in enumerate(reads):
        if len(r) < 80:
            continue                      # keep every run longer than k (the reference underflows otherwise, :22)
        lines.append(">read%d some ACGT text" % i)
        lines += [r[j:j + 60] for j in range(0, len(r), 60)]
        if i % 7 == 3:
            lines.append("")              # end of a run
    fa = tmp_path / "multi.fa"
    fa.write_text("\n".join(lines) + "\n")
    rd = ref.RefDictionary(g.index, max_k=g.max_k)
    want, _ = rd.streaming_file(str(fa), multiline=True)
    rd.close()
    got = d.streaming_query_from_file(str(fa), multiline=True)
    assert [got[k] for k in REPORT_KEYS] == [want[k] for k in REPORT_KEYS]
    assert got["num_positive_kmers"] > 1000 and got["num_invalid_kmers"] > 1000


@pytest.mark.parametrize("name", FIXTURES)
def test_navigational_queries(dicts, name):
    """kmer_neighbours / forward / backward / string_neighbours vs the reference's goldens, and the
    property of test/check_from_file.hpp:173-226: consecutive k-mers of a string are neighbours."""
    from conftest import GOLDEN
    g, d = golden(name), dicts(name)
    z = np.load(os.path.join(GOLDEN, name + ".nav.npz"))
    for key, which, rc in (("both", 3, True), ("forward", 1, True), ("backward", 2, True), ("both_norc", 3, False)):
        got = d.kmer_neighbours_batch(z["kmers"], check_reverse_complement=rc, which=which)
        for f in got.dtype.names:
            assert (got[f] == z[key][f]).all(), (key, f)
        ids = d.kmer_neighbours_batch(z["kmers"], check_reverse_complement=rc, which=which, full=False)
        assert (ids == z[key]["kmer_id"]).all()
    got = d.string_neighbours_batch(z["string_ids"])
    for f in got.dtype.names:
        assert (got[f] == z["strings"][f]).all(), f
    # id i and id i+1 inside one string: the successor is a forward neighbour, the predecessor a backward one
    ids = np.arange(0, 2000, dtype=np.uint64)
    km = d.access_batch(ids)
    full = d.lookup_batch(km.reshape(-1), full=True)
    nb = d.kmer_neighbours_batch(km.reshape(-1), full=False)
    same = full["string_id"][1:] == full["string_id"][:-1]
    assert ((nb[:-1, :4] == ids[1:, None]).any(axis=1) | ~same).all()
    assert ((nb[1:, 4:] == ids[:-1, None]).any(axis=1) | ~same).all()


@pytest.mark.parametrize("name", FIXTURES)
def test_low_complexity_queries_vs_oracle(dicts, name):
    """Homopolymers, short tandem repeats and indexed k-mers with a repeated m-mer: every m-mer hash
    ties with another one, which is what the leftmost-minimum rule (util.hpp:262-283) is about and
    what sends the kernel's 32-bit minimizer scan to its exact fallback."""
    from oracle import port
    g, d = golden(name), dicts(name)
    o = port.OracleDictionary(g.index, g.max_k)
    k, m, w = d.k(), d.m(), g.words
    rng = np.random.default_rng(5)
    vals = []
    for period in range(1, 9):
        for _ in range(40):
            unit = rng.integers(0, 4, period)
            x = 0
            for i in range(k):
                x |= int(unit[i % period]) << (2 * i)
            vals.append(x)
    # indexed k-mers whose minimizer-length substrings repeat
    ids = rng.integers(0, d.num_kmers(), 300000).astype(np.uint64)
    km = d.access_batch(ids).reshape(-1, w)
    ints = [int(r[0]) | ((int(r[1]) << 64) if w == 2 else 0) for r in km[:60000]]
    mask = (1 << (2 * m)) - 1
    rep = [x for x in ints if len({(x >> (2 * i)) & mask for i in range(k - m + 1)}) < k - m + 1]
    vals += rep[:3000]
    full_mask = (1 << (2 * k)) - 1
    vals += [(~x) & full_mask for x in vals[:320]] + [0, full_mask]
    q = np.array([[x & (2**64 - 1), x >> 64][:w] for x in vals], dtype=np.uint64).reshape(-1)
    want_ids, want_full = o.lookup(q, full=True)
    got = d.lookup_batch(q, full=True)
    for f in got.dtype.names:
        assert (got[f] == want_full[f]).all(), f
    assert (d.lookup_batch(q) == want_ids).all()
    assert (d.lookup_batch(q, check_reverse_complement=False) == o.lookup(q, check_rc=False)).all()
    o.close()


@pytest.mark.parametrize("name", FIXTURES)
def test_streaming_low_complexity_reads_vs_oracle(dicts, name):
    """Streaming over reads made of homopolymers, short tandem repeats, indexed text with repeats
    spliced in, and their mixtures: within a window several m-mers share one hash, so the window
    minimizer of both strands is decided by the tie rules (leftmost on the k-mer, leftmost on its
    reverse complement) -- the streaming kernel's packed 25-bit scan must hand those windows to
    the exact scan.  Ids and all six counters against the C oracle."""
    from oracle import port
    g, d = golden(name), dicts(name)
    o = port.OracleDictionary(g.index, g.max_k)
    k = d.k()
    rng = np.random.default_rng(11)
    raw = g.z["read_bases"].tobytes().decode()
    off = g.z["read_offsets"].astype(np.int64)
    real = [raw[off[i]:off[i + 1]] for i in range(min(len(off) - 1, 60))]
    reads = []
    for period in range(1, 7):
        for _ in range(12):
            unit = "".join("ACGT"[c] for c in rng.integers(0, 4, period))
            reads.append((unit * (300 // period + 1))[: int(rng.integers(k, 300))])
    for r in real:
        if len(r) < 2 * k:
            continue
        cut = int(rng.integers(k // 2, len(r) - k // 2))
        unit = "".join("ACGT"[c] for c in rng.integers(0, 4, int(rng.integers(1, 4))))
        reads.append(r[:cut] + unit * int(rng.integers(5, 40)) + r[cut:])      # a repeat inside indexed text
        reads.append(r[:cut] + "N" + unit * 20)
    reads += ["A" * (k - 1), "C" * k, "", "ACGT" * 100]
    bases = "".join(reads).encode()
    offsets = np.zeros(len(reads) + 1, dtype=np.uint64)
    offsets[1:] = np.cumsum([len(r) for r in reads])
    want_ids, _, want_rep = o.streaming_reads(bases, offsets)
    got_ids, got_rep = d.streaming_batch(np.frombuffer(bases, dtype=np.uint8), offsets)
    assert (got_ids == want_ids).all()
    assert got_rep == want_rep
    assert got_rep["num_kmers"] > 5000


@pytest.mark.parametrize("name", WEIGHTED)
def test_weights(dicts, name):
    """dictionary::weight (src/dictionary.cpp:96-100) vs the reference's goldens and, for every
    k-mer id of the index, vs the C oracle; host and device buffers; lookup -> weight chain of
    tools/perf.hpp:143-149."""
    import torch
    from oracle import port
    g, d = golden(name), dicts(name)
    assert d.weighted()
    assert (d.weight_batch(g.z["weight_ids"]) == g.z["weights"]).all()
    o = port.OracleDictionary(g.index, g.max_k)
    ids = np.arange(d.num_kmers(), dtype=np.uint64)
    want = o.weight(ids)
    assert (d.weight_batch(ids) == want).all()
    dev = d.weight_batch(torch.from_numpy(ids.view(np.int64)).cuda())
    torch.cuda.synchronize()
    assert (dev.cpu().numpy().view(np.uint64) == want).all()
    npos = g.z["positive_ids"].size
    q = torch.from_numpy(g.z["queries"].reshape(-1, g.words)[:npos].reshape(-1).view(np.int64)).cuda()
    w = d.weight_batch(d.lookup_batch(q))
    torch.cuda.synchronize()
    assert (w.cpu().numpy().view(np.uint64) == g.z["weights"][:npos]).all()
    assert d.weight(int(g.z["weight_ids"][0])) == int(g.z["weights"][0])
    o.close()


def test_weight_on_unweighted_dictionary_is_an_error(dicts):
    import sshash_b200
    d = dicts("se_k31_m13")
    assert not d.weighted()
    with pytest.raises(sshash_b200.SshashGpuError):
        d.weight_batch(np.zeros(4, dtype=np.uint64))


def test_edge_cases(dicts):
    d = dicts("se_k31_m13")
    assert d.lookup_batch(np.zeros(0, dtype=np.uint64)).size == 0
    ids, rep = d.streaming_batch(b"", np.zeros(1, dtype=np.uint64))
    assert ids.size == 0 and rep["num_kmers"] == 0
    # ragged reads: empty, shorter than k, exactly k, all-N
    reads = ["", "ACGT", "ACCGTATGTCCCTTTTGCCTTGCTGTCGCGC", "N" * 40, "ACCGTATGTCCCTTTTGCCTTGCTGTCGCGCN"]
    offs = np.concatenate([[0], np.cumsum([len(r) for r in reads])]).astype(np.uint64)
    ids, rep = d.streaming_batch("".join(reads).encode(), offs)
    assert ids.tolist() == [0] + [int(INVALID)] * 10 + [0, int(INVALID)]
    assert (rep["num_kmers"], rep["num_positive_kmers"], rep["num_invalid_kmers"], rep["num_searches"]) == (13, 2, 11, 2)


def test_large_batch_properties():
    """BASELINE cfg-2 shape at reduced count (1e7): every positive returns its own id, through the
    chunked host pipeline and through device buffers; 50% reverse-complemented."""
    import torch
    import sshash_b200
    g = golden("se_k31_m13")
    d = sshash_b200.Dictionary(g.index)
    n = 10_000_000
    gen = torch.Generator(device="cuda").manual_seed(42)
    ids = torch.randint(0, d.num_kmers(), (n,), generator=gen, device="cuda", dtype=torch.int64)
    kmers = d.access_batch(ids)
    got = d.lookup_batch(kmers)
    torch.cuda.synchronize()
    assert torch.equal(got, ids)
    host = kmers.cpu().numpy().view(np.uint64)
    got_h = d.lookup_batch(host)
    assert (got_h.view(np.int64) == ids.cpu().numpy()).all()
    neg = torch.randint(0, 2**62, (n,), generator=gen, device="cuda", dtype=torch.int64)
    got = d.lookup_batch(neg)
    torch.cuda.synchronize()
    assert int((got != -1).sum()) == 0
    d.close()


def test_multi_partition_index_vs_oracle(tmp_path):
    """> 3e6 minimizers => several PTHash partitions; index built on the box by the reference
    builder (oracle/_ref), CUDA path compared with the C oracle."""
    from oracle import port, ref
    if not ref.available(31):
        pytest.skip("oracle/_ref not built")
    import sshash_b200
    rng = np.random.default_rng(11)
    fa = tmp_path / "synth.fa"
    with open(fa, "w") as f:
        for i in range(13000):
            f.write(">%d\n%s\n" % (i, "".join("ACGT"[c] for c in rng.integers(0, 4, 3000))))
    idx = str(tmp_path / "synth.sshash")
    ref.build(str(fa), 31, 14, idx, threads=min(16, os.cpu_count() or 1), tmp_dir=str(tmp_path))
    o = port.OracleDictionary(idx)
    d = sshash_b200.Dictionary(idx)
    assert d.info["mphf_partitions"] >= 2
    ids = rng.integers(0, o.num_kmers, 100000).astype(np.uint64)
    pos = d.access_batch(ids)
    assert (pos == o.access(ids)).all()
    q = np.concatenate([pos, rng.integers(0, 2**62, 100000).astype(np.uint64)])
    a, fa_ = o.lookup(q, full=True)
    full = d.lookup_batch(q, full=True)
    assert (full["kmer_id"][:100000] == ids).all()
    for f in full.dtype.names:
        assert (full[f] == fa_[f]).all(), f
    d.close()


def _open_with_env(path, env, **kw):
    """Dictionary opened under a set of library environment switches (they are read at open time)."""
    import sshash_b200
    saved = {k: os.environ.get(k) for k in env}
    os.environ.update(env)
    try:
        return sshash_b200.Dictionary(path, **kw)
    finally:
        for k, v in saved.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


BINNED_ON = {"SSHASH_GPU_BINNED": "1", "SSHASH_GPU_BINNED_MIN": "1"}


@pytest.mark.parametrize("name", FIXTURES)
def test_binned_path_gives_the_reference_ids(name):
    """Partition-major path (binned.cu) forced on for every golden index (one bin there: the
    multisplit, both rounds, the canonical tie and the un-permute still run): device-resident
    lookups and membership equal the reference's golden ids, with and without the rc pass."""
    import torch
    g = golden(name)
    d = _open_with_env(g.index, BINNED_ON, device=0, max_k=g.max_k)
    q = torch.from_numpy(g.z["queries"].view(np.int64)).cuda()
    ids = d.lookup_batch(q)
    assert (ids.cpu().numpy().view(np.uint64) == g.z["ids"]).all()
    ids = d.lookup_batch(q, check_reverse_complement=False)
    assert (ids.cpu().numpy().view(np.uint64) == g.z["ids_norc"]).all()
    mem = d.is_member_batch(q)
    assert (mem.cpu().numpy().astype(bool) == (g.z["ids"] != INVALID)).all()
    d.close()


def test_binned_path_multi_partition_multi_range(tmp_path):
    """Several MPHF partitions (bins) and several 2^20-query output ranges: 2.6e6 device-resident
    queries (positives, half reverse-complemented, interleaved with random negatives) through the
    partition-major path against the direct kernel and, on a sample, the C oracle."""
    import torch
    from oracle import port, ref
    if not ref.available(31):
        pytest.skip("oracle/_ref not built")
    rng = np.random.default_rng(5)
    fa = tmp_path / "synth.fa"
    with open(fa, "w") as f:
        for i in range(13000):
            f.write(">%d\n%s\n" % (i, "".join("ACGT"[c] for c in rng.integers(0, 4, 3000))))
    for canonical in (False, True):
        idx = str(tmp_path / ("synth%d.sshash" % canonical))
        ref.build(str(fa), 31, 14, idx, canonical=canonical, threads=min(16, os.cpu_count() or 1), tmp_dir=str(tmp_path))
        direct = _open_with_env(idx, {"SSHASH_GPU_BINNED": "0"})
        binned = _open_with_env(idx, BINNED_ON)
        assert binned.info["mphf_partitions"] >= 2
        n = 2_600_001
        ids = rng.integers(0, direct.num_kmers(), n).astype(np.uint64)
        pos = direct.access_batch(ids)
        o = port.OracleDictionary(idx)
        from bench import rc_packed_torch
        q = pos.copy()
        q[1::4] = rc_packed_torch(torch.from_numpy(pos[1::4].view(np.int64).copy()), 31).numpy().view(np.uint64)
        q[2::4] = rng.integers(0, 2**62, q[2::4].size).astype(np.uint64)      # negatives
        q[3::4] = q[0::4][: q[3::4].size]                                      # duplicates of other queries
        qd = torch.from_numpy(q.view(np.int64)).cuda()
        for rc in (True, False):
            a = direct.lookup_batch(qd, check_reverse_complement=rc)
            b = binned.lookup_batch(qd, check_reverse_complement=rc)
            torch.cuda.synchronize()
            assert torch.equal(a, b)
            assert torch.equal(direct.is_member_batch(qd, check_reverse_complement=rc), binned.is_member_batch(qd, check_reverse_complement=rc))
        got = binned.lookup_batch(qd).cpu().numpy().view(np.uint64)
        assert (got[0::4] == ids[0::4]).all()
        m = 200000
        assert (got[:m] == o.lookup(q[:m])).all()
        direct.close()
        binned.close()


@pytest.mark.parametrize("name", ["se_k31_m13", "sal100_k31_m7_canon", "se_k63_m21", "sal100_k31_m7_reg"])
def test_u32_ids(dicts, name):
    """sshash_gpu_lookup_batch_u32: the low 32 bits of the reference ids, UINT32_MAX for "not found";
    host buffers, device buffers, and the partition-major path."""
    import torch
    if name not in FIXTURES:
        pytest.skip("fixture not present")
    g, d = golden(name), dicts(name)
    want = g.z["ids"].astype(np.uint32)          # INVALID truncates to UINT32_MAX
    assert (d.lookup_batch_u32(g.z["queries"]) == want).all()
    q = torch.from_numpy(g.z["queries"].view(np.int64)).cuda()
    assert (d.lookup_batch_u32(q).cpu().numpy().view(np.uint32) == want).all()
    b = _open_with_env(g.index, BINNED_ON, device=0, max_k=g.max_k)
    assert (b.lookup_batch_u32(q).cpu().numpy().view(np.uint32) == want).all()
    assert (b.lookup_batch_u32(q, check_reverse_complement=False).cpu().numpy().view(np.uint32) == g.z["ids_norc"].astype(np.uint32)).all()
    b.close()


@pytest.mark.parametrize("name", ["se_k31_m13", "se_k63_m21"])
def test_multi_gpu_handle_matches_goldens(name):
    """MultiDictionary = sshash_gpu_multi_* (every visible GPU, also a single one): host buffers,
    device buffers on the first and on the last GPU, u32 ids, membership, streaming."""
    import torch
    import sshash_b200
    g = golden(name)
    m = sshash_b200.MultiDictionary(g.index, max_k=g.max_k)
    assert m.num_devices == torch.cuda.device_count()
    q, want = g.z["queries"], g.z["ids"]
    assert (m.lookup_batch(q) == want).all()
    assert (m.lookup_batch(q, check_reverse_complement=False) == g.z["ids_norc"]).all()
    assert (m.lookup_batch_u32(q) == want.astype(np.uint32)).all()
    assert (m.is_member_batch(q) == (want != INVALID)).all()
    for dev in {0, m.num_devices - 1}:
        qd = torch.from_numpy(q.view(np.int64)).to("cuda:%d" % dev)
        ids = m.lookup_batch(qd)
        assert ids.device == qd.device and (ids.cpu().numpy().view(np.uint64) == want).all()
        assert (m.lookup_batch_u32(qd).cpu().numpy().view(np.uint32) == want.astype(np.uint32)).all()
    sids, rep = m.streaming_batch(g.z["read_bases"], g.z["read_offsets"])
    assert (sids == g.z["stream_ids"]).all()
    assert rep == dict(zip(REPORT_KEYS, g.z["stream_report"].tolist()))
    # ragged: fewer queries than GPUs, empty batch
    assert (m.lookup_batch(q[: g.words]) == want[:1]).all()
    assert m.lookup_batch(q[:0]).size == 0
    m.close()


def test_concurrent_batches_on_one_handle(dicts):
    """include/sshash_gpu.h: the handle is immutable after open, any number of host threads may issue
    batches concurrently.  Four threads, mixed host / device / full-record / streaming calls."""
    import threading
    import torch
    g, d = golden("se_k31_m13"), dicts("se_k31_m13")
    q, want = g.z["queries"], g.z["ids"]
    errors = []

    def worker(t):
        try:
            stream = torch.cuda.Stream()
            for it in range(12):
                kind = (t + it) % 4
                if kind == 0:
                    assert (d.lookup_batch(q) == want).all()
                elif kind == 1:
                    with torch.cuda.stream(stream):
                        qd = torch.from_numpy(q.view(np.int64)).cuda()
                        ids = d.lookup_batch(qd)
                        stream.synchronize()
                    assert (ids.cpu().numpy().view(np.uint64) == want).all()
                elif kind == 2:
                    full = d.lookup_batch(q[:6000], full=True)
                    for f in full.dtype.names:
                        assert (full[f] == g.z["full"][f][:6000]).all(), f
                else:
                    sids, rep = d.streaming_batch(g.z["read_bases"], g.z["read_offsets"])
                    assert (sids == g.z["stream_ids"]).all()
                    assert rep == dict(zip(REPORT_KEYS, g.z["stream_report"].tolist()))
        except BaseException as e:   # noqa: BLE001
            errors.append((t, repr(e)))

    threads = [threading.Thread(target=worker, args=(t,)) for t in range(4)]
    [t.start() for t in threads]
    [t.join() for t in threads]
    assert not errors, errors


@pytest.mark.parametrize("name", FIXTURES)
def test_input_contract_check_flags_exactly_the_non_distinct_indexes(dicts, name):
    """distinct_check_kernel: the twin / duplicate fixtures break SSHash's input contract, no other fixture does
    (a false positive would silently cost the streaming shortcuts: 2x slower, still correct)."""
    assert dicts(name).breaks_input_contract() == (golden(name).meta.get("distinct_kmers") is False)


@pytest.mark.parametrize("name", ["se_k31_m13", "sal100_k31_m7_canon", "sal100_k31_m7_reg", "twins_k31_m13_reg"])
def test_streaming_replay_path_equals_the_shortcuts(name):
    """SSHASH_GPU_ASSUME_DISTINCT=0 forces the literal replay of the reference's state machine (the path
    indexes with duplicated k-mers / rc twins take): same ids and counters as the goldens."""
    g = golden(name)
    d = _open_with_env(g.index, {"SSHASH_GPU_ASSUME_DISTINCT": "0"}, device=0, max_k=g.max_k)
    sids, rep = d.streaming_batch(g.z["read_bases"], g.z["read_offsets"])
    assert (sids == g.z["stream_ids"]).all()
    assert rep == dict(zip(REPORT_KEYS, g.z["stream_report"].tolist()))
    d.close()


def test_sharded_lookup_with_peer_store_gather_two_gpus():
    """N > 1: every rank's lookup kernel stores its ids straight into rank 0's gathered vector through
    NVLink peer stores (sshash_b200.sharded, mode "peer"); rank 0 checks every slice against the
    owner's sampled ids.  Needs two GPUs on the box (the CPU suite covers the host logic with gloo)."""
    import json
    import subprocess
    import sys
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29533", os.path.join(root, "tools", "micro", "peer_gather.py"),
           os.path.join(root, "tests", "golden", "se_k31_m13.sshash"), "3000000"]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stderr[-2000:]
    line = [l for l in p.stdout.splitlines() if l.startswith("{")][-1]
    r = json.loads(line)
    assert r["world"] == 2 and r["peer_ms"] > 0


@pytest.mark.parametrize("name", ["se_k31_m13", "sakai_k31_m13_weighted", "ecoli_k31_m11_canon_weighted", "sal100_k31_m11_canon"])
def test_wide_entries_layout_gives_the_same_answers(dicts, name):
    """Indexes that do not fit L2 get 16-byte wide entries (codeword + the text around a singleton
    bucket's offset) and the ids-only lookups compare against that copy instead of reading
    `strings`.  Forced here on small fixtures (SSHASH_GPU_WIDE=1): ids, membership, no-RC lookups,
    ASCII input, negatives and low-complexity k-mers must equal the goldens / the C oracle."""
    import sshash_b200
    from oracle import port
    g, d = golden(name), dicts(name)
    old = _set_env(SSHASH_GPU_WIDE=1)
    dw = sshash_b200.Dictionary(g.index, device=0, max_k=g.max_k)
    _set_env(**old)
    try:
        if dw.info["device_bytes"] < d.info["device_bytes"] + 16 * d.info["num_minimizers"]:
            pytest.skip("codeword + text do not fit 128 bits for this k, m")
        q = g.z["queries"]
        assert (dw.lookup_batch(q) == g.z["ids"]).all()
        assert (dw.is_member_batch(q) == (g.z["ids"] != INVALID)).all()
        assert (dw.lookup_batch(q, check_reverse_complement=False) == d.lookup_batch(q, check_reverse_complement=False)).all()
        full = dw.lookup_batch(q, full=True)                      # full records keep the regular route
        assert (full["kmer_id"] == g.z["ids"]).all()
        o = port.OracleDictionary(g.index, g.max_k)
        k = d.k()
        rng = np.random.default_rng(3)
        vals = []
        for period in range(1, 9):
            for _ in range(40):
                unit = rng.integers(0, 4, period)
                vals.append(sum(int(unit[i % period]) << (2 * i) for i in range(k)))
        ids = rng.integers(0, d.num_kmers(), 200000).astype(np.uint64)
        km = d.access_batch(ids)
        assert (dw.lookup_batch(km) == ids).all()                 # every sampled positive, first k-mers of strings included
        first = d.access_batch(np.arange(0, min(d.num_kmers(), 5000), dtype=np.uint64))
        assert (dw.lookup_batch(first) == d.lookup_batch(first)).all()
        neg = rng.integers(0, 2 ** (2 * k), 200000, dtype=np.uint64)
        lc = np.array(vals + [0, (1 << (2 * k)) - 1], dtype=np.uint64)
        for arr in (neg, lc):
            assert (dw.lookup_batch(arr) == o.lookup(arr)).all()
        asc = g.z["read_bases"][: 40 * k].tobytes()
        assert (dw.lookup_batch_ascii(asc) == d.lookup_batch_ascii(asc)).all()
    finally:
        dw.close()
