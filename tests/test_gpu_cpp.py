"""GPU test of the C++ host layer: the header-only dictionary wrapper (sshash_b200/csrc/dictionary.hpp)
compiled with g++ against libsshash_gpu.so, run like the reference's own check loop."""
import os
import subprocess

import pytest

from conftest import ROOT, golden

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", ["se_k31_m13", "se_k63_m21"])
def test_cpp_example(tmp_path, name):
    exe = str(tmp_path / "lookup_example")
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    lib_dir = os.path.join(ROOT, "sshash_b200")
    subprocess.check_call([cxx, "-std=c++17", "-O2", os.path.join(ROOT, "examples", "lookup_example.cpp"), "-o", exe,
                           os.path.join(lib_dir, "libsshash_gpu.so"), "-Wl,-rpath," + lib_dir])
    out = subprocess.run([exe, golden(name).index, "200000"], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "0 mismatches" in out.stdout


@pytest.mark.parametrize("name", ["se_k31_m13", "se_k63_m21"])
def test_c_multi_gpu_example(tmp_path, name):
    """examples/multi_gpu_example.c: plain C through the multi-GPU entry points of the C ABI
    (sshash_gpu_multi_*), every visible GPU, host buffers; self-checking."""
    exe = str(tmp_path / "multi_gpu_example")
    cc = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else "gcc"
    lib_dir = os.path.join(ROOT, "sshash_b200")
    subprocess.check_call([cc, "-std=c11", "-O2", os.path.join(ROOT, "examples", "multi_gpu_example.c"), "-o", exe,
                           os.path.join(lib_dir, "libsshash_gpu.so"), "-Wl,-rpath," + lib_dir])
    out = subprocess.run([exe, golden(name).index, "300001"], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "0 mismatches" in out.stdout.splitlines()[-1]


def test_cpp_query_tool_prints_the_reference_report(tmp_path):
    """examples/query_example.cpp = the reference's `sshash query` (tools/query.cpp:5-70): report
    lines on stdout, one json line on stderr, counters equal to the reference's golden report."""
    import json
    import numpy as np
    from conftest import REPORT_KEYS
    g = golden("se_k31_m13")
    exe = str(tmp_path / "query_example")
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    lib_dir = os.path.join(ROOT, "sshash_b200")
    subprocess.check_call([cxx, "-std=c++17", "-O2", os.path.join(ROOT, "examples", "query_example.cpp"), "-o", exe,
                           os.path.join(lib_dir, "libsshash_gpu.so"), "-Wl,-rpath," + lib_dir])
    raw = g.z["read_bases"].tobytes().decode()
    o = g.z["read_offsets"].astype(np.int64)
    fq = tmp_path / "reads.fastq"
    fq.write_text("".join("@r%d\n%s\n+\n%s\n" % (i, raw[o[i]:o[i + 1]], "I" * int(o[i + 1] - o[i])) for i in range(len(o) - 1)))
    out = subprocess.run([exe, "-i", g.index, "-q", str(fq)], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    want = dict(zip(REPORT_KEYS, g.z["stream_report"].tolist()))
    assert out.stdout.startswith("==== query report:\nnum_kmers = %d\n" % want["num_kmers"])
    assert "num_searches = %d/%d (" % (want["num_searches"], want["num_positive_kmers"]) in out.stdout
    line = json.loads([l for l in out.stderr.splitlines() if l.startswith("{")][-1])
    assert {k: int(line[k]) for k in REPORT_KEYS} == want
    assert line["index_filename"] == g.index and "elapsed_millisec" in line
