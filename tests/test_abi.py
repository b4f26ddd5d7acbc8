"""CPU tests of the boundary: the C-ABI library loads, exports every symbol include/sshash_gpu.h
declares, and fails loudly (no CPU fallback) when there is no GPU."""
import os
import re

import pytest

from conftest import ROOT, golden


def header_symbols():
    text = open(os.path.join(ROOT, "include", "sshash_gpu.h")).read()
    return sorted(set(re.findall(r"SSHASH_GPU_API[^;(]*?\b(sshash_gpu_\w+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from sshash_b200 import _lib
    lib = _lib.lib()
    syms = header_symbols()
    assert len(syms) >= 12
    for s in syms:
        assert hasattr(lib, s), s
    assert set(syms) == set(_lib.SYMBOLS), "python binding and header disagree"
    assert b"sm_100a" in lib.sshash_gpu_build_info()


def test_library_is_sm100a_only():
    import subprocess
    from sshash_b200 import _lib
    out = subprocess.run(["cuobjdump", "-lelf", _lib.LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


def test_open_errors_without_cpu_fallback(tmp_path):
    import torch
    import sshash_b200
    with pytest.raises(sshash_b200.SshashGpuError, match="EIO"):
        sshash_b200.Dictionary(str(tmp_path / "missing.sshash"))
    bad = tmp_path / "bad.sshash"
    bad.write_bytes(b"\x04\x01\x01" + b"\0" * 100)
    with pytest.raises(sshash_b200.SshashGpuError, match="MAJOR index version mismatch"):
        sshash_b200.Dictionary(str(bad))
    good = open(golden("se_k47_m8").index, "rb").read()
    bad.write_bytes(good[: len(good) // 3])
    with pytest.raises(sshash_b200.SshashGpuError, match="EFORMAT"):
        sshash_b200.Dictionary(str(bad))
    with pytest.raises(sshash_b200.SshashGpuError, match="EINVAL"):
        sshash_b200.Dictionary(golden("se_k63_m21").index, max_k=31)
    if not torch.cuda.is_available():
        # the product path must refuse to run without a GPU rather than fall back to a CPU path
        with pytest.raises(sshash_b200.SshashGpuError, match="no CUDA device"):
            sshash_b200.Dictionary(golden("se_k31_m13").index)


def test_multi_gpu_handle_errors_without_a_gpu():
    import ctypes as C
    import torch
    import sshash_b200
    from sshash_b200 import _lib
    lib = _lib.lib()
    assert lib.sshash_gpu_multi_num_devices(None) == 0 and lib.sshash_gpu_multi_dict(None, 0) is None
    assert lib.sshash_gpu_multi_lookup_batch(None, None, 1, 1, None) == 1          # EINVAL: null handle
    assert b"null multi-GPU handle" in lib.sshash_gpu_last_error()
    assert lib.sshash_gpu_multi_close(None) == 0
    if not torch.cuda.is_available():
        with pytest.raises(sshash_b200.SshashGpuError, match="no CUDA device"):
            sshash_b200.MultiDictionary(golden("se_k31_m13").index)
        h = C.c_void_p()
        devs = (C.c_int * 2)(0, 0)
        assert lib.sshash_gpu_multi_open(golden("se_k31_m13").index.encode(), devs, 2, 0, C.byref(h)) == 5 and not h


def test_header_is_plain_c_and_examples_compile(tmp_path):
    """include/sshash_gpu.h is a C header (gcc -std=c99 -pedantic); the C and C++ examples build against the
    library and, on a box without a GPU, fail loudly with the library's message instead of computing on the CPU."""
    import subprocess
    import torch
    lib_dir = os.path.join(ROOT, "sshash_b200")
    probe = tmp_path / "probe.c"
    probe.write_text('#include "sshash_gpu.h"\nint main(void) { return sizeof(sshash_lookup_result) == 64 ? 0 : 1; }\n')
    subprocess.check_call(["gcc", "-std=c99", "-pedantic", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"),
                           str(probe), "-o", str(tmp_path / "probe")])
    assert subprocess.run([str(tmp_path / "probe")]).returncode == 0
    builds = [("gcc", ["-std=c11"], "multi_gpu_example.c"), ("g++", ["-std=c++17"], "lookup_example.cpp"),
              ("g++", ["-std=c++17"], "query_example.cpp")]
    for cc, flags, src in builds:
        cc = "/usr/bin/" + cc if os.path.exists("/usr/bin/" + cc) else cc
        exe = str(tmp_path / src.split(".")[0])
        subprocess.check_call([cc, *flags, "-O1", "-Wall", "-Wextra", os.path.join(ROOT, "examples", src), "-o", exe,
                               os.path.join(lib_dir, "libsshash_gpu.so"), "-Wl,-rpath," + lib_dir])
    if not torch.cuda.is_available():
        out = subprocess.run([str(tmp_path / "multi_gpu_example"), golden("se_k31_m13").index, "1000"], capture_output=True, text=True)
        assert out.returncode == 1 and "no CUDA device" in out.stderr
        out = subprocess.run([str(tmp_path / "lookup_example"), golden("se_k31_m13").index, "1000"], capture_output=True, text=True)
        assert out.returncode == 1 and "no CUDA device" in out.stderr


def test_product_never_imports_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "sshash_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in text.lower() or f == "_lib.py" and False, os.path.join(dirpath, f)
