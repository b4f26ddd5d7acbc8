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
