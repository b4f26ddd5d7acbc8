"""CPU test of the index loader (the host side of sshash_gpu_open, include/sshash_gpu.h): damaged
files must come back as a status code -- EFORMAT / EVERSION / EINVAL, or ECUDA when the damage is
invisible to the structural parse and only the missing GPU stops the open -- never as a crash, a
hang or an out-of-memory.  The reference throws on a short read (essentials.hpp:413-417) and
otherwise trusts the file; so does the device path once a file has passed these checks."""
import ast
import os
import subprocess
import sys

import pytest

from conftest import ROOT, golden


def run(index, seed, n, mode, span=None):
    env = dict(os.environ)
    if span:
        env["FUZZ_SPAN"] = str(span)
    p = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "_fuzz_open.py"), index, str(seed), str(n), mode],
                       capture_output=True, text=True, timeout=600, env=env)
    assert p.returncode == 0, "loader crashed: rc=%d\n%s" % (p.returncode, p.stderr[-2000:])
    assert "slow" not in p.stdout, p.stdout
    return ast.literal_eval(p.stdout.strip().splitlines()[-1])


@pytest.mark.parametrize("name", ["se_k31_m13", "se_k63_m8_canon", "ecoli_k31_m11_canon_weighted"])
def test_truncated_index_files_are_rejected(name):
    stats = run(golden(name).index, 11, 60, "truncate")
    assert set(stats) <= {"EFORMAT", "EVERSION", "EIO"}, stats


@pytest.mark.parametrize("name", ["se_k31_m13", "sal100_k31_m7_reg"])
def test_corrupted_header_fields_never_crash_the_loader(name):
    import torch
    if torch.cuda.is_available():
        pytest.skip("with a GPU a structurally valid but corrupted file would be uploaded and queried")
    stats = run(golden(name).index, 12, 150, "flip", span=48)
    assert "opened" not in stats and set(stats) <= {"EFORMAT", "EVERSION", "EINVAL", "ECUDA"}, stats
    assert stats.get("EFORMAT", 0) > 0


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["sal100_k31_m7_reg", "se_k31_m13"])
def test_corrupted_files_on_a_gpu_open_cleanly_or_are_rejected(name):
    """With a GPU the open goes all the way: parse, upload, the open-time validation kernel (every control
    codeword and bucket offset is range-checked), the fingerprint build.  Bit flips anywhere in the file must
    give `opened` or a status code -- never a crash or a sticky CUDA error -- and a good file must still open
    and answer correctly in the same process afterwards (the helper does that last step itself)."""
    import numpy as np
    import sshash_b200
    g = golden(name)
    stats = run(g.index, 13, 80, "flip", span=1 << 30)
    assert set(stats) <= {"opened", "EFORMAT", "EVERSION", "EINVAL"}, stats
    d = sshash_b200.Dictionary(g.index, max_k=g.max_k)
    assert (d.lookup_batch(g.z["queries"]) == g.z["ids"]).all()
    d.close()
