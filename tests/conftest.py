import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _cuda_device_count() -> int:
    """Number of CUDA devices, asked of the driver directly (no torch import, no library call)."""
    import ctypes
    for name in ("libcuda.so.1", "libcuda.so"):
        try:
            cu = ctypes.CDLL(name)
        except OSError:
            continue
        n = ctypes.c_int(0)
        if cu.cuInit(0) == 0 and cu.cuDeviceGetCount(ctypes.byref(n)) == 0:
            return n.value
        return 0
    return 0


def pytest_collection_modifyitems(config, items):
    """`gpu` tests are skipped (not failed) on a box without a CUDA device or without the built library,
    so that a plain `pytest tests` on a CPU-only machine stays green."""
    lib = os.path.join(ROOT, "sshash_b200", "libsshash_gpu.so")
    reason = None
    if not os.path.exists(lib):
        reason = "sshash_b200/libsshash_gpu.so is not built"
    elif _cuda_device_count() == 0:
        reason = "no CUDA device"
    if reason is None:
        return
    skip = pytest.mark.skip(reason=reason)
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def load_manifest():
    with open(os.path.join(GOLDEN, "manifest.json")) as f:
        return json.load(f)


MANIFEST = load_manifest()
FIXTURES = sorted(MANIFEST)
WEIGHTED = [n for n in FIXTURES if MANIFEST[n].get("weighted")]
# indexes built from inputs with duplicated k-mers / reverse-complement twins: lookup(access(id))
# need not be id there, everything else must still equal the reference
NON_DISTINCT = [n for n in FIXTURES if MANIFEST[n].get("distinct_kmers") is False]


@pytest.fixture(scope="session")
def manifest():
    return MANIFEST


class Golden:
    def __init__(self, name):
        self.name = name
        self.meta = MANIFEST[name]
        self.index = os.path.join(GOLDEN, name + ".sshash")
        z = np.load(os.path.join(GOLDEN, name + ".npz"))
        self.z = {k: z[k] for k in z.files}
        self.max_k = self.meta["max_k"]
        self.words = 1 if self.max_k == 31 else 2


_cache = {}


def golden(name):
    if name not in _cache:
        _cache[name] = Golden(name)
    return _cache[name]


REPORT_KEYS = ("num_kmers", "num_positive_kmers", "num_negative_kmers", "num_invalid_kmers", "num_searches",
               "num_extensions")
