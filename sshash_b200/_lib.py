"""ctypes binding of libsshash_gpu.so (the C ABI of include/sshash_gpu.h).

The shared library is built in-tree by sshash_b200/csrc/Makefile (nvcc, sm_100a only).  There is
no Python or CPU implementation of the lookup path in this package: if the library is missing
the import of `sshash_b200.dictionary` fails loudly.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
# SSHASH_GPU_LIB: A/B builds of the same library during kernel work (tools/, not the product default)
LIB_PATH = os.environ.get("SSHASH_GPU_LIB") or os.path.join(_HERE, "libsshash_gpu.so")

STATUS = {0: "OK", 1: "EINVAL", 2: "EIO", 3: "EFORMAT", 4: "EVERSION", 5: "ECUDA", 6: "ENOMEM"}


class SshashGpuError(RuntimeError):
    def __init__(self, status: int, message: str):
        super().__init__("sshash_gpu %s: %s" % (STATUS.get(status, status), message))
        self.status = status


class LookupResult(C.Structure):
    """lookup_result, reference include/util.hpp:38-62"""
    _fields_ = [("kmer_id", C.c_uint64), ("kmer_id_in_string", C.c_uint64), ("kmer_offset", C.c_uint64),
                ("kmer_orientation", C.c_int64), ("string_id", C.c_uint64), ("string_begin", C.c_uint64),
                ("string_end", C.c_uint64), ("minimizer_found", C.c_uint64)]


class StreamingReport(C.Structure):
    """streaming_query_report, reference include/util.hpp:21-36"""
    _fields_ = [(n, C.c_uint64) for n in ("num_kmers", "num_positive_kmers", "num_negative_kmers",
                                          "num_invalid_kmers", "num_searches", "num_extensions")]

    def as_dict(self):
        return {n: int(getattr(self, n)) for n, _ in self._fields_}


class Info(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in ("num_kmers", "num_strings", "k", "m", "canonical", "weighted", "max_k",
                                          "version", "num_minimizers", "mphf_partitions", "skew_partitions",
                                          "index_file_bytes", "device_bytes")] + [("device", C.c_int64)]


# every symbol include/sshash_gpu.h declares: (restype, argtypes)
SYMBOLS = {
    "sshash_gpu_last_error": (C.c_char_p, []),
    "sshash_gpu_build_info": (C.c_char_p, []),
    "sshash_gpu_launch_count": (C.c_uint64, []),
    "sshash_gpu_open": (C.c_int, [C.c_char_p, C.c_int, C.c_int, C.POINTER(C.c_void_p)]),
    "sshash_gpu_close": (C.c_int, [C.c_void_p]),
    "sshash_gpu_info": (C.c_int, [C.c_void_p, C.POINTER(Info)]),
    "sshash_gpu_set_peer_inplace": (C.c_int, [C.c_void_p, C.c_int]),
    "sshash_gpu_lookup_batch": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint64, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    "sshash_gpu_lookup_batch_u32": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint64, C.c_int, C.c_void_p, C.c_void_p]),
    "sshash_gpu_lookup_batch_ascii": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint64, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    "sshash_gpu_is_member_batch": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint64, C.c_int, C.c_void_p, C.c_void_p]),
    "sshash_gpu_minimizer_partition_batch": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p]),
    "sshash_gpu_access_batch": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p]),
    "sshash_gpu_weight_batch": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p]),
    "sshash_gpu_kmer_neighbours_batch": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint64, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                                   C.c_void_p]),
    "sshash_gpu_string_neighbours_batch": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint64, C.c_int, C.c_void_p, C.c_void_p,
                                                     C.c_void_p]),
    "sshash_gpu_check_input_contract": (C.c_int, [C.c_void_p, C.POINTER(C.c_int)]),
    "sshash_gpu_streaming_batch": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p,
                                             C.POINTER(StreamingReport), C.c_void_p]),
    "sshash_gpu_streaming_query_from_file": (C.c_int, [C.c_void_p, C.c_char_p, C.c_int, C.POINTER(StreamingReport)]),
    "sshash_gpu_multi_open": (C.c_int, [C.c_char_p, C.POINTER(C.c_int), C.c_int, C.c_int, C.POINTER(C.c_void_p)]),
    "sshash_gpu_multi_close": (C.c_int, [C.c_void_p]),
    "sshash_gpu_multi_num_devices": (C.c_int, [C.c_void_p]),
    "sshash_gpu_multi_dict": (C.c_void_p, [C.c_void_p, C.c_int]),
    "sshash_gpu_multi_lookup_batch": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint64, C.c_int, C.c_void_p]),
    "sshash_gpu_multi_lookup_batch_u32": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint64, C.c_int, C.c_void_p]),
    "sshash_gpu_multi_is_member_batch": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint64, C.c_int, C.c_void_p]),
    "sshash_gpu_multi_streaming_batch": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p,
                                                   C.POINTER(StreamingReport)]),
}

_lib = None


def build(force: bool = False) -> str:
    """Compile the CUDA library in-tree (nvcc cross-compiles sm_100a without a GPU)."""
    src_dir = os.path.join(_HERE, "csrc")
    if force and os.path.exists(LIB_PATH):
        os.remove(LIB_PATH)
    subprocess.check_call(["make", "-j4", "-C", src_dir], stdout=subprocess.DEVNULL)
    return LIB_PATH


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                "sshash_b200: %s is missing -- build it with `make -C sshash_b200/csrc` or "
                "`python -c 'import __graft_entry__ as g; g.build()'`. There is no CPU fallback." % LIB_PATH)
        l = C.CDLL(LIB_PATH)
        for name, (restype, argtypes) in SYMBOLS.items():
            fn = getattr(l, name)  # AttributeError if the library does not export it
            fn.restype = restype
            fn.argtypes = argtypes
        _lib = l
    return _lib


def check(status: int) -> None:
    if status != 0:
        raise SshashGpuError(status, lib().sshash_gpu_last_error().decode())
