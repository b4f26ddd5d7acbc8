"""ctypes wrapper around liboracle.so (oracle/sshash_oracle.c) -- TEST INFRASTRUCTURE ONLY."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from .ref import RESULT_DTYPE, Report

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "liboracle.so")
_lib = None


def build(force: bool = False) -> str:
    """Compile oracle/sshash_oracle.c with gcc (a few seconds)."""
    src = os.path.join(_HERE, "sshash_oracle.c")
    if force or not os.path.exists(_LIB) or os.path.getmtime(_LIB) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "liboracle.so"], stdout=subprocess.DEVNULL)
    return _LIB


def lib():
    global _lib
    if _lib is None:
        build()
        l = C.CDLL(_LIB)
        l.oracle_open.restype = C.c_void_p
        l.oracle_open.argtypes = [C.c_char_p, C.c_int, C.c_char_p, C.c_uint64]
        l.oracle_close.argtypes = [C.c_void_p]
        l.oracle_info.argtypes = [C.c_void_p, C.c_void_p]
        l.oracle_lookup_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_int, C.c_void_p, C.c_void_p]
        l.oracle_lookup_batch_ascii.argtypes = [C.c_void_p, C.c_char_p, C.c_uint64, C.c_int, C.c_void_p, C.c_void_p]
        l.oracle_access_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p]
        l.oracle_weight_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p]
        l.oracle_weighted.argtypes = [C.c_void_p]
        l.oracle_kmer_neighbours_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_int, C.c_int, C.c_void_p]
        l.oracle_string_neighbours_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_int, C.c_void_p]
        l.oracle_streaming_reads.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p,
                                             C.POINTER(Report)]
        _lib = l
    return _lib


_INFO = ("num_kmers", "num_strings", "k", "m", "canonical", "magic", "mphf_seed", "mphf_partitions",
         "num_minimizers", "skew_partitions", "heavy_size", "mid_load_size", "codeword_width",
         "strings_bits", "weights_bytes", "kmer_bits")


class OracleDictionary:
    def __init__(self, path: str, max_k: int = 0):
        err = C.create_string_buffer(256)
        self.h = lib().oracle_open(path.encode(), max_k, err, 256)
        if not self.h:
            raise RuntimeError(err.value.decode())
        info = np.zeros(16, dtype=np.uint64)
        lib().oracle_info(self.h, info.ctypes.data)
        for n, v in zip(_INFO, info):
            setattr(self, n, int(v))
        self.words = self.kmer_bits // 64

    def close(self):
        if self.h:
            lib().oracle_close(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def lookup(self, kmers, check_rc: bool = True, full: bool = False):
        a = np.ascontiguousarray(kmers, dtype=np.uint64)
        n = a.size // self.words
        ids = np.empty(n, dtype=np.uint64)
        res = np.empty(n, dtype=RESULT_DTYPE) if full else None
        lib().oracle_lookup_batch(self.h, a.ctypes.data, n, int(check_rc), ids.ctypes.data,
                                  res.ctypes.data if full else None)
        return (ids, res) if full else ids

    def lookup_ascii(self, strings: bytes, check_rc: bool = True):
        n = len(strings) // self.k
        ids = np.empty(n, dtype=np.uint64)
        lib().oracle_lookup_batch_ascii(self.h, strings, n, int(check_rc), ids.ctypes.data, None)
        return ids

    def access(self, ids):
        ids = np.ascontiguousarray(ids, dtype=np.uint64)
        out = np.empty(ids.size * self.words, dtype=np.uint64)
        lib().oracle_access_batch(self.h, ids.ctypes.data, ids.size, out.ctypes.data)
        return out if self.words == 1 else out.reshape(-1, 2)

    def weighted(self) -> bool:
        return bool(lib().oracle_weighted(self.h))

    def weight(self, ids):
        ids = np.ascontiguousarray(ids, dtype=np.uint64)
        out = np.empty(ids.size, dtype=np.uint64)
        lib().oracle_weight_batch(self.h, ids.ctypes.data, ids.size, out.ctypes.data)
        return out

    def kmer_neighbours(self, kmers, check_rc: bool = True, which: int = 3):
        a = np.ascontiguousarray(kmers, dtype=np.uint64)
        n = a.size // self.words
        out = np.empty((n, 8), dtype=RESULT_DTYPE)
        lib().oracle_kmer_neighbours_batch(self.h, a.ctypes.data, n, int(check_rc), which, out.ctypes.data)
        return out

    def string_neighbours(self, string_ids, check_rc: bool = True):
        ids = np.ascontiguousarray(string_ids, dtype=np.uint64)
        out = np.empty((ids.size, 8), dtype=RESULT_DTYPE)
        lib().oracle_string_neighbours_batch(self.h, ids.ctypes.data, ids.size, int(check_rc), out.ctypes.data)
        return out

    def streaming_reads(self, bases: bytes, offsets, full: bool = False):
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        nreads = offsets.size - 1
        lens = np.diff(offsets.astype(np.int64))
        nwin = int(np.maximum(lens - self.k + 1, 0).sum())
        ids = np.empty(nwin, dtype=np.uint64)
        res = np.empty(nwin, dtype=RESULT_DTYPE) if full else None
        rep = Report()
        buf = np.frombuffer(bases, dtype=np.uint8)
        lib().oracle_streaming_reads(self.h, buf.ctypes.data, offsets.ctypes.data, nreads, ids.ctypes.data,
                                     res.ctypes.data if full else None, C.byref(rep))
        return ids, res, rep.as_dict()
