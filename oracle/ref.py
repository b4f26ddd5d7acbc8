"""ctypes wrapper around oracle/_ref/libsshash_ref{31,63}.so -- TEST INFRASTRUCTURE ONLY.

The shared objects are the UNMODIFIED reference (jermp/sshash) compiled by oracle/Makefile from
the sources under /root/reference through oracle/ref_harness.cpp.  Only tests/, the smoke check
and bench.py's cpu_baseline / --impl reference legs may import this module; the product package
(sshash_b200/) never does.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))

RESULT_DTYPE = np.dtype(
    [
        ("kmer_id", "<u8"),
        ("kmer_id_in_string", "<u8"),
        ("kmer_offset", "<u8"),
        ("kmer_orientation", "<i8"),
        ("string_id", "<u8"),
        ("string_begin", "<u8"),
        ("string_end", "<u8"),
        ("minimizer_found", "<u8"),
    ]
)


class Info(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in ("num_kmers", "num_strings", "k", "m", "canonical", "weighted", "max_k")]


class Report(C.Structure):
    _fields_ = [
        (n, C.c_uint64)
        for n in (
            "num_kmers",
            "num_positive_kmers",
            "num_negative_kmers",
            "num_invalid_kmers",
            "num_searches",
            "num_extensions",
        )
    ]

    def as_dict(self):
        return {n: int(getattr(self, n)) for n, _ in self._fields_}


def lib_path(max_k: int) -> str:
    return os.path.join(_HERE, "_ref", "libsshash_ref%d.so" % max_k)


def available(max_k: int = 31) -> bool:
    return os.path.exists(lib_path(max_k))


_libs = {}


def _lib(max_k: int):
    if max_k not in _libs:
        lib = C.CDLL(lib_path(max_k))
        lib.ref_last_error.restype = C.c_char_p
        lib.ref_open.restype = C.c_void_p
        lib.ref_open.argtypes = [C.c_char_p]
        lib.ref_close.argtypes = [C.c_void_p]
        lib.ref_info.argtypes = [C.c_void_p, C.POINTER(Info)]
        lib.ref_build.argtypes = [C.c_char_p, C.c_uint64, C.c_uint64, C.c_int, C.c_uint64, C.c_uint64,
                                  C.c_char_p, C.c_char_p, C.c_int, C.c_int]
        lib.ref_lookup_batch.restype = C.c_double
        lib.ref_lookup_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_int, C.c_void_p, C.c_void_p,
                                         C.c_uint64]
        lib.ref_lookup_batch_ascii.argtypes = [C.c_void_p, C.c_char_p, C.c_uint64, C.c_int, C.c_void_p,
                                               C.c_void_p]
        lib.ref_access_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p]
        lib.ref_weight_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p]
        lib.ref_kmer_neighbours_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_int, C.c_int, C.c_void_p]
        lib.ref_string_neighbours_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_int, C.c_void_p]
        lib.ref_streaming_file.argtypes = [C.c_void_p, C.c_char_p, C.c_int, C.POINTER(Report),
                                           C.POINTER(C.c_double)]
        lib.ref_streaming_reads.restype = C.c_double
        lib.ref_streaming_reads.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p,
                                            C.c_void_p, C.POINTER(Report)]
        _libs[max_k] = lib
    return _libs[max_k]


def build(input_path: str, k: int, m: int, output: str, canonical: bool = False, threads: int = 1,
          seed: int = 0, tmp_dir: str = "", max_k: int | None = None, verbose: bool = False,
          weighted: bool = False) -> None:
    """sshash build -i input -k k -m m [--canonical] [--weighted] -o output (tools/build.cpp)."""
    if max_k is None:
        max_k = 31 if k <= 31 else 63
    lib = _lib(max_k)
    rc = lib.ref_build(input_path.encode(), k, m, int(canonical), threads, seed, tmp_dir.encode(),
                       output.encode(), int(verbose), int(weighted))
    if rc != 0:
        raise RuntimeError("reference build failed: " + lib.ref_last_error().decode())


class RefDictionary:
    """The reference dictionary_type loaded from an index file."""

    def __init__(self, path: str, max_k: int = 31):
        self.max_k = max_k
        self.lib = _lib(max_k)
        self.h = self.lib.ref_open(path.encode())
        if not self.h:
            raise RuntimeError("reference open failed: " + self.lib.ref_last_error().decode())
        info = Info()
        self.lib.ref_info(self.h, C.byref(info))
        for n, _ in Info._fields_:
            setattr(self, n, int(getattr(info, n)))
        self.words = 1 if max_k == 31 else 2

    def close(self):
        if self.h:
            self.lib.ref_close(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _kmers(self, kmers):
        a = np.ascontiguousarray(kmers, dtype=np.uint64)
        n = a.size // self.words
        return a, n

    def lookup(self, kmers, check_rc: bool = True, full: bool = False, threads: int = 1):
        a, n = self._kmers(kmers)
        ids = np.empty(n, dtype=np.uint64)
        res = np.empty(n, dtype=RESULT_DTYPE) if full else None
        self.lib.ref_lookup_batch(self.h, a.ctypes.data, n, int(check_rc), ids.ctypes.data,
                                  res.ctypes.data if full else None, threads)
        return (ids, res) if full else ids

    def time_lookup(self, kmers, check_rc: bool = True, threads: int = 1) -> float:
        a, n = self._kmers(kmers)
        return float(self.lib.ref_lookup_batch(self.h, a.ctypes.data, n, int(check_rc), None, None, threads))

    def lookup_ascii(self, strings: bytes, check_rc: bool = True):
        n = len(strings) // self.k
        ids = np.empty(n, dtype=np.uint64)
        self.lib.ref_lookup_batch_ascii(self.h, strings, n, int(check_rc), ids.ctypes.data, None)
        return ids

    def access(self, ids):
        ids = np.ascontiguousarray(ids, dtype=np.uint64)
        out = np.empty(ids.size * self.words, dtype=np.uint64)
        self.lib.ref_access_batch(self.h, ids.ctypes.data, ids.size, out.ctypes.data)
        return out if self.words == 1 else out.reshape(-1, 2)

    def weight(self, ids):
        """dictionary::weight(kmer_id) (src/dictionary.cpp:96-100)"""
        a = np.ascontiguousarray(ids, dtype=np.uint64)
        out = np.empty(a.size, dtype=np.uint64)
        self.lib.ref_weight_batch(self.h, a.ctypes.data, a.size, out.ctypes.data)
        return out

    def kmer_neighbours(self, kmers, check_rc: bool = True, which: int = 3):
        """(n, 8) lookup_result records: forward[A,C,T,G], backward[A,C,T,G]."""
        a, n = self._kmers(kmers)
        out = np.empty((n, 8), dtype=RESULT_DTYPE)
        self.lib.ref_kmer_neighbours_batch(self.h, a.ctypes.data, n, int(check_rc), which, out.ctypes.data)
        return out

    def string_neighbours(self, string_ids, check_rc: bool = True):
        ids = np.ascontiguousarray(string_ids, dtype=np.uint64)
        out = np.empty((ids.size, 8), dtype=RESULT_DTYPE)
        self.lib.ref_string_neighbours_batch(self.h, ids.ctypes.data, ids.size, int(check_rc), out.ctypes.data)
        return out

    def streaming_file(self, path: str, multiline: bool = False):
        rep = Report()
        secs = C.c_double(0)
        rc = self.lib.ref_streaming_file(self.h, path.encode(), int(multiline), C.byref(rep), C.byref(secs))
        if rc != 0:
            raise RuntimeError(self.lib.ref_last_error().decode())
        return rep.as_dict(), secs.value

    def streaming_reads(self, bases: bytes, offsets, full: bool = False):
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        nreads = offsets.size - 1
        lens = np.diff(offsets.astype(np.int64))
        nwin = int(np.maximum(lens - self.k + 1, 0).sum())
        ids = np.empty(nwin, dtype=np.uint64)
        res = np.empty(nwin, dtype=RESULT_DTYPE) if full else None
        rep = Report()
        buf = np.frombuffer(bases, dtype=np.uint8)
        secs = self.lib.ref_streaming_reads(self.h, buf.ctypes.data, offsets.ctypes.data, nreads, ids.ctypes.data,
                                            res.ctypes.data if full else None, C.byref(rep))
        return ids, res, rep.as_dict(), float(secs)
