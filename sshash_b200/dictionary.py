"""Python mirror of the reference's `dictionary<Kmer, Offsets>` for the lookup path.

Same method names and argument meaning as the reference class (include/dictionary.hpp:10-181):
`k() m() canonical() num_kmers() num_strings() weighted()`, `lookup`, `is_member`, `access`,
`streaming_query_from_file`, plus the batched forms every caller of the reference loops to get
(tools/perf.hpp:55-60, test/check.hpp:29-31).  Everything executes in the CUDA library through the
C ABI (include/sshash_gpu.h); numpy arrays are host buffers, torch CUDA tensors are used in place.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import numpy as np

from . import _lib
from ._lib import Info, LookupResult, StreamingReport, SshashGpuError, check  # noqa: F401

INVALID = np.uint64(2**64 - 1)  # constants::invalid_uint64, include/constants.hpp:5

RESULT_DTYPE = np.dtype([("kmer_id", "<u8"), ("kmer_id_in_string", "<u8"), ("kmer_offset", "<u8"),
                         ("kmer_orientation", "<i8"), ("string_id", "<u8"), ("string_begin", "<u8"),
                         ("string_end", "<u8"), ("minimizer_found", "<u8")])

_NUC = {"A": 0, "C": 1, "T": 2, "G": 3}


def _is_torch(x) -> bool:
    return type(x).__module__.startswith("torch")


def _stream_for(x, stream) -> int:
    """cudaStream_t for a call on `x`: an explicit `stream` wins; torch CUDA tensors default to torch's
    CURRENT stream on their device (side streams are non-blocking, so the legacy default stream would
    race the caller's producer/consumer kernels); host buffers ignore the argument (0)."""
    if stream is not None:
        return int(stream)
    if _is_torch(x) and x.is_cuda:
        import torch
        return int(torch.cuda.current_stream(x.device).cuda_stream)
    return 0


def _ptr(x) -> int:
    if x is None:
        return 0
    if _is_torch(x):
        return x.data_ptr()
    return x.ctypes.data


def string_to_uint_kmer(s: str) -> int:
    """util::string_to_uint_kmer (include/util.hpp:207-213): (c >> 1) & 3, base 0 in the low bits."""
    x = 0
    for i, c in enumerate(s.encode()):
        x |= ((c >> 1) & 3) << (2 * i)
    return x


def uint_kmer_to_string(x: int, k: int) -> str:
    return "".join("ACTG"[(x >> (2 * i)) & 3] for i in range(k))


class Dictionary:
    """An SSHash index resident in the HBM of one GPU."""

    def __init__(self, index_filename: str, device: int = 0, max_k: int = 0):
        self._lib = _lib.lib()
        self._h = C.c_void_p()
        check(self._lib.sshash_gpu_open(index_filename.encode(), device, max_k, C.byref(self._h)))
        info = Info()
        check(self._lib.sshash_gpu_info(self._h, C.byref(info)))
        self.info = {n: int(getattr(info, n)) for n, _ in Info._fields_}
        self.words = 1 if self.info["max_k"] == 31 else 2
        self.device = device

    # ---- lifetime ---------------------------------------------------------------------------
    def close(self) -> None:
        if getattr(self, "_h", None) is not None and self._h:
            self._lib.sshash_gpu_close(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def set_peer_inplace(self, inplace: bool) -> None:
        """Device pointers of OTHER GPUs: staged with copy engines (False, default) or dereferenced by the
        kernels over NVLink (True; the caller guarantees peer access / a peer-mapped allocation)."""
        check(self._lib.sshash_gpu_set_peer_inplace(self._h, int(inplace)))

    # ---- accessors, include/dictionary.hpp:31-38 ----------------------------------------------
    def k(self) -> int: return self.info["k"]
    def m(self) -> int: return self.info["m"]
    def canonical(self) -> bool: return bool(self.info["canonical"])
    def weighted(self) -> bool: return bool(self.info["weighted"])
    def num_kmers(self) -> int: return self.info["num_kmers"]
    def num_strings(self) -> int: return self.info["num_strings"]

    # ---- helpers ------------------------------------------------------------------------------
    def _count(self, kmers) -> int:
        n = kmers.numel() if _is_torch(kmers) else kmers.size
        if n % self.words:
            raise ValueError("packed k-mer buffer length must be a multiple of %d words" % self.words)
        return n // self.words

    def _prep_in(self, a, dtype=np.uint64):
        if _is_torch(a):
            if not a.is_contiguous():
                a = a.contiguous()
            return a
        return np.ascontiguousarray(a, dtype=dtype)

    def _alloc_like(self, ref, n, dtype, shape=None):
        shape = (n,) if shape is None else shape
        if _is_torch(ref):
            import torch
            tdt = {np.uint64: torch.int64, np.uint8: torch.uint8}[dtype]
            return torch.empty(shape, dtype=tdt, device=ref.device)
        return np.empty(shape, dtype=dtype)

    # ---- lookup, include/dictionary.hpp:41-42 -------------------------------------------------
    def lookup_batch(self, kmers, check_reverse_complement: bool = True, out=None, full: bool = False,
                     stream: Optional[int] = None):
        """Batched dictionary::lookup(Kmer, bool).  kmers: uint64 array (numpy = host, torch CUDA
        tensor = device), `words` words per k-mer.  Returns kmer ids (uint64; torch: int64 bit
        patterns) or, with full=True, a structured array of complete lookup_result records."""
        kmers = self._prep_in(kmers)
        n = self._count(kmers)
        if full:
            if _is_torch(kmers):
                import torch
                res = torch.empty((n, 8), dtype=torch.int64, device=kmers.device)
            else:
                res = np.empty(n, dtype=RESULT_DTYPE)
            check(self._lib.sshash_gpu_lookup_batch(self._h, _ptr(kmers), n, int(check_reverse_complement), None,
                                                    _ptr(res), _stream_for(kmers, stream)))
            return res
        ids = out if out is not None else self._alloc_like(kmers, n, np.uint64)
        check(self._lib.sshash_gpu_lookup_batch(self._h, _ptr(kmers), n, int(check_reverse_complement), _ptr(ids),
                                                None, _stream_for(kmers, stream)))
        return ids

    def lookup_batch_u32(self, kmers, check_reverse_complement: bool = True, out=None, stream: Optional[int] = None):
        """lookup_batch with 32-bit ids (dictionaries with < 2^32 - 1 k-mers): "not found" is
        UINT32_MAX (numpy uint32; torch: int32 bit patterns, i.e. -1)."""
        kmers = self._prep_in(kmers)
        n = self._count(kmers)
        if out is None:
            if _is_torch(kmers):
                import torch
                out = torch.empty(n, dtype=torch.int32, device=kmers.device)
            else:
                out = np.empty(n, dtype=np.uint32)
        check(self._lib.sshash_gpu_lookup_batch_u32(self._h, _ptr(kmers), n, int(check_reverse_complement), _ptr(out),
                                                    _stream_for(kmers, stream)))
        return out

    def lookup_batch_ascii(self, strings: bytes, check_reverse_complement: bool = True, full: bool = False):
        """Batched dictionary::lookup(char const*, bool): n*k characters, no validation."""
        buf = np.frombuffer(strings, dtype=np.uint8)
        if buf.size % self.k():
            raise ValueError("ASCII buffer length must be a multiple of k")
        n = buf.size // self.k()
        if full:
            res = np.empty(n, dtype=RESULT_DTYPE)
            check(self._lib.sshash_gpu_lookup_batch_ascii(self._h, _ptr(buf), n, int(check_reverse_complement), None,
                                                          _ptr(res), None))
            return res
        ids = np.empty(n, dtype=np.uint64)
        check(self._lib.sshash_gpu_lookup_batch_ascii(self._h, _ptr(buf), n, int(check_reverse_complement), _ptr(ids),
                                                      None, None))
        return ids

    def lookup(self, kmer, check_reverse_complement: bool = True) -> dict:
        """Scalar dictionary::lookup (string or packed int) -> lookup_result as a dict."""
        if isinstance(kmer, str):
            if len(kmer) != self.k():
                raise ValueError("k-mer string must have length k")
            kmer = string_to_uint_kmer(kmer)
        q = np.array([(kmer >> (64 * i)) & (2**64 - 1) for i in range(self.words)], dtype=np.uint64)
        r = self.lookup_batch(q, check_reverse_complement, full=True)[0]
        return {n: int(r[n]) for n in RESULT_DTYPE.names}

    # ---- membership, include/dictionary.hpp:75-76 ---------------------------------------------
    def is_member_batch(self, kmers, check_reverse_complement: bool = True, stream: Optional[int] = None, out=None):
        """member[i] = 1 iff found.  `out`: optional uint8 buffer (e.g. pinned) to receive the bytes as they are."""
        kmers = self._prep_in(kmers)
        n = self._count(kmers)
        if out is not None:
            check(self._lib.sshash_gpu_is_member_batch(self._h, _ptr(kmers), n, int(check_reverse_complement), _ptr(out),
                                                       _stream_for(kmers, stream)))
            return out
        out = self._alloc_like(kmers, n, np.uint8)
        check(self._lib.sshash_gpu_is_member_batch(self._h, _ptr(kmers), n, int(check_reverse_complement), _ptr(out),
                                                   _stream_for(kmers, stream)))
        return out if _is_torch(out) else out.astype(bool)

    def is_member(self, kmer, check_reverse_complement: bool = True) -> bool:
        return self.lookup(kmer, check_reverse_complement)["kmer_id"] != int(INVALID)

    # ---- access, include/dictionary.hpp:71 ----------------------------------------------------
    def access_batch(self, kmer_ids, stream: Optional[int] = None):
        kmer_ids = self._prep_in(kmer_ids)
        n = kmer_ids.numel() if _is_torch(kmer_ids) else kmer_ids.size
        shape = (n,) if self.words == 1 else (n, 2)
        out = self._alloc_like(kmer_ids, n, np.uint64, shape)
        check(self._lib.sshash_gpu_access_batch(self._h, _ptr(kmer_ids), n, _ptr(out), _stream_for(kmer_ids, stream)))
        return out

    def access(self, kmer_id: int) -> str:
        w = self.access_batch(np.array([kmer_id], dtype=np.uint64)).reshape(-1)
        x = int(w[0]) | (int(w[1]) << 64 if self.words == 2 else 0)
        return uint_kmer_to_string(x, self.k())

    # ---- weight, include/dictionary.hpp:65-66 ---------------------------------------------------
    def weight_batch(self, kmer_ids, stream: Optional[int] = None):
        """dictionary::weight for a batch of k-mer ids (weighted dictionaries only)."""
        kmer_ids = self._prep_in(kmer_ids)
        n = kmer_ids.numel() if _is_torch(kmer_ids) else kmer_ids.size
        out = self._alloc_like(kmer_ids, n, np.uint64, (n,))
        check(self._lib.sshash_gpu_weight_batch(self._h, _ptr(kmer_ids), n, _ptr(out), _stream_for(kmer_ids, stream)))
        return out

    def weight(self, kmer_id: int) -> int:
        return int(self.weight_batch(np.array([kmer_id], dtype=np.uint64))[0])

    # ---- navigational queries, include/dictionary.hpp:50-66 ------------------------------------
    def kmer_neighbours_batch(self, kmers, check_reverse_complement: bool = True, which: int = 3, full: bool = True):
        """Batched kmer_neighbours (which=3), kmer_forward_neighbours (1), kmer_backward_neighbours (2):
        (n, 8) results per k-mer = forward[A,C,T,G] + backward[A,C,T,G]; host (numpy) buffers."""
        kmers = np.ascontiguousarray(kmers, dtype=np.uint64)
        n = self._count(kmers)
        out = np.empty((n, 8), dtype=RESULT_DTYPE if full else np.uint64)
        check(self._lib.sshash_gpu_kmer_neighbours_batch(self._h, _ptr(kmers), n, int(check_reverse_complement), which,
                                                         None if full else _ptr(out), _ptr(out) if full else None, None))
        return out

    def string_neighbours_batch(self, string_ids, check_reverse_complement: bool = True, full: bool = True):
        string_ids = np.ascontiguousarray(string_ids, dtype=np.uint64)
        out = np.empty((string_ids.size, 8), dtype=RESULT_DTYPE if full else np.uint64)
        check(self._lib.sshash_gpu_string_neighbours_batch(self._h, _ptr(string_ids), string_ids.size,
                                                           int(check_reverse_complement), None if full else _ptr(out),
                                                           _ptr(out) if full else None, None))
        return out

    def breaks_input_contract(self) -> bool:
        """True when the index holds duplicated k-mers or (regular index) reverse-complement twins: streaming
        then replays the reference's state machine instead of using the shortcuts (checked on the device once)."""
        flag = C.c_int(0)
        check(self._lib.sshash_gpu_check_input_contract(self._h, C.byref(flag)))
        return bool(flag.value)

    # ---- streaming, include/streaming_query.hpp + src/query.cpp -------------------------------
    def streaming_batch(self, bases, read_offsets, want_ids: bool = True, stream: Optional[int] = None):
        """Streaming membership over a batch of reads (concatenated characters + offsets).
        Returns (kmer_ids or None, report dict)."""
        if isinstance(bases, (bytes, bytearray)):
            bases = np.frombuffer(bases, dtype=np.uint8)
        bases = self._prep_in(bases, np.uint8)
        read_offsets = self._prep_in(read_offsets)
        nreads = (read_offsets.numel() if _is_torch(read_offsets) else read_offsets.size) - 1
        ids = None
        if want_ids:
            if _is_torch(read_offsets):
                lens = read_offsets[1:] - read_offsets[:-1]
                nwin = int((lens - self.k() + 1).clamp(min=0).sum().item())
            else:
                lens = np.diff(read_offsets.astype(np.int64))
                nwin = int(np.maximum(lens - self.k() + 1, 0).sum())
            ids = self._alloc_like(read_offsets, max(nwin, 1), np.uint64)[:nwin]
        rep = StreamingReport()
        check(self._lib.sshash_gpu_streaming_batch(self._h, _ptr(bases), _ptr(read_offsets), max(nreads, 0),
                                                   _ptr(ids) if want_ids else None, C.byref(rep), _stream_for(bases, stream)))
        return ids, rep.as_dict()

    def streaming_query_from_file(self, filename: str, multiline: bool = False) -> dict:
        rep = StreamingReport()
        check(self._lib.sshash_gpu_streaming_query_from_file(self._h, filename.encode(), int(multiline), C.byref(rep)))
        return rep.as_dict()


class MultiDictionary:
    """The same index replicated on several GPUs of one box behind ONE handle of the C ABI
    (sshash_gpu_multi_*): batches are sharded by query across the GPUs inside the library, results
    come back in query order.  numpy arrays = host buffers; a torch CUDA tensor on any GPU of the
    box is used in place by its GPU and reached peer-to-peer by the others."""

    def __init__(self, index_filename: str, devices=None, max_k: int = 0):
        self._lib = _lib.lib()
        self._h = C.c_void_p()
        if devices is None:
            arr, n = None, 0
        else:
            devices = list(devices)
            arr, n = (C.c_int * len(devices))(*devices), len(devices)
        check(self._lib.sshash_gpu_multi_open(index_filename.encode(), arr, n, max_k, C.byref(self._h)))
        self.num_devices = int(self._lib.sshash_gpu_multi_num_devices(self._h))
        info = Info()
        check(self._lib.sshash_gpu_info(self._lib.sshash_gpu_multi_dict(self._h, 0), C.byref(info)))
        self.info = {n_: int(getattr(info, n_)) for n_, _ in Info._fields_}
        self.words = 1 if self.info["max_k"] == 31 else 2

    def close(self) -> None:
        if getattr(self, "_h", None) is not None and self._h:
            self._lib.sshash_gpu_multi_close(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def k(self) -> int: return self.info["k"]
    def num_kmers(self) -> int: return self.info["num_kmers"]

    def _n(self, kmers) -> int:
        n = kmers.numel() if _is_torch(kmers) else kmers.size
        if n % self.words:
            raise ValueError("packed k-mer buffer length must be a multiple of %d words" % self.words)
        return n // self.words

    def _in(self, kmers):
        if _is_torch(kmers):
            return kmers if kmers.is_contiguous() else kmers.contiguous()
        return np.ascontiguousarray(kmers, dtype=np.uint64)

    def _out(self, ref, n, np_dtype):
        if _is_torch(ref):
            import torch
            return torch.empty(n, dtype={np.uint64: torch.int64, np.uint32: torch.int32, np.uint8: torch.uint8}[np_dtype],
                               device=ref.device)
        return np.empty(n, dtype=np_dtype)

    def _sync(self, kmers):
        if _is_torch(kmers) and kmers.is_cuda:      # the library works on its own streams
            import torch
            torch.cuda.current_stream(kmers.device).synchronize()

    def lookup_batch(self, kmers, check_reverse_complement: bool = True, out=None):
        kmers = self._in(kmers)
        n = self._n(kmers)
        ids = out if out is not None else self._out(kmers, n, np.uint64)
        self._sync(kmers)
        check(self._lib.sshash_gpu_multi_lookup_batch(self._h, _ptr(kmers), n, int(check_reverse_complement), _ptr(ids)))
        return ids

    def lookup_batch_u32(self, kmers, check_reverse_complement: bool = True, out=None):
        kmers = self._in(kmers)
        n = self._n(kmers)
        ids = out if out is not None else self._out(kmers, n, np.uint32)
        self._sync(kmers)
        check(self._lib.sshash_gpu_multi_lookup_batch_u32(self._h, _ptr(kmers), n, int(check_reverse_complement), _ptr(ids)))
        return ids

    def is_member_batch(self, kmers, check_reverse_complement: bool = True):
        kmers = self._in(kmers)
        n = self._n(kmers)
        out = self._out(kmers, n, np.uint8)
        self._sync(kmers)
        check(self._lib.sshash_gpu_multi_is_member_batch(self._h, _ptr(kmers), n, int(check_reverse_complement), _ptr(out)))
        return out if _is_torch(out) else out.astype(bool)

    def streaming_batch(self, bases, read_offsets, want_ids: bool = True):
        if isinstance(bases, (bytes, bytearray)):
            bases = np.frombuffer(bases, dtype=np.uint8)
        bases = np.ascontiguousarray(bases, dtype=np.uint8)
        read_offsets = np.ascontiguousarray(read_offsets, dtype=np.uint64)
        nreads = read_offsets.size - 1
        ids = None
        if want_ids:
            lens = np.diff(read_offsets.astype(np.int64))
            ids = np.empty(max(int(np.maximum(lens - self.k() + 1, 0).sum()), 1), dtype=np.uint64)
        rep = StreamingReport()
        check(self._lib.sshash_gpu_multi_streaming_batch(self._h, _ptr(bases), _ptr(read_offsets), max(nreads, 0),
                                                         _ptr(ids) if want_ids else None, C.byref(rep)))
        if want_ids:
            lens = np.diff(read_offsets.astype(np.int64))
            ids = ids[: int(np.maximum(lens - self.k() + 1, 0).sum())]
        return ids, rep.as_dict()


def launch_count() -> int:
    return int(_lib.lib().sshash_gpu_launch_count())
