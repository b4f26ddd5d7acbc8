"""sshash_b200 -- B200-native batched k-mer Lookup / streaming membership over SSHash indexes.

The product is the CUDA library `libsshash_gpu.so` (sshash_b200/csrc, C ABI in
include/sshash_gpu.h).  This package is the thin host-side mirror of the reference's
`dictionary<>` interface on top of that ABI, plus the query-sharded multi-GPU driver.
"""
from .dictionary import Dictionary, MultiDictionary, INVALID, RESULT_DTYPE, SshashGpuError, launch_count  # noqa: F401
from .dictionary import string_to_uint_kmer, uint_kmer_to_string  # noqa: F401
