/*
 * sshash_gpu.h -- C ABI of the B200-native SSHash lookup path.
 *
 * The reference (jermp/sshash @ afff26dc) has no FFI/plugin layer: its boundary for this path is
 * the C++ class template `dictionary<Kmer, Offsets>` (include/dictionary.hpp:10-181).  Each entry
 * point below names the reference interface it replaces.  The C++ drop-in wrapper that keeps the
 * reference's method names on top of this ABI is sshash_b200/csrc/dictionary.hpp; the Python
 * mirror is sshash_b200/dictionary.py.  INTEGRATION.md shows the reference-side binding.
 *
 * Conventions
 *   - plain pointers and sizes only; no exceptions cross the boundary: every call returns an
 *     sshash_gpu_status and sshash_gpu_last_error() (thread-local) describes the last failure;
 *   - every data pointer may be a HOST pointer (pageable or pinned) or a DEVICE pointer on the
 *     dictionary's GPU; the library detects which (cudaPointerGetAttributes).  Host buffers are
 *     streamed through the GPU in chunks with copies overlapped with the kernels; device buffers
 *     are used in place (for lookup / membership / access / weight a pointer into ANOTHER GPU's
 *     memory is accepted too and staged like a host buffer, peer-to-peer when peer access is on).
 *     Device buffers of packed k-mers must be 8-byte aligned, 16-byte aligned when max_k = 63 or when
 *     full lookup_result records are requested (128-bit loads / stores);
 *   - `stream` is a cudaStream_t passed as void* (NULL = the legacy default stream).  With device
 *     buffers the call is asynchronous on that stream; with host buffers `stream` is ignored, the
 *     library pipelines the batch on its own streams and returns when the outputs are complete;
 *   - packed k-mers: 2 bits per base, base i at bits [2i, 2i+1], A=0 C=1 T=2 G=3
 *     (include/kmer.hpp:194); one little-endian uint64 per k-mer when the dictionary was opened
 *     with max_k = 31, two (low word first) when max_k = 63 (include/kmer.hpp:304-308);
 *   - "not found" is kmer_id == UINT64_MAX (include/constants.hpp:5);
 *   - the handle is immutable after open: any number of host threads may issue batches on
 *     distinct streams concurrently (mirrors the const, re-entrant reference lookup,
 *     include/spectrum_preserving_string_set.hpp:37-39).
 * There is NO CPU fallback: without a CUDA device every compute entry point fails with
 * SSHASH_GPU_ECUDA.
 */
#ifndef SSHASH_GPU_H
#define SSHASH_GPU_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SSHASH_GPU_INVALID UINT64_MAX

#if defined(__GNUC__)
#define SSHASH_GPU_API __attribute__((visibility("default")))
#else
#define SSHASH_GPU_API
#endif

typedef enum {
    SSHASH_GPU_OK = 0,
    SSHASH_GPU_EINVAL = 1,   /* bad argument */
    SSHASH_GPU_EIO = 2,      /* cannot open / read the file (essentials.hpp:413-417) */
    SSHASH_GPU_EFORMAT = 3,  /* malformed or unsupported index file */
    SSHASH_GPU_EVERSION = 4, /* MAJOR index version mismatch (include/util.hpp:191-195) */
    SSHASH_GPU_ECUDA = 5,    /* CUDA runtime error / no device */
    SSHASH_GPU_ENOMEM = 6
} sshash_gpu_status;

typedef struct sshash_gpu_dict sshash_gpu_dict;

/* lookup_result, include/util.hpp:38-62 (bool widened to 8 bytes; 64 bytes per record) */
typedef struct {
    uint64_t kmer_id;
    uint64_t kmer_id_in_string;
    uint64_t kmer_offset;
    int64_t kmer_orientation; /* +1 forward, -1 backward (include/constants.hpp:19-20) */
    uint64_t string_id;
    uint64_t string_begin;
    uint64_t string_end;
    uint64_t minimizer_found;
} sshash_lookup_result;

/* streaming_query_report, include/util.hpp:21-36 */
typedef struct {
    uint64_t num_kmers;
    uint64_t num_positive_kmers;
    uint64_t num_negative_kmers;
    uint64_t num_invalid_kmers;
    uint64_t num_searches;
    uint64_t num_extensions;
} sshash_streaming_report;

/* accessors of dictionary (include/dictionary.hpp:31-38) plus placement facts */
typedef struct {
    uint64_t num_kmers;
    uint64_t num_strings;
    uint64_t k;
    uint64_t m;
    uint64_t canonical;
    uint64_t weighted;
    uint64_t max_k;           /* 31 or 63: k-mer word width the index was built with */
    uint64_t version;         /* x<<16 | y<<8 | z */
    uint64_t num_minimizers;
    uint64_t mphf_partitions;
    uint64_t skew_partitions;
    uint64_t index_file_bytes;
    uint64_t device_bytes;    /* HBM resident bytes of the device mirrors */
    int64_t device;
} sshash_gpu_info_t;

/* thread-local description of the last error returned on this thread */
SSHASH_GPU_API const char* sshash_gpu_last_error(void);

/* library build facts: "sm_100a" etc. */
SSHASH_GPU_API const char* sshash_gpu_build_info(void);

/* number of CUDA kernels this library has launched since it was loaded (diagnostics / bench) */
SSHASH_GPU_API uint64_t sshash_gpu_launch_count(void);

/*
 * Replaces essentials::load(dict, path) / open_dictionary (tools/common.hpp:19-29).
 * Parses the reference's on-disk format 5.x unchanged (include/dictionary.hpp:139-152) and uploads
 * GPU-friendly mirrors of its arrays to `device`.  max_k: 31 or 63 = k-mer width of the reference
 * build that wrote the file (it is not recorded in the file); 0 = infer from k (k <= 31 -> 31).
 */
SSHASH_GPU_API int sshash_gpu_open(const char* index_path, int device, int max_k, sshash_gpu_dict** out);
SSHASH_GPU_API int sshash_gpu_close(sshash_gpu_dict* dict);
SSHASH_GPU_API int sshash_gpu_info(const sshash_gpu_dict* dict, sshash_gpu_info_t* out);

/* How lookup / membership / access / weight treat a device pointer that lives on ANOTHER GPU.
   0 (default): staged chunk by chunk with copy-engine transfers, like a host buffer.  1: used in place --
   the kernels load / store it directly over NVLink; the caller guarantees that it is dereferenceable
   from the dictionary's GPU (peer access enabled, or a peer-mapped symmetric / IPC allocation).  This
   is how a per-GPU process delivers its ids straight into another rank's result vector. */
SSHASH_GPU_API int sshash_gpu_set_peer_inplace(sshash_gpu_dict* dict, int inplace);

/*
 * Batched dictionary::lookup(Kmer uint_kmer, bool check_reverse_complement)
 * (include/dictionary.hpp:42, src/dictionary.cpp:64-78).  kmers: n packed k-mers.
 * kmer_ids (nullable): n ids; full (nullable): n complete lookup_result records.
 */
SSHASH_GPU_API int sshash_gpu_lookup_batch(const sshash_gpu_dict* dict, const uint64_t* kmers, uint64_t n,
                            int check_reverse_complement, uint64_t* kmer_ids,
                            sshash_lookup_result* full, void* stream);

/*
 * The same lookup with 32-bit ids, for dictionaries with fewer than 2^32 - 1 k-mers (every index of
 * up to ~4.29e9 k-mers): kmer_ids32[i] = (uint32_t)lookup(...).kmer_id, "not found" = UINT32_MAX.
 * Halves the id bytes that cross PCIe / NVLink; SSHASH_GPU_EINVAL when the dictionary is larger.
 */
SSHASH_GPU_API int sshash_gpu_lookup_batch_u32(const sshash_gpu_dict* dict, const uint64_t* kmers, uint64_t n,
                                               int check_reverse_complement, uint32_t* kmer_ids32, void* stream);

/*
 * Batched dictionary::lookup(char const* string_kmer, bool) (include/dictionary.hpp:41,
 * src/dictionary.cpp:58-63): n strings of exactly k characters, back to back, no terminators,
 * no validation (non-ACGT bytes alias through (c>>1)&3 exactly as in the reference).
 */
SSHASH_GPU_API int sshash_gpu_lookup_batch_ascii(const sshash_gpu_dict* dict, const char* kmers, uint64_t n,
                                  int check_reverse_complement, uint64_t* kmer_ids,
                                  sshash_lookup_result* full, void* stream);

/* Batched dictionary::is_member (include/dictionary.hpp:75-76, src/dictionary.cpp:80-88):
   member[i] = 1 iff found. */
SSHASH_GPU_API int sshash_gpu_is_member_batch(const sshash_gpu_dict* dict, const uint64_t* kmers, uint64_t n,
                               int check_reverse_complement, uint8_t* member, void* stream);

/* Diagnostics: partitions[i] = the partition of the minimizer MPHF that k-mer i's forward minimizer
   hashes to (external/pthash/include/partitioned_phf.hpp:145-149): the key the partition-major
   lookup path bins queries by. */
SSHASH_GPU_API int sshash_gpu_minimizer_partition_batch(const sshash_gpu_dict* dict, const uint64_t* kmers, uint64_t n,
                                                        uint32_t* partitions, void* stream);

/* Batched dictionary::access(kmer_id, char*) (include/dictionary.hpp:71, src/dictionary.cpp:90-94),
   returning packed k-mers instead of strings.  Ids must be < num_kmers. */
SSHASH_GPU_API int sshash_gpu_access_batch(const sshash_gpu_dict* dict, const uint64_t* kmer_ids, uint64_t n,
                            uint64_t* kmers_out, void* stream);

/* Batched dictionary::weight(kmer_id) (include/dictionary.hpp:65-66, src/dictionary.cpp:96-100,
   include/weights.hpp:148-153): the abundance stored for each k-mer id of a weighted dictionary
   (built with --weighted).  Ids must be < num_kmers.  SSHASH_GPU_EINVAL if the dictionary is not
   weighted (the reference asserts). */
SSHASH_GPU_API int sshash_gpu_weight_batch(const sshash_gpu_dict* dict, const uint64_t* kmer_ids, uint64_t n,
                            uint64_t* weights_out, void* stream);

/*
 * Navigational queries (include/dictionary.hpp:50-66, src/dictionary.cpp:112-201).  For each of the
 * n inputs 8 results are produced: forward[A,C,T,G] then backward[A,C,T,G] (neighbourhood<Kmer>,
 * include/util.hpp:77-81; alphabet order of include/kmer.hpp:115-119).  which: 1 =
 * kmer_forward_neighbours, 2 = kmer_backward_neighbours, 3 = kmer_neighbours; slots not asked for
 * hold the default lookup_result (not found), exactly like the reference's default-constructed
 * arrays.  kmer_ids (nullable) and full (nullable) have 8*n entries.  String ids must be
 * < num_strings.  With device buffers these two calls synchronise `stream` before returning.
 */
SSHASH_GPU_API int sshash_gpu_kmer_neighbours_batch(const sshash_gpu_dict* dict, const uint64_t* kmers, uint64_t n,
                                                    int check_reverse_complement, int which, uint64_t* kmer_ids,
                                                    sshash_lookup_result* full, void* stream);
SSHASH_GPU_API int sshash_gpu_string_neighbours_batch(const sshash_gpu_dict* dict, const uint64_t* string_ids, uint64_t n,
                                                      int check_reverse_complement, uint64_t* kmer_ids,
                                                      sshash_lookup_result* full, void* stream);

/*
 * Does the index keep SSHash's input contract (README: "without duplicate k-mers") -- every
 * k-mer once and, on a regular index, never together with its reverse complement?  Checked on the
 * device (two lookups per text offset), once per handle, lazily at the first streaming call or here.
 * *breaks_contract = 1: streaming over this index replays the reference's state machine literally
 * (include/streaming_query.hpp:56-197), because the reference's own answers then depend on the state
 * of the stream; 0: the alignment / classification shortcuts are exact and are used.
 */
SSHASH_GPU_API int sshash_gpu_check_input_contract(const sshash_gpu_dict* dict, int* breaks_contract);

/*
 * Streaming membership over a batch of reads: replaces streaming_query<Dict,canonical>::lookup
 * driven per read as in streaming_query_from_fastq_file (include/streaming_query.hpp:56-115,
 * src/query.cpp:78-108): reset per read, reads shorter than k skipped, a window containing a
 * non-ACGTacgt byte is invalid.  bases: concatenated read characters; read_offsets: num_reads+1
 * offsets into bases.  kmer_ids (nullable): one id per window, reads in order, sum over reads of
 * max(0, len-k+1) entries.  report: the six counters of streaming_query_report.
 */
SSHASH_GPU_API int sshash_gpu_streaming_batch(const sshash_gpu_dict* dict, const char* bases,
                               const uint64_t* read_offsets, uint64_t num_reads,
                               uint64_t* kmer_ids, sshash_streaming_report* report, void* stream);

/* dictionary::streaming_query_from_file(filename, multiline) (include/dictionary.hpp:81-82,
   src/query.cpp:118-175): .fa/.fasta/.fq/.fastq, optionally .gz. */
SSHASH_GPU_API int sshash_gpu_streaming_query_from_file(const sshash_gpu_dict* dict, const char* filename,
                                         int multiline, sshash_streaming_report* report);

/*
 * Several GPUs of one box behind one handle (BASELINE.json configs[4]).  The reference
 * has one `dictionary` object per process (include/dictionary.hpp:10-181) and shards nothing; here
 * the index is REPLICATED on every listed GPU and each batch is SHARDED by query in contiguous
 * slices, slice j on devices[j] (sizes differ by at most one; results keep the query order).  One
 * host thread per GPU drives its slice through the single-device entry points above:
 *   - host buffers: every GPU moves its own slice over its own PCIe link, the ids land in the
 *     caller's array -- no inter-GPU traffic at all;
 *   - device buffers (on any GPU of the box): the owning GPU works in place, the others stage their
 *     slices with peer-to-peer copy-engine transfers over NVLink, so all ids arrive in the owner's
 *     vector (the "gather of ids" of configs[4]) without using SMs for communication.
 * devices == NULL: ordinals 0 .. n_devices-1; n_devices <= 0: every visible GPU.  Calls return when
 * the outputs are complete.  sshash_gpu_multi_dict(m, i) exposes the i-th replica for the
 * single-device calls (access, weight, neighbours, file streaming ...).
 */
typedef struct sshash_gpu_multi sshash_gpu_multi;
SSHASH_GPU_API int sshash_gpu_multi_open(const char* index_path, const int* devices, int n_devices, int max_k,
                                         sshash_gpu_multi** out);
SSHASH_GPU_API int sshash_gpu_multi_close(sshash_gpu_multi* m);
SSHASH_GPU_API int sshash_gpu_multi_num_devices(const sshash_gpu_multi* m);
SSHASH_GPU_API const sshash_gpu_dict* sshash_gpu_multi_dict(const sshash_gpu_multi* m, int i);
/* dictionary::lookup / is_member over a sharded batch (src/dictionary.cpp:64-88) */
SSHASH_GPU_API int sshash_gpu_multi_lookup_batch(const sshash_gpu_multi* m, const uint64_t* kmers, uint64_t n,
                                                 int check_reverse_complement, uint64_t* kmer_ids);
SSHASH_GPU_API int sshash_gpu_multi_lookup_batch_u32(const sshash_gpu_multi* m, const uint64_t* kmers, uint64_t n,
                                                     int check_reverse_complement, uint32_t* kmer_ids32);
SSHASH_GPU_API int sshash_gpu_multi_is_member_batch(const sshash_gpu_multi* m, const uint64_t* kmers, uint64_t n,
                                                    int check_reverse_complement, uint8_t* member);
/* streaming membership (include/streaming_query.hpp:56-115) over HOST reads sharded by read; the
   six counters are per-read sums, so the shards' reports add up */
SSHASH_GPU_API int sshash_gpu_multi_streaming_batch(const sshash_gpu_multi* m, const char* bases, const uint64_t* read_offsets,
                                                    uint64_t num_reads, uint64_t* kmer_ids, sshash_streaming_report* report);

#ifdef __cplusplus
}
#endif
#endif /* SSHASH_GPU_H */
