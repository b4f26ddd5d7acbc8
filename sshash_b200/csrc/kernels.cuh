// kernels.cuh -- launchers of the sm_100a kernels (kernels.cu).
#pragma once

#include <cuda_runtime.h>

#include <atomic>
#include <cstdint>

#include "../../include/sshash_gpu.h"
#include "device_index.cuh"

namespace sshash_b200 {

// Partition-major path (binned.cu): what one bin (= 2^bin_shift consecutive partitions of the minimizer
// MPHF) touches -- byte ranges inside the pilots pool and the control-codeword vector, 16-byte aligned
struct BinRegion {
    uint64_t pilots_off, pilots_bytes;
    uint64_t cw_off, cw_bytes;
};
struct BinPlan {
    uint32_t n_bins = 0, bin_shift = 0;   // n_bins == 0: the path is unavailable for this index
    const BinRegion* regions = nullptr;   // device array, n_bins entries
    bool enabled = false;                 // chosen at open (SSHASH_GPU_BINNED=0/1 overrides the size rule)
    bool prefetch = true;                 // bulk L2 prefetch of the next bins' regions
    uint32_t lookahead = 1;               // how many bins ahead the prefetch runs
    uint64_t min_queries = 1ull << 22;    // smaller device-resident batches take the direct kernel
    uint64_t window_bytes = 0;            // L2 window of this path: the locate tables only
    float hit_ratio = 1.0f;
};

// per-dictionary launch facts: grid sizing and the L2 access-policy window over the hot slab
struct LaunchCtx {
    int sm_count = 148;
    const void* hot_base = nullptr;
    uint64_t hot_bytes = 0;
    uint64_t hot_prefix_bytes = 0;  // the slab without its last piece (the pilots pool)
    uint64_t window_bytes = 0;      // 0 = no window
    float hit_ratio = 1.0f;
    uint64_t max_window_bytes = 0, max_persist_bytes = 0;
    BinPlan bins;
};

// number of kernels launched by this library since it was loaded (bench.py's `gpu_launches`)
uint64_t kernel_launch_count();

// Batched dictionary::lookup.  `queries`: packed k-mers (ascii = false) or n*k characters.
// Exactly one of {ids and/or full, member, ids32} is produced; all pointers are DEVICE pointers.
// ids32: 32-bit ids (UINT32_MAX = not found) for dictionaries with fewer than 2^32 - 1 k-mers.
cudaError_t launch_lookup(const DeviceIndex& ix, const LaunchCtx& ctx, const void* queries, bool ascii, uint64_t n, bool check_rc,
                          uint64_t* ids, sshash_lookup_result* full, uint8_t* member, cudaStream_t stream, uint32_t* ids32 = nullptr);

cudaError_t launch_access(const DeviceIndex& ix, const LaunchCtx& ctx, const uint64_t* ids, uint64_t n, uint64_t* kmers_out,
                          cudaStream_t stream);

// Navigational queries: expand n inputs (packed k-mers, or string ids when strings = true) into 8n
// neighbour k-mers in `expanded` (8n * kmer_words words), look them up, reset the slots `which`
// (bit 0 forward, bit 1 backward) does not ask for.
cudaError_t launch_neighbours(const DeviceIndex& ix, const LaunchCtx& ctx, const uint64_t* in, bool strings, uint64_t n,
                              bool check_rc, int which, uint64_t* expanded, uint64_t* ids, sshash_lookup_result* full,
                              cudaStream_t stream);

// Partition-major batched lookup (binned.cu): ids (u64), ids32 (u32, UINT32_MAX = not found) or member
// bytes for n <= binned_max_batch() packed k-mers, all DEVICE pointers; `scratch` holds
// binned_scratch_bytes(ix, ctx, n) bytes.  Fully asynchronous on `stream`.
uint64_t binned_max_batch();
uint64_t binned_scratch_bytes(const DeviceIndex& ix, const LaunchCtx& ctx, uint64_t n);
cudaError_t launch_lookup_binned(const DeviceIndex& ix, const LaunchCtx& ctx, const uint64_t* queries, uint64_t n, bool check_rc,
                                 uint64_t* ids, uint32_t* ids32, uint8_t* member, void* scratch, cudaStream_t stream);

// diagnostics: MPHF partition of each query's forward minimizer
cudaError_t launch_minimizer_partition(const DeviceIndex& ix, const LaunchCtx& ctx, const uint64_t* kmers, uint64_t n, uint32_t* out,
                                       cudaStream_t stream);

// dictionary::weight for n k-mer ids (weighted indexes only)
cudaError_t launch_weight(const DeviceIndex& ix, const LaunchCtx& ctx, const uint64_t* ids, uint64_t n, uint64_t* weights_out,
                          cudaStream_t stream);

// open time: range-check every control codeword and bucket offset (verbatim codewords); *flag (zeroed
// by the caller) receives a non-zero bit mask when the file is inconsistent
cudaError_t launch_validate_index(const DeviceIndex& ix, const LaunchCtx& ctx, uint32_t* flag, cudaStream_t stream);

// open time: re-encode ix.codewords (verbatim) into `out` (zeroed, width + fp_bits per entry) with
// a fingerprint of each slot's minimizer above the codeword; `filter` (nullable, zeroed, 2^(32 - filter_shift)
// words) receives the blocked Bloom filter over the same minimizers
cudaError_t launch_build_fingerprints(const DeviceIndex& ix, const LaunchCtx& ctx, uint32_t fp_bits, uint64_t* out,
                                      uint32_t* filter, uint32_t filter_shift, cudaStream_t stream);

// open time: the 16-byte wide entries (codeword + the text around a singleton bucket's offset), one per
// MPHF slot, into `out` (ix.codewords.size entries); see DeviceIndex::wide
cudaError_t launch_build_wide(const DeviceIndex& ix, const LaunchCtx& ctx, void* out, cudaStream_t stream);

// Reads are spans of `bases`: read r = [read_begins[r], read_ends[r]) (contiguous reads:
// read_offsets and read_offsets + 1).
// win_offsets[r] = number of windows in reads [0, r), computed on the device from the spans
cudaError_t launch_window_offsets(uint32_t k, const uint64_t* read_begins, const uint64_t* read_ends, uint64_t num_reads,
                                  uint64_t* win_offsets, uint64_t* block_sums, cudaStream_t stream);
uint64_t window_offsets_scratch_words(uint64_t num_reads);

// Streaming membership over a batch of reads: per-window lookups, then the classification of the
// windows into searches / extensions (one thread per window for 64-bit k-mers, the per-read replay
// of the reference state machine for 128-bit ones).  total_windows_bound >= number of windows.
// counters[5] += {num_kmers, searches, extensions, negative, invalid}.
cudaError_t launch_streaming(const DeviceIndex& ix, const LaunchCtx& ctx, const char* bases, const uint64_t* read_begins,
                             const uint64_t* read_ends, const uint64_t* win_offsets, uint64_t num_reads, void* anchors, uint64_t* win_id, uint64_t* win_aux,
                             uint64_t* ids_out, uint64_t total_windows_bound, unsigned long long* counters, cudaStream_t stream,
                             bool replay = false);
// open-time / first-use check of SSHash's input contract (every k-mer once; regular index: never with its
// reverse complement): *flag (zeroed by the caller) becomes non-zero when the index breaks it.  Streaming
// over such an index must pass replay = true above (no anchors, the reference's state machine replayed
// literally), because the reference's own answers then depend on the state of the stream.
cudaError_t launch_distinct_check(const DeviceIndex& ix, const LaunchCtx& ctx, uint32_t* flag, cudaStream_t stream);
// scratch for the per-read alignment anchors (pass nullptr as `anchors` to look every window up)
uint64_t streaming_anchor_bytes(uint64_t num_reads);

// Device-side FASTA/FASTQ record parsing over a chunk of raw file bytes (16-byte padded buffer).
//   launch_count_lines: tile_counts has parse_tiles(n) + 1 entries, the last one zeroed by the caller;
//     on return entry [t] = newlines before tile t and entry [parse_tiles(n)] = number of newlines.
//   launch_read_spans: line_start (lines + 1 entries) and the spans of the sequence lines of
//     num_records records of lines_per_record lines each (FASTQ 4, FASTA 2).
uint64_t parse_tiles(uint64_t n_bytes);
cudaError_t launch_count_lines(const uint8_t* raw, uint64_t n_bytes, uint64_t* tile_counts, cudaStream_t stream);
cudaError_t launch_read_spans(const uint8_t* raw, uint64_t n_bytes, const uint64_t* tile_offsets, uint64_t* line_start,
                              uint64_t num_records, uint32_t lines_per_record, uint64_t* read_begins, uint64_t* read_ends,
                              int sm_count, cudaStream_t stream);

}  // namespace sshash_b200
