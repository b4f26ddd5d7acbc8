// binned.cu -- the PARTITION-MAJOR lookup path: queries are binned on the device by the partition of
// the minimizer MPHF their minimizer hashes to (external/pthash/include/partitioned_phf.hpp:145-149),
// looked up bin by bin, and the ids are put back into query order.
//
// Why.  A lookup on an HBM-resident index is a chain of random accesses: pilot (single_phf.hpp:68-78),
// control codeword (minimizers_control_map.hpp:36-39 -> sparse_and_skew_index.hpp:112-137), strings
// (spss.hpp:213-235), end-points.  The MPHF position of a minimizer is  partition.offset + position
// inside the partition, so BOTH the pilots (~1 MB per partition) and the control codewords (~15 MB
// per partition of ~3e6 minimizers) of one partition are contiguous: when all queries of a partition
// run together, those two accesses become L2 hits (the region is streamed into L2 once, ahead of
// time, by bulk L2 prefetches -- the copy engine of TMA, cp.async.bulk.prefetch.L2) instead of one
// random 64-byte DRAM fetch each.  What stays random is the one access the index layout cannot
// coalesce: the k-mer comparison in `strings`, whose offset is unrelated to the MPHF position.
//
// Pipeline (all kernels asynchronous on one stream, no host round trip; exact, no overflow paths):
//   A1  bin_count_kernel    minimizer + CityHash + partition per query -> meta1[i] = bin | pos,
//                           histogram counts[range][bin] (range = 2^20 consecutive query indices)
//   A2  bin_scan_kernel     exclusive scan in (bin, range) order -> every (range, bin) sub-run's slot
//   A3  bin_scatter_kernel  records {k-mer, idx | pos | bin} written bin-major: a tile of 8192 records is
//                           sorted by bin in shared memory and leaves as coalesced runs (one global atomic
//                           per (tile, bin))
//   B   lookup_binned_kernel  warps claim 128 consecutive records; each record is ONE pass of the
//                           reference's lookup with the minimizer given (device_index.cuh); result ids
//                           stored in record order; on a regular index misses are counted per sub-run
//   C   unpermute_kernel    CTAs claim 2048-record chunks of the (range, bin) sub-runs in RANGE-major order,
//                           so the whole grid works inside one or two ranges at a time: hits are stored to
//                           ids[idx] (an 8 MB window of ids per range: the scattered stores merge in L2);
//                           misses of a regular index with check_reverse_complement are appended,
//                           reverse-complemented, to the round-2 list (src/dictionary.cpp:71-76) together
//                           with their round-2 bin (A1 of round 2 is fused in here)
//   round 2 = A2..C over the miss list; what still misses is stored as "not found".
// Canonical indexes take one round (src/dictionary.cpp:24-42); the minimizer tie case runs its
// second attempt inline.
#include <algorithm>

#include "kernels.cuh"
#include "launch.cuh"

namespace sshash_b200 {

namespace {

constexpr int kTile = 2048;                    // records per counting tile (kBlock threads x 8)
constexpr int kTileItems = kTile / kBlock;
constexpr int kSortTile = 8192;                // records per scatter tile (sorted in shared memory)
constexpr int kSortThreads = 1024;
constexpr int kSortItems = kSortTile / kSortThreads;
constexpr int kChunk = 2048;                   // records per un-permute chunk
constexpr int kChunkItems = kChunk / kBlock;
constexpr uint32_t kRangeShift = 20;           // 2^20 query indices per output range (8 MB of u64 ids)
constexpr uint32_t kPadIdx = 0xffffffffu;      // round-2 list: padding slot (ranges start on scatter-tile boundaries)
constexpr int kClaimItems = 4;                 // records per lane and claim in phase B
constexpr uint32_t kClaim = 32 * kClaimItems;

// meta1 (u32): bin [0,16) | minimizer pos [16,22) | strand (canonical: minimizer taken from the rc) 22 | tie 23
// record meta (u64): idx [0,32) | (meta1 >> 16) [32,40) | bin [40,56)

struct Control {                 // device-resident control block of one round (u32 unless noted); s = range * n_bins + bin
    uint32_t* counts;            // [s]   records per sub-run
    uint32_t* base;              // [s]   first slot of the sub-run in the bin-major record arrays
    uint32_t* cursor;            // [s]   scatter cursor (starts at base)
    uint32_t* chunk_start;       // [s+1] first un-permute chunk of the sub-run (range-major order)
    uint32_t* miss_counts;       // [s]   misses per sub-run (round 1 of a regular index)
    uint32_t* mcursor;           // [s]   next free slot of the sub-run's misses in the round-2 list
    uint32_t* bin_start;         // [n_bins + 1]
    uint64_t* mstart;            // [n_ranges + 1] first slot of every range in the round-2 list; [n_ranges] = its length
    unsigned long long* claims;  // [0] phase B record cursor, [1] phase C chunk cursor
};

__device__ __forceinline__ uint32_t range_of_tile(uint64_t first_record, const uint64_t* __restrict__ mstart, uint32_t n_ranges) {
    if (!mstart) return (uint32_t)(first_record >> kRangeShift);
    uint32_t lo = 0, hi = n_ranges;                    // largest r with mstart[r] <= first_record (empty ranges share a start)
    while (hi - lo > 1) {
        const uint32_t mid = (lo + hi) / 2;
        if (mstart[mid] <= first_record) lo = mid; else hi = mid;
    }
    return lo;
}

// exclusive scan of one value per thread over a CTA of NT threads (NT a multiple of 32, <= 1024)
template <int NT>
__device__ __forceinline__ uint32_t cta_exclusive_scan(uint32_t v, uint32_t* warp_sums /* NT / 32 + 1 */) {
    const uint32_t lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    uint32_t inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= (uint32_t)o) inc += t; }
    if (lane == 31) warp_sums[wid] = inc;
    __syncthreads();
    if (wid == 0) {
        uint32_t w = lane < NT / 32 ? warp_sums[lane] : 0;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, w, o); if (lane >= (uint32_t)o) w += t; }
        if (lane < NT / 32) warp_sums[lane] = w;        // inclusive
        if (lane == 31) warp_sums[NT / 32] = w;         // grand total (lanes >= NT/32 added zeros)
    }
    __syncthreads();
    const uint32_t before = wid ? warp_sums[wid - 1] : 0;
    return before + inc - v;
}

// bin | pos | flags of one query k-mer (A1; also run by C for the round-2 list)
template <int W, bool CANON>
__device__ __forceinline__ uint32_t bin_meta_of(const DeviceIndex& ix, Kmer<W> x, uint32_t bin_shift) {
    Minimizer mi = compute_minimizer(ix, x);
    uint32_t flags = 0;
    if (CANON) {                                       // src/dictionary.cpp:24-42: the smaller minimizer decides
        const Minimizer mr = compute_minimizer(ix, kmer_rc(x, ix.k));
        if (mr.value < mi.value) { mi = mr; flags = 1u << 6; }
        else if (mr.value == mi.value) flags = 1u << 7;          // tie: forward info first, then the rc info
    }
    const uint32_t bin = mphf_partition(ix.mphf, city_hash_u64(ix.mphf, mi.value)) >> bin_shift;
    return bin | ((mi.pos | flags) << 16);
}

// ---- A1 -------------------------------------------------------------------------------------------
template <int W, bool CANON>
__global__ void __launch_bounds__(kBlock)
bin_count_kernel(const __grid_constant__ DeviceIndex ix, const uint64_t* __restrict__ kmers, uint64_t n_records,
                 uint32_t n_bins, uint32_t bin_shift, uint32_t* __restrict__ meta1, uint32_t* __restrict__ counts) {
    extern __shared__ uint32_t hist[];
    const uint64_t n_tiles = (n_records + kTile - 1) / kTile;
    for (uint64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        for (uint32_t b = threadIdx.x; b < n_bins; b += kBlock) hist[b] = 0;
        __syncthreads();
#pragma unroll 1
        for (int t = 0; t < kTileItems; ++t) {
            const uint64_t i = tile * kTile + (uint64_t)t * kBlock + threadIdx.x;
            if (i >= n_records) continue;
            const uint32_t meta = bin_meta_of<W, CANON>(ix, load_kmer<W>(kmers, i), bin_shift);
            atomicAdd(&hist[meta & 0xffffu], 1u);
            meta1[i] = meta;
        }
        __syncthreads();
        const uint32_t r = (uint32_t)((tile * kTile) >> kRangeShift);
        for (uint32_t b = threadIdx.x; b < n_bins; b += kBlock)
            if (hist[b]) atomicAdd(&counts[(uint64_t)r * n_bins + b], hist[b]);
        __syncthreads();
    }
}

// ---- A2 (one CTA of kSortThreads) -------------------------------------------------------------------
__global__ void __launch_bounds__(kSortThreads)
bin_scan_kernel(Control c, uint32_t n_ranges, uint32_t n_bins) {
    extern __shared__ uint32_t sh[];                   // totals[n_bins + 1] + scan scratch [kSortThreads / 32 + 1]
    uint32_t* totals = sh;
    uint32_t* scratch = sh + n_bins + 1;
    // per-bin totals, then their exclusive scan (n_bins <= kSortThreads)
    uint32_t mine = 0;
    if (threadIdx.x < n_bins)
        for (uint32_t r = 0; r < n_ranges; ++r) mine += c.counts[(uint64_t)r * n_bins + threadIdx.x];
    const uint32_t ex = cta_exclusive_scan<kSortThreads>(mine, scratch);
    if (threadIdx.x < n_bins) totals[threadIdx.x] = ex;
    if (threadIdx.x == 0) totals[n_bins] = scratch[kSortThreads / 32];
    __syncthreads();
    for (uint32_t b = threadIdx.x; b <= n_bins; b += kSortThreads) c.bin_start[b] = totals[b];
    if (threadIdx.x < n_bins) {
        uint32_t run = totals[threadIdx.x];
        for (uint32_t r = 0; r < n_ranges; ++r) {
            const uint64_t s = (uint64_t)r * n_bins + threadIdx.x;
            c.base[s] = run; c.cursor[s] = run;
            run += c.counts[s];
        }
    }
    __syncthreads();
    // chunk_start: exclusive scan of ceil(counts / kChunk) over the sub-runs in range-major order
    const uint64_t runs = (uint64_t)n_ranges * n_bins;
    const uint64_t per = (runs + kSortThreads - 1) / kSortThreads, s0 = per * threadIdx.x, s1 = s0 + per < runs ? s0 + per : runs;
    uint32_t local = 0;
    for (uint64_t s = s0; s < s1; ++s) local += (c.counts[s] + kChunk - 1) / kChunk;
    uint32_t run = cta_exclusive_scan<kSortThreads>(local, scratch);
    for (uint64_t s = s0; s < s1; ++s) { c.chunk_start[s] = run; run += (c.counts[s] + kChunk - 1) / kChunk; }
    if (threadIdx.x == kSortThreads - 1) c.chunk_start[runs] = scratch[kSortThreads / 32];
}

// ---- A3 -------------------------------------------------------------------------------------------
__device__ __forceinline__ void store_plain(uint64_t* out, uint64_t i, Kmer<1> x) { out[i] = x.lo; }
__device__ __forceinline__ void store_plain(uint64_t* out, uint64_t i, Kmer<2> x) {
    reinterpret_cast<ulonglong2*>(out)[i] = make_ulonglong2(x.lo, x.hi);
}
__device__ __forceinline__ Kmer<1> load_plain(const uint64_t* in, uint64_t i, Kmer<1>*) { return {in[i]}; }
__device__ __forceinline__ Kmer<2> load_plain(const uint64_t* in, uint64_t i, Kmer<2>*) {
    const ulonglong2 v = reinterpret_cast<const ulonglong2*>(in)[i];
    return {v.x, v.y};
}

// src_idx == nullptr: record i is query i (round 1); otherwise the round-2 list (kPadIdx = padding slot)
template <int W>
__global__ void __launch_bounds__(kSortThreads, 1)
bin_scatter_kernel(const uint64_t* __restrict__ kmers, const uint32_t* __restrict__ src_idx, uint64_t n_records_arg,
                   const uint64_t* __restrict__ mstart, uint32_t n_ranges, uint32_t n_bins, const uint32_t* __restrict__ meta1,
                   uint32_t* __restrict__ cursor, uint64_t* __restrict__ rec_kmer, uint64_t* __restrict__ rec_meta) {
    extern __shared__ __align__(16) uint8_t smem[];
    uint64_t* s_kmer = reinterpret_cast<uint64_t*>(smem);                        // kSortTile * W words
    uint64_t* s_meta = s_kmer + (size_t)kSortTile * W;                           // kSortTile
    uint32_t* hist = reinterpret_cast<uint32_t*>(s_meta + kSortTile);            // n_bins
    uint32_t* toff = hist + n_bins;                                              // n_bins
    uint32_t* gbase = toff + n_bins;                                             // n_bins
    uint32_t* scratch = gbase + n_bins;                                          // kSortThreads / 32 + 1
    __shared__ uint32_t s_range;
    const uint64_t n_records = mstart ? mstart[n_ranges] : n_records_arg;
    const uint64_t n_tiles = (n_records + kSortTile - 1) / kSortTile;
    for (uint64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        for (uint32_t b = threadIdx.x; b < n_bins; b += kSortThreads) hist[b] = 0;
        if (threadIdx.x == 0) s_range = range_of_tile(tile * kSortTile, mstart, n_ranges);
        __syncthreads();
        uint32_t meta[kSortItems], rank[kSortItems];
#pragma unroll
        for (int t = 0; t < kSortItems; ++t) {
            const uint64_t i = tile * kSortTile + (uint64_t)t * kSortThreads + threadIdx.x;
            meta[t] = 0xffffffffu;
            if (i < n_records && (!src_idx || src_idx[i] != kPadIdx)) meta[t] = meta1[i];
            rank[t] = meta[t] != 0xffffffffu ? atomicAdd(&hist[meta[t] & 0xffffu], 1u) : 0u;
        }
        __syncthreads();
        // tile-local offsets of the bins (n_bins <= kSortThreads) + one global reservation per (tile, bin)
        const uint32_t mine = threadIdx.x < n_bins ? hist[threadIdx.x] : 0;
        const uint32_t ex = cta_exclusive_scan<kSortThreads>(mine, scratch);
        if (threadIdx.x < n_bins) {
            toff[threadIdx.x] = ex;
            gbase[threadIdx.x] = mine ? atomicAdd(&cursor[(uint64_t)s_range * n_bins + threadIdx.x], mine) : 0;
        }
        const uint32_t total = scratch[kSortThreads / 32];
        __syncthreads();
#pragma unroll
        for (int t = 0; t < kSortItems; ++t) {
            if (meta[t] == 0xffffffffu) continue;
            const uint64_t i = tile * kSortTile + (uint64_t)t * kSortThreads + threadIdx.x;
            const uint32_t bin = meta[t] & 0xffffu, p = toff[bin] + rank[t];
            const uint32_t idx = src_idx ? src_idx[i] : (uint32_t)i;
            store_plain(s_kmer, p, load_kmer<W>(kmers, i));
            s_meta[p] = (uint64_t)idx | ((uint64_t)(meta[t] >> 16) << 32) | ((uint64_t)bin << 40);
        }
        __syncthreads();
        for (uint32_t p = threadIdx.x; p < total; p += kSortThreads) {     // consecutive p of a bin -> consecutive slots
            const uint64_t m = s_meta[p];
            const uint32_t bin = (uint32_t)(m >> 40) & 0xffffu;
            const uint64_t dest = (uint64_t)gbase[bin] + (p - toff[bin]);
            store_plain(rec_kmer, dest, load_plain(s_kmer, p, (Kmer<W>*)nullptr));
            rec_meta[dest] = m;
        }
        __syncthreads();
    }
}

// ---- B --------------------------------------------------------------------------------------------
// Bulk L2 prefetch (the TMA unit's copy-less form): [p, p + bytes) is pulled into L2 as one streaming
// transfer.  p 16-byte aligned, bytes a multiple of 16.
__device__ __forceinline__ void prefetch_l2_bulk(const void* p, uint32_t bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}

// piece `chunk` of `n_chunks` of the region [base, base + bytes), in 16 KB bulk prefetches
__device__ __forceinline__ void prefetch_piece(const uint8_t* base, uint64_t bytes, uint64_t chunk, uint64_t n_chunks) {
    uint64_t lo = (bytes * chunk / n_chunks) & ~127ull, hi = (bytes * (chunk + 1) / n_chunks) & ~127ull;
    if (chunk + 1 == n_chunks) hi = bytes & ~15ull;
    while (lo < hi) {
        const uint32_t step = (uint32_t)(hi - lo < 16384 ? hi - lo : 16384);
        prefetch_l2_bulk(base + lo, step);
        lo += step;
    }
}

template <int W, bool CANON>
__global__ void __launch_bounds__(kBlock, 6)
lookup_binned_kernel(const __grid_constant__ DeviceIndex ix, const uint64_t* __restrict__ rec_kmer, const uint64_t* __restrict__ rec_meta,
                     uint32_t n_bins, const uint32_t* __restrict__ bin_start, const BinRegion* __restrict__ regions, uint32_t lookahead,
                     uint64_t* __restrict__ res_id, uint32_t* __restrict__ miss_counts, unsigned long long* __restrict__ claim) {
    const uint32_t lane = threadIdx.x & 31;
    const uint64_t n_records = bin_start[n_bins];
    for (;;) {
        unsigned long long first = 0;
        if (lane == 0) first = atomicAdd(claim, (unsigned long long)kClaim);
        first = __shfl_sync(0xffffffffu, first, 0);
        if (first >= n_records) break;
        if (regions && lane == 0) {
            // this claim's share of the prefetch of bin + lookahead: the bin's pilots and control codewords
            uint32_t lo = 0, hi = n_bins;                   // bin of the first record: largest b with bin_start[b] <= first
            while (hi - lo > 1) { const uint32_t mid = (lo + hi) / 2; if (bin_start[mid] <= first) lo = mid; else hi = mid; }
            const uint32_t tgt = lo + lookahead;
            if (tgt < n_bins) {
                const uint64_t b0 = bin_start[lo], cnt = bin_start[lo + 1] - b0;
                const uint64_t n_chunks = (cnt + kClaim - 1) / kClaim, chunk = (first - b0) / kClaim;
                const BinRegion rg = regions[tgt];
                if (chunk < n_chunks) {
                    prefetch_piece(reinterpret_cast<const uint8_t*>(ix.pilots) + rg.pilots_off, rg.pilots_bytes, chunk, n_chunks);
                    prefetch_piece(reinterpret_cast<const uint8_t*>(ix.codewords.data) + rg.cw_off, rg.cw_bytes, chunk, n_chunks);
                }
            }
        }
#pragma unroll 1
        for (int t = 0; t < kClaimItems; ++t) {
            const uint64_t j = first + (uint64_t)t * 32 + lane;
            const bool active = j < n_records;
            bool found = false;
            uint64_t meta = 0;
            LookupResult res;
            res.kmer_id = ~0ull;
            if (active) {
                const Kmer<W> x = load_kmer<W>(rec_kmer, j);
                meta = __ldcs(rec_meta + j);
                const uint32_t pos = (uint32_t)(meta >> 32) & 63u;
                if (CANON) {
                    const Kmer<W> xr = kmer_rc(x, ix.k);
                    const bool from_rc = (meta >> 38) & 1;
                    Minimizer mi{kmer_bits_at(from_rc ? xr : x, 2 * pos) & ix.mmer_mask, pos};
                    found = lookup_canonical_with<W, false, false, true>(ix, x, xr, mi, res);
                    if (!found && ((meta >> 39) & 1)) {       // tie: the rc info is tried second (dictionary.cpp:35-41)
                        mi = compute_minimizer(ix, xr);
                        found = lookup_canonical_with<W, false, false, true>(ix, x, xr, mi, res);
                    }
                } else {
                    const Minimizer mi{kmer_bits_at(x, 2 * pos) & ix.mmer_mask, pos};
                    found = lookup_regular_with<W, false, false, true>(ix, x, mi, res);
                }
                __stcs(res_id + j, found ? res.kmer_id : ~0ull);
            }
            if (miss_counts) {                                   // warp-uniform
                const bool miss = active && !found;
                const uint32_t s = miss ? ((uint32_t)meta >> kRangeShift) * n_bins + (uint32_t)((meta >> 40) & 0xffffu) : 0xffffffffu;
                const uint32_t peers = __match_any_sync(0xffffffffu, s);
                if (miss && lane == (uint32_t)__ffs(peers) - 1) atomicAdd(&miss_counts[s], (uint32_t)__popc(peers));
            }
        }
    }
}

// ---- between B and C of round 1 (one CTA): slots of the round-2 list -------------------------------
__global__ void __launch_bounds__(kBlock)
miss_scan_kernel(Control c, uint32_t n_ranges, uint32_t n_bins) {
    for (uint32_t r = threadIdx.x; r < n_ranges; r += kBlock) {
        uint64_t s = 0;
        for (uint32_t b = 0; b < n_bins; ++b) s += c.miss_counts[(uint64_t)r * n_bins + b];
        c.mstart[r] = s;                                      // the range's total for now
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        uint64_t run = 0;
        for (uint32_t r = 0; r < n_ranges; ++r) {
            const uint64_t t = c.mstart[r];
            c.mstart[r] = run;
            run += (t + kSortTile - 1) / kSortTile * kSortTile;    // every range starts on a scatter-tile boundary
        }
        c.mstart[n_ranges] = run;
    }
    __syncthreads();
    for (uint32_t r = threadIdx.x; r < n_ranges; r += kBlock) {
        uint32_t run = (uint32_t)c.mstart[r];
        for (uint32_t b = 0; b < n_bins; ++b) {
            const uint64_t s = (uint64_t)r * n_bins + b;
            c.mcursor[s] = run;
            run += c.miss_counts[s];
        }
    }
}

// ---- C --------------------------------------------------------------------------------------------
// MODE 0: u64 ids, 2: membership bytes, 3: u32 ids.  SECOND: a round follows -- misses go to its list
// (reverse-complemented, with their bin in that round: next_meta1 / next_counts) instead of the output.
template <int W, int MODE, bool SECOND>
__global__ void __launch_bounds__(kBlock)
unpermute_kernel(const __grid_constant__ DeviceIndex ix, const uint64_t* __restrict__ rec_kmer, const uint64_t* __restrict__ rec_meta,
                 const uint64_t* __restrict__ res_id, Control c, uint32_t n_ranges, uint32_t n_bins, uint32_t bin_shift,
                 void* __restrict__ out, uint64_t* __restrict__ miss_kmer, uint32_t* __restrict__ miss_idx,
                 uint32_t* __restrict__ next_meta1, uint32_t* __restrict__ next_counts) {
    extern __shared__ uint32_t hist[];                    // SECOND: n_bins counters of the chunk's misses per round-2 bin
    __shared__ uint32_t s_run, s_first, s_cnt, s_slot, scratch[kBlock / 32 + 1];
    const uint64_t runs = (uint64_t)n_ranges * n_bins;
    const uint32_t n_chunks = c.chunk_start[runs];
    for (;;) {
        __syncthreads();
        if (threadIdx.x == 0) {
            const unsigned long long ch = atomicAdd(c.claims + 1, 1ull);
            if (ch >= n_chunks) s_cnt = 0xffffffffu;
            else {
                uint64_t lo = 0, hi = runs;                 // sub-run of the chunk: largest s with chunk_start[s] <= ch
                while (hi - lo > 1) { const uint64_t mid = (lo + hi) / 2; if (c.chunk_start[mid] <= ch) lo = mid; else hi = mid; }
                // sub-runs without records share a chunk_start with their successor: take the last one of the tie
                const uint32_t off = ((uint32_t)ch - c.chunk_start[lo]) * kChunk, cnt = c.counts[lo];
                s_run = (uint32_t)lo;
                s_first = c.base[lo] + off;
                s_cnt = cnt - off < (uint32_t)kChunk ? cnt - off : (uint32_t)kChunk;
            }
        }
        if (SECOND) for (uint32_t b = threadIdx.x; b < n_bins; b += kBlock) hist[b] = 0;
        __syncthreads();
        if (s_cnt == 0xffffffffu) break;
        const uint32_t first = s_first, cnt = s_cnt, run = s_run;
        uint32_t n_miss = 0, idxs[kChunkItems];
        bool miss[kChunkItems];
#pragma unroll
        for (int t = 0; t < kChunkItems; ++t) {
            const uint32_t o = (uint32_t)t * kBlock + threadIdx.x;
            miss[t] = false;
            idxs[t] = 0;
            if (o < cnt) {
                const uint64_t j = (uint64_t)first + o;
                const uint64_t id = __ldcs(res_id + j);
                const uint32_t idx = (uint32_t)__ldcs(rec_meta + j);
                const bool hit = id != ~0ull;
                idxs[t] = idx;
                if (hit || !SECOND) {
                    if (MODE == 2) static_cast<uint8_t*>(out)[idx] = hit;
                    else if (MODE == 3) static_cast<uint32_t*>(out)[idx] = (uint32_t)id;   // not found: UINT32_MAX
                    else static_cast<uint64_t*>(out)[idx] = id;
                } else { miss[t] = true; ++n_miss; }
            }
        }
        if (SECOND) {
            // slots of this chunk's misses: one reservation per chunk from the sub-run's cursor
            const uint32_t before = cta_exclusive_scan<kBlock>(n_miss, scratch);
            if (threadIdx.x == kBlock - 1) s_slot = (before + n_miss) ? atomicAdd(&c.mcursor[run], before + n_miss) : 0;
            __syncthreads();
            uint32_t slot = s_slot + before;
#pragma unroll
            for (int t = 0; t < kChunkItems; ++t) {
                if (!miss[t]) continue;
                const uint64_t j = (uint64_t)first + (uint32_t)t * kBlock + threadIdx.x;
                const Kmer<W> xr = kmer_rc(load_kmer<W>(rec_kmer, j), ix.k);               // src/dictionary.cpp:72
                const uint32_t meta = bin_meta_of<W, false>(ix, xr, bin_shift);
                store_kmer(miss_kmer, slot, xr);
                miss_idx[slot] = idxs[t];
                next_meta1[slot] = meta;
                atomicAdd(&hist[meta & 0xffffu], 1u);
                ++slot;
            }
            __syncthreads();
            const uint32_t r = run / n_bins;
            for (uint32_t b = threadIdx.x; b < n_bins; b += kBlock)
                if (hist[b]) atomicAdd(&next_counts[(uint64_t)r * n_bins + b], hist[b]);
        }
    }
}

uint64_t align_up(uint64_t x, uint64_t a) { return (x + a - 1) / a * a; }

struct Plan {                      // carve-up of the scratch buffer for a batch of n queries
    uint32_t n_ranges, n_bins;
    uint64_t cap;                  // records any array holds: n + one scatter tile of padding per range
    uint64_t runs;                 // n_ranges * n_bins
    uint64_t off_meta1, off_rec_kmer, off_rec_meta, off_res, off_miss_kmer, off_miss_idx, off_ctl[2], ctl_bytes, total;
};

Plan make_plan(uint32_t kmer_words, uint32_t n_bins, uint64_t n) {
    Plan p{};
    p.n_bins = n_bins;
    p.n_ranges = (uint32_t)((n + (1ull << kRangeShift) - 1) >> kRangeShift);
    p.cap = align_up(n, kSortTile) + (uint64_t)p.n_ranges * kSortTile;
    p.runs = (uint64_t)p.n_ranges * n_bins;
    uint64_t o = 0;
    auto take = [&](uint64_t bytes) { const uint64_t at = o; o += align_up(bytes, 256); return at; };
    p.off_meta1 = take(p.cap * 4);
    p.off_rec_kmer = take(p.cap * 8 * kmer_words);
    p.off_rec_meta = take(p.cap * 8);
    p.off_res = take(p.cap * 8);
    p.off_miss_kmer = take(p.cap * 8 * kmer_words);
    p.off_miss_idx = take(p.cap * 4);
    // control block: 6 arrays of runs (+1) u32, bin_start (n_bins + 1), mstart (n_ranges + 1, u64), claims (2 x u64)
    p.ctl_bytes = align_up((6 * p.runs + 1 + n_bins + 1) * 4, 8) + (p.n_ranges + 1) * 8 + 16;
    p.off_ctl[0] = take(p.ctl_bytes);
    p.off_ctl[1] = take(p.ctl_bytes);
    p.total = o;
    return p;
}

Control control_at(uint8_t* base, const Plan& p) {
    Control c{};
    uint32_t* u = reinterpret_cast<uint32_t*>(base);
    c.counts = u; c.base = u + p.runs; c.cursor = u + 2 * p.runs; c.miss_counts = u + 3 * p.runs; c.mcursor = u + 4 * p.runs;
    c.chunk_start = u + 5 * p.runs;                        // runs + 1 entries
    c.bin_start = u + 6 * p.runs + 1;
    uint8_t* q = base + align_up((6 * p.runs + 1 + p.n_bins + 1) * 4, 8);
    c.mstart = reinterpret_cast<uint64_t*>(q);
    c.claims = reinterpret_cast<unsigned long long*>(q + (p.n_ranges + 1) * 8);
    return c;
}

size_t scatter_smem_bytes(uint32_t kmer_words, uint32_t n_bins) {
    return (size_t)kSortTile * 8 * (kmer_words + 1) + (3 * (size_t)n_bins + kSortThreads / 32 + 1) * 4;
}

}  // namespace

uint64_t binned_max_batch() { return 1ull << 27; }

uint64_t binned_scratch_bytes(const DeviceIndex& ix, const LaunchCtx& ctx, uint64_t n) {
    return make_plan(ix.kmer_words, ctx.bins.n_bins, n).total;
}

cudaError_t launch_lookup_binned(const DeviceIndex& ix, const LaunchCtx& ctx, const uint64_t* queries, uint64_t n, bool check_rc,
                                 uint64_t* ids, uint32_t* ids32, uint8_t* member, void* scratch, cudaStream_t stream) {
    if (n == 0) return cudaSuccess;
    if (n > binned_max_batch() || !ctx.bins.n_bins || ctx.bins.n_bins > (uint32_t)kSortThreads) return cudaErrorInvalidValue;
    const Plan p = make_plan(ix.kmer_words, ctx.bins.n_bins, n);
    uint8_t* s = static_cast<uint8_t*>(scratch);
    uint32_t* meta1 = reinterpret_cast<uint32_t*>(s + p.off_meta1);
    uint64_t* rec_kmer = reinterpret_cast<uint64_t*>(s + p.off_rec_kmer);
    uint64_t* rec_meta = reinterpret_cast<uint64_t*>(s + p.off_rec_meta);
    uint64_t* res = reinterpret_cast<uint64_t*>(s + p.off_res);
    uint64_t* miss_kmer = reinterpret_cast<uint64_t*>(s + p.off_miss_kmer);
    uint32_t* miss_idx = reinterpret_cast<uint32_t*>(s + p.off_miss_idx);
    const bool canon = ix.canonical != 0, two_rounds = !canon && check_rc;
    const int mode = member ? 2 : (ids32 ? 3 : 0);
    void* out = member ? static_cast<void*>(member) : ids32 ? static_cast<void*>(ids32) : static_cast<void*>(ids);
    const uint32_t nb = p.n_bins, nr = p.n_ranges, shift = ctx.bins.bin_shift;
    const BinRegion* regions = ctx.bins.prefetch ? ctx.bins.regions : nullptr;
    const int sm = ctx.sm_count;
    const bool w1 = ix.kmer_words == 1;
    cudaError_t e = cudaMemsetAsync(s + p.off_ctl[0], 0, two_rounds ? 2 * align_up(p.ctl_bytes, 256) : p.ctl_bytes, stream);
    if (e != cudaSuccess) return e;
    if (two_rounds) {
        e = cudaMemsetAsync(miss_idx, 0xff, p.cap * 4, stream);        // every slot is padding until a miss lands in it
        if (e != cudaSuccess) return e;
    }
    const size_t hist_bytes = nb * sizeof(uint32_t), sort_smem = scatter_smem_bytes(ix.kmer_words, nb);
    // opt-in shared memory above 48 KB is a per-device function attribute: set it on every call (microseconds)
    cudaFuncSetAttribute(bin_scatter_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)scatter_smem_bytes(1, kSortThreads));
    cudaFuncSetAttribute(bin_scatter_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)scatter_smem_bytes(2, kSortThreads));
    auto cfg_launch = [&](auto kernel, int grid, int block, size_t smem, auto... args) -> cudaError_t {
        g_launches.fetch_add(1);
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3((unsigned)grid); cfg.blockDim = dim3((unsigned)block); cfg.dynamicSmemBytes = smem; cfg.stream = stream;
        cudaLaunchAttribute attr[1];
        cfg.attrs = attr; cfg.numAttrs = 0;
        if (ctx.bins.window_bytes) {                      // only the locate tables (the slab's prefix) stay persisting in L2:
            attr[0].id = cudaLaunchAttributeAccessPolicyWindow;   // pilots are L2 hits by schedule here
            attr[0].val.accessPolicyWindow.base_ptr = const_cast<void*>(ctx.hot_base);
            attr[0].val.accessPolicyWindow.num_bytes = ctx.bins.window_bytes;
            attr[0].val.accessPolicyWindow.hitRatio = ctx.bins.hit_ratio;
            attr[0].val.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
            attr[0].val.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
            cfg.numAttrs = 1;
        }
        return cudaLaunchKernelEx(&cfg, kernel, args...);
    };
    const Control c0 = control_at(s + p.off_ctl[0], p), c1 = control_at(s + p.off_ctl[1], p);
    // ---- A1 (round 1 only: round 2's bins are computed by C of round 1) ----
    {
        const int grid = (int)std::min<uint64_t>((n + kTile - 1) / kTile, (uint64_t)sm * 8);
        if (canon) e = w1 ? cfg_launch(bin_count_kernel<1, true>, grid, kBlock, hist_bytes, ix, queries, n, nb, shift, meta1, c0.counts)
                          : cfg_launch(bin_count_kernel<2, true>, grid, kBlock, hist_bytes, ix, queries, n, nb, shift, meta1, c0.counts);
        else e = w1 ? cfg_launch(bin_count_kernel<1, false>, grid, kBlock, hist_bytes, ix, queries, n, nb, shift, meta1, c0.counts)
                    : cfg_launch(bin_count_kernel<2, false>, grid, kBlock, hist_bytes, ix, queries, n, nb, shift, meta1, c0.counts);
        if (e != cudaSuccess) return e;
    }
    for (int round = 0; round < (two_rounds ? 2 : 1); ++round) {
        const Control c = round ? c1 : c0;
        const bool r2 = round == 1, second = two_rounds && !r2;
        const uint64_t* src_kmer = r2 ? miss_kmer : queries;
        const uint32_t* src_idx = r2 ? miss_idx : nullptr;
        const uint64_t* mstart = r2 ? c0.mstart : nullptr;
        const uint64_t bound = r2 ? p.cap : n;              // round 2's exact length lives on the device (mstart[n_ranges])
        e = cfg_launch(bin_scan_kernel, 1, kSortThreads, (nb + 1 + kSortThreads / 32 + 1) * sizeof(uint32_t), c, nr, nb);
        if (e != cudaSuccess) return e;
        const int sgrid = (int)std::min<uint64_t>((bound + kSortTile - 1) / kSortTile, (uint64_t)sm);
        e = w1 ? cfg_launch(bin_scatter_kernel<1>, sgrid, kSortThreads, sort_smem, src_kmer, src_idx, n, mstart, nr, nb, (const uint32_t*)meta1, c.cursor, rec_kmer, rec_meta)
               : cfg_launch(bin_scatter_kernel<2>, sgrid, kSortThreads, sort_smem, src_kmer, src_idx, n, mstart, nr, nb, (const uint32_t*)meta1, c.cursor, rec_kmer, rec_meta);
        if (e != cudaSuccess) return e;
        uint32_t* miss_counts = second ? c.miss_counts : nullptr;
        const int lgrid = sm * 6;
#define SSHASH_LOOKUP_BINNED(W, CANON) \
        cfg_launch(lookup_binned_kernel<W, CANON>, lgrid, kBlock, 0, ix, (const uint64_t*)rec_kmer, (const uint64_t*)rec_meta, nb, (const uint32_t*)c.bin_start, regions, ctx.bins.lookahead, res, miss_counts, c.claims)
        if (canon) e = w1 ? SSHASH_LOOKUP_BINNED(1, true) : SSHASH_LOOKUP_BINNED(2, true);
        else e = w1 ? SSHASH_LOOKUP_BINNED(1, false) : SSHASH_LOOKUP_BINNED(2, false);
#undef SSHASH_LOOKUP_BINNED
        if (e != cudaSuccess) return e;
        if (second) {
            e = cfg_launch(miss_scan_kernel, 1, kBlock, 0, c, nr, nb);
            if (e != cudaSuccess) return e;
        }
        const int ugrid = (int)std::min<uint64_t>((bound + kChunk - 1) / kChunk + p.runs, (uint64_t)sm * 8);
#define SSHASH_UNPERMUTE(W, MODE, SECOND) \
        cfg_launch(unpermute_kernel<W, MODE, SECOND>, ugrid, kBlock, SECOND ? hist_bytes : 0, ix, (const uint64_t*)rec_kmer, (const uint64_t*)rec_meta, (const uint64_t*)res, c, nr, nb, shift, out, miss_kmer, miss_idx, meta1, c1.counts)
#define SSHASH_UNPERMUTE_MODE(W, SECOND) (mode == 2 ? SSHASH_UNPERMUTE(W, 2, SECOND) : mode == 3 ? SSHASH_UNPERMUTE(W, 3, SECOND) : SSHASH_UNPERMUTE(W, 0, SECOND))
        if (second) e = w1 ? SSHASH_UNPERMUTE_MODE(1, true) : SSHASH_UNPERMUTE_MODE(2, true);
        else e = w1 ? SSHASH_UNPERMUTE_MODE(1, false) : SSHASH_UNPERMUTE_MODE(2, false);
#undef SSHASH_UNPERMUTE_MODE
#undef SSHASH_UNPERMUTE
        if (e != cudaSuccess) return e;
    }
    return cudaSuccess;
}

}  // namespace sshash_b200
