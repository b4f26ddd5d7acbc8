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
// Pipeline (all kernels asynchronous on one stream, no host round trip; exact, no overflow paths; every
// global store of the reordering steps leaves shared memory as full, coalesced sectors -- scattered
// 8-byte stores were measured at 3.6x write amplification plus read-for-fill on B200's L2,
// profiles/r2_human_binned_v2_scatter_unpermute_ncu_full.txt):
//   A1  bin_count_kernel    minimizer + CityHash + partition per query -> meta1[i] = bin | pos,
//                           histogram counts[range][bin] (range = 2^20 consecutive query indices)
//   A2  bin_scan_kernel     exclusive scan in (bin, range) order -> every (range, bin) sub-run's slot
//   A3  bin_scatter_kernel  one CTA per TILE of 4096 consecutive queries: the tile's records {k-mer, idx |
//                           pos | bin} are sorted by bin in shared memory and leave as coalesced runs
//                           (one global reservation per (tile, bin)); slot_of[] records where every record
//                           of the tile went so that the tile's results can be found again
//   B   lookup_binned_kernel  warps claim 128 consecutive records of the bin-major arrays; each record
//                           is ONE pass of the reference's lookup with the minimizer given
//                           (device_index.cuh); result ids stored in record order
//   C1  gather_tile_kernel  one CTA per tile: the tile's results are collected through slot_of[] (reads run
//                           along the tile's <= n_bins runs) into a shared-memory image of ids[4096 t .. 4096 t + 4096) and stored as
//                           one coalesced block; misses of a regular index with check_reverse_complement
//                           are appended, reverse-complemented, to the round-2 list (src/dictionary.cpp:71-76)
//                           together with their round-2 bin (A1 of round 2 is fused in here)
//   round 2 = A2, A3, B over each tile's misses, then
//   C2  patch_tile_kernel   one CTA per tile: the ids block is read back, patched with the round-2
//                           results and stored again; what still misses stays "not found".
// Canonical indexes take one round (src/dictionary.cpp:24-42); the minimizer tie case runs its
// second attempt inline.
#include <algorithm>

#include "kernels.cuh"
#include "launch.cuh"

namespace sshash_b200 {

namespace {

constexpr int kTile = 2048;                    // records per counting tile (kBlock threads x 8)
constexpr int kTileItems = kTile / kBlock;
constexpr int kOutTile = 4096;                 // queries per scatter / gather tile (32 KB of u64 ids)
constexpr int kSortThreads = 512;
constexpr int kSortItems = kOutTile / kSortThreads;
constexpr uint32_t kRangeShift = 20;           // 2^20 query indices per range of the (range, bin) histogram
constexpr int kClaimItems = 4;                 // records per lane and claim in phase B
constexpr uint32_t kClaim = 32 * kClaimItems;
constexpr uint32_t kMaxBins = 1024;

// meta1 (u32): bin [0,16) | minimizer pos [16,22) | strand (canonical: minimizer taken from the rc) 22 | tie 23
// record meta (u64): idx [0,32) | (meta1 >> 16) [32,40) | bin [40,56)

struct Control {                 // device-resident control block of one round (u32 unless noted); s = range * n_bins + bin
    uint32_t* counts;            // [s]   records per sub-run
    uint32_t* cursor;            // [s]   scatter cursor: first free slot of the sub-run in the bin-major record arrays
    uint32_t* bin_start;         // [n_bins + 1]
    unsigned long long* claims;  // [0] phase B record cursor, [1] length of the round-2 list so far
};

// exclusive scan of one value per thread over a CTA of NT threads (NT a multiple of 32, <= 1024);
// scratch[NT / 32] holds the grand total afterwards
template <int NT>
__device__ __forceinline__ uint32_t cta_exclusive_scan(uint32_t v, uint32_t* scratch /* NT / 32 + 1 */) {
    const uint32_t lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    uint32_t inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= (uint32_t)o) inc += t; }
    __syncthreads();                                   // the previous use of scratch is over
    if (lane == 31) scratch[wid] = inc;
    __syncthreads();
    if (wid == 0) {
        uint32_t w = lane < NT / 32 ? scratch[lane] : 0;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, w, o); if (lane >= (uint32_t)o) w += t; }
        if (lane < NT / 32) scratch[lane] = w;          // inclusive
        if (lane == 31) scratch[NT / 32] = w;           // grand total (lanes >= NT/32 added zeros)
    }
    __syncthreads();
    const uint32_t before = wid ? scratch[wid - 1] : 0;
    return before + inc - v;
}

// bin | pos | flags of one query k-mer (A1; also run by C1 for the round-2 list)
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

// ---- A2 (one CTA of kMaxBins threads) ----------------------------------------------------------------
__global__ void __launch_bounds__(kMaxBins)
bin_scan_kernel(Control c, uint32_t n_ranges, uint32_t n_bins) {
    __shared__ uint32_t scratch[kMaxBins / 32 + 1];
    uint32_t mine = 0;                                 // this thread's bin: its total over the ranges
    if (threadIdx.x < n_bins)
        for (uint32_t r = 0; r < n_ranges; ++r) mine += c.counts[(uint64_t)r * n_bins + threadIdx.x];
    const uint32_t ex = cta_exclusive_scan<kMaxBins>(mine, scratch);
    if (threadIdx.x < n_bins) {
        c.bin_start[threadIdx.x] = ex;
        uint32_t run = ex;
        for (uint32_t r = 0; r < n_ranges; ++r) {
            const uint64_t s = (uint64_t)r * n_bins + threadIdx.x;
            c.cursor[s] = run;
            run += c.counts[s];
        }
    }
    if (threadIdx.x == 0) c.bin_start[n_bins] = scratch[kMaxBins / 32];
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

// One CTA per tile t.  Round 1 (seg == nullptr): the tile's records are queries [4096 t, 4096 t + 4096), idx = i.
// Round 2: the tile's records are its misses, seg[t] = {first slot, count} in the round-2 list (src_idx = their idx).
// slot_of[first + p] = slot, in the bin-major record arrays, of the p-th record of the tile in bin order: the
// way back for C (consecutive p of one bin are consecutive slots).
template <int W>
__global__ void __launch_bounds__(kSortThreads, 2)
bin_scatter_kernel(const uint64_t* __restrict__ kmers, const uint32_t* __restrict__ src_idx, const uint2* __restrict__ seg,
                   uint64_t n_queries, uint32_t n_tiles, uint32_t n_bins, const uint32_t* __restrict__ meta1,
                   uint32_t* __restrict__ cursor, uint64_t* __restrict__ rec_kmer, uint64_t* __restrict__ rec_meta,
                   uint32_t* __restrict__ slot_of) {
    extern __shared__ __align__(16) uint8_t smem[];
    uint64_t* s_kmer = reinterpret_cast<uint64_t*>(smem);                        // kOutTile * W words
    uint64_t* s_meta = s_kmer + (size_t)kOutTile * W;                            // kOutTile
    uint32_t* hist = reinterpret_cast<uint32_t*>(s_meta + kOutTile);             // n_bins
    uint32_t* toff = hist + n_bins;                                              // n_bins
    uint32_t* gbase = toff + n_bins;                                             // n_bins
    uint32_t* scratch = gbase + n_bins;                                          // kSortThreads / 32 + 1
    for (uint32_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        uint64_t first;
        uint32_t cnt;
        if (seg) { const uint2 sg = seg[tile]; first = sg.x; cnt = sg.y; }
        else { first = (uint64_t)tile * kOutTile; cnt = (uint32_t)(n_queries - first < (uint64_t)kOutTile ? n_queries - first : kOutTile); }
        for (uint32_t b = threadIdx.x; b < n_bins; b += kSortThreads) hist[b] = 0;
        __syncthreads();
        uint32_t meta[kSortItems], rank[kSortItems];
#pragma unroll
        for (int t = 0; t < kSortItems; ++t) {
            const uint32_t o = (uint32_t)t * kSortThreads + threadIdx.x;
            meta[t] = o < cnt ? meta1[first + o] : 0xffffffffu;
            rank[t] = o < cnt ? atomicAdd(&hist[meta[t] & 0xffffu], 1u) : 0u;
        }
        __syncthreads();
        // tile-local offsets of the bins + one global reservation per (tile, bin)
        const uint32_t r = (uint32_t)(((uint64_t)tile * kOutTile) >> kRangeShift);
        for (uint32_t b0 = 0; b0 < n_bins; b0 += kSortThreads) {                 // n_bins <= 1024: at most two sweeps
            const uint32_t b = b0 + threadIdx.x;
            const uint32_t mine = b < n_bins ? hist[b] : 0;
            const uint32_t ex = cta_exclusive_scan<kSortThreads>(mine, scratch);
            const uint32_t carry = b0 ? toff[b0 - 1] + hist[b0 - 1] : 0;
            if (b < n_bins) {
                toff[b] = carry + ex;
                const uint32_t g = mine ? atomicAdd(&cursor[(uint64_t)r * n_bins + b], mine) : 0;
                gbase[b] = g;
            }
            __syncthreads();
        }
#pragma unroll
        for (int t = 0; t < kSortItems; ++t) {
            const uint32_t o = (uint32_t)t * kSortThreads + threadIdx.x;
            if (o >= cnt) continue;
            const uint32_t bin = meta[t] & 0xffffu, p = toff[bin] + rank[t];
            const uint32_t idx = src_idx ? src_idx[first + o] : (uint32_t)(first + o);
            store_plain(s_kmer, p, load_kmer<W>(kmers, first + o));
            s_meta[p] = (uint64_t)idx | ((uint64_t)(meta[t] >> 16) << 32) | ((uint64_t)bin << 40);
        }
        __syncthreads();
        for (uint32_t p = threadIdx.x; p < cnt; p += kSortThreads) {          // consecutive p of a bin -> consecutive slots
            const uint64_t m = s_meta[p];
            const uint32_t bin = (uint32_t)(m >> 40) & 0xffffu;
            const uint64_t dest = (uint64_t)gbase[bin] + (p - toff[bin]);
            store_plain(rec_kmer, dest, load_plain(s_kmer, p, (Kmer<W>*)nullptr));
            rec_meta[dest] = m;
            slot_of[first + p] = (uint32_t)dest;
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
                     uint64_t* __restrict__ res_id, unsigned long long* __restrict__ claim) {
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
            if (j >= n_records) break;
            const Kmer<W> x = load_kmer<W>(rec_kmer, j);
            const uint64_t meta = __ldcs(rec_meta + j);
            const uint32_t pos = (uint32_t)(meta >> 32) & 63u;
            bool found;
            LookupResult res;
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
    }
}

// ---- C --------------------------------------------------------------------------------------------
// MODE 0: u64 ids, 2: membership bytes, 3: u32 ids
template <int MODE> struct OutT;
template <> struct OutT<0> { using type = uint64_t; };
template <> struct OutT<2> { using type = uint8_t; };
template <> struct OutT<3> { using type = uint32_t; };
template <int MODE>
__device__ __forceinline__ typename OutT<MODE>::type out_value(uint64_t id) {
    if (MODE == 2) return id != ~0ull;
    return (typename OutT<MODE>::type)id;                 // u32: "not found" truncates to UINT32_MAX
}

// C1: one CTA per tile, one thread per record of the tile in bin order (slot_of).  SECOND: a round follows --
// misses go to its list (reverse-complemented, with their bin in that round) and read "not found" until
// C2 patches them.
template <int W, int MODE, bool SECOND>
__global__ void __launch_bounds__(kBlock)
gather_tile_kernel(const __grid_constant__ DeviceIndex ix, const uint64_t* __restrict__ rec_kmer, const uint64_t* __restrict__ rec_meta,
                   const uint64_t* __restrict__ res_id, const uint32_t* __restrict__ slot_of, uint64_t n_queries, uint32_t n_tiles,
                   uint32_t n_bins, uint32_t bin_shift, void* __restrict__ out, uint64_t* __restrict__ miss_kmer,
                   uint32_t* __restrict__ miss_idx, uint32_t* __restrict__ miss_meta1, uint2* __restrict__ seg,
                   uint32_t* __restrict__ next_counts, unsigned long long* __restrict__ miss_total) {
    using T = typename OutT<MODE>::type;
    extern __shared__ __align__(16) uint8_t smem[];
    T* image = reinterpret_cast<T*>(smem);                                       // kOutTile results
    uint32_t* hist = reinterpret_cast<uint32_t*>(smem + (size_t)kOutTile * sizeof(T));   // SECOND: n_bins
    __shared__ uint32_t scratch[kBlock / 32 + 1], s_seg;
    for (uint32_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const uint64_t first = (uint64_t)tile * kOutTile;
        const uint32_t cnt = (uint32_t)(n_queries - first < (uint64_t)kOutTile ? n_queries - first : kOutTile);
        if (SECOND) for (uint32_t b = threadIdx.x; b < n_bins; b += kBlock) hist[b] = 0;
        uint32_t my_miss = 0;
#pragma unroll 4
        for (uint32_t p = threadIdx.x; p < cnt; p += kBlock) {
            const uint64_t j = __ldcs(slot_of + first + p);
            const uint64_t id = __ldcs(res_id + j);
            const uint32_t idx = (uint32_t)rec_meta[j];
            image[idx - (uint32_t)first] = out_value<MODE>(id);
            if (SECOND && id == ~0ull) ++my_miss;
        }
        if (SECOND) {
            // slots of the tile's misses in the round-2 list: one reservation per tile, one sub-range per warp,
            // filled in ballot order so that a warp's stores are adjacent
            const uint32_t lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
            uint32_t warp_total = my_miss;
            for (int o = 16; o > 0; o >>= 1) warp_total += __shfl_xor_sync(0xffffffffu, warp_total, o);
            const uint32_t warp_before = cta_exclusive_scan<kBlock>(lane == 0 ? warp_total : 0u, scratch);   // valid in lane 0
            const uint32_t total = scratch[kBlock / 32];
            if (threadIdx.x == 0) {
                s_seg = total ? (uint32_t)atomicAdd(miss_total, (unsigned long long)total) : 0;
                seg[tile] = make_uint2(s_seg, total);
            }
            __syncthreads();
            if (total) {                                          // CTA-uniform
                uint32_t at = s_seg + __shfl_sync(0xffffffffu, warp_before, 0);
                for (uint32_t p0 = wid * 32; p0 < cnt; p0 += kBlock) {     // warp-uniform trip count
                    const uint32_t p = p0 + lane;
                    uint64_t j = 0;
                    bool miss = false;
                    if (p < cnt) { j = slot_of[first + p]; miss = res_id[j] == ~0ull; }
                    const uint32_t mask = __ballot_sync(0xffffffffu, miss);
                    if (miss) {
                        const uint32_t slot = at + __popc(mask & ((1u << lane) - 1));
                        const Kmer<W> xr = kmer_rc(load_plain(rec_kmer, j, (Kmer<W>*)nullptr), ix.k);     // src/dictionary.cpp:72
                        const uint32_t meta = bin_meta_of<W, false>(ix, xr, bin_shift);
                        store_plain(miss_kmer, slot, xr);
                        miss_idx[slot] = (uint32_t)rec_meta[j];
                        miss_meta1[slot] = meta;
                        atomicAdd(&hist[meta & 0xffffu], 1u);
                    }
                    at += __popc(mask);
                }
            }
        }
        __syncthreads();
        // the tile's results leave as one coalesced block
        T* dst = static_cast<T*>(out) + first;
        for (uint32_t i = threadIdx.x; i < cnt; i += kBlock) dst[i] = image[i];
        if (SECOND) {
            const uint32_t r = (uint32_t)(first >> kRangeShift);
            for (uint32_t b = threadIdx.x; b < n_bins; b += kBlock)
                if (hist[b]) atomicAdd(&next_counts[(uint64_t)r * n_bins + b], hist[b]);
        }
        __syncthreads();
    }
}

// C2: one CTA per tile that had misses: read the ids block back, patch the round-2 results in, store it
template <int MODE>
__global__ void __launch_bounds__(kBlock)
patch_tile_kernel(const uint64_t* __restrict__ rec_meta, const uint64_t* __restrict__ res_id, const uint32_t* __restrict__ slot_of,
                  const uint2* __restrict__ seg, uint64_t n_queries, uint32_t n_tiles, void* __restrict__ out) {
    using T = typename OutT<MODE>::type;
    extern __shared__ __align__(16) uint8_t smem[];
    T* image = reinterpret_cast<T*>(smem);
    for (uint32_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const uint2 sg = seg[tile];
        if (sg.y == 0) continue;                              // CTA-uniform: nothing of this tile went to round 2
        const uint64_t first = (uint64_t)tile * kOutTile;
        const uint32_t cnt = (uint32_t)(n_queries - first < (uint64_t)kOutTile ? n_queries - first : kOutTile);
        T* dst = static_cast<T*>(out) + first;
        for (uint32_t i = threadIdx.x; i < cnt; i += kBlock) image[i] = dst[i];
        __syncthreads();
#pragma unroll 4
        for (uint32_t p = threadIdx.x; p < sg.y; p += kBlock) {
            const uint64_t j = __ldcs(slot_of + sg.x + p);
            image[(uint32_t)rec_meta[j] - (uint32_t)first] = out_value<MODE>(__ldcs(res_id + j));
        }
        __syncthreads();
        for (uint32_t i = threadIdx.x; i < cnt; i += kBlock) dst[i] = image[i];
        __syncthreads();
    }
}

uint64_t align_up(uint64_t x, uint64_t a) { return (x + a - 1) / a * a; }

struct Plan {                      // carve-up of the scratch buffer for a batch of n queries
    uint32_t n_ranges, n_bins, n_tiles;
    uint64_t cap;                  // records any array holds
    uint64_t runs;                 // n_ranges * n_bins
    uint64_t off_meta1, off_rec_kmer, off_rec_meta, off_res, off_miss_kmer, off_miss_idx, off_miss_meta1, off_slot, off_seg,
        off_ctl[2], ctl_bytes, total;
};

Plan make_plan(uint32_t kmer_words, uint32_t n_bins, uint64_t n) {
    Plan p{};
    p.n_bins = n_bins;
    p.n_ranges = (uint32_t)((n + (1ull << kRangeShift) - 1) >> kRangeShift);
    p.n_tiles = (uint32_t)((n + kOutTile - 1) / kOutTile);
    p.cap = align_up(n, kOutTile);
    p.runs = (uint64_t)p.n_ranges * n_bins;
    uint64_t o = 0;
    auto take = [&](uint64_t bytes) { const uint64_t at = o; o += align_up(bytes, 256); return at; };
    p.off_meta1 = take(p.cap * 4);
    p.off_rec_kmer = take(p.cap * 8 * kmer_words);
    p.off_rec_meta = take(p.cap * 8);
    p.off_res = take(p.cap * 8);
    p.off_miss_kmer = take(p.cap * 8 * kmer_words);
    p.off_miss_idx = take(p.cap * 4);
    p.off_miss_meta1 = take(p.cap * 4);
    p.off_slot = take(p.cap * 4);
    p.off_seg = take((uint64_t)p.n_tiles * 8);
    // control block: counts, cursor (runs each), bin_start (n_bins + 1), claims (2 x u64)
    p.ctl_bytes = align_up((2 * p.runs + n_bins + 1) * 4, 8) + 16;
    p.off_ctl[0] = take(p.ctl_bytes);
    p.off_ctl[1] = take(p.ctl_bytes);
    p.total = o;
    return p;
}

Control control_at(uint8_t* base, const Plan& p) {
    Control c{};
    uint32_t* u = reinterpret_cast<uint32_t*>(base);
    c.counts = u; c.cursor = u + p.runs; c.bin_start = u + 2 * p.runs;
    c.claims = reinterpret_cast<unsigned long long*>(base + align_up((2 * p.runs + p.n_bins + 1) * 4, 8));
    return c;
}

size_t scatter_smem_bytes(uint32_t kmer_words, uint32_t n_bins) {
    return (size_t)kOutTile * 8 * (kmer_words + 1) + (3 * (size_t)n_bins + kSortThreads / 32 + 1) * 4;
}

}  // namespace

uint64_t binned_max_batch() { return 1ull << 27; }

uint64_t binned_scratch_bytes(const DeviceIndex& ix, const LaunchCtx& ctx, uint64_t n) {
    return make_plan(ix.kmer_words, ctx.bins.n_bins, n).total;
}

cudaError_t launch_lookup_binned(const DeviceIndex& ix, const LaunchCtx& ctx, const uint64_t* queries, uint64_t n, bool check_rc,
                                 uint64_t* ids, uint32_t* ids32, uint8_t* member, void* scratch, cudaStream_t stream) {
    if (n == 0) return cudaSuccess;
    if (n > binned_max_batch() || !ctx.bins.n_bins || ctx.bins.n_bins > kMaxBins) return cudaErrorInvalidValue;
    const Plan p = make_plan(ix.kmer_words, ctx.bins.n_bins, n);
    uint8_t* s = static_cast<uint8_t*>(scratch);
    uint32_t* meta1 = reinterpret_cast<uint32_t*>(s + p.off_meta1);
    uint64_t* rec_kmer = reinterpret_cast<uint64_t*>(s + p.off_rec_kmer);
    uint64_t* rec_meta = reinterpret_cast<uint64_t*>(s + p.off_rec_meta);
    uint64_t* res = reinterpret_cast<uint64_t*>(s + p.off_res);
    uint64_t* miss_kmer = reinterpret_cast<uint64_t*>(s + p.off_miss_kmer);
    uint32_t* miss_idx = reinterpret_cast<uint32_t*>(s + p.off_miss_idx);
    uint32_t* miss_meta1 = reinterpret_cast<uint32_t*>(s + p.off_miss_meta1);
    uint32_t* slot_of = reinterpret_cast<uint32_t*>(s + p.off_slot);
    uint2* seg = reinterpret_cast<uint2*>(s + p.off_seg);
    const bool canon = ix.canonical != 0, two_rounds = !canon && check_rc;
    const int mode = member ? 2 : (ids32 ? 3 : 0);
    void* out = member ? static_cast<void*>(member) : ids32 ? static_cast<void*>(ids32) : static_cast<void*>(ids);
    const uint32_t nb = p.n_bins, nr = p.n_ranges, nt = p.n_tiles, shift = ctx.bins.bin_shift;
    const BinRegion* regions = ctx.bins.prefetch ? ctx.bins.regions : nullptr;
    const int sm = ctx.sm_count;
    const bool w1 = ix.kmer_words == 1;
    cudaError_t e = cudaMemsetAsync(s + p.off_ctl[0], 0, two_rounds ? 2 * align_up(p.ctl_bytes, 256) : p.ctl_bytes, stream);
    if (e != cudaSuccess) return e;
    const size_t hist_bytes = nb * sizeof(uint32_t), sort_smem = scatter_smem_bytes(ix.kmer_words, nb);
    const size_t out_elem = mode == 2 ? 1 : mode == 3 ? 4 : 8;
    // opt-in shared memory above 48 KB is a per-device function attribute: set it on every call (microseconds)
    cudaFuncSetAttribute(bin_scatter_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)scatter_smem_bytes(1, kMaxBins));
    cudaFuncSetAttribute(bin_scatter_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)scatter_smem_bytes(2, kMaxBins));
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
    const int tile_grid = (int)std::min<uint64_t>(nt, (uint64_t)sm * 8);
    // ---- A1 (round 1 only: round 2's bins are computed by C1) ----
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
        e = cfg_launch(bin_scan_kernel, 1, (int)kMaxBins, 0, c, nr, nb);
        if (e != cudaSuccess) return e;
        const uint64_t* src_kmer = r2 ? miss_kmer : queries;
        const uint32_t* src_idx = r2 ? miss_idx : nullptr;
        const uint32_t* src_meta1 = r2 ? miss_meta1 : meta1;
        const uint2* src_seg = r2 ? seg : nullptr;
        const int sgrid = (int)std::min<uint64_t>(nt, (uint64_t)sm * 2);
        e = w1 ? cfg_launch(bin_scatter_kernel<1>, sgrid, kSortThreads, sort_smem, src_kmer, src_idx, src_seg, n, nt, nb, src_meta1, c.cursor, rec_kmer, rec_meta, slot_of)
               : cfg_launch(bin_scatter_kernel<2>, sgrid, kSortThreads, sort_smem, src_kmer, src_idx, src_seg, n, nt, nb, src_meta1, c.cursor, rec_kmer, rec_meta, slot_of);
        if (e != cudaSuccess) return e;
        const int lgrid = sm * 6;
#define SSHASH_LOOKUP_BINNED(W, CANON) \
        cfg_launch(lookup_binned_kernel<W, CANON>, lgrid, kBlock, 0, ix, (const uint64_t*)rec_kmer, (const uint64_t*)rec_meta, nb, (const uint32_t*)c.bin_start, regions, ctx.bins.lookahead, res, c.claims)
        if (canon) e = w1 ? SSHASH_LOOKUP_BINNED(1, true) : SSHASH_LOOKUP_BINNED(2, true);
        else e = w1 ? SSHASH_LOOKUP_BINNED(1, false) : SSHASH_LOOKUP_BINNED(2, false);
#undef SSHASH_LOOKUP_BINNED
        if (e != cudaSuccess) return e;
        if (!r2) {
#define SSHASH_GATHER(W, MODE, SECOND) \
            cfg_launch(gather_tile_kernel<W, MODE, SECOND>, tile_grid, kBlock, kOutTile * out_elem + (SECOND ? hist_bytes : 0), ix, (const uint64_t*)rec_kmer, (const uint64_t*)rec_meta, (const uint64_t*)res, (const uint32_t*)slot_of, n, nt, nb, shift, out, miss_kmer, miss_idx, miss_meta1, seg, c1.counts, c0.claims + 1)
#define SSHASH_GATHER_MODE(W, SECOND) (mode == 2 ? SSHASH_GATHER(W, 2, SECOND) : mode == 3 ? SSHASH_GATHER(W, 3, SECOND) : SSHASH_GATHER(W, 0, SECOND))
            if (second) e = w1 ? SSHASH_GATHER_MODE(1, true) : SSHASH_GATHER_MODE(2, true);
            else e = w1 ? SSHASH_GATHER_MODE(1, false) : SSHASH_GATHER_MODE(2, false);
#undef SSHASH_GATHER_MODE
#undef SSHASH_GATHER
        } else {
#define SSHASH_PATCH(MODE) \
            cfg_launch(patch_tile_kernel<MODE>, tile_grid, kBlock, kOutTile * out_elem, (const uint64_t*)rec_meta, (const uint64_t*)res, (const uint32_t*)slot_of, (const uint2*)seg, n, nt, out)
            e = mode == 2 ? SSHASH_PATCH(2) : mode == 3 ? SSHASH_PATCH(3) : SSHASH_PATCH(0);
#undef SSHASH_PATCH
        }
        if (e != cudaSuccess) return e;
    }
    return cudaSuccess;
}

}  // namespace sshash_b200
