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
//   A3  bin_scatter_kernel  records {k-mer, idx | pos | bin} written bin-major (tile-local ranks in
//                           shared memory, one global atomic per (tile, bin))
//   B   lookup_binned_kernel  warps claim 128 consecutive records; each record is ONE pass of the
//                           reference's lookup with the minimizer given (device_index.cuh); result ids
//                           stored in record order; on a regular index misses are counted per sub-run
//   C   unpermute_kernel    one warp per (range, bin) sub-run in RANGE-major order: hits are stored to
//                           ids[idx] (an 8 MB window of ids per range: the scattered stores merge in
//                           L2); misses of a regular index with check_reverse_complement are appended,
//                           reverse-complemented, to the round-2 list (src/dictionary.cpp:71-76)
//   round 2 = A1..C over the miss list; what still misses is stored as "not found".
// Canonical indexes take one round (src/dictionary.cpp:24-42); the minimizer tie case runs its
// second attempt inline.
#include <algorithm>

#include "kernels.cuh"
#include "launch.cuh"

namespace sshash_b200 {

namespace {

constexpr int kTile = 2048;                    // records per multisplit tile
constexpr int kTileItems = kTile / kBlock;     // records per thread and tile
constexpr uint32_t kRangeShift = 20;           // 2^20 query indices per output range (8 MB of u64 ids)
constexpr uint32_t kPadIdx = 0xffffffffu;      // round-2 list: padding slot (ranges start on tile boundaries)
constexpr int kClaimItems = 4;                 // records per lane and claim in phase B
constexpr uint32_t kClaim = 32 * kClaimItems;

// meta1 (u32): bin [0,16) | minimizer pos [16,22) | strand (canonical: minimizer taken from the rc) 22 | tie 23
// record meta (u64): idx [0,32) | (meta1 >> 16) [32,40) | bin [40,56)

struct Control {                 // device-resident control block of one round (all u32 unless noted)
    uint32_t* counts;            // [range * n_bins + bin]   records per sub-run
    uint32_t* base;              // first slot of the sub-run in the bin-major record arrays
    uint32_t* cursor;            // scatter cursor (starts at base)
    uint32_t* bin_start;         // [n_bins + 1]
    uint32_t* miss_counts;       // [range * n_bins + bin]   misses per sub-run (round 1 of a regular index)
    uint32_t* mbase;             // first slot of the sub-run's misses in the round-2 list
    uint64_t* mstart;            // [n_ranges + 1] first slot of every range in the round-2 list; [n_ranges] = its length
    unsigned long long* claims;  // [0] phase B record cursor, [1] phase C sub-run cursor
};

__device__ __forceinline__ uint32_t range_of_tile(uint64_t tile, const uint64_t* __restrict__ mstart, uint32_t n_ranges) {
    const uint64_t t0 = tile * kTile;
    if (!mstart) return (uint32_t)(t0 >> kRangeShift);
    uint32_t lo = 0, hi = n_ranges;                    // largest r with mstart[r] <= t0 (empty ranges share a start)
    while (hi - lo > 1) {
        const uint32_t mid = (lo + hi) / 2;
        if (mstart[mid] <= t0) lo = mid; else hi = mid;
    }
    return lo;
}

// ---- A1 -------------------------------------------------------------------------------------------
template <int W, bool CANON>
__global__ void __launch_bounds__(kBlock)
bin_count_kernel(const __grid_constant__ DeviceIndex ix, const uint64_t* __restrict__ kmers, const uint32_t* __restrict__ src_idx,
                 uint64_t n_records_arg, const uint64_t* __restrict__ mstart, uint32_t n_ranges, uint32_t n_bins, uint32_t bin_shift,
                 uint32_t* __restrict__ meta1, uint32_t* __restrict__ counts) {
    extern __shared__ uint32_t hist[];
    __shared__ uint32_t s_range;
    const uint64_t n_records = mstart ? mstart[n_ranges] : n_records_arg;
    const uint64_t n_tiles = (n_records + kTile - 1) / kTile;
    for (uint64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        for (uint32_t b = threadIdx.x; b < n_bins; b += kBlock) hist[b] = 0;
        if (threadIdx.x == 0) s_range = range_of_tile(tile, mstart, n_ranges);
        __syncthreads();
#pragma unroll 1
        for (int t = 0; t < kTileItems; ++t) {
            const uint64_t i = tile * kTile + (uint64_t)t * kBlock + threadIdx.x;
            if (i >= n_records) continue;
            uint32_t meta = 0xffffffffu;
            if (!src_idx || src_idx[i] != kPadIdx) {
                const Kmer<W> x = load_kmer<W>(kmers, i);
                Minimizer mi = compute_minimizer(ix, x);
                uint32_t flags = 0;
                if (CANON) {                                   // src/dictionary.cpp:24-42: the smaller minimizer decides
                    const Minimizer mr = compute_minimizer(ix, kmer_rc(x, ix.k));
                    if (mr.value < mi.value) { mi = mr; flags = 1u << 6; }
                    else if (mr.value == mi.value) flags = 1u << 7;      // tie: forward info first, then the rc info
                }
                const uint32_t bin = mphf_partition(ix.mphf, city_hash_u64(ix.mphf, mi.value)) >> bin_shift;
                meta = bin | ((mi.pos | flags) << 16);
                atomicAdd(&hist[bin], 1u);
            }
            meta1[i] = meta;
        }
        __syncthreads();
        const uint32_t r = s_range;
        for (uint32_t b = threadIdx.x; b < n_bins; b += kBlock)
            if (hist[b]) atomicAdd(&counts[(uint64_t)r * n_bins + b], hist[b]);
        __syncthreads();
    }
}

// ---- A2 (one CTA) ---------------------------------------------------------------------------------
__global__ void __launch_bounds__(kBlock)
bin_scan_kernel(Control c, uint32_t n_ranges, uint32_t n_bins) {
    extern __shared__ uint32_t totals[];               // n_bins + 1
    for (uint32_t b = threadIdx.x; b < n_bins; b += kBlock) {
        uint32_t s = 0;
        for (uint32_t r = 0; r < n_ranges; ++r) s += c.counts[(uint64_t)r * n_bins + b];
        totals[b] = s;
    }
    __syncthreads();
    if (threadIdx.x == 0) {                            // <= 1024 bins: a serial scan costs a few microseconds
        uint32_t run = 0;
        for (uint32_t b = 0; b < n_bins; ++b) { const uint32_t t = totals[b]; totals[b] = run; run += t; }
        totals[n_bins] = run;
    }
    __syncthreads();
    for (uint32_t b = threadIdx.x; b <= n_bins; b += kBlock) c.bin_start[b] = totals[b];
    for (uint32_t b = threadIdx.x; b < n_bins; b += kBlock) {
        uint32_t run = totals[b];
        for (uint32_t r = 0; r < n_ranges; ++r) {
            const uint64_t s = (uint64_t)r * n_bins + b;
            c.base[s] = run; c.cursor[s] = run;
            run += c.counts[s];
        }
    }
}

// ---- A3 -------------------------------------------------------------------------------------------
__device__ __forceinline__ void store_record_kmer(uint64_t* out, uint64_t i, Kmer<1> x) { out[i] = x.lo; }
__device__ __forceinline__ void store_record_kmer(uint64_t* out, uint64_t i, Kmer<2> x) {
    reinterpret_cast<ulonglong2*>(out)[i] = make_ulonglong2(x.lo, x.hi);
}

template <int W>
__global__ void __launch_bounds__(kBlock)
bin_scatter_kernel(const uint64_t* __restrict__ kmers, const uint32_t* __restrict__ src_idx, uint64_t n_records_arg,
                   const uint64_t* __restrict__ mstart, uint32_t n_ranges, uint32_t n_bins, const uint32_t* __restrict__ meta1,
                   uint32_t* __restrict__ cursor, uint64_t* __restrict__ rec_kmer, uint64_t* __restrict__ rec_meta) {
    extern __shared__ uint32_t sh[];                   // hist[n_bins] + tile_base[n_bins]
    uint32_t* hist = sh;
    uint32_t* tile_base = sh + n_bins;
    __shared__ uint32_t s_range;
    const uint64_t n_records = mstart ? mstart[n_ranges] : n_records_arg;
    const uint64_t n_tiles = (n_records + kTile - 1) / kTile;
    for (uint64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        for (uint32_t b = threadIdx.x; b < n_bins; b += kBlock) hist[b] = 0;
        if (threadIdx.x == 0) s_range = range_of_tile(tile, mstart, n_ranges);
        __syncthreads();
        uint32_t meta[kTileItems], rank[kTileItems];
#pragma unroll
        for (int t = 0; t < kTileItems; ++t) {
            const uint64_t i = tile * kTile + (uint64_t)t * kBlock + threadIdx.x;
            meta[t] = i < n_records ? meta1[i] : 0xffffffffu;
            rank[t] = meta[t] != 0xffffffffu ? atomicAdd(&hist[meta[t] & 0xffffu], 1u) : 0u;
        }
        __syncthreads();
        const uint32_t r = s_range;
        for (uint32_t b = threadIdx.x; b < n_bins; b += kBlock)
            if (hist[b]) tile_base[b] = atomicAdd(&cursor[(uint64_t)r * n_bins + b], hist[b]);
        __syncthreads();
#pragma unroll
        for (int t = 0; t < kTileItems; ++t) {
            if (meta[t] == 0xffffffffu) continue;
            const uint64_t i = tile * kTile + (uint64_t)t * kBlock + threadIdx.x;
            const uint32_t bin = meta[t] & 0xffffu;
            const uint64_t dest = (uint64_t)tile_base[bin] + rank[t];
            const uint32_t idx = src_idx ? src_idx[i] : (uint32_t)i;
            const Kmer<W> x = load_kmer<W>(kmers, i);
            // plain stores: the ~n/bins records a tile sends to one bin are adjacent, L2 merges them into full sectors
            store_record_kmer(rec_kmer, dest, x);
            rec_meta[dest] = (uint64_t)idx | ((uint64_t)(meta[t] >> 16) << 32) | ((uint64_t)bin << 40);
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
                     uint64_t n_records_arg, const uint64_t* __restrict__ mstart, uint32_t n_ranges, uint32_t n_bins,
                     const uint32_t* __restrict__ bin_start, const BinRegion* __restrict__ regions, uint32_t lookahead,
                     uint64_t* __restrict__ res_id, uint32_t* __restrict__ miss_counts, unsigned long long* __restrict__ claim) {
    const uint32_t lane = threadIdx.x & 31;
    const uint64_t n_records = mstart ? mstart[n_ranges] : n_records_arg;   // round 2 holds padding slots, but none were scattered:
    const uint64_t n_scattered = bin_start[n_bins];                         // the record arrays hold bin_start[n_bins] records
    (void)n_records;
    for (;;) {
        unsigned long long first = 0;
        if (lane == 0) first = atomicAdd(claim, (unsigned long long)kClaim);
        first = __shfl_sync(0xffffffffu, first, 0);
        if (first >= n_scattered) break;
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
            const bool active = j < n_scattered;
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
            run += (t + kTile - 1) / kTile * kTile;          // every range starts on a tile boundary
        }
        c.mstart[n_ranges] = run;
    }
    __syncthreads();
    for (uint32_t r = threadIdx.x; r < n_ranges; r += kBlock) {
        uint32_t run = (uint32_t)c.mstart[r];
        for (uint32_t b = 0; b < n_bins; ++b) {
            const uint64_t s = (uint64_t)r * n_bins + b;
            c.mbase[s] = run;
            run += c.miss_counts[s];
        }
    }
}

// ---- C --------------------------------------------------------------------------------------------
// MODE 0: u64 ids, 2: membership bytes, 3: u32 ids
template <int W, int MODE>
__global__ void __launch_bounds__(kBlock)
unpermute_kernel(uint32_t k, const uint64_t* __restrict__ rec_kmer, const uint64_t* __restrict__ rec_meta, const uint64_t* __restrict__ res_id,
                 Control c, uint32_t n_ranges, uint32_t n_bins, void* __restrict__ out, uint64_t* __restrict__ miss_kmer,
                 uint32_t* __restrict__ miss_idx) {
    const uint32_t lane = threadIdx.x & 31;
    const uint64_t n_runs = (uint64_t)n_ranges * n_bins;
    for (;;) {
        unsigned long long s = 0;
        if (lane == 0) s = atomicAdd(c.claims + 1, 1ull);
        s = __shfl_sync(0xffffffffu, s, 0);
        if (s >= n_runs) break;
        const uint32_t beg = c.base[s], cnt = c.counts[s];
        uint32_t mb = miss_kmer ? c.mbase[s] : 0;
        for (uint32_t off = 0; off < cnt; off += 32) {
            const bool valid = off + lane < cnt;
            const uint64_t j = (uint64_t)beg + off + lane;
            uint64_t id = ~0ull;
            uint32_t idx = 0;
            if (valid) { id = __ldcs(res_id + j); idx = (uint32_t)__ldcs(rec_meta + j); }
            const bool hit = id != ~0ull;
            if (valid && (hit || !miss_kmer)) {
                if (MODE == 2) static_cast<uint8_t*>(out)[idx] = hit;
                else if (MODE == 3) static_cast<uint32_t*>(out)[idx] = (uint32_t)id;   // not found: UINT32_MAX
                else static_cast<uint64_t*>(out)[idx] = id;
            }
            if (miss_kmer) {
                const bool miss = valid && !hit;
                const uint32_t mask = __ballot_sync(0xffffffffu, miss);
                if (miss) {
                    const uint32_t slot = mb + __popc(mask & ((1u << lane) - 1));
                    store_kmer(miss_kmer, slot, kmer_rc(load_kmer<W>(rec_kmer, j), k));   // src/dictionary.cpp:72
                    miss_idx[slot] = idx;
                }
                mb += __popc(mask);
            }
        }
    }
}

uint64_t align_up(uint64_t x, uint64_t a) { return (x + a - 1) / a * a; }

struct Plan {                      // carve-up of the scratch buffer for a batch of n queries
    uint32_t n_ranges, n_bins;
    uint64_t cap;                  // records any array holds: n + one tile of padding per range
    uint64_t runs;                 // n_ranges * n_bins
    uint64_t off_meta1, off_rec_kmer, off_rec_meta, off_res, off_miss_kmer, off_miss_idx, off_ctl[2], ctl_bytes, total;
};

Plan make_plan(uint32_t kmer_words, uint32_t n_bins, uint64_t n) {
    Plan p{};
    p.n_bins = n_bins;
    p.n_ranges = (uint32_t)((n + (1ull << kRangeShift) - 1) >> kRangeShift);
    p.cap = align_up(n, kTile) + (uint64_t)p.n_ranges * kTile;
    p.runs = (uint64_t)p.n_ranges * n_bins;
    uint64_t o = 0;
    auto take = [&](uint64_t bytes) { const uint64_t at = o; o += align_up(bytes, 256); return at; };
    p.off_meta1 = take(p.cap * 4);
    p.off_rec_kmer = take(p.cap * 8 * kmer_words);
    p.off_rec_meta = take(p.cap * 8);
    p.off_res = take(p.cap * 8);
    p.off_miss_kmer = take(p.cap * 8 * kmer_words);
    p.off_miss_idx = take(p.cap * 4);
    // control block: counts, base, cursor, miss_counts, mbase (runs each), bin_start (n_bins + 1), mstart (n_ranges + 1, u64), claims (2 x u64)
    p.ctl_bytes = align_up(5 * p.runs * 4 + (n_bins + 1) * 4, 8) + (p.n_ranges + 1) * 8 + 16;
    p.off_ctl[0] = take(p.ctl_bytes);
    p.off_ctl[1] = take(p.ctl_bytes);
    p.total = o;
    return p;
}

Control control_at(uint8_t* base, const Plan& p) {
    Control c{};
    uint32_t* u = reinterpret_cast<uint32_t*>(base);
    c.counts = u; c.base = u + p.runs; c.cursor = u + 2 * p.runs; c.miss_counts = u + 3 * p.runs; c.mbase = u + 4 * p.runs;
    c.bin_start = u + 5 * p.runs;
    uint8_t* q = base + align_up(5 * p.runs * 4 + (p.n_bins + 1) * 4, 8);
    c.mstart = reinterpret_cast<uint64_t*>(q);
    c.claims = reinterpret_cast<unsigned long long*>(q + (p.n_ranges + 1) * 8);
    return c;
}

}  // namespace

uint64_t binned_max_batch() { return 1ull << 27; }

uint64_t binned_scratch_bytes(const DeviceIndex& ix, const LaunchCtx& ctx, uint64_t n) {
    return make_plan(ix.kmer_words, ctx.bins.n_bins, n).total;
}

cudaError_t launch_lookup_binned(const DeviceIndex& ix, const LaunchCtx& ctx, const uint64_t* queries, uint64_t n, bool check_rc,
                                 uint64_t* ids, uint32_t* ids32, uint8_t* member, void* scratch, cudaStream_t stream) {
    if (n == 0) return cudaSuccess;
    if (n > binned_max_batch() || !ctx.bins.n_bins || ctx.bins.n_bins > 1024) return cudaErrorInvalidValue;
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
    cudaError_t e = cudaMemsetAsync(s + p.off_ctl[0], 0, two_rounds ? 2 * align_up(p.ctl_bytes, 256) : p.ctl_bytes, stream);
    if (e != cudaSuccess) return e;
    if (two_rounds) {
        e = cudaMemsetAsync(miss_idx, 0xff, p.cap * 4, stream);
        if (e != cudaSuccess) return e;
    }
    const size_t hist_bytes = nb * sizeof(uint32_t);
    auto cfg_launch = [&](auto kernel, int grid, size_t smem, auto... args) -> cudaError_t {
        g_launches.fetch_add(1);
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3((unsigned)grid); cfg.blockDim = dim3(kBlock); cfg.dynamicSmemBytes = smem; cfg.stream = stream;
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
    for (int round = 0; round < (two_rounds ? 2 : 1); ++round) {
        const Control c = control_at(s + p.off_ctl[round], p);
        const Control c0 = control_at(s + p.off_ctl[0], p);
        const bool r2 = round == 1;
        const uint64_t* src_kmer = r2 ? miss_kmer : queries;
        const uint32_t* src_idx = r2 ? miss_idx : nullptr;
        const uint64_t* mstart = r2 ? c0.mstart : nullptr;
        const uint64_t bound = r2 ? p.cap : n;              // round 2's exact length lives on the device (mstart[n_ranges])
        const int tiles_grid = (int)std::min<uint64_t>((bound + kTile - 1) / kTile, (uint64_t)sm * 8);
#define SSHASH_BY_W(CALL_1, CALL_2) (ix.kmer_words == 1 ? (CALL_1) : (CALL_2))
        if (canon)
            e = SSHASH_BY_W((cfg_launch(bin_count_kernel<1, true>, tiles_grid, hist_bytes, ix, src_kmer, src_idx, n, mstart, nr, nb, shift, meta1, c.counts)),
                            (cfg_launch(bin_count_kernel<2, true>, tiles_grid, hist_bytes, ix, src_kmer, src_idx, n, mstart, nr, nb, shift, meta1, c.counts)));
        else
            e = SSHASH_BY_W((cfg_launch(bin_count_kernel<1, false>, tiles_grid, hist_bytes, ix, src_kmer, src_idx, n, mstart, nr, nb, shift, meta1, c.counts)),
                            (cfg_launch(bin_count_kernel<2, false>, tiles_grid, hist_bytes, ix, src_kmer, src_idx, n, mstart, nr, nb, shift, meta1, c.counts)));
        if (e != cudaSuccess) return e;
        e = cfg_launch(bin_scan_kernel, 1, (nb + 1) * sizeof(uint32_t), c, nr, nb);
        if (e != cudaSuccess) return e;
        e = SSHASH_BY_W((cfg_launch(bin_scatter_kernel<1>, tiles_grid, 2 * hist_bytes, src_kmer, src_idx, n, mstart, nr, nb, (const uint32_t*)meta1, c.cursor, rec_kmer, rec_meta)),
                        (cfg_launch(bin_scatter_kernel<2>, tiles_grid, 2 * hist_bytes, src_kmer, src_idx, n, mstart, nr, nb, (const uint32_t*)meta1, c.cursor, rec_kmer, rec_meta)));
        if (e != cudaSuccess) return e;
        uint32_t* miss_counts = (two_rounds && !r2) ? c.miss_counts : nullptr;
        const int lgrid = sm * 6;
        if (canon)
            e = SSHASH_BY_W((cfg_launch(lookup_binned_kernel<1, true>, lgrid, 0, ix, (const uint64_t*)rec_kmer, (const uint64_t*)rec_meta, n, mstart, nr, nb, (const uint32_t*)c.bin_start, regions, ctx.bins.lookahead, res, miss_counts, c.claims)),
                            (cfg_launch(lookup_binned_kernel<2, true>, lgrid, 0, ix, (const uint64_t*)rec_kmer, (const uint64_t*)rec_meta, n, mstart, nr, nb, (const uint32_t*)c.bin_start, regions, ctx.bins.lookahead, res, miss_counts, c.claims)));
        else
            e = SSHASH_BY_W((cfg_launch(lookup_binned_kernel<1, false>, lgrid, 0, ix, (const uint64_t*)rec_kmer, (const uint64_t*)rec_meta, n, mstart, nr, nb, (const uint32_t*)c.bin_start, regions, ctx.bins.lookahead, res, miss_counts, c.claims)),
                            (cfg_launch(lookup_binned_kernel<2, false>, lgrid, 0, ix, (const uint64_t*)rec_kmer, (const uint64_t*)rec_meta, n, mstart, nr, nb, (const uint32_t*)c.bin_start, regions, ctx.bins.lookahead, res, miss_counts, c.claims)));
        if (e != cudaSuccess) return e;
        uint64_t* mk = nullptr;
        uint32_t* mi = nullptr;
        if (miss_counts) {
            e = cfg_launch(miss_scan_kernel, 1, 0, c, nr, nb);
            if (e != cudaSuccess) return e;
            mk = miss_kmer; mi = miss_idx;
        }
        const int ugrid = (int)std::min<uint64_t>((p.runs + kBlock / 32 - 1) / (kBlock / 32), (uint64_t)sm * 8);
#define SSHASH_UNPERMUTE(MODE)                                                                                                            \
        SSHASH_BY_W((cfg_launch(unpermute_kernel<1, MODE>, ugrid, 0, ix.k, (const uint64_t*)rec_kmer, (const uint64_t*)rec_meta, (const uint64_t*)res, c, nr, nb, out, mk, mi)), \
                    (cfg_launch(unpermute_kernel<2, MODE>, ugrid, 0, ix.k, (const uint64_t*)rec_kmer, (const uint64_t*)rec_meta, (const uint64_t*)res, c, nr, nb, out, mk, mi)))
        e = mode == 2 ? SSHASH_UNPERMUTE(2) : mode == 3 ? SSHASH_UNPERMUTE(3) : SSHASH_UNPERMUTE(0);
#undef SSHASH_UNPERMUTE
#undef SSHASH_BY_W
        if (e != cudaSuccess) return e;
    }
    return cudaSuccess;
}

}  // namespace sshash_b200
