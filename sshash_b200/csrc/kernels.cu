// kernels.cu -- sm_100a kernels of the lookup path and their launchers.
//
// All kernels are integer / indexing work bound by random 32-byte-sector accesses to the index
// arrays (L2 when the index fits its 126 MB, HBM otherwise); there is nothing GEMM-shaped here,
// so no tensor cores.  Parallelisation: one THREAD per query k-mer -- a lookup is a chain of 4-6
// dependent loads, and the only way to cover ~600-800 ns of HBM latency per link is to keep tens
// of thousands of independent chains in flight (148 SMs x up to 2048 threads).
#include "kernels.cuh"
#include "launch.cuh"

#include <cuda_runtime.h>

#include <algorithm>
#include <cstdlib>

namespace sshash_b200 {

namespace {

// 6 resident CTAs/SM (<= 40 registers, 75 % occupancy): measured best on B200 for both the L2-resident
// and the HBM-resident case (sweep over 3/4/5/6/8 in profiles/r1_notes.md)
#ifndef SSHASH_LOOKUP_MINB
#define SSHASH_LOOKUP_MINB 6
#endif
constexpr int kLookupMinBlocks = SSHASH_LOOKUP_MINB;
// the canonical flow on 128-bit k-mers keeps x, its reverse complement and two candidate k-mers
// live and spills ~100 bytes at 40 registers; measured on a 5e8-k-mer k=63 canonical index it is
// still fastest with 6 resident CTAs (14.2 G lookups/s vs 13.4 with 5, 13.3 with 4)
#ifndef SSHASH_WIDE_CANON_MINB
#define SSHASH_WIDE_CANON_MINB 6
#endif
constexpr int kLookupMinBlocksWideCanon = SSHASH_WIDE_CANON_MINB;

__device__ __forceinline__ void store_full(sshash_lookup_result* full, uint64_t i, const LookupResult& r) {
    // 64-byte record written as four 16-byte streaming stores
    ulonglong2* p = reinterpret_cast<ulonglong2*>(full + i);
    __stcs(p + 0, make_ulonglong2(r.kmer_id, r.kmer_id_in_string));
    __stcs(p + 1, make_ulonglong2(r.kmer_offset, (uint64_t)r.kmer_orientation));
    __stcs(p + 2, make_ulonglong2(r.string_id, r.string_begin));
    __stcs(p + 3, make_ulonglong2(r.string_end, r.minimizer_found));
}

// util::string_to_uint_kmer (include/util.hpp:207-213) with char_to_uint = (c >> 1) & 3
// (include/kmer.hpp:194); no validation, exactly like the reference.
template <int W>
__device__ __forceinline__ Kmer<W> pack_ascii(const char* __restrict__ s, uint32_t k);
template <>
__device__ __forceinline__ Kmer<1> pack_ascii<1>(const char* __restrict__ s, uint32_t k) {
    uint64_t x = 0;
    for (uint32_t i = 0; i < k; ++i) x |= (uint64_t)(((uint8_t)s[i] >> 1) & 3) << (2 * i);
    return {x};
}
template <>
__device__ __forceinline__ Kmer<2> pack_ascii<2>(const char* __restrict__ s, uint32_t k) {
    uint64_t lo = 0, hi = 0;
    for (uint32_t i = 0; i < k; ++i) {
        uint64_t c = ((uint8_t)s[i] >> 1) & 3;
        if (i < 32) lo |= c << (2 * i); else hi |= c << (2 * (i - 32));
    }
    return {lo, hi};
}

// ------------------------------------------------------------------------------------------------
// batched dictionary::lookup.  MODE 0: ids, 1: ids + full records, 2: membership bytes, 3: 32-bit ids (`ids` reinterpreted)
//
// One thread per query.  On a regular (non-canonical) index a k-mer stored in the other
// orientation misses the forward pass and needs a second pass on its reverse complement
// (src/dictionary.cpp:71-76) -- with the reference's 50 % RC query mix that is every other lane,
// and running the second pass under divergence would idle half of each warp.  Instead each warp
// parks its missed queries in a small shared-memory queue and runs the reverse-complement pass
// only on FULL groups of 32 (plus one drain at the end), so both passes execute with all lanes on.
// ------------------------------------------------------------------------------------------------
template <int MODE>
__device__ __forceinline__ void emit(uint64_t i, const LookupResult& r, uint64_t* __restrict__ ids,
                                     sshash_lookup_result* __restrict__ full, uint8_t* __restrict__ member) {
    if (MODE == 2) member[i] = r.kmer_id != ~0ull;
    else if (MODE == 3) __stcs(reinterpret_cast<uint32_t*>(ids) + i, (uint32_t)r.kmer_id);   // u32 ids: UINT32_MAX = not found
    else {
        if (ids) __stcs(ids + i, r.kmer_id);
        if (MODE == 1) store_full(full, i, r);
    }
}

template <int W> struct RcQueue;
template <> struct RcQueue<1> {
    uint64_t kmer[64]; uint64_t idx[64];
    __device__ void put(uint32_t s, Kmer<1> x, uint64_t i) { kmer[s] = x.lo; idx[s] = i; }
    __device__ Kmer<1> get(uint32_t s) const { return {kmer[s]}; }
};
template <> struct RcQueue<2> {
    uint64_t lo[64]; uint64_t hi[64]; uint64_t idx[64];
    __device__ void put(uint32_t s, Kmer<2> x, uint64_t i) { lo[s] = x.lo; hi[s] = x.hi; idx[s] = i; }
    __device__ Kmer<2> get(uint32_t s) const { return {lo[s], hi[s]}; }
};

// The loop has ONE lookup call site: each trip a warp takes either 32 fresh queries or, as soon as
// 32 are parked (or the input is exhausted), 32 parked reverse complements.  CANON selects the
// canonical (src/dictionary.cpp:24-56) or the regular (:7-22, :64-78) flow at compile time so that
// each instantiation carries only its own path.
template <int W, int MODE, bool ASCII, bool CANON, int MINB, bool WIDE = false>
__global__ void __launch_bounds__(kBlock, MINB)
lookup_kernel(const __grid_constant__ DeviceIndex ix, const void* __restrict__ queries, uint64_t n, int check_rc,
              uint64_t* __restrict__ ids, sshash_lookup_result* __restrict__ full, uint8_t* __restrict__ member) {
    constexpr bool FULL = MODE == 1;
    __shared__ RcQueue<W> queues[kBlock / 32];
    RcQueue<W>& q = queues[threadIdx.x >> 5];
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t k = ix.k;
    const bool two_pass = !CANON && check_rc != 0;
    uint32_t queued = 0;                                  // warp-uniform
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    uint64_t tile = (uint64_t)blockIdx.x * blockDim.x + (threadIdx.x & ~31u);
    for (;;) {                                            // warp-uniform control flow
        const bool fresh = tile < n && queued < 32;
        if (!fresh && queued == 0) break;
        Kmer<W> x{};
        uint64_t i;
        bool active;
        if (fresh) {
            i = tile + lane;
            tile += stride;
            active = i < n;
            if (active) {
                if (ASCII) x = pack_ascii<W>(static_cast<const char*>(queries) + i * k, k);
                else x = load_kmer<W>(static_cast<const uint64_t*>(queries), i);
            }
        } else {                                          // the parked reverse complements, newest first
            const uint32_t cnt = queued < 32 ? queued : 32;
            queued -= cnt;
            active = lane < cnt;
            x = q.get(queued + lane);
            i = q.idx[queued + lane];
            __syncwarp();
        }
        bool found = false;
        LookupResult r;
        if (active) {
            if constexpr (WIDE) {                              // wide entries: ids-only, 64-bit k-mers (device_index.cuh)
                if (CANON) found = lookup_canonical_wide(ix, x, r);
                else found = lookup_regular_wide(ix, x, r);
            } else {
                if (CANON) found = lookup_canonical<W, FULL>(ix, x, r);
                else found = lookup_regular<W, FULL>(ix, x, r);
            }
        }
        const bool park = active && fresh && two_pass && !found;
        if (active && !park) {
            if (!CANON && !fresh) r.kmer_orientation = -1;    // dictionary.cpp:74-75 (also for a miss)
            emit<MODE>(i, r, ids, full, member);
        }
        if (fresh && two_pass) {
            const uint32_t mask = __ballot_sync(0xffffffffu, park);
            if (park) q.put(queued + __popc(mask & ((1u << lane) - 1)), kmer_rc(x, k), i);
            queued += __popc(mask);
            __syncwarp();
        }
    }
}

// ------------------------------------------------------------------------------------------------
// diagnostics: MPHF partition of every query's forward minimizer (partitioned_phf.hpp:145-149), the key
// the partition-major path bins by
// ------------------------------------------------------------------------------------------------
template <int W>
__global__ void __launch_bounds__(kBlock)
minimizer_partition_kernel(const __grid_constant__ DeviceIndex ix, const uint64_t* __restrict__ kmers, uint64_t n,
                           uint32_t* __restrict__ out) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const Minimizer mi = compute_minimizer(ix, load_kmer<W>(kmers, i));
        out[i] = mphf_partition(ix.mphf, city_hash_u64(ix.mphf, mi.value));
    }
}

// ------------------------------------------------------------------------------------------------
// open-time validation of what the lookup kernels trust without checking (a corrupted or crafted file
// must be rejected at open, not fault later): every control codeword's bucket reference and every
// offset stored in the bucket arrays.  Runs on the verbatim codewords (cw_fp_bits == 0).
// flag bits: 1 singleton offset, 2 mid-load range, 4 heavy part / begin, 8 mid-load offset, 16 heavy offset
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kBlock)
validate_index_kernel(const __grid_constant__ DeviceIndex ix, uint32_t* __restrict__ flag) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x, t0 = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t n_bases = ix.strings_bits / 2;
    uint32_t bad = 0;
    for (uint64_t id = t0; id < ix.codewords.size; id += stride) {
        uint64_t code = compact_get<false>(ix.codewords, id);
        if ((code & 1) == 0) { if ((code >> 1) >= n_bases) bad |= 1; }
        else if ((code & 3) == 1) {
            code >>= 2;
            const uint64_t size = (code & 63) + 2;
            const uint64_t first = ix.begin_buckets_of_size[size] + (code >> 6) * size;
            if (first + size > ix.mid_load.size || first + size < first) bad |= 2;
        } else {
            code >>= 2;
            if ((code & 7) >= ix.n_skew || (code >> 3) >= ix.heavy.size) bad |= 4;
        }
    }
    for (uint64_t i = t0; i < ix.mid_load.size; i += stride) if (compact_get<false>(ix.mid_load, i) >= n_bases) bad |= 8;
    for (uint64_t i = t0; i < ix.heavy.size; i += stride) if (compact_get<false>(ix.heavy, i) >= n_bases) bad |= 16;
    if (bad) atomicOr(flag, bad);
}

// ------------------------------------------------------------------------------------------------
// open-time re-encoding of the control codewords with minimizer fingerprints: one thread per MPHF
// slot reads the slot's codeword from the verbatim vector (still in ix.codewords, cw_fp_bits == 0),
// follows it to the first offset of its bucket (sparse_and_skew_index.hpp:112-137), fingerprints the
// m-mer found there in `strings` and ORs  codeword | fp << w  into the zeroed output vector.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kBlock)
build_fingerprints_kernel(const __grid_constant__ DeviceIndex ix, uint32_t fp_bits, unsigned long long* __restrict__ out,
                          uint32_t* __restrict__ filter, uint32_t filter_shift) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    const uint32_t w = ix.codewords.width, wo = w + fp_bits;
    DeviceIndex fx = ix;                      // only canonical / m / cw_fp_bits are read by the fingerprint
    fx.cw_fp_bits = fp_bits;
    for (uint64_t id = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; id < ix.codewords.size; id += stride) {
        const uint64_t entry = compact_get<false>(ix.codewords, id);
        uint64_t code = entry, off;
        if ((code & 1) == 0) off = code >> 1;
        else if ((code & 3) == 1) {
            code >>= 2;
            const uint32_t size = (uint32_t)(code & 63) + 2;
            off = compact_get<false>(ix.mid_load, ix.begin_buckets_of_size[size] + (code >> 6) * size);
        } else off = compact_get<false>(ix.heavy, code >> 5);           // (code >> 2) >> 3 = begin of the heavy bucket
        const uint32_t key32 = minimizer_key32(fx, read_mmer(ix, off, ix.m));
        if (filter) {                                    // the minimizer filter (device_index.cuh), same pass
            uint32_t word, mask;
            filter_slot(key32, filter_shift, word, mask);
            atomicOr(filter + word, mask);
        }
        const uint64_t v = entry | ((uint64_t)fingerprint_of_key(fx, key32) << w);
        const uint64_t pos = id * wo, word = pos >> 6;
        const uint32_t s = (uint32_t)pos & 63u;
        atomicOr(out + word, (unsigned long long)(v << s));
        if (s + wo > 64) atomicOr(out + word + 1, (unsigned long long)(v >> (64 - s)));
    }
}

// ------------------------------------------------------------------------------------------------
// open-time construction of the WIDE ENTRIES (device_index.cuh): one thread per MPHF slot copies the
// slot's codeword and, for a SINGLETON bucket, the 2(2k - m) bits of `strings` around its offset
// ([offset - (k - m), offset + k); positions before the start of the text read as zero).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kBlock)
build_wide_kernel(const __grid_constant__ DeviceIndex ix, ulonglong2* __restrict__ out) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    const uint32_t w = ix.cw_code_bits, span = ix.k - ix.m, text_bits = 2 * (2 * ix.k - ix.m);
    for (uint64_t id = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; id < ix.codewords.size; id += stride) {
        const uint64_t code = compact_get<false>(ix.codewords, id) & low_mask(w);
        uint64_t lo = 0, hi = 0;
        if ((code & 1) == 0) {
            const uint64_t off = code >> 1;
            if (off >= span) {
                lo = read_word64(ix.strings, 2 * (off - span));
                hi = read_word64(ix.strings, 2 * (off - span) + 64);
            } else {
                const uint32_t sh = 2 * (uint32_t)(span - off);          // in [2, 60]
                const uint64_t a = read_word64(ix.strings, 0), b = read_word64(ix.strings, 64);
                lo = a << sh;
                hi = (b << sh) | (a >> (64 - sh));
            }
            if (text_bits <= 64) { lo &= low_mask(text_bits); hi = 0; }
            else hi &= low_mask(text_bits - 64);
        }
        out[id] = make_ulonglong2(code | (lo << w), (lo >> (64 - w)) | (hi << w));   // w in [1, 63], w + text_bits <= 128
    }
}

// ------------------------------------------------------------------------------------------------
// navigational queries (src/dictionary.cpp:112-201): a neighbourhood is 8 lookups -- the k-1
// suffix extended by A,C,T,G (forward) and the k-1 prefix preceded by A,C,T,G (backward); alphabet
// order "ACTG" = codes 0..3 (include/kmer.hpp:115-119,194).  These kernels only EXPAND the 8 query
// k-mers per input; the lookups themselves run in lookup_kernel.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ Kmer<1> with_last(Kmer<1> suffix, uint32_t c, uint32_t k) { return {suffix.lo | ((uint64_t)c << (2 * (k - 1)))}; }
__device__ __forceinline__ Kmer<2> with_last(Kmer<2> suffix, uint32_t c, uint32_t k) {
    const uint32_t s = 2 * (k - 1);
    if (s < 64) suffix.lo |= (uint64_t)c << s; else suffix.hi |= (uint64_t)c << (s - 64);
    return suffix;
}
__device__ __forceinline__ Kmer<1> drop_first(Kmer<1> x) { return {x.lo >> 2}; }                       // get_suffix :139-143
__device__ __forceinline__ Kmer<2> drop_first(Kmer<2> x) { return {(x.lo >> 2) | (x.hi << 62), x.hi >> 2}; }
__device__ __forceinline__ Kmer<1> pad_first(Kmer<1> x, uint32_t k) { return {(x.lo << 2) & low_mask(2 * k)}; }   // get_prefix :160-165
__device__ __forceinline__ Kmer<2> pad_first(Kmer<2> x, uint32_t k) {
    Kmer<2> r{x.lo << 2, (x.hi << 2) | (x.lo >> 62)};
    if (2 * k <= 64) { r.lo &= low_mask(2 * k); r.hi = 0; } else r.hi &= low_mask(2 * k - 64);
    return r;
}
__device__ __forceinline__ Kmer<1> or_first(Kmer<1> x, uint32_t c) { return {x.lo | c}; }
__device__ __forceinline__ Kmer<2> or_first(Kmer<2> x, uint32_t c) { return {x.lo | c, x.hi}; }

template <int W, bool STRINGS>
__global__ void __launch_bounds__(kBlock)
expand_neighbours_kernel(const __grid_constant__ DeviceIndex ix, const uint64_t* __restrict__ in, uint64_t n,
                         uint64_t* __restrict__ out) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    const uint32_t k = ix.k;
    for (uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; t < 8 * n; t += stride) {
        const uint64_t i = t >> 3;
        const uint32_t j = (uint32_t)t & 7;
        Kmer<W> suffix, prefix;
        if (STRINGS) {                                   // string_neighbours :189-201, spss.hpp:19-27
            const uint64_t sid = in[i];
            const uint64_t begin = ld64<true>(ix.ends + sid), end = ld64<true>(ix.ends + sid + 1);
            suffix = read_kmer(ix, end - k + 1, k - 1, (Kmer<W>*)nullptr);
            prefix = pad_first(read_kmer(ix, begin, k - 1, (Kmer<W>*)nullptr), k);
        } else {
            const Kmer<W> x = load_kmer<W>(in, i);
            suffix = drop_first(x);
            prefix = pad_first(x, k);
        }
        store_kmer(out, t, j < 4 ? with_last(suffix, j, k) : or_first(prefix, j - 4));
    }
}

// slots the reference leaves default-constructed (e.g. backward[] of kmer_forward_neighbours)
__global__ void __launch_bounds__(kBlock)
reset_neighbour_slots_kernel(uint64_t n, int which, uint64_t* __restrict__ ids, sshash_lookup_result* __restrict__ full) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; t < 8 * n; t += stride) {
        const bool fwd_slot = (t & 7) < 4;
        if ((fwd_slot && (which & 1)) || (!fwd_slot && (which & 2))) continue;
        if (ids) ids[t] = ~0ull;
        if (full) { LookupResult r; result_clear(r, true); store_full(full, t, r); }
    }
}

// ------------------------------------------------------------------------------------------------
// batched dictionary::access: offsets::id_to_offset (include/offsets.hpp:41-65) restated as a
// binary search over the decoded end-points for the last string whose first k-mer id
// (= begin - string_id * (k-1)) is <= id, then spss::access (spss.hpp:114-118).
// ------------------------------------------------------------------------------------------------
template <int W>
__global__ void __launch_bounds__(kBlock)
access_kernel(const __grid_constant__ DeviceIndex ix, const uint64_t* __restrict__ ids, uint64_t n,
              uint64_t* __restrict__ kmers_out) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    const uint64_t km1 = ix.k - 1;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        uint64_t id = __ldcs(ids + i);
        uint64_t lo = 0, hi = ix.n_ends - 1;
        while (hi - lo > 1) {
            uint64_t mid = lo + (hi - lo) / 2;
            if (ld64<true>(ix.ends + mid) - mid * km1 <= id) lo = mid; else hi = mid;
        }
        store_kmer(kmers_out, i, read_kmer(ix, id + lo * km1, ix.k, (Kmer<W>*)nullptr));
    }
}

// ------------------------------------------------------------------------------------------------
// batched dictionary::weight (src/dictionary.cpp:96-100 -> weights::weight, include/weights.hpp:148-153):
// interval = prev_leq(kmer_id) over the interval starts (elias_fano.hpp:254-258; the starts are
// strictly increasing, so it is the last start <= id), then two compact-vector reads.  The sampled
// directory narrows the binary search to the intervals that begin inside one 2^shift block of ids.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kBlock)
weight_kernel(const __grid_constant__ DeviceIndex ix, const uint64_t* __restrict__ ids, uint64_t n,
              uint64_t* __restrict__ weights_out) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        uint64_t id = __ldcs(ids + i);
        if (id >= ix.num_kmers) id = ix.num_kmers - 1;   // the reference's prev_leq saturates (:252); ids must be < num_kmers
        const uint64_t h = id >> ix.weight_dir_shift;
        uint64_t lo = ld32<true>(ix.weight_dir + h), hi = ld32<true>(ix.weight_dir + h + 1);
        while (lo < hi) {                                // last start <= id inside [lo, hi]
            const uint64_t mid = lo + (hi - lo + 1) / 2;
            if (ld64<true>(ix.weight_starts + mid) <= id) lo = mid; else hi = mid - 1;
        }
        const uint64_t wid = compact_get<true>(ix.weight_values, lo);
        __stcs(weights_out + i, compact_get<true>(ix.weight_dict, wid));
    }
}

// ------------------------------------------------------------------------------------------------
// streaming membership, step 1: one lookup per window.  The reference defines every streamed
// result as equal to dict->lookup(kmer) (include/streaming_query.hpp:107), which is what makes the
// windows independent.  One WARP per read, one lane per window, 32 windows (a "tile") at a time;
// the work that neighbouring windows share is done once per tile:
//   * characters: every lane loads ONE character per 32-character block, converts it to 2 bits
//     (include/kmer.hpp:194) and the warp ORs the shifted codes together with redux.sync: the tile's
//     2-bit packed text appears in all lanes in 2 instructions per 32 bases; lane j's k-mer is a
//     funnel shift of it.  Validity (kmer.hpp:209-219) is a ballot of the per-character flags.
//   * minimizers: the hash of every m-mer position of the tile is computed once (forward and
//     reverse-complement strand) into shared memory; each window then takes the minimum over its
//     k-m+1 positions (leftmost for the forward strand, rightmost for the RC strand, which is the
//     leftmost of the reversed order -- include/util.hpp:262-283 on kmer and on kmer_rc).
//   * on a regular index, windows that miss the forward pass are parked in a per-warp queue and
//     their reverse-complement pass runs on full groups of 32 lanes (as in lookup_kernel).
// Window record: id (u64) + aux = string_id (low 59 bits) | flags: bit 63 backward, bit 62 valid,
// bit 61 first window of its read, bit 60 / 59 the k-mer is the first / last one of its string
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ bool valid_base(uint8_t c) {  // canonicalize_basepair_forward_map, kmer.hpp:209-219
    uint8_t u = c & 0xDF;
    return u == 'A' || u == 'C' || u == 'G' || u == 'T';
}

// 32 characters (one per lane) -> 64-bit word of 2-bit codes, replicated in every lane
__device__ __forceinline__ uint64_t pack_tile_word(uint32_t code, uint32_t lane) {
    const uint32_t sh = 2 * (lane & 15);
    const uint32_t lo = __reduce_or_sync(0xffffffffu, lane < 16 ? code << sh : 0u);
    const uint32_t hi = __reduce_or_sync(0xffffffffu, lane < 16 ? 0u : code << sh);
    return ((uint64_t)hi << 32) | lo;
}
// bits [2*j, 2*j + 64) of the 192-bit little-endian value (w0, w1, w2); j < 32
__device__ __forceinline__ uint64_t funnel64(uint64_t w0, uint64_t w1, uint32_t j) {
    const uint32_t s = 2 * j;
    return s == 0 ? w0 : (w0 >> s) | (w1 << (64 - s));
}

constexpr int kTilePositions = 96;   // m-mer start positions covered by one tile: 32 + (k - m) <= 94

template <int W> struct StreamQueue;
template <> struct StreamQueue<1> {
    uint64_t kmer[64], mini[64], widx[64]; uint32_t pos[64];
    __device__ void put(uint32_t s, Kmer<1> x, Minimizer mi, uint64_t w) { kmer[s] = x.lo; mini[s] = mi.value; pos[s] = mi.pos; widx[s] = w; }
    __device__ Kmer<1> get(uint32_t s) const { return {kmer[s]}; }
};
template <> struct StreamQueue<2> {
    uint64_t lo[64], hi[64], mini[64], widx[64]; uint32_t pos[64];
    __device__ void put(uint32_t s, Kmer<2> x, Minimizer mi, uint64_t w) { lo[s] = x.lo; hi[s] = x.hi; mini[s] = mi.value; pos[s] = mi.pos; widx[s] = w; }
    __device__ Kmer<2> get(uint32_t s) const { return {lo[s], hi[s]}; }
};

constexpr uint64_t kAuxBackward = 1ull << 63, kAuxValid = 1ull << 62, kAuxFirstOfRead = 1ull << 61,
                   kAuxFirstInString = 1ull << 60, kAuxLastInString = 1ull << 59, kAuxSidMask = (1ull << 59) - 1;

__device__ __forceinline__ void store_window(uint64_t* __restrict__ win_id, uint64_t* __restrict__ win_aux, uint64_t w,
                                             const LookupResult& r, bool backward, bool first_of_read, uint32_t k) {
    win_id[w] = r.kmer_id;
    uint64_t aux = kAuxValid | (backward ? kAuxBackward : 0) | (first_of_read ? kAuxFirstOfRead : 0);
    if (r.kmer_id != ~0ull) {
        aux |= r.string_id & kAuxSidMask;
        if (r.kmer_id_in_string == 0) aux |= kAuxFirstInString;
        if (r.kmer_id_in_string + k == r.string_end - r.string_begin) aux |= kAuxLastInString;
    }
    win_aux[w] = aux;
}

template <int W>
__device__ __forceinline__ Kmer<W> tile_kmer(uint64_t a, uint64_t b, uint64_t c, uint32_t lane, uint32_t k);
template <>
__device__ __forceinline__ Kmer<1> tile_kmer<1>(uint64_t a, uint64_t b, uint64_t, uint32_t lane, uint32_t k) {
    return {funnel64(a, b, lane) & low_mask(2 * k)};
}
template <>
__device__ __forceinline__ Kmer<2> tile_kmer<2>(uint64_t a, uint64_t b, uint64_t c, uint32_t lane, uint32_t k) {
    Kmer<2> r{funnel64(a, b, lane), funnel64(b, c, lane)};
    if (2 * k <= 64) { r.lo &= low_mask(2 * k); r.hi = 0; } else r.hi &= low_mask(2 * k - 64);
    return r;
}
__device__ __forceinline__ uint64_t kmer_bits_at(Kmer<1> x, uint32_t pos, uint32_t m) { return (x.lo >> (2 * pos)) & low_mask(2 * m); }
__device__ __forceinline__ uint64_t kmer_bits_at(Kmer<2> x, uint32_t pos, uint32_t m) {
    const uint32_t s = 2 * pos;
    uint64_t w = s == 0 ? x.lo : (s < 64 ? ((x.lo >> s) | (x.hi << (64 - s))) : (x.hi >> (s - 64)));
    return w & low_mask(2 * m);
}

// ------------------------------------------------------------------------------------------------
// streaming membership, step 0: ANCHORS.  Reads that come from the indexed strings are resolved
// without per-window lookups: a few sample windows per read (kAnchorsPerRead) are looked up here,
// one thread each; a positive sample fixes an ungapped alignment of the whole read against one
// string ("diagonal"), and stream_windows_kernel then only has to compare each window's k-mer with
// the string's k-mer at the aligned offset -- the same final comparison a lookup performs
// (spss.hpp:222,259-261), minus minimizer, MPHF and bucket.  A window whose aligned comparison
// succeeds inside the string holds the k-mer with id  offset - string_id*(k-1); because the index
// stores every (canonical) k-mer once -- SSHash's input contract, README "without duplicate
// k-mers" -- that is the id dictionary::lookup returns for it.  Windows that do not match (read
// errors, other strings, absent k-mers) fall through to the regular per-window lookup.
//   forward alignment:  read base t <-> string base diag + t
//   backward alignment: read base t <-> complement of string base diag - t
// ------------------------------------------------------------------------------------------------
constexpr int kAnchorsPerRead = 3;
struct Anchor { int64_t diag; uint64_t info; };   // info: bit 63 valid, bit 62 backward, low 62 bits string id

template <int W>
__global__ void __launch_bounds__(kBlock, 4)
stream_anchor_kernel(const __grid_constant__ DeviceIndex ix, const char* __restrict__ bases,
                     const uint64_t* __restrict__ read_begins, const uint64_t* __restrict__ read_ends, uint64_t num_reads,
                     Anchor* __restrict__ anchors) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    const uint32_t k = ix.k;
    for (uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; t < num_reads * kAnchorsPerRead; t += stride) {
        const uint64_t r = t / kAnchorsPerRead, s = t % kAnchorsPerRead;
        const uint64_t b = read_begins[r], len = read_ends[r] - b;
        Anchor a{0, 0};
        if (len >= k) {
            const uint64_t nwin = len - k + 1, p = s * (nwin / kAnchorsPerRead);
            const char* q = bases + b + p;
            bool valid = true;
            for (uint32_t j = 0; j < k; ++j) valid &= valid_base((uint8_t)q[j]);
            if (valid) {
                LookupResult res;
                lookup_kmer<W, false>(ix, pack_ascii<W>(q, k), true, res);
                if (res.kmer_id != ~0ull) {
                    const bool back = res.kmer_orientation < 0;
                    a.diag = back ? (int64_t)(res.kmer_offset + k - 1 + p) : (int64_t)res.kmer_offset - (int64_t)p;
                    a.info = (1ull << 63) | (back ? (1ull << 62) : 0) | res.string_id;
                }
            }
        }
        anchors[t] = a;
    }
}

template <int W>
__global__ void __launch_bounds__(kBlock, 4)
stream_windows_kernel(const __grid_constant__ DeviceIndex ix, const char* __restrict__ bases,
                      const uint64_t* __restrict__ read_begins, const uint64_t* __restrict__ read_ends,
                      const uint64_t* __restrict__ win_offsets,
                      uint64_t num_reads, const Anchor* __restrict__ anchors, uint64_t* __restrict__ win_id,
                      uint64_t* __restrict__ win_aux, unsigned long long* __restrict__ next_read) {
    __shared__ uint64_t hash_f[kBlock / 32][kTilePositions];
    __shared__ uint64_t hash_r[kBlock / 32][kTilePositions];
    __shared__ StreamQueue<W> queues[kBlock / 32];
    __shared__ int64_t anchor_s[kBlock / 32][kAnchorsPerRead][4];   // diag, string begin, string end, sid | flags
    const uint32_t lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    uint64_t* hf = hash_f[wib];
    uint64_t* hr = hash_r[wib];
    StreamQueue<W>& q = queues[wib];
    const uint32_t k = ix.k, m = ix.m, n = k - m + 1;
    const uint64_t mmask = low_mask(2 * m), magic = ix.magic;
    const bool canonical = ix.canonical != 0;
    constexpr int NBLK = W == 1 ? 2 : 3;     // 32-character blocks a tile needs: 32 + k - 1 <= 62 / 94
    uint32_t queued = 0;
    // Reads cost very different amounts of work (a read resolved by its anchor is ~10x cheaper than a
    // read whose every window needs a lookup), so warps claim reads dynamically, kReadsPerClaim at a time.
    constexpr uint64_t kReadsPerClaim = 8;
    for (;;) {
        uint64_t claim = 0;
        if (lane == 0) claim = atomicAdd(next_read, (unsigned long long)kReadsPerClaim);
        claim = __shfl_sync(0xffffffffu, claim, 0);
        if (claim >= num_reads) break;
        const uint64_t claim_end = claim + kReadsPerClaim < num_reads ? claim + kReadsPerClaim : num_reads;
    for (uint64_t r = claim; r < claim_end; ++r) {
        const uint64_t b = read_begins[r], len = read_ends[r] - b;
        if (len < k) continue;
        const uint64_t nwin = len - k + 1, w0 = win_offsets[r];
        const char* s = bases + b;
        // this read's distinct alignments, kept in shared memory (warp-uniform data)
        bool any_back = false, any_anchor = false;
        if (anchors) {
            __syncwarp();
            bool ok = false, back = false;
            int64_t diag = 0;
            uint64_t sid = 0;
            if (lane < kAnchorsPerRead) {
                const Anchor a = anchors[r * kAnchorsPerRead + lane];
                ok = (a.info >> 63) != 0;
                back = (a.info >> 62 & 1) != 0;
                diag = a.diag;
                sid = a.info & ((1ull << 62) - 1);
            }
#pragma unroll
            for (int aj = 0; aj < kAnchorsPerRead - 1; ++aj) {   // same string + diagonal + strand = same alignment
                const bool ok_j = __shfl_sync(0xffffffffu, ok, aj), back_j = __shfl_sync(0xffffffffu, back, aj);
                const int64_t diag_j = __shfl_sync(0xffffffffu, diag, aj);
                const uint64_t sid_j = __shfl_sync(0xffffffffu, sid, aj);
                if ((int)lane > aj && ok_j && back_j == back && diag_j == diag && sid_j == sid) ok = false;
            }
            if (lane < kAnchorsPerRead) {
                int64_t* a = anchor_s[wib][lane];
                a[0] = diag;
                a[1] = ok ? (int64_t)ld64<true>(ix.ends + sid) : 0;
                a[2] = ok ? (int64_t)ld64<true>(ix.ends + sid + 1) : 0;
                a[3] = (int64_t)(sid | (ok ? 1ull << 63 : 0) | (back ? 1ull << 62 : 0));
            }
            any_anchor = __ballot_sync(0xffffffffu, ok) != 0;
            any_back = __ballot_sync(0xffffffffu, ok && back) != 0;
            __syncwarp();
        }
        for (uint64_t wb = 0; wb < nwin; wb += 32) {
            // ---- tile text: 2-bit packed words + invalid-character masks ------------------------------
            uint64_t word[NBLK + 1];
            uint32_t inval[NBLK];
#pragma unroll
            for (int blk = 0; blk < NBLK; ++blk) {
                const uint64_t p = wb + 32 * blk + lane;
                const uint8_t c = p < len ? (uint8_t)s[p] : (uint8_t)0;
                word[blk] = pack_tile_word((c >> 1) & 3, lane);
                inval[blk] = __ballot_sync(0xffffffffu, !valid_base(c));
            }
            word[NBLK] = 0;
            const uint64_t w = wb + lane;
            const Kmer<W> x = tile_kmer<W>(word[0], word[1], W == 1 ? 0 : word[2], lane, k);
            // window valid <=> no invalid character among its k characters
            bool valid = w < nwin;
            {
                const uint64_t v01 = ((uint64_t)inval[1] << 32) | inval[0];
                uint64_t bad = (v01 >> lane) & low_mask(k < 64 - lane ? k : 64 - lane);
                if (W == 2 && k + lane > 64) bad |= (uint64_t)inval[NBLK - 1] & low_mask(k + lane - 64);
                valid = valid && bad == 0;
            }
            // ---- windows resolved by an anchor's alignment: one k-mer comparison, no lookup ---------------
            Kmer<W> xr = x;
            bool have_xr = false;
            if (any_anchor) {
                bool resolved = false;
                if (any_back) { xr = kmer_rc(x, k); have_xr = true; }
#pragma unroll
                for (int ai = 0; ai < kAnchorsPerRead; ++ai) {
                    const int64_t* a = anchor_s[wib][ai];
                    const uint64_t info = (uint64_t)a[3];
                    if (!(info >> 63)) continue;         // warp-uniform
                    const bool back = (info >> 62 & 1) != 0;
                    const uint64_t sid = info & ((1ull << 62) - 1);
                    const int64_t oj = back ? a[0] - (int64_t)w - (int64_t)(k - 1) : a[0] + (int64_t)w;
                    if (valid && !resolved && oj >= a[1] && oj + (int64_t)k <= a[2]) {
                        const Kmer<W> sk = read_kmer(ix, (uint64_t)oj, k, (Kmer<W>*)nullptr);
                        if (kmer_eq(sk, back ? xr : x)) {
                            resolved = true;
                            win_id[w0 + w] = (uint64_t)oj - sid * (k - 1);
                            win_aux[w0 + w] = kAuxValid | sid | (back ? kAuxBackward : 0) | (w == 0 ? kAuxFirstOfRead : 0) |
                                              (oj == a[1] ? kAuxFirstInString : 0) | (oj + (int64_t)k == a[2] ? kAuxLastInString : 0);
                        }
                    }
                }
                if (w < nwin && !valid) { win_id[w0 + w] = ~0ull; win_aux[w0 + w] = w == 0 ? kAuxFirstOfRead : 0; }
                valid = valid && !resolved;              // from here on: windows that still need a lookup
                if (__ballot_sync(0xffffffffu, valid) == 0) continue;
            } else if (w < nwin && !valid) { win_id[w0 + w] = ~0ull; win_aux[w0 + w] = w == 0 ? kAuxFirstOfRead : 0; }
            // ---- m-mer hashes of the tile, both strands (mixer_64::hash, hash_util.hpp:91) -------------
            __syncwarp();
#ifdef SSHASH_STREAM_EXACT_MIN
#pragma unroll
            for (int blk = 0; blk < NBLK; ++blk) {
                const uint32_t p = 32 * blk + lane;
                if (p < 32 + n - 1) {
                    const uint64_t mm = funnel64(word[blk], word[blk + 1], lane) & mmask;
                    hf[p] = (mm * SSHASH_MIX_C) ^ magic;
                    hr[p] = (mmer_rc(mm, m) * SSHASH_MIX_C) ^ magic;
                }
            }
            __syncwarp();
            Minimizer mf{~0ull, 0}, mr{~0ull, 0};
            if (valid) {
                uint64_t bf = ~0ull, br = ~0ull;
                uint32_t pf = 0, pr = 0;
                for (uint32_t i = 0; i < n; ++i) {
                    const uint64_t a = hf[lane + i], c = hr[lane + i];
                    if (a < bf) { bf = a; pf = i; }            // leftmost minimum of kmer
                    if (c <= br) { br = c; pr = i; }           // rightmost here = leftmost of kmer_rc
                }
                if (!have_xr) xr = kmer_rc(x, k);
                mf.pos = pf; mf.value = bf == ~0ull ? ~0ull : kmer_bits_at(x, pf, m);
                // util.hpp:268-270: with no hash below UINT64_MAX the reference keeps (all ones, pos 0)
                if (br == ~0ull) { mr.pos = 0; mr.value = ~0ull; }
                else { mr.pos = n - 1 - pr; mr.value = kmer_bits_at(xr, mr.pos, m); }
                if (bf == ~0ull) mf.pos = 0;
            }
#else
            // Per tile position p the two strands' hashes are reduced to their 25 high bits and packed with
            // the position both ways: lo = h25 << 7 | p, hi = h25 << 7 | (127 - p).  A window's minimum over
            // the lo keys names the LEFTMOST position of its smallest h25, over the hi keys the RIGHTMOST;
            // when they agree exactly one m-mer has the smallest 25 bits, hence the smallest 64-bit hash,
            // whichever tie rule applies (leftmost on the k-mer, include/util.hpp:262-283; rightmost =
            // leftmost on the reverse complement).  Otherwise -- a repeated m-mer, a 2^-25 collision, or
            // h25 all ones (the reference's "no hash below UINT64_MAX" case) -- the exact scan decides.
#pragma unroll
            for (int blk = 0; blk < NBLK; ++blk) {
                const uint32_t p = 32 * blk + lane;
                if (p < 32 + n - 1) {
                    const uint64_t mm = funnel64(word[blk], word[blk + 1], lane) & mmask;
                    const uint32_t f25 = (uint32_t)((((mm * SSHASH_MIX_C) ^ magic) >> 39) << 7);
                    const uint32_t r25 = (uint32_t)((((mmer_rc(mm, m) * SSHASH_MIX_C) ^ magic) >> 39) << 7);
                    hf[p] = ((uint64_t)(f25 | (127u - p)) << 32) | (f25 | p);
                    hr[p] = ((uint64_t)(r25 | (127u - p)) << 32) | (r25 | p);
                }
            }
            __syncwarp();
            Minimizer mf{~0ull, 0}, mr{~0ull, 0};
            if (valid) {
                uint32_t fl = ~0u, fr = ~0u, rl = ~0u, rr = ~0u;
                for (uint32_t i = 0; i < n; ++i) {
                    const uint64_t a = hf[lane + i], c = hr[lane + i];
                    fl = min(fl, (uint32_t)a); fr = min(fr, (uint32_t)(a >> 32));
                    rl = min(rl, (uint32_t)c); rr = min(rr, (uint32_t)(c >> 32));
                }
                if (!have_xr) xr = kmer_rc(x, k);
                const uint32_t pf = (fl & 127u) - lane, pr = (rl & 127u) - lane;     // positions inside the window
                if (pf == 127u - (fr & 127u) - lane && (fl >> 7) != 0x1ffffffu) {
                    mf.pos = pf; mf.value = kmer_bits_at(x, pf, m);
                } else {
                    mf = ix.m <= 16 ? compute_minimizer_exact<true>(x, k, m, magic) : compute_minimizer_exact<false>(x, k, m, magic);
                }
                if (pr == 127u - (rr & 127u) - lane && (rl >> 7) != 0x1ffffffu) {
                    mr.pos = n - 1 - pr; mr.value = kmer_bits_at(xr, mr.pos, m);
                } else {
                    mr = ix.m <= 16 ? compute_minimizer_exact<true>(xr, k, m, magic) : compute_minimizer_exact<false>(xr, k, m, magic);
                }
            }
#endif
            // ---- lookups ----------------------------------------------------------------------------
            bool park = false;
            if (valid) {
                LookupResult res;
                if (canonical) {                                // dictionary.cpp:24-42
                    bool found;
                    if (mf.value < mr.value) found = lookup_canonical_with<W, false, true>(ix, x, xr, mf, res);
                    else if (mr.value < mf.value) found = lookup_canonical_with<W, false, true>(ix, x, xr, mr, res);
                    else {
                        found = lookup_canonical_with<W, false, true>(ix, x, xr, mf, res);
                        if (!found) found = lookup_canonical_with<W, false, true>(ix, x, xr, mr, res);
                    }
                    store_window(win_id, win_aux, w0 + w, res, found && res.kmer_orientation < 0, w == 0, k);
                } else if (lookup_regular_with<W, false, true>(ix, x, mf, res)) {
                    store_window(win_id, win_aux, w0 + w, res, false, w == 0, k);
                } else {
                    park = true;
                }
            }
            if (canonical) continue;
            const uint32_t mask = __ballot_sync(0xffffffffu, park);
            if (park) q.put(queued + __popc(mask & ((1u << lane) - 1)), xr, mr, (w0 + w) | (w == 0 ? kAuxFirstOfRead : 0));
            queued += __popc(mask);
            __syncwarp();
            if (queued >= 32) {
                queued -= 32;
                const uint32_t e = queued + lane;
                LookupResult res;
                const bool found = lookup_regular_with<W, false, true>(ix, q.get(e), Minimizer{q.mini[e], q.pos[e]}, res);
                store_window(win_id, win_aux, q.widx[e] & ~kAuxFirstOfRead, res, found, (q.widx[e] & kAuxFirstOfRead) != 0, k);
                __syncwarp();
            }
        }
    }
    }
    if (!canonical && lane < queued) {
        LookupResult res;
        const bool found = lookup_regular_with<W, false, true>(ix, q.get(lane), Minimizer{q.mini[lane], q.pos[lane]}, res);
        store_window(win_id, win_aux, q.widx[lane] & ~kAuxFirstOfRead, res, found, (q.widx[lane] & kAuxFirstOfRead) != 0, k);
    }
}

// ------------------------------------------------------------------------------------------------
// streaming membership, step 2: the reference's per-read state machine
// (include/streaming_query.hpp:56-197) replayed over the window records by one THREAD per read to
// produce the searches / extensions / negative / invalid counters and the streamed ids.  A window
// is an EXTENSION when the previous window left `remaining > 0` and the next k-mer of the string
// (read through a literal restatement of kmer_iterator, include/kmer_iterator.hpp:8-86) equals the
// window's k-mer or its reverse complement; otherwise it is a seed(): positive -> SEARCH.
// The restated iterator keeps the reference's word order in fill_buff_reverse (:72-79), which for
// the 128-bit k-mer type swaps the two halves: in a max_k = 63 build backward extensions therefore
// (almost) never match and are counted as searches, exactly as the reference does.
// ------------------------------------------------------------------------------------------------
template <int W> struct KmerIter;
template <> struct KmerIter<2> {
    uint64_t pos, avail, lo, hi;  // buff = hi:lo
    __device__ void at(uint64_t p) { pos = p; avail = 0; lo = hi = 0; }
    __device__ uint64_t word(const DeviceIndex& ix, uint64_t p) const { return read_word64(ix.strings, p); }
    __device__ void fill(const DeviceIndex& ix) { hi = word(ix, pos + 64); lo = word(ix, pos); avail = 128; }
    __device__ void fill_reverse(const DeviceIndex& ix) {
        uint64_t base = pos > 128 ? pos : 128;
        hi = word(ix, base - 128);  // appended first -> ends up in the HIGH half (reference quirk)
        lo = word(ix, base - 64);
        avail = pos < 128 ? pos : 128;
        uint64_t pad = 128 - avail;
        if (pad >= 128) { lo = hi = 0; }
        else if (pad >= 64) { hi = lo << (pad - 64); lo = 0; }
        else if (pad) { hi = (hi << pad) | (lo >> (64 - pad)); lo <<= pad; }
    }
    __device__ Kmer<2> get(const DeviceIndex& ix) {
        if (avail < 2 * ix.k) fill(ix);
        uint32_t b = 2 * ix.k;
        return b <= 64 ? Kmer<2>{lo & low_mask(b), 0} : Kmer<2>{lo, hi & low_mask(b - 64)};
    }
    __device__ Kmer<2> get_reverse(const DeviceIndex& ix) {
        if (avail < 2 * ix.k) fill_reverse(ix);
        uint32_t s = 128 - 2 * ix.k;  // >= 2
        return s >= 64 ? Kmer<2>{hi >> (s - 64), 0} : Kmer<2>{(lo >> s) | (hi << (64 - s)), hi >> s};
    }
    __device__ void next(const DeviceIndex& ix) {
        if (avail < 2) fill(ix);
        lo = (lo >> 2) | (hi << 62); hi >>= 2; avail -= 2; pos += 2;
    }
    __device__ void next_reverse(const DeviceIndex& ix) {
        if (avail < 2) fill_reverse(ix);
        hi = (hi << 2) | (lo >> 62); lo <<= 2; avail -= 2; pos -= 2;
    }
};

// 64-bit k-mers: the reference iterator (include/kmer_iterator.hpp:8-86) has no word-order quirk there; after
// at(p) and s steps, get() is the k-mer at bits [p + 2s, p + 2s + 2k) and get_reverse() the one at
// bits [p - 2s - 2k, p - 2s): plain reads of `strings`.
template <> struct KmerIter<1> {
    uint64_t pos;
    __device__ void at(uint64_t p) { pos = p; }
    __device__ void next(const DeviceIndex&) { pos += 2; }
    __device__ void next_reverse(const DeviceIndex&) { pos -= 2; }
    __device__ Kmer<1> get(const DeviceIndex& ix) const { return read_kmer(ix, pos >> 1, (Kmer<1>*)nullptr); }
    __device__ Kmer<1> get_reverse(const DeviceIndex& ix) const { return read_kmer(ix, (pos >> 1) - ix.k, (Kmer<1>*)nullptr); }
};

template <int W> struct RollingKmer;
template <> struct RollingKmer<1> {
    Kmer<1> x, xr;
    __device__ void init(const char* s, uint32_t k) {
        x.lo = 0; xr.lo = 0;
        for (uint32_t i = 0; i + 1 < k; ++i) push((uint8_t)s[i], k);
    }
    __device__ void push(uint8_t c, uint32_t k) {
        const uint64_t code = (c >> 1) & 3;
        x.lo = (x.lo >> 2) | (code << (2 * (k - 1)));
        xr.lo = ((xr.lo << 2) | (code ^ 2)) & low_mask(2 * k);
    }
};
template <> struct RollingKmer<2> {
    Kmer<2> x, xr;
    __device__ void init(const char* s, uint32_t k) {
        x.lo = x.hi = 0; xr.lo = xr.hi = 0;
        for (uint32_t i = 0; i + 1 < k; ++i) push((uint8_t)s[i], k);
    }
    __device__ void push(uint8_t c, uint32_t k) {
        const uint64_t code = (c >> 1) & 3;
        x.lo = (x.lo >> 2) | (x.hi << 62);
        x.hi >>= 2;
        const uint32_t top = 2 * (k - 1);
        if (top < 64) x.lo |= code << top; else x.hi |= code << (top - 64);
        xr.hi = (xr.hi << 2) | (xr.lo >> 62);
        xr.lo = (xr.lo << 2) | (code ^ 2);
        if (2 * k <= 64) { xr.lo &= low_mask(2 * k); xr.hi = 0; } else xr.hi &= low_mask(2 * k - 64);
    }
};

template <int W>
__global__ void __launch_bounds__(kBlock)
stream_scan_kernel(const __grid_constant__ DeviceIndex ix, const char* __restrict__ bases,
                   const uint64_t* __restrict__ read_begins, const uint64_t* __restrict__ read_ends,
                   const uint64_t* __restrict__ win_offsets,
                   uint64_t num_reads, const uint64_t* __restrict__ win_id, const uint64_t* __restrict__ win_aux,
                   uint64_t* __restrict__ ids_out, unsigned long long* __restrict__ counters) {
    const uint32_t k = ix.k;
    unsigned long long n_search = 0, n_ext = 0, n_neg = 0, n_inv = 0, n_kmers = 0;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; r < num_reads; r += stride) {
        const uint64_t b = read_begins[r], len = read_ends[r] - b;
        if (len < k) continue;
        const uint64_t nwin = len - k + 1, w0 = win_offsets[r];
        n_kmers += nwin;
        uint64_t remaining = 0, cur_id = ~0ull;
        bool backward = false;
        KmerIter<W> it; it.at(0);
        // rolling window k-mer and its reverse complement (streaming_query.hpp:68-80)
        RollingKmer<W> roll;
        roll.init(bases + b, k);
        for (uint64_t i = 0; i < nwin; ++i) {
            roll.push((uint8_t)bases[b + i + k - 1], k);
            const uint64_t aux = win_aux[w0 + i];
            if (!(aux >> 62 & 1)) {                      // invalid window: reset() (streaming_query.hpp:59-65)
                n_inv += 1; remaining = 0; cur_id = ~0ull;
                if (ids_out) ids_out[w0 + i] = ~0ull;
                continue;
            }
            bool extended = false;
            if (remaining != 0) {                        // :88-99
                const Kmer<W> x = roll.x, xr = roll.xr;
                Kmer<W> e;
                if (!backward) { it.next(ix); e = it.get(ix); }
                else { it.next_reverse(ix); e = it.get_reverse(ix); }
                if (kmer_eq(e, x) || kmer_eq(e, xr)) {
                    n_ext += 1;
                    cur_id += backward ? ~0ull : 1ull;   // kmer_id += orientation
                    remaining -= 1;
                    extended = true;
                }
            }
            if (!extended) {                             // seed() :144-197
                remaining = 0;
                cur_id = win_id[w0 + i];
                if (cur_id == ~0ull) { n_neg += 1; }
                else {
                    n_search += 1;
                    backward = (aux >> 63) != 0;
                    const uint64_t sid = aux & kAuxSidMask;
                    const uint64_t sb = ld64<true>(ix.ends + sid), se = ld64<true>(ix.ends + sid + 1);
                    const uint64_t ko = cur_id + sid * (k - 1);          // kmer_offset in bases
                    const uint64_t in_string = ko - sb;
                    uint64_t bitpos = 2 * ko;
                    remaining = (se - sb - k) - in_string;
                    if (backward) { bitpos += 2 * k; remaining = in_string; }
                    it.at(bitpos);
                }
            }
            if (ids_out) ids_out[w0 + i] = cur_id;
        }
    }
    // block reduction -> 5 atomics per block
    __shared__ unsigned long long sh[5];
    if (threadIdx.x < 5) sh[threadIdx.x] = 0;
    __syncthreads();
    unsigned long long v[5] = {n_kmers, n_search, n_ext, n_neg, n_inv};
#pragma unroll
    for (int j = 0; j < 5; ++j) {
        unsigned long long x = v[j];
        for (int o = 16; o > 0; o >>= 1) x += __shfl_down_sync(0xffffffffu, x, o);
        if ((threadIdx.x & 31) == 0 && x) atomicAdd(&sh[j], x);
    }
    __syncthreads();
    if (threadIdx.x < 5 && sh[threadIdx.x]) atomicAdd(&counters[threadIdx.x], sh[threadIdx.x]);
}


// ------------------------------------------------------------------------------------------------
// streaming membership, step 2 for 64-bit k-mers: the same counters WITHOUT the sequential replay.
// With the 64-bit kmer_iterator the string k-mer the reference compares against (:88-99) is exactly
// the next k-mer of the string in the previous window's direction, and the index holds every k-mer
// once, so
//     window i is an EXTENSION  <=>  windows i-1 and i are positive, window i-1's k-mer is not the
//     last (forward) / first (backward) k-mer of its string (remaining > 0, :189-195), and
//     id(i) == id(i-1) + orientation(i-1);
// every other positive window is a SEARCH.  The state the reference carries (id, orientation,
// remaining) is a function of window i-1's lookup result alone (:107 asserts streamed == looked-up),
// so windows are classified independently, one thread each; the streamed ids ARE the window ids.
// (The 128-bit build keeps stream_scan_kernel: its fill_buff_reverse word-order quirk makes the
// outcome depend on the iterator's history.)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kBlock)
stream_classify_kernel(const uint64_t* __restrict__ win_offsets, uint64_t num_reads, const uint64_t* __restrict__ win_id,
                       const uint64_t* __restrict__ win_aux, unsigned long long* __restrict__ counters) {
    const uint64_t total = win_offsets[num_reads];
    unsigned long long n_search = 0, n_ext = 0, n_neg = 0, n_inv = 0, n_kmers = 0;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; g < total; g += stride) {
        const uint64_t aux = win_aux[g], id = win_id[g];
        n_kmers += 1;
        if (!(aux & kAuxValid)) { n_inv += 1; continue; }
        if (id == ~0ull) { n_neg += 1; continue; }
        bool ext = false;
        if (!(aux & kAuxFirstOfRead)) {
            const uint64_t paux = win_aux[g - 1], pid = win_id[g - 1];
            if ((paux & kAuxValid) && pid != ~0ull) {
                ext = (paux & kAuxBackward) ? (!(paux & kAuxFirstInString) && id + 1 == pid)
                                            : (!(paux & kAuxLastInString) && id == pid + 1);
            }
        }
        if (ext) n_ext += 1; else n_search += 1;
    }
    __shared__ unsigned long long sh[5];
    if (threadIdx.x < 5) sh[threadIdx.x] = 0;
    __syncthreads();
    unsigned long long v[5] = {n_kmers, n_search, n_ext, n_neg, n_inv};
#pragma unroll
    for (int j = 0; j < 5; ++j) {
        unsigned long long x = v[j];
        for (int o = 16; o > 0; o >>= 1) x += __shfl_down_sync(0xffffffffu, x, o);
        if ((threadIdx.x & 31) == 0 && x) atomicAdd(&sh[j], x);
    }
    __syncthreads();
    if (threadIdx.x < 5 && sh[threadIdx.x]) atomicAdd(&counters[threadIdx.x], sh[threadIdx.x]);
}

// ------------------------------------------------------------------------------------------------
// Does the index keep SSHash's input contract -- every k-mer once, and on a regular index never
// together with its reverse complement (README)?  The streaming shortcuts (anchors,
// per-window classification) are exact only then; the reference itself still answers such indexes,
// with results that depend on the state of the stream, so they take the literal replay instead.
// One thread per text offset: the k-mer there must look up to its own id, and (regular index) its
// reverse complement must be absent.
// ------------------------------------------------------------------------------------------------
template <int W>
__global__ void __launch_bounds__(kBlock)
distinct_check_kernel(const __grid_constant__ DeviceIndex ix, uint32_t* __restrict__ flag) {
    // the text ends at the last end-point; `strings` may hold a few more (sentinel / padding) bits
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x, n_bases = ix.ends[ix.n_ends - 1];
    const uint32_t k = ix.k;
    bool bad = false;
    for (uint64_t o = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; o + k <= n_bases; o += stride) {
        uint64_t sb, se;
        const uint64_t sid = locate_string(ix, o, sb, se);
        if (o + k > se) continue;                         // window straddles two strings
        const Kmer<W> x = read_kmer(ix, o, (Kmer<W>*)nullptr);
        LookupResult r;
        const bool found = ix.canonical ? lookup_canonical<W, false>(ix, x, r) : lookup_regular<W, false>(ix, x, r);
        if (!found || r.kmer_id != o - sid * (k - 1)) bad = true;
        if (!ix.canonical && lookup_regular<W, false>(ix, kmer_rc(x, k), r)) bad = true;
    }
    if (bad) atomicOr(flag, 1u);
}

// ------------------------------------------------------------------------------------------------
// win_offsets = exclusive prefix sum of max(0, len_r - k + 1) over the reads (three small kernels:
// per-block totals, a single-block scan of the totals, per-block exclusive scan + offset).
// win_offsets has num_reads + 1 entries; the last one is the total number of windows.
// ------------------------------------------------------------------------------------------------
constexpr int kScanItems = 8;                       // reads per thread
constexpr int kScanTile = kBlock * kScanItems;      // reads per block

// Reads are given as spans: read r = bases[read_begins[r], read_ends[r]).  Contiguous reads pass
// (read_offsets, read_offsets + 1); the device-side FASTA/FASTQ parser passes the spans of the
// sequence lines inside the raw file bytes.
__device__ __forceinline__ uint64_t windows_of(const uint64_t* __restrict__ rb, const uint64_t* __restrict__ re, uint64_t r,
                                               uint64_t num_reads, uint32_t k) {
    if (r >= num_reads) return 0;
    uint64_t len = re[r] - rb[r];
    return len >= k ? len - k + 1 : 0;
}

__device__ __forceinline__ uint64_t block_exclusive_scan(uint64_t v, uint64_t& total) {
    __shared__ uint64_t warp_sums[kBlock / 32];
    const uint32_t lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    uint64_t inc = v;
    for (int o = 1; o < 32; o <<= 1) { uint64_t t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
    if (lane == 31) warp_sums[wid] = inc;
    __syncthreads();
    if (wid == 0) {
        uint64_t w = lane < kBlock / 32 ? warp_sums[lane] : 0;
        for (int o = 1; o < 32; o <<= 1) { uint64_t t = __shfl_up_sync(0xffffffffu, w, o); if (lane >= o) w += t; }
        if (lane < kBlock / 32) warp_sums[lane] = w;
    }
    __syncthreads();
    uint64_t base = wid ? warp_sums[wid - 1] : 0;
    total = warp_sums[kBlock / 32 - 1];
    __syncthreads();
    return base + inc - v;
}

__global__ void __launch_bounds__(kBlock)
win_block_sums_kernel(uint32_t k, const uint64_t* __restrict__ rb, const uint64_t* __restrict__ re, uint64_t num_reads,
                      uint64_t* __restrict__ block_sums) {
    uint64_t r0 = (uint64_t)blockIdx.x * kScanTile + (uint64_t)threadIdx.x * kScanItems, s = 0;
    for (int j = 0; j < kScanItems; ++j) s += windows_of(rb, re, r0 + j, num_reads, k);
    uint64_t total;
    block_exclusive_scan(s, total);
    if (threadIdx.x == 0) block_sums[blockIdx.x] = total;
}

__global__ void __launch_bounds__(kBlock)
win_scan_sums_kernel(uint64_t* __restrict__ block_sums, uint64_t nblocks) {   // single block
    uint64_t carry = 0;
    for (uint64_t b0 = 0; b0 < nblocks; b0 += kBlock) {
        uint64_t i = b0 + threadIdx.x;
        uint64_t v = i < nblocks ? block_sums[i] : 0, total;
        uint64_t ex = block_exclusive_scan(v, total);
        if (i < nblocks) block_sums[i] = carry + ex;
        carry += total;
    }
}

__global__ void __launch_bounds__(kBlock)
win_offsets_kernel(uint32_t k, const uint64_t* __restrict__ rb, const uint64_t* __restrict__ re, uint64_t num_reads,
                   const uint64_t* __restrict__ block_sums, uint64_t* __restrict__ win_offsets) {
    uint64_t r0 = (uint64_t)blockIdx.x * kScanTile + (uint64_t)threadIdx.x * kScanItems, s = 0;
    uint64_t c[kScanItems];
    for (int j = 0; j < kScanItems; ++j) { c[j] = windows_of(rb, re, r0 + j, num_reads, k); s += c[j]; }
    uint64_t total;
    uint64_t off = block_sums[blockIdx.x] + block_exclusive_scan(s, total);
    for (int j = 0; j < kScanItems; ++j) {
        if (r0 + j <= num_reads) win_offsets[r0 + j] = off;   // entry num_reads = grand total
        off += c[j];
    }
}


// ------------------------------------------------------------------------------------------------
// Device-side FASTA / FASTQ record parsing (the step before the path; the reference's drivers,
// src/query.cpp:53-108, call std::getline per line on the host).  Input: a chunk of raw file bytes
// in HBM.  Records are POSITIONAL, exactly as in the reference: FASTQ = 4 lines per record with the
// sequence on the 2nd, FASTA = 2 lines per record with the sequence on the 2nd; no character of a
// header/quality line is ever interpreted.  Three small kernels index the newlines
//   newline_count_kernel     '\n' per 4 KB tile (16 bytes per thread, one 128-bit load)
//   win_scan_sums_kernel     exclusive scan of the tile counts (+ grand total = number of lines)
//   newline_positions_kernel line_start[j + 1] = position after the j-th newline
// and read_spans_kernel turns lines into read spans [begin, end) over the raw bytes, which the
// streaming kernels consume in place: bases are never copied or compacted.
// ------------------------------------------------------------------------------------------------
constexpr int kParseBytesPerThread = 16;
constexpr int kParseTile = kBlock * kParseBytesPerThread;

// bit j set <=> raw[p + j] == '\n' (p is 16-byte aligned; bytes at or beyond n do not count)
__device__ __forceinline__ uint32_t newline_mask16(const uint8_t* __restrict__ raw, uint64_t p, uint64_t n) {
    if (p >= n) return 0;
    const uint4 v = __ldg(reinterpret_cast<const uint4*>(raw + p));   // the buffer is padded to 16 bytes
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
    uint32_t mask = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const uint32_t e = __vcmpeq4(w[i], 0x0a0a0a0au);             // 0xff in every matching byte
        mask |= (((e >> 7) & 1u) | ((e >> 14) & 2u) | ((e >> 21) & 4u) | ((e >> 28) & 8u)) << (4 * i);
    }
    const uint64_t left = n - p;
    return left >= 16 ? mask : mask & ((1u << left) - 1u);
}

__global__ void __launch_bounds__(kBlock)
newline_count_kernel(const uint8_t* __restrict__ raw, uint64_t n, uint64_t* __restrict__ tile_counts) {
    const uint64_t p = (uint64_t)blockIdx.x * kParseTile + (uint64_t)threadIdx.x * kParseBytesPerThread;
    uint64_t total;
    block_exclusive_scan(__popc(newline_mask16(raw, p, n)), total);
    if (threadIdx.x == 0) tile_counts[blockIdx.x] = total;
}

__global__ void __launch_bounds__(kBlock)
newline_positions_kernel(const uint8_t* __restrict__ raw, uint64_t n, const uint64_t* __restrict__ tile_offsets,
                         uint64_t* __restrict__ line_start) {
    const uint64_t p = (uint64_t)blockIdx.x * kParseTile + (uint64_t)threadIdx.x * kParseBytesPerThread;
    uint32_t mask = newline_mask16(raw, p, n);
    uint64_t total;
    uint64_t rank = tile_offsets[blockIdx.x] + block_exclusive_scan(__popc(mask), total);
    if (blockIdx.x == 0 && threadIdx.x == 0) line_start[0] = 0;
    while (mask) {
        const uint32_t j = __ffs(mask) - 1;
        mask &= mask - 1;
        line_start[++rank] = p + j + 1;
    }
}

// record r = lines [stride * r, stride * (r + 1)); its sequence is line stride * r + 1, without the '\n'
__global__ void __launch_bounds__(kBlock)
read_spans_kernel(const uint64_t* __restrict__ line_start, uint64_t num_records, uint32_t stride,
                  uint64_t* __restrict__ read_begins, uint64_t* __restrict__ read_ends) {
    const uint64_t step = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; r < num_records; r += step) {
        read_begins[r] = line_start[stride * r + 1];
        read_ends[r] = line_start[stride * r + 2] - 1;
    }
}

}  // namespace

std::atomic<uint64_t> g_launches{0};
uint64_t kernel_launch_count() { return g_launches.load(); }

cudaError_t launch_lookup(const DeviceIndex& ix, const LaunchCtx& ctx, const void* queries, bool ascii, uint64_t n, bool check_rc,
                          uint64_t* ids, sshash_lookup_result* full, uint8_t* member, cudaStream_t stream, uint32_t* ids32) {
    if (n == 0) return cudaSuccess;
    if (ids32) {                                           // 32-bit ids: packed queries, ids only
        if (ascii || ids || full || member) return cudaErrorInvalidValue;
        const int g32 = grid_for(n, ctx.sm_count, 2 * (ix.canonical && ix.kmer_words == 2 ? kLookupMinBlocksWideCanon : kLookupMinBlocks));
        uint64_t* out = reinterpret_cast<uint64_t*>(ids32);
        const int rcf = check_rc ? 1 : 0;
        if (ix.kmer_words == 1)
            return ix.canonical ? launch(lookup_kernel<1, 3, false, true, kLookupMinBlocks>, g32, stream, ctx, ix, queries, n, rcf, out, full, member)
                                : launch(lookup_kernel<1, 3, false, false, kLookupMinBlocks>, g32, stream, ctx, ix, queries, n, rcf, out, full, member);
        return ix.canonical ? launch(lookup_kernel<2, 3, false, true, kLookupMinBlocksWideCanon>, g32, stream, ctx, ix, queries, n, rcf, out, full, member)
                            : launch(lookup_kernel<2, 3, false, false, kLookupMinBlocks>, g32, stream, ctx, ix, queries, n, rcf, out, full, member);
    }
    const int grid = grid_for(n, ctx.sm_count, 2 * (ix.canonical && ix.kmer_words == 2 ? kLookupMinBlocksWideCanon : kLookupMinBlocks));
    const int mode = member ? 2 : (full ? 1 : 0);
    const int crc = check_rc ? 1 : 0;
    cudaError_t err = cudaSuccess;
#define SSHASH_LAUNCH(W, MODE, ASCII) \
    err = ix.canonical ? launch(lookup_kernel<W, MODE, ASCII, true, (W == 2 ? kLookupMinBlocksWideCanon : kLookupMinBlocks)>, grid, stream, ctx, ix, queries, n, crc, ids, full, member) \
                       : launch(lookup_kernel<W, MODE, ASCII, false, kLookupMinBlocks>, grid, stream, ctx, ix, queries, n, crc, ids, full, member)
#define SSHASH_LAUNCH_WIDE(MODE, ASCII) \
    err = ix.canonical ? launch(lookup_kernel<1, MODE, ASCII, true, kLookupMinBlocks, true>, grid, stream, ctx, ix, queries, n, crc, ids, full, member) \
                       : launch(lookup_kernel<1, MODE, ASCII, false, kLookupMinBlocks, true>, grid, stream, ctx, ix, queries, n, crc, ids, full, member)
    if (ix.wide && ix.kmer_words == 1 && mode != 1) {      // ids / membership over wide entries
        if (mode == 0) { if (ascii) SSHASH_LAUNCH_WIDE(0, true); else SSHASH_LAUNCH_WIDE(0, false); }
        else { if (ascii) SSHASH_LAUNCH_WIDE(2, true); else SSHASH_LAUNCH_WIDE(2, false); }
        return err;
    }
#undef SSHASH_LAUNCH_WIDE
#define SSHASH_DISPATCH_MODE(W, ASCII)                     \
    do {                                                   \
        if (mode == 0) SSHASH_LAUNCH(W, 0, ASCII);         \
        else if (mode == 1) SSHASH_LAUNCH(W, 1, ASCII);    \
        else SSHASH_LAUNCH(W, 2, ASCII);                   \
    } while (0)
    if (ix.kmer_words == 1) { if (ascii) SSHASH_DISPATCH_MODE(1, true); else SSHASH_DISPATCH_MODE(1, false); }
    else { if (ascii) SSHASH_DISPATCH_MODE(2, true); else SSHASH_DISPATCH_MODE(2, false); }
#undef SSHASH_DISPATCH_MODE
#undef SSHASH_LAUNCH
    return err;
}

cudaError_t launch_access(const DeviceIndex& ix, const LaunchCtx& ctx, const uint64_t* ids, uint64_t n, uint64_t* kmers_out,
                          cudaStream_t stream) {
    if (n == 0) return cudaSuccess;
    const int grid = grid_for(n, ctx.sm_count, 8);
    if (ix.kmer_words == 1) return launch(access_kernel<1>, grid, stream, ctx, ix, ids, n, kmers_out);
    return launch(access_kernel<2>, grid, stream, ctx, ix, ids, n, kmers_out);
}

cudaError_t launch_minimizer_partition(const DeviceIndex& ix, const LaunchCtx& ctx, const uint64_t* kmers, uint64_t n, uint32_t* out,
                                       cudaStream_t stream) {
    if (n == 0) return cudaSuccess;
    const int grid = grid_for(n, ctx.sm_count, 8);
    if (ix.kmer_words == 1) return launch(minimizer_partition_kernel<1>, grid, stream, ctx, ix, kmers, n, out);
    return launch(minimizer_partition_kernel<2>, grid, stream, ctx, ix, kmers, n, out);
}

cudaError_t launch_weight(const DeviceIndex& ix, const LaunchCtx& ctx, const uint64_t* ids, uint64_t n, uint64_t* weights_out,
                          cudaStream_t stream) {
    if (n == 0) return cudaSuccess;
    return launch(weight_kernel, grid_for(n, ctx.sm_count, 8), stream, ctx, ix, ids, n, weights_out);
}

cudaError_t launch_neighbours(const DeviceIndex& ix, const LaunchCtx& ctx, const uint64_t* in, bool strings, uint64_t n,
                              bool check_rc, int which, uint64_t* expanded, uint64_t* ids, sshash_lookup_result* full,
                              cudaStream_t stream) {
    if (n == 0) return cudaSuccess;
    const int grid = grid_for(8 * n, ctx.sm_count, 8);
    cudaError_t e;
    if (ix.kmer_words == 1)
        e = strings ? launch(expand_neighbours_kernel<1, true>, grid, stream, ctx, ix, in, n, expanded)
                    : launch(expand_neighbours_kernel<1, false>, grid, stream, ctx, ix, in, n, expanded);
    else
        e = strings ? launch(expand_neighbours_kernel<2, true>, grid, stream, ctx, ix, in, n, expanded)
                    : launch(expand_neighbours_kernel<2, false>, grid, stream, ctx, ix, in, n, expanded);
    if (e != cudaSuccess) return e;
    e = launch_lookup(ix, ctx, expanded, false, 8 * n, check_rc, ids, full, nullptr, stream);
    if (e != cudaSuccess) return e;
    if ((which & 3) != 3) e = launch(reset_neighbour_slots_kernel, grid, stream, ctx, n, which, ids, full);
    return e;
}

cudaError_t launch_validate_index(const DeviceIndex& ix, const LaunchCtx& ctx, uint32_t* flag, cudaStream_t stream) {
    const uint64_t n = std::max(ix.codewords.size, std::max(ix.mid_load.size, ix.heavy.size));
    if (n == 0) return cudaSuccess;
    return launch(validate_index_kernel, grid_for(n, ctx.sm_count, 8), stream, ctx, ix, flag);
}

cudaError_t launch_build_fingerprints(const DeviceIndex& ix, const LaunchCtx& ctx, uint32_t fp_bits, uint64_t* out,
                                      uint32_t* filter, uint32_t filter_shift, cudaStream_t stream) {
    if (ix.codewords.size == 0) return cudaSuccess;
    return launch(build_fingerprints_kernel, grid_for(ix.codewords.size, ctx.sm_count, 8), stream, ctx, ix, fp_bits,
                  reinterpret_cast<unsigned long long*>(out), filter, filter_shift);
}

cudaError_t launch_build_wide(const DeviceIndex& ix, const LaunchCtx& ctx, void* out, cudaStream_t stream) {
    if (ix.codewords.size == 0) return cudaSuccess;
    return launch(build_wide_kernel, grid_for(ix.codewords.size, ctx.sm_count, 8), stream, ctx, ix, static_cast<ulonglong2*>(out));
}

uint64_t streaming_anchor_bytes(uint64_t num_reads) { return num_reads * kAnchorsPerRead * sizeof(Anchor); }

uint64_t window_offsets_scratch_words(uint64_t num_reads) { return (num_reads + 1 + kScanTile - 1) / kScanTile + 1; }

cudaError_t launch_window_offsets(uint32_t k, const uint64_t* read_begins, const uint64_t* read_ends, uint64_t num_reads,
                                  uint64_t* win_offsets, uint64_t* block_sums, cudaStream_t stream) {
    const uint64_t nblocks = (num_reads + 1 + kScanTile - 1) / kScanTile;
    win_block_sums_kernel<<<(unsigned)nblocks, kBlock, 0, stream>>>(k, read_begins, read_ends, num_reads, block_sums);
    win_scan_sums_kernel<<<1, kBlock, 0, stream>>>(block_sums, nblocks);
    win_offsets_kernel<<<(unsigned)nblocks, kBlock, 0, stream>>>(k, read_begins, read_ends, num_reads, block_sums, win_offsets);
    g_launches.fetch_add(3);
    return cudaGetLastError();
}

cudaError_t launch_distinct_check(const DeviceIndex& ix, const LaunchCtx& ctx, uint32_t* flag, cudaStream_t stream) {
    const int grid = grid_for(ix.strings_bits / 2, ctx.sm_count, 8);
    return ix.kmer_words == 1 ? launch(distinct_check_kernel<1>, grid, stream, ctx, ix, flag)
                              : launch(distinct_check_kernel<2>, grid, stream, ctx, ix, flag);
}

cudaError_t launch_streaming(const DeviceIndex& ix, const LaunchCtx& ctx, const char* bases, const uint64_t* read_begins,
                             const uint64_t* read_ends, const uint64_t* win_offsets, uint64_t num_reads, void* anchors, uint64_t* win_id, uint64_t* win_aux,
                             uint64_t* ids_out, uint64_t total_windows_bound, unsigned long long* counters, cudaStream_t stream, bool replay) {
    if (num_reads == 0) return cudaSuccess;
    if (replay) anchors = nullptr;                       // every window is an independent lookup, then the literal state machine
    // 64-bit k-mers on a distinct index: streamed ids == window ids, so the window kernel writes the caller's buffer directly
    if (ix.kmer_words == 1 && ids_out && !replay) win_id = ids_out;
    Anchor* an = static_cast<Anchor*>(anchors);
    if (an) {
        const int grid = grid_for(num_reads * kAnchorsPerRead, ctx.sm_count, 8);
        cudaError_t e = ix.kmer_words == 1 ? launch(stream_anchor_kernel<1>, grid, stream, ctx, ix, bases, read_begins, read_ends, num_reads, an)
                                           : launch(stream_anchor_kernel<2>, grid, stream, ctx, ix, bases, read_begins, read_ends, num_reads, an);
        if (e != cudaSuccess) return e;
    }
    {
        uint64_t threads = num_reads * 32;
        const int grid = grid_for(threads, ctx.sm_count, 4);
        cudaError_t ez = cudaMemsetAsync(counters + 5, 0, sizeof(unsigned long long), stream);   // work-claim counter
        if (ez != cudaSuccess) return ez;
        cudaError_t e;
        if (ix.kmer_words == 1)
            e = launch(stream_windows_kernel<1>, grid, stream, ctx, ix, bases, read_begins, read_ends, win_offsets, num_reads,
                       (const Anchor*)an, win_id, win_aux, counters + 5);
        else
            e = launch(stream_windows_kernel<2>, grid, stream, ctx, ix, bases, read_begins, read_ends, win_offsets, num_reads,
                       (const Anchor*)an, win_id, win_aux, counters + 5);
        if (e != cudaSuccess) return e;
    }
    if (ix.kmer_words == 1 && replay)
        return launch(stream_scan_kernel<1>, grid_for(num_reads, ctx.sm_count, 8), stream, ctx, ix, bases, read_begins, read_ends,
                      win_offsets, num_reads, win_id, win_aux, ids_out, counters);
    if (ix.kmer_words == 1) {
        // 64-bit k-mers: the windows kernel wrote the streamed ids straight into ids_out (see below)
        return launch(stream_classify_kernel, grid_for(total_windows_bound, ctx.sm_count, 8), stream, ctx, win_offsets, num_reads,
                      (const uint64_t*)win_id, (const uint64_t*)win_aux, counters);
    }
    return launch(stream_scan_kernel<2>, grid_for(num_reads, ctx.sm_count, 8), stream, ctx, ix, bases, read_begins, read_ends,
                  win_offsets, num_reads, win_id, win_aux, ids_out, counters);
}

uint64_t parse_tiles(uint64_t n_bytes) { return (n_bytes + kParseTile - 1) / kParseTile; }

cudaError_t launch_count_lines(const uint8_t* raw, uint64_t n_bytes, uint64_t* tile_counts, cudaStream_t stream) {
    const uint64_t tiles = parse_tiles(n_bytes);
    newline_count_kernel<<<(unsigned)tiles, kBlock, 0, stream>>>(raw, n_bytes, tile_counts);
    // scan tiles + 1 entries (the caller zeroes the last one): entry [tiles] becomes the number of lines
    win_scan_sums_kernel<<<1, kBlock, 0, stream>>>(tile_counts, tiles + 1);
    g_launches.fetch_add(2);
    return cudaGetLastError();
}

cudaError_t launch_read_spans(const uint8_t* raw, uint64_t n_bytes, const uint64_t* tile_offsets, uint64_t* line_start,
                              uint64_t num_records, uint32_t lines_per_record, uint64_t* read_begins, uint64_t* read_ends,
                              int sm_count, cudaStream_t stream) {
    newline_positions_kernel<<<(unsigned)parse_tiles(n_bytes), kBlock, 0, stream>>>(raw, n_bytes, tile_offsets, line_start);
    read_spans_kernel<<<grid_for(num_records, sm_count, 8), kBlock, 0, stream>>>(line_start, num_records, lines_per_record,
                                                                                   read_begins, read_ends);
    g_launches.fetch_add(2);
    return cudaGetLastError();
}

}  // namespace sshash_b200
