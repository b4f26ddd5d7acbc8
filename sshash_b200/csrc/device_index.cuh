// device_index.cuh -- the HBM-resident mirror of an SSHash index and the device-side lookup.
//
// Layout in HBM (one cudaMalloc per array, 256-byte aligned, each padded with >= 2 zero words so
// the two-word funnel reads below never leave the allocation):
//   strings        u64[]  2 bits/base, all input strings concatenated (+ the reference's sentinel)
//                         -- verbatim copy of spectrum_preserving_string_set::strings
//   pilots         u64[]  every PTHash partition's `compact` pilot vector, packed back to back
//                         (word-aligned), minimizer MPHF first, then the <= 8 skew-index MPHFs
//   free_slots     u32[]  every partition's Elias-Fano free-slot list DECODED to plain integers
//                         (replaces elias_fano::access + darray select: one 4-byte load)
//   phf_parts      DevPhfPart[] per-partition constants + offsets into the two pools above
//   codewords      u64[]  control_codewords compact vector, RE-ENCODED at open time: entry =
//                         codeword | fingerprint(minimizer owning the slot) << cw_code_bits.  A
//                         minimizer that is NOT in the index hashes to an arbitrary slot, and all but
//                         2^-cw_fp_bits of those are rejected right after this read, before the cold
//                         strings read the reference needs to find out (verbatim when cw_fp_bits == 0)
//   mid_load       u64[]  mid_load_buckets compact vector, verbatim
//   heavy          u64[]  heavy_load_buckets compact vector, verbatim
//   skew_pos[i]    u64[]  positions[i] compact vectors, verbatim
//   ends           u64[]  string end-points DECODED from the endpoints_sequence (n = strings+1)
//   ends_dir       u32[]  ends_dir[h] = index of the last end-point < (h << dir_shift) (0 if none):
//                         replaces hints_0 + the unary scan of high_bits with one 4-byte load + a
//                         short linear scan of `ends`
//   weight_starts  u64[]  (weighted indexes) first k-mer id of every weight interval, DECODED from
//                         the Elias-Fano weights::m_weight_interval_lengths (+ a sampled directory
//                         weight_dir like ends_dir); weight_values / weight_dict: verbatim
// Everything else (k, m, seeds, widths, begin_buckets_of_size[65]) travels in the kernel
// parameter block (`DeviceIndex`, < 1 KB, constant-cached).
//
// Reference semantics (file:line under /root/reference) are cited per function.
#pragma once

#include <cstdint>

namespace sshash_b200 {

struct DevCompact {            // bits::compact_vector read side, compact_vector.hpp:253-260
    const uint64_t* data;
    uint64_t size;
    uint64_t mask;
    uint32_t width;
    uint32_t pad_;
};

struct DevPhfPart {            // one pthash::single_phf (single_phf.hpp:116-142) + its offset: 32 bytes, two 128-bit loads
    uint64_t offset;           // partitioned_phf::partition::offset, partitioned_phf.hpp:20-38
    uint32_t pilots_word;      // first word of this partition's pilots inside the pilots pool
    uint32_t free_off;         // first free slot of this partition inside the free-slot pool
    uint32_t num_keys;         // the three sizes are < 2^32 per partition (checked at open time), which
    uint32_t table_size;       // turns the 64x64-bit high multiplies of the bucketer / of fastmod-free
    uint32_t num_buckets;      // `position` into 64x32-bit ones
    uint32_t pilot_width;
};

struct DevPhf {                // pthash::partitioned_phf, partitioned_phf.hpp:139-149
    uint64_t seed_hi;          // ~seed  (CityHash seed pair {seed, ~seed}, hash_util.hpp:12-16)
    uint64_t city_a;           // ShiftMix(seed * k1) * k1      -- seed-only terms of CityMurmur
    uint64_t city_cb;          // (~seed) * k1                     hoisted to the host
    uint64_t num_partitions;
    const DevPhfPart* parts;
    uint64_t first_part_;      // host-side scratch while the partition table is being placed
};

struct DeviceIndex {
    uint32_t k, m;
    uint32_t canonical;
    uint32_t kmer_words;       // 1 (max_k 31 build) or 2 (max_k 63 build)
    uint64_t magic;            // mixer_64::m_magic
    uint64_t num_kmers;
    uint64_t num_strings;
    const uint64_t* strings;
    uint64_t strings_bits;
    const uint64_t* pilots;
    const uint32_t* free_slots;
    DevPhf mphf;
    DevCompact codewords;
    DevCompact mid_load;
    DevCompact heavy;
    uint32_t n_skew;
    uint32_t dir_shift;
    DevPhf skew[8];
    DevCompact skew_pos[8];
    const uint64_t* ends;
    uint64_t n_ends;
    const uint32_t* ends_dir;
    const uint32_t* ends32;            // the same end-points as u32 when the last one is < 2^32 - 1 (else nullptr):
                                       // locate_string reads these, half the bytes to keep L2-resident
    uint32_t pad3_[2];
    // weights (include/weights.hpp:148-153,182-187); n_weight_intervals == 0 <=> not weighted
    const uint64_t* weight_starts;     // n_weight_intervals + 1 entries (+ sentinels)
    const uint32_t* weight_dir;        // weight_dir[h] = index of the last start <= (h << weight_dir_shift)
    uint64_t n_weight_intervals;
    DevCompact weight_values;          // m_weight_interval_values: interval -> id in the dictionary
    DevCompact weight_dict;            // m_weight_dictionary: id -> weight
    uint32_t weight_dir_shift;
    uint32_t begin_buckets_of_size[65];
    uint32_t cw_code_bits;     // width of the reference's codeword inside a `codewords` entry
    uint32_t cw_fp_bits;       // fingerprint bits above it (0 = verbatim vector)
    uint32_t pad_;
    // per-position constants of the minimizer scan (compute_minimizer_fast): (magic >> 32) with the
    // low 6 bits replaced by J / 63 - J, J = position inside a 16-position word.  Indexed with
    // compile-time J, so each is a constant-bank operand of the LOP3 that builds the key.
    uint32_t mini_left[16], mini_right[16];
    uint64_t kmer_mask_lo, kmer_mask_hi;   // low 2k bits set (second word: bits 64..2k-1)
    uint64_t mmer_mask;                    // low 2m bits set
    // MINIMIZER FILTER (small indexes only, nullptr otherwise): a blocked Bloom filter over the
    // minimizers of the index, 16 bits per minimizer, two bits per key inside one 32-bit word, built at
    // open time next to the fingerprints.  The STREAMING window lookups probe it right after the
    // minimizer: a pass whose minimizer is not in the index ends there, before CityHash + PTHash +
    // the codeword read (+23 % windows/s on cfg3, where half of the reads are negative and whole
    // tiles of windows fail together).  lookup_kernel does not use it: measured -3 % on positives
    // (the probe) and +1 % on uniform random negatives -- ~15 % of random k-mers have a minimizer
    // that IS in the index (both are biased to small hashes), so a warp almost never fails as a whole.
    // Only kept when it is small enough to stay L2-resident with the rest.
    const uint32_t* minimizer_filter;
    uint32_t filter_shift;                 // word index = hash >> filter_shift
    uint32_t pad2_;
    // WIDE ENTRIES (opt-in, SSHASH_GPU_WIDE=1; k <= 31 in the 64-bit build; nullptr otherwise): one
    // 16-byte record per MPHF slot = codeword | TEXT << cw_code_bits, where TEXT is the 2(2k - m) bits
    // of `strings` around the slot's bucket offset, [offset - (k - m), offset + k): every k-mer that
    // has this minimizer at any position lies inside it.  For a SINGLETON bucket (94-97 % of the
    // minimizers) the ids-only lookup compares against this copy instead of reading `strings`: one
    // cold fetch per pass instead of two (codeword + strings).  Built on the GPU at open time from
    // the arrays above; other bucket types fall back to the regular route.
    const ulonglong2* wide;
};

#ifdef __CUDACC__

// ------------------------------------------------------------------------------------------------
// k-mer word types.  W = number of 64-bit words (1: k <= 31, 2: k <= 63).
// ------------------------------------------------------------------------------------------------
template <int W> struct Kmer;
template <> struct Kmer<1> { uint64_t lo; };
template <> struct Kmer<2> { uint64_t lo, hi; };

__device__ __forceinline__ bool kmer_eq(Kmer<1> a, Kmer<1> b) { return a.lo == b.lo; }
__device__ __forceinline__ bool kmer_eq(Kmer<2> a, Kmer<2> b) { return a.lo == b.lo && a.hi == b.hi; }
__device__ __forceinline__ bool kmer_lt(Kmer<1> a, Kmer<1> b) { return a.lo < b.lo; }
__device__ __forceinline__ bool kmer_lt(Kmer<2> a, Kmer<2> b) { return a.hi < b.hi || (a.hi == b.hi && a.lo < b.lo); }

__device__ __forceinline__ uint64_t low_mask(uint32_t bits) {  // bits in [0,64]
    return bits >= 64 ? ~0ull : ((1ull << bits) - 1ull);
}

// ------------------------------------------------------------------------------------------------
// Loads with an L2 eviction policy.  The index arrays fall into two classes:
//   HOT  (pilots, free slots, end-points + directory: tens of MB, touched by every lookup) are
//        loaded evict_last so that they stay resident in the 126 MB L2;
//   COLD (control codewords, strings, bucket arrays: up to GBs, one random sector per lookup)
//        are loaded evict_first so that they do not push the hot arrays out.
// createpolicy with constant operands folds into a uniform descriptor register (no per-thread cost).
// ------------------------------------------------------------------------------------------------
// POLICY: 0 = COLD, 1 = HOT (the historical bool spelling still works: false / true), 2 = NORMAL: no hint, full
// 128-byte fills -- used by the partition-major path for pilots and codewords, which it visits one
// partition at a time so that they are L2 hits by schedule (binned.cu).
constexpr int kCold = 0, kHot = 1, kNormal = 2, kHot64 = 3;   // kHot64: evict_last with 64-byte fills
template <int POLICY>
__device__ __forceinline__ uint64_t l2_policy() {
    uint64_t p;
    if (POLICY == kHot || POLICY == kHot64) asm("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    else asm("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
// COLD loads additionally carry .L2::64B: measured on B200 (tools/micro/gather.cu) a default load
// that misses L2 pulls 128 B from HBM, with .L2::64B it pulls 64 B -- the minimum -- which halves
// the DRAM traffic of the random codeword / strings accesses.
template <int POLICY>
__device__ __forceinline__ uint64_t ld64(const uint64_t* __restrict__ p) {
    uint64_t v;
    if (POLICY == kHot) asm("ld.global.nc.L2::cache_hint.b64 %0, [%1], %2;" : "=l"(v) : "l"(p), "l"(l2_policy<kHot>()));
    else if (POLICY == kNormal) asm("ld.global.nc.b64 %0, [%1];" : "=l"(v) : "l"(p));
    else if (POLICY == kHot64) asm("ld.global.nc.L2::cache_hint.L2::64B.b64 %0, [%1], %2;" : "=l"(v) : "l"(p), "l"(l2_policy<kHot64>()));
    else asm("ld.global.nc.L2::cache_hint.L2::64B.b64 %0, [%1], %2;" : "=l"(v) : "l"(p), "l"(l2_policy<kCold>()));
    return v;
}
template <int POLICY>
__device__ __forceinline__ uint32_t ld32(const uint32_t* __restrict__ p) {
    uint32_t v;
    if (POLICY == kHot) asm("ld.global.nc.L2::cache_hint.b32 %0, [%1], %2;" : "=r"(v) : "l"(p), "l"(l2_policy<kHot>()));
    else if (POLICY == kNormal) asm("ld.global.nc.b32 %0, [%1];" : "=r"(v) : "l"(p));
    else asm("ld.global.nc.L2::cache_hint.L2::64B.b32 %0, [%1], %2;" : "=r"(v) : "l"(p), "l"(l2_policy<kCold>()));
    return v;
}

__device__ __forceinline__ ulonglong2 ld128_cold(const ulonglong2* __restrict__ p) {
    ulonglong2 v;
    asm("ld.global.nc.L2::cache_hint.L2::64B.v2.b64 {%0, %1}, [%2], %3;" : "=l"(v.x), "=l"(v.y) : "l"(p), "l"(l2_policy<kCold>()));
    return v;
}

// dna_uint_kmer_t::crc64, include/kmer.hpp:141-157: complement (xor 0b10 per base), then reverse
// the order of the 32 two-bit groups.  __brevll reverses single bits, so the two bits inside each
// group are swapped back afterwards.
__device__ __forceinline__ uint64_t rc64(uint64_t x) {
    uint64_t r = __brevll(x ^ 0xaaaaaaaaaaaaaaaaull);
    return ((r & 0x5555555555555555ull) << 1) | ((r >> 1) & 0x5555555555555555ull);
}
// reverse_complement_inplace, include/kmer.hpp:159-165
__device__ __forceinline__ Kmer<1> kmer_rc(Kmer<1> x, uint32_t k) { return {rc64(x.lo) >> (64 - 2 * k)}; }
__device__ __forceinline__ Kmer<2> kmer_rc(Kmer<2> x, uint32_t k) {
    uint64_t hi = rc64(x.lo), lo = rc64(x.hi);  // halves swap
    uint32_t s = 128 - 2 * k;                   // in [2, 126]
    Kmer<2> r;
    if (s >= 64) { r.lo = hi >> (s - 64); r.hi = 0; }
    else { r.lo = (lo >> s) | (hi << (64 - s)); r.hi = hi >> s; }
    return r;
}
// the reverse complement of an m-mer held in 64 bits (spss.hpp:95-98 does it through the k-mer type)
__device__ __forceinline__ uint64_t mmer_rc(uint64_t x, uint32_t m) { return rc64(x) >> (64 - 2 * m); }

// 64 bits of `data` starting at bit `pos` (bit_vector::get_word64, bit_vector.hpp:186-193; the
// device copy is zero-padded, which is what the reference's bounds test amounts to)
__device__ __forceinline__ uint64_t read_word64(const uint64_t* __restrict__ data, uint64_t pos) {
    uint64_t w = pos >> 6;
    uint32_t s = (uint32_t)pos & 63u;
    uint64_t a = ld64<false>(data + w);
    if (s == 0) return a;
    uint64_t b = ld64<false>(data + w + 1);
    return (a >> s) | (b << (64 - s));
}

// util::read_kmer_at(strings, k, 2*offset), include/util.hpp:248-257
__device__ __forceinline__ Kmer<1> read_kmer(const DeviceIndex& ix, uint64_t base_offset, uint32_t k, Kmer<1>*) {
    return {read_word64(ix.strings, 2 * base_offset) & low_mask(2 * k)};
}
// the same for k == ix.k, with the mask precomputed on the host
__device__ __forceinline__ Kmer<1> read_kmer(const DeviceIndex& ix, uint64_t base_offset, Kmer<1>*) {
    return {read_word64(ix.strings, 2 * base_offset) & ix.kmer_mask_lo};
}
__device__ __forceinline__ Kmer<2> read_kmer(const DeviceIndex& ix, uint64_t base_offset, Kmer<2>*) {
    uint64_t pos = 2 * base_offset, w = pos >> 6;
    uint32_t s = (uint32_t)pos & 63u;
    uint64_t a = ld64<false>(ix.strings + w), b = ld64<false>(ix.strings + w + 1);
    Kmer<2> r;
    if (s == 0) { r.lo = a; r.hi = b; }
    else {
        uint64_t c = ld64<false>(ix.strings + w + 2);
        r.lo = (a >> s) | (b << (64 - s));
        r.hi = (b >> s) | (c << (64 - s));
    }
    r.lo &= ix.kmer_mask_lo; r.hi &= ix.kmer_mask_hi;
    return r;
}
__device__ __forceinline__ Kmer<2> read_kmer(const DeviceIndex& ix, uint64_t base_offset, uint32_t k, Kmer<2>*) {
    uint64_t pos = 2 * base_offset, w = pos >> 6;
    uint32_t s = (uint32_t)pos & 63u;
    uint64_t a = ld64<false>(ix.strings + w), b = ld64<false>(ix.strings + w + 1);
    Kmer<2> r;
    if (s == 0) { r.lo = a; r.hi = b; }
    else {
        uint64_t c = ld64<false>(ix.strings + w + 2);
        r.lo = (a >> s) | (b << (64 - s));
        r.hi = (b >> s) | (c << (64 - s));
    }
    if (2 * k <= 64) { r.lo &= low_mask(2 * k); r.hi = 0; }
    else r.hi &= low_mask(2 * k - 64);
    return r;
}
__device__ __forceinline__ uint64_t read_mmer(const DeviceIndex& ix, uint64_t base_offset, uint32_t m) {
    (void)m;
    return read_word64(ix.strings, 2 * base_offset) & ix.mmer_mask;
}

// compact_vector::access, compact_vector.hpp:253-260 (two aligned word loads instead of one
// unaligned 8-byte load)
template <int HOT>
__device__ __forceinline__ uint64_t compact_get(const uint64_t* __restrict__ data, uint32_t width, uint64_t mask, uint64_t i) {
    uint64_t pos = i * width, w = pos >> 6;
    uint32_t s = (uint32_t)pos & 63u;
    uint64_t v = ld64<HOT>(data + w) >> s;
    if (s + width > 64) v |= ld64<HOT>(data + w + 1) << (64 - s);
    return v & mask;
}
template <int HOT>
__device__ __forceinline__ uint64_t compact_get(const DevCompact& c, uint64_t i) {
    return compact_get<HOT>(c.data, c.width, c.mask, i);
}

// ------------------------------------------------------------------------------------------------
// minimizer: util::compute_minimizer, include/util.hpp:262-283 with mixer_64::hash,
// include/hash_util.hpp:91.  Leftmost strict minimum.
// ------------------------------------------------------------------------------------------------
struct Minimizer { uint64_t value; uint32_t pos; };

#define SSHASH_MIX_C 0x517cc1b727220a95ull

// Only the running minimum hash and its position are tracked inside the loop; the minimizer value
// is re-extracted from the k-mer afterwards (two selects fewer per m-mer).  When m <= 16 the m-mer
// fits 32 bits and the 64-bit multiply by the mixer constant needs two IMADs instead of four.
// If no hash is < UINT64_MAX the reference leaves minimizer = all ones, pos = 0 (util.hpp:268-270).
template <bool SMALL_M>
__device__ __noinline__ Minimizer compute_minimizer_exact(Kmer<1> x, uint32_t k, uint32_t m, uint64_t magic) {
    uint64_t min_hash = ~0ull, v = x.lo;
    uint32_t pos = 0;
    const uint32_t n = k - m + 1;
    if (SMALL_M) {
        const uint32_t mm = (uint32_t)low_mask(2 * m);
#pragma unroll 4
        for (uint32_t i = 0; i < n; ++i) {
            uint64_t h = ((uint64_t)((uint32_t)v & mm) * SSHASH_MIX_C) ^ magic;
            if (h < min_hash) { min_hash = h; pos = i; }
            v >>= 2;
        }
    } else {
        const uint64_t mm = low_mask(2 * m);
#pragma unroll 4
        for (uint32_t i = 0; i < n; ++i) {
            uint64_t h = ((v & mm) * SSHASH_MIX_C) ^ magic;
            if (h < min_hash) { min_hash = h; pos = i; }
            v >>= 2;
        }
    }
    uint64_t mini = (x.lo >> (2 * pos)) & low_mask(2 * m);
    if (min_hash == ~0ull) mini = ~0ull;
    return {mini, pos};
}
template <bool SMALL_M>
__device__ __noinline__ Minimizer compute_minimizer_exact(Kmer<2> x, uint32_t k, uint32_t m, uint64_t magic) {
    uint64_t min_hash = ~0ull, lo = x.lo, hi = x.hi;
    uint32_t pos = 0;
    const uint32_t n = k - m + 1;
    const uint64_t mm = low_mask(2 * m);
#pragma unroll 4
    for (uint32_t i = 0; i < n; ++i) {
        uint64_t h = SMALL_M ? (((uint64_t)((uint32_t)lo & (uint32_t)mm) * SSHASH_MIX_C) ^ magic)
                             : (((lo & mm) * SSHASH_MIX_C) ^ magic);
        if (h < min_hash) { min_hash = h; pos = i; }
        lo = (lo >> 2) | (hi << 62);
        hi >>= 2;
    }
    const uint32_t s = 2 * pos;   // < 128
    uint64_t w = s == 0 ? x.lo : (s < 64 ? ((x.lo >> s) | (x.hi << (64 - s))) : (x.hi >> (s - 64)));
    uint64_t mini = w & mm;
    if (min_hash == ~0ull) mini = ~0ull;
    return {mini, pos};
}

// FAST PATH.  x -> (x * C) mod 2^64 is a bijection (C is odd), so two m-mers have equal hashes only
// if they are equal; the scan therefore orders the m-mers by the HIGH 26 bits of their hashes only
// (high half of the product, one xor) and proves afterwards that this was enough:
//     key_left (p) = hash26(p) << 6 | p            min -> leftmost  position among the smallest hash26
//     key_right(p) = hash26(p) << 6 | (63 - p)     min -> rightmost position among the smallest hash26
// If both minima name the same position, exactly one m-mer has the smallest 26 high bits, hence
// the smallest 64-bit hash: it is the minimizer.  Otherwise (a repeated m-mer, or two m-mers whose
// hashes agree on 26 bits: 2^-26 per pair) the exact 64-bit scan above decides.
// Cost per m-mer: funnel shift + mask + 2 IMAD (hash) + xor/mask + 2 adds + 2 min, positions and
// shift amounts are immediates: the scan is unrolled 16 positions (= one 32-bit word of text) at a
// time and the n mod 16 positions of the tail are entered through a switch.
__device__ __forceinline__ void kmer_words32(Kmer<1> x, uint32_t (&r)[4]) {
    r[0] = (uint32_t)x.lo; r[1] = (uint32_t)(x.lo >> 32); r[2] = 0; r[3] = 0;
}
__device__ __forceinline__ void kmer_words32(Kmer<2> x, uint32_t (&r)[4]) {
    r[0] = (uint32_t)x.lo; r[1] = (uint32_t)(x.lo >> 32); r[2] = (uint32_t)x.hi; r[3] = (uint32_t)(x.hi >> 32);
}
__device__ __forceinline__ uint64_t kmer_bits_at(Kmer<1> x, uint32_t s) { return x.lo >> s; }
__device__ __forceinline__ uint64_t kmer_bits_at(Kmer<2> x, uint32_t s) {   // s < 128
    return s == 0 ? x.lo : (s < 64 ? ((x.lo >> s) | (x.hi << (64 - s))) : (x.hi >> (s - 64)));
}

template <bool SMALL_M, int J>
__device__ __forceinline__ void minimizer_step(const DeviceIndex& ix, const uint32_t (&r)[4], uint32_t mm,
                                               uint32_t& left, uint32_t& right) {
    constexpr uint32_t c_lo = (uint32_t)SSHASH_MIX_C, c_hi = (uint32_t)(SSHASH_MIX_C >> 32);
    uint32_t hh;
    if (SMALL_M) {
        const uint32_t w = __funnelshift_r(r[0], r[1], 2 * J) & mm;
        hh = __umulhi(w, c_lo) + w * c_hi;
    } else {
        const uint32_t wl = __funnelshift_r(r[0], r[1], 2 * J);
        const uint32_t wh = __funnelshift_r(r[1], r[2], 2 * J) & mm;
        hh = __umulhi(wl, c_lo) + wl * c_hi + wh * c_lo;
    }
    // ((hh ^ magic_hi) & ~63) | J  ==  (hh & ~63) ^ mini_left[J]
    left = min(left, (hh & ~63u) ^ ix.mini_left[J]);
    right = min(right, (hh & ~63u) ^ ix.mini_right[J]);
}

template <int W, bool SMALL_M>
__device__ __forceinline__ Minimizer compute_minimizer_fast(const DeviceIndex& ix, Kmer<W> x) {
    const uint32_t k = ix.k, m = ix.m;
    const uint32_t n = k - m + 1;                      // <= 63
    const uint32_t mm = SMALL_M ? (uint32_t)ix.mmer_mask : (uint32_t)(ix.mmer_mask >> 32);
    uint32_t r[4];
    kmer_words32(x, r);
    uint32_t left = 0xffffffffu, right = 0xffffffffu;   // keys over global positions
    uint32_t base = 0;
#define SSHASH_STEP(J) minimizer_step<SMALL_M, J>(ix, r, mm, l, rt)
    for (; base + 16 <= n; base += 16) {
        uint32_t l = 0xffffffffu, rt = 0xffffffffu;    // keys over positions within this word
        SSHASH_STEP(0); SSHASH_STEP(1); SSHASH_STEP(2); SSHASH_STEP(3); SSHASH_STEP(4); SSHASH_STEP(5);
        SSHASH_STEP(6); SSHASH_STEP(7); SSHASH_STEP(8); SSHASH_STEP(9); SSHASH_STEP(10); SSHASH_STEP(11);
        SSHASH_STEP(12); SSHASH_STEP(13); SSHASH_STEP(14); SSHASH_STEP(15);
        left = min(left, l + base);
        right = min(right, rt - base);
        r[0] = r[1]; r[1] = r[2]; r[2] = r[3]; r[3] = 0;
    }
    if (n & 15) {
        uint32_t l = 0xffffffffu, rt = 0xffffffffu;
        switch (n & 15) {                              // falls through: positions (n & 15) - 1 .. 0
            case 15: SSHASH_STEP(14); case 14: SSHASH_STEP(13); case 13: SSHASH_STEP(12); case 12: SSHASH_STEP(11);
            case 11: SSHASH_STEP(10); case 10: SSHASH_STEP(9); case 9: SSHASH_STEP(8); case 8: SSHASH_STEP(7);
            case 7: SSHASH_STEP(6); case 6: SSHASH_STEP(5); case 5: SSHASH_STEP(4); case 4: SSHASH_STEP(3);
            case 3: SSHASH_STEP(2); case 2: SSHASH_STEP(1); default: SSHASH_STEP(0);
        }
        left = min(left, l + base);
        right = min(right, rt - base);
    }
#undef SSHASH_STEP
    const uint32_t pos = left & 63u;
    if (pos != 63u - (right & 63u)) return compute_minimizer_exact<SMALL_M>(x, k, m, ix.magic);
    return {kmer_bits_at(x, 2 * pos) & ix.mmer_mask, pos};
}

template <int W>
__device__ __forceinline__ Minimizer compute_minimizer(const DeviceIndex& ix, Kmer<W> x) {
#ifdef SSHASH_EXACT_MINIMIZER_ONLY   // A/B switch for measurements
    return ix.m <= 16 ? compute_minimizer_exact<true>(x, ix.k, ix.m, ix.magic) : compute_minimizer_exact<false>(x, ix.k, ix.m, ix.magic);
#else
    return ix.m <= 16 ? compute_minimizer_fast<W, true>(ix, x) : compute_minimizer_fast<W, false>(ix, x);
#endif
}

// ------------------------------------------------------------------------------------------------
// PTHash evaluation.  CityHash128WithSeed restricted to 8- and 16-byte keys:
// external/cityhash/cityhash.cpp:238-266 (CityMurmur, len <= 16 branch), :116-125 (HashLen0to16),
// cityhash.hpp:90-99 (Hash128to64).
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint64_t city_h16(uint64_t u, uint64_t v) {
    const uint64_t kMul = 0x9ddfea08eb382d69ull;
    uint64_t a = (u ^ v) * kMul; a ^= (a >> 47);
    uint64_t b = (v ^ a) * kMul; b ^= (b >> 47);
    return b * kMul;
}
struct Hash128 { uint64_t first, second; };

__device__ __forceinline__ Hash128 city_tail(const DevPhf& f, uint64_t h, uint64_t key_lo) {
    uint64_t a = f.city_a;
    uint64_t c = f.city_cb + h;
    uint64_t d = a + key_lo; d ^= (d >> 47);
    a = city_h16(a, c);
    uint64_t b = city_h16(d, f.seed_hi);
    return {a ^ b, city_h16(b, a)};
}
__device__ __forceinline__ Hash128 city_hash_u64(const DevPhf& f, uint64_t key) {          // len = 8
    return city_tail(f, city_h16(8 + ((key & 0xffffffffull) << 3), key >> 32), key);
}
__device__ __forceinline__ Hash128 city_hash_u128(const DevPhf& f, uint64_t lo, uint64_t hi) {  // len = 16
    uint64_t t = hi + 16;
    uint64_t rot = (t >> 16) | (t << 48);
    return city_tail(f, city_h16(lo, rot) ^ hi, lo);
}

// partitioned_phf::position (partitioned_phf.hpp:145-149) -> single_phf::position
// (single_phf.hpp:68-78) with opt_bucketer::bucket (utils/bucketers.hpp:38-39), range_bucketer
// (:129-131), compact pilots (utils/encoders.hpp:33-35), mix (utils/hasher.hpp:41-43) and the
// minimal remap through the (decoded) free slots.
// (x * y) >> 64 for a 32-bit y
__device__ __forceinline__ uint32_t mulhi_64x32(uint64_t x, uint32_t y) {
    const uint64_t lo = (uint64_t)(uint32_t)x * y;
    return (uint32_t)(((x >> 32) * y + (lo >> 32)) >> 32);
}

// partitioned_phf::position's partition choice, partitioned_phf.hpp:145-149
__device__ __forceinline__ uint32_t mphf_partition(const DevPhf& f, Hash128 h) {
    return f.num_partitions > 1 ? (uint32_t)((((h.first ^ h.second) >> 32) * f.num_partitions) >> 32) : 0u;
}

template <bool BINNED = false>
__device__ __forceinline__ uint64_t phf_position(const DeviceIndex& ix, const DevPhf& f, Hash128 h) {
    const uint64_t part = mphf_partition(f, h);
    const uint4* __restrict__ rec = reinterpret_cast<const uint4*>(f.parts + part);
    const uint4 a = __ldg(rec), b = __ldg(rec + 1);
    const uint64_t offset = ((uint64_t)a.y << 32) | a.x;
    const uint32_t pilots_word = a.z, free_off = a.w, num_keys = b.x, table_size = b.y, num_buckets = b.z, pilot_width = b.w;
    const uint64_t h1 = h.first;
    const uint64_t H = __umul64hi(__umul64hi(h1, h1), (h1 >> 1) | (1ull << 63)) / 8 * 7 + h1 / 8;
    const uint32_t bucket = mulhi_64x32(H, num_buckets);
    // pilots: evict_last in the direct kernels (a partially resident pool beats the cold policy even when it
    // outgrows L2: measured), plain loads in the partition-major path (L2 hits by schedule)
    const uint64_t pilot = compact_get<BINNED ? kNormal : kHot>(ix.pilots + pilots_word, pilot_width, low_mask(pilot_width), bucket);
    uint32_t pos = mulhi_64x32((h.second ^ (pilot * SSHASH_MIX_C)) * SSHASH_MIX_C, table_size);
    if (pos >= num_keys) pos = ld32<true>(ix.free_slots + free_off + (pos - num_keys));
    return offset + pos;
}

// ------------------------------------------------------------------------------------------------
// string end-points: decoded_offsets::offset_to_id (include/offsets.hpp:138-154) ->
// endpoints_sequence::locate (endpoints_sequence.hpp:182-198).  Returns the index of the largest
// end-point <= x; begin/end are that end-point and the next.
// ------------------------------------------------------------------------------------------------
// texts of 2^32 - 1 bases or more keep 64-bit end-points: out of line, so that the common path stays small
static __device__ __noinline__ uint64_t locate_string_u64(const DeviceIndex& ix, uint64_t x, uint64_t& begin, uint64_t& end) {
    uint64_t i = ld32<true>(ix.ends_dir + (x >> ix.dir_shift));
    uint64_t cur = ld64<true>(ix.ends + i), next = ld64<true>(ix.ends + i + 1);
    while (next <= x) { cur = next; ++i; next = ld64<true>(ix.ends + i + 1); }
    begin = cur; end = next;
    return i;
}
__device__ __forceinline__ uint64_t locate_string(const DeviceIndex& ix, uint64_t x, uint64_t& begin, uint64_t& end) {
    if (!ix.ends32) return locate_string_u64(ix, x, begin, end);
    // ends_dir[h] = index of the last end-point < (h << dir_shift) (0 if none): ends[i] <= x holds;
    // advance to the last end-point <= x (about half a step on average: two directory blocks per string)
    const uint32_t x32 = (uint32_t)x;                       // the text has < 2^32 - 1 bases here
    uint32_t i = ld32<true>(ix.ends_dir + (x32 >> ix.dir_shift));
    uint32_t cur = ld32<true>(ix.ends32 + i), next = ld32<true>(ix.ends32 + i + 1);
    while (next <= x32) { cur = next; ++i; next = ld32<true>(ix.ends32 + i + 1); }
    begin = cur; end = next;
    return i;
}
// ------------------------------------------------------------------------------------------------
// one lookup
// ------------------------------------------------------------------------------------------------
struct LookupResult {          // include/util.hpp:38-62
    uint64_t kmer_id, kmer_id_in_string, kmer_offset;
    int64_t kmer_orientation;
    uint64_t string_id, string_begin, string_end, minimizer_found;
};

__device__ __forceinline__ void result_clear(LookupResult& r, bool minimizer_found) {
    r.kmer_id = r.kmer_id_in_string = r.kmer_offset = ~0ull;
    r.kmer_orientation = 1;
    r.string_id = r.string_begin = r.string_end = ~0ull;
    r.minimizer_found = minimizer_found ? 1 : 0;
}

// kmers_city_hasher_128::hash, include/hash_util.hpp:59-67: the key is sizeof(x.bits) = 8 or 16 bytes
__device__ __forceinline__ Hash128 skew_hash(const DevPhf& f, Kmer<1> x) { return city_hash_u64(f, x.lo); }
__device__ __forceinline__ Hash128 skew_hash(const DevPhf& f, Kmer<2> x) { return city_hash_u128(f, x.lo, x.hi); }

// Fingerprint of a minimizer (cw_fp_bits <= 32 bits).  In a canonical index the text at a bucket
// offset may hold the reverse complement of the minimizer (compute_minimizer_tuples.cpp:76-86), so
// the fingerprint is taken of the smaller of the two forms there.
// 32-bit key of a minimizer for the fingerprint and the filter: folded halves of its canonical form
__device__ __forceinline__ uint32_t minimizer_key32(const DeviceIndex& ix, uint64_t minimizer) {
    if (ix.canonical) { uint64_t r = mmer_rc(minimizer, ix.m); minimizer = r < minimizer ? r : minimizer; }
    return (uint32_t)minimizer ^ (uint32_t)(minimizer >> 32);
}
__device__ __forceinline__ uint32_t fingerprint_of_key(const DeviceIndex& ix, uint32_t key32) {
    return (key32 * 0x9e3779b1u) >> (32 - ix.cw_fp_bits);
}
__device__ __forceinline__ uint32_t minimizer_fingerprint(const DeviceIndex& ix, uint64_t minimizer) {
    return fingerprint_of_key(ix, minimizer_key32(ix, minimizer));
}
// filter slot of a key: word index and the two bits inside the word
__device__ __forceinline__ void filter_slot(uint32_t key32, uint32_t shift, uint32_t& word, uint32_t& mask) {
    word = (key32 * 0x9e3779b1u) >> shift;
    const uint32_t g = (key32 ^ (key32 >> 15)) * 0x85ebca6bu;
    mask = (1u << (g >> 27)) | (1u << ((g >> 22) & 31u));
}

// sparse_and_skew_index::lookup (sparse_and_skew_index.hpp:112-137): minimizer -> bucket.
// Returns the number of offsets in the bucket; `first` = the single offset (SINGLETON/HEAVYLOAD)
// or the index of the first entry in mid_load_buckets (MIDLOAD).
// USE_FP: reject a minimizer whose slot belongs to a different minimizer (returns 0 = no bucket).
// Only the ids-only paths may use it: a full lookup_result needs the bucket type of the (wrong)
// slot to reproduce minimizer_found (spss.hpp:51-65).
template <int W, bool USE_FP, bool USE_FILTER = false, bool BINNED = false>
__device__ __forceinline__ uint32_t bucket_of(const DeviceIndex& ix, uint64_t minimizer, Kmer<W> skew_key,
                                              uint64_t& first, bool& heavy) {
    uint32_t fp = 0;
    if (USE_FP) {
        const uint32_t key32 = minimizer_key32(ix, minimizer);
        if (USE_FILTER && ix.minimizer_filter) {
            uint32_t word, mask;
            filter_slot(key32, ix.filter_shift, word, mask);
            if ((ld32<true>(ix.minimizer_filter + word) & mask) != mask) return 0;   // not a minimizer of this index
        }
        fp = fingerprint_of_key(ix, key32);                   // before the loads: only 32 bits stay live across them
    }
    uint64_t id = phf_position<BINNED>(ix, ix.mphf, city_hash_u64(ix.mphf, minimizer));   // minimizers_control_map.hpp:36-39
    // entries are exactly 32 bits wide whenever the reference's codeword has <= 24 bits (api.cu)
    constexpr int CW = BINNED ? kNormal : kCold;
    uint64_t code = ix.codewords.width == 32 ? (uint64_t)ld32<CW>(reinterpret_cast<const uint32_t*>(ix.codewords.data) + id)
                                             : compact_get<CW>(ix.codewords, id);
    heavy = false;
    if (ix.cw_fp_bits) {
        if (USE_FP && (uint32_t)(code >> ix.cw_code_bits) != fp) return 0;
        code &= low_mask(ix.cw_code_bits);
    }
    if ((code & 1) == 0) { first = code >> 1; return 1; }                         // SINGLETON
    if ((code & 3) == 1) {                                                        // MIDLOAD
        code >>= 2;
        uint32_t size = (uint32_t)(code & 63) + 2;
        first = ix.begin_buckets_of_size[size] + (code >> 6) * size;
        return size;
    }
    // HEAVYLOAD: skew_index::lookup, sparse_and_skew_index.hpp:34-44
    code >>= 2;
    uint32_t part = (uint32_t)code & 7;
    uint64_t begin = code >> 3;
    uint64_t kid = phf_position(ix, ix.skew[part], skew_hash(ix.skew[part], skew_key));
    uint64_t pos_in_bucket = compact_get<false>(ix.skew_pos[part], kid);
    uint64_t idx = begin + pos_in_bucket;
    // A k-mer that was never a key gets an arbitrary slot; the reference then reads past the bucket
    // (spss.hpp:51-63).  Whatever is read cannot make the k-mer comparison succeed for an absent
    // k-mer, so clamping yields the same answer without the out-of-bounds access.
    if (idx >= ix.heavy.size) idx = ix.heavy.size - 1;
    first = compact_get<false>(ix.heavy, idx);
    heavy = true;
    return 1;
}

// Regular pass: dictionary::lookup_regular (src/dictionary.cpp:7-22) + spss::lookup_regular
// (spectrum_preserving_string_set.hpp:29-73, _lookup_regular :213-235).
// FULL = also produce minimizer_found exactly (needs the m-mer check of spss.hpp:46-65); without
// it the k-mer comparison alone decides, which yields the same ids (a k-mer match implies the
// m-mer match because the minimizer is a substring of the k-mer at pos_in_kmer).
template <int W, bool FULL, bool FILTER = false, bool BINNED = false>
__device__ __forceinline__ bool lookup_regular_with(const DeviceIndex& ix, Kmer<W> x, Minimizer mi, LookupResult& res) {
    const uint32_t k = ix.k, m = ix.m;
    uint64_t first; bool heavy;
    uint32_t n = bucket_of<W, !FULL, FILTER, BINNED>(ix, mi.value, x, first, heavy);
    if (!FULL && n == 0) { result_clear(res, false); return false; }
    uint64_t off0 = (n == 1) ? first : compact_get<false>(ix.mid_load, first);
    if (FULL) {
        if (read_mmer(ix, off0, m) != mi.value) { result_clear(res, heavy); return false; }
    }
    // The candidate scan only COMPARES; the string is located once, after the scan, so that the
    // lanes of a warp (most have a single candidate, a few a long mid-load bucket) reconverge
    // before the end-point loads instead of each running them inside its own divergent iteration.
    // (Issuing the end-point probe of the first candidate speculatively, next to its k-mer read,
    // was measured too: +-0.5 % on a 5e8-k-mer index, -2.5 % on an L2-resident one: not kept.)
    for (uint32_t i = 0;; ++i) {
        uint64_t ko = 0;
        bool hit = false;
        for (; i < n; ++i) {
            uint64_t off = (i == 0) ? off0 : compact_get<false>(ix.mid_load, first + i);
            if (off < mi.pos) continue;
            ko = off - mi.pos;
            if (kmer_eq(read_kmer(ix, ko, (Kmer<W>*)nullptr), x)) { hit = true; break; }
        }
        if (!hit) break;
        uint64_t sb, se;
        uint64_t sid = locate_string(ix, ko, sb, se);
        if (ko < se - k + 1) {                          // spss.hpp:233: reject k-mers spanning two strings
            res.kmer_id = ko - sid * (k - 1);
            res.kmer_id_in_string = ko - sb;
            res.kmer_offset = ko;
            res.kmer_orientation = 1;
            res.string_id = sid; res.string_begin = sb; res.string_end = se;
            res.minimizer_found = 1;
            return true;
        }
    }
    result_clear(res, true);
    return false;
}

template <int W, bool FULL>
__device__ __forceinline__ bool lookup_regular(const DeviceIndex& ix, Kmer<W> x, LookupResult& res) {
    return lookup_regular_with<W, FULL>(ix, x, compute_minimizer(ix, x), res);
}

// Canonical pass: dictionary::lookup_canonical(kmer, kmer_rc, mini_info) (src/dictionary.cpp:44-56)
// + spss::lookup_canonical (spss.hpp:75-112, _lookup_canonical :237-247, __lookup_canonical :249-275)
template <int W, bool FULL, bool FILTER = false, bool BINNED = false>
__device__ __forceinline__ bool lookup_canonical_with(const DeviceIndex& ix, Kmer<W> x, Kmer<W> xr, Minimizer mi,
                                                      LookupResult& res) {
    const uint32_t k = ix.k, m = ix.m;
    Kmer<W> canon = kmer_lt(x, xr) ? x : xr;            // std::min(uint_kmer, uint_kmer_rc), dictionary.cpp:53
    uint64_t first; bool heavy;
    uint32_t n = bucket_of<W, !FULL, FILTER, BINNED>(ix, mi.value, canon, first, heavy);
    if (!FULL && n == 0) { result_clear(res, false); return false; }
    uint64_t off0 = (n == 1) ? first : compact_get<false>(ix.mid_load, first);
    if (FULL) {
        uint64_t rm = read_mmer(ix, off0, m);
        if (rm != mi.value && rm != mmer_rc(mi.value, m)) { result_clear(res, heavy); return false; }
    }
    // Candidates in the reference's order: for each offset the forward position pa, then the mirrored
    // one pb (spss.hpp:249-275).  The two k-mers of an offset lie within k - m bases of each other
    // (almost always the same 64-byte sector), so both are read up front -- independent loads, one
    // latency -- and compared in order.  Compare-only scan, located after reconvergence; (i, t0)
    // = where to resume should a match be rejected by the string-boundary test (spss.hpp:266).
    const uint32_t pa = mi.pos, pb = k - m - mi.pos;
    uint32_t i = 0, t0 = 0;
    for (;;) {
        uint64_t ko = 0;
        bool hit = false, eq_r = false, second = false;
        for (; i < n; ++i, t0 = 0) {
            const uint64_t off = (i == 0) ? off0 : compact_get<false>(ix.mid_load, first + i);
            const bool va = t0 == 0 && off >= pa, vb = off >= pb;
            Kmer<W> ra = x, rb = x;
            if (va) ra = read_kmer(ix, off - pa, (Kmer<W>*)nullptr);
            if (vb) rb = read_kmer(ix, off - pb, (Kmer<W>*)nullptr);
            if (va && (kmer_eq(ra, x) || kmer_eq(ra, xr))) { ko = off - pa; eq_r = kmer_eq(ra, xr); hit = true; break; }
            if (vb && (kmer_eq(rb, x) || kmer_eq(rb, xr))) { ko = off - pb; eq_r = kmer_eq(rb, xr); hit = true; second = true; break; }
        }
        if (!hit) break;
        uint64_t sb, se;
        uint64_t sid = locate_string(ix, ko, sb, se);
        if (ko < se - k + 1) {
            res.kmer_id = ko - sid * (k - 1);
            res.kmer_id_in_string = ko - sb;
            res.kmer_offset = ko;
            res.kmer_orientation = eq_r ? -1 : 1;       // spss.hpp:263-264 (rc wins when both equal: impossible for odd k)
            res.string_id = sid; res.string_begin = sb; res.string_end = se;
            res.minimizer_found = 1;
            return true;
        }
        if (second) { ++i; t0 = 0; } else t0 = 1;
    }
    result_clear(res, true);
    return false;
}

// dictionary::lookup_canonical(Kmer), src/dictionary.cpp:24-42
template <int W, bool FULL>
__device__ __forceinline__ bool lookup_canonical(const DeviceIndex& ix, Kmer<W> x, LookupResult& res) {
    const uint32_t k = ix.k;
    Kmer<W> xr = kmer_rc(x, k);
    Minimizer mf = compute_minimizer(ix, x);
    Minimizer mr = compute_minimizer(ix, xr);
    // the smaller minimizer decides; on a tie the forward info is tried first, then the rc info (:35-41)
    const bool tie = mf.value == mr.value;
    Minimizer mi = mr.value < mf.value ? mr : mf;
#pragma unroll 1
    for (int t = 0;; ++t) {                             // one inlined copy of the pass
        if (lookup_canonical_with<W, FULL>(ix, x, xr, mi, res)) return true;
        if (!tie || t == 1) return false;
        mi = mr;
    }
}

// ------------------------------------------------------------------------------------------------
// WIDE-ENTRY passes (ids-only, 64-bit k-mers).  Same decisions as lookup_regular_with /
// lookup_canonical_with; for a SINGLETON bucket the candidate k-mers come out of the entry's copy of
// the text instead of `strings`.  ko = offset - pos sits at bit 2 * ((k - m) - pos) of TEXT.
// ------------------------------------------------------------------------------------------------
struct WideEntry { uint64_t code, text_lo, text_hi; };
__device__ __forceinline__ WideEntry load_wide(const DeviceIndex& ix, uint64_t minimizer) {
    const uint64_t id = phf_position(ix, ix.mphf, city_hash_u64(ix.mphf, minimizer));
    const ulonglong2 e = ld128_cold(ix.wide + id);
    const uint32_t w = ix.cw_code_bits;                  // in [1, 63]
    return {e.x & low_mask(w), (e.x >> w) | (e.y << (64 - w)), e.y >> w};
}
__device__ __forceinline__ uint64_t wide_kmer(const DeviceIndex& ix, const WideEntry& e, uint32_t pos) {
    const uint32_t s = 2 * (ix.k - ix.m - pos);          // <= 2 (k - m) <= 60
    return (s == 0 ? e.text_lo : (e.text_lo >> s) | (e.text_hi << (64 - s))) & ix.kmer_mask_lo;
}
__device__ __forceinline__ bool wide_finish(const DeviceIndex& ix, uint64_t ko, int64_t orientation, LookupResult& res) {
    const uint32_t k = ix.k;
    uint64_t sb, se;
    const uint64_t sid = locate_string(ix, ko, sb, se);
    if (!(ko < se - k + 1)) return false;                // spss.hpp:233 / :266
    res.kmer_id = ko - sid * (k - 1);
    res.kmer_id_in_string = ko - sb;
    res.kmer_offset = ko;
    res.kmer_orientation = orientation;
    res.string_id = sid; res.string_begin = sb; res.string_end = se;
    res.minimizer_found = 1;
    return true;
}

__device__ __forceinline__ bool lookup_regular_wide(const DeviceIndex& ix, Kmer<1> x, LookupResult& res) {
    const Minimizer mi = compute_minimizer(ix, x);
    const WideEntry e = load_wide(ix, mi.value);
    if (e.code & 1) return lookup_regular_with<1, false>(ix, x, mi, res);     // MIDLOAD / HEAVYLOAD: the regular route
    const uint64_t off = e.code >> 1;
    if (off >= mi.pos && wide_kmer(ix, e, mi.pos) == x.lo && wide_finish(ix, off - mi.pos, 1, res)) return true;
    result_clear(res, true);
    return false;
}

__device__ __forceinline__ bool lookup_canonical_wide_with(const DeviceIndex& ix, Kmer<1> x, Kmer<1> xr, Minimizer mi,
                                                           LookupResult& res) {
    const WideEntry e = load_wide(ix, mi.value);
    if (e.code & 1) return lookup_canonical_with<1, false>(ix, x, xr, mi, res);
    const uint64_t off = e.code >> 1;
    const uint32_t pa = mi.pos, pb = ix.k - ix.m - mi.pos;                      // spss.hpp:243-246: pa first, then pb
    if (off >= pa) {
        const uint64_t r = wide_kmer(ix, e, pa);
        if ((r == x.lo || r == xr.lo) && wide_finish(ix, off - pa, r == xr.lo ? -1 : 1, res)) return true;
    }
    if (off >= pb) {
        const uint64_t r = wide_kmer(ix, e, pb);
        if ((r == x.lo || r == xr.lo) && wide_finish(ix, off - pb, r == xr.lo ? -1 : 1, res)) return true;
    }
    result_clear(res, true);
    return false;
}

// dictionary::lookup_canonical(Kmer), src/dictionary.cpp:24-42, over wide entries
__device__ __forceinline__ bool lookup_canonical_wide(const DeviceIndex& ix, Kmer<1> x, LookupResult& res) {
    const Kmer<1> xr = kmer_rc(x, ix.k);
    const Minimizer mf = compute_minimizer(ix, x);
    const Minimizer mr = compute_minimizer(ix, xr);
    const bool tie = mf.value == mr.value;
    Minimizer mi = mr.value < mf.value ? mr : mf;
#pragma unroll 1
    for (int t = 0;; ++t) {
        if (lookup_canonical_wide_with(ix, x, xr, mi, res)) return true;
        if (!tie || t == 1) return false;
        mi = mr;
    }
}

// dictionary::lookup(Kmer, check_reverse_complement), src/dictionary.cpp:64-78
template <int W, bool FULL>
__device__ __forceinline__ void lookup_kmer(const DeviceIndex& ix, Kmer<W> x, bool check_rc, LookupResult& res) {
    if (ix.canonical) { lookup_canonical<W, FULL>(ix, x, res); return; }
    if (lookup_regular<W, FULL>(ix, x, res)) return;
    if (check_rc) {
        lookup_regular<W, FULL>(ix, kmer_rc(x, ix.k), res);
        res.kmer_orientation = -1;                      // dictionary.cpp:74-75: also for a miss
    }
}

#endif  // __CUDACC__

}  // namespace sshash_b200
