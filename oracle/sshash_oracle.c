/*
 * oracle/sshash_oracle.c -- TEST INFRASTRUCTURE ONLY.  NOT part of the product.
 *
 * A plain-C, CPU, single-file restatement of the reference's (jermp/sshash @ afff26dc, index
 * format 5.1.1) k-mer Lookup / streaming-membership path, written from the reference's
 * algorithm; every function cites the reference file:line it follows (paths relative to
 * /root/reference).  It exists so that the CUDA path can be checked bit-for-bit on a box where
 * the reference itself is absent.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load it; sshash_b200/ never does.
 *
 * Parity status: PINNED.  tests/test_oracle.py checks this file against (a) the literal
 * known-answer vectors of SURVEY.md section 8c, (b) golden vectors produced by the unmodified
 * reference compiled here (oracle/_ref, see tests/golden/make_golden.py), on regular / canonical /
 * k=63 / heavy-bucket / multi-partition indexes, and (c) the reference's README Example-2/3
 * streaming report.
 *
 * Deliberately simple: it mirrors the reference's data structures (Elias-Fano end-points with
 * hints, darray select, compact vectors) instead of the flattened device mirrors the CUDA path
 * uses, so that the two implementations are independent.
 */
#define _GNU_SOURCE
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

typedef unsigned __int128 u128;
#define INVALID UINT64_MAX

/* ------------------------------------------------------------------------------------------
 * containers (external/pthash/external/bits/include)
 * ---------------------------------------------------------------------------------------- */
typedef struct { uint64_t size, width, mask; uint64_t* data; uint64_t nwords; } cvec_t;   /* compact_vector.hpp:286-304 */
typedef struct { uint64_t num_bits; uint64_t* data; uint64_t nwords; } bitvec_t;          /* bit_vector.hpp:343-352 */
typedef struct {                                                                           /* darray.hpp:210-227 */
    uint64_t positions;
    int64_t* block_inventory; uint64_t n_blocks;
    uint16_t* subblock_inventory; uint64_t n_subblocks;
    uint64_t* overflow_positions; uint64_t n_overflow;
} darray_t;
typedef struct { uint64_t back; bitvec_t high; darray_t d1, d0; cvec_t low; } ef_t;        /* elias_fano.hpp:366-379 */
typedef struct {                                                                           /* endpoints_sequence.hpp:222-240 */
    uint64_t back; bitvec_t high; darray_t d1; cvec_t hints0; uint8_t* low; uint64_t n;
} endpoints_t;
typedef struct {                                                                           /* single_phf.hpp:116-142 */
    uint64_t seed, num_keys, table_size, num_buckets; cvec_t pilots; ef_t free_slots;
} sphf_t;
typedef struct {                                                                           /* partitioned_phf.hpp:193-206 */
    uint64_t seed, num_keys, table_size, partitioner_buckets, nparts; uint64_t* offsets; sphf_t* parts;
} pphf_t;

typedef struct {
    /* dictionary.hpp:139-162 */
    uint8_t vx, vy, vz;
    uint64_t num_kmers, num_strings, k, m, canonical, magic;
    /* spectrum_preserving_string_set.hpp:200-211 */
    endpoints_t ends;
    bitvec_t strings;
    /* sparse_and_skew_index.hpp:149-167, minimizers_control_map.hpp:49-63 */
    pphf_t mphf;
    cvec_t codewords;
    uint32_t* begin_buckets_of_size; uint64_t n_bbos;
    cvec_t mid_load;
    /* skew_index, sparse_and_skew_index.hpp:60-76 */
    uint64_t n_ski; pphf_t* ski_mphfs; uint64_t n_pos; cvec_t* positions; cvec_t heavy;
    uint64_t weights_bytes;
    /* weights.hpp:182-187 */
    cvec_t w_values; ef_t w_lengths; cvec_t w_dict;
    int kmer_bits;          /* 64 or 128: which reference build wrote the file */
    uint8_t* file; uint64_t file_size;
} oracle_dict;

typedef struct {
    uint64_t kmer_id, kmer_id_in_string, kmer_offset; int64_t kmer_orientation;
    uint64_t string_id, string_begin, string_end, minimizer_found;
} oracle_result; /* util.hpp:38-62 */

typedef struct {
    uint64_t num_kmers, num_positive_kmers, num_negative_kmers, num_invalid_kmers, num_searches,
        num_extensions;
} oracle_report; /* util.hpp:21-36 */

/* ------------------------------------------------------------------------------------------
 * file parsing: essentials visitor format (essentials.hpp:329-407): PODs raw, vectors as
 * {u64 n; n*sizeof(T)}; no padding.  Every array is copied to its own aligned allocation with
 * two spare words so the 8-byte unaligned reads of compact_vector::access stay in bounds.
 * ---------------------------------------------------------------------------------------- */
typedef struct { const uint8_t* p; const uint8_t* end; int fail; } rd_t;

static uint64_t rd_u64(rd_t* r) { uint64_t v = 0; if (r->p + 8 > r->end) { r->fail = 1; return 0; } memcpy(&v, r->p, 8); r->p += 8; return v; }
static uint16_t rd_u16(rd_t* r) { uint16_t v = 0; if (r->p + 2 > r->end) { r->fail = 1; return 0; } memcpy(&v, r->p, 2); r->p += 2; return v; }
static uint8_t rd_u8(rd_t* r) { if (r->p + 1 > r->end) { r->fail = 1; return 0; } return *r->p++; }
static void* rd_vec(rd_t* r, uint64_t elem, uint64_t* n_out) {
    uint64_t n = rd_u64(r);
    *n_out = n;
    if (r->fail || n > (uint64_t)(r->end - r->p) / elem) { r->fail = 1; *n_out = 0; return calloc(2, 8); }
    uint8_t* a = (uint8_t*)calloc(n * elem + 16, 1);
    memcpy(a, r->p, n * elem);
    r->p += n * elem;
    return a;
}
static void rd_cvec(rd_t* r, cvec_t* c) { c->size = rd_u64(r); c->width = rd_u64(r); c->mask = rd_u64(r); c->data = (uint64_t*)rd_vec(r, 8, &c->nwords); }
static void rd_bitvec(rd_t* r, bitvec_t* b) { b->num_bits = rd_u64(r); b->data = (uint64_t*)rd_vec(r, 8, &b->nwords); }
static void rd_darray(rd_t* r, darray_t* d) {
    d->positions = rd_u64(r);
    d->block_inventory = (int64_t*)rd_vec(r, 8, &d->n_blocks);
    d->subblock_inventory = (uint16_t*)rd_vec(r, 2, &d->n_subblocks);
    d->overflow_positions = (uint64_t*)rd_vec(r, 8, &d->n_overflow);
}
static void rd_ef(rd_t* r, ef_t* e) { e->back = rd_u64(r); rd_bitvec(r, &e->high); rd_darray(r, &e->d1); rd_darray(r, &e->d0); rd_cvec(r, &e->low); }
static void rd_sphf(rd_t* r, sphf_t* f) {
    f->seed = rd_u64(r); f->num_keys = rd_u64(r); f->table_size = rd_u64(r); f->num_buckets = rd_u64(r);
    rd_cvec(r, &f->pilots); rd_ef(r, &f->free_slots);
}
static void rd_pphf(rd_t* r, pphf_t* f) {
    f->seed = rd_u64(r); f->num_keys = rd_u64(r); f->table_size = rd_u64(r); f->partitioner_buckets = rd_u64(r);
    f->nparts = rd_u64(r);
    if (r->fail || f->nparts > (1u << 24)) { r->fail = 1; f->nparts = 0; }
    f->offsets = (uint64_t*)calloc(f->nparts + 1, 8);
    f->parts = (sphf_t*)calloc(f->nparts + 1, sizeof(sphf_t));
    for (uint64_t i = 0; i != f->nparts; ++i) { f->offsets[i] = rd_u64(r); rd_sphf(r, &f->parts[i]); }
}

static void free_cvec(cvec_t* c) { free(c->data); }
static void free_darray(darray_t* d) { free(d->block_inventory); free(d->subblock_inventory); free(d->overflow_positions); }
static void free_pphf(pphf_t* f) {
    for (uint64_t i = 0; i != f->nparts; ++i) {
        free_cvec(&f->parts[i].pilots);
        free(f->parts[i].free_slots.high.data); free_darray(&f->parts[i].free_slots.d1);
        free_darray(&f->parts[i].free_slots.d0); free_cvec(&f->parts[i].free_slots.low);
    }
    free(f->offsets); free(f->parts);
}

void oracle_close(oracle_dict* d) {
    if (!d) return;
    free(d->ends.high.data); free_darray(&d->ends.d1); free_cvec(&d->ends.hints0); free(d->ends.low);
    free(d->strings.data);
    free_pphf(&d->mphf); free_cvec(&d->codewords); free(d->begin_buckets_of_size); free_cvec(&d->mid_load);
    for (uint64_t i = 0; i != d->n_ski; ++i) free_pphf(&d->ski_mphfs[i]);
    free(d->ski_mphfs);
    for (uint64_t i = 0; i != d->n_pos; ++i) free_cvec(&d->positions[i]);
    free(d->positions); free_cvec(&d->heavy);
    free_cvec(&d->w_values); free(d->w_lengths.high.data); free_darray(&d->w_lengths.d1); free_darray(&d->w_lengths.d0);
    free_cvec(&d->w_lengths.low); free_cvec(&d->w_dict);
    free(d);
}

/* max_k: 31 or 63 = which reference build wrote the file (not recorded in it);
   0 = infer (k <= 31 -> default 64-bit build). */
oracle_dict* oracle_open(const char* path, int max_k, char* err, uint64_t errlen) {
    FILE* f = fopen(path, "rb");
    if (!f) { snprintf(err, errlen, "cannot open '%s'", path); return NULL; }
    fseek(f, 0, SEEK_END); long sz = ftell(f); fseek(f, 0, SEEK_SET);
    uint8_t* buf = (uint8_t*)malloc(sz > 0 ? sz : 1);
    if (fread(buf, 1, sz, f) != (size_t)sz) { fclose(f); free(buf); snprintf(err, errlen, "short read"); return NULL; }
    fclose(f);
    oracle_dict* d = (oracle_dict*)calloc(1, sizeof(oracle_dict));
    rd_t r = {buf, buf + sz, 0};
    d->vx = rd_u8(&r); d->vy = rd_u8(&r); d->vz = rd_u8(&r);           /* essentials.hpp:791-803 */
    if (d->vx != 5) {                                                    /* util.hpp:191-195 */
        snprintf(err, errlen, "MAJOR index version mismatch: SSHash index needs rebuilding");
        free(buf); free(d); return NULL;
    }
    d->num_kmers = rd_u64(&r); d->num_strings = rd_u64(&r);
    d->k = rd_u16(&r); d->m = rd_u16(&r); d->canonical = rd_u8(&r); d->magic = rd_u64(&r);
    uint64_t k2 = rd_u16(&r), m2 = rd_u16(&r);                           /* spss k, m */
    (void)rd_u64(&r);                                                     /* m_num_bits_per_relative_offset: garbage (offsets.hpp:104-112) */
    d->ends.back = rd_u64(&r); rd_bitvec(&r, &d->ends.high); rd_darray(&r, &d->ends.d1);
    rd_cvec(&r, &d->ends.hints0); d->ends.low = (uint8_t*)rd_vec(&r, 1, &d->ends.n);
    rd_bitvec(&r, &d->strings);
    rd_pphf(&r, &d->mphf); rd_cvec(&r, &d->codewords);
    d->begin_buckets_of_size = (uint32_t*)rd_vec(&r, 4, &d->n_bbos);
    rd_cvec(&r, &d->mid_load);
    d->n_ski = rd_u64(&r);
    if (r.fail || d->n_ski > 64) { r.fail = 1; d->n_ski = 0; }
    d->ski_mphfs = (pphf_t*)calloc(d->n_ski + 1, sizeof(pphf_t));
    for (uint64_t i = 0; i != d->n_ski; ++i) rd_pphf(&r, &d->ski_mphfs[i]);
    d->n_pos = rd_u64(&r);
    if (r.fail || d->n_pos > 64) { r.fail = 1; d->n_pos = 0; }
    d->positions = (cvec_t*)calloc(d->n_pos + 1, sizeof(cvec_t));
    for (uint64_t i = 0; i != d->n_pos; ++i) rd_cvec(&r, &d->positions[i]);
    rd_cvec(&r, &d->heavy);
    d->weights_bytes = (uint64_t)(r.end - r.p);
    rd_cvec(&r, &d->w_values); rd_ef(&r, &d->w_lengths); rd_cvec(&r, &d->w_dict);     /* weights.hpp:182-187 */
    d->file_size = sz;
    free(buf);
    if (r.fail || k2 != d->k || m2 != d->m) { snprintf(err, errlen, "malformed index file"); oracle_close(d); return NULL; }
    if (max_k == 0) max_k = d->k <= 31 ? 31 : 63;
    if (d->k > (uint64_t)max_k) { snprintf(err, errlen, "k=%lu needs the max_k=63 build", (unsigned long)d->k); oracle_close(d); return NULL; }
    d->kmer_bits = max_k == 31 ? 64 : 128;
    return d;
}

void oracle_info(const oracle_dict* d, uint64_t* out /* 16 */) {
    out[0] = d->num_kmers; out[1] = d->num_strings; out[2] = d->k; out[3] = d->m; out[4] = d->canonical;
    out[5] = d->magic; out[6] = d->mphf.seed; out[7] = d->mphf.nparts; out[8] = d->mphf.num_keys;
    out[9] = d->n_ski; out[10] = d->heavy.size; out[11] = d->mid_load.size; out[12] = d->codewords.width;
    out[13] = d->strings.num_bits; out[14] = d->weights_bytes; out[15] = (uint64_t)d->kmer_bits;
}

/* ------------------------------------------------------------------------------------------
 * container reads
 * ---------------------------------------------------------------------------------------- */
/* compact_vector::access, compact_vector.hpp:253-260 */
static uint64_t cvec_access(const cvec_t* c, uint64_t i) {
    uint64_t pos = i * c->width, word;
    memcpy(&word, (const char*)c->data + (pos >> 3), 8);
    return (word >> (pos & 7)) & c->mask;
}
/* bit_vector::get_word64, bit_vector.hpp:186-193 */
static uint64_t get_word64(const bitvec_t* b, uint64_t pos) {
    uint64_t block = pos >> 6, shift = pos & 63;
    uint64_t word = b->data[block] >> shift;
    if (shift && block + 1 < b->nwords) word |= b->data[block + 1] << (64 - shift);
    return word;
}
/* select_in_word, bits/include/util.hpp:68-106 (portable form) */
static uint64_t select_in_word(uint64_t word, uint64_t i) {
    for (uint64_t j = 0; j != i; ++j) word &= word - 1;
    return (uint64_t)__builtin_ctzll(word);
}
/* darray::select (darray1: WordGetter = identity), darray.hpp:163-188 */
static uint64_t darray_select(const darray_t* d, const bitvec_t* B, uint64_t i) {
    uint64_t block = i / 1024;
    int64_t block_pos = d->block_inventory[block];
    if (block_pos < 0) {
        uint64_t overflow_pos = (uint64_t)(-block_pos - 1);
        return d->overflow_positions[overflow_pos + (i & 1023)];
    }
    uint64_t subblock = i / 32;
    uint64_t start_pos = (uint64_t)block_pos + d->subblock_inventory[subblock];
    uint64_t reminder = i & 31;
    if (!reminder) return start_pos;
    uint64_t word_idx = start_pos >> 6, word_shift = start_pos & 63;
    uint64_t word = B->data[word_idx] & (UINT64_MAX << word_shift);
    for (;;) {
        uint64_t popcnt = (uint64_t)__builtin_popcountll(word);
        if (reminder < popcnt) break;
        reminder -= popcnt;
        word = B->data[++word_idx];
    }
    return (word_idx << 6) + select_in_word(word, reminder);
}
/* elias_fano<false,false>::access, elias_fano.hpp:181-185 */
static uint64_t ef_access(const ef_t* e, uint64_t i) {
    return ((darray_select(&e->d1, &e->high, i) - i) << e->low.width) | cvec_access(&e->low, i);
}
/* endpoints_sequence::access, endpoints_sequence.hpp:160-163 */
static uint64_t ends_access(const endpoints_t* e, uint64_t i) {
    return ((darray_select(&e->d1, &e->high, i) - i) << 8) | e->low[i];
}
/* position of the next set bit at or after `pos` (bit_vector::iterator::next, bit_vector.hpp:276-295) */
static uint64_t next_one(const bitvec_t* b, uint64_t pos) {
    uint64_t block = pos >> 6;
    uint64_t word = b->data[block] & (UINT64_MAX << (pos & 63));
    while (word == 0) word = b->data[++block];
    return (block << 6) + (uint64_t)__builtin_ctzll(word);
}
/* endpoints_sequence::locate, endpoints_sequence.hpp:182-198 with next_geq_helper :242-277:
   start from the hint, walk the ones of high_bits until value >= x; lo = largest end-point <= x,
   hi = the next one. */
static void ends_locate(const endpoints_t* e, uint64_t x, uint64_t* lo_pos, uint64_t* lo_val, uint64_t* hi_val) {
    uint64_t h_x = x >> 8, p = 0, begin = 0;
    if (h_x > 0) { p = cvec_access(&e->hints0, h_x - 1); begin = p - h_x + 1; }
    uint64_t pos = begin;
    uint64_t hb = next_one(&e->high, p);
    uint64_t val = ((hb - pos) << 8) | e->low[pos];
    uint64_t prev_val = 0; int have_prev = 0;
    while (val < x) {
        prev_val = val; have_prev = 1;
        ++pos;
        hb = next_one(&e->high, hb + 1);
        val = ((hb - pos) << 8) | e->low[pos];
    }
    if (val > x) {                       /* step back one (iterator::prev_value :127-139) */
        *lo_pos = pos - 1;
        *lo_val = have_prev ? prev_val : ends_access(e, pos - 1);
        *hi_val = val;
    } else {
        *lo_pos = pos; *lo_val = val;
        hb = next_one(&e->high, hb + 1);
        *hi_val = ((hb - (pos + 1)) << 8) | e->low[pos + 1];
    }
}

/* ------------------------------------------------------------------------------------------
 * k-mer primitives (include/kmer.hpp, include/util.hpp, include/hash_util.hpp)
 * ---------------------------------------------------------------------------------------- */
static u128 mask_bits(unsigned b) { return b >= 128 ? ~(u128)0 : (((u128)1 << b) - 1); }

/* dna_uint_kmer_t::crc64, kmer.hpp:141-157 (default alphabet A0 C1 T2 G3) */
static uint64_t crc64(uint64_t x) {
    uint64_t c = x ^ 0xaaaaaaaaaaaaaaaaULL;
    uint64_t res = __builtin_bswap64(c);
    const uint64_t c1 = 0x0f0f0f0f0f0f0f0fULL, c2 = 0x3333333333333333ULL;
    res = ((res & c1) << 4) | ((res & (c1 << 4)) >> 4);
    res = ((res & c2) << 2) | ((res & (c2 << 2)) >> 2);
    return res;
}
/* reverse_complement_inplace, kmer.hpp:159-165: pop 64-bit words low first, append crc64 of each
   (so the halves swap), then drop the unused low bits */
static u128 rc_kmer(u128 x, uint64_t k, int kmer_bits) {
    if (kmer_bits == 64) return (u128)(crc64((uint64_t)x) >> (64 - 2 * k));
    u128 rev = ((u128)crc64((uint64_t)x) << 64) | (u128)crc64((uint64_t)(x >> 64));
    return rev >> (128 - 2 * k);
}
/* util::read_kmer_at, util.hpp:248-257 */
static u128 read_kmer_at(const oracle_dict* d, uint64_t k, uint64_t pos) {
    u128 kmer = 0;
    for (int i = d->kmer_bits - 64; i >= 0; i -= 64) {
        if (pos + (uint64_t)i < d->strings.num_bits) {
            uint64_t w = get_word64(&d->strings, pos + (uint64_t)i);
            kmer = d->kmer_bits == 64 ? (u128)w : ((kmer << 64) | (u128)w); /* append64, kmer.hpp:69-76 */
        }
    }
    return kmer & mask_bits((unsigned)(2 * k));
}
/* util::compute_minimizer, util.hpp:262-283; mixer_64::hash, hash_util.hpp:91 */
static void compute_minimizer(const oracle_dict* d, u128 kmer, uint64_t* minimizer, uint64_t* pos_out) {
    uint64_t min_hash = INVALID, mini = INVALID, pos = 0;
    const u128 mm = mask_bits((unsigned)(2 * d->m));
    for (uint64_t i = 0; i != d->k - d->m + 1; ++i) {
        uint64_t mmer = (uint64_t)(kmer & mm);
        uint64_t hash = (mmer * 0x517cc1b727220a95ULL) ^ d->magic;
        if (hash < min_hash) { min_hash = hash; mini = mmer; pos = i; }
        kmer >>= 2;
    }
    *minimizer = mini; *pos_out = pos;
}

/* ------------------------------------------------------------------------------------------
 * CityHash128WithSeed for 8- and 16-byte keys (external/cityhash/cityhash.cpp:238-269, only the
 * len <= 16 branch of CityMurmur is reachable) and PTHash evaluation
 * ---------------------------------------------------------------------------------------- */
static const uint64_t K1 = 0xb492b66fbe98f273ULL;
static uint64_t shift_mix(uint64_t v) { return v ^ (v >> 47); }                       /* cityhash.cpp:111 */
static uint64_t hash128to64(uint64_t lo, uint64_t hi) {                               /* cityhash.hpp:90-99 */
    const uint64_t kMul = 0x9ddfea08eb382d69ULL;
    uint64_t a = (lo ^ hi) * kMul; a ^= (a >> 47);
    uint64_t b = (hi ^ a) * kMul; b ^= (b >> 47); b *= kMul;
    return b;
}
static uint64_t rotr64(uint64_t v, int s) { return (v >> s) | (v << (64 - s)); }     /* cityhash.cpp:107-109 */
/* key = 8 or 16 little-endian bytes given as (lo, hi); seed pair = {seed, ~seed} (hash_util.hpp:12-16,62-66) */
static void city128(uint64_t lo, uint64_t hi, int len, uint64_t seed, uint64_t* first, uint64_t* second) {
    uint64_t a = seed, b = ~seed, c, dd, h;
    if (len == 8) {                                                                   /* HashLen0to16: 4 <= len <= 8, cityhash.cpp:122-125 */
        h = hash128to64(8 + ((lo & 0xffffffffULL) << 3), lo >> 32);
    } else {                                                                          /* len == 16 > 8, cityhash.cpp:117-121 */
        h = hash128to64(lo, rotr64(hi + 16, 16)) ^ hi;
    }
    a = shift_mix(a * K1) * K1;                                                       /* cityhash.cpp:244-247 */
    c = b * K1 + h;
    dd = shift_mix(a + lo);                                                           /* len >= 8: Fetch64(s) */
    a = hash128to64(a, c);                                                            /* cityhash.cpp:263-265 */
    b = hash128to64(dd, b);
    *first = a ^ b;
    *second = hash128to64(b, a);
}
static uint64_t mul_high(uint64_t x, uint64_t y) { return (uint64_t)(((u128)x * (u128)y) >> 64); } /* pthash utils/util.hpp:39-41 */
static uint64_t mix64(uint64_t v) { return v * 0x517cc1b727220a95ULL; }                            /* pthash utils/hasher.hpp:41-43 */

/* partitioned_phf::position partitioned_phf.hpp:145-149 + single_phf::position single_phf.hpp:68-78
   + opt_bucketer::bucket bucketers.hpp:38-39 + range_bucketer::bucket :129-131 */
static uint64_t pphf_eval(const pphf_t* f, uint64_t key_lo, uint64_t key_hi, int key_len) {
    uint64_t h1, h2;
    city128(key_lo, key_hi, key_len, f->seed, &h1, &h2);
    uint64_t mixh = h1 ^ h2;                                                          /* hash128::mix, hasher.hpp:79-81 */
    uint64_t part = ((mixh >> 32) * f->partitioner_buckets) >> 32;
    const sphf_t* p = &f->parts[part];
    uint64_t H = mul_high(mul_high(h1, h1), (h1 >> 1) | (1ULL << 63)) / 8 * 7 + h1 / 8;
    uint64_t bucket = mul_high(H, p->num_buckets);
    uint64_t pilot = cvec_access(&p->pilots, bucket);
    uint64_t pos = mul_high(mix64(h2 ^ mix64(pilot)), p->table_size);
    if (pos >= p->num_keys) pos = ef_access(&p->free_slots, pos - p->num_keys);
    return f->offsets[part] + pos;
}

/* ------------------------------------------------------------------------------------------
 * lookup
 * ---------------------------------------------------------------------------------------- */
static void result_init(oracle_result* r, int minimizer_found) {                     /* util.hpp:39-49 */
    r->kmer_id = r->kmer_id_in_string = r->kmer_offset = INVALID;
    r->kmer_orientation = 1;
    r->string_id = r->string_begin = r->string_end = INVALID;
    r->minimizer_found = (uint64_t)minimizer_found;
}

typedef struct { uint64_t n; uint64_t offs[64]; int heavy; } bucket_t;

/* sparse_and_skew_index::lookup :112-137, skew_index::lookup :34-44, bucket_iterator :82-110 */
static void ssi_lookup(const oracle_dict* d, u128 skew_key, uint64_t minimizer, bucket_t* b) {
    uint64_t mid = pphf_eval(&d->mphf, minimizer, 0, 8);                             /* minimizers_control_map.hpp:36-39 */
    uint64_t code = cvec_access(&d->codewords, mid);
    b->heavy = 0;
    if ((code & 1) == 0) { b->n = 1; b->offs[0] = code >> 1; return; }              /* SINGLETON */
    if ((code & 3) == 1) {                                                           /* MIDLOAD */
        code >>= 2;
        uint64_t size = (code & 63) + 2, id = code >> 6;
        uint64_t begin = d->begin_buckets_of_size[size] + id * size;
        b->n = size;
        for (uint64_t i = 0; i != size; ++i) b->offs[i] = cvec_access(&d->mid_load, begin + i);
        return;
    }
    code >>= 2;                                                                      /* HEAVYLOAD */
    uint64_t part = code & 7, begin = code >> 3;
    uint64_t kid = pphf_eval(&d->ski_mphfs[part], (uint64_t)skew_key, (uint64_t)(skew_key >> 64), d->kmer_bits / 8);
    uint64_t pos_in_bucket = cvec_access(&d->positions[part], kid);
    uint64_t idx = begin + pos_in_bucket;
    /* a k-mer that was never a key can index past the array; the reference reads out of bounds
       there (spectrum_preserving_string_set.hpp:51-63).  The value read can never make the m-mer
       AND the k-mer comparison succeed for an absent k-mer, so any in-range substitute gives the
       same answer: clamp. */
    if (idx >= d->heavy.size) idx = d->heavy.size - 1;
    b->n = 1; b->offs[0] = cvec_access(&d->heavy, idx); b->heavy = 1;
}

/* decoded_offsets::offset_to_id offsets.hpp:138-154 + the acceptance test of
   _lookup_regular spectrum_preserving_string_set.hpp:224-234 */
static int finish_candidate(const oracle_dict* d, uint64_t kmer_offset, oracle_result* res) {
    res->kmer_offset = kmer_offset;
    if (!(res->string_begin != INVALID && kmer_offset >= res->string_begin && kmer_offset < res->string_end - d->k + 1)) {
        uint64_t lo_pos, lo_val, hi_val;
        ends_locate(&d->ends, kmer_offset, &lo_pos, &lo_val, &hi_val);
        res->string_id = lo_pos; res->string_begin = lo_val; res->string_end = hi_val;
    }
    res->kmer_id = kmer_offset - res->string_id * (d->k - 1);
    res->kmer_id_in_string = kmer_offset - res->string_begin;
    return kmer_offset < res->string_end - d->k + 1;
}

/* dictionary::lookup_regular dictionary.cpp:7-22 + spss::lookup_regular spss.hpp:29-73 */
static void lookup_regular(const oracle_dict* d, u128 kmer, oracle_result* out) {
    uint64_t mini, pos; bucket_t b;
    compute_minimizer(d, kmer, &mini, &pos);
    ssi_lookup(d, kmer, mini, &b);
    uint64_t read_mmer = (uint64_t)read_kmer_at(d, d->m, 2 * b.offs[0]);
    if (read_mmer != mini) { result_init(out, b.heavy ? 1 : 0); return; }
    oracle_result res; result_init(&res, 1);
    for (uint64_t i = 0; i != b.n; ++i) {                                            /* _lookup_regular spss.hpp:213-235 */
        if (b.offs[i] < pos) continue;
        uint64_t ko = b.offs[i] - pos;
        res.kmer_offset = ko;
        if (kmer != read_kmer_at(d, d->k, 2 * ko)) continue;
        if (finish_candidate(d, ko, &res)) { *out = res; return; }
    }
    result_init(out, 1);
}

/* dictionary::lookup_canonical(kmer, kmer_rc, mini_info) dictionary.cpp:44-56 + spss::lookup_canonical
   spss.hpp:75-112 + _lookup_canonical/__lookup_canonical :237-275 */
static void lookup_canonical_with(const oracle_dict* d, u128 kmer, u128 kmer_rc, uint64_t mini, uint64_t pos, oracle_result* out) {
    bucket_t b;
    u128 canon = kmer < kmer_rc ? kmer : kmer_rc;
    ssi_lookup(d, canon, mini, &b);
    uint64_t read_mmer = (uint64_t)read_kmer_at(d, d->m, 2 * b.offs[0]);
    if (read_mmer != mini) {
        uint64_t mini_rc = (uint64_t)rc_kmer((u128)mini, d->m, d->kmer_bits);
        if (read_mmer != mini_rc) { result_init(out, b.heavy ? 1 : 0); return; }
    }
    oracle_result res; result_init(&res, 1);
    for (uint64_t i = 0; i != b.n; ++i) {
        for (int t = 0; t != 2; ++t) {
            uint64_t p = t == 0 ? pos : d->k - d->m - pos;
            if (b.offs[i] < p) continue;
            uint64_t ko = b.offs[i] - p;
            res.kmer_offset = ko;
            u128 rk = read_kmer_at(d, d->k, 2 * ko);
            if (rk != kmer && rk != kmer_rc) continue;
            res.kmer_orientation = rk == kmer_rc ? -1 : 1;
            if (finish_candidate(d, ko, &res)) { *out = res; return; }
        }
    }
    result_init(out, 1);
}

/* dictionary::lookup_canonical(kmer) dictionary.cpp:24-42 */
static void lookup_canonical(const oracle_dict* d, u128 kmer, oracle_result* out) {
    u128 kmer_rc = rc_kmer(kmer, d->k, d->kmer_bits);
    uint64_t mini, pos, mini_rc, pos_rc;
    compute_minimizer(d, kmer, &mini, &pos);
    compute_minimizer(d, kmer_rc, &mini_rc, &pos_rc);
    if (mini < mini_rc) { lookup_canonical_with(d, kmer, kmer_rc, mini, pos, out); }
    else if (mini_rc < mini) { lookup_canonical_with(d, kmer, kmer_rc, mini_rc, pos_rc, out); }
    else {
        lookup_canonical_with(d, kmer, kmer_rc, mini, pos, out);
        if (out->kmer_id == INVALID) lookup_canonical_with(d, kmer, kmer_rc, mini_rc, pos_rc, out);
    }
}

/* dictionary::lookup(Kmer, check_rc) dictionary.cpp:64-78 */
static void lookup(const oracle_dict* d, u128 kmer, int check_rc, oracle_result* out) {
    if (d->canonical) { lookup_canonical(d, kmer, out); return; }
    lookup_regular(d, kmer, out);
    if (check_rc && out->kmer_id == INVALID) {
        lookup_regular(d, rc_kmer(kmer, d->k, d->kmer_bits), out);
        out->kmer_orientation = -1;
    }
}

static u128 load_kmer(const oracle_dict* d, const uint64_t* p, uint64_t i) {
    if (d->kmer_bits == 64) return (u128)p[i];
    return ((u128)p[2 * i + 1] << 64) | (u128)p[2 * i];
}

void oracle_lookup_batch(const oracle_dict* d, const uint64_t* kmers, uint64_t n, int check_rc,
                         uint64_t* ids, oracle_result* full) {
    for (uint64_t i = 0; i != n; ++i) {
        oracle_result r;
        lookup(d, load_kmer(d, kmers, i), check_rc, &r);
        if (ids) ids[i] = r.kmer_id;
        if (full) full[i] = r;
    }
}

/* util::string_to_uint_kmer util.hpp:207-213 with char_to_uint kmer.hpp:194 -- no validation */
static u128 string_to_kmer(const char* s, uint64_t k) {
    u128 x = 0;
    for (uint64_t i = 0; i != k; ++i) x |= (u128)(((uint64_t)s[i] >> 1) & 3) << (2 * i);
    return x;
}

void oracle_lookup_batch_ascii(const oracle_dict* d, const char* kmers, uint64_t n, int check_rc,
                               uint64_t* ids, oracle_result* full) {
    for (uint64_t i = 0; i != n; ++i) {
        oracle_result r;
        lookup(d, string_to_kmer(kmers + i * d->k, d->k), check_rc, &r);
        if (ids) ids[i] = r.kmer_id;
        if (full) full[i] = r;
    }
}

/* ------------------------------------------------------------------------------------------
 * navigational queries: dictionary::kmer_forward/backward_neighbours, kmer_neighbours,
 * string_neighbours (src/dictionary.cpp:112-201).  Alphabet order A,C,T,G = codes 0..3
 * (include/kmer.hpp:115-119,194).  Slots that the reference leaves default-constructed keep the
 * default lookup_result (invalid ids, minimizer_found = true, util.hpp:39-49).
 * ---------------------------------------------------------------------------------------- */
static void neighbours_from(const oracle_dict* d, u128 suffix, int do_fwd, u128 prefix, int do_bwd, int check_rc,
                            oracle_result* out /* 8 */) {
    for (int j = 0; j != 8; ++j) result_init(out + j, 1);
    if (do_fwd)                                              /* forward_neighbours, dictionary.cpp:112-120 */
        for (u128 c = 0; c != 4; ++c) lookup(d, suffix | (c << (2 * (d->k - 1))), check_rc, out + (int)c);
    if (do_bwd)                                              /* backward_neighbours, dictionary.cpp:121-129 */
        for (u128 c = 0; c != 4; ++c) lookup(d, prefix | c, check_rc, out + 4 + (int)c);
}

void oracle_kmer_neighbours_batch(const oracle_dict* d, const uint64_t* kmers, uint64_t n, int check_rc, int which,
                                  oracle_result* out) {
    for (uint64_t i = 0; i != n; ++i) {
        u128 x = load_kmer(d, kmers, i);
        u128 suffix = x >> 2;                                               /* get_suffix: drop_char, :139-143 */
        u128 prefix = (x << 2) & mask_bits((unsigned)(2 * d->k));           /* get_prefix: pad_char + take_chars(k), :160-165 */
        neighbours_from(d, suffix, which & 1, prefix, which & 2, check_rc, out + 8 * i);
    }
}

/* string_neighbours dictionary.cpp:189-201 with string_prefix / string_suffix spss.hpp:19-27 */
void oracle_string_neighbours_batch(const oracle_dict* d, const uint64_t* string_ids, uint64_t n, int check_rc,
                                    oracle_result* out) {
    for (uint64_t i = 0; i != n; ++i) {
        uint64_t begin = ends_access(&d->ends, string_ids[i]), end = ends_access(&d->ends, string_ids[i] + 1);
        u128 suffix = read_kmer_at(d, d->k - 1, 2 * (end - d->k + 1));
        u128 prefix = read_kmer_at(d, d->k - 1, 2 * begin) << 2;
        neighbours_from(d, suffix, 1, prefix, 1, check_rc, out + 8 * i);
    }
}

/* ------------------------------------------------------------------------------------------
 * Access: offsets::id_to_offset offsets.hpp:41-65 (binary search restated without the linear
 * tail) + spss::access spss.hpp:114-118
 * ---------------------------------------------------------------------------------------- */
void oracle_access_batch(const oracle_dict* d, const uint64_t* ids, uint64_t n, uint64_t* kmers_out) {
    for (uint64_t q = 0; q != n; ++q) {
        uint64_t id = ids[q], lo = 0, hi = d->ends.n - 1;
        /* largest string s with first-kmer-id(s) = begin(s) - s*(k-1) <= id */
        while (hi - lo > 1) {
            uint64_t mid = lo + (hi - lo) / 2;
            if (ends_access(&d->ends, mid) - mid * (d->k - 1) <= id) lo = mid; else hi = mid;
        }
        uint64_t offset = id + lo * (d->k - 1);
        u128 x = read_kmer_at(d, d->k, 2 * offset);
        if (d->kmer_bits == 64) kmers_out[q] = (uint64_t)x;
        else { kmers_out[2 * q] = (uint64_t)x; kmers_out[2 * q + 1] = (uint64_t)(x >> 64); }
    }
}

/* ------------------------------------------------------------------------------------------
 * Weights: dictionary::weight dictionary.cpp:96-100 -> weights::weight weights.hpp:148-153:
 *   i = interval_lengths.prev_leq(kmer_id).pos; weight = dictionary[ interval_values[i] ]
 * elias_fano::prev_leq (elias_fano.hpp:236-258) = the rightmost element <= x, saturating to the last
 * one for x >= back(); restated as a binary search over elias_fano::access (:181-185) instead of
 * the reference's select0 + forward scan (:385-430) -- same answer for a sorted sequence.
 * ---------------------------------------------------------------------------------------- */
static uint64_t ef_prev_leq_pos(const ef_t* e, uint64_t x) {
    uint64_t n = e->low.size;
    if (x >= e->back) return n - 1;
    if (ef_access(e, 0) > x) return ~0ull;
    uint64_t lo = 0, hi = n - 1;                 /* access(lo) <= x < access(hi) */
    while (hi - lo > 1) {
        uint64_t mid = lo + (hi - lo) / 2;
        if (ef_access(e, mid) <= x) lo = mid; else hi = mid;
    }
    return lo;
}
int oracle_weighted(const oracle_dict* d) { return d->w_dict.size != 0; }   /* weights::empty, weights.hpp:146 */
void oracle_weight_batch(const oracle_dict* d, const uint64_t* ids, uint64_t n, uint64_t* weights_out) {
    for (uint64_t q = 0; q != n; ++q) {
        uint64_t i = ef_prev_leq_pos(&d->w_lengths, ids[q]);
        uint64_t id = cvec_access(&d->w_values, i);
        weights_out[q] = cvec_access(&d->w_dict, id);
    }
}

/* ------------------------------------------------------------------------------------------
 * streaming membership: a literal restatement of the reference state machine
 * (streaming_query.hpp:56-197) with the rolling minimizers replaced by from-scratch ones (the
 * reference asserts they are equal, minimizer_iterator.hpp:56-57,138-139); the string iterator
 * (kmer_iterator.hpp) is restated literally above.
 * Drivers: FASTQ/FASTA one read per record, reset per read, reads shorter than k skipped
 * (query.cpp:53-108).
 * ---------------------------------------------------------------------------------------- */
/* kmer_iterator<Kmer, bit_vector>, kmer_iterator.hpp:8-86, restated literally -- including the
   order in which fill_buff_reverse appends its 64-bit words (:72-79): for a 128-bit k-mer type
   the two halves end up swapped, so in a max_k=63 build a backward extension practically never
   matches and the reference falls back to a search.  Only the searches/extensions counters see
   this; the oracle (and the CUDA path) reproduce it bit-for-bit. */
typedef struct { uint64_t pos, avail; u128 buff; } kmer_iter_t;

static void kit_at(kmer_iter_t* it, uint64_t pos) { it->pos = pos; it->avail = 0; it->buff = 0; }
static u128 kit_word(const oracle_dict* d, uint64_t pos) {
    /* get_word64 without the num_bits guard of read_kmer_at; out-of-range words read as zero
       (the reference has sentinel words there, encode_strings.cpp:183-188) */
    if ((pos >> 6) >= d->strings.nwords) return 0;
    return (u128)get_word64(&d->strings, pos);
}
static void kit_append64(const oracle_dict* d, kmer_iter_t* it, u128 w) {
    it->buff = d->kmer_bits == 64 ? w : ((it->buff << 64) | w);
}
static void kit_fill(const oracle_dict* d, kmer_iter_t* it) {                        /* :65-70 */
    for (int i = d->kmer_bits - 64; i >= 0; i -= 64) kit_append64(d, it, kit_word(d, it->pos + (uint64_t)i));
    it->avail = (uint64_t)d->kmer_bits;
}
static void kit_fill_reverse(const oracle_dict* d, kmer_iter_t* it) {                /* :72-79 */
    uint64_t W = (uint64_t)d->kmer_bits;
    uint64_t base = it->pos > W ? it->pos : W;
    for (int i = d->kmer_bits; i > 0; i -= 64) kit_append64(d, it, kit_word(d, base - (uint64_t)i));
    it->avail = it->pos < W ? it->pos : W;
    if (W - it->avail) it->buff = (W - it->avail) >= 128 ? 0 : it->buff << (W - it->avail);
    if (d->kmer_bits == 64) it->buff &= (u128)UINT64_MAX;
}
static void kit_trim(const oracle_dict* d, kmer_iter_t* it) { if (d->kmer_bits == 64) it->buff &= (u128)UINT64_MAX; }
static u128 kit_get(const oracle_dict* d, kmer_iter_t* it) {                         /* :27-32 */
    if (it->avail < 2 * d->k) kit_fill(d, it);
    return it->buff & mask_bits((unsigned)(2 * d->k));
}
static u128 kit_get_reverse(const oracle_dict* d, kmer_iter_t* it) {                 /* :34-39 */
    if (it->avail < 2 * d->k) kit_fill_reverse(d, it);
    return it->buff >> ((uint64_t)d->kmer_bits - 2 * d->k);
}
static void kit_next(const oracle_dict* d, kmer_iter_t* it) {                        /* :41-46 */
    if (it->avail < 2) kit_fill(d, it);
    it->buff >>= 2; it->avail -= 2; it->pos += 2;
}
static void kit_next_reverse(const oracle_dict* d, kmer_iter_t* it) {                /* :48-53 */
    if (it->avail < 2) kit_fill_reverse(d, it);
    it->buff <<= 2; kit_trim(d, it); it->avail -= 2; it->pos -= 2;
}

static int is_valid_base(char c) {                                                   /* kmer.hpp:209-219,253-255 */
    switch (c) { case 'A': case 'C': case 'G': case 'T': case 'a': case 'c': case 'g': case 't': return 1; default: return 0; }
}

void oracle_streaming_reads(const oracle_dict* d, const char* bases, const uint64_t* read_offsets,
                            uint64_t num_reads, uint64_t* kmer_ids, oracle_result* full, oracle_report* rep) {
    const uint64_t k = d->k;
    uint64_t n_search = 0, n_ext = 0, n_inv = 0, n_neg = 0, n_kmers = 0, w = 0;
    for (uint64_t r = 0; r != num_reads; ++r) {
        const char* line = bases + read_offsets[r];
        uint64_t len = read_offsets[r + 1] - read_offsets[r];
        if (len < k) continue;
        uint64_t nk = len - k + 1;
        n_kmers += nk;
        /* reset(), streaming_query.hpp:48-54 */
        int start = 1; uint64_t remaining = 0; oracle_result res; result_init(&res, 1);
        uint64_t prev_mini = INVALID, prev_mini_rc = INVALID;
        kmer_iter_t it; kit_at(&it, 0);
        for (uint64_t i = 0; i != nk; ++i, ++w) {
            const char* s = line + i;
            int valid = 1;
            if (start) { for (uint64_t j = 0; j != k; ++j) if (!is_valid_base(s[j])) { valid = 0; break; } }
            else valid = is_valid_base(s[k - 1]);
            if (!valid) {                                                             /* :59-65 */
                n_inv += 1; start = 1; remaining = 0; result_init(&res, 1);
                prev_mini = prev_mini_rc = INVALID;
                /* the rolling iterators are reset too; their "previous" values only matter through the
                   unchanged-minimizer shortcut, which needs res.minimizer_found == false -- impossible
                   right after a reset (result_init sets it true) */
                if (kmer_ids) kmer_ids[w] = res.kmer_id;
                if (full) full[w] = res;
                continue;
            }
            u128 kmer = string_to_kmer(s, k), kmer_rc = rc_kmer(kmer, k, d->kmer_bits);
            uint64_t mini, pos, mini_rc, pos_rc;
            compute_minimizer(d, kmer, &mini, &pos);
            compute_minimizer(d, kmer_rc, &mini_rc, &pos_rc);
            int do_seed = 1;
            if (remaining != 0) {                                                     /* :88-99 */
                u128 expected;
                if (res.kmer_orientation == 1) { kit_next(d, &it); expected = kit_get(d, &it); }
                else { kit_next_reverse(d, &it); expected = kit_get_reverse(d, &it); }
                if (expected == kmer || expected == kmer_rc) {
                    n_ext += 1;
                    res.kmer_id += (uint64_t)res.kmer_orientation;
                    res.kmer_id_in_string += (uint64_t)res.kmer_orientation;
                    remaining -= 1;
                    do_seed = 0;
                }
            }
            if (do_seed) {                                                            /* seed() :144-197 */
                remaining = 0;
                if (mini == prev_mini && mini_rc == prev_mini_rc && res.minimizer_found == 0) {
                    n_neg += 1;
                } else {
                    if (d->canonical) {
                        if (mini < mini_rc) lookup_canonical_with(d, kmer, kmer_rc, mini, pos, &res);
                        else if (mini_rc < mini) lookup_canonical_with(d, kmer, kmer_rc, mini_rc, pos_rc, &res);
                        else {
                            lookup_canonical_with(d, kmer, kmer_rc, mini, pos, &res);
                            if (res.kmer_id == INVALID) lookup_canonical_with(d, kmer, kmer_rc, mini_rc, pos_rc, &res);
                        }
                    } else {
                        lookup_regular(d, kmer, &res);
                        uint64_t mf = res.minimizer_found;
                        if (res.kmer_id == INVALID) {
                            lookup_regular(d, kmer_rc, &res);
                            res.kmer_orientation = -1;
                            res.minimizer_found = res.minimizer_found || mf;
                        }
                    }
                    if (res.kmer_id == INVALID) { n_neg += 1; }
                    else {
                        n_search += 1;
                        uint64_t kmer_offset = 2 * (res.kmer_id + res.string_id * (k - 1));
                        remaining = (res.string_end - res.string_begin - k) - res.kmer_id_in_string;
                        if (res.kmer_orientation == -1) { kmer_offset += 2 * k; remaining = res.kmer_id_in_string; }
                        kit_at(&it, kmer_offset);
                    }
                }
            }
            prev_mini = mini; prev_mini_rc = mini_rc; start = 0;
            if (kmer_ids) kmer_ids[w] = res.kmer_id;
            if (full) full[w] = res;
        }
    }
    rep->num_kmers = n_kmers; rep->num_searches = n_search; rep->num_extensions = n_ext;
    rep->num_positive_kmers = n_search + n_ext; rep->num_negative_kmers = n_neg; rep->num_invalid_kmers = n_inv;
}
