/*
 * oracle/ref_harness.cpp -- TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * A thin extern "C" shim around the UNMODIFIED reference (jermp/sshash @ afff26dc) so that the
 * C restatement (oracle/sshash_oracle.c) and the CUDA path can be checked against the real thing.
 * It #includes the reference's translation units where they lie under /root/reference, exactly
 * like the reference's own unity build does (tools/sshash.cpp:9-16); no reference source is
 * copied into this repository.  Built by oracle/Makefile into oracle/_ref/libsshash_ref{31,63}.so
 * (the 63 variant adds -DSSHASH_USE_MAX_KMER_LENGTH_63, CMakeLists.txt:18-23).
 *
 * Entry points (all plain C types):
 *   ref_build            dictionary::build + essentials::save   (tools/build.cpp:62-95)
 *   ref_open/ref_close   essentials::load                        (tools/common.hpp:19-29)
 *   ref_info             k(), m(), canonical(), num_kmers(), ... (include/dictionary.hpp:31-38)
 *   ref_lookup_batch     dictionary::lookup(Kmer, check_rc)      (src/dictionary.cpp:64-78)
 *   ref_access_batch     dictionary::access                      (src/dictionary.cpp:90-94)
 *   ref_weight_batch     dictionary::weight                      (src/dictionary.cpp:96-100)
 *   ref_kmer_neighbours_batch / ref_string_neighbours_batch
 *                        dictionary::kmer_neighbours etc.        (src/dictionary.cpp:112-201)
 *   ref_streaming_file   dictionary::streaming_query_from_file   (src/query.cpp:118-175)
 *   ref_streaming_reads  streaming_query<>::lookup per read      (include/streaming_query.hpp:56-109)
 */
#include <iostream>
#include <thread>
#include <vector>
#include <cstring>
#include <chrono>

#include "include/dictionary_types.hpp"
#include "include/streaming_query.hpp"

#include "src/builder/build.cpp"
#include "src/dictionary.cpp"
#include "src/query.cpp"
#include "src/info.cpp"

using namespace sshash;

extern "C" {

struct ref_lookup_result {
    uint64_t kmer_id;
    uint64_t kmer_id_in_string;
    uint64_t kmer_offset;
    int64_t kmer_orientation;
    uint64_t string_id;
    uint64_t string_begin;
    uint64_t string_end;
    uint64_t minimizer_found;
};

struct ref_info_t {
    uint64_t num_kmers, num_strings, k, m, canonical, weighted, max_k;
};

struct ref_report_t {
    uint64_t num_kmers, num_positive_kmers, num_negative_kmers, num_invalid_kmers, num_searches,
        num_extensions;
};

static thread_local std::string g_err;
const char* ref_last_error() { return g_err.c_str(); }

int ref_max_k() { return default_kmer_t::max_k; }

int ref_build(const char* input, uint64_t k, uint64_t m, int canonical, uint64_t threads,
              uint64_t seed, const char* tmp_dir, const char* output, int verbose, int weighted) {
    try {
        build_configuration cfg;
        cfg.k = k;
        cfg.m = m;
        cfg.canonical = canonical != 0;
        cfg.num_threads = threads ? threads : 1;
        if (seed) cfg.seed = seed;
        cfg.verbose = verbose != 0;
        cfg.weighted = weighted != 0;       // tools/build.cpp:63
        if (tmp_dir && *tmp_dir) {
            cfg.tmp_dirname = tmp_dir;
            essentials::create_directory(cfg.tmp_dirname);
        }
        dictionary_type dict;
        dict.build(input, cfg);
        essentials::save(dict, output);
        return 0;
    } catch (std::exception const& e) {
        g_err = e.what();
        return 1;
    }
}

void* ref_open(const char* path) {
    try {
        auto* d = new dictionary_type();
        essentials::load(*d, path);
        return d;
    } catch (std::exception const& e) {
        g_err = e.what();
        return nullptr;
    }
}

void ref_close(void* h) { delete static_cast<dictionary_type*>(h); }

void ref_info(void* h, ref_info_t* out) {
    auto* d = static_cast<dictionary_type*>(h);
    out->num_kmers = d->num_kmers();
    out->num_strings = d->num_strings();
    out->k = d->k();
    out->m = d->m();
    out->canonical = d->canonical();
    out->weighted = d->weighted();
    out->max_k = default_kmer_t::max_k;
}

static inline default_kmer_t load_kmer(const uint64_t* p, uint64_t i) {
    default_kmer_t x;
#ifdef SSHASH_USE_MAX_KMER_LENGTH_63
    x.bits = (__uint128_t(p[2 * i + 1]) << 64) | __uint128_t(p[2 * i]);
#else
    x.bits = p[i];
#endif
    return x;
}

static inline void store_kmer(uint64_t* p, uint64_t i, default_kmer_t x) {
#ifdef SSHASH_USE_MAX_KMER_LENGTH_63
    p[2 * i] = uint64_t(x.bits);
    p[2 * i + 1] = uint64_t(x.bits >> 64);
#else
    p[i] = x.bits;
#endif
}

static inline void copy_result(ref_lookup_result* o, lookup_result const& r) {
    o->kmer_id = r.kmer_id;
    o->kmer_id_in_string = r.kmer_id_in_string;
    o->kmer_offset = r.kmer_offset;
    o->kmer_orientation = r.kmer_orientation;
    o->string_id = r.string_id;
    o->string_begin = r.string_begin;
    o->string_end = r.string_end;
    o->minimizer_found = r.minimizer_found;
}

/* kmers: n packed k-mers (1 or 2 little-endian u64 words each). ids/full may be null.
   Returns elapsed seconds of the lookup loop (all threads). */
double ref_lookup_batch(void* h, const uint64_t* kmers, uint64_t n, int check_rc, uint64_t* ids,
                        ref_lookup_result* full, uint64_t threads) {
    auto* d = static_cast<dictionary_type*>(h);
    if (threads == 0) threads = 1;
    auto work = [&](uint64_t lo, uint64_t hi) {
        for (uint64_t i = lo; i != hi; ++i) {
            auto r = d->lookup(load_kmer(kmers, i), check_rc != 0);
            if (ids) ids[i] = r.kmer_id;
            if (full) copy_result(full + i, r);
            if (!ids && !full) essentials::do_not_optimize_away(r.kmer_id);
        }
    };
    auto t0 = std::chrono::steady_clock::now();
    if (threads == 1) {
        work(0, n);
    } else {
        std::vector<std::thread> pool;
        uint64_t chunk = (n + threads - 1) / threads;
        for (uint64_t t = 0; t != threads; ++t) {
            uint64_t lo = std::min(n, t * chunk), hi = std::min(n, lo + chunk);
            if (lo < hi) pool.emplace_back(work, lo, hi);
        }
        for (auto& th : pool) th.join();
    }
    auto t1 = std::chrono::steady_clock::now();
    return std::chrono::duration<double>(t1 - t0).count();
}

/* ASCII variant: n strings of exactly k chars, back to back (src/dictionary.cpp:58-63). */
void ref_lookup_batch_ascii(void* h, const char* kmers, uint64_t n, int check_rc, uint64_t* ids,
                            ref_lookup_result* full) {
    auto* d = static_cast<dictionary_type*>(h);
    const uint64_t k = d->k();
    for (uint64_t i = 0; i != n; ++i) {
        auto r = d->lookup(kmers + i * k, check_rc != 0);
        if (ids) ids[i] = r.kmer_id;
        if (full) copy_result(full + i, r);
    }
}

void ref_access_batch(void* h, const uint64_t* ids, uint64_t n, uint64_t* kmers_out) {
    auto* d = static_cast<dictionary_type*>(h);
    const uint64_t k = d->k();
    std::string s(k, 0);
    for (uint64_t i = 0; i != n; ++i) {
        d->access(ids[i], s.data());
        store_kmer(kmers_out, i, util::string_to_uint_kmer<default_kmer_t>(s.data(), k));
    }
}

/* which: 1 = kmer_forward_neighbours, 2 = kmer_backward_neighbours, 3 = kmer_neighbours.
   out: 8 records per k-mer: forward[A,C,T,G] then backward[A,C,T,G] (include/util.hpp:77-81). */
void ref_weight_batch(void* h, const uint64_t* ids, uint64_t n, uint64_t* weights_out) {
    auto* d = static_cast<dictionary_type*>(h);
    for (uint64_t i = 0; i != n; ++i) weights_out[i] = d->weight(ids[i]);
}

void ref_kmer_neighbours_batch(void* h, const uint64_t* kmers, uint64_t n, int check_rc, int which,
                               ref_lookup_result* out) {
    auto* d = static_cast<dictionary_type*>(h);
    for (uint64_t i = 0; i != n; ++i) {
        auto x = load_kmer(kmers, i);
        neighbourhood<default_kmer_t> nb = which == 1   ? d->kmer_forward_neighbours(x, check_rc != 0)
                                           : which == 2 ? d->kmer_backward_neighbours(x, check_rc != 0)
                                                        : d->kmer_neighbours(x, check_rc != 0);
        for (int j = 0; j != 4; ++j) {
            copy_result(out + 8 * i + j, nb.forward[j]);
            copy_result(out + 8 * i + 4 + j, nb.backward[j]);
        }
    }
}

void ref_string_neighbours_batch(void* h, const uint64_t* string_ids, uint64_t n, int check_rc,
                                 ref_lookup_result* out) {
    auto* d = static_cast<dictionary_type*>(h);
    for (uint64_t i = 0; i != n; ++i) {
        auto nb = d->string_neighbours(string_ids[i], check_rc != 0);
        for (int j = 0; j != 4; ++j) {
            copy_result(out + 8 * i + j, nb.forward[j]);
            copy_result(out + 8 * i + 4 + j, nb.backward[j]);
        }
    }
}

int ref_streaming_file(void* h, const char* path, int multiline, ref_report_t* out,
                       double* seconds) {
    auto* d = static_cast<dictionary_type*>(h);
    try {
        auto t0 = std::chrono::steady_clock::now();
        auto r = d->streaming_query_from_file(path, multiline != 0);
        auto t1 = std::chrono::steady_clock::now();
        if (seconds) *seconds = std::chrono::duration<double>(t1 - t0).count();
        out->num_kmers = r.num_kmers;
        out->num_positive_kmers = r.num_positive_kmers;
        out->num_negative_kmers = r.num_negative_kmers;
        out->num_invalid_kmers = r.num_invalid_kmers;
        out->num_searches = r.num_searches;
        out->num_extensions = r.num_extensions;
        return 0;
    } catch (std::exception const& e) {
        g_err = e.what();
        return 1;
    }
}

}  // extern "C"

template <bool canonical>
static void streaming_reads_impl(dictionary_type const* d, const char* bases,
                                 const uint64_t* read_offsets, uint64_t num_reads,
                                 uint64_t* kmer_ids, ref_lookup_result* full, ref_report_t* out) {
    streaming_query<dictionary_type, canonical> q(d);
    const uint64_t k = d->k();
    uint64_t w = 0, num_kmers = 0;
    for (uint64_t r = 0; r != num_reads; ++r) {
        q.reset();
        const char* line = bases + read_offsets[r];
        uint64_t len = read_offsets[r + 1] - read_offsets[r];
        if (len < k) continue;
        uint64_t nk = len - k + 1;
        num_kmers += nk;
        for (uint64_t i = 0; i != nk; ++i, ++w) {
            auto res = q.lookup(line + i);
            if (kmer_ids) kmer_ids[w] = res.kmer_id;
            if (full) copy_result(full + w, res);
        }
    }
    out->num_kmers = num_kmers;
    out->num_searches = q.num_searches();
    out->num_extensions = q.num_extensions();
    out->num_positive_kmers = q.num_positive_lookups();
    out->num_negative_kmers = q.num_negative_lookups();
    out->num_invalid_kmers = q.num_invalid_lookups();
}

extern "C" {

/* Reads given as one concatenated char buffer + (num_reads+1) offsets; follows the FASTQ driver
   (src/query.cpp:78-108): reset per read, reads shorter than k skipped. Per-window ids are
   written densely in window order (sum over reads of max(0,len-k+1) entries). */
double ref_streaming_reads(void* h, const char* bases, const uint64_t* read_offsets,
                           uint64_t num_reads, uint64_t* kmer_ids, ref_lookup_result* full,
                           ref_report_t* out) {
    auto* d = static_cast<dictionary_type*>(h);
    auto t0 = std::chrono::steady_clock::now();
    if (d->canonical())
        streaming_reads_impl<true>(d, bases, read_offsets, num_reads, kmer_ids, full, out);
    else
        streaming_reads_impl<false>(d, bases, read_offsets, num_reads, kmer_ids, full, out);
    auto t1 = std::chrono::steady_clock::now();
    return std::chrono::duration<double>(t1 - t0).count();
}

}  // extern "C"
