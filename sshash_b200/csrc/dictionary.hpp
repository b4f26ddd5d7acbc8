// dictionary.hpp -- C++ host-side drop-in for the reference's `sshash::dictionary<Kmer, Offsets>`
// on the lookup path (reference include/dictionary.hpp:10-181), header-only, on top of the C ABI
// (include/sshash_gpu.h).  Same method names, argument meaning and error behaviour:
//
//   reference                                             this wrapper
//   essentials::load(dict, path) / open_dictionary        sshash_b200::dictionary dict(path[, device, max_k])
//   dict.k() m() canonical() num_kmers() num_strings()    same            (dictionary.hpp:31-38)
//   lookup_result lookup(char const*, bool = true)        same            (dictionary.hpp:41, dictionary.cpp:58-63)
//   lookup_result lookup(Kmer, bool = true)               lookup(kmer_t)  (dictionary.hpp:42, dictionary.cpp:64-78)
//   bool is_member(char const* / Kmer, bool = true)       same            (dictionary.hpp:75-76)
//   void access(uint64_t kmer_id, char* string_kmer)      same            (dictionary.hpp:71)
//   uint64_t weight(uint64_t kmer_id)                     same            (dictionary.hpp:65-66, weights.hpp:148-153)
//   kmer_neighbours / kmer_forward_neighbours /           same            (dictionary.hpp:50-66, dictionary.cpp:112-201)
//     kmer_backward_neighbours / string_neighbours
//   streaming_query_from_file(filename, multiline)        same            (dictionary.hpp:81-82)
//   -- callers that loop over lookup (tools/perf.hpp:55-60, test/check.hpp:29-31) --
//                                                         lookup_batch / is_member_batch / access_batch
//
// Open failures and a major-version mismatch throw std::runtime_error like the reference
// (essentials.hpp:413-417, util.hpp:191-195); lookups never throw for absent k-mers: not found is
// kmer_id == constants::invalid_uint64.  Scalar calls are one-element batches: use them for
// parity / debugging, the batch calls for throughput.
#pragma once

#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/sshash_gpu.h"

namespace sshash_b200 {

namespace constants {
constexpr uint64_t invalid_uint64 = uint64_t(-1);     // include/constants.hpp:5
constexpr int forward_orientation = 1;                // include/constants.hpp:19
constexpr int backward_orientation = -1;              // include/constants.hpp:20
}  // namespace constants

using lookup_result = sshash_lookup_result;           // include/util.hpp:38-62
using streaming_query_report = sshash_streaming_report;  // include/util.hpp:21-36

struct kmer_t {                                       // 2 bits/base, base 0 in the low bits
    uint64_t lo = 0, hi = 0;
};

// util::string_to_uint_kmer, include/util.hpp:207-213 (no validation, like the reference)
inline kmer_t string_to_uint_kmer(char const* str, uint64_t k) {
    kmer_t x;
    for (uint64_t i = 0; i != k; ++i) {
        uint64_t c = (static_cast<uint8_t>(str[i]) >> 1) & 3;
        if (i < 32) x.lo |= c << (2 * i); else x.hi |= c << (2 * (i - 32));
    }
    return x;
}
// util::uint_kmer_to_string, include/util.hpp:215-219 (alphabet "ACTG", include/kmer.hpp:118)
inline void uint_kmer_to_string(kmer_t x, char* str, uint64_t k) {
    for (uint64_t i = 0; i != k; ++i)
        str[i] = "ACTG"[i < 32 ? (x.lo >> (2 * i)) & 3 : (x.hi >> (2 * (i - 32))) & 3];
}

class dictionary {
public:
    explicit dictionary(std::string const& index_filename, int device = 0, int max_k = 0) {
        check(sshash_gpu_open(index_filename.c_str(), device, max_k, &m_dict));
        check(sshash_gpu_info(m_dict, &m_info));
        m_words = m_info.max_k == 31 ? 1 : 2;
    }
    ~dictionary() { sshash_gpu_close(m_dict); }
    dictionary(dictionary const&) = delete;
    dictionary& operator=(dictionary const&) = delete;

    uint64_t num_kmers() const { return m_info.num_kmers; }
    uint64_t num_strings() const { return m_info.num_strings; }
    uint64_t k() const { return m_info.k; }
    uint64_t m() const { return m_info.m; }
    bool canonical() const { return m_info.canonical != 0; }
    bool weighted() const { return m_info.weighted != 0; }
    uint64_t words_per_kmer() const { return m_words; }
    sshash_gpu_info_t const& info() const { return m_info; }
    sshash_gpu_dict* handle() const { return m_dict; }

    /* Lookup queries. */
    lookup_result lookup(char const* string_kmer, bool check_reverse_complement = true) const {
        lookup_result r;
        check(sshash_gpu_lookup_batch_ascii(m_dict, string_kmer, 1, check_reverse_complement, nullptr, &r, nullptr));
        return r;
    }
    lookup_result lookup(kmer_t uint_kmer, bool check_reverse_complement = true) const {
        uint64_t w[2] = {uint_kmer.lo, uint_kmer.hi};
        lookup_result r;
        check(sshash_gpu_lookup_batch(m_dict, w, 1, check_reverse_complement, nullptr, &r, nullptr));
        return r;
    }
    bool is_member(char const* string_kmer, bool check_reverse_complement = true) const {
        return lookup(string_kmer, check_reverse_complement).kmer_id != constants::invalid_uint64;
    }
    bool is_member(kmer_t uint_kmer, bool check_reverse_complement = true) const {
        return lookup(uint_kmer, check_reverse_complement).kmer_id != constants::invalid_uint64;
    }
    void access(uint64_t kmer_id, char* string_kmer) const {
        uint64_t w[2] = {0, 0};
        check(sshash_gpu_access_batch(m_dict, &kmer_id, 1, w, nullptr));
        uint_kmer_to_string(kmer_t{w[0], w[1]}, string_kmer, k());
    }

    uint64_t weight(uint64_t kmer_id) const {
        uint64_t w = 0;
        check(sshash_gpu_weight_batch(m_dict, &kmer_id, 1, &w, nullptr));
        return w;
    }

    /* Batched forms: host or device pointers; `stream` only matters for device pointers. */
    void lookup_batch(uint64_t const* kmers, uint64_t n, uint64_t* kmer_ids, bool check_reverse_complement = true,
                      lookup_result* full = nullptr, void* stream = nullptr) const {
        check(sshash_gpu_lookup_batch(m_dict, kmers, n, check_reverse_complement, kmer_ids, full, stream));
    }
    void lookup_batch(char const* string_kmers, uint64_t n, uint64_t* kmer_ids, bool check_reverse_complement = true,
                      lookup_result* full = nullptr, void* stream = nullptr) const {
        check(sshash_gpu_lookup_batch_ascii(m_dict, string_kmers, n, check_reverse_complement, kmer_ids, full, stream));
    }
    /* 32-bit ids (dictionaries with < 2^32 - 1 k-mers): "not found" is UINT32_MAX */
    void lookup_batch_u32(uint64_t const* kmers, uint64_t n, uint32_t* kmer_ids32, bool check_reverse_complement = true,
                          void* stream = nullptr) const {
        check(sshash_gpu_lookup_batch_u32(m_dict, kmers, n, check_reverse_complement, kmer_ids32, stream));
    }
    void is_member_batch(uint64_t const* kmers, uint64_t n, uint8_t* member, bool check_reverse_complement = true,
                         void* stream = nullptr) const {
        check(sshash_gpu_is_member_batch(m_dict, kmers, n, check_reverse_complement, member, stream));
    }
    void access_batch(uint64_t const* kmer_ids, uint64_t n, uint64_t* kmers_out, void* stream = nullptr) const {
        check(sshash_gpu_access_batch(m_dict, kmer_ids, n, kmers_out, stream));
    }

    void weight_batch(uint64_t const* kmer_ids, uint64_t n, uint64_t* weights_out, void* stream = nullptr) const {
        check(sshash_gpu_weight_batch(m_dict, kmer_ids, n, weights_out, stream));
    }

    /* Navigational queries (include/dictionary.hpp:50-66): forward[A,C,T,G] then backward[A,C,T,G]. */
    struct neighbourhood { lookup_result forward[4]; lookup_result backward[4]; };   // include/util.hpp:77-81
    neighbourhood kmer_neighbours(kmer_t uint_kmer, bool check_reverse_complement = true) const {
        return neighbours_of(uint_kmer, check_reverse_complement, 3);
    }
    neighbourhood kmer_forward_neighbours(kmer_t uint_kmer, bool check_reverse_complement = true) const {
        return neighbours_of(uint_kmer, check_reverse_complement, 1);
    }
    neighbourhood kmer_backward_neighbours(kmer_t uint_kmer, bool check_reverse_complement = true) const {
        return neighbours_of(uint_kmer, check_reverse_complement, 2);
    }
    neighbourhood kmer_neighbours(char const* string_kmer, bool check_reverse_complement = true) const {
        return neighbours_of(string_to_uint_kmer(string_kmer, k()), check_reverse_complement, 3);
    }
    neighbourhood string_neighbours(uint64_t string_id, bool check_reverse_complement = true) const {
        neighbourhood nb;
        check(sshash_gpu_string_neighbours_batch(m_dict, &string_id, 1, check_reverse_complement, nullptr, nb.forward, nullptr));
        return nb;
    }
    void kmer_neighbours_batch(uint64_t const* kmers, uint64_t n, uint64_t* kmer_ids /* 8n */, bool check_reverse_complement = true,
                               int which = 3, lookup_result* full = nullptr, void* stream = nullptr) const {
        check(sshash_gpu_kmer_neighbours_batch(m_dict, kmers, n, check_reverse_complement, which, kmer_ids, full, stream));
    }

    /* Streaming membership. */
    streaming_query_report streaming_query_from_file(std::string const& filename, bool multiline = false) const {
        streaming_query_report r;
        check(sshash_gpu_streaming_query_from_file(m_dict, filename.c_str(), multiline, &r));
        return r;
    }
    streaming_query_report streaming_query(char const* bases, uint64_t const* read_offsets, uint64_t num_reads,
                                           uint64_t* kmer_ids = nullptr, void* stream = nullptr) const {
        streaming_query_report r;
        check(sshash_gpu_streaming_batch(m_dict, bases, read_offsets, num_reads, kmer_ids, &r, stream));
        return r;
    }

private:
    neighbourhood neighbours_of(kmer_t x, bool check_rc, int which) const {
        static_assert(sizeof(neighbourhood) == 8 * sizeof(lookup_result), "neighbourhood must be 8 packed records");
        uint64_t w[2] = {x.lo, x.hi};
        neighbourhood nb;
        check(sshash_gpu_kmer_neighbours_batch(m_dict, w, 1, check_rc, which, nullptr, nb.forward, nullptr));
        return nb;
    }
    static void check(int status) {
        if (status != SSHASH_GPU_OK) throw std::runtime_error(sshash_gpu_last_error());
    }
    sshash_gpu_dict* m_dict = nullptr;
    sshash_gpu_info_t m_info{};
    uint64_t m_words = 1;
};

// The same index replicated on several GPUs of one box behind ONE handle (sshash_gpu_multi_*): batches are
// sharded by query inside the library, results come back in query order.  devices empty = every GPU.
class multi_dictionary {
public:
    explicit multi_dictionary(std::string const& index_filename, std::vector<int> const& devices = {}, int max_k = 0) {
        check(sshash_gpu_multi_open(index_filename.c_str(), devices.empty() ? nullptr : devices.data(), (int)devices.size(), max_k, &m_multi));
        check(sshash_gpu_info(sshash_gpu_multi_dict(m_multi, 0), &m_info));
    }
    ~multi_dictionary() { sshash_gpu_multi_close(m_multi); }
    multi_dictionary(multi_dictionary const&) = delete;
    multi_dictionary& operator=(multi_dictionary const&) = delete;

    int num_devices() const { return sshash_gpu_multi_num_devices(m_multi); }
    uint64_t num_kmers() const { return m_info.num_kmers; }
    uint64_t k() const { return m_info.k; }
    uint64_t words_per_kmer() const { return m_info.max_k == 31 ? 1 : 2; }
    sshash_gpu_dict const* replica(int i) const { return sshash_gpu_multi_dict(m_multi, i); }

    void lookup_batch(uint64_t const* kmers, uint64_t n, uint64_t* kmer_ids, bool check_reverse_complement = true) const {
        check(sshash_gpu_multi_lookup_batch(m_multi, kmers, n, check_reverse_complement, kmer_ids));
    }
    void lookup_batch_u32(uint64_t const* kmers, uint64_t n, uint32_t* kmer_ids32, bool check_reverse_complement = true) const {
        check(sshash_gpu_multi_lookup_batch_u32(m_multi, kmers, n, check_reverse_complement, kmer_ids32));
    }
    void is_member_batch(uint64_t const* kmers, uint64_t n, uint8_t* member, bool check_reverse_complement = true) const {
        check(sshash_gpu_multi_is_member_batch(m_multi, kmers, n, check_reverse_complement, member));
    }
    streaming_query_report streaming_query(char const* bases, uint64_t const* read_offsets, uint64_t num_reads,
                                           uint64_t* kmer_ids = nullptr) const {
        streaming_query_report r;
        check(sshash_gpu_multi_streaming_batch(m_multi, bases, read_offsets, num_reads, kmer_ids, &r));
        return r;
    }

private:
    static void check(int status) {
        if (status != SSHASH_GPU_OK) throw std::runtime_error(sshash_gpu_last_error());
    }
    sshash_gpu_multi* m_multi = nullptr;
    sshash_gpu_info_t m_info{};
};

// Stateful per-k-mer streaming object with the reference's interface (include/streaming_query.hpp:36-115):
// reset(), lookup(kmer) one k-mer at a time, the four counters.  For callers that interleave their own
// logic with lookups; every call is a one-element batch (tens of microseconds) -- whole reads belong in
// dictionary::streaming_query / streaming_query_from_file.  The record returned is dictionary::lookup's
// (the reference asserts they are equal, :107); a lookup counts as an EXTENSION when the previous one
// was positive and this k-mer is the next one of the same string in the previous orientation
// (:88-99, :189-195), else as a search -- exact on indexes that keep SSHash's distinct-k-mer contract.
class streaming_query {
public:
    explicit streaming_query(dictionary const* dict) : m_dict(dict) { reset(); }
    void reset() { m_prev.kmer_id = constants::invalid_uint64; }
    lookup_result lookup(char const* kmer) {
        ++m_num_kmers;
        for (uint64_t i = 0; i != m_dict->k(); ++i) {        // kmer.hpp:209-219: only ACGTacgt are valid
            const char c = kmer[i] & 0xDF;
            if (c != 'A' && c != 'C' && c != 'G' && c != 'T') {
                ++m_num_invalid;
                reset();
                lookup_result r{};
                r.kmer_id = r.kmer_id_in_string = r.kmer_offset = r.string_id = r.string_begin = r.string_end = constants::invalid_uint64;
                r.kmer_orientation = constants::forward_orientation;
                return r;
            }
        }
        lookup_result r = m_dict->lookup(kmer, true);
        if (r.kmer_id == constants::invalid_uint64) { ++m_num_negative; reset(); return r; }
        bool extension = false;
        if (m_prev.kmer_id != constants::invalid_uint64 && m_prev.string_id == r.string_id) {
            const uint64_t last = m_prev.string_end - m_prev.string_begin - m_dict->k();   // id_in_string of the string's last k-mer
            if (m_prev.kmer_orientation > 0) extension = m_prev.kmer_id_in_string != last && r.kmer_id == m_prev.kmer_id + 1;
            else extension = m_prev.kmer_id_in_string != 0 && r.kmer_id + 1 == m_prev.kmer_id;
        }
        if (extension) ++m_num_extensions; else ++m_num_searches;
        m_prev = r;
        return r;
    }
    uint64_t num_searches() const { return m_num_searches; }
    uint64_t num_extensions() const { return m_num_extensions; }
    uint64_t num_positive_lookups() const { return m_num_searches + m_num_extensions; }
    uint64_t num_negative_lookups() const { return m_num_negative; }
    uint64_t num_invalid_lookups() const { return m_num_invalid; }
    streaming_query_report report() const {
        return {m_num_kmers, m_num_searches + m_num_extensions, m_num_negative, m_num_invalid, m_num_searches, m_num_extensions};
    }

private:
    dictionary const* m_dict;
    lookup_result m_prev{};
    uint64_t m_num_kmers = 0, m_num_searches = 0, m_num_extensions = 0, m_num_negative = 0, m_num_invalid = 0;
};

}  // namespace sshash_b200
