// Host C++ above the C ABI: mirrors how the reference's tools call the dictionary
// (tools/perf.hpp:41-64 positive lookups, test/check.hpp:29-49 lookup(access(id)) == id).
//   g++ -std=c++17 -O2 examples/lookup_example.cpp -o lookup_example sshash_b200/libsshash_gpu.so -Wl,-rpath,$PWD/sshash_b200
//   ./lookup_example tests/golden/se_k31_m13.sshash [queries]
#include <chrono>
#include <cstdio>
#include <random>
#include <string>
#include <vector>

#include "../sshash_b200/csrc/dictionary.hpp"

int main(int argc, char** argv) {
    if (argc < 2) { std::fprintf(stderr, "usage: %s <index.sshash> [num_queries]\n", argv[0]); return 2; }
    const uint64_t n = argc > 2 ? std::stoull(argv[2]) : 1000000;
    try {
        sshash_b200::dictionary dict(argv[1]);
        std::printf("k=%lu m=%lu canonical=%d num_kmers=%lu num_strings=%lu\n", dict.k(), dict.m(), dict.canonical(),
                    dict.num_kmers(), dict.num_strings());
        const uint64_t w = dict.words_per_kmer();
        std::mt19937_64 rng(42);
        std::vector<uint64_t> ids(n), kmers(n * w), got(n);
        for (auto& id : ids) id = rng() % dict.num_kmers();
        dict.access_batch(ids.data(), n, kmers.data());
        auto t0 = std::chrono::steady_clock::now();
        dict.lookup_batch(kmers.data(), n, got.data());
        double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        uint64_t bad = 0;
        for (uint64_t i = 0; i != n; ++i) bad += got[i] != ids[i];
        std::printf("%lu positive lookups (host buffers) in %.3f ms, %lu mismatches\n", n, s * 1e3, bad);
        std::string kmer(dict.k(), 'A');
        dict.access(0, kmer.data());
        auto r = dict.lookup(kmer.c_str());
        std::printf("access(0) = %s -> lookup id %lu orientation %ld string [%lu,%lu)\n", kmer.c_str(), r.kmer_id,
                    r.kmer_orientation, r.string_begin, r.string_end);
        dict.access(1, kmer.data());                       // k-mer 1 follows k-mer 0 in string 0
        auto nb = dict.kmer_neighbours(kmer.c_str());
        bool back_ok = false;
        for (auto const& b : nb.backward) back_ok |= b.kmer_id == 0;
        std::printf("kmer_neighbours(access(1)): backward contains id 0: %s\n", back_ok ? "yes" : "no");
        // the stateful per-k-mer object (include/streaming_query.hpp:36-115): 10 consecutive k-mers of string 0
        // walked forward = 1 search + 9 extensions, then an invalid k-mer, then a k-mer of another string = a search
        sshash_b200::streaming_query sq(&dict);
        uint64_t sq_bad = 0;
        for (uint64_t id = 0; id != 10; ++id) {
            dict.access(id, kmer.data());
            sq_bad += sq.lookup(kmer.c_str()).kmer_id != id;
        }
        std::string inval(dict.k(), 'N');
        sq_bad += sq.lookup(inval.c_str()).kmer_id != sshash_b200::constants::invalid_uint64;
        dict.access(dict.num_kmers() - 1, kmer.data());
        sq_bad += sq.lookup(kmer.c_str()).kmer_id != dict.num_kmers() - 1;
        sq_bad += !(sq.num_searches() == 2 && sq.num_extensions() == 9 && sq.num_invalid_lookups() == 1 && sq.num_negative_lookups() == 0);
        std::printf("streaming_query object: %lu searches, %lu extensions, %lu invalid -> %s\n", sq.num_searches(), sq.num_extensions(),
                    sq.num_invalid_lookups(), sq_bad ? "MISMATCH" : "ok");
        // every GPU of the box behind one handle: same ids, 32-bit ids, membership
        sshash_b200::multi_dictionary multi(argv[1]);
        std::vector<uint64_t> got_m(n);
        std::vector<uint32_t> got32(n);
        std::vector<uint8_t> mem(n);
        multi.lookup_batch(kmers.data(), n, got_m.data());
        multi.lookup_batch_u32(kmers.data(), n, got32.data());
        multi.is_member_batch(kmers.data(), n, mem.data());
        uint64_t multi_bad = 0;
        for (uint64_t i = 0; i != n; ++i) multi_bad += (got_m[i] != ids[i]) + (got32[i] != (uint32_t)ids[i]) + (mem[i] != 1);
        std::printf("multi_dictionary over %d GPU(s): %lu mismatches\n", multi.num_devices(), multi_bad);
        return bad == 0 && r.kmer_id == 0 && back_ok && sq_bad == 0 && multi_bad == 0 ? 0 : 1;
    } catch (std::exception const& e) {
        std::fprintf(stderr, "error: %s\n", e.what());
        return 1;
    }
}
