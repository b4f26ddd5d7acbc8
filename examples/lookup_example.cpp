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
        return bad == 0 && r.kmer_id == 0 && back_ok ? 0 : 1;
    } catch (std::exception const& e) {
        std::fprintf(stderr, "error: %s\n", e.what());
        return 1;
    }
}
