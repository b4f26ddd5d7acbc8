// Drop-in for the reference's `sshash query -i <index> -q <reads.fa|fq[.gz]> [--multiline]`
// (tools/query.cpp:5-70): same call -- dict.streaming_query_from_file(query_filename, multiline) --
// same report on stdout and the same json line on stderr (essentials::json_lines, all values quoted).
//   g++ -std=c++17 -O2 examples/query_example.cpp -o query_example sshash_b200/libsshash_gpu.so -Wl,-rpath,$PWD/sshash_b200
//   ./query_example -i tests/golden/se_k31_m13.sshash -q reads.fastq.gz
#include <chrono>
#include <cstring>
#include <iostream>
#include <string>

#include "../sshash_b200/csrc/dictionary.hpp"

int main(int argc, char** argv) {
    std::string index_filename, query_filename;
    bool multiline = false;
    for (int i = 1; i < argc; ++i) {
        if (!std::strcmp(argv[i], "-i") && i + 1 < argc) index_filename = argv[++i];
        else if (!std::strcmp(argv[i], "-q") && i + 1 < argc) query_filename = argv[++i];
        else if (!std::strcmp(argv[i], "--multiline")) multiline = true;
    }
    if (index_filename.empty() || query_filename.empty()) {
        std::cerr << "usage: " << argv[0] << " -i <index.sshash> -q <reads.fa|fq[.gz]> [--multiline]\n";
        return 2;
    }
    try {
        sshash_b200::dictionary dict(index_filename);
        auto t0 = std::chrono::high_resolution_clock::now();
        auto report = dict.streaming_query_from_file(query_filename, multiline);
        const double ms = std::chrono::duration<double, std::milli>(std::chrono::high_resolution_clock::now() - t0).count();

        std::cout << "==== query report:\n";
        std::cout << "num_kmers = " << report.num_kmers << std::endl;
        std::cout << "num_positive_kmers = " << report.num_positive_kmers << " ("
                  << (report.num_positive_kmers * 100.0) / report.num_kmers << "%)" << std::endl;
        std::cout << "num_negative_kmers = " << report.num_negative_kmers << " ("
                  << (report.num_negative_kmers * 100.0) / report.num_kmers << "%)" << std::endl;
        std::cout << "num_invalid_kmers = " << report.num_invalid_kmers << " ("
                  << (report.num_invalid_kmers * 100.0) / report.num_kmers << "%)" << std::endl;
        std::cout << "num_searches = " << report.num_searches << "/" << report.num_positive_kmers << " ("
                  << (report.num_searches * 100.0) / report.num_positive_kmers << "%)" << std::endl;
        std::cout << "num_extensions = " << report.num_extensions << "/" << report.num_positive_kmers << " ("
                  << (report.num_extensions * 100.0) / report.num_positive_kmers << "%)" << std::endl;
        std::cout << "elapsed = " << ms / 1000 << " sec / " << ms / 1000 / 60 << " min / " << (ms * 1e6) / report.num_kmers
                  << " ns/kmer" << std::endl;

        auto q = [](std::string const& name, std::string const& value) { return "\"" + name + "\": \"" + value + "\""; };
        std::cerr << "{" << q("index_filename", index_filename) << ", " << q("query_filename", query_filename) << ", "
                  << q("num_kmers", std::to_string(report.num_kmers)) << ", "
                  << q("num_positive_kmers", std::to_string(report.num_positive_kmers)) << ", "
                  << q("num_negative_kmers", std::to_string(report.num_negative_kmers)) << ", "
                  << q("num_invalid_kmers", std::to_string(report.num_invalid_kmers)) << ", "
                  << q("num_searches", std::to_string(report.num_searches)) << ", "
                  << q("num_extensions", std::to_string(report.num_extensions)) << ", "
                  << q("elapsed_millisec", std::to_string((uint64_t)ms)) << "}\n";
        return 0;
    } catch (std::exception const& e) {
        std::cerr << "error: " << e.what() << "\n";
        return 1;
    }
}
