// index_file.hpp -- host-side parser of the reference's on-disk index (format 5.x).
//
// The file is the byte stream the reference's visitor writes in declaration order
// (include/dictionary.hpp:139-152; essentials.hpp:329-407): PODs raw little-endian, every
// vector / owning_span as {u64 n; n * sizeof(T) bytes}, no padding.  This parser only records
// WHERE each array lives inside the memory-mapped file (a `Span`), so that the uploader can copy
// straight from the page cache to HBM; nothing is unpacked on the host except the two small
// structures that get a GPU-friendly re-encoding (MPHF free slots, string end-points).
#pragma once

#include <cstdint>
#include <string>
#include <vector>

namespace sshash_b200 {

struct Span {                 // n elements of `elem` bytes at file offset `off`
    uint64_t off = 0, n = 0, elem = 0;
    uint64_t bytes() const { return n * elem; }
};

struct CompactVectorView {    // bits::compact_vector, compact_vector.hpp:286-304
    uint64_t size = 0, width = 0, mask = 0;
    Span data;                // u64 words (the builder leaves >= 1 spare word, :102-104)
};

struct BitVectorView {        // bits::bit_vector, bit_vector.hpp:343-352
    uint64_t num_bits = 0;
    Span data;
};

struct EliasFanoView {        // bits::elias_fano<false,false>, elias_fano.hpp:366-379
    uint64_t back = 0;
    BitVectorView high_bits;
    CompactVectorView low_bits;   // the two darrays are skipped: select is not needed (see decode)
};

struct SinglePhfView {        // pthash::single_phf, single_phf.hpp:116-142
    uint64_t seed = 0, num_keys = 0, table_size = 0, num_buckets = 0;
    CompactVectorView pilots;
    EliasFanoView free_slots;
};

struct PartitionedPhfView {   // pthash::partitioned_phf, partitioned_phf.hpp:193-206
    uint64_t seed = 0, num_keys = 0, table_size = 0, num_partitions = 0;
    std::vector<uint64_t> offsets;
    std::vector<SinglePhfView> parts;
};

struct EndpointsView {        // bits::endpoints_sequence, endpoints_sequence.hpp:222-240
    uint64_t back = 0;
    BitVectorView high_bits;
    CompactVectorView hints_0;
    Span low_bits;            // u8
};

struct IndexFile {
    // header, dictionary.hpp:139-162
    uint8_t version[3] = {0, 0, 0};
    uint64_t num_kmers = 0, num_strings = 0;
    uint16_t k = 0, m = 0;
    bool canonical = false;
    uint64_t hasher_magic = 0;                       // mixer_64::m_magic, hash_util.hpp:84-105
    // spectrum_preserving_string_set.hpp:200-211
    EndpointsView endpoints;
    BitVectorView strings;
    // sparse_and_skew_index.hpp:149-167
    PartitionedPhfView minimizers_mphf;              // minimizers_control_map.hpp:49-63
    CompactVectorView control_codewords;
    Span begin_buckets_of_size;                      // u32 x 65
    CompactVectorView mid_load_buckets;
    std::vector<PartitionedPhfView> skew_mphfs;      // skew_index, :60-76
    std::vector<CompactVectorView> skew_positions;
    CompactVectorView heavy_load_buckets;
    uint64_t weights_off = 0, weights_bytes = 0;
    bool weighted = false;
    // weights.hpp:182-187 (populated iff weighted)
    CompactVectorView weight_interval_values;
    EliasFanoView weight_interval_lengths;           // elias_fano<true,false>: same byte layout
    CompactVectorView weight_dictionary;

    // mapping
    const uint8_t* base = nullptr;
    uint64_t file_bytes = 0;
    int fd = -1;

    IndexFile() = default;
    IndexFile(const IndexFile&) = delete;
    IndexFile& operator=(const IndexFile&) = delete;
    ~IndexFile();

    // Returns "" on success, otherwise an error message; *status_out gets an sshash_gpu_status.
    std::string open(const char* path, int* status_out);

    const uint8_t* ptr(const Span& s) const { return base + s.off; }

    // Decoders for the two re-encoded structures (host side, sequential, run once at open).
    // Elias-Fano i-th value = ((position of the i-th set bit of high_bits - i) << l) | low[i]
    // (elias_fano.hpp:181-185); walking the set bits in order needs no select structure.
    void decode_elias_fano(const EliasFanoView& ef, uint64_t n, std::vector<uint32_t>& out) const;
    void decode_elias_fano(const EliasFanoView& ef, uint64_t n, std::vector<uint64_t>& out) const;
    // end-points: same with l = 8 and byte-wide low parts (endpoints_sequence.hpp:160-163)
    void decode_endpoints(std::vector<uint64_t>& out) const;
};

}  // namespace sshash_b200
