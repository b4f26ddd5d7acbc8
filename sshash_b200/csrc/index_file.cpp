// index_file.cpp -- see index_file.hpp.
#include "index_file.hpp"

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <cstring>

#include "../../include/sshash_gpu.h"

namespace sshash_b200 {

namespace {

struct Reader {
    const uint8_t* base;
    uint64_t size;
    uint64_t pos = 0;
    bool fail = false;

    template <typename T>
    T pod() {
        T v{};
        if (pos + sizeof(T) > size) { fail = true; return v; }
        std::memcpy(&v, base + pos, sizeof(T));
        pos += sizeof(T);
        return v;
    }
    Span vec(uint64_t elem) {  // {u64 n; n*elem bytes}, essentials.hpp:346-393
        Span s;
        uint64_t n = pod<uint64_t>();
        if (fail || n > (size - pos) / elem) { fail = true; return s; }
        s.off = pos; s.n = n; s.elem = elem;
        pos += n * elem;
        return s;
    }
    CompactVectorView compact_vector() {
        CompactVectorView c;
        c.size = pod<uint64_t>(); c.width = pod<uint64_t>(); c.mask = pod<uint64_t>();
        c.data = vec(8);
        // an empty compact_vector is {0,0,0,[]}; a populated one needs width in [1,64] and enough words
        if (!fail && c.size != 0) {
            // overflow-safe: data.n <= file_bytes / 8, so data.n * 64 cannot wrap, size * width can
            if (c.width > 64 || (c.width && c.size > c.data.n * 64 / c.width)) fail = true;
        }
        return c;
    }
    BitVectorView bit_vector() {
        BitVectorView b;
        b.num_bits = pod<uint64_t>();
        b.data = vec(8);
        if (!fail && b.data.n * 64 < b.num_bits) fail = true;
        return b;
    }
    void skip_darray() {  // darray.hpp:221-227
        (void)pod<uint64_t>();
        (void)vec(8); (void)vec(2); (void)vec(8);
    }
    EliasFanoView elias_fano() {
        EliasFanoView e;
        e.back = pod<uint64_t>();
        e.high_bits = bit_vector();
        skip_darray();  // d1
        skip_darray();  // d0 (present even when unused)
        e.low_bits = compact_vector();
        return e;
    }
    SinglePhfView single_phf() {
        SinglePhfView f;
        f.seed = pod<uint64_t>(); f.num_keys = pod<uint64_t>(); f.table_size = pod<uint64_t>();
        f.num_buckets = pod<uint64_t>();
        f.pilots = compact_vector();
        f.free_slots = elias_fano();
        if (!fail && f.num_keys != 0) {
            if (f.pilots.size != f.num_buckets || f.table_size < f.num_keys) fail = true;
        }
        return f;
    }
    PartitionedPhfView partitioned_phf() {
        PartitionedPhfView f;
        f.seed = pod<uint64_t>(); f.num_keys = pod<uint64_t>(); f.table_size = pod<uint64_t>();
        f.num_partitions = pod<uint64_t>();
        uint64_t n = pod<uint64_t>();
        if (fail || n > (1u << 24) || n != f.num_partitions) {
            // an empty (default-constructed) MPHF serialises as all zeros with n == 0
            if (!(n == 0 && !fail)) { fail = true; return f; }
        }
        f.offsets.reserve(n); f.parts.reserve(n);
        for (uint64_t i = 0; i != n && !fail; ++i) {
            f.offsets.push_back(pod<uint64_t>());
            f.parts.push_back(single_phf());
        }
        return f;
    }
};

uint64_t load_word(const uint8_t* p, uint64_t i) {
    uint64_t w;
    std::memcpy(&w, p + 8 * i, 8);
    return w;
}

}  // namespace

IndexFile::~IndexFile() {
    if (base) ::munmap(const_cast<uint8_t*>(base), file_bytes);
    if (fd >= 0) ::close(fd);
}

std::string IndexFile::open(const char* path, int* status_out) {
    auto err = [&](int st, std::string msg) { *status_out = st; return msg; };
    fd = ::open(path, O_RDONLY);
    if (fd < 0) return err(SSHASH_GPU_EIO, std::string("cannot open index file '") + path + "'");
    struct stat sb;
    if (fstat(fd, &sb) != 0 || sb.st_size < 40) return err(SSHASH_GPU_EFORMAT, "index file too small");
    file_bytes = static_cast<uint64_t>(sb.st_size);
    void* p = ::mmap(nullptr, file_bytes, PROT_READ, MAP_PRIVATE, fd, 0);
    if (p == MAP_FAILED) return err(SSHASH_GPU_EIO, "mmap of the index file failed");
    base = static_cast<const uint8_t*>(p);
    ::madvise(p, file_bytes, MADV_SEQUENTIAL);

    Reader r{base, file_bytes};
    version[0] = r.pod<uint8_t>(); version[1] = r.pod<uint8_t>(); version[2] = r.pod<uint8_t>();
    if (version[0] != 5)  // util::check_version_number, util.hpp:191-195
        return err(SSHASH_GPU_EVERSION, "MAJOR index version mismatch: SSHash index needs rebuilding");
    num_kmers = r.pod<uint64_t>(); num_strings = r.pod<uint64_t>();
    k = r.pod<uint16_t>(); m = r.pod<uint16_t>();
    canonical = r.pod<uint8_t>() != 0;
    hasher_magic = r.pod<uint64_t>();
    uint16_t k2 = r.pod<uint16_t>(), m2 = r.pod<uint16_t>();
    (void)r.pod<uint64_t>();  // m_num_bits_per_relative_offset: uninitialised for decoded_offsets (offsets.hpp:104-112)
    endpoints.back = r.pod<uint64_t>();
    endpoints.high_bits = r.bit_vector();
    r.skip_darray();
    endpoints.hints_0 = r.compact_vector();
    endpoints.low_bits = r.vec(1);
    strings = r.bit_vector();
    minimizers_mphf = r.partitioned_phf();
    control_codewords = r.compact_vector();
    begin_buckets_of_size = r.vec(4);
    mid_load_buckets = r.compact_vector();
    uint64_t n_mphfs = r.pod<uint64_t>();
    if (r.fail || n_mphfs > 8) return err(SSHASH_GPU_EFORMAT, "malformed index file (skew index)");
    for (uint64_t i = 0; i != n_mphfs; ++i) skew_mphfs.push_back(r.partitioned_phf());
    uint64_t n_pos = r.pod<uint64_t>();
    if (r.fail || n_pos != n_mphfs) return err(SSHASH_GPU_EFORMAT, "malformed index file (skew positions)");
    for (uint64_t i = 0; i != n_pos; ++i) skew_positions.push_back(r.compact_vector());
    heavy_load_buckets = r.compact_vector();
    if (r.fail) return err(SSHASH_GPU_EFORMAT, "malformed index file (truncated or inconsistent sections)");
    weights_off = r.pos;
    weights_bytes = file_bytes - r.pos;
    {   // weights: compact_vector interval_values first (weights.hpp:182-187); non-empty <=> weighted
        Reader w{base, file_bytes, r.pos};
        weight_interval_values = w.compact_vector();
        weight_interval_lengths = w.elias_fano();
        weight_dictionary = w.compact_vector();
        weighted = !w.fail && weight_interval_values.size != 0;
        if (weighted && (weight_interval_lengths.low_bits.size != weight_interval_values.size + 1 ||
                         weight_dictionary.size == 0 || weight_interval_values.width > 57 ||
                         weight_dictionary.width > 57))
            return err(SSHASH_GPU_EFORMAT, "malformed index file (weights)");
    }
    if (k2 != k || m2 != m || k == 0 || m == 0 || m > k || k > 63 || m > 31)
        return err(SSHASH_GPU_EFORMAT, "malformed index file (k/m)");
    if (endpoints.low_bits.n != num_strings + 1 || begin_buckets_of_size.n > 65 ||
        minimizers_mphf.num_keys != control_codewords.size || minimizers_mphf.parts.empty())
        return err(SSHASH_GPU_EFORMAT, "malformed index file (inconsistent sizes)");
    if (control_codewords.width > 57 || mid_load_buckets.width > 57 || heavy_load_buckets.width > 57)
        return err(SSHASH_GPU_EFORMAT, "unsupported index: compact vector wider than 57 bits");
    *status_out = SSHASH_GPU_OK;
    return "";
}

namespace {
template <typename T>
void decode_ef(const IndexFile& f, const EliasFanoView& ef, uint64_t n, std::vector<T>& out) {
    const uint8_t* high = f.ptr(ef.high_bits.data);
    const uint8_t* low = f.ptr(ef.low_bits.data);
    const uint64_t l = ef.low_bits.width, lmask = ef.low_bits.mask;
    const uint64_t nwords = ef.high_bits.data.n;
    if (l && ef.low_bits.size < n) return;   // fewer low parts than values: the caller sees a short result and rejects the file
    uint64_t i = 0;
    for (uint64_t w = 0; w != nwords && i != n; ++w) {
        uint64_t word = load_word(high, w);
        while (word && i != n) {
            uint64_t pos = (w << 6) + static_cast<uint64_t>(__builtin_ctzll(word));
            word &= word - 1;
            uint64_t lo = 0;
            if (l) {
                uint64_t bit = i * l, wi = bit >> 6, sh = bit & 63;
                lo = load_word(low, wi) >> sh;
                if (sh + l > 64) lo |= load_word(low, wi + 1) << (64 - sh);
                lo &= lmask;
            }
            out.push_back(static_cast<T>(((pos - i) << l) | lo));
            ++i;
        }
    }
}
}  // namespace

void IndexFile::decode_elias_fano(const EliasFanoView& ef, uint64_t n, std::vector<uint32_t>& out) const {
    decode_ef(*this, ef, n, out);
}
void IndexFile::decode_elias_fano(const EliasFanoView& ef, uint64_t n, std::vector<uint64_t>& out) const {
    decode_ef(*this, ef, n, out);
}

void IndexFile::decode_endpoints(std::vector<uint64_t>& out) const {
    const uint8_t* high = ptr(endpoints.high_bits.data);
    const uint8_t* low = ptr(endpoints.low_bits);
    const uint64_t n = endpoints.low_bits.n, nwords = endpoints.high_bits.data.n;
    out.clear();
    out.reserve(n + 2);
    uint64_t i = 0;
    for (uint64_t w = 0; w != nwords && i != n; ++w) {
        uint64_t word = load_word(high, w);
        while (word && i != n) {
            uint64_t pos = (w << 6) + static_cast<uint64_t>(__builtin_ctzll(word));
            word &= word - 1;
            out.push_back(((pos - i) << 8) | low[i]);
            ++i;
        }
    }
}

}  // namespace sshash_b200
