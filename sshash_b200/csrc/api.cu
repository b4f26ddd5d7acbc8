// api.cu -- the C ABI declared in include/sshash_gpu.h: index upload, host<->device pipelines.
//
// There is no CPU fallback anywhere in this file: every compute entry point launches the sm_100a
// kernels of kernels.cu or fails with SSHASH_GPU_ECUDA.
#include <cuda_runtime.h>
#include <zlib.h>

#include <fcntl.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <new>
#include <stdexcept>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/sshash_gpu.h"
#include "api_internal.hpp"
#include "index_file.hpp"
#include "kernels.cuh"

using namespace sshash_b200;

#define SSHASH_STR2(x) #x
#define SSHASH_STR(x) SSHASH_STR2(x)

namespace {

thread_local std::string g_last_error;

int fail(int status, const std::string& msg) {
    g_last_error = msg;
    return status;
}
int cuda_fail(cudaError_t e, const char* what) {
    return fail(SSHASH_GPU_ECUDA, std::string(what) + ": " + cudaGetErrorString(e));
}
#define CU(call)                                              \
    do {                                                      \
        cudaError_t e__ = (call);                             \
        if (e__ != cudaSuccess) return cuda_fail(e__, #call); \
    } while (0)

// No exception crosses the C ABI (include/sshash_gpu.h): every entry point runs its body through
// guarded(), which turns std::bad_alloc into SSHASH_GPU_ENOMEM and anything else into EINVAL.
template <typename F>
int guarded(F&& body) noexcept {
    try { return body(); }
    catch (const std::bad_alloc&) { return fail(SSHASH_GPU_ENOMEM, "out of host memory"); }
    catch (const std::exception& e) { return fail(SSHASH_GPU_EINVAL, std::string("internal error: ") + e.what()); }
    catch (...) { return fail(SSHASH_GPU_EINVAL, "internal error: unknown exception"); }
}
#define SSHASH_ENTRY(name, params, args)                                   \
    static int name##_impl params;                                         \
    int name params { return guarded([&]() -> int { return name##_impl args; }); } \
    static int name##_impl params

constexpr uint64_t kPadBytes = 64;   // zero tail after every array (>= 2 words for funnel reads)

// One in-flight chunk of a host<->device pipeline.
struct Slot {
    cudaStream_t stream = nullptr;
    void* d_in = nullptr; uint64_t in_cap = 0;
    void* d_out = nullptr; uint64_t out_cap = 0;
    void* d_tmp = nullptr; uint64_t tmp_cap = 0;
};

// Device scratch owned by one call at a time (taken from / returned to the dictionary's pool).
struct Workspace {
    static constexpr int kSlots = 3;
    Slot slots[kSlots];
    // streaming scratch
    uint64_t* d_win_offsets = nullptr; uint64_t wo_cap = 0;
    uint64_t* d_block_sums = nullptr; uint64_t bs_cap = 0;
    uint64_t* d_win_id = nullptr; uint64_t* d_win_aux = nullptr; uint64_t win_cap = 0;
    uint64_t* d_ids = nullptr; uint64_t ids_cap = 0;
    void* d_anchors = nullptr; uint64_t anchors_cap = 0;
    unsigned long long* d_counters = nullptr;
    unsigned long long* h_counters = nullptr;   // pinned
    // host-buffer streaming pipeline: two staging sets so that the H2D copy of chunk c+1 overlaps the
    // kernels of chunk c (the window scratch above is shared: kernels stay on one stream)
    struct StreamSlot {
        void* d_bases = nullptr; uint64_t bases_cap = 0;
        uint64_t* d_ro = nullptr; uint64_t ro_cap = 0;
        uint64_t* d_ids = nullptr; uint64_t ids_cap = 0;
        cudaEvent_t ready = nullptr, free = nullptr;
    } sslots[2];
    cudaStream_t copy_stream = nullptr;
    // file driver with device-side record parsing
    uint8_t* d_raw[2] = {nullptr, nullptr}; uint64_t raw_cap[2] = {0, 0};          // two chunks in flight: chunk c+1 is copied
    uint64_t* d_tiles[2] = {nullptr, nullptr}; uint64_t tiles_cap[2] = {0, 0};    // and line-counted while chunk c is streamed
    cudaEvent_t file_counted[2] = {nullptr, nullptr}, file_done[2] = {nullptr, nullptr};
    uint64_t* d_line_start = nullptr; uint64_t ls_cap = 0;
    uint64_t* d_spans = nullptr; uint64_t spans_cap = 0;
    uint8_t* h_file[2] = {nullptr, nullptr}; uint64_t h_file_cap = 0;   // pinned
    // partition-major lookups (binned.cu)
    uint8_t* d_bin = nullptr; uint64_t bin_cap = 0;

    ~Workspace() {
        for (auto& s : slots) {
            if (s.d_in) cudaFree(s.d_in);
            if (s.d_out) cudaFree(s.d_out);
            if (s.d_tmp) cudaFree(s.d_tmp);
            if (s.stream) cudaStreamDestroy(s.stream);
        }
        cudaFree(d_win_offsets); cudaFree(d_block_sums);
        cudaFree(d_win_id); cudaFree(d_win_aux); cudaFree(d_ids); cudaFree(d_counters); cudaFree(d_anchors);
        if (h_counters) cudaFreeHost(h_counters);
        for (int i = 0; i < 2; ++i) {
            cudaFree(d_raw[i]); cudaFree(d_tiles[i]);
            if (file_counted[i]) cudaEventDestroy(file_counted[i]);
            if (file_done[i]) cudaEventDestroy(file_done[i]);
        }
        cudaFree(d_line_start); cudaFree(d_spans); cudaFree(d_bin);
        for (auto& ss : sslots) {
            cudaFree(ss.d_bases); cudaFree(ss.d_ro); cudaFree(ss.d_ids);
            if (ss.ready) cudaEventDestroy(ss.ready);
            if (ss.free) cudaEventDestroy(ss.free);
        }
        if (copy_stream) cudaStreamDestroy(copy_stream);
        for (auto* h : h_file) if (h) cudaFreeHost(h);
    }
};

template <typename T>
cudaError_t ensure(T*& p, uint64_t& cap, uint64_t need_bytes) {
    if (need_bytes <= cap && p) return cudaSuccess;
    if (p) cudaFree(p);
    p = nullptr; cap = 0;
    uint64_t bytes = need_bytes + need_bytes / 8 + 256;
    cudaError_t e = cudaMalloc(reinterpret_cast<void**>(&p), bytes);
    if (e == cudaSuccess) cap = bytes;
    return e;
}

uint64_t env_bytes(const char* name, uint64_t dflt) {
    const char* e = std::getenv(name);
    if (!e || !*e) return dflt;
    const unsigned long long v = std::strtoull(e, nullptr, 10);
    return v ? v : dflt;
}

bool is_device_pointer(const void* p) {
    if (!p) return false;
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}

// device memory that lives on `device` itself (or managed memory): kernels use it in place.  Memory
// of ANOTHER GPU is treated like host memory by the batched pipeline: staged chunk by chunk with
// copy-engine transfers (cudaMemcpyDefault: peer-to-peer over NVLink when peer access is enabled).
bool is_local_device_pointer(const void* p, int device) {
    if (!p) return false;
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeManaged || (a.type == cudaMemoryTypeDevice && a.device == device);
}

}  // namespace

int sshash_b200::set_last_error(int status, const std::string& msg) { return fail(status, msg); }

struct sshash_gpu_dict {
    int device = 0;
    LaunchCtx ctx{};
    DeviceIndex ix{};
    sshash_gpu_info_t info{};
    std::vector<void*> allocs;
    cudaStream_t stream = nullptr;
    std::mutex mu;
    std::vector<std::unique_ptr<Workspace>> pool;
    std::atomic<int> peer_inplace{0};   // 1: device pointers of OTHER GPUs are dereferenced by the kernels (peer access / symmetric memory)
    std::mutex contract_mu;
    int breaks_contract = -1;   // -1 not checked yet; 1 = duplicated k-mers / rc twins: streaming replays the state machine

    std::unique_ptr<Workspace> take() {
        std::lock_guard<std::mutex> g(mu);
        if (pool.empty()) return std::make_unique<Workspace>();
        auto w = std::move(pool.back());
        pool.pop_back();
        return w;
    }
    void give(std::unique_ptr<Workspace> w) {
        std::lock_guard<std::mutex> g(mu);
        pool.push_back(std::move(w));
    }
    ~sshash_gpu_dict() {
        cudaSetDevice(device);
        pool.clear();
        for (void* p : allocs) cudaFree(p);
        if (stream) cudaStreamDestroy(stream);
    }
};

namespace {

struct WorkspaceLease {
    sshash_gpu_dict* d;
    std::unique_ptr<Workspace> w;
    explicit WorkspaceLease(const sshash_gpu_dict* dict) : d(const_cast<sshash_gpu_dict*>(dict)), w(d->take()) {}
    ~WorkspaceLease() { d->give(std::move(w)); }
    Workspace* operator->() { return w.get(); }
};

// ---- index upload ------------------------------------------------------------------------------
struct Uploader {
    sshash_gpu_dict* d;
    uint64_t bytes = 0;
    std::string error;

    void* raw(const void* src, uint64_t n_bytes) {
        void* p = nullptr;
        cudaError_t e = cudaMalloc(&p, n_bytes + kPadBytes);
        if (e != cudaSuccess) { error = std::string("cudaMalloc: ") + cudaGetErrorString(e); return nullptr; }
        d->allocs.push_back(p);
        bytes += n_bytes + kPadBytes;
        e = cudaMemset(static_cast<uint8_t*>(p) + n_bytes, 0, kPadBytes);
        if (e == cudaSuccess && n_bytes) e = cudaMemcpy(p, src, n_bytes, cudaMemcpyHostToDevice);
        if (e != cudaSuccess) { error = std::string("cudaMemcpy: ") + cudaGetErrorString(e); return nullptr; }
        return p;
    }
    DevCompact compact(const IndexFile& f, const CompactVectorView& c) {
        DevCompact o{};
        o.data = static_cast<const uint64_t*>(raw(f.ptr(c.data), c.data.bytes()));
        o.size = c.size; o.mask = c.mask; o.width = (uint32_t)c.width;
        return o;
    }
};

uint64_t shift_mix(uint64_t v) { return v ^ (v >> 47); }

int upload_index(sshash_gpu_dict* d, const IndexFile& f, int max_k) {
    Uploader up{d, 0, {}};
    DeviceIndex& ix = d->ix;
    ix.k = f.k; ix.m = f.m; ix.canonical = f.canonical ? 1 : 0;
    ix.kmer_words = max_k == 31 ? 1 : 2;
    ix.magic = f.hasher_magic;
    for (uint32_t j = 0; j < 16; ++j) {
        const uint32_t hi = static_cast<uint32_t>(f.hasher_magic >> 32) & ~63u;
        ix.mini_left[j] = hi | j;
        ix.mini_right[j] = hi | (63u - j);
    }
    ix.kmer_mask_lo = 2 * f.k >= 64 ? ~0ull : (1ull << (2 * f.k)) - 1;
    ix.kmer_mask_hi = 2 * f.k <= 64 ? 0ull : (1ull << (2 * f.k - 64)) - 1;
    ix.mmer_mask = 2 * f.m >= 64 ? ~0ull : (1ull << (2 * f.m)) - 1;
    ix.num_kmers = f.num_kmers; ix.num_strings = f.num_strings;
    ix.strings = static_cast<const uint64_t*>(up.raw(f.ptr(f.strings.data), f.strings.data.bytes()));
    ix.strings_bits = f.strings.num_bits;

    // MPHFs: one pilots pool + one decoded free-slot pool + one table of partitions
    std::vector<const PartitionedPhfView*> phfs;
    phfs.push_back(&f.minimizers_mphf);
    for (auto const& s : f.skew_mphfs) phfs.push_back(&s);
    uint64_t pilot_words = 0, n_parts = 0;
    for (auto* p : phfs) for (auto const& sp : p->parts) { pilot_words += sp.pilots.data.n; ++n_parts; }
    std::vector<DevPhfPart> parts; parts.reserve(n_parts);
    std::vector<uint32_t> free_pool;
    // host staging of the pilots pool (copied into the hot slab below)
    std::vector<uint64_t> pilots_host(pilot_words);
    uint64_t word = 0;
    std::vector<uint64_t> first_part;
    for (size_t pi = 0; pi != phfs.size(); ++pi) {
        const PartitionedPhfView* p = phfs[pi];
        first_part.push_back(parts.size());
        for (size_t i = 0; i != p->parts.size(); ++i) {
            auto const& sp = p->parts[i];
            if ((sp.num_keys | sp.table_size | sp.num_buckets) >> 32)
                return fail(SSHASH_GPU_EFORMAT, "unsupported index: MPHF partition with >= 2^32 keys / slots / buckets");
            if ((word | free_pool.size()) >> 32) return fail(SSHASH_GPU_EFORMAT, "unsupported index: MPHF pools exceed 2^32 entries");
            if (sp.pilots.width > 64 || sp.table_size < sp.num_keys) return fail(SSHASH_GPU_EFORMAT, "malformed index file (MPHF partition)");
            DevPhfPart o{};
            o.offset = p->offsets[i];
            o.num_keys = (uint32_t)sp.num_keys; o.table_size = (uint32_t)sp.table_size; o.num_buckets = (uint32_t)sp.num_buckets;
            o.pilots_word = (uint32_t)word; o.pilot_width = (uint32_t)sp.pilots.width;
            o.free_off = (uint32_t)free_pool.size();
            if (sp.pilots.data.n) std::memcpy(pilots_host.data() + word, f.ptr(sp.pilots.data), sp.pilots.data.bytes());
            word += sp.pilots.data.n;
            const uint64_t n_free = sp.table_size - sp.num_keys;
            if (n_free) {
                size_t before = free_pool.size();
                f.decode_elias_fano(sp.free_slots, n_free, free_pool);
                if (free_pool.size() - before != n_free)
                    return fail(SSHASH_GPU_EFORMAT, "malformed index file (free slots)");
                for (size_t j = before; j != free_pool.size(); ++j)      // a free slot maps to a position < num_keys
                    if (free_pool[j] >= sp.num_keys) return fail(SSHASH_GPU_EFORMAT, "malformed index file (free slot out of range)");
            }
            {   // positions of this partition index the codewords (minimizer MPHF) / the skew positions
                const uint64_t limit = pi == 0 ? f.control_codewords.size : f.skew_positions[pi - 1].size;
                if (o.offset > limit || sp.num_keys > limit - o.offset)
                    return fail(SSHASH_GPU_EFORMAT, "malformed index file (MPHF partition offset)");
            }
            parts.push_back(o);
        }
    }
    auto make_phf = [&](const PartitionedPhfView& p, uint64_t first) {
        DevPhf o{};
        const uint64_t k1 = 0xb492b66fbe98f273ull;
        o.seed_hi = ~p.seed;
        o.city_a = shift_mix(p.seed * k1) * k1;     // cityhash.cpp:245
        o.city_cb = (~p.seed) * k1;                 // cityhash.cpp:246
        o.num_partitions = p.parts.size();
        o.first_part_ = first;
        return o;
    };
    ix.mphf = make_phf(f.minimizers_mphf, first_part[0]);
    ix.n_skew = (uint32_t)f.skew_mphfs.size();
    for (uint32_t i = 0; i != ix.n_skew; ++i) {
        ix.skew[i] = make_phf(f.skew_mphfs[i], first_part[1 + i]);
        ix.skew_pos[i] = up.compact(f, f.skew_positions[i]);
    }
    ix.codewords = up.compact(f, f.control_codewords);
    ix.mid_load = up.compact(f, f.mid_load_buckets);
    ix.heavy = up.compact(f, f.heavy_load_buckets);
    std::memset(ix.begin_buckets_of_size, 0, sizeof(ix.begin_buckets_of_size));
    std::memcpy(ix.begin_buckets_of_size, f.ptr(f.begin_buckets_of_size), f.begin_buckets_of_size.bytes());

    // string end-points: decoded + a sampled directory
    std::vector<uint64_t> ends;
    f.decode_endpoints(ends);
    if (ends.size() != f.num_strings + 1 || ends.empty() || ends[0] != 0)
        return fail(SSHASH_GPU_EFORMAT, "malformed index file (string end-points)");
    for (size_t i = 1; i < ends.size(); ++i)
        if (ends[i] <= ends[i - 1]) return fail(SSHASH_GPU_EFORMAT, "malformed index file (end-points not increasing)");
    if (ends.size() >> 32) return fail(SSHASH_GPU_EFORMAT, "unsupported index: >= 2^32 strings");
    const uint64_t U = ends.back();
    if (2 * U > f.strings.num_bits) return fail(SSHASH_GPU_EFORMAT, "malformed index file (end-points beyond the strings)");
    // Directory granularity: about two blocks per string (a block then holds 0.5 end-points on average, so
    // locate_string's scan takes half a step), but never fewer than 2^20 blocks while blocks stay >= 2^6
    // bases: a small index gets a fine directory (4 MB at most, no scan steps at all), a huge one a compact one
    // (2.5e9 k-mers: 60 MB -> 30 MB of locate tables).  SSHASH_GPU_LOCATE=legacy keeps round 1's fixed 2^8
    // blocks and 64-bit end-points (A/B switch).
    const char* loc_env = std::getenv("SSHASH_GPU_LOCATE");
    const bool legacy_locate = loc_env && std::strcmp(loc_env, "legacy") == 0;
    ix.dir_shift = 8;
    if (!legacy_locate) {
        const uint64_t max_blocks = std::max<uint64_t>(2 * ends.size(), 1ull << 20);
        ix.dir_shift = 6;
        while (ix.dir_shift < 16 && (U >> ix.dir_shift) > max_blocks) ++ix.dir_shift;
    }
    std::vector<uint32_t> dir(((f.strings.num_bits / 2) >> ix.dir_shift) + 2);   // every offset the validated buckets can hold
    {   // dir[h] = index of the last end-point < (h << shift), 0 if none
        uint64_t j = 0;
        for (uint64_t h = 0; h != dir.size(); ++h) {
            const uint64_t lim = h << ix.dir_shift;
            while (j + 1 < ends.size() && ends[j + 1] < lim) ++j;
            dir[h] = (uint32_t)j;
        }
    }
    std::vector<uint32_t> ends32;
    if (!legacy_locate && U < 0xffffffffull) {
        ends32.resize(ends.size() + 2);
        for (size_t i = 0; i != ends.size(); ++i) ends32[i] = (uint32_t)ends[i];
        ends32[ends.size()] = ends32[ends.size() + 1] = 0xffffffffu;   // scan sentinels (> any offset)
    }
    ix.n_ends = ends.size();
    ends.push_back(~0ull); ends.push_back(~0ull);   // scan sentinels

    // weights (include/weights.hpp:182-187): interval starts decoded from their Elias-Fano sequence +
    // a sampled directory; both tables go into the hot slab, the two compact vectors are verbatim
    std::vector<uint64_t> wstarts;
    std::vector<uint32_t> wdir;
    ix.n_weight_intervals = 0;
    ix.weight_dir_shift = 0;
    if (f.weighted) {
        const uint64_t nw = f.weight_interval_values.size;
        f.decode_elias_fano(f.weight_interval_lengths, nw + 1, wstarts);
        if (wstarts.size() != nw + 1 || wstarts[0] != 0 || wstarts.back() != f.num_kmers || (nw >> 32))
            return fail(SSHASH_GPU_EFORMAT, "malformed index file (weight intervals)");
        for (uint64_t i = 1; i <= nw; ++i)
            if (wstarts[i] <= wstarts[i - 1]) return fail(SSHASH_GPU_EFORMAT, "malformed index file (weight intervals not increasing)");
        uint32_t shift = 0;
        while ((f.num_kmers >> shift) > 2 * nw + 64) ++shift;
        wdir.resize((f.num_kmers >> shift) + 3);
        uint64_t j = 0;
        for (uint64_t h = 0; h != wdir.size(); ++h) {      // index of the last start <= (h << shift), interval starts only
            const uint64_t lim = h << shift;
            while (j + 1 < nw && wstarts[j + 1] <= lim) ++j;
            wdir[h] = (uint32_t)j;
        }
        ix.n_weight_intervals = nw;
        ix.weight_dir_shift = shift;
        ix.weight_values = up.compact(f, f.weight_interval_values);
        ix.weight_dict = up.compact(f, f.weight_dictionary);
    }

    // HOT SLAB: the arrays every lookup touches (pilots, free slots, partition table, end-points and
    // their directory) live in ONE allocation so that a single L2 access-policy window can keep
    // them persistent in L2 while the cold, much larger arrays (codewords, strings) stream through.
    // Order: the locate tables first, the pilots pool last -- when the slab is larger than the part of L2
    // that can be set aside (human-scale indexes: > 100 MB of pilots), the persisting window covers
    // the PREFIX that fits (configure_l2) and the pilots take the cold load policy instead.
    struct Piece { const void* src; uint64_t bytes; uint64_t off; };
    Piece pieces[8] = {{ends32.data(), ends32.size() * 4, 0}, {dir.data(), dir.size() * 4, 0},
                       {parts.data(), parts.size() * sizeof(DevPhfPart), 0}, {free_pool.data(), free_pool.size() * 4, 0},
                       {ends.data(), ends.size() * 8, 0}, {wstarts.data(), wstarts.size() * 8, 0},
                       {wdir.data(), wdir.size() * 4, 0}, {pilots_host.data(), pilots_host.size() * 8, 0}};
    uint64_t slab_bytes = 0;
    for (auto& p : pieces) { p.off = slab_bytes; slab_bytes += (p.bytes + kPadBytes + 255) & ~255ull; }
    uint8_t* slab = nullptr;
    {
        cudaError_t e = cudaMalloc(reinterpret_cast<void**>(&slab), slab_bytes);
        if (e != cudaSuccess) return cuda_fail(e, "cudaMalloc(hot slab)");
        d->allocs.push_back(slab);
        up.bytes += slab_bytes;
        CU(cudaMemset(slab, 0, slab_bytes));
        for (auto& p : pieces) if (p.bytes) CU(cudaMemcpy(slab + p.off, p.src, p.bytes, cudaMemcpyHostToDevice));
    }
    ix.ends32 = ends32.empty() ? nullptr : reinterpret_cast<const uint32_t*>(slab + pieces[0].off);
    ix.ends_dir = reinterpret_cast<const uint32_t*>(slab + pieces[1].off);
    const DevPhfPart* d_parts = reinterpret_cast<const DevPhfPart*>(slab + pieces[2].off);
    ix.free_slots = reinterpret_cast<const uint32_t*>(slab + pieces[3].off);
    ix.ends = reinterpret_cast<const uint64_t*>(slab + pieces[4].off);
    ix.weight_starts = reinterpret_cast<const uint64_t*>(slab + pieces[5].off);
    ix.weight_dir = reinterpret_cast<const uint32_t*>(slab + pieces[6].off);
    ix.pilots = reinterpret_cast<const uint64_t*>(slab + pieces[7].off);
    ix.mphf.parts = d_parts + ix.mphf.first_part_;
    for (uint32_t i = 0; i != ix.n_skew; ++i) ix.skew[i].parts = d_parts + ix.skew[i].first_part_;
    d->ctx.hot_prefix_bytes = pieces[7].off;     // everything but the pilots
    d->ctx.hot_base = slab;
    d->ctx.hot_bytes = slab_bytes;
    if (!up.error.empty()) return fail(SSHASH_GPU_ECUDA, up.error);

    {   // reject files whose codewords / bucket offsets point outside their arrays (kernels.cu)
        uint32_t* d_flag = nullptr;
        CU(cudaMalloc(reinterpret_cast<void**>(&d_flag), 4));
        cudaError_t e = cudaMemset(d_flag, 0, 4);
        if (e == cudaSuccess) e = launch_validate_index(ix, d->ctx, d_flag, nullptr);
        uint32_t flag = 0;
        if (e == cudaSuccess) e = cudaMemcpy(&flag, d_flag, 4, cudaMemcpyDeviceToHost);
        cudaFree(d_flag);
        if (e != cudaSuccess) return cuda_fail(e, "index validation");
        if (flag) return fail(SSHASH_GPU_EFORMAT, "malformed index file (bucket reference out of range, mask " + std::to_string(flag) + ")");
    }
    // FINGERPRINTED CODEWORDS: the control codewords are re-encoded on the device as a compact vector
    // of width w + f whose entries carry, above the reference's w-bit codeword, an f-bit fingerprint
    // of the minimizer that owns the slot (read back from `strings` through the bucket).  A minimizer
    // that is not in the index hashes to an arbitrary slot; the ids-only lookup rejects it right
    // after the codeword read, without the strings read the reference needs to find out.  The
    // verbatim copy is dropped afterwards.  SSHASH_GPU_FP=0 keeps the verbatim vector.
    ix.cw_code_bits = ix.codewords.width;
    ix.cw_fp_bits = 0;
    const char* fp_env = std::getenv("SSHASH_GPU_FP");
    if (!(fp_env && fp_env[0] == '0') && ix.codewords.size && ix.codewords.width + 8 <= 57) {
        const uint32_t w = ix.codewords.width;
        const uint32_t fpb = w + 8 <= 32 ? 32 - w : 8;            // round the entry up to 32 bits when that is cheap
        const uint64_t words = (ix.codewords.size * (uint64_t)(w + fpb) + 63) / 64;
        uint64_t* fresh = nullptr;
        cudaError_t e = cudaMalloc(reinterpret_cast<void**>(&fresh), words * 8 + kPadBytes);
        if (e != cudaSuccess) return cuda_fail(e, "cudaMalloc(fingerprinted codewords)");
        CU(cudaMemset(fresh, 0, words * 8 + kPadBytes));
        // minimizer filter: 16 bits per minimizer rounded up to a power of two of 32-bit words, only
        // while it stays a few MB (L2-resident next to the hot slab); SSHASH_GPU_FILTER=0 disables it
        uint32_t* filter = nullptr;
        uint32_t filter_shift = 0;
        const char* fl_env = std::getenv("SSHASH_GPU_FILTER");
        if (!(fl_env && fl_env[0] == '0') && ix.codewords.size <= (4ull << 20)) {
            uint32_t log_words = 4;
            while ((32ull << log_words) < 16 * ix.codewords.size) ++log_words;
            CU(cudaMalloc(reinterpret_cast<void**>(&filter), (4ull << log_words) + kPadBytes));
            d->allocs.push_back(filter);
            up.bytes += (4ull << log_words) + kPadBytes;
            CU(cudaMemset(filter, 0, (4ull << log_words) + kPadBytes));
            filter_shift = 32 - log_words;
        }
        CU(launch_build_fingerprints(ix, d->ctx, fpb, fresh, filter, filter_shift, nullptr));
        ix.minimizer_filter = filter;
        ix.filter_shift = filter_shift;
        CU(cudaDeviceSynchronize());
        void* old = const_cast<uint64_t*>(ix.codewords.data);
        d->allocs.erase(std::find(d->allocs.begin(), d->allocs.end(), old));
        cudaFree(old);
        d->allocs.push_back(fresh);
        up.bytes += words * 8 - f.control_codewords.data.bytes();
        ix.codewords.data = fresh;
        ix.codewords.width = w + fpb;
        ix.codewords.mask = (w + fpb) >= 64 ? ~0ull : ((1ull << (w + fpb)) - 1);
        ix.cw_fp_bits = fpb;
    }
    // WIDE ENTRIES (DeviceIndex::wide): opt-in with SSHASH_GPU_WIDE=1.  Measured on the 5e8-k-mer index
    // (same box): forward-only positives 22.8 -> 26.5 G lookups/s, 50 % RC mix 17.6 -> 18.4, uniform
    // random negatives 20.1 -> 17.3 (the 1 GB entry table loses the partial L2 residency the 290 MB
    // codeword vector has), at 3x the device bytes: a win for positive-heavy batches only.
    {
        const char* we = std::getenv("SSHASH_GPU_WIDE");
        const bool want = we && we[0] == '1';
        const uint64_t text_bits = 2ull * (2ull * f.k - f.m);
        if (want && up.error.empty() && ix.kmer_words == 1 && ix.codewords.size && f.k >= f.m && ix.cw_code_bits >= 1 &&
            ix.cw_code_bits + text_bits <= 128) {
            void* wide = nullptr;
            const uint64_t bytes = ix.codewords.size * 16 + kPadBytes;
            cudaError_t e = cudaMalloc(&wide, bytes);
            if (e != cudaSuccess) return cuda_fail(e, "cudaMalloc(wide entries)");
            d->allocs.push_back(wide);
            up.bytes += bytes;
            CU(cudaMemset(static_cast<uint8_t*>(wide) + ix.codewords.size * 16, 0, kPadBytes));
            CU(launch_build_wide(ix, d->ctx, wide, nullptr));
            CU(cudaDeviceSynchronize());
            ix.wide = static_cast<const ulonglong2*>(wide);
        }
    }
    {   // PARTITION-MAJOR PATH (binned.cu): per-bin byte ranges of the pilots pool and of the codeword vector
        BinPlan& bp = d->ctx.bins;
        const uint64_t P = f.minimizers_mphf.parts.size();
        bp.bin_shift = 0;
        while ((P >> bp.bin_shift) > 1024) ++bp.bin_shift;
        bp.n_bins = (uint32_t)((P + (1ull << bp.bin_shift) - 1) >> bp.bin_shift);
        std::vector<BinRegion> regions(bp.n_bins);
        const uint64_t cw_total = (ix.codewords.size * (uint64_t)ix.codewords.width + 7) / 8;
        uint64_t pilot_bytes_total = 0;
        for (uint32_t b = 0; b != bp.n_bins; ++b) {
            const uint64_t p0 = (uint64_t)b << bp.bin_shift, p1 = std::min<uint64_t>(P, p0 + (1ull << bp.bin_shift)) - 1;
            const DevPhfPart& a = parts[first_part[0] + p0];
            const DevPhfPart& z = parts[first_part[0] + p1];
            BinRegion r{};
            r.pilots_off = ((uint64_t)a.pilots_word * 8) & ~15ull;     // bulk prefetches want 16-byte alignment
            r.pilots_bytes = ((uint64_t)z.pilots_word + f.minimizers_mphf.parts[p1].pilots.data.n) * 8 - r.pilots_off;
            const uint64_t lo = (a.offset * ix.codewords.width / 8) & ~127ull;
            const uint64_t hi = std::min<uint64_t>(cw_total, (((z.offset + z.num_keys) * ix.codewords.width + 7) / 8 + 127) & ~127ull);
            r.cw_off = lo;
            r.cw_bytes = hi > lo ? hi - lo : 0;
            regions[b] = r;
            pilot_bytes_total += r.pilots_bytes;
        }
        void* d_regions = up.raw(regions.data(), regions.size() * sizeof(BinRegion));
        if (!d_regions) return fail(SSHASH_GPU_ECUDA, up.error);
        bp.regions = static_cast<const BinRegion*>(d_regions);
        // Off by default: measured on 5e8-, 2.5e9- and 3e9-k-mer indexes (profiles/r2_binned_ab_v4.jsonl)
        // the exact reordering around the bin-ordered lookups costs more than the L2 hits save.
        // SSHASH_GPU_BINNED=1 turns it on, SSHASH_GPU_BINNED_MIN sets the smallest batch that takes it.
        (void)pilot_bytes_total;
        bp.enabled = false;
        if (const char* be = std::getenv("SSHASH_GPU_BINNED")) bp.enabled = be[0] == '1';
        bp.min_queries = env_bytes("SSHASH_GPU_BINNED_MIN", 1ull << 22);
        if (const char* pe = std::getenv("SSHASH_GPU_BIN_PREFETCH")) bp.prefetch = pe[0] != '0';
        bp.lookahead = (uint32_t)env_bytes("SSHASH_GPU_BIN_LOOKAHEAD", 1);
    }
    if (!up.error.empty()) return fail(SSHASH_GPU_ECUDA, up.error);
    d->info.device_bytes = up.bytes;
    return SSHASH_GPU_OK;
}

// L2 residency of the hot slab: reserve persisting L2 for it and describe the window that every
// kernel launch carries as a launch attribute (kernels.cu).  SSHASH_GPU_L2_PERSIST=0 disables it;
// SSHASH_GPU_L2_FETCH={32,64,128} additionally sets the device's L2 fetch granularity hint.
void configure_l2(sshash_gpu_dict* d) {
    LaunchCtx& c = d->ctx;
    const char* e = std::getenv("SSHASH_GPU_L2_PERSIST");
    const bool want = !(e && e[0] == '0');
    c.window_bytes = 0;
    // Slabs beyond the persisting capacity (human-scale pilots).  Measured on the 2.5e9-k-mer index (profiles/r2_exp_locality_v1.jsonl): a partially resident pilots pool
    // (evict_last, window over the whole slab with hitRatio = persisting / window) beats a cold-policy pool
    // (13.5 vs 12.8 G lookups/s forward, 13.3 vs 10.7 negative) and evict_last with 64-byte fills makes no
    // difference, so the pilots keep the hot policy at every index size (the A/B switches are gone).
    // SSHASH_GPU_L2_WINDOW=prefix: the window covers only the locate tables (the pilots keep their evict_last hint)
    const char* we = std::getenv("SSHASH_GPU_L2_WINDOW");
    const bool prefix_only = we && std::strcmp(we, "prefix") == 0;
    const uint64_t span = prefix_only ? std::max<uint64_t>(c.hot_prefix_bytes, 256) : c.hot_bytes;
    if (want && c.max_window_bytes && c.max_persist_bytes && span) {
        size_t cur = 0;
        cudaDeviceGetLimit(&cur, cudaLimitPersistingL2CacheSize);
        const uint64_t need = std::min<uint64_t>(span, c.max_persist_bytes);
        if (cur < need && cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, need) != cudaSuccess) cudaGetLastError();
        cudaDeviceGetLimit(&cur, cudaLimitPersistingL2CacheSize);
        if (cur) {
            c.window_bytes = std::min<uint64_t>(span, c.max_window_bytes);
            c.hit_ratio = c.window_bytes <= cur ? 1.0f : (float)((double)cur / (double)c.window_bytes);
        }
    }
    {   // window of the partition-major path: the slab's prefix (locate tables, free slots), never the pilots
        size_t cur = 0;
        cudaDeviceGetLimit(&cur, cudaLimitPersistingL2CacheSize);
        const uint64_t prefix = std::min<uint64_t>(c.hot_prefix_bytes, c.max_window_bytes);
        c.bins.window_bytes = (want && cur && prefix) ? prefix : 0;
        c.bins.hit_ratio = prefix <= cur ? 1.0f : (float)((double)cur / (double)prefix);
    }
    if (const char* g = std::getenv("SSHASH_GPU_L2_FETCH")) {
        size_t v = (size_t)std::atoi(g);
        if (v == 32 || v == 64 || v == 128)
            if (cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, v) != cudaSuccess) cudaGetLastError();
    }
}

int check_dict(const sshash_gpu_dict* d) {
    if (!d) return fail(SSHASH_GPU_EINVAL, "null dictionary handle");
    cudaError_t e = cudaSetDevice(d->device);
    if (e != cudaSuccess) return cuda_fail(e, "cudaSetDevice");
    return SSHASH_GPU_OK;
}

cudaError_t ensure_slot(Slot& s, uint64_t in_bytes, uint64_t out_bytes) {
    cudaError_t e;
    if (!s.stream) { e = cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking); if (e != cudaSuccess) return e; }
    if (in_bytes) { e = ensure(s.d_in, s.in_cap, in_bytes); if (e != cudaSuccess) return e; }
    if (out_bytes) { e = ensure(s.d_out, s.out_cap, out_bytes); if (e != cudaSuccess) return e; }
    return cudaSuccess;
}

// Generic batched pipeline for the per-query kernels (lookup / membership / access).
//   in: n elements of in_elem bytes, out: n elements of out_elem bytes (either side host or device).
// All-device: one asynchronous launch on the caller's stream.  Otherwise the batch is cut into
// chunks that cycle through three slots (stream + staging buffers) so that the H2D copy of chunk
// c+1, the kernel of chunk c and the D2H copy of chunk c-1 overlap.
template <typename Launch>
int run_batched(const sshash_gpu_dict* d, const void* in, uint64_t in_elem, void* out, uint64_t out_elem, uint64_t n,
                void* user_stream, Launch launch) {
    if (n == 0) return SSHASH_GPU_OK;
    const bool any_dev = d->peer_inplace.load(std::memory_order_relaxed) != 0;
    const bool in_dev = any_dev ? is_device_pointer(in) : is_local_device_pointer(in, d->device);
    const bool out_dev = any_dev ? is_device_pointer(out) : is_local_device_pointer(out, d->device);
    if (in_dev && out_dev) {
        cudaStream_t s = static_cast<cudaStream_t>(user_stream);   // NULL = the legacy default stream
        CU(launch(in, out, n, s));
        return SSHASH_GPU_OK;
    }
    WorkspaceLease ws(d);
    // SSHASH_GPU_BATCH_CHUNK: bytes per pipeline chunk of the larger of the two element streams
    static const uint64_t chunk_bytes = env_bytes("SSHASH_GPU_BATCH_CHUNK", 32ull << 20);   // read once
    const uint64_t chunk = std::max<uint64_t>(1, chunk_bytes / std::max(in_elem, out_elem));
    int c = 0;
    for (uint64_t off = 0; off < n; off += chunk, ++c) {
        const uint64_t cn = std::min(chunk, n - off);
        Slot& s = ws->slots[c % Workspace::kSlots];
        CU(ensure_slot(s, in_dev ? 0 : chunk * in_elem, out_dev ? 0 : chunk * out_elem));
        const void* din = static_cast<const uint8_t*>(in) + off * in_elem;
        void* dout = static_cast<uint8_t*>(out) + off * out_elem;
        if (!in_dev) { CU(cudaMemcpyAsync(s.d_in, din, cn * in_elem, cudaMemcpyDefault, s.stream)); din = s.d_in; }
        void* kout = out_dev ? dout : s.d_out;
        CU(launch(din, kout, cn, s.stream));
        if (!out_dev) CU(cudaMemcpyAsync(dout, s.d_out, cn * out_elem, cudaMemcpyDefault, s.stream));
    }
    for (auto& s : ws->slots) if (s.stream) CU(cudaStreamSynchronize(s.stream));
    return SSHASH_GPU_OK;
}

}  // namespace

extern "C" {

const char* sshash_gpu_last_error(void) { return g_last_error.c_str(); }

const char* sshash_gpu_build_info(void) {
    return "sshash_b200 CUDA " SSHASH_STR(CUDART_VERSION) " sm_100a";
}

uint64_t sshash_gpu_launch_count(void) { return kernel_launch_count(); }

SSHASH_ENTRY(sshash_gpu_open, (const char* index_path, int device, int max_k, sshash_gpu_dict** out), (index_path, device, max_k, out)) {
    if (!index_path || !out) return fail(SSHASH_GPU_EINVAL, "null argument");
    *out = nullptr;
    if (max_k != 0 && max_k != 31 && max_k != 63) return fail(SSHASH_GPU_EINVAL, "max_k must be 0, 31 or 63");
    IndexFile f;
    int st = SSHASH_GPU_OK;
    std::string msg = f.open(index_path, &st);
    if (st != SSHASH_GPU_OK) return fail(st, msg);
    if (max_k == 0) max_k = f.k <= 31 ? 31 : 63;
    if (f.k > (uint32_t)max_k) return fail(SSHASH_GPU_EINVAL, "k = " + std::to_string(f.k) + " needs max_k = 63");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return fail(SSHASH_GPU_ECUDA, "no CUDA device available (this library has no CPU fallback)");
    }
    if (device < 0 || device >= ndev) return fail(SSHASH_GPU_EINVAL, "invalid device ordinal");
    CU(cudaSetDevice(device));
    auto d = std::make_unique<sshash_gpu_dict>();
    d->device = device;
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, device));
    d->ctx.sm_count = prop.multiProcessorCount;
    d->ctx.max_window_bytes = (uint64_t)std::max(prop.accessPolicyMaxWindowSize, 0);
    d->ctx.max_persist_bytes = (uint64_t)std::max(prop.persistingL2CacheMaxSize, 0);
    CU(cudaStreamCreateWithFlags(&d->stream, cudaStreamNonBlocking));
    sshash_gpu_info_t& in = d->info;
    in.num_kmers = f.num_kmers; in.num_strings = f.num_strings; in.k = f.k; in.m = f.m;
    in.canonical = f.canonical; in.weighted = f.weighted; in.max_k = (uint64_t)max_k;
    in.version = (uint64_t)f.version[0] << 16 | (uint64_t)f.version[1] << 8 | f.version[2];
    in.num_minimizers = f.minimizers_mphf.num_keys;
    in.mphf_partitions = f.minimizers_mphf.parts.size();
    in.skew_partitions = f.skew_mphfs.size();
    in.index_file_bytes = f.file_bytes;
    in.device = device;
    st = upload_index(d.get(), f, max_k);
    if (st != SSHASH_GPU_OK) return st;
    configure_l2(d.get());
    CU(cudaDeviceSynchronize());
    *out = d.release();
    return SSHASH_GPU_OK;
}

SSHASH_ENTRY(sshash_gpu_close, (sshash_gpu_dict* dict), (dict)) {
    delete dict;
    return SSHASH_GPU_OK;
}

SSHASH_ENTRY(sshash_gpu_set_peer_inplace, (sshash_gpu_dict* dict, int inplace), (dict, inplace)) {
    if (!dict) return fail(SSHASH_GPU_EINVAL, "null dictionary handle");
    dict->peer_inplace.store(inplace ? 1 : 0);
    return SSHASH_GPU_OK;
}

SSHASH_ENTRY(sshash_gpu_info, (const sshash_gpu_dict* dict, sshash_gpu_info_t* out), (dict, out)) {
    if (!dict || !out) return fail(SSHASH_GPU_EINVAL, "null argument");
    *out = dict->info;
    return SSHASH_GPU_OK;
}

// Device-resident batch through the partition-major path (binned.cu), in launches of at most
// binned_max_batch() queries.  The scratch belongs to a pooled workspace, so the call waits for the
// stream before handing it back (the direct kernel stays fully asynchronous).
static bool use_binned(const sshash_gpu_dict* dict, const void* in, const void* out, uint64_t n) {
    const BinPlan& bp = dict->ctx.bins;
    return bp.enabled && bp.n_bins && n >= bp.min_queries && is_local_device_pointer(in, dict->device) &&
           is_local_device_pointer(out, dict->device);
}

static int lookup_binned_device(const sshash_gpu_dict* dict, const uint64_t* kmers, uint64_t n, bool check_rc, uint64_t* ids,
                                uint32_t* ids32, uint8_t* member, void* stream) {
    const DeviceIndex& ix = dict->ix;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    WorkspaceLease ws(dict);
    const uint64_t step = std::min<uint64_t>(n, binned_max_batch());
    CU(ensure(ws->d_bin, ws->bin_cap, binned_scratch_bytes(ix, dict->ctx, step)));
    for (uint64_t off = 0; off < n; off += step) {
        const uint64_t cn = std::min(step, n - off);
        CU(launch_lookup_binned(ix, dict->ctx, kmers + off * ix.kmer_words, cn, check_rc, ids ? ids + off : nullptr,
                                ids32 ? ids32 + off : nullptr, member ? member + off : nullptr, ws->d_bin, s));
    }
    CU(cudaStreamSynchronize(s));
    return SSHASH_GPU_OK;
}

static int lookup_common(const sshash_gpu_dict* dict, const void* queries, bool ascii, uint64_t n, int check_rc,
                         uint64_t* kmer_ids, sshash_lookup_result* full, void* stream) {
    int st = check_dict(dict);
    if (st) return st;
    if (n == 0) return SSHASH_GPU_OK;
    if (!queries || (!kmer_ids && !full)) return fail(SSHASH_GPU_EINVAL, "null argument");
    const uint64_t in_elem = ascii ? dict->ix.k : 8ull * dict->ix.kmer_words;
    const DeviceIndex& ix = dict->ix;
    const LaunchCtx& sms = dict->ctx;
    const bool rc = check_rc != 0;
    if (kmer_ids && full) {
        // both outputs: ids are a column of the full records; produce the records, then the ids
        if (is_device_pointer(queries) != is_device_pointer(kmer_ids) || is_device_pointer(queries) != is_device_pointer(full))
            return fail(SSHASH_GPU_EINVAL, "kmers, kmer_ids and full must all be host or all be device pointers");
        if (is_device_pointer(queries)) {
            cudaStream_t s = static_cast<cudaStream_t>(stream);
            CU(launch_lookup(ix, sms, queries, ascii, n, rc, kmer_ids, full, nullptr, s));
            return SSHASH_GPU_OK;
        }
        st = run_batched(dict, queries, in_elem, full, sizeof(sshash_lookup_result), n, stream,
                         [&](const void* in, void* out, uint64_t cn, cudaStream_t s) {
                             return launch_lookup(ix, sms, in, ascii, cn, rc, nullptr, static_cast<sshash_lookup_result*>(out), nullptr, s);
                         });
        if (st) return st;
        for (uint64_t i = 0; i != n; ++i) kmer_ids[i] = full[i].kmer_id;
        return SSHASH_GPU_OK;
    }
    if (full)
        return run_batched(dict, queries, in_elem, full, sizeof(sshash_lookup_result), n, stream,
                           [&](const void* in, void* out, uint64_t cn, cudaStream_t s) {
                               return launch_lookup(ix, sms, in, ascii, cn, rc, nullptr, static_cast<sshash_lookup_result*>(out), nullptr, s);
                           });
    if (!ascii && use_binned(dict, queries, kmer_ids, n))
        return lookup_binned_device(dict, static_cast<const uint64_t*>(queries), n, rc, kmer_ids, nullptr, nullptr, stream);
    return run_batched(dict, queries, in_elem, kmer_ids, 8, n, stream,
                       [&](const void* in, void* out, uint64_t cn, cudaStream_t s) {
                           return launch_lookup(ix, sms, in, ascii, cn, rc, static_cast<uint64_t*>(out), nullptr, nullptr, s);
                       });
}

SSHASH_ENTRY(sshash_gpu_lookup_batch, (const sshash_gpu_dict* dict, const uint64_t* kmers, uint64_t n, int check_reverse_complement,
                            uint64_t* kmer_ids, sshash_lookup_result* full, void* stream), (dict, kmers, n, check_reverse_complement, kmer_ids, full, stream)) {
    return lookup_common(dict, kmers, false, n, check_reverse_complement, kmer_ids, full, stream);
}

SSHASH_ENTRY(sshash_gpu_lookup_batch_u32, (const sshash_gpu_dict* dict, const uint64_t* kmers, uint64_t n, int check_reverse_complement,
                                        uint32_t* kmer_ids32, void* stream), (dict, kmers, n, check_reverse_complement, kmer_ids32, stream)) {
    int st = check_dict(dict);
    if (st) return st;
    if (dict->ix.num_kmers >= 0xffffffffull) return fail(SSHASH_GPU_EINVAL, "32-bit ids need a dictionary with fewer than 2^32 - 1 k-mers");
    if (n == 0) return SSHASH_GPU_OK;
    if (!kmers || !kmer_ids32) return fail(SSHASH_GPU_EINVAL, "null argument");
    const DeviceIndex& ix = dict->ix;
    const LaunchCtx& sms = dict->ctx;
    const bool rc = check_reverse_complement != 0;
    if (use_binned(dict, kmers, kmer_ids32, n)) return lookup_binned_device(dict, kmers, n, rc, nullptr, kmer_ids32, nullptr, stream);
    return run_batched(dict, kmers, 8ull * ix.kmer_words, kmer_ids32, 4, n, stream,
                       [&](const void* in, void* out, uint64_t cn, cudaStream_t s) {
                           return launch_lookup(ix, sms, in, false, cn, rc, nullptr, nullptr, nullptr, s, static_cast<uint32_t*>(out));
                       });
}

SSHASH_ENTRY(sshash_gpu_lookup_batch_ascii, (const sshash_gpu_dict* dict, const char* kmers, uint64_t n, int check_reverse_complement,
                                  uint64_t* kmer_ids, sshash_lookup_result* full, void* stream), (dict, kmers, n, check_reverse_complement, kmer_ids, full, stream)) {
    return lookup_common(dict, kmers, true, n, check_reverse_complement, kmer_ids, full, stream);
}

SSHASH_ENTRY(sshash_gpu_is_member_batch, (const sshash_gpu_dict* dict, const uint64_t* kmers, uint64_t n, int check_reverse_complement,
                               uint8_t* member, void* stream), (dict, kmers, n, check_reverse_complement, member, stream)) {
    int st = check_dict(dict);
    if (st) return st;
    if (n == 0) return SSHASH_GPU_OK;
    if (!kmers || !member) return fail(SSHASH_GPU_EINVAL, "null argument");
    const DeviceIndex& ix = dict->ix;
    const LaunchCtx& sms = dict->ctx;
    const bool rc = check_reverse_complement != 0;
    if (use_binned(dict, kmers, member, n)) return lookup_binned_device(dict, kmers, n, rc, nullptr, nullptr, member, stream);
    return run_batched(dict, kmers, 8ull * ix.kmer_words, member, 1, n, stream,
                       [&](const void* in, void* out, uint64_t cn, cudaStream_t s) {
                           return launch_lookup(ix, sms, in, false, cn, rc, nullptr, nullptr, static_cast<uint8_t*>(out), s);
                       });
}

SSHASH_ENTRY(sshash_gpu_minimizer_partition_batch, (const sshash_gpu_dict* dict, const uint64_t* kmers, uint64_t n, uint32_t* partitions,
                                         void* stream), (dict, kmers, n, partitions, stream)) {
    int st = check_dict(dict);
    if (st) return st;
    if (n == 0) return SSHASH_GPU_OK;
    if (!kmers || !partitions) return fail(SSHASH_GPU_EINVAL, "null argument");
    const DeviceIndex& ix = dict->ix;
    const LaunchCtx& sms = dict->ctx;
    return run_batched(dict, kmers, 8ull * ix.kmer_words, partitions, 4, n, stream,
                       [&](const void* in, void* out, uint64_t cn, cudaStream_t s) {
                           return launch_minimizer_partition(ix, sms, static_cast<const uint64_t*>(in), cn, static_cast<uint32_t*>(out), s);
                       });
}

SSHASH_ENTRY(sshash_gpu_access_batch, (const sshash_gpu_dict* dict, const uint64_t* kmer_ids, uint64_t n, uint64_t* kmers_out,
                            void* stream), (dict, kmer_ids, n, kmers_out, stream)) {
    int st = check_dict(dict);
    if (st) return st;
    if (n == 0) return SSHASH_GPU_OK;
    if (!kmer_ids || !kmers_out) return fail(SSHASH_GPU_EINVAL, "null argument");
    const DeviceIndex& ix = dict->ix;
    const LaunchCtx& sms = dict->ctx;
    return run_batched(dict, kmer_ids, 8, kmers_out, 8ull * ix.kmer_words, n, stream,
                       [&](const void* in, void* out, uint64_t cn, cudaStream_t s) {
                           return launch_access(ix, sms, static_cast<const uint64_t*>(in), cn, static_cast<uint64_t*>(out), s);
                       });
}

SSHASH_ENTRY(sshash_gpu_weight_batch, (const sshash_gpu_dict* dict, const uint64_t* kmer_ids, uint64_t n, uint64_t* weights_out,
                            void* stream), (dict, kmer_ids, n, weights_out, stream)) {
    int st = check_dict(dict);
    if (st) return st;
    if (!dict->ix.n_weight_intervals) return fail(SSHASH_GPU_EINVAL, "the dictionary is not weighted");
    if (n == 0) return SSHASH_GPU_OK;
    if (!kmer_ids || !weights_out) return fail(SSHASH_GPU_EINVAL, "null argument");
    const DeviceIndex& ix = dict->ix;
    const LaunchCtx& sms = dict->ctx;
    return run_batched(dict, kmer_ids, 8, weights_out, 8, n, stream,
                       [&](const void* in, void* out, uint64_t cn, cudaStream_t s) {
                           return launch_weight(ix, sms, static_cast<const uint64_t*>(in), cn, static_cast<uint64_t*>(out), s);
                       });
}

// kmer_neighbours / string_neighbours: n inputs -> 8n results (forward A,C,T,G then backward A,C,T,G)
static int neighbours_common(const sshash_gpu_dict* dict, const uint64_t* in, bool strings, uint64_t n, int check_rc, int which,
                             uint64_t* kmer_ids, sshash_lookup_result* full, void* stream) {
    int st = check_dict(dict);
    if (st) return st;
    if (n == 0) return SSHASH_GPU_OK;
    if (!in || (!kmer_ids && !full) || which < 1 || which > 3) return fail(SSHASH_GPU_EINVAL, "bad argument");
    const DeviceIndex& ix = dict->ix;
    const LaunchCtx& ctx = dict->ctx;
    const uint64_t W = ix.kmer_words, in_elem = strings ? 8 : 8 * W;
    const bool rc = check_rc != 0;
    const bool dev = is_device_pointer(in);
    if ((kmer_ids && dev != is_device_pointer(kmer_ids)) || (full && dev != is_device_pointer(full)))
        return fail(SSHASH_GPU_EINVAL, "inputs and outputs must all be host or all be device pointers");
    WorkspaceLease ws(dict);
    if (dev) {
        Slot& s = ws->slots[0];
        CU(ensure_slot(s, 0, 0));
        CU(ensure(s.d_tmp, s.tmp_cap, 8 * n * W * 8));
        cudaStream_t cs = static_cast<cudaStream_t>(stream);
        CU(launch_neighbours(ix, ctx, in, strings, n, rc, which, static_cast<uint64_t*>(s.d_tmp), kmer_ids, full, cs));
        CU(cudaStreamSynchronize(cs));      // the expansion scratch belongs to the workspace
        return SSHASH_GPU_OK;
    }
    const uint64_t out_elem = full ? 8 * sizeof(sshash_lookup_result) : 8 * 8;
    const uint64_t chunk = std::max<uint64_t>(1, (32ull << 20) / out_elem);
    uint8_t* out = full ? reinterpret_cast<uint8_t*>(full) : reinterpret_cast<uint8_t*>(kmer_ids);
    int c = 0;
    for (uint64_t off = 0; off < n; off += chunk, ++c) {
        const uint64_t cn = std::min(chunk, n - off);
        Slot& s = ws->slots[c % Workspace::kSlots];
        CU(ensure_slot(s, chunk * in_elem, chunk * out_elem));
        CU(ensure(s.d_tmp, s.tmp_cap, 8 * chunk * W * 8));
        CU(cudaMemcpyAsync(s.d_in, reinterpret_cast<const uint8_t*>(in) + off * in_elem, cn * in_elem, cudaMemcpyHostToDevice, s.stream));
        CU(launch_neighbours(ix, ctx, static_cast<const uint64_t*>(s.d_in), strings, cn, rc, which, static_cast<uint64_t*>(s.d_tmp),
                             full ? nullptr : static_cast<uint64_t*>(s.d_out),
                             full ? static_cast<sshash_lookup_result*>(s.d_out) : nullptr, s.stream));
        CU(cudaMemcpyAsync(out + off * out_elem, s.d_out, cn * out_elem, cudaMemcpyDeviceToHost, s.stream));
    }
    for (auto& s : ws->slots) if (s.stream) CU(cudaStreamSynchronize(s.stream));
    if (full && kmer_ids) for (uint64_t i = 0; i != 8 * n; ++i) kmer_ids[i] = full[i].kmer_id;
    return SSHASH_GPU_OK;
}

SSHASH_ENTRY(sshash_gpu_kmer_neighbours_batch, (const sshash_gpu_dict* dict, const uint64_t* kmers, uint64_t n, int check_reverse_complement,
                                     int which, uint64_t* kmer_ids, sshash_lookup_result* full, void* stream), (dict, kmers, n, check_reverse_complement, which, kmer_ids, full, stream)) {
    return neighbours_common(dict, kmers, false, n, check_reverse_complement, which, kmer_ids, full, stream);
}

SSHASH_ENTRY(sshash_gpu_string_neighbours_batch, (const sshash_gpu_dict* dict, const uint64_t* string_ids, uint64_t n,
                                       int check_reverse_complement, uint64_t* kmer_ids, sshash_lookup_result* full, void* stream), (dict, string_ids, n, check_reverse_complement, kmer_ids, full, stream)) {
    return neighbours_common(dict, string_ids, true, n, check_reverse_complement, 3, kmer_ids, full, stream);
}

// Lazily (first streaming call) checks SSHash's input contract on the device (kernels.cu
// distinct_check_kernel): 1 = the index holds duplicated k-mers or reverse-complement twins, and
// streaming must replay the reference's state machine literally.  SSHASH_GPU_ASSUME_DISTINCT=1 skips
// the check, =0 forces the replay path (tests).
static int needs_replay(const sshash_gpu_dict* dict, bool* replay) {
    sshash_gpu_dict* d = const_cast<sshash_gpu_dict*>(dict);
    std::lock_guard<std::mutex> g(d->contract_mu);
    if (d->breaks_contract < 0) {
        const char* e = std::getenv("SSHASH_GPU_ASSUME_DISTINCT");
        if (e && (e[0] == '0' || e[0] == '1')) d->breaks_contract = e[0] == '0';
        else {
            uint32_t* d_flag = nullptr;
            uint32_t flag = 0;
            CU(cudaMalloc(reinterpret_cast<void**>(&d_flag), 4));
            cudaError_t err = cudaMemsetAsync(d_flag, 0, 4, d->stream);
            if (err == cudaSuccess) err = launch_distinct_check(d->ix, d->ctx, d_flag, d->stream);
            if (err == cudaSuccess) err = cudaMemcpyAsync(&flag, d_flag, 4, cudaMemcpyDeviceToHost, d->stream);
            if (err == cudaSuccess) err = cudaStreamSynchronize(d->stream);
            cudaFree(d_flag);
            if (err != cudaSuccess) return cuda_fail(err, "input-contract check");
            d->breaks_contract = flag ? 1 : 0;
        }
    }
    *replay = d->breaks_contract == 1;
    return SSHASH_GPU_OK;
}

extern "C" {
SSHASH_ENTRY(sshash_gpu_check_input_contract, (const sshash_gpu_dict* dict, int* breaks_contract), (dict, breaks_contract)) {
    int st = check_dict(dict);
    if (st) return st;
    if (!breaks_contract) return fail(SSHASH_GPU_EINVAL, "null argument");
    bool replay = false;
    st = needs_replay(dict, &replay);
    *breaks_contract = replay ? 1 : 0;
    return st;
}
}

// One device-resident batch of reads: offsets scan, window lookups, state-machine replay.
static int streaming_device(const sshash_gpu_dict* dict, Workspace& w, const char* d_bases, const uint64_t* d_read_begins,
                            const uint64_t* d_read_ends, uint64_t num_reads, uint64_t max_windows, uint64_t* d_ids_out, cudaStream_t s) {
    const DeviceIndex& ix = dict->ix;
    CU(ensure(w.d_win_offsets, w.wo_cap, (num_reads + 2) * 8));
    CU(ensure(w.d_block_sums, w.bs_cap, window_offsets_scratch_words(num_reads) * 8));
    if (max_windows > w.win_cap / 8 || !w.d_win_id || !w.d_win_aux) {
        cudaFree(w.d_win_id); cudaFree(w.d_win_aux);
        w.d_win_id = w.d_win_aux = nullptr; w.win_cap = 0;
        uint64_t bytes = (max_windows + max_windows / 8 + 32) * 8;
        CU(cudaMalloc(reinterpret_cast<void**>(&w.d_win_id), bytes));
        CU(cudaMalloc(reinterpret_cast<void**>(&w.d_win_aux), bytes));
        w.win_cap = bytes;
    }
    // SSHASH_GPU_STREAM_ALIGN=0 disables the anchor/alignment shortcut (every window is looked up)
    bool replay = false;
    {
        const int rs = needs_replay(dict, &replay);
        if (rs) return rs;
    }
    static const bool anchors_on = !(std::getenv("SSHASH_GPU_STREAM_ALIGN") && std::getenv("SSHASH_GPU_STREAM_ALIGN")[0] == '0');
    const bool use_anchors = anchors_on && !replay;
    if (use_anchors) CU(ensure(w.d_anchors, w.anchors_cap, streaming_anchor_bytes(num_reads)));
    CU(launch_window_offsets(ix.k, d_read_begins, d_read_ends, num_reads, w.d_win_offsets, w.d_block_sums, s));
    CU(launch_streaming(ix, dict->ctx, d_bases, d_read_begins, d_read_ends, w.d_win_offsets, num_reads, use_anchors ? w.d_anchors : nullptr,
                        w.d_win_id, w.d_win_aux, d_ids_out, max_windows, w.d_counters, s, replay));
    return SSHASH_GPU_OK;
}

SSHASH_ENTRY(sshash_gpu_streaming_batch, (const sshash_gpu_dict* dict, const char* bases, const uint64_t* read_offsets, uint64_t num_reads,
                               uint64_t* kmer_ids, sshash_streaming_report* report, void* stream), (dict, bases, read_offsets, num_reads, kmer_ids, report, stream)) {
    int st = check_dict(dict);
    if (st) return st;
    if (!report) return fail(SSHASH_GPU_EINVAL, "null report");
    std::memset(report, 0, sizeof(*report));
    if (num_reads == 0) return SSHASH_GPU_OK;
    if (!bases || !read_offsets) return fail(SSHASH_GPU_EINVAL, "null argument");
    const bool dev = is_device_pointer(bases);
    if (dev != is_device_pointer(read_offsets) || (kmer_ids && dev != is_device_pointer(kmer_ids)))
        return fail(SSHASH_GPU_EINVAL, "bases, read_offsets and kmer_ids must all be host or all be device pointers");
    WorkspaceLease ws(dict);
    Workspace& w = *ws.w;
    const uint32_t k = dict->ix.k;
    if (!w.d_counters) {
        CU(cudaMalloc(reinterpret_cast<void**>(&w.d_counters), 8 * sizeof(unsigned long long)));
        CU(cudaMallocHost(reinterpret_cast<void**>(&w.h_counters), 8 * sizeof(unsigned long long)));
    }
    // device buffers: the caller's stream (NULL = legacy default stream); host buffers: our own
    cudaStream_t s = dev ? static_cast<cudaStream_t>(stream) : dict->stream;
    CU(cudaMemsetAsync(w.d_counters, 0, 8 * sizeof(unsigned long long), s));
    if (dev) {
        // total bases bound the number of windows; two 8-byte reads tell us how many
        uint64_t first = 0, last = 0;
        CU(cudaMemcpyAsync(&first, read_offsets, 8, cudaMemcpyDeviceToHost, s));
        CU(cudaMemcpyAsync(&last, read_offsets + num_reads, 8, cudaMemcpyDeviceToHost, s));
        CU(cudaStreamSynchronize(s));
        st = streaming_device(dict, w, bases, read_offsets, read_offsets + 1, num_reads, last - first + 1, kmer_ids, s);
        if (st) return st;
    } else {
        // Chunks of reads (<= 16 MB of bases, <= 1 Mi reads each) cycle through two staging sets: the
        // H2D copy of chunk c+1 runs on a copy stream while the kernels of chunk c run on the compute
        // stream.  Offsets are copied as they are (absolute); the kernels get the staging pointer
        // rebased by the chunk's first offset instead of a rewritten offsets array.
        // SSHASH_GPU_STREAM_CHUNK: bases per chunk (tests use tiny chunks to exercise the pipeline)
        const uint64_t max_bases = env_bytes("SSHASH_GPU_STREAM_CHUNK", 16ull << 20), max_reads = 1ull << 20;
        if (!w.copy_stream) CU(cudaStreamCreateWithFlags(&w.copy_stream, cudaStreamNonBlocking));
        for (auto& ss : w.sslots) {
            if (!ss.ready) CU(cudaEventCreateWithFlags(&ss.ready, cudaEventDisableTiming));
            if (!ss.free) CU(cudaEventCreateWithFlags(&ss.free, cudaEventDisableTiming));
        }
        uint64_t r0 = 0, win_done = 0;
        for (int c = 0; r0 < num_reads; ++c) {
            uint64_t r1 = r0 + 1;
            while (r1 < num_reads && r1 - r0 < max_reads && read_offsets[r1 + 1] - read_offsets[r0] <= max_bases) ++r1;
            const uint64_t first = read_offsets[r0], nb = read_offsets[r1] - first, nr = r1 - r0;
            uint64_t nwin = 0;
            for (uint64_t i = r0; i < r1; ++i) { const uint64_t len = read_offsets[i + 1] - read_offsets[i]; if (len >= k) nwin += len - k + 1; }
            Workspace::StreamSlot& ss = w.sslots[c & 1];
            if (c >= 2) CU(cudaEventSynchronize(ss.free));   // chunk c-2 is done with this staging set (ensure() may reallocate it)
            CU(ensure(ss.d_bases, ss.bases_cap, nb + kPadBytes));
            CU(ensure(ss.d_ro, ss.ro_cap, (nr + 1) * 8));
            if (kmer_ids) CU(ensure(ss.d_ids, ss.ids_cap, (nwin + 1) * 8));
            CU(cudaMemcpyAsync(ss.d_bases, bases + first, nb, cudaMemcpyHostToDevice, w.copy_stream));
            CU(cudaMemcpyAsync(ss.d_ro, read_offsets + r0, (nr + 1) * 8, cudaMemcpyHostToDevice, w.copy_stream));
            CU(cudaEventRecord(ss.ready, w.copy_stream));
            CU(cudaStreamWaitEvent(s, ss.ready, 0));
            const char* rebased = static_cast<const char*>(ss.d_bases) - first;
            st = streaming_device(dict, w, rebased, ss.d_ro, ss.d_ro + 1, nr, nwin + 1, kmer_ids ? ss.d_ids : nullptr, s);
            if (st) return st;
            if (kmer_ids && nwin) CU(cudaMemcpyAsync(kmer_ids + win_done, ss.d_ids, nwin * 8, cudaMemcpyDeviceToHost, s));
            CU(cudaEventRecord(ss.free, s));
            win_done += nwin;
            r0 = r1;
        }
    }
    CU(cudaMemcpyAsync(w.h_counters, w.d_counters, 5 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, s));
    CU(cudaStreamSynchronize(s));
    report->num_kmers = w.h_counters[0];
    report->num_searches = w.h_counters[1];
    report->num_extensions = w.h_counters[2];
    report->num_negative_kmers = w.h_counters[3];
    report->num_invalid_kmers = w.h_counters[4];
    report->num_positive_kmers = report->num_searches + report->num_extensions;   // streaming_query.hpp:113
    return SSHASH_GPU_OK;
}

}  // extern "C"

namespace {

// ------------------------------------------------------------------------------------------------
// File driver with DEVICE-side record parsing (FASTQ, single-line FASTA; plain or gzip).
// The host only moves bytes: a reader thread fills two pinned buffers (parallel pread for plain
// files, zlib inflate for .gz), each chunk is copied to HBM as is, the GPU finds the lines and
// the sequence spans (kernels.cu, "record parsing") and the streaming kernels run on the raw bytes.
// Per chunk the host learns one number (how many lines the chunk holds) and carries the bytes of
// the unfinished last record over to the next chunk.
// ------------------------------------------------------------------------------------------------
class ChunkReader {
public:
    ~ChunkReader() { if (gz_) gzclose(gz_); else if (fd_ >= 0) ::close(fd_); }
    bool open(const char* filename) {
        fd_ = ::open(filename, O_RDONLY);
        if (fd_ < 0) return false;
        unsigned char magic[2] = {0, 0};
        const ssize_t got = ::pread(fd_, magic, 2, 0);
        if (got == 2 && magic[0] == 0x1f && magic[1] == 0x8b) {
            gz_ = gzdopen(fd_, "rb");
            if (!gz_) return false;
            gzbuffer(gz_, 1 << 20);
        }
        return true;
    }
    // up to n bytes into dst; returns the number read (< n only at end of file), -1 on error
    int64_t read(uint8_t* dst, uint64_t n) {
        if (gz_) {
            uint64_t done = 0;
            while (done < n) {
                const int r = gzread(gz_, dst + done, (unsigned)std::min<uint64_t>(n - done, 1u << 30));
                if (r < 0) return -1;
                if (r == 0) break;
                done += (uint64_t)r;
            }
            return (int64_t)done;
        }
        // plain file: the page-cache copy is the bottleneck of one thread, so slices are read in parallel
        const unsigned hw = std::thread::hardware_concurrency();
        const uint64_t nt = n >= (4u << 20) ? std::max(1u, std::min(16u, hw ? hw : 1u)) : 1;
        const uint64_t slice = (n + nt - 1) / nt;
        std::vector<int64_t> got(nt, 0);
        auto work = [&](uint64_t t) {
            const uint64_t b = t * slice, e = std::min(n, b + slice);
            uint64_t done = b;
            while (done < e) {
                const ssize_t r = ::pread(fd_, dst + done, e - done, (off_t)(pos_ + done));
                if (r < 0) { got[t] = -1; return; }
                if (r == 0) break;
                done += (uint64_t)r;
            }
            got[t] = (int64_t)(done - b);
        };
        std::vector<std::thread> th;
        for (uint64_t t = 1; t < nt; ++t) th.emplace_back(work, t);
        work(0);
        for (auto& x : th) x.join();
        uint64_t total = 0;
        for (uint64_t t = 0; t < nt; ++t) {
            if (got[t] < 0) return -1;
            total += (uint64_t)got[t];
            if ((uint64_t)got[t] < std::min(n, (t + 1) * slice) - std::min(n, t * slice)) break;   // end of file inside this slice
        }
        pos_ += total;
        return (int64_t)total;
    }
private:
    int fd_ = -1;
    gzFile gz_ = nullptr;
    uint64_t pos_ = 0;
};

// *need_host_parser = true (and OK returned) when a single record does not fit a chunk: the caller
// then runs the host line parser over the whole file instead.
int stream_file_device_parse(const sshash_gpu_dict* dict, const char* filename, bool fastq, sshash_streaming_report* report,
                             bool* need_host_parser) {
    *need_host_parser = false;
    // SSHASH_GPU_FILE_CHUNK: bytes per chunk (tests use tiny chunks to exercise the record carry-over)
    const uint64_t chunk = std::max<uint64_t>(64, env_bytes("SSHASH_GPU_FILE_CHUNK", 32ull << 20));
    const uint64_t carry_max = chunk;
    const uint32_t stride = fastq ? 4 : 2;
    ChunkReader reader;
    if (!reader.open(filename)) return fail(SSHASH_GPU_EIO, std::string("error in opening the file '") + filename + "'");

    WorkspaceLease ws(dict);
    Workspace& w = *ws.w;
    const uint64_t host_cap = carry_max + chunk + 64;
    if (w.h_file_cap < host_cap) {
        for (auto*& h : w.h_file) { if (h) cudaFreeHost(h); h = nullptr; }
        w.h_file_cap = 0;
        for (auto*& h : w.h_file) CU(cudaMallocHost(reinterpret_cast<void**>(&h), host_cap));
        w.h_file_cap = host_cap;
    }
    if (!w.d_counters) {
        CU(cudaMalloc(reinterpret_cast<void**>(&w.d_counters), 8 * sizeof(unsigned long long)));
        CU(cudaMallocHost(reinterpret_cast<void**>(&w.h_counters), 8 * sizeof(unsigned long long)));
    }
    // Two streams: the copy stream moves chunk c+1 to HBM and counts its lines (the one number the host needs)
    // while the compute stream runs the span + streaming kernels of chunk c.
    cudaStream_t s = dict->stream;
    if (!w.copy_stream) CU(cudaStreamCreateWithFlags(&w.copy_stream, cudaStreamNonBlocking));
    cudaStream_t cs = w.copy_stream;
    for (int i = 0; i < 2; ++i) {
        if (!w.file_counted[i]) CU(cudaEventCreateWithFlags(&w.file_counted[i], cudaEventDisableTiming));
        if (!w.file_done[i]) CU(cudaEventCreateWithFlags(&w.file_done[i], cudaEventDisableTiming));
    }
    CU(cudaMemsetAsync(w.d_counters, 0, 8 * sizeof(unsigned long long), s));
    bool slot_used[2] = {false, false};

    // reader thread: fills slot 0, 1, 0, ... at offset carry_max, hands each over with its byte count
    struct Handoff {
        std::mutex mu; std::condition_variable cv;
        int64_t filled[2] = {-2, -2};   // -2 = free, -1 = read error, >= 0 = bytes read
        bool stop = false;
    } ho;
    std::thread rd([&] {
        for (int slot = 0;; slot ^= 1) {
            {
                std::unique_lock<std::mutex> lk(ho.mu);
                ho.cv.wait(lk, [&] { return ho.filled[slot] == -2 || ho.stop; });
                if (ho.stop) return;
            }
            const int64_t n = reader.read(w.h_file[slot] + carry_max, chunk);
            {
                std::lock_guard<std::mutex> lk(ho.mu);
                ho.filled[slot] = n;
            }
            ho.cv.notify_all();
            if (n < (int64_t)chunk) return;   // error or end of file
        }
    });
    struct Joiner {
        std::thread& t; Handoff& h;
        ~Joiner() { { std::lock_guard<std::mutex> lk(h.mu); h.stop = true; } h.cv.notify_all(); if (t.joinable()) t.join(); }
    } joiner{rd, ho};

    std::vector<uint8_t> carry;
    int st = SSHASH_GPU_OK;
    for (int slot = 0;; slot ^= 1) {
        int64_t got;
        {
            std::unique_lock<std::mutex> lk(ho.mu);
            ho.cv.wait(lk, [&] { return ho.filled[slot] != -2; });
            got = ho.filled[slot];
        }
        if (got < 0) return fail(SSHASH_GPU_EIO, std::string("error in reading the file '") + filename + "'");
        const bool eof = (uint64_t)got < chunk;
        uint8_t* begin = w.h_file[slot] + carry_max - carry.size();
        if (!carry.empty()) std::memcpy(begin, carry.data(), carry.size());
        uint64_t n = carry.size() + (uint64_t)got;
        if (eof) {   // terminate the last line and complete the last record (missing lines read as empty, query.cpp:86-107)
            std::memset(begin + n, '\n', stride + 1);
            n += stride + 1;
        }
        const uint64_t tiles = parse_tiles(n);
        if (slot_used[slot]) CU(cudaEventSynchronize(w.file_done[slot]));   // chunk c-2 has been streamed: its buffers are free (ensure may reallocate)
        CU(ensure(w.d_raw[slot], w.raw_cap[slot], n + 64));
        CU(ensure(w.d_tiles[slot], w.tiles_cap[slot], (tiles + 2) * 8));
        uint8_t* d_raw = w.d_raw[slot];
        uint64_t* d_tiles = w.d_tiles[slot];
        CU(cudaMemcpyAsync(d_raw, begin, n, cudaMemcpyHostToDevice, cs));
        CU(cudaMemsetAsync(d_tiles + tiles, 0, 8, cs));
        CU(launch_count_lines(d_raw, n, d_tiles, cs));
        CU(cudaMemcpyAsync(w.h_counters + 6, d_tiles + tiles, 8, cudaMemcpyDeviceToHost, cs));
        CU(cudaEventRecord(w.file_counted[slot], cs));
        CU(cudaStreamSynchronize(cs));                   // waits for this chunk's copy + count only; chunk c-1 keeps streaming
        const uint64_t lines = w.h_counters[6], records = lines / stride;
        if (records == 0 && !eof) { *need_host_parser = true; return SSHASH_GPU_OK; }
        // the unfinished record: everything after the newline that ends line stride * records - 1
        carry.clear();
        if (!eof) {
            uint64_t cut = n;
            for (uint64_t q = lines - stride * records + 1; q; --q) {
                const void* nl = memrchr(begin, '\n', cut);
                cut = (uint64_t)(static_cast<const uint8_t*>(nl) - begin);
            }
            cut += 1;
            carry.assign(begin + cut, begin + n);
            if (carry.size() > carry_max) { *need_host_parser = true; return SSHASH_GPU_OK; }
        }
        {   // the pinned slot has been copied and the carry saved: hand it back to the reader
            std::lock_guard<std::mutex> lk(ho.mu);
            ho.filled[slot] = -2;
        }
        ho.cv.notify_all();
        if (records) {
            CU(cudaStreamWaitEvent(s, w.file_counted[slot], 0));
            CU(ensure(w.d_line_start, w.ls_cap, (lines + 2) * 8));   // growing = cudaFree + cudaMalloc: cudaFree waits for the device, so chunk c-1 cannot be using the old buffer
            CU(ensure(w.d_spans, w.spans_cap, 2 * records * 8));
            CU(launch_read_spans(d_raw, n, d_tiles, w.d_line_start, records, stride, w.d_spans, w.d_spans + records,
                                 dict->ctx.sm_count, s));
            st = streaming_device(dict, w, reinterpret_cast<const char*>(d_raw), w.d_spans, w.d_spans + records, records, n + 1,
                                  nullptr, s);
            if (st) return st;
        }
        CU(cudaEventRecord(w.file_done[slot], s));
        slot_used[slot] = true;
        if (eof) break;
    }
    CU(cudaMemcpyAsync(w.h_counters, w.d_counters, 5 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, s));
    CU(cudaStreamSynchronize(s));
    report->num_kmers = w.h_counters[0];
    report->num_searches = w.h_counters[1];
    report->num_extensions = w.h_counters[2];
    report->num_negative_kmers = w.h_counters[3];
    report->num_invalid_kmers = w.h_counters[4];
    report->num_positive_kmers = report->num_searches + report->num_extensions;
    return SSHASH_GPU_OK;
}

}  // namespace

extern "C" {

// Host-side file driver: same record structure as the reference's drivers (src/query.cpp:53-108):
// FASTA = header line + one sequence line per record, FASTQ = 4 lines per record; gz through zlib.
SSHASH_ENTRY(sshash_gpu_streaming_query_from_file, (const sshash_gpu_dict* dict, const char* filename, int multiline,
                                         sshash_streaming_report* report), (dict, filename, multiline, report)) {
    int st = check_dict(dict);
    if (st) return st;
    if (!filename || !report) return fail(SSHASH_GPU_EINVAL, "null argument");
    std::memset(report, 0, sizeof(*report));
    std::string fn(filename);
    auto ends_with = [&](const char* suf) {
        size_t n = std::strlen(suf);
        return fn.size() >= n && fn.compare(fn.size() - n, n, suf) == 0;
    };
    bool fastq;
    if (ends_with(".fa.gz") || ends_with(".fasta.gz") || ends_with(".fa") || ends_with(".fasta")) fastq = false;
    else if (ends_with(".fq.gz") || ends_with(".fastq.gz") || ends_with(".fq") || ends_with(".fastq")) fastq = true;
    else {   // query.cpp:169-171: only a message on stderr, empty report
        std::fprintf(stderr, "unsupported query file format\n");
        return SSHASH_GPU_OK;
    }
    if (multiline && fastq) multiline = 0;   // query.cpp:154-163: the flag only affects FASTA input
    // default: records are parsed on the GPU.  SSHASH_GPU_HOST_PARSER=1 (or multiline FASTA, or a
    // record that does not fit one chunk) takes the host line parser below.
    if (!multiline && !(std::getenv("SSHASH_GPU_HOST_PARSER") && std::getenv("SSHASH_GPU_HOST_PARSER")[0] == '1')) {
        bool need_host_parser = false;
        st = stream_file_device_parse(dict, filename, fastq, report, &need_host_parser);
        if (st || !need_host_parser) return st;
        std::memset(report, 0, sizeof(*report));
    }
    gzFile gz = gzopen(filename, "rb");   // transparently reads uncompressed files too
    if (!gz) return fail(SSHASH_GPU_EIO, "error in opening the file '" + fn + "'");
    gzbuffer(gz, 1 << 20);
    std::string bases;
    std::vector<uint64_t> offsets{0};
    std::vector<char> line(1 << 16);
    auto getline = [&](std::string* dst) -> bool {   // false at EOF with nothing read
        bool any = false;
        for (;;) {
            if (!gzgets(gz, line.data(), (int)line.size())) return any;
            any = true;
            size_t len = std::strlen(line.data());
            bool eol = len && line[len - 1] == '\n';
            if (eol) --len;
            if (dst) dst->append(line.data(), len);
            if (eol) return true;
        }
    };
    sshash_streaming_report total{};
    auto flush = [&]() -> int {
        if (offsets.size() <= 1) return SSHASH_GPU_OK;
        sshash_streaming_report r{};
        int rc = sshash_gpu_streaming_batch_impl(dict, bases.data(), offsets.data(), offsets.size() - 1, nullptr, &r, nullptr);
        if (rc) return rc;
        total.num_kmers += r.num_kmers; total.num_positive_kmers += r.num_positive_kmers;
        total.num_negative_kmers += r.num_negative_kmers; total.num_invalid_kmers += r.num_invalid_kmers;
        total.num_searches += r.num_searches; total.num_extensions += r.num_extensions;
        bases.clear(); offsets.assign(1, 0);
        return SSHASH_GPU_OK;
    };
    if (multiline) {
        // streaming_query_from_fasta_file_multiline (query.cpp:9-51) + buffered_lines_iterator
        // (util.hpp:287-340): the reference concatenates ALL lines -- header lines included, their
        // characters simply make windows invalid -- and slides over the concatenation without
        // resetting the query; only an EMPTY line ends the run (buffer cleared, query.reset()).  The
        // 1024-character refills keep the last k-1 characters, so every window of a run is visited
        // exactly once: a run is one read.  (A run shorter than k-1 characters underflows num_kmers
        // in the reference, :22; here it contributes nothing.)
        for (;;) {
            const size_t before = bases.size();
            if (!getline(&bases)) break;
            if (bases.size() == before) {            // empty line: end of the run
                if (bases.size() > offsets.back()) offsets.push_back(bases.size());
                if (bases.size() >= (256u << 20)) { st = flush(); if (st) { gzclose(gz); return st; } }
            }
        }
        if (bases.size() > offsets.back()) offsets.push_back(bases.size());
    } else
    for (;;) {
        if (!getline(nullptr)) break;                 // header
        if (!getline(&bases)) { /* header without sequence: empty read */ }
        offsets.push_back(bases.size());
        if (fastq) { getline(nullptr); getline(nullptr); }   // '+' and quality
        if (bases.size() >= (256u << 20)) { st = flush(); if (st) { gzclose(gz); return st; } }
    }
    gzclose(gz);
    st = flush();
    if (st) return st;
    *report = total;
    return SSHASH_GPU_OK;
}

}  // extern "C"
