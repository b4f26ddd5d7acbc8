// multi.cu -- the multi-GPU handle of the C ABI (include/sshash_gpu.h, "several GPUs of one box").
//
// lookups are independent against a static index that fits one GPU many times over, so
// the index is REPLICATED on every GPU and a batch is SHARDED by query: device j gets the contiguous
// slice [lo_j, hi_j) of the batch and runs the single-GPU path on it.  There is no exchange step in
// the algorithm.  One host thread per device drives its slice through the single-device entry points:
//   * HOST buffers: every GPU copies its own slice in and its own results out (its own PCIe link,
//     its own copy engines); the "gather" is the fact that all slices land in one host array;
//   * DEVICE buffers (on any one GPU of the box): the owning GPU works in place; every other GPU stages
//     its slice chunk by chunk with peer-to-peer copy-engine transfers over NVLink (api.cu run_batched:
//     chunk c+1 in, kernel of chunk c, ids of chunk c-1 out overlap), i.e. the ids arrive in the
//     destination GPU's vector without any SM of any GPU spent on communication.
#include <cuda_runtime.h>

#include <algorithm>
#include <string>
#include <thread>
#include <vector>

#include "../../include/sshash_gpu.h"
#include "api_internal.hpp"

using sshash_b200::set_last_error;

struct sshash_gpu_multi {
    std::vector<int> devices;
    std::vector<sshash_gpu_dict*> dicts;
    ~sshash_gpu_multi() {
        for (auto* d : dicts) if (d) sshash_gpu_close(d);
    }
};

namespace {

// contiguous slice of n items owned by shard r of w: sizes differ by at most one, rank order = item order
void shard_range(uint64_t n, uint64_t r, uint64_t w, uint64_t* lo, uint64_t* hi) {
    const uint64_t base = n / w, rem = n % w;
    *lo = r * base + std::min(r, rem);
    *hi = *lo + base + (r < rem ? 1 : 0);
}

// run fn(j) for every device on its own host thread; first failure wins
template <typename F>
int for_each_device(const sshash_gpu_multi* m, F&& fn) {
    const size_t w = m->dicts.size();
    std::vector<int> status(w, SSHASH_GPU_OK);
    std::vector<std::string> message(w);
    auto run = [&](size_t j) {
        try {
            status[j] = fn(j);
            if (status[j] != SSHASH_GPU_OK) message[j] = sshash_gpu_last_error();   // this thread's message
        } catch (const std::exception& e) {
            status[j] = SSHASH_GPU_EINVAL;
            message[j] = std::string("internal error: ") + e.what();
        }
    };
    std::vector<std::thread> th;
    th.reserve(w);
    for (size_t j = 1; j < w; ++j) th.emplace_back(run, j);
    run(0);
    for (auto& t : th) t.join();
    for (size_t j = 0; j < w; ++j)
        if (status[j] != SSHASH_GPU_OK)
            return set_last_error(status[j], "device " + std::to_string(m->devices[j]) + ": " + message[j]);
    return SSHASH_GPU_OK;
}

template <typename F>
int guarded(F&& body) noexcept {
    try { return body(); }
    catch (const std::bad_alloc&) { return set_last_error(SSHASH_GPU_ENOMEM, "out of host memory"); }
    catch (const std::exception& e) { return set_last_error(SSHASH_GPU_EINVAL, std::string("internal error: ") + e.what()); }
    catch (...) { return set_last_error(SSHASH_GPU_EINVAL, "internal error: unknown exception"); }
}

// a device-buffer call on the owning GPU is asynchronous on the legacy default stream: wait for it
int finish(const sshash_gpu_multi* m, size_t j, int st) {
    if (st != SSHASH_GPU_OK) return st;
    cudaError_t e = cudaSetDevice(m->devices[j]);
    if (e == cudaSuccess) e = cudaStreamSynchronize(nullptr);
    return e == cudaSuccess ? SSHASH_GPU_OK : set_last_error(SSHASH_GPU_ECUDA, cudaGetErrorString(e));
}

uint64_t kmer_words(const sshash_gpu_multi* m) {
    sshash_gpu_info_t info;
    sshash_gpu_info(m->dicts[0], &info);
    return info.max_k == 31 ? 1 : 2;
}

}  // namespace

extern "C" {

int sshash_gpu_multi_open(const char* index_path, const int* devices, int n_devices, int max_k, sshash_gpu_multi** out) {
    return guarded([&]() -> int {
        if (!index_path || !out) return set_last_error(SSHASH_GPU_EINVAL, "null argument");
        *out = nullptr;
        int ndev = 0;
        if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
            cudaGetLastError();
            return set_last_error(SSHASH_GPU_ECUDA, "no CUDA device available (this library has no CPU fallback)");
        }
        auto m = std::make_unique<sshash_gpu_multi>();
        if (n_devices <= 0) n_devices = ndev;                     // all visible devices
        for (int i = 0; i < n_devices; ++i) {
            const int dev = devices ? devices[i] : i;
            if (dev < 0 || dev >= ndev) return set_last_error(SSHASH_GPU_EINVAL, "invalid device ordinal");
            if (std::find(m->devices.begin(), m->devices.end(), dev) != m->devices.end())
                return set_last_error(SSHASH_GPU_EINVAL, "duplicate device ordinal");
            m->devices.push_back(dev);
        }
        m->dicts.assign(m->devices.size(), nullptr);
        // replicas are uploaded concurrently, one host thread per GPU
        int st = for_each_device(m.get(), [&](size_t j) { return sshash_gpu_open(index_path, m->devices[j], max_k, &m->dicts[j]); });
        if (st != SSHASH_GPU_OK) return st;
        // peer access between every pair: slices of a device-resident batch move GPU to GPU over NVLink
        for (int a : m->devices)
            for (int b : m->devices) {
                if (a == b) continue;
                int can = 0;
                if (cudaDeviceCanAccessPeer(&can, a, b) != cudaSuccess || !can) { cudaGetLastError(); continue; }
                if (cudaSetDevice(a) != cudaSuccess) { cudaGetLastError(); continue; }
                const cudaError_t e = cudaDeviceEnablePeerAccess(b, 0);
                if (e != cudaSuccess) cudaGetLastError();             // already enabled counts as fine
            }
        *out = m.release();
        return SSHASH_GPU_OK;
    });
}

int sshash_gpu_multi_close(sshash_gpu_multi* m) {
    delete m;
    return SSHASH_GPU_OK;
}

int sshash_gpu_multi_num_devices(const sshash_gpu_multi* m) { return m ? (int)m->dicts.size() : 0; }

const sshash_gpu_dict* sshash_gpu_multi_dict(const sshash_gpu_multi* m, int i) {
    return (m && i >= 0 && (size_t)i < m->dicts.size()) ? m->dicts[(size_t)i] : nullptr;
}

int sshash_gpu_multi_lookup_batch(const sshash_gpu_multi* m, const uint64_t* kmers, uint64_t n, int check_reverse_complement,
                                  uint64_t* kmer_ids) {
    return guarded([&]() -> int {
        if (!m) return set_last_error(SSHASH_GPU_EINVAL, "null multi-GPU handle");
        if (n == 0) return SSHASH_GPU_OK;
        if (!kmers || !kmer_ids) return set_last_error(SSHASH_GPU_EINVAL, "null argument");
        const uint64_t W = kmer_words(m), w = m->dicts.size();
        return for_each_device(m, [&](size_t j) {
            uint64_t lo, hi;
            shard_range(n, j, w, &lo, &hi);
            return finish(m, j, sshash_gpu_lookup_batch(m->dicts[j], kmers + lo * W, hi - lo, check_reverse_complement, kmer_ids + lo,
                                                        nullptr, nullptr));
        });
    });
}

int sshash_gpu_multi_lookup_batch_u32(const sshash_gpu_multi* m, const uint64_t* kmers, uint64_t n, int check_reverse_complement,
                                      uint32_t* kmer_ids32) {
    return guarded([&]() -> int {
        if (!m) return set_last_error(SSHASH_GPU_EINVAL, "null multi-GPU handle");
        if (n == 0) return SSHASH_GPU_OK;
        if (!kmers || !kmer_ids32) return set_last_error(SSHASH_GPU_EINVAL, "null argument");
        const uint64_t W = kmer_words(m), w = m->dicts.size();
        return for_each_device(m, [&](size_t j) {
            uint64_t lo, hi;
            shard_range(n, j, w, &lo, &hi);
            return finish(m, j, sshash_gpu_lookup_batch_u32(m->dicts[j], kmers + lo * W, hi - lo, check_reverse_complement,
                                                            kmer_ids32 + lo, nullptr));
        });
    });
}

int sshash_gpu_multi_is_member_batch(const sshash_gpu_multi* m, const uint64_t* kmers, uint64_t n, int check_reverse_complement,
                                     uint8_t* member) {
    return guarded([&]() -> int {
        if (!m) return set_last_error(SSHASH_GPU_EINVAL, "null multi-GPU handle");
        if (n == 0) return SSHASH_GPU_OK;
        if (!kmers || !member) return set_last_error(SSHASH_GPU_EINVAL, "null argument");
        const uint64_t W = kmer_words(m), w = m->dicts.size();
        return for_each_device(m, [&](size_t j) {
            uint64_t lo, hi;
            shard_range(n, j, w, &lo, &hi);
            return finish(m, j, sshash_gpu_is_member_batch(m->dicts[j], kmers + lo * W, hi - lo, check_reverse_complement, member + lo,
                                                           nullptr));
        });
    });
}

// Streaming membership over HOST reads, sharded by read: device j takes a contiguous run of reads
// holding ~1/w of the bases; ids land at the run's window offset; the six counters add up (they are
// per-read sums: streaming_query is reset for every read, src/query.cpp:78-108).
int sshash_gpu_multi_streaming_batch(const sshash_gpu_multi* m, const char* bases, const uint64_t* read_offsets, uint64_t num_reads,
                                     uint64_t* kmer_ids, sshash_streaming_report* report) {
    return guarded([&]() -> int {
        if (!m) return set_last_error(SSHASH_GPU_EINVAL, "null multi-GPU handle");
        if (!report) return set_last_error(SSHASH_GPU_EINVAL, "null report");
        *report = sshash_streaming_report{};
        if (num_reads == 0) return SSHASH_GPU_OK;
        if (!bases || !read_offsets) return set_last_error(SSHASH_GPU_EINVAL, "null argument");
        cudaPointerAttributes a;
        if (cudaPointerGetAttributes(&a, bases) == cudaSuccess && a.type == cudaMemoryTypeDevice)
            return set_last_error(SSHASH_GPU_EINVAL, "the multi-GPU streaming call takes host buffers");
        cudaGetLastError();
        sshash_gpu_info_t info;
        sshash_gpu_info(m->dicts[0], &info);
        const uint64_t k = info.k, w = m->dicts.size();
        // read boundaries at equal shares of the bases
        std::vector<uint64_t> cut(w + 1, num_reads);
        cut[0] = 0;
        const uint64_t total = read_offsets[num_reads] - read_offsets[0];
        for (uint64_t j = 1; j < w; ++j) {
            const uint64_t target = read_offsets[0] + total / w * j;
            cut[j] = (uint64_t)(std::lower_bound(read_offsets, read_offsets + num_reads, target) - read_offsets);
            cut[j] = std::max(cut[j], cut[j - 1]);
        }
        std::vector<uint64_t> win0(w + 1, 0);
        if (kmer_ids) {
            for (uint64_t j = 0; j < w; ++j) {
                uint64_t s = 0;
                for (uint64_t r = cut[j]; r < cut[j + 1]; ++r) { const uint64_t len = read_offsets[r + 1] - read_offsets[r]; if (len >= k) s += len - k + 1; }
                win0[j + 1] = win0[j] + s;
            }
        }
        std::vector<sshash_streaming_report> rep(w);
        int st = for_each_device(m, [&](size_t j) {
            rep[j] = sshash_streaming_report{};
            if (cut[j + 1] == cut[j]) return (int)SSHASH_GPU_OK;
            return sshash_gpu_streaming_batch(m->dicts[j], bases, read_offsets + cut[j], cut[j + 1] - cut[j],
                                              kmer_ids ? kmer_ids + win0[j] : nullptr, &rep[j], nullptr);
        });
        if (st != SSHASH_GPU_OK) return st;
        for (auto const& r : rep) {
            report->num_kmers += r.num_kmers; report->num_positive_kmers += r.num_positive_kmers;
            report->num_negative_kmers += r.num_negative_kmers; report->num_invalid_kmers += r.num_invalid_kmers;
            report->num_searches += r.num_searches; report->num_extensions += r.num_extensions;
        }
        return SSHASH_GPU_OK;
    });
}

}  // extern "C"
