// launch.cuh -- internal helpers shared by the kernel translation units (kernels.cu, binned.cu).
#pragma once

#include <cuda_runtime.h>

#include <atomic>
#include <cstdint>

#include "kernels.cuh"

namespace sshash_b200 {

constexpr int kBlock = 256;

extern std::atomic<uint64_t> g_launches;   // kernels.cu

template <int W>
__device__ __forceinline__ Kmer<W> load_kmer(const uint64_t* __restrict__ kmers, uint64_t i);
template <>
__device__ __forceinline__ Kmer<1> load_kmer<1>(const uint64_t* __restrict__ kmers, uint64_t i) {
    return {__ldcs(kmers + i)};
}
template <>
__device__ __forceinline__ Kmer<2> load_kmer<2>(const uint64_t* __restrict__ kmers, uint64_t i) {
    ulonglong2 v = __ldcs(reinterpret_cast<const ulonglong2*>(kmers) + i);
    return {v.x, v.y};
}
__device__ __forceinline__ void store_kmer(uint64_t* out, uint64_t i, Kmer<1> x) { __stcs(out + i, x.lo); }
__device__ __forceinline__ void store_kmer(uint64_t* out, uint64_t i, Kmer<2> x) {
    __stcs(reinterpret_cast<ulonglong2*>(out) + i, make_ulonglong2(x.lo, x.hi));
}


// every kernel of the lookup path is launched through here: <<<grid, kBlock>>> plus the L2
// access-policy window that keeps the hot slab (locate tables, pilots, ...) persistent in L2
template <typename... KArgs, typename... Args>
inline cudaError_t launch(void (*kernel)(KArgs...), int grid, cudaStream_t stream, const LaunchCtx& ctx, Args... args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3(kBlock);
    cfg.dynamicSmemBytes = 0;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    cfg.attrs = attr;
    cfg.numAttrs = 0;
    if (ctx.window_bytes) {
        attr[0].id = cudaLaunchAttributeAccessPolicyWindow;
        attr[0].val.accessPolicyWindow.base_ptr = const_cast<void*>(ctx.hot_base);
        attr[0].val.accessPolicyWindow.num_bytes = ctx.window_bytes;
        attr[0].val.accessPolicyWindow.hitRatio = ctx.hit_ratio;
        attr[0].val.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
        attr[0].val.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
        cfg.numAttrs = 1;
    }
    g_launches.fetch_add(1);
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

inline int grid_for(uint64_t n, int sm_count, int blocks_per_sm) {
    uint64_t need = (n + kBlock - 1) / kBlock;
    uint64_t cap = (uint64_t)sm_count * blocks_per_sm;
    if (need < 1) need = 1;
    return (int)(need < cap ? need : cap);
}


}  // namespace sshash_b200
