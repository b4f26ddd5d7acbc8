// Micro-benchmark: DRAM bytes moved per random 8-byte gather on B200, by load flavour.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gather gather.cu ; ncu --metrics dram__bytes_read.sum,gpu__time_duration.sum ./gather
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ uint64_t mix(uint64_t x) { x ^= x >> 33; x *= 0xff51afd7ed558ccdull; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ull; x ^= x >> 33; return x; }

template <int MODE>
__global__ void gather(const uint64_t* __restrict__ a, uint64_t nwords, uint64_t n, uint64_t* out) {
    uint64_t acc = 0;
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        uint64_t j = mix(i * 0x9e3779b97f4a7c15ull + 12345) % (nwords - 8);
        const uint64_t* p = a + j;
        uint64_t v;
        if (MODE == 0) v = *p;
        else if (MODE == 1) v = __ldg(p);
        else if (MODE == 2) v = __ldcs(p);
        else if (MODE == 3) { uint64_t pol; asm("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol)); asm volatile("ld.global.nc.L2::cache_hint.b64 %0, [%1], %2;" : "=l"(v) : "l"(p), "l"(pol)); }
        else if (MODE == 4) { asm volatile("ld.global.nc.L1::no_allocate.b64 %0, [%1];" : "=l"(v) : "l"(p)); }
        else if (MODE == 5) { asm volatile("ld.global.nc.L2::64B.b64 %0, [%1];" : "=l"(v) : "l"(p)); }
        else if (MODE == 6) { v = __ldg(p) + __ldg(p + 1); }   // two adjacent words like the funnel reads
        else if (MODE == 7) { asm volatile("ld.global.cv.b64 %0, [%1];" : "=l"(v) : "l"(p)); }
        else if (MODE == 8) { asm volatile("ld.global.L1::no_allocate.L2::64B.b64 %0, [%1];" : "=l"(v) : "l"(p)); }
        acc += v;
    }
    if (acc == 0x123456789) out[0] = acc;
}

int main() {
    const uint64_t nwords = 1ull << 27;  // 1 GiB
    const uint64_t n = 1ull << 26;       // 64M gathers
    uint64_t *a, *out;
    cudaMalloc(&a, nwords * 8); cudaMalloc(&out, 8);
    cudaMemset(a, 1, nwords * 8);
    size_t g = 0; cudaDeviceGetLimit(&g, cudaLimitMaxL2FetchGranularity); printf("L2 fetch granularity limit: %zu\n", g);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
#define RUN(M) { gather<M><<<148 * 8, 256>>>(a, nwords, n, out); cudaEventRecord(e0); gather<M><<<148 * 8, 256>>>(a, nwords, n, out); cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); printf("mode %d: %.3f ms  %.2f Ggather/s\n", M, ms, n / ms / 1e6); }
    RUN(0) RUN(1) RUN(2) RUN(3) RUN(4) RUN(5) RUN(6) RUN(7) RUN(8)
    cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, 32);
    cudaDeviceGetLimit(&g, cudaLimitMaxL2FetchGranularity); printf("L2 fetch granularity limit now: %zu\n", g);
    RUN(0) RUN(1)
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
