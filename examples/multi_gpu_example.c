/* Plain C above the C ABI: one index replicated on every visible GPU behind ONE handle, batches sharded
 * by query inside the library (include/sshash_gpu.h, sshash_gpu_multi_*).  Self-checking like the
 * reference's own check loop (test/check.hpp:29-49): lookup(access(id)) == id, the 32-bit ids and
 * the membership bytes agree, streamed ids equal looked-up ids.
 *   gcc -std=c11 -O2 examples/multi_gpu_example.c -o multi_gpu_example sshash_b200/libsshash_gpu.so -Wl,-rpath,$PWD/sshash_b200
 *   ./multi_gpu_example tests/golden/se_k31_m13.sshash [num_queries] [num_gpus]
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../include/sshash_gpu.h"

#define CHECK(call)                                                                         \
    do {                                                                                    \
        int st__ = (call);                                                                  \
        if (st__ != SSHASH_GPU_OK) {                                                        \
            fprintf(stderr, "%s failed (%d): %s\n", #call, st__, sshash_gpu_last_error()); \
            return 1;                                                                       \
        }                                                                                   \
    } while (0)

static uint64_t rng_state = 42;
static uint64_t next_u64(void) {   /* splitmix64 */
    uint64_t z = (rng_state += 0x9e3779b97f4a7c15ull);
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
    return z ^ (z >> 31);
}

int main(int argc, char** argv) {
    if (argc < 2) { fprintf(stderr, "usage: %s <index.sshash> [num_queries] [num_gpus]\n", argv[0]); return 2; }
    const uint64_t n = argc > 2 ? strtoull(argv[2], NULL, 10) : 1000000;
    const int want_gpus = argc > 3 ? atoi(argv[3]) : 0;            /* 0 = every visible GPU */
    sshash_gpu_multi* m = NULL;
    CHECK(sshash_gpu_multi_open(argv[1], NULL, want_gpus, 0, &m));
    const int gpus = sshash_gpu_multi_num_devices(m);
    sshash_gpu_info_t info;
    CHECK(sshash_gpu_info(sshash_gpu_multi_dict(m, 0), &info));
    const uint64_t w = info.max_k == 31 ? 1 : 2;
    printf("%d GPU(s), k=%lu m=%lu num_kmers=%lu\n", gpus, (unsigned long)info.k, (unsigned long)info.m, (unsigned long)info.num_kmers);

    uint64_t* ids = malloc(n * 8);
    uint64_t* kmers = malloc(n * w * 8);
    uint64_t* got = malloc(n * 8);
    uint32_t* got32 = malloc(n * 4);
    uint8_t* member = malloc(n);
    for (uint64_t i = 0; i != n; ++i) ids[i] = next_u64() % info.num_kmers;
    /* positives from replica 0, every 5th query replaced by a random (absent) k-mer */
    CHECK(sshash_gpu_access_batch(sshash_gpu_multi_dict(m, 0), ids, n, kmers, NULL));
    for (uint64_t i = 4; i < n; i += 5) {
        kmers[i * w] = next_u64() >> 2;
        if (w == 2) kmers[i * w + 1] = next_u64() & ((1ull << (2 * info.k - 64)) - 1);
        else if (info.k < 32) kmers[i * w] &= (1ull << (2 * info.k)) - 1;
    }
    CHECK(sshash_gpu_multi_lookup_batch(m, kmers, n, 1, got));
    CHECK(sshash_gpu_multi_lookup_batch_u32(m, kmers, n, 1, got32));
    CHECK(sshash_gpu_multi_is_member_batch(m, kmers, n, 1, member));
    uint64_t bad = 0, negatives = 0;
    for (uint64_t i = 0; i != n; ++i) {
        if (i % 5 != 4) bad += got[i] != ids[i];
        else negatives += got[i] == SSHASH_GPU_INVALID;
        bad += got32[i] != (uint32_t)got[i];
        bad += member[i] != (got[i] != SSHASH_GPU_INVALID);
    }
    /* the single-GPU path gives the same answers on the same buffer */
    uint64_t* ref = malloc(n * 8);
    CHECK(sshash_gpu_lookup_batch(sshash_gpu_multi_dict(m, gpus - 1), kmers, n, 1, ref, NULL, NULL));
    for (uint64_t i = 0; i != n; ++i) bad += ref[i] != got[i];
    printf("%lu lookups sharded over %d GPU(s): %lu mismatches, %lu of %lu random k-mers absent\n", (unsigned long)n, gpus,
           (unsigned long)bad, (unsigned long)negatives, (unsigned long)(n / 5));

    /* streaming over reads cut from the positives' strings is covered by the Python tests; here: reads made of k-mers */
    if (w == 1) {
        const uint64_t reads = n < 20000 ? n : 20000, k = info.k;
        char* bases = malloc(reads * k);
        uint64_t* offs = malloc((reads + 1) * 8);
        for (uint64_t r = 0; r != reads; ++r) {
            offs[r] = r * k;
            for (uint64_t j = 0; j != k; ++j) bases[r * k + j] = "ACTG"[(kmers[r] >> (2 * j)) & 3];
        }
        offs[reads] = reads * k;
        uint64_t* sids = malloc(reads * 8);
        sshash_streaming_report rep;
        CHECK(sshash_gpu_multi_streaming_batch(m, bases, offs, reads, sids, &rep));
        uint64_t sbad = 0;
        for (uint64_t r = 0; r != reads; ++r) sbad += sids[r] != got[r];
        printf("%lu one-window reads streamed over %d GPU(s): %lu mismatches, %lu positive + %lu negative\n", (unsigned long)reads,
               gpus, (unsigned long)sbad, (unsigned long)rep.num_positive_kmers, (unsigned long)rep.num_negative_kmers);
        bad += sbad + (rep.num_kmers != reads);
        free(bases); free(offs); free(sids);
    }
    CHECK(sshash_gpu_multi_close(m));
    free(ids); free(kmers); free(got); free(got32); free(member); free(ref);
    printf("%s\n", bad == 0 ? "0 mismatches" : "MISMATCHES");
    return bad == 0 ? 0 : 1;
}
