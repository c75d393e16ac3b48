// Host-side half of the end-to-end path, exercised without a GPU: the tapered upload chunks and the block-scheduled pack
// pool of tarok_rollout_host_packed (tarok_b200/csrc/tarok_host.cpp) against the single-threaded serialiser.
//   g++ -O2 -std=c++17 -pthread host_pool_harness.cpp ../tarok_b200/csrc/tarok_host.cpp -o harness && ./harness <rows> <threads>
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <vector>

#include "../include/tarok_b200.h"
#include "../tarok_b200/csrc/tarok_host.h"

int main(int argc, char** argv) {
    const uint64_t n = argc > 1 ? strtoull(argv[1], nullptr, 10) : 300007;
    const int threads = argc > 2 ? atoi(argv[2]) : 4;
    std::vector<uint8_t> perm(n * 54), c(n, 3), d(n), k(n);
    std::mt19937 rng(1);
    for (uint64_t g = 0; g < n; g++) {
        uint8_t* r = &perm[g * 54];
        for (int i = 0; i < 54; i++) r[i] = (uint8_t)i;
        std::shuffle(r, r + 54, rng);
        d[g] = g & 3; k[g] = g % 5;
    }
    if (n > 8) perm[54 * 7 + 3] = perm[54 * 7 + 4];                           // one invalid row
    const size_t bytes = n * TAROK_RECORD_BYTES;
    uint8_t* ref = (uint8_t*)aligned_alloc(64, (bytes + 127) / 64 * 64);
    uint8_t* out = (uint8_t*)aligned_alloc(64, (bytes + 127) / 64 * 64);
    const int64_t bad = tarok_pack_records_mt(perm.data(), c.data(), d.data(), k.data(), n, ref, 1);
    tarok_pack_pool* p = tarok_pack_pool_create(threads);
    for (int rep = 0; rep < 5; rep++) {
        memset(out, 0xEE, bytes);
        uint64_t bounds[40];
        const int want = 1 + 7 * rep;                                         // 1, 8, 15, 22, 29 chunks asked for
        const int nch = tarok_chunk_bounds(n, want, tarok_pack_block_rows(), bounds);
        if (bounds[0] != 0 || bounds[nch] != n || nch < 1 || nch > want) { printf("BAD BOUNDS\n"); return 1; }
        for (int i = 0; i < nch; i++) {
            if (bounds[i + 1] <= bounds[i]) { printf("EMPTY CHUNK\n"); return 1; }
            if (i + 1 < nch && bounds[i + 1] % tarok_pack_block_rows()) { printf("UNALIGNED CHUNK\n"); return 1; }
        }
        if (nch >= 4 && !(bounds[1] - bounds[0] <= bounds[nch / 2 + 1] - bounds[nch / 2])) { printf("NOT TAPERED\n"); return 1; }
        tarok_pack_pool_begin(p, perm.data(), c.data(), d.data(), k.data(), n, bounds, nch, out);
        for (int ch = 0; ch < nch; ch++) {
            tarok_pack_pool_wait_chunk(p, ch);                                // chunk ch must be complete NOW, later ones need not be
            if (memcmp(out + bounds[ch] * TAROK_RECORD_BYTES, ref + bounds[ch] * TAROK_RECORD_BYTES,
                       (bounds[ch + 1] - bounds[ch]) * TAROK_RECORD_BYTES)) { printf("MISMATCH chunk %d of %d\n", ch, nch); return 1; }
        }
        if (tarok_pack_pool_bad(p) != bad) { printf("BAD COUNT %lld != %lld\n", (long long)tarok_pack_pool_bad(p), (long long)bad); return 1; }
    }
    tarok_pack_pool_destroy(p);
    printf("ok rows %llu threads %d bad %lld\n", (unsigned long long)n, threads, (long long)bad);
    return 0;
}
