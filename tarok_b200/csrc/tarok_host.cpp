// Host-side half of libtarok_b200.so (plain C++, compiled by g++): the serialiser that turns the permutation rows a
// patched Igra.shuffle produces (Igra.py:65-73: uint8 [n,54]) + forced contracts into 24-byte deal records
// (layout in include/tarok_b200.h), so the host-buffer entry moves 24 instead of 57 bytes per deal over PCIe.
// No device work here.  tarok_pack_records_mt splits the rows over `threads` std::threads (the callers of the
// host-buffer entries are launched by torchrun with OMP_NUM_THREADS=1, so the thread count is an explicit argument).
#include "../../include/tarok_b200.h"

#include <immintrin.h>

#include <atomic>
#include <condition_variable>
#include <cstring>
#include <mutex>
#include <thread>
#include <vector>

#include "tarok_host.h"

namespace {

constexpr uint64_t ALL54 = (1ull << 54) - 1;

// One row -> three words.  The five owner sets are built with independent accumulators (the loop over a 12-card
// segment is a load, a shift and an OR per card); the bit planes are then unions of the sets:
// code 0-3 = seat, 4 = talon -> plane 0 = seats 1|3, plane 1 = seats 2|3, plane 2 = talon.
__attribute__((always_inline)) inline bool pack_row(const uint8_t* row, unsigned contract, unsigned declarer, unsigned king, uint64_t* w) {
    uint64_t h[4] = {0, 0, 0, 0}, talon = 0;
    unsigned over = 0;
    for (int s = 0; s < 4; s++) {
        uint64_t a = 0, b = 0;
        const uint8_t* p = row + 12 * s;
        for (int i = 0; i < 12; i += 2) {
            over |= p[i] | p[i + 1];
            a |= 1ull << (p[i] & 63);
            b |= 1ull << (p[i + 1] & 63);
        }
        h[s] = a | b;
    }
    uint64_t ranks = 0;
    for (int i = 0; i < 6; i++) { over |= row[48 + i]; talon |= 1ull << (row[48 + i] & 63); }
    for (int i = 0; i < 6; i++) {                         // position in the talon of its cards taken in ascending id
        const unsigned c = row[48 + i] & 63;
        const int below = __builtin_popcountll(talon & ((1ull << c) - 1));
        ranks |= (uint64_t)i << (3 * below);
    }
    // a permutation of 0..53 <=> the five sets are disjoint, have 12/12/12/12/6 members and cover ALL54 (ids < 64 checked
    // through `over`: an id of 54..63 breaks the cover, one >= 64 sets bit 6 or 7 of `over`)
    const bool perm_ok = over < 64 && (h[0] | h[1] | h[2] | h[3] | talon) == ALL54
        && __builtin_popcountll(h[0]) == 12 && __builtin_popcountll(h[1]) == 12 && __builtin_popcountll(h[2]) == 12
        && __builtin_popcountll(h[3]) == 12 && __builtin_popcountll(talon) == 6;
    const bool ok = perm_ok && contract <= 15u && declarer <= 3u;
    uint64_t w0 = h[1] | h[3], w1 = h[2] | h[3], w2 = talon;
    if (!ok) { w0 = w1 = w2 = ALL54; ranks = 0; }         // decodes to an error game, like the row itself would
    const unsigned k = king < 7u ? king : 7u;
    w[0] = w0 | (ranks & 0x1FF) << 54;
    w[1] = w1 | ((ranks >> 9) & 0x1FF) << 54;
    w[2] = w2 | (uint64_t)(contract & 15u) << 54 | (uint64_t)(declarer & 3u) << 58 | (uint64_t)(k & 7u) << 60;
    return ok;
}

// Built twice (function multi-versioning, resolved once at load time): with BMI2 the variable shifts are single SHLX
// micro-ops instead of the three-uop SHL-by-CL and the eleven popcounts are instructions instead of library calls.
__attribute__((target_clones("default", "arch=x86-64-v3")))
int64_t pack_range_scalar(const uint8_t* perm, const uint8_t* contract, const uint8_t* declarer, const uint8_t* king, uint64_t a,
                          uint64_t b, uint64_t* records) {
    int64_t bad = 0;
    for (uint64_t g = a; g < b; g++)
        bad += pack_row(perm + g * 54, contract[g], declarer[g], king ? king[g] : 7u, records + g * 3) ? 0 : 1;
    return bad;
}

// AVX-512 version: the row's 54 ids become one-hot 64-bit words eight at a time (VPMOVZXBQ + VPSLLVQ; a shift count >= 64
// gives 0, so an out-of-range id can only LOSE a bit and the cover / count checks below reject the row), the words are OR-ed
// per owner -- the twelve cards of a seat are one and a half vectors, hence the lane masks -- and reduced two owners at a time.
#define TK_AVX512 __attribute__((target("avx512f,avx512bw,avx512vl,avx512dq,bmi2,popcnt")))
TK_AVX512 static inline __m128i or_reduce_pair(__m512i a, __m512i b) {          // -> [OR of a's lanes, OR of b's lanes]
    const __m512i t = _mm512_or_si512(_mm512_unpacklo_epi64(a, b), _mm512_unpackhi_epi64(a, b));   // per 128-bit lane [a, b]
    const __m256i u = _mm256_or_si256(_mm512_castsi512_si256(t), _mm512_extracti64x4_epi64(t, 1));
    return _mm_or_si128(_mm256_castsi256_si128(u), _mm256_extracti128_si256(u, 1));
}
TK_AVX512 static inline __m512i one_hot8(const uint8_t* p) {                    // eight ids -> eight one-hot 64-bit words
    return _mm512_sllv_epi64(_mm512_set1_epi64(1), _mm512_cvtepu8_epi64(_mm_loadl_epi64((const __m128i*)p)));
}
TK_AVX512 static inline bool pack_row_avx512(const uint8_t* row, unsigned contract, unsigned declarer, unsigned king, uint64_t* w) {
    const __m512i one = _mm512_set1_epi64(1);
    const __m512i v0 = one_hot8(row), v1 = one_hot8(row + 8), v2 = one_hot8(row + 16), v3 = one_hot8(row + 24),
                  v4 = one_hot8(row + 32), v5 = one_hot8(row + 40);
    // the six talon ids: a masked load (never reads past the row: the last row of a buffer ends there)
    const __m128i t6 = _mm_maskz_loadu_epi8((__mmask16)0x3F, row + 48);
    const __m512i v6 = _mm512_maskz_sllv_epi64((__mmask8)0x3F, one, _mm512_cvtepu8_epi64(t6));
    const __m512i a0 = _mm512_or_si512(v0, _mm512_maskz_mov_epi64((__mmask8)0x0F, v1));            // seat 0: ids 0..11
    const __m512i a1 = _mm512_or_si512(_mm512_maskz_mov_epi64((__mmask8)0xF0, v1), v2);            // seat 1: ids 12..23
    const __m512i a2 = _mm512_or_si512(v3, _mm512_maskz_mov_epi64((__mmask8)0x0F, v4));            // seat 2
    const __m512i a3 = _mm512_or_si512(_mm512_maskz_mov_epi64((__mmask8)0xF0, v4), v5);            // seat 3
    const __m128i r01 = or_reduce_pair(a0, a1), r23 = or_reduce_pair(a2, a3), r4 = or_reduce_pair(v6, v6);
    const uint64_t h0 = (uint64_t)_mm_cvtsi128_si64(r01), h1 = (uint64_t)_mm_extract_epi64(r01, 1);
    const uint64_t h2 = (uint64_t)_mm_cvtsi128_si64(r23), h3 = (uint64_t)_mm_extract_epi64(r23, 1);
    const uint64_t talon = (uint64_t)_mm_cvtsi128_si64(r4);
    uint64_t ranks = 0;
    for (int i = 0; i < 6; i++) {                         // position in the talon of its cards taken in ascending id
        const unsigned c = row[48 + i];
        const int below = __builtin_popcountll(_bzhi_u64(talon, c < 64 ? c : 0));
        ranks |= (uint64_t)i << (3 * below);
    }
    const bool perm_ok = (h0 | h1 | h2 | h3 | talon) == ALL54 && __builtin_popcountll(h0) == 12
        && __builtin_popcountll(h1) == 12 && __builtin_popcountll(h2) == 12 && __builtin_popcountll(h3) == 12
        && __builtin_popcountll(talon) == 6;
    const bool ok = perm_ok && contract <= 15u && declarer <= 3u;
    uint64_t w0 = h1 | h3, w1 = h2 | h3, w2 = talon;
    if (!ok) { w0 = w1 = w2 = ALL54; ranks = 0; }
    const unsigned k = king < 7u ? king : 7u;
    w[0] = w0 | (ranks & 0x1FF) << 54;
    w[1] = w1 | ((ranks >> 9) & 0x1FF) << 54;
    w[2] = w2 | (uint64_t)(contract & 15u) << 54 | (uint64_t)(declarer & 3u) << 58 | (uint64_t)(k & 7u) << 60;
    return ok;
}
TK_AVX512 static int64_t pack_range_avx512(const uint8_t* perm, const uint8_t* contract, const uint8_t* declarer,
                                           const uint8_t* king, uint64_t a, uint64_t b, uint64_t* records) {
    int64_t bad = 0;
    for (uint64_t g = a; g < b; g++)
        bad += pack_row_avx512(perm + g * 54, contract[g], declarer[g], king ? king[g] : 7u, records + g * 3) ? 0 : 1;
    return bad;
}

std::atomic<int> g_force_scalar{0};

int64_t pack_range(const uint8_t* perm, const uint8_t* contract, const uint8_t* declarer, const uint8_t* king, uint64_t a,
                   uint64_t b, uint64_t* records) {
    static const bool wide = __builtin_cpu_supports("avx512f") && __builtin_cpu_supports("avx512bw")
                          && __builtin_cpu_supports("avx512vl") && __builtin_cpu_supports("avx512dq") && __builtin_cpu_supports("bmi2");
    return (wide && !g_force_scalar.load(std::memory_order_relaxed)) ? pack_range_avx512(perm, contract, declarer, king, a, b, records)
                : pack_range_scalar(perm, contract, declarer, king, a, b, records);
}

}  // namespace

// ---- persistent pack pool: the chunked host pipeline packs one chunk at a time, so the workers are kept alive between
// chunks (spawning threads per chunk would cost more than the packing) -------------------------------------------------
struct tarok_pack_pool {
    std::vector<std::thread> workers;
    std::mutex mu;
    std::condition_variable cv_job, cv_done;
    uint64_t generation = 0;
    int pending = 0;
    bool stop = false;
    // the current job
    const uint8_t *perm = nullptr, *contract = nullptr, *declarer = nullptr, *king = nullptr;
    uint64_t a = 0, b = 0;
    uint64_t* records = nullptr;
    std::atomic<int64_t> bad{0};
    int parts = 1;

    void slice(int t) {
        const uint64_t len = b - a, per = (len + (uint64_t)parts - 1) / (uint64_t)parts;
        const uint64_t lo = a + per * (uint64_t)t, hi = lo + per < b ? lo + per : b;
        if (lo < hi) bad.fetch_add(pack_range(perm, contract, declarer, king, lo, hi, records));
    }
    void loop(int t) {
        uint64_t seen = 0;
        for (;;) {
            std::unique_lock<std::mutex> lk(mu);
            cv_job.wait(lk, [&] { return stop || generation != seen; });
            if (stop) return;
            seen = generation;
            lk.unlock();
            slice(t);
            lk.lock();
            if (--pending == 0) cv_done.notify_one();
        }
    }
};

tarok_pack_pool* tarok_pack_pool_create(int threads) {
    if (threads < 1) threads = 1;
    tarok_pack_pool* p = new (std::nothrow) tarok_pack_pool();
    if (!p) return nullptr;
    p->parts = threads;
    try {
        for (int t = 1; t < threads; t++) p->workers.emplace_back([p, t] { p->loop(t); });
    } catch (...) {
        p->parts = (int)p->workers.size() + 1;          // fewer threads than asked for: still correct
    }
    return p;
}

void tarok_pack_pool_destroy(tarok_pack_pool* p) {
    if (!p) return;
    {
        std::lock_guard<std::mutex> lk(p->mu);
        p->stop = true;
    }
    p->cv_job.notify_all();
    for (auto& w : p->workers) w.join();
    delete p;
}

int tarok_pack_pool_threads(const tarok_pack_pool* p) { return p ? p->parts : 0; }

// Packs rows [a, b) with every thread of the pool (the caller's thread takes slice 0); returns the number of bad rows.
int64_t tarok_pack_pool_run(tarok_pack_pool* p, const uint8_t* perm, const uint8_t* contract, const uint8_t* declarer,
                            const uint8_t* king, uint64_t a, uint64_t b, uint64_t* records) {
    {
        std::lock_guard<std::mutex> lk(p->mu);
        p->perm = perm; p->contract = contract; p->declarer = declarer; p->king = king;
        p->a = a; p->b = b; p->records = records;
        p->bad.store(0);
        p->pending = (int)p->workers.size();
        p->generation++;
    }
    p->cv_job.notify_all();
    p->slice(0);
    std::unique_lock<std::mutex> lk(p->mu);
    p->cv_done.wait(lk, [&] { return p->pending == 0; });
    return p->bad.load();
}

extern "C" {

int64_t tarok_pack_records_mt(const uint8_t* perm, const uint8_t* contract, const uint8_t* declarer, const uint8_t* king,
                              uint64_t n, uint64_t* records, int threads) {
    if (!perm || !contract || !declarer || !records) return -1;
    if (threads < 1) threads = 1;
    if ((uint64_t)threads > n / 4096 + 1) threads = (int)(n / 4096 + 1);      // not worth a thread below 4096 rows
    if (threads == 1) return pack_range(perm, contract, declarer, king, 0, n, records);
    std::vector<int64_t> bad((size_t)threads, 0);
    std::vector<std::thread> pool;
    pool.reserve((size_t)threads - 1);
    const uint64_t per = (n + (uint64_t)threads - 1) / (uint64_t)threads;
    for (int t = 1; t < threads; t++) {
        const uint64_t a = per * (uint64_t)t, b = a + per < n ? a + per : n;
        if (a >= n) break;
        pool.emplace_back([=, &bad] { bad[(size_t)t] = pack_range(perm, contract, declarer, king, a, b, records); });
    }
    bad[0] = pack_range(perm, contract, declarer, king, 0, per < n ? per : n, records);
    int64_t total = 0;
    for (auto& th : pool) th.join();
    for (int64_t v : bad) total += v;
    return total;
}

int tarok_pack_force_scalar(int on) { return g_force_scalar.exchange(on ? 1 : 0); }

int tarok_pack_uses_avx512(void) {
    if (g_force_scalar.load()) return 0;
    return __builtin_cpu_supports("avx512f") && __builtin_cpu_supports("avx512bw") && __builtin_cpu_supports("avx512vl")
        && __builtin_cpu_supports("avx512dq") && __builtin_cpu_supports("bmi2");
}

int64_t tarok_pack_records(const uint8_t* perm, const uint8_t* contract, const uint8_t* declarer, const uint8_t* king,
                           uint64_t n, uint64_t* records) {
    return tarok_pack_records_mt(perm, contract, declarer, king, n, records, 1);
}

}  // extern "C"
