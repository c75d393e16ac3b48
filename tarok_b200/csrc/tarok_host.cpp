// Host-side half of libtarok_b200.so (plain C++, compiled by g++): the serialiser that turns the permutation rows a
// patched Igra.shuffle produces (Igra.py:65-73: uint8 [n,54]) + forced contracts into 24-byte deal records
// (layout in include/tarok_b200.h), so the host-buffer entry moves 24 instead of 57 bytes per deal over PCIe.
// No device work here.  tarok_pack_records_mt splits the rows over `threads` std::threads (the callers of the
// host-buffer entries are launched by torchrun with OMP_NUM_THREADS=1, so the thread count is an explicit argument).
#include "../../include/tarok_b200.h"

#include <immintrin.h>

#include <atomic>
#include <condition_variable>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <thread>
#include <vector>

#include "tarok_host.h"

namespace {

constexpr uint64_t ALL54 = (1ull << 54) - 1;

// One row -> three words (layout "deal records" in include/tarok_b200.h): the owner sets are built with independent
// accumulators (a load, a shift and an OR per card); plane 0 = seats 1|3, plane 1 = seats 2|3; the six talon ids go into
// w2 as they stand (6 bits each, talon order), followed by the forced contract.
constexpr uint64_t ERR_PLANE = ALL54;                      // both planes full: every card "belongs to seat 3" -> decodes to an error game
__attribute__((always_inline)) inline bool pack_row(const uint8_t* row, unsigned contract, unsigned declarer, unsigned king, uint64_t* w) {
    uint64_t h[4] = {0, 0, 0, 0}, talon = 0, order = 0;
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
    for (int i = 0; i < 6; i++) {
        const unsigned c = row[48 + i];
        over |= c;
        talon |= 1ull << (c & 63);
        order |= (uint64_t)(c & 63) << (6 * i);
    }
    // 54 positions covering all 54 ids <=> a permutation of 0..53 (an id of 54..63 breaks the cover, one >= 64 sets bit 6
    // or 7 of `over`)
    const bool ok = over < 64 && (h[0] | h[1] | h[2] | h[3] | talon) == ALL54 && contract <= 15u && declarer <= 3u;
    const unsigned k = king < 7u ? king : 7u;
    w[0] = ok ? (h[1] | h[3]) : ERR_PLANE;
    w[1] = ok ? (h[2] | h[3]) : ERR_PLANE;
    w[2] = ok ? (order | (uint64_t)contract << 36 | (uint64_t)declarer << 40 | (uint64_t)k << 42) : 0ull;
    return ok;
}

// Built twice (function multi-versioning, resolved once at load time): with BMI2 the variable shifts are single SHLX
// micro-ops instead of the three-uop SHL-by-CL.
__attribute__((target_clones("default", "arch=x86-64-v3")))
int64_t pack_range_scalar(const uint8_t* perm, const uint8_t* contract, const uint8_t* declarer, const uint8_t* king, uint64_t a,
                          uint64_t b, uint64_t* records) {
    int64_t bad = 0;
    for (uint64_t g = a; g < b; g++)
        bad += pack_row(perm + g * 54, contract[g], declarer[g], king ? king[g] : 7u, records + g * 3) ? 0 : 1;
    return bad;
}

// AVX-512 version, EIGHT ROWS AT A TIME, one row per 64-bit lane, so nothing is ever reduced across lanes: seven gathers
// fetch bytes 8j..8j+7 of each of the eight rows; position 8j+k of every row is then one shift (byte k to the bottom), one
// variable rotate of the constant 1 (VPROLVQ takes its count modulo 64 from the low six bits, so the other bytes of the
// lane need no masking) and one OR into the accumulator of the position's owner.  ids >= 64 are caught by OR-ing the raw
// bytes, ids 54..63 and duplicates by the cover test.  The last gather of a row reads two bytes of the next row, so the
// caller keeps at least one row behind every group (the tail goes through the scalar code).
#define TK_AVX512 __attribute__((target("avx512f,avx512bw,avx512vl,avx512dq,bmi2,popcnt")))
template <int K> TK_AVX512 static inline __m512i one_hot_at(__m512i v) {        // 1 << (byte K of each lane & 63)
    // the rotate runs on port 0 only (512-bit), so the byte is brought down by a byte shuffle (port 5) rather than a shift
    // (port 0) for all but two of the eight positions: the two ports end up evenly loaded
    if (K == 0) return _mm512_rolv_epi64(_mm512_set1_epi64(1), v);
    if (K == 7 || K == 4) return _mm512_rolv_epi64(_mm512_set1_epi64(1), _mm512_srli_epi64(v, 8 * K));
    return _mm512_rolv_epi64(_mm512_set1_epi64(1), _mm512_shuffle_epi8(v, _mm512_broadcast_i32x4(_mm_set_epi64x(0x8080808080808008ll | K, 0x8080808080808000ll | K))));
}
#define TK_OR3(a, b, c) _mm512_ternarylogic_epi64((a), (b), (c), 0xFE)
TK_AVX512 static inline __m512i or8(__m512i v) {                                // the eight positions of a gathered word, one owner
    return TK_OR3(TK_OR3(one_hot_at<0>(v), one_hot_at<1>(v), one_hot_at<2>(v)), TK_OR3(one_hot_at<3>(v), one_hot_at<4>(v), one_hot_at<5>(v)),
                  _mm512_or_si512(one_hot_at<6>(v), one_hot_at<7>(v)));
}
TK_AVX512 static inline __m512i or4lo_into(__m512i acc, __m512i v) {
    return TK_OR3(TK_OR3(acc, one_hot_at<0>(v), one_hot_at<1>(v)), one_hot_at<2>(v), one_hot_at<3>(v));
}
TK_AVX512 static inline __m512i or4hi_into(__m512i acc, __m512i v) {
    return TK_OR3(TK_OR3(acc, one_hot_at<4>(v), one_hot_at<5>(v)), one_hot_at<6>(v), one_hot_at<7>(v));
}
TK_AVX512 static int64_t pack_range_avx512(const uint8_t* perm, const uint8_t* contract, const uint8_t* declarer,
                                           const uint8_t* king, uint64_t a, uint64_t b, uint64_t* records) {
    int64_t bad = 0;
    uint64_t g = a;
    // a group's three stores are 64-byte aligned when the buffer is and the group starts at a multiple of eight rows: the
    // records then leave with non-temporal stores (no read-for-ownership of lines the CPU never reads back; the DMA engine does)
    const bool aligned = ((uintptr_t)records & 63) == 0;
    if (aligned && (g & 7)) {
        const uint64_t peel = (g + 7) & ~7ull;
        const uint64_t upto = peel < b ? peel : b;
        bad += pack_range_scalar(perm, contract, declarer, king, g, upto, records);
        g = upto;
    }
    const __m512i row_off = _mm512_setr_epi64(0, 54, 108, 162, 216, 270, 324, 378);
    const __m512i all54 = _mm512_set1_epi64((long long)ALL54);
    // records of eight rows interleaved: [w0 w1 w2] x 8 = three stores
    const __m512i ia0 = _mm512_setr_epi64(0, 8, 0, 1, 9, 0, 2, 10), ib0 = _mm512_setr_epi64(0, 0, 0, 0, 0, 1, 0, 0);
    const __m512i ia1 = _mm512_setr_epi64(0, 3, 11, 0, 4, 12, 0, 5), ib1 = _mm512_setr_epi64(2, 0, 0, 3, 0, 0, 4, 0);
    const __m512i ia2 = _mm512_setr_epi64(13, 0, 6, 14, 0, 7, 15, 0), ib2 = _mm512_setr_epi64(0, 5, 0, 0, 6, 0, 0, 7);
    for (; g + 8 < b; g += 8) {
        const uint8_t* base = perm + g * 54;
        for (int i = 0; i < 7; i++) _mm_prefetch((const char*)(base + 432 * 12 + 64 * i), _MM_HINT_T0);   // the gathers' lines, 12 groups ahead
        const __m512i v0 = _mm512_i64gather_epi64(row_off, base, 1), v1 = _mm512_i64gather_epi64(row_off, base + 8, 1),
                      v2 = _mm512_i64gather_epi64(row_off, base + 16, 1), v3 = _mm512_i64gather_epi64(row_off, base + 24, 1),
                      v4 = _mm512_i64gather_epi64(row_off, base + 32, 1), v5 = _mm512_i64gather_epi64(row_off, base + 40, 1);
        const __m512i v6 = _mm512_and_si512(_mm512_i64gather_epi64(row_off, base + 48, 1), _mm512_set1_epi64(0x0000FFFFFFFFFFFFll));
        const __m512i s0 = or4lo_into(or8(v0), v1);                             // positions 0..11
        const __m512i s1 = or4hi_into(or8(v2), v1);                             // 12..23
        const __m512i s2 = or4lo_into(or8(v3), v4);                             // 24..35
        const __m512i s3 = or4hi_into(or8(v5), v4);                             // 36..47
        const __m512i tal = _mm512_or_si512(TK_OR3(one_hot_at<0>(v6), one_hot_at<1>(v6), one_hot_at<2>(v6)),
                                            TK_OR3(one_hot_at<3>(v6), one_hot_at<4>(v6), one_hot_at<5>(v6)));
        const __m512i cover = _mm512_or_si512(TK_OR3(s0, s1, s2), _mm512_or_si512(s3, tal));
        const __m512i raw = _mm512_or_si512(TK_OR3(v0, v1, v2), _mm512_or_si512(TK_OR3(v3, v4, v5), v6));
        const __m512i c8 = _mm512_cvtepu8_epi64(_mm_loadl_epi64((const __m128i*)(contract + g)));
        const __m512i d8 = _mm512_cvtepu8_epi64(_mm_loadl_epi64((const __m128i*)(declarer + g)));
        const __m512i k8 = king ? _mm512_min_epu64(_mm512_cvtepu8_epi64(_mm_loadl_epi64((const __m128i*)(king + g))), _mm512_set1_epi64(7))
                                : _mm512_set1_epi64(7);
        const __mmask8 ok = _mm512_cmpeq_epi64_mask(cover, all54)
                          & _mm512_testn_epi64_mask(raw, _mm512_set1_epi64((long long)0xC0C0C0C0C0C0C0C0ull))
                          & _mm512_cmple_epu64_mask(c8, _mm512_set1_epi64(15)) & _mm512_cmple_epu64_mask(d8, _mm512_set1_epi64(3));
        // the six talon ids -> 36 contiguous bits: byte pairs (x1, x64) -> 12-bit fields, field pairs (x1, x4096) -> 24 bits
        const __m512i t12 = _mm512_maddubs_epi16(v6, _mm512_set1_epi16(0x4001));
        const __m512i t24 = _mm512_madd_epi16(t12, _mm512_set1_epi32(0x10000001));
        const __m512i order = _mm512_or_si512(_mm512_and_si512(t24, _mm512_set1_epi64(0xFFFFFF)),
                                              _mm512_slli_epi64(_mm512_srli_epi64(t24, 32), 24));
        const __m512i meta = TK_OR3(_mm512_slli_epi64(c8, 36), _mm512_slli_epi64(d8, 40), _mm512_slli_epi64(k8, 42));
        const __m512i w0 = _mm512_mask_blend_epi64(ok, all54, _mm512_or_si512(s1, s3));
        const __m512i w1 = _mm512_mask_blend_epi64(ok, all54, _mm512_or_si512(s2, s3));
        const __m512i w2 = _mm512_maskz_or_epi64(ok, order, meta);
        uint64_t* out = records + g * 3;
        const __m512i o0 = _mm512_mask_permutexvar_epi64(_mm512_permutex2var_epi64(w0, ia0, w1), (__mmask8)0x24, ib0, w2);
        const __m512i o1 = _mm512_mask_permutexvar_epi64(_mm512_permutex2var_epi64(w0, ia1, w1), (__mmask8)0x49, ib1, w2);
        const __m512i o2 = _mm512_mask_permutexvar_epi64(_mm512_permutex2var_epi64(w0, ia2, w1), (__mmask8)0x92, ib2, w2);
        if (aligned) { _mm512_stream_si512((__m512i*)out, o0); _mm512_stream_si512((__m512i*)(out + 8), o1); _mm512_stream_si512((__m512i*)(out + 16), o2); }
        else { _mm512_storeu_si512(out, o0); _mm512_storeu_si512(out + 8, o1); _mm512_storeu_si512(out + 16, o2); }
        bad += 8 - __builtin_popcount((unsigned)ok);
    }
    if (aligned) _mm_sfence();
    return bad + pack_range_scalar(perm, contract, declarer, king, g, b, records);
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

// ---- persistent pack pool: the chunked host pipeline of tarok_rollout_host_packed ---------------------------------------
// One job = the whole batch, cut into blocks of PACK_BLOCK rows that the workers (and the calling thread, while it waits)
// take in order from an atomic counter; a per-chunk counter tells the caller when every block of upload chunk c is packed, so
// it can enqueue that chunk's copy while the workers are already in chunk c+1 -- no barrier per chunk, no idle thread while
// the slowest one finishes.  Idle workers spin for a short while before they park on the condition variable: the jobs of
// consecutive calls arrive back to back.
constexpr uint64_t PACK_BLOCK = 2048;                      // rows; a multiple of 8 (the vector code's group)
constexpr int PACK_MAX_CHUNKS = 32;
constexpr uint64_t PACK_CLOSED = 1ull << 62;

struct tarok_pack_pool {
    std::vector<std::thread> workers;
    std::mutex mu;
    std::condition_variable cv_job;
    std::atomic<uint64_t> generation{0};
    std::atomic<bool> stop{false};
    int parts = 1;
    // the current job
    const uint8_t *perm = nullptr, *contract = nullptr, *declarer = nullptr, *king = nullptr;
    uint64_t* records = nullptr;
    uint64_t n = 0, nblocks = 0;
    int nchunks = 0;
    uint64_t bounds[PACK_MAX_CHUNKS + 1] = {};              // upload chunk c = rows [bounds[c], bounds[c + 1]); whole blocks except the last
    std::atomic<uint64_t> next{0};
    std::atomic<uint64_t> done[PACK_MAX_CHUNKS];            // blocks finished per upload chunk
    std::atomic<int64_t> bad{0};
    std::atomic<int> active{0};                             // workers inside work(): the job's fields must not change under them

    bool take_block() {                                     // packs one block; false when none is left
        const uint64_t i = next.fetch_add(1, std::memory_order_acq_rel);      // acquire: the job's fields were written before next = 0
        if (i >= nblocks) return false;
        const uint64_t lo = i * PACK_BLOCK, hi = lo + PACK_BLOCK < n ? lo + PACK_BLOCK : n;
        const int64_t nb = pack_range(perm, contract, declarer, king, lo, hi, records);
        if (nb) bad.fetch_add(nb, std::memory_order_relaxed);
        int c = 0;
        while (c + 1 < nchunks && lo >= bounds[c + 1]) c++;
        done[c].fetch_add(1, std::memory_order_release);
        return true;
    }
    void loop() {
        uint64_t seen = 0;
        for (;;) {
            int spins = 0;
            while (generation.load(std::memory_order_acquire) == seen && !stop.load(std::memory_order_relaxed)) {
                if (++spins < 20000) { _mm_pause(); continue; }
                std::unique_lock<std::mutex> lk(mu);
                cv_job.wait(lk, [&] { return stop.load() || generation.load() != seen; });
            }
            if (stop.load()) return;
            active.fetch_add(1, std::memory_order_acq_rel);
            seen = generation.load(std::memory_order_acquire);
            while (take_block()) {}                            // a job being set up has next = CLOSED: nothing to take
            active.fetch_sub(1, std::memory_order_release);
        }
    }
};

tarok_pack_pool* tarok_pack_pool_create(int threads) {
    if (threads < 1) threads = 1;
    tarok_pack_pool* p = new (std::nothrow) tarok_pack_pool();
    if (!p) return nullptr;
    p->parts = threads;
    for (auto& d : p->done) d.store(0);
    try {
        for (int t = 1; t < threads; t++) p->workers.emplace_back([p] { p->loop(); });
    } catch (...) {
        p->parts = (int)p->workers.size() + 1;          // fewer threads than asked for: still correct
    }
    return p;
}

void tarok_pack_pool_destroy(tarok_pack_pool* p) {
    if (!p) return;
    {
        std::lock_guard<std::mutex> lk(p->mu);
        p->stop.store(true);
    }
    p->cv_job.notify_all();
    for (auto& w : p->workers) w.join();
    delete p;
}

int tarok_pack_pool_threads(const tarok_pack_pool* p) { return p ? p->parts : 0; }

// The upload chunks of the host pipelines: `want` chunks over n rows, TAPERED -- the first and the last chunk get one unit,
// their neighbours two, the rest four -- because the pipeline's fill is the first chunk's upload and its drain the last
// chunk's play + download: small ends, large middle.  Boundaries are multiples of `quantum` rows (the pack block, itself a
// multiple of the CTA tile); bounds[0] = 0, bounds[returned count] = n.  A small batch is one chunk.
int tarok_chunk_bounds(uint64_t n, int want, uint64_t quantum, uint64_t* bounds) {
    if (want > PACK_MAX_CHUNKS) want = PACK_MAX_CHUNKS;
    bounds[0] = 0;
    if (want < 2 || n < (1ull << 18)) { bounds[1] = n; return 1; }
    uint64_t units[PACK_MAX_CHUNKS], total = 0;
    for (int c = 0; c < want; c++) {
        const int edge = c < want - 1 - c ? c : want - 1 - c;           // distance from the nearer end
        units[c] = want < 4 ? 1 : edge == 0 ? 1 : edge == 1 ? 2 : 4;    // steeper tapers (up to 16) measure the same
        total += units[c];
    }
    int count = 0;
    uint64_t acc = 0;
    for (int c = 0; c < want; c++) {
        acc += units[c];
        uint64_t b = (n * acc / total + quantum - 1) / quantum * quantum;
        if (b > n || c == want - 1) b = n;
        if (b > bounds[count]) bounds[++count] = b;
        if (b == n) break;
    }
    return count;
}

// Starts packing rows [0, n) into `records`; bounds[0..nchunks] (tarok_chunk_bounds with quantum = tarok_pack_block_rows())
// are the upload chunks that tarok_pack_pool_wait_chunk reports on.  Returns at once; the workers run in the background.
void tarok_pack_pool_begin(tarok_pack_pool* p, const uint8_t* perm, const uint8_t* contract, const uint8_t* declarer,
                           const uint8_t* king, uint64_t n, const uint64_t* bounds, int nchunks, uint64_t* records) {
    p->next.store(PACK_CLOSED, std::memory_order_release);                  // nobody can take a block while the job changes
    while (p->active.load(std::memory_order_acquire) != 0) _mm_pause();      // stragglers of the previous job
    p->perm = perm; p->contract = contract; p->declarer = declarer; p->king = king; p->records = records;
    p->n = n;
    p->nchunks = nchunks;
    for (int c = 0; c <= nchunks; c++) p->bounds[c] = bounds[c];
    p->nblocks = (n + PACK_BLOCK - 1) / PACK_BLOCK;
    for (auto& d : p->done) d.store(0, std::memory_order_relaxed);
    p->bad.store(0, std::memory_order_relaxed);
    p->next.store(0, std::memory_order_release);                            // opens the job
    {
        std::lock_guard<std::mutex> lk(p->mu);
        p->generation.fetch_add(1, std::memory_order_release);
    }
    p->cv_job.notify_all();
}

uint64_t tarok_pack_block_rows(void) { return PACK_BLOCK; }

// Blocks (packing blocks itself meanwhile) until every row of upload chunk c is in `records`.
void tarok_pack_pool_wait_chunk(tarok_pack_pool* p, int c) {
    const uint64_t lo = p->bounds[c], hi = p->bounds[c + 1];
    const uint64_t want = lo < hi ? (hi - lo + PACK_BLOCK - 1) / PACK_BLOCK : 0;
    while (p->done[c].load(std::memory_order_acquire) < want)
        if (!p->take_block()) _mm_pause();
}

int64_t tarok_pack_pool_bad(const tarok_pack_pool* p) { return p->bad.load(); }

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
