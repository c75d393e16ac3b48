// Host-side half of libtarok_b200.so (plain C++, compiled by g++): the serialiser that turns the permutation rows a
// patched Igra.shuffle produces (Igra.py:65-73: uint8 [n,54]) + forced contracts into 20-byte deal records
// (layout in include/tarok_b200.h), so the host-buffer entry moves 20 instead of 57 bytes per deal over PCIe.
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

constexpr int REC = TAROK_RECORD_BYTES;                    // 20

// One row -> one record (layout "deal records" in include/tarok_b200.h): the owner sets are built with independent
// accumulators (a load, a shift and an OR per card); plane 0 = seats 1|3, plane 1 = seats 2|3; the six talon ids as they
// stand (6 bits each, talon order) and the forced contract make up the 45-bit word m, which fills the ten spare bits of
// each plane word and a third, 32-bit word.
constexpr uint64_t ERR_PLANE = ALL54;                      // both planes full: every card "belongs to seat 3" -> decodes to an error game
inline void store_record(uint8_t* out, uint64_t p0, uint64_t p1, uint64_t m) {
    const uint64_t w0 = p0 | (m & 0x3FF) << 54, w1 = p1 | ((m >> 10) & 0x3FF) << 54;
    const uint32_t w2 = (uint32_t)(m >> 20);
    memcpy(out, &w0, 8); memcpy(out + 8, &w1, 8); memcpy(out + 16, &w2, 4);
}
__attribute__((always_inline)) inline bool pack_row(const uint8_t* row, unsigned contract, unsigned declarer, unsigned king, uint8_t* out) {
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
    if (ok) store_record(out, h[1] | h[3], h[2] | h[3], order | (uint64_t)contract << 36 | (uint64_t)declarer << 40 | (uint64_t)k << 42);
    else store_record(out, ERR_PLANE, ERR_PLANE, 0);
    return ok;
}

// Built twice (function multi-versioning, resolved once at load time): with BMI2 the variable shifts are single SHLX
// micro-ops instead of the three-uop SHL-by-CL.
__attribute__((target_clones("default", "arch=x86-64-v3")))
int64_t pack_range_scalar(const uint8_t* perm, const uint8_t* contract, const uint8_t* declarer, const uint8_t* king, uint64_t a,
                          uint64_t b, uint8_t* records) {
    int64_t bad = 0;
    for (uint64_t g = a; g < b; g++)
        bad += pack_row(perm + g * 54, contract[g], declarer[g], king ? king[g] : 7u, records + g * REC) ? 0 : 1;
    return bad;
}

// AVX-512 version, one row per 64-bit lane, so nothing is ever reduced across lanes: seven gathers fetch bytes 8j..8j+7 of
// each of eight rows; position 8j+k of every row is then one byte move to the bottom of the lane (a byte shuffle on port
// 5, for two positions in eight a shift on port 0), one variable rotate of the constant 1 (VPROLVQ takes its count modulo
// 64 from the low six bits, so the other bytes of the lane need no masking) and one OR into the accumulator of the
// position's owner.  ids >= 64 are caught by OR-ing the raw bytes, ids 54..63 and duplicates by the cover test.  Sixteen
// rows (two such groups) make 320 bytes of records = five whole cache lines, assembled with two-source byte permutes
// (AVX512-VBMI) and written with non-temporal stores when the buffer is 64-byte aligned (no read-for-ownership of lines the
// CPU never reads back; the DMA engine does).  The last gather of a row reads two bytes of the next row, so the caller
// keeps at least one row behind every group (the tail goes through the scalar code).
#define TK_AVX512 __attribute__((target("avx512f,avx512bw,avx512vl,avx512dq,avx512vbmi,bmi2,popcnt")))
template <int K> TK_AVX512 static inline __m512i one_hot_at(__m512i v) {        // 1 << (byte K of each lane & 63)
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
// Eight rows -> the three record words per lane (w2 in the low 32 bits of its lane); returns the mask of valid rows.
TK_AVX512 static inline __mmask8 pack8(const uint8_t* base, const uint8_t* contract, const uint8_t* declarer, const uint8_t* king,
                                       __m512i& w0, __m512i& w1, __m512i& w2) {
    const __m512i row_off = _mm512_setr_epi64(0, 54, 108, 162, 216, 270, 324, 378);
    const __m512i all54 = _mm512_set1_epi64((long long)ALL54);
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
    const __m512i c8 = _mm512_cvtepu8_epi64(_mm_loadl_epi64((const __m128i*)contract));
    const __m512i d8 = _mm512_cvtepu8_epi64(_mm_loadl_epi64((const __m128i*)declarer));
    const __m512i k8 = king ? _mm512_min_epu64(_mm512_cvtepu8_epi64(_mm_loadl_epi64((const __m128i*)king)), _mm512_set1_epi64(7))
                            : _mm512_set1_epi64(7);
    const __mmask8 ok = _mm512_cmpeq_epi64_mask(cover, all54)
                      & _mm512_testn_epi64_mask(raw, _mm512_set1_epi64((long long)0xC0C0C0C0C0C0C0C0ull))
                      & _mm512_cmple_epu64_mask(c8, _mm512_set1_epi64(15)) & _mm512_cmple_epu64_mask(d8, _mm512_set1_epi64(3));
    // the six talon ids -> 36 contiguous bits: byte pairs (x1, x64) -> 12-bit fields, field pairs (x1, x4096) -> 24 bits
    const __m512i t12 = _mm512_maddubs_epi16(v6, _mm512_set1_epi16(0x4001));
    const __m512i t24 = _mm512_madd_epi16(t12, _mm512_set1_epi32(0x10000001));
    const __m512i order = _mm512_or_si512(_mm512_and_si512(t24, _mm512_set1_epi64(0xFFFFFF)),
                                          _mm512_slli_epi64(_mm512_srli_epi64(t24, 32), 24));
    const __m512i m = _mm512_maskz_or_epi64(ok, order, TK_OR3(_mm512_slli_epi64(c8, 36), _mm512_slli_epi64(d8, 40), _mm512_slli_epi64(k8, 42)));
    const __m512i top = _mm512_set1_epi64((long long)(0x3FFull << 54));
    w0 = _mm512_or_si512(_mm512_mask_blend_epi64(ok, all54, _mm512_or_si512(s1, s3)), _mm512_and_si512(_mm512_slli_epi64(m, 54), top));
    w1 = _mm512_or_si512(_mm512_mask_blend_epi64(ok, all54, _mm512_or_si512(s2, s3)), _mm512_and_si512(_mm512_slli_epi64(m, 44), top));
    w2 = _mm512_srli_epi64(m, 20);
    return ok;
}

// Byte-permute tables that lay sixteen rows' (w0, w1, w2) out as 16 x 20 bytes = five 64-byte vectors: output byte k of
// the block belongs to row k / 20 (half = row / 8, lane = row % 8) at offset k % 20 -- 0..7 from w0, 8..15 from w1, 16..19
// from w2.  Per output vector and half: ab = index into the concatenation (w0, w1) for VPERMI2B, c = index into w2, cm =
// bytes taken from w2, hm = bytes that belong to the second half.
struct RecTables { alignas(64) uint8_t ab[5][2][64]; alignas(64) uint8_t c[5][2][64]; uint64_t cm[5][2]; uint64_t hm[5]; };
constexpr RecTables make_rec_tables() {
    RecTables t{};
    for (int j = 0; j < 5; j++)
        for (int i = 0; i < 64; i++) {
            const int k = 64 * j + i, row = k / REC, off = k % REC, half = row / 8, lane = row % 8;
            if (half) t.hm[j] |= 1ull << i;
            if (off < 8) t.ab[j][half][i] = (uint8_t)(lane * 8 + off);
            else if (off < 16) t.ab[j][half][i] = (uint8_t)(64 + lane * 8 + off - 8);
            else { t.c[j][half][i] = (uint8_t)(lane * 8 + off - 16); t.cm[j][half] |= 1ull << i; }
        }
    return t;
}
alignas(64) static const RecTables REC_TABLES = make_rec_tables();

template <int J, int HALF> TK_AVX512 static inline __m512i rec_bytes(__m512i w0, __m512i w1, __m512i w2) {
    const __m512i t = _mm512_permutex2var_epi8(w0, _mm512_load_si512(REC_TABLES.ab[J][HALF]), w1);
    return _mm512_mask_permutexvar_epi8(t, (__mmask64)REC_TABLES.cm[J][HALF], _mm512_load_si512(REC_TABLES.c[J][HALF]), w2);
}

TK_AVX512 static int64_t pack_range_avx512(const uint8_t* perm, const uint8_t* contract, const uint8_t* declarer,
                                           const uint8_t* king, uint64_t a, uint64_t b, uint8_t* records) {
    int64_t bad = 0;
    uint64_t g = a;
    // a block of sixteen records is five whole cache lines when the buffer is 64-byte aligned and the block starts at a
    // multiple of sixteen rows
    const bool aligned = ((uintptr_t)records & 63) == 0;
    if (aligned && (g & 15)) {
        const uint64_t peel = (g + 15) & ~15ull;
        const uint64_t upto = peel < b ? peel : b;
        bad += pack_range_scalar(perm, contract, declarer, king, g, upto, records);
        g = upto;
    }
    for (; g + 16 < b; g += 16) {
        const uint8_t* base = perm + g * 54;
        for (int i = 0; i < 14; i++) _mm_prefetch((const char*)(base + 864 * 6 + 64 * i), _MM_HINT_T0);   // the gathers' lines, 6 blocks ahead
        __m512i a0, b0, c0, a1, b1, c1;
        const __mmask8 ok0 = pack8(base, contract + g, declarer + g, king ? king + g : nullptr, a0, b0, c0);
        const __mmask8 ok1 = pack8(base + 432, contract + g + 8, declarer + g + 8, king ? king + g + 8 : nullptr, a1, b1, c1);
        const __m512i o0 = rec_bytes<0, 0>(a0, b0, c0), o1 = rec_bytes<1, 0>(a0, b0, c0);
        const __m512i o2 = _mm512_mask_blend_epi8((__mmask64)REC_TABLES.hm[2], rec_bytes<2, 0>(a0, b0, c0), rec_bytes<2, 1>(a1, b1, c1));
        const __m512i o3 = rec_bytes<3, 1>(a1, b1, c1), o4 = rec_bytes<4, 1>(a1, b1, c1);
        uint8_t* out = records + g * REC;
        if (aligned) {
            _mm512_stream_si512((__m512i*)out, o0); _mm512_stream_si512((__m512i*)(out + 64), o1); _mm512_stream_si512((__m512i*)(out + 128), o2);
            _mm512_stream_si512((__m512i*)(out + 192), o3); _mm512_stream_si512((__m512i*)(out + 256), o4);
        } else {
            _mm512_storeu_si512(out, o0); _mm512_storeu_si512(out + 64, o1); _mm512_storeu_si512(out + 128, o2);
            _mm512_storeu_si512(out + 192, o3); _mm512_storeu_si512(out + 256, o4);
        }
        bad += 16 - __builtin_popcount((unsigned)ok0) - __builtin_popcount((unsigned)ok1);
    }
    if (aligned) _mm_sfence();
    return bad + pack_range_scalar(perm, contract, declarer, king, g, b, records);
}

std::atomic<int> g_force_scalar{0};

bool cpu_is_wide() {
    return __builtin_cpu_supports("avx512f") && __builtin_cpu_supports("avx512bw") && __builtin_cpu_supports("avx512vl")
        && __builtin_cpu_supports("avx512dq") && __builtin_cpu_supports("avx512vbmi") && __builtin_cpu_supports("bmi2");
}

int64_t pack_range(const uint8_t* perm, const uint8_t* contract, const uint8_t* declarer, const uint8_t* king, uint64_t a,
                   uint64_t b, uint8_t* records) {
    static const bool wide = cpu_is_wide();
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
    uint8_t* records = nullptr;
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
                           const uint8_t* king, uint64_t n, const uint64_t* bounds, int nchunks, void* records) {
    p->next.store(PACK_CLOSED, std::memory_order_release);                  // nobody can take a block while the job changes
    while (p->active.load(std::memory_order_acquire) != 0) _mm_pause();      // stragglers of the previous job
    p->perm = perm; p->contract = contract; p->declarer = declarer; p->king = king; p->records = (uint8_t*)records;
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
                              uint64_t n, void* records_out, int threads) {
    if (!perm || !contract || !declarer || !records_out) return -1;
    uint8_t* records = (uint8_t*)records_out;
    if (threads < 1) threads = 1;
    if ((uint64_t)threads > n / 4096 + 1) threads = (int)(n / 4096 + 1);      // not worth a thread below 4096 rows
    if (threads == 1) return pack_range(perm, contract, declarer, king, 0, n, records);
    std::vector<int64_t> bad((size_t)threads, 0);
    std::vector<std::thread> pool;
    pool.reserve((size_t)threads - 1);
    const uint64_t per = ((n + (uint64_t)threads - 1) / (uint64_t)threads + 15) & ~15ull;   // slices start on 16-row blocks
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

int tarok_pack_uses_avx512(void) { return g_force_scalar.load() ? 0 : (cpu_is_wide() ? 1 : 0); }

int64_t tarok_pack_records(const uint8_t* perm, const uint8_t* contract, const uint8_t* declarer, const uint8_t* king,
                           uint64_t n, void* records) {
    return tarok_pack_records_mt(perm, contract, declarer, king, n, records, 1);
}

}  // extern "C"
