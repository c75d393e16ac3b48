// Counter-based Philox4x32-10 (Salmon et al., "Parallel random numbers: as easy as 1, 2, 3") and the
// unbiased bounded draw used for every synthetic input (DESIGN.md "Synthetic inputs"):
//
//   word(gid, stream, idx)    = philox4x32_10(key = seed, ctr = {gid.lo, gid.hi, stream | attempt<<16, idx>>2})[idx&3]
//   draw(gid, stream, idx, n) = Lemire multiply-shift with rejection -> exactly uniform on [0, n)
//
// Stateless: any (game, stream, index) can be regenerated anywhere, so results do not depend on how
// games are sharded over GPUs.  The reference has no seeds at all; parity is by
// exporting the deal/actions and replaying them through the oracle.
//
// Issue-slot notes (B200: the integer ALU pipe retires one warp instruction every 2 cycles per SM
// sub-partition and is what bounds the step kernel): the ten round keys depend only on the seed, so the
// host precomputes them into the kernel parameters (constant bank operands, no per-lane adds), and the
// 32x32 multiplies are written as mul.hi/mul.lo so they land on the FMA pipe.
#pragma once
#include <cstdint>

namespace tk {

using u64 = unsigned long long;
using u32 = unsigned int;

enum : u32 { ST_DEAL = 0, ST_BID = 1, ST_KING = 2, ST_EXCH = 3, ST_PLAY = 4, ST_FORCE = 5, ST_PLAY_RETRY = 6 };

constexpr u32 PHILOX_M0 = 0xD2511F53u, PHILOX_M1 = 0xCD9E8D57u, PHILOX_W0 = 0x9E3779B9u, PHILOX_W1 = 0xBB67AE85u;

// Seed + the 10 round keys (k0 + r*W0, k1 + r*W1); filled on the host (philox_keys_init).
struct Rng {
    u64 seed;
    u32 rk[20];
};

inline void philox_keys_init(Rng& r, u64 seed) {
    r.seed = seed;
    u32 k0 = (u32)seed, k1 = (u32)(seed >> 32);
    for (int i = 0; i < 10; i++) {
        r.rk[2 * i] = k0; r.rk[2 * i + 1] = k1;
        k0 += PHILOX_W0; k1 += PHILOX_W1;
    }
}

struct Words4 { u32 w[4]; };

__device__ __forceinline__ Words4 philox_block(const Rng& rng, u64 ctr01, u32 stream_attempt, u32 block) {
    u32 c0 = (u32)ctr01, c1 = (u32)(ctr01 >> 32), c2 = stream_attempt, c3 = block;
#pragma unroll
    for (int r = 0; r < 10; r++) {
        const u32 hi0 = __umulhi(PHILOX_M0, c0), lo0 = PHILOX_M0 * c0;
        const u32 hi1 = __umulhi(PHILOX_M1, c2), lo1 = PHILOX_M1 * c2;
        c0 = hi1 ^ c1 ^ rng.rk[2 * r];
        c2 = hi0 ^ c3 ^ rng.rk[2 * r + 1];
        c1 = lo1; c3 = lo0;
    }
    Words4 o;
    o.w[0] = c0; o.w[1] = c1; o.w[2] = c2; o.w[3] = c3;
    return o;
}

// Lemire: accept unless the low product word falls in the biased sliver (probability < n / 2^32).
__device__ __forceinline__ bool lemire(u32 x, u32 n, u32& out) {
    u64 m = (u64)x * n;
    u32 lo = (u32)m;
    out = (u32)(m >> 32);
    if (lo >= n) return true;
    u32 t = (0u - n) % n;
    return lo >= t;
}

// Cold path (never inlined): redraw with attempt = first, first+1, ... until accepted.  Bumps the key itself.
__device__ __noinline__ u32 draw_loop(u64 seed, u64 gid, u32 stream, u32 idx, u32 n, u32 first) {
    for (u32 attempt = first;; attempt++) {
        u32 c0 = (u32)gid, c1 = (u32)(gid >> 32), c2 = stream | (attempt << 16), c3 = idx >> 2;
        u32 k0 = (u32)seed, k1 = (u32)(seed >> 32);
        for (int r = 0; r < 10; r++) {
            const u32 hi0 = __umulhi(PHILOX_M0, c0), lo0 = PHILOX_M0 * c0;
            const u32 hi1 = __umulhi(PHILOX_M1, c2), lo1 = PHILOX_M1 * c2;
            c0 = hi1 ^ c1 ^ k0; c2 = hi0 ^ c3 ^ k1; c1 = lo1; c3 = lo0;
            k0 += PHILOX_W0; k1 += PHILOX_W1;
        }
        const u32 k = idx & 3u;
        const u32 x = k == 0 ? c0 : k == 1 ? c1 : k == 2 ? c2 : c3;
        u32 r;
        if (lemire(x, n, r)) return r;
    }
}

__device__ __forceinline__ u32 draw_from_word(u32 x, const Rng& rng, u64 gid, u32 stream, u32 idx, u32 n) {
    u32 r;
    if (__builtin_expect(!lemire(x, n, r), 0)) r = draw_loop(rng.seed, gid, stream, idx, n, 1u);
    return r;
}

// One-off draw (computes its own Philox block).
__device__ __forceinline__ u32 draw(const Rng& rng, u64 gid, u32 stream, u32 idx, u32 n) {
    Words4 b = philox_block(rng, gid, stream, idx >> 2);
    const u32 k = idx & 3u;
    const u32 x = k == 0 ? b.w[0] : k == 1 ? b.w[1] : k == 2 ? b.w[2] : b.w[3];
    return draw_from_word(x, rng, gid, stream, idx, n);
}

// ---- batched bounded draws -------------------------------------------------------------------------
// Several exact draws from ONE 32-bit word (Brackett-Rozinsky & Lemire, "Batched ranged random integer generation"):
//   r_1 = hi32(x * n_1), x <- lo32(x * n_1);  r_2 = hi32(x * n_2), x <- lo32(x * n_2); ...
// By induction floor(x0 * B / 2^32) = the mixed-radix number (r_1, ..., r_k) with B = n_1 * ... * n_k and the final x is
// (x0 * B) mod 2^32, so this is Lemire's multiply-shift for the bound B read digit by digit: the tuple is exactly uniform
// iff the final x >= 2^32 mod B.  A rejected word (probability < B / 2^32) is redrawn whole on the retry stream with
// attempt = 1, 2, ...  Two multiplies per draw, both on the FMA pipe, and one compare per word.
__device__ __forceinline__ u32 bdraw(u32& x, u32 n) { const u32 r = __umulhi(x, n); x *= n; return r; }

// Cold path: redraw word `word` of (gid, stream) until accepted; bounds8 = the k <= 8 bounds, 8 bits each (each <= 63);
// returns the k results, 6 bits each.
__device__ __noinline__ u64 bdraw_retry(u64 seed, u64 gid, u32 stream, u32 word, u64 bounds8, u32 k) {
    for (u32 attempt = 1;; attempt++) {
        u32 c0 = (u32)gid, c1 = (u32)(gid >> 32), c2 = stream | (attempt << 16), c3 = word >> 2;
        u32 k0 = (u32)seed, k1 = (u32)(seed >> 32);
        for (int r = 0; r < 10; r++) {
            const u32 hi0 = __umulhi(PHILOX_M0, c0), lo0 = PHILOX_M0 * c0;
            const u32 hi1 = __umulhi(PHILOX_M1, c2), lo1 = PHILOX_M1 * c2;
            c0 = hi1 ^ c1 ^ k0; c2 = hi0 ^ c3 ^ k1; c1 = lo1; c3 = lo0;
            k0 += PHILOX_W0; k1 += PHILOX_W1;
        }
        const u32 j = word & 3u;
        u32 x = j == 0 ? c0 : j == 1 ? c1 : j == 2 ? c2 : c3;
        u64 out = 0;
        u32 prod = 1;
        for (u32 i = 0; i < k; i++) {
            const u32 n = (u32)(bounds8 >> (8u * i)) & 0xFFu;
            out |= (u64)bdraw(x, n) << (6u * i);
            prod *= n;
        }
        if (x >= (0u - prod) % prod) return out;
    }
}

// ---- play draws -------------------------------------------------------------------------------------
// A legal mask holds at most 12 cards, so a play needs far fewer than 32 random bits.  One Philox block is
// shared by TWO consecutive games (pair = gid >> 1) and the FOUR plays of one trick: eight 16-bit lanes,
//   lane(g, t) = (g & 1) * 4 + (t & 3),   block = philox(key = seed, ctr = {pair.lo, pair.hi, ST_PLAY, t >> 2}).
// The stepwise kernel (one lane = one game pair) therefore needs ONE block per lane per step, the fused
// kernel (one lane = one game) one block per trick.  16-bit Lemire with rejection keeps the draw exactly
// uniform; the rejected sliver (probability < n / 65536) falls back to a 32-bit draw on stream ST_PLAY_RETRY.
__device__ __forceinline__ Words4 play_block(const Rng& rng, u64 gid, u32 trick) {
    return philox_block(rng, gid >> 1, ST_PLAY, trick);
}

// The 16-bit lane of (game, position in the trick) within the pair's block.
template <int POS = -1>
__device__ __forceinline__ u32 play_lane(const Words4& b, u64 gid, u32 t) {
    u32 w;
    if (POS >= 0) {                                        // position known at compile time (lock-step batches): one select
        w = ((u32)gid & 1u) ? b.w[2 + (POS >> 1)] : b.w[POS >> 1];
        return (POS & 1) ? (w >> 16) : (w & 0xFFFFu);
    }
    const u32 lane = ((u32)gid & 1u) * 4u + (t & 3u);
    w = (lane >> 1) == 0 ? b.w[0] : (lane >> 1) == 1 ? b.w[1] : (lane >> 1) == 2 ? b.w[2] : b.w[3];
    return (lane & 1u) ? (w >> 16) : (w & 0xFFFFu);
}

// Uniform draw on [0, n) from the 16-bit lane x (n <= 12 on this path).
__device__ __forceinline__ u32 play_pick(u32 x, const Rng& rng, u64 gid, u32 t, u32 n) {
    const u32 m = x * n;
    const u32 lo = m & 0xFFFFu;
    if (__builtin_expect(lo < n, 0)) {
        // 65536 % n for n = 0..15, four bits each
        const u32 rem = (u32)((0x1234967024101000ull >> (4u * (n & 15u))) & 15ull);
        if (lo < rem) return draw_loop(rng.seed, gid, ST_PLAY_RETRY, t, n, 0u);
    }
    return m >> 16;
}

template <int POS = -1>
__device__ __forceinline__ u32 play_draw(const Words4& b, const Rng& rng, u64 gid, u32 t, u32 n) {
    return play_pick(play_lane<POS>(b, gid, t), rng, gid, t, n);
}

// The lanes of trick positions 1..3 for the two games of a pair (even game in the low half): what the position-0 launch
// of a lock-step chain leaves behind for the three launches that follow (tarok_kernels.cuh "draw cache").
__device__ __forceinline__ u32 play_lanes_of_pair(const Words4& b, int pos) {
    return pos == 1 ? (b.w[0] >> 16) | (b.w[2] & 0xFFFF0000u)
         : pos == 2 ? (b.w[1] & 0xFFFFu) | (b.w[3] << 16)
                    : (b.w[1] >> 16) | (b.w[3] & 0xFFFF0000u);
}

}  // namespace tk
