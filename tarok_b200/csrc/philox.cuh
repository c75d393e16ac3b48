// Counter-based Philox4x32-10 (Salmon et al., "Parallel random numbers: as easy as 1, 2, 3") and the
// unbiased bounded draw used for every synthetic input (DESIGN.md "Synthetic inputs"):
//
//   word(gid, stream, idx)    = philox4x32_10(key = seed, ctr = {gid.lo, gid.hi, stream | attempt<<16, idx>>2})[idx&3]
//   draw(gid, stream, idx, n) = Lemire multiply-shift with rejection -> exactly uniform on [0, n)
//
// Stateless: any (game, stream, index) can be regenerated anywhere, so results do not depend on how
// games are sharded over GPUs.  The reference has no seeds at all; parity is by
// exporting the deal/actions and replaying them through the oracle.
#pragma once
#include <cstdint>

namespace tk {

using u64 = unsigned long long;
using u32 = unsigned int;

enum : u32 { ST_DEAL = 0, ST_BID = 1, ST_KING = 2, ST_EXCH = 3, ST_PLAY = 4, ST_FORCE = 5 };

__device__ __forceinline__ void philox4x32_10(u32& c0, u32& c1, u32& c2, u32& c3, u32 k0, u32 k1) {
#pragma unroll
    for (int r = 0; r < 10; r++) {
        u64 p0 = (u64)0xD2511F53u * c0;
        u64 p1 = (u64)0xCD9E8D57u * c2;
        u32 n0 = (u32)(p1 >> 32) ^ c1 ^ k0;
        u32 n2 = (u32)(p0 >> 32) ^ c3 ^ k1;
        c1 = (u32)p1; c3 = (u32)p0; c0 = n0; c2 = n2;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
}

struct Words4 { u32 w[4]; };

__device__ __forceinline__ Words4 philox_block(u64 seed, u64 gid, u32 stream_attempt, u32 block) {
    Words4 o;
    o.w[0] = (u32)gid; o.w[1] = (u32)(gid >> 32); o.w[2] = stream_attempt; o.w[3] = block;
    philox4x32_10(o.w[0], o.w[1], o.w[2], o.w[3], (u32)seed, (u32)(seed >> 32));
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

// Cold path: redraw with attempt = first, first+1, ... until accepted.
__device__ __noinline__ u32 draw_loop(u64 seed, u64 gid, u32 stream, u32 idx, u32 n, u32 first) {
    for (u32 attempt = first;; attempt++) {
        Words4 b = philox_block(seed, gid, stream | (attempt << 16), idx >> 2);
        u32 r;
        u32 x = (idx & 3u) == 0 ? b.w[0] : (idx & 3u) == 1 ? b.w[1] : (idx & 3u) == 2 ? b.w[2] : b.w[3];
        if (lemire(x, n, r)) return r;
    }
}

__device__ __forceinline__ u32 draw_retry(u64 seed, u64 gid, u32 stream, u32 idx, u32 n) { return draw_loop(seed, gid, stream, idx, n, 1u); }
__device__ __forceinline__ u32 draw_retry0(u64 seed, u64 gid, u32 stream, u32 idx, u32 n) { return draw_loop(seed, gid, stream, idx, n, 0u); }

__device__ __forceinline__ u32 draw_from_word(u32 x, u64 seed, u64 gid, u32 stream, u32 idx, u32 n) {
    u32 r;
    if (__builtin_expect(!lemire(x, n, r), 0)) r = draw_retry(seed, gid, stream, idx, n);
    return r;
}

// One-off draw (computes its own Philox block).
__device__ __forceinline__ u32 draw(u64 seed, u64 gid, u32 stream, u32 idx, u32 n) {
    Words4 b = philox_block(seed, gid, stream, idx >> 2);
    u32 k = idx & 3u;
    u32 x = k == 0 ? b.w[0] : k == 1 ? b.w[1] : k == 2 ? b.w[2] : b.w[3];
    return draw_from_word(x, seed, gid, stream, idx, n);
}

// ---- play draws -------------------------------------------------------------------------------------
// A legal mask holds at most 12 cards, so a play needs far fewer than 32 random bits.  One Philox block is
// shared by TWO consecutive games (pair = gid >> 1) and the FOUR plays of one trick: eight 16-bit lanes,
//   lane(g, t) = (g & 1) * 4 + (t & 3),   block = philox(key = seed, ctr = {pair.lo, pair.hi, ST_PLAY, t >> 2}).
// The stepwise kernel (one lane = one game pair) therefore needs ONE block per lane per step, the fused
// kernel (one lane = one game) one block per trick.  16-bit Lemire with rejection keeps the draw exactly
// uniform; the rejected sliver (probability < n / 65536) falls back to a 32-bit draw on stream ST_PLAY_RETRY.
enum : u32 { ST_PLAY_RETRY = 6 };

__device__ __forceinline__ Words4 play_block(u64 seed, u64 gid, u32 trick) {
    return philox_block(seed, gid >> 1, ST_PLAY, trick);
}

__device__ __forceinline__ u32 play_draw(const Words4& b, u64 seed, u64 gid, u32 t, u32 n) {
    u32 lane = ((u32)gid & 1u) * 4u + (t & 3u);
    u32 w = (lane >> 1) == 0 ? b.w[0] : (lane >> 1) == 1 ? b.w[1] : (lane >> 1) == 2 ? b.w[2] : b.w[3];
    u32 x = (lane & 1u) ? (w >> 16) : (w & 0xFFFFu);
    u32 m = x * n;
    u32 lo = m & 0xFFFFu;
    if (__builtin_expect(lo < n, 0)) {
        if (lo < (65536u % n)) return draw_retry0(seed, gid, ST_PLAY_RETRY, t, n);
    }
    return m >> 16;
}

}  // namespace tk
