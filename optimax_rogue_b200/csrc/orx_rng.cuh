// Philox4x32-10 and the shared draw schedule (see DESIGN.md "Draw schedule").
// Every reference random call (randombot.py:21, updater.py:114,127, worldgen.py:39-40,
// world.py:62) consumes exactly one 32-bit word, mapped with bounded(w, n) = (w*n) >> 32.
#pragma once
#include <stdint.h>

namespace orx {

enum : uint32_t { DOM_TICK = 0, DOM_LEVEL = 1, DOM_RESET = 2 };
enum : uint32_t { SUB_MAIN = 0, SUB_NPC = 1, SUB_DESCEND = 64, MAX_TRIES = 256 };

// The ten round keys (seed + r * Weyl constants) are the same for every game of a launch: the host
// computes them once into the kernel parameter block, so a round is two wide multiplies and two
// three-input XORs against constant-bank operands.
struct RoundKeys { uint32_t k[20]; };   // k[2r], k[2r+1]

struct Stream {
    const RoundKeys* rk;
    uint32_t g0, g1;      // global game id (g1 < 2^22)
    uint32_t episode;
};

__host__ __device__ inline void make_round_keys(RoundKeys& rk, uint32_t k0, uint32_t k1)
{
    for (int r = 0; r < 10; ++r) { rk.k[2 * r] = k0; rk.k[2 * r + 1] = k1; k0 += 0x9E3779B9u; k1 += 0xBB67AE85u; }
}

__device__ __forceinline__ uint4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, const RoundKeys& rk)
{
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        c0 = hi1 ^ c1 ^ rk.k[2 * r];
        c1 = lo1;
        c2 = hi0 ^ c3 ^ rk.k[2 * r + 1];
        c3 = lo0;
    }
    return make_uint4(c0, c1, c2, c3);
}

__device__ __forceinline__ uint4 draw_block(const Stream& s, uint32_t domain, uint32_t sub, uint32_t index)
{
    const uint32_t c1 = (s.g1 & 0x3FFFFFu) | ((sub & 0xFFu) << 22) | (domain << 30);
    return philox4x32_10(s.g0, c1, s.episode, index, *s.rk);
}

__device__ __forceinline__ uint32_t bounded(uint32_t w, uint32_t n) { return __umulhi(w, n); }

__device__ __forceinline__ uint32_t word_of(const uint4& b, int k)
{
    return k == 0 ? b.x : k == 1 ? b.y : k == 2 ? b.z : b.w;
}

// bounded value of draw q of the sequence that starts at block sub_base (rare paths only)
__device__ __forceinline__ uint32_t seq_bounded(const Stream& s, uint32_t domain, uint32_t sub_base,
                                                uint32_t index, int q, uint32_t n)
{
    if (q < (int)MAX_TRIES) {
        const uint4 b = draw_block(s, domain, sub_base + (uint32_t)(q >> 2), index);
        return bounded(word_of(b, q & 3), n);
    }
    const uint4 b = draw_block(s, domain, sub_base + ((MAX_TRIES - 1) >> 2), index);
    return (bounded(word_of(b, (MAX_TRIES - 1) & 3), n) + (uint32_t)(q - (int)(MAX_TRIES - 1))) % n;
}

}  // namespace orx
