// Ruleset R1 (docs/RULESET_R1.md): the README-only rules of optimax_rogue (readme.md:44-48,69-74).
// PARITY UNPINNED with respect to the reference (it has no code for them); bit-exact against
// oracle/orx_r1_oracle.c.
//
// Mapping: sixteen lanes (half a warp) own one game -- lane 0/1 the players, 2..9 the enemy slots,
// 10..13 the ground items. Entity planes are [n][16] words, so a warp reads two games as one
// 128-byte line. Everything that couples entities is a warp primitive on the half-warp mask:
// "who stands on my target tile" is a broadcast-and-compare sweep over the ten movers, "who else
// wants my target tile" is __match_any_sync on the packed (depth, y, x) key, damage is gathered by
// shuffles, free-slot and occupancy queries are ballots. Levels are not stored: walls are a hash of
// (x, y, level key), the key and the staircase come from the Philox LEVEL block.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/orx.h"
#include "orx_rng.cuh"

using namespace orx;

namespace {

constexpr int kThreadsR1 = 256;
constexpr int NM = 10;
enum : uint32_t { SUB_SPAWN = 8, SUB_DROP = 10, SUB_SPAWN_TRY = 16, R1_MAX_TRIES = 64 };

struct R1Params {
    int W, H, max_ticks, auto_reset, wall_density;
    RoundKeys rk;
    uint32_t *ent_loc, *ent_stat, *pl_a, *pl_b, *pl_c, *lvl_stairs, *lvl_key, *sep, *episode;
    int *ent_depth, *tick;
    uint8_t* status;
    unsigned int n;
    unsigned long long gid_base;
};

struct Grp {                 // half-warp context
    uint32_t mask;           // participating lanes of the warp
    int base;                // first warp lane of the group
    int l;                   // lane inside the game, 0..15
    template <typename T> __device__ __forceinline__ T bc(T v, int j) const { return __shfl_sync(mask, v, base + j); }
    __device__ __forceinline__ uint32_t ballot(bool p) const { return (__ballot_sync(mask, p) >> base) & 0xFFFFu; }
};

// per-lane state + replicated game-level state
struct R1Lane {
    int x, y, depth, alive, hp, aux;                                   // every lane
    int max_hp, max_mana, xp, level, n_items, cd, damage, armor;       // lanes 0, 1
    int sx0, sy0, sx1, sy1; uint32_t key0, key1;                       // replicated
    int sep, tick; uint32_t episode;
};

__device__ __forceinline__ uint32_t mix(uint32_t x, uint32_t y, uint32_t key)
{
    uint32_t h = x * 0x9E3779B1u ^ y * 0x85EBCA77u ^ key;
    h ^= h >> 15; h *= 0x2C1B3C6Du; h ^= h >> 12; h *= 0x297A2D39u; h ^= h >> 15;
    return h;
}
__device__ __forceinline__ bool is_wall(const R1Params& P, uint32_t key, int sx, int sy, int x, int y)
{
    if (x <= 0 || y <= 0 || x >= P.W - 1 || y >= P.H - 1) return true;
    if (x == sx && y == sy) return false;
    return (int)(mix((uint32_t)x, (uint32_t)y, key) & 255u) < P.wall_density;
}
__device__ __forceinline__ void level_init(const R1Params& P, const Stream& s, int depth, int& sx, int& sy, uint32_t& key)
{
    const uint4 b = draw_block(s, DOM_LEVEL, 0, (uint32_t)depth);
    sx = 1 + (int)bounded(b.x, (uint32_t)(P.W - 3));
    sy = 1 + (int)bounded(b.y, (uint32_t)(P.H - 3));
    key = b.z;
}
// Every lane of the group runs this with the same arguments; `blocker` is this lane's own
// "I am a living player/enemy" flag. Returns x | y << 8.
__device__ __noinline__ uint32_t free_tile(const R1Params& P, const Grp g, Stream s, uint32_t domain, uint32_t sub_base,
                                           uint32_t index, int depth, uint32_t key, int sx, int sy,
                                           bool blocker, int bx, int by, int bdepth)
{
    for (uint32_t r = 0; r < R1_MAX_TRIES; ++r) {
        const uint4 b = draw_block(s, domain, sub_base + (r >> 1), index);
        const int x = 1 + (int)bounded((r & 1) ? b.z : b.x, (uint32_t)(P.W - 2));
        const int y = 1 + (int)bounded((r & 1) ? b.w : b.y, (uint32_t)(P.H - 2));
        const bool bad = is_wall(P, key, sx, sy, x, y) || (x == sx && y == sy);
        const uint32_t occ = g.ballot(blocker && bdepth == depth && bx == x && by == y);
        if (!bad && occ == 0) return (uint32_t)x | ((uint32_t)y << 8);
    }
    for (int x = 1; x < P.W - 1; ++x)
        for (int y = 1; y < P.H - 1; ++y) {
            const bool bad = is_wall(P, key, sx, sy, x, y) || (x == sx && y == sy);
            const uint32_t occ = g.ballot(blocker && bdepth == depth && bx == x && by == y);
            if (!bad && occ == 0) return (uint32_t)x | ((uint32_t)y << 8);
        }
    return 1u | (1u << 8);
}

__device__ __forceinline__ void setup_game(const R1Params& P, const Grp g, R1Lane& L, const Stream& s)
{
    level_init(P, s, 0, L.sx0, L.sy0, L.key0);
    L.sx1 = L.sx0; L.sy1 = L.sy0; L.key1 = L.key0;
    L.alive = 0; L.x = 0; L.y = 0; L.depth = 0; L.hp = 0; L.aux = 0;
    for (int p = 0; p < 2; ++p) {
        const uint32_t t = free_tile(P, g, s, DOM_RESET, 32u * p, 0u, 0, L.key0, L.sx0, L.sy0,
                                     L.alive && g.l < NM, L.x, L.y, L.depth);
        if (g.l == p) {
            L.x = t & 255; L.y = t >> 8; L.depth = 0; L.alive = 1; L.hp = 10; L.aux = 9;
            L.max_hp = 10; L.max_mana = 9; L.xp = 0; L.level = 1; L.n_items = 0; L.cd = 0; L.damage = 2; L.armor = 1;
        }
    }
    L.sep = 0; L.tick = 1;
}

__device__ __forceinline__ uint32_t pos_key(int depth, int x, int y) { return ((uint32_t)depth << 16) | ((uint32_t)y << 8) | (uint32_t)x; }

struct R1Counters { unsigned int ticks, p1, p2, ties, descents, hits; };

// One tick for the group's game. c1/c2 are the players' commands (group-uniform values).
__device__ __forceinline__ int r1_tick(const R1Params& P, const Grp g, R1Lane& L, const Stream& s, int c1, int c2, R1Counters& cnt)
{
    const int l = g.l;
    const bool mover = l < NM;
    const int cd_pre = L.cd;
    int dx = 0, dy = 0;
    // player positions / depths, known to every lane
    const int p0x = g.bc(L.x, 0), p0y = g.bc(L.y, 0), p0d = g.bc(L.depth, 0);
    const int p1x = g.bc(L.x, 1), p1y = g.bc(L.y, 1), p1d = g.bc(L.depth, 1);
    // 1. heal + 2. intents
    if (l < 2) {
        int c = l == 0 ? c1 : c2;
        if (c == ORX_MOVE_HEAL) {
            const int h = min(L.aux, L.max_mana / 3);
            L.hp = min(L.max_hp, L.hp + h);
            L.aux -= h;
            c = ORX_MOVE_STAY;
        }
        const int ddx = (c == ORX_MOVE_RIGHT) - (c == ORX_MOVE_LEFT), ddy = (c == ORX_MOVE_DOWN) - (c == ORX_MOVE_UP);
        const uint32_t key = l == 0 ? L.key0 : L.key1;
        const int sx = l == 0 ? L.sx0 : L.sx1, sy = l == 0 ? L.sy0 : L.sy1;
        if ((ddx | ddy) != 0 && !is_wall(P, key, sx, sy, L.x + ddx, L.y + ddy)) { dx = ddx; dy = ddy; }
    } else if (mover && L.alive) {      // decide_npc_move: chase the nearest player on this depth
        const int m0 = p0d == L.depth ? abs(p0x - L.x) + abs(p0y - L.y) : (1 << 30);
        const int m1 = p1d == L.depth ? abs(p1x - L.x) + abs(p1y - L.y) : (1 << 30);
        if (min(m0, m1) < (1 << 30)) {
            const bool t1 = m1 < m0;
            const int ex = (t1 ? p1x : p0x) - L.x, ey = (t1 ? p1y : p0y) - L.y;
            if (max(abs(ex), abs(ey)) <= 6) {
                int ddx = 0, ddy = 0;
                if (abs(ex) > abs(ey)) ddx = ex > 0 ? 1 : -1; else ddy = ey > 0 ? 1 : -1;
                const bool lp0 = p0d == L.depth;
                const uint32_t key = lp0 ? L.key0 : L.key1;
                const int sx = lp0 ? L.sx0 : L.sx1, sy = lp0 ? L.sy0 : L.sy1;
                const int tx = L.x + ddx, ty = L.y + ddy;
                if (!is_wall(P, key, sx, sy, tx, ty) && !(tx == sx && ty == sy)) { dx = ddx; dy = ddy; }
            }
        }
    }
    bool has = mover && L.alive && (dx | dy) != 0;
    const uint32_t pkey = (mover && L.alive) ? pos_key(L.depth, L.x, L.y) : 0xFFFFFFFFu;
    uint32_t tkey = has ? pos_key(L.depth, L.x + dx, L.y + dy) : (0xFFFF0000u | (uint32_t)l);
    // occupant of my target tile at tick start: sweep the ten movers
    int occ = -1;
#pragma unroll
    for (int j = 0; j < NM; ++j) {
        const uint32_t pj = g.bc(pkey, j);
        if (has && tkey == pj) occ = j;
    }
    // 3. cooldown conversion
    if (l < 2 && cd_pre > 0 && has && occ >= 0) { has = false; dx = 0; dy = 0; tkey = 0xFFFF0000u | (uint32_t)l; occ = -1; }
    // 4. attacks
    const int my_dmg = l < 2 ? L.damage + min(L.aux, L.max_mana / 3) : 2 + L.depth / 4;
    const int my_arm = l < 2 ? L.armor : 0;
    const int src = occ >= 0 ? occ : l;
    const bool o_has = g.bc(has, src) != 0;
    const uint32_t o_tkey = g.bc(tkey, src);
    const int o_cd = g.bc(cd_pre, src);
    const uint32_t same_t = (__match_any_sync(g.mask, tkey) >> g.base) & 0x3FFu & ~(1u << l);
    int victim = -1, amount = 0, newcd = 0;
    bool spend = false, moves = false;
    if (has) {
        if (occ >= 0) {
            if (!(l >= 2 && occ >= 2)) {
                if (!o_has) {
                    if (occ < 2 && o_cd == 0) { if (l < 2) { newcd = 1; spend = true; } }     // negated
                    else { victim = occ; amount = 1; spend = l < 2; }                        // full
                } else if (o_tkey == pkey) { victim = occ; amount = 2; if (l < 2) { newcd = 3; spend = true; } }   // mutual: half
            }
        } else if (same_t == 0) {
            moves = true;
        } else {
            const uint32_t cand = l < 2 ? same_t : (same_t & 3u);
            if (cand != 0 && !(l < 2 && cd_pre > 0)) { victim = __ffs(cand) - 1; amount = 1; spend = l < 2; }
        }
    }
    // damage amount needs the victim's armor
    const int v_arm = g.bc(my_arm, victim >= 0 ? victim : l);
    int dealt = 0;
    if (victim >= 0) { const int full = max(0, my_dmg - v_arm); dealt = amount == 2 ? full / 2 : full; }
    // gather: what do I take, and which players damaged me
    int taken = 0, credit = 0;
    uint32_t attackers = g.ballot(victim >= 0 && dealt > 0);
    while (attackers) {                       // usually empty: most ticks have no combat
        const int j = __ffs(attackers) - 1;
        attackers &= attackers - 1;
        const int vj = g.bc(victim, j), aj = g.bc(dealt, j);
        if (vj == l) { taken += aj; if (j < 2) credit |= 1 << j; }
    }
    // 5. apply
    if (l < 2 && spend) L.aux -= min(L.aux, L.max_mana / 3);
    if (mover && L.alive) {
        L.hp -= taken;
        if (taken > 0) ++cnt.hits;
        if (moves) { L.x += dx; L.y += dy; }
    }
    // descents (player 0 then 1), then pickups
    bool descended = false;
#pragma unroll
    for (int p = 0; p < 2; ++p) {
        const int sx = p == 0 ? L.sx0 : L.sx1, sy = p == 0 ? L.sy0 : L.sy1;
        const bool want = g.bc((int)(l == p && moves && L.x == sx && L.y == sy), p) != 0;
        if (want) {
            const int nd = g.bc(L.depth, p) + 1;
            int nsx, nsy; uint32_t nkey;
            level_init(P, s, nd, nsx, nsy, nkey);
            const uint32_t t = free_tile(P, g, s, DOM_TICK, SUB_DESCEND + 64u * p, (uint32_t)L.tick, nd, nkey, nsx, nsy,
                                         mover && L.alive && l != p, L.x, L.y, L.depth);
            if (p == 0) { L.sx0 = nsx; L.sy0 = nsy; L.key0 = nkey; } else { L.sx1 = nsx; L.sy1 = nsy; L.key1 = nkey; }
            if (l == p) { L.depth = nd; L.x = t & 255; L.y = t >> 8; descended = true; ++cnt.descents; }
        }
    }
#pragma unroll
    for (int p = 0; p < 2; ++p) {
        const bool can = g.bc((int)(l == p && moves && !descended), p) != 0;
        if (!can) continue;
        const int px = g.bc(L.x, p), py = g.bc(L.y, p), pd = g.bc(L.depth, p);
        const bool here = l >= NM && l < NM + ORX_R1_ITEMS && L.alive && L.depth == pd && L.x == px && L.y == py;
        if (g.ballot(here) == 0) continue;
#pragma unroll
        for (int i = NM; i < NM + ORX_R1_ITEMS; ++i) {
            const bool hit = g.bc((int)here, i) != 0;
            const int kind = g.bc(L.aux, i);
            const int room = g.bc((int)(L.n_items < 4), p);
            if (hit && room) {
                if (l == p) {
                    if (kind == 0) L.damage += 1; else if (kind == 1) L.armor += 1; else { L.max_hp += 2; L.hp += 2; }
                    L.n_items += 1;
                }
                if (l == i) L.alive = 0;
            }
        }
    }
    // 6. enemy deaths in slot order: xp, level-ups, drops
    uint32_t dead = g.ballot(l >= 2 && mover && L.alive && L.hp <= 0);
    while (dead) {
        const int m = __ffs(dead) - 1;
        dead &= dead - 1;
        const int cr = g.bc(credit, m), mx = g.bc(L.x, m), my = g.bc(L.y, m), md = g.bc(L.depth, m);
        if (l == m) L.alive = 0;
        if (l < 2 && ((cr >> l) & 1)) {
            L.xp += 1;
            while (L.xp >= 3) { L.xp -= 3; L.level += 1; L.hp = L.max_hp; L.aux = L.max_mana; }
        }
        const uint4 b = draw_block(s, DOM_TICK, SUB_DROP + (uint32_t)((m - 2) >> 1), (uint32_t)L.tick);
        const uint32_t chance = ((m - 2) & 1) ? b.z : b.x, kind = (((m - 2) & 1) ? b.w : b.y) % 3u;
        const uint32_t freei = g.ballot(l >= NM && l < NM + ORX_R1_ITEMS && !L.alive);
        if (chance < (1u << 30) && freei != 0 && l == __ffs(freei) - 1) {
            L.alive = 1; L.depth = md; L.x = mx; L.y = my; L.aux = (int)kind; L.hp = 0;
        }
    }
    // 7. levels without a player vanish; spawns
    {
        const int d0 = g.bc(L.depth, 0), d1 = g.bc(L.depth, 1);
        if (l >= 2 && L.alive && L.depth != d0 && L.depth != d1) L.alive = 0;
        uint4 sb = make_uint4(0xFFFFFFFFu, 0xFFFFFFFFu, 0u, 0u);
        if ((L.tick & 3) == 0) sb = draw_block(s, DOM_TICK, SUB_SPAWN, (uint32_t)L.tick);   // spawn roll every 4th tick
#pragma unroll
        for (int p = 0; p < 2; ++p) {
            if (p == 1 && d1 == d0) continue;
            if ((p == 0 ? sb.x : sb.y) >= (1u << 30)) continue;
            const uint32_t freee = g.ballot(l >= 2 && mover && !L.alive);
            if (freee == 0) continue;
            const int slot = __ffs(freee) - 1;
            const int d = p == 0 ? d0 : d1;
            const uint32_t key = p == 0 ? L.key0 : L.key1;
            const int sx = p == 0 ? L.sx0 : L.sx1, sy = p == 0 ? L.sy0 : L.sy1;
            const uint32_t t = free_tile(P, g, s, DOM_TICK, SUB_SPAWN_TRY + 32u * p, (uint32_t)L.tick, d, key, sx, sy,
                                         mover && L.alive, L.x, L.y, L.depth);
            if (l == slot) { L.alive = 1; L.depth = d; L.x = t & 255; L.y = t >> 8; L.hp = min(20, 2 + d / 2); L.aux = 0; }
        }
    }
    // 8. mana, 9. separation, 10. cooldowns
    if (l < 2 && (L.tick & 3) == 0) L.aux = min(L.max_mana, L.aux + 1);
    {
        const int d0 = g.bc(L.depth, 0), d1 = g.bc(L.depth, 1);
        if (d0 != d1) {
            L.sep += 1;
            const int behind = d0 < d1 ? 0 : 1;
            if (l == behind) L.hp -= L.sep / 16;
        } else L.sep = 0;
    }
    if (l < 2) L.cd = cd_pre > 0 ? cd_pre - 1 : newcd;
    const int tick_pre = L.tick;
    L.tick += 1;
    const bool dead0 = g.bc(L.hp, 0) <= 0, dead1 = g.bc(L.hp, 1) <= 0;
    int res = ORX_RESULT_IN_PROGRESS;
    if (P.max_ticks != 0 && L.tick >= P.max_ticks) res = ORX_RESULT_TIE;
    if (dead1) res = ORX_RESULT_PLAYER1_WIN;
    if (dead0) res = ORX_RESULT_PLAYER2_WIN;
    if (dead0 && dead1) {
        const uint4 b = draw_block(s, DOM_TICK, SUB_MAIN, (uint32_t)tick_pre);
        res = (b.z >> 31) ? ORX_RESULT_PLAYER1_WIN : ORX_RESULT_PLAYER2_WIN;
    }
    return res;
}

__device__ __forceinline__ uint32_t ldg32(const void* p)
{
    uint32_t v;
    asm volatile("ld.global.u32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}

// All of a lane's loads are issued back to back (asm volatile keeps ptxas from sinking them below
// the first use), so one DRAM round trip covers the whole group state.
__device__ __forceinline__ void load_group(const R1Params& P, const Grp g, unsigned int game, R1Lane& L)
{
    const size_t e = (size_t)game * 16 + g.l;
    const size_t q = (size_t)game * 2 + (g.l & 1);          // lanes >= 2 read a valid word and drop it
    const uint32_t loc = ldg32(P.ent_loc + e), stat = ldg32(P.ent_stat + e), dep = ldg32(P.ent_depth + e);
    const uint32_t a = ldg32(P.pl_a + q), b = ldg32(P.pl_b + q), c = ldg32(P.pl_c + q);
    const uint32_t sw = ldg32(P.lvl_stairs + game), k0 = ldg32(P.lvl_key + (size_t)game * 2), k1 = ldg32(P.lvl_key + (size_t)game * 2 + 1);
    const uint32_t sep = ldg32(P.sep + game), tick = ldg32(P.tick + game), ep = ldg32(P.episode + game);
    L.x = loc & 255; L.y = (loc >> 8) & 255; L.alive = (loc >> 16) & 1; L.depth = (int)dep;
    L.hp = (int)(int16_t)(stat & 0xFFFF);
    L.aux = g.l >= NM ? (int)((loc >> 17) & 3) : (int)(int16_t)(stat >> 16);
    L.max_hp = L.max_mana = L.xp = L.level = L.n_items = L.cd = L.damage = L.armor = 0;
    if (g.l < 2) {
        L.max_hp = (int)(int16_t)(a & 0xFFFF); L.max_mana = (int)(int16_t)(a >> 16);
        L.xp = b & 255; L.level = (b >> 8) & 255; L.n_items = (b >> 16) & 255; L.cd = b >> 24;
        L.damage = c & 255; L.armor = (c >> 8) & 255;
    }
    L.sx0 = sw & 255; L.sy0 = (sw >> 8) & 255; L.sx1 = (sw >> 16) & 255; L.sy1 = sw >> 24;
    L.key0 = k0; L.key1 = k1;
    L.sep = (int)sep; L.tick = (int)tick; L.episode = ep;
}

__device__ __forceinline__ void store_group(const R1Params& P, const Grp g, unsigned int game, const R1Lane& L, int status)
{
    const size_t e = (size_t)game * 16 + g.l;
    uint32_t loc = (uint32_t)(L.x & 255) | ((uint32_t)(L.y & 255) << 8) | ((uint32_t)(L.alive & 1) << 16);
    uint32_t stat = (uint32_t)L.hp & 0xFFFFu;
    if (g.l >= NM) loc |= (uint32_t)(L.aux & 3) << 17; else stat |= (uint32_t)L.aux << 16;
    const bool blank = (!L.alive && g.l >= 2) || g.l >= NM + ORX_R1_ITEMS;
    P.ent_loc[e] = blank ? 0u : loc;
    P.ent_stat[e] = blank ? 0u : stat;
    P.ent_depth[e] = blank ? 0 : L.depth;
    if (g.l < 2) {
        const size_t q = (size_t)game * 2 + g.l;
        P.pl_a[q] = ((uint32_t)L.max_hp & 0xFFFFu) | ((uint32_t)L.max_mana << 16);
        P.pl_b[q] = (uint32_t)(L.xp & 255) | ((uint32_t)(L.level & 255) << 8) | ((uint32_t)(L.n_items & 255) << 16) | ((uint32_t)L.cd << 24);
        P.pl_c[q] = (uint32_t)(L.damage & 255) | ((uint32_t)(L.armor & 255) << 8);
        P.lvl_key[q] = g.l == 0 ? L.key0 : L.key1;
    }
    if (g.l == 0) {
        P.lvl_stairs[game] = (uint32_t)L.sx0 | ((uint32_t)L.sy0 << 8) | ((uint32_t)L.sx1 << 16) | ((uint32_t)L.sy1 << 24);
        P.sep[game] = (uint32_t)L.sep; P.tick[game] = L.tick; P.episode[game] = L.episode; P.status[game] = (uint8_t)status;
    }
}

__device__ __forceinline__ Grp make_group()
{
    Grp g;
    g.base = threadIdx.x & 16;
    g.mask = 0xFFFFu << g.base;
    g.l = threadIdx.x & 15;
    return g;
}
__device__ __forceinline__ Stream make_stream(const R1Params& P, unsigned int game, uint32_t episode)
{
    const unsigned long long gid = P.gid_base + game;
    Stream s;
    s.rk = &P.rk; s.g0 = (uint32_t)gid; s.g1 = (uint32_t)(gid >> 32); s.episode = episode;
    return s;
}

__global__ void __launch_bounds__(kThreadsR1)
k_r1_reset(const __grid_constant__ R1Params P, const uint8_t* __restrict__ mask, int bump)
{
    const Grp g = make_group();
    const unsigned int game = (blockIdx.x * kThreadsR1 + threadIdx.x) >> 4;
    if (game >= P.n) return;
    if (mask != nullptr && mask[game] == 0) return;
    R1Lane L;
    memset(&L, 0, sizeof(L));
    L.episode = P.episode[game] + (bump ? 1u : 0u);
    const Stream s = make_stream(P, game, L.episode);
    setup_game(P, g, L, s);
    store_group(P, g, game, L, ORX_RESULT_IN_PROGRESS);
}

#ifndef ORX_R1_MINBLOCKS
#define ORX_R1_MINBLOCKS 3
#endif
__global__ void __launch_bounds__(kThreadsR1, ORX_R1_MINBLOCKS)
k_r1_step(const __grid_constant__ R1Params P, const uint8_t* __restrict__ moves, uint8_t* __restrict__ result)
{
    const Grp g = make_group();
    const unsigned int game = (blockIdx.x * kThreadsR1 + threadIdx.x) >> 4;
    if (game >= P.n) return;
    R1Lane L;
    load_group(P, g, game, L);
    uint32_t mvw, status;
    asm volatile("ld.global.u16 %0, [%1];" : "=r"(mvw) : "l"(moves + 2 * (size_t)game));
    asm volatile("ld.global.u8 %0, [%1];" : "=r"(status) : "l"(P.status + game));
    if (status != ORX_RESULT_IN_PROGRESS) { if (g.l == 0) result[game] = (uint8_t)status; return; }
    Stream s = make_stream(P, game, L.episode);
    R1Counters cnt{};
    int res = r1_tick(P, g, L, s, (int)(mvw & 255u), (int)(mvw >> 8), cnt);
    if (g.l == 0) result[game] = (uint8_t)res;
    if (res != ORX_RESULT_IN_PROGRESS && P.auto_reset) {
        L.episode += 1;
        s.episode = L.episode;
        setup_game(P, g, L, s);
        res = ORX_RESULT_IN_PROGRESS;
    }
    store_group(P, g, game, L, res);
}

__device__ __forceinline__ unsigned int warp_sum(unsigned int v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__global__ void __launch_bounds__(kThreadsR1, ORX_R1_MINBLOCKS)
k_r1_rollout(const __grid_constant__ R1Params P, int n_ticks, unsigned long long* __restrict__ stats)
{
    __shared__ unsigned int s_cnt[ORX_STAT_COUNT];
    if (threadIdx.x < ORX_STAT_COUNT) s_cnt[threadIdx.x] = 0;
    __syncthreads();
    const Grp g = make_group();
    const unsigned int game = (blockIdx.x * kThreadsR1 + threadIdx.x) >> 4;
    R1Counters cnt{};
    if (game < P.n) {
        int status = P.status[game];
        if (status == ORX_RESULT_IN_PROGRESS) {
            R1Lane L;
            load_group(P, g, game, L);
            Stream s = make_stream(P, game, L.episode);
            for (int t = 0; t < n_ticks; ++t) {
                const uint4 b = draw_block(s, DOM_TICK, SUB_MAIN, (uint32_t)L.tick);
                const int res = r1_tick(P, g, L, s, 1 + (int)bounded(b.x, 6u), 1 + (int)bounded(b.y, 6u), cnt);
                if (g.l == 0) {
                    ++cnt.ticks;
                    cnt.p1 += res == ORX_RESULT_PLAYER1_WIN; cnt.p2 += res == ORX_RESULT_PLAYER2_WIN; cnt.ties += res == ORX_RESULT_TIE;
                }
                if (res != ORX_RESULT_IN_PROGRESS) {
                    if (P.auto_reset) { L.episode += 1; s.episode = L.episode; setup_game(P, g, L, s); }
                    else { status = res; break; }
                }
            }
            store_group(P, g, game, L, status);
        }
    }
    if (stats != nullptr) {
        unsigned int v[ORX_STAT_COUNT] = {cnt.ticks, cnt.p1, cnt.p2, cnt.ties, 0u, cnt.descents, cnt.hits, 0u};
#pragma unroll
        for (int k = 0; k < ORX_STAT_COUNT; ++k) {
            const unsigned int w = warp_sum(v[k]);
            if ((threadIdx.x & 31) == 0 && w != 0) atomicAdd(&s_cnt[k], w);
        }
        __syncthreads();
        if (threadIdx.x < ORX_STAT_COUNT && s_cnt[threadIdx.x] != 0)
            atomicAdd(&stats[threadIdx.x], (unsigned long long)s_cnt[threadIdx.x]);
    }
}

}  // namespace

namespace {
#include "orx_r1t.cuh"
}  // namespace

namespace {

// Observation: lanes 0/1 write the scalar block of their player, every entity lane writes its own
// triple into both players' rows, and the 7x7 wall window is hashed 16 tiles at a time by the
// whole group -- one ballot is one output word.
__global__ void __launch_bounds__(kThreadsR1)
k_r1_observe(const __grid_constant__ R1Params P, int16_t* __restrict__ obs, int radius)
{
    const Grp g = make_group();
    const unsigned int game = (blockIdx.x * kThreadsR1 + threadIdx.x) >> 4;
    if (game >= P.n) return;
    R1Lane L;
    load_group(P, g, game, L);
    const int status = P.status[game];
    const int l = g.l;
#pragma unroll
    for (int p = 0; p < 2; ++p) {
        int16_t* o = obs + ((size_t)game * 2 + p) * ORX_R1_OBS_LEN;
        const int mx = g.bc(L.x, p), my = g.bc(L.y, p), md = g.bc(L.depth, p);
        const int ox = g.bc(L.x, 1 - p), oy = g.bc(L.y, 1 - p), od = g.bc(L.depth, 1 - p), oh = g.bc(L.hp, 1 - p);
        const int sx = p == 0 ? L.sx0 : L.sx1, sy = p == 0 ? L.sy0 : L.sy1;
        const uint32_t key = p == 0 ? L.key0 : L.key1;
        if (l == p) {
            o[0] = (int16_t)L.x; o[1] = (int16_t)L.y; o[2] = (int16_t)min(L.depth, 32767); o[3] = (int16_t)L.hp;
            o[4] = (int16_t)L.aux; o[5] = (int16_t)L.cd; o[6] = (int16_t)L.damage; o[7] = (int16_t)L.armor;
            o[8] = (int16_t)L.max_hp; o[9] = (int16_t)L.max_mana; o[10] = (int16_t)L.level; o[11] = (int16_t)L.xp;
            o[12] = (int16_t)L.n_items; o[13] = (int16_t)min(L.sep, 32767); o[14] = (int16_t)min(L.tick, 32767); o[15] = (int16_t)status;
            const bool same = od == md;
            o[16] = (int16_t)same; o[17] = (int16_t)(same ? ox : -1); o[18] = (int16_t)(same ? oy : -1); o[19] = (int16_t)(same ? oh : 0);
            const bool vis = radius < 0 || max(abs(sx - mx), abs(sy - my)) <= radius;
            o[20] = (int16_t)vis; o[21] = (int16_t)(vis ? sx : -1); o[22] = (int16_t)(vis ? sy : -1);
            o[63] = 0;
        }
        const bool here = L.alive && L.depth == md;
        if (l >= 2 && l < NM) {
            int16_t* e = o + 23 + 3 * (l - 2);
            e[0] = (int16_t)(here ? L.x : -1); e[1] = (int16_t)(here ? L.y : -1); e[2] = (int16_t)(here ? L.hp : 0);
        } else if (l >= NM && l < NM + ORX_R1_ITEMS) {
            int16_t* e = o + 47 + 3 * (l - NM);
            e[0] = (int16_t)(here ? L.x : -1); e[1] = (int16_t)(here ? L.y : -1); e[2] = (int16_t)(here ? L.aux : -1);
        }
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const int t = 16 * r + l;
            const bool w = t < 49 && is_wall(P, key, sx, sy, mx + t % 7 - 3, my + t / 7 - 3);
            const uint32_t bits = g.ballot(w);
            if (l == 0) o[59 + r] = (int16_t)bits;
        }
    }
}

int r1_check(const OrxR1Config* c, const OrxR1State* st, int64_t n)
{
    if (c == nullptr || st == nullptr || n < 0) return ORX_ERR_BAD_ARG;
    if (c->struct_size != sizeof(OrxR1Config)) return ORX_ERR_BAD_ARG;
    if (n > (1ll << 27)) return ORX_ERR_UNSUPPORTED;
    if (c->width < 5 || c->height < 5 || c->width > ORX_MAX_DIM || c->height > ORX_MAX_DIM) return ORX_ERR_BAD_ARG;
    if (c->max_ticks < 0 || c->wall_density < 0 || c->wall_density > 128) return ORX_ERR_BAD_ARG;
    if (!st->ent_loc || !st->ent_depth || !st->ent_stat || !st->pl_a || !st->pl_b || !st->pl_c || !st->lvl_stairs ||
        !st->lvl_key || !st->sep || !st->tick || !st->episode || !st->status) return ORX_ERR_BAD_ARG;
    return ORX_OK;
}

R1Params r1_params(const OrxR1Config* c, const OrxR1State* st, int64_t n, uint64_t base)
{
    R1Params P;
    memset(&P, 0, sizeof(P));
    P.W = c->width; P.H = c->height; P.max_ticks = c->max_ticks; P.auto_reset = c->auto_reset; P.wall_density = c->wall_density;
    make_round_keys(P.rk, (uint32_t)c->seed, (uint32_t)(c->seed >> 32));
    P.ent_loc = st->ent_loc; P.ent_stat = st->ent_stat; P.ent_depth = st->ent_depth;
    P.pl_a = st->pl_a; P.pl_b = st->pl_b; P.pl_c = st->pl_c;
    P.lvl_stairs = st->lvl_stairs; P.lvl_key = st->lvl_key; P.sep = st->sep;
    P.tick = st->tick; P.episode = st->episode; P.status = st->status;
    P.n = (unsigned int)n; P.gid_base = base;
    return P;
}

int r1_done()
{
    const cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? ORX_OK : ORX_ERR_CUDA_BASE - (int)e;
}

int r1_grid(int64_t n) { return (int)((n * 16 + kThreadsR1 - 1) / kThreadsR1); }

}  // namespace

extern "C" {

int orx_r1_reset(const OrxR1Config* cfg, const OrxR1State* st, const uint8_t* mask, int bump_episode,
                 int64_t n, uint64_t game_id_base, void* cuda_stream)
{
    const int rc = r1_check(cfg, st, n);
    if (rc != ORX_OK) return rc;
    if ((game_id_base >> 54) != 0 || ((game_id_base + (uint64_t)n) >> 54) != 0) return ORX_ERR_BAD_ARG;
    if (n == 0) return ORX_OK;
    if (st->sched != nullptr && st->sched_words > 0) {      // the hand-over words are balanced between launches; a reset re-establishes zero
        const cudaError_t e = cudaMemsetAsync(st->sched, 0, (size_t)st->sched_words * sizeof(unsigned int), static_cast<cudaStream_t>(cuda_stream));
        if (e != cudaSuccess) return ORX_ERR_CUDA_BASE - (int)e;
    }
    k_r1_reset<<<r1_grid(n), kThreadsR1, 0, static_cast<cudaStream_t>(cuda_stream)>>>(r1_params(cfg, st, n, game_id_base), mask, bump_episode);
    return r1_done();
}

// Common body of orx_r1_step / orx_r1_step_events / orx_r1_step_host_sync.
static int r1_step_impl(const OrxR1Config* cfg, const OrxR1State* st, const uint8_t* moves, uint8_t* result, OrxEvent* events,
                        int max_events, int64_t n, uint64_t game_id_base, void* cuda_stream)
{
    const int rc = r1_check(cfg, st, n);
    if (rc != ORX_OK) return rc;
    if ((game_id_base >> 54) != 0 || ((game_id_base + (uint64_t)n) >> 54) != 0) return ORX_ERR_BAD_ARG;
    if (moves == nullptr || result == nullptr || (reinterpret_cast<uintptr_t>(moves) & 1)) return ORX_ERR_BAD_ARG;
    if (events != nullptr && (max_events < 1 || max_events > 255 || (reinterpret_cast<uintptr_t>(events) & 7))) return ORX_ERR_BAD_ARG;
    if (n == 0) return ORX_OK;
    // Default: one thread per game (orx_r1t.cuh). ORX_R1_PATH_HALFWARP selects the sixteen-lanes-per-game
    // kernels instead (same results; kept as the warp-primitive formulation and as a cross-check).
    cudaStream_t s = static_cast<cudaStream_t>(cuda_stream);
    if (cfg->path_flags & ORX_R1_PATH_HALFWARP) {
        if (events != nullptr) return ORX_ERR_UNSUPPORTED;
        k_r1_step<<<r1_grid(n), kThreadsR1, 0, s>>>(r1_params(cfg, st, n, game_id_base), moves, result);
        return r1_done();
    }
    static_assert(r1t::kThreads == ORX_R1_BLOCK, "a hand-over block is the games of one CTA");
    const unsigned int grid = (unsigned)((n + r1t::kThreads - 1) / r1t::kThreads);
    const R1Params P = r1_params(cfg, st, n, game_id_base);
    const uint16_t* mv = reinterpret_cast<const uint16_t*>(moves);
    uint2* evp = reinterpret_cast<uint2*>(events);
    const bool flagged = st->sched != nullptr && (reinterpret_cast<uintptr_t>(st->sched) & 3) == 0 &&
                         st->sched_words >= ORX_R1_SCHED_WORDS(n) && (cfg->path_flags & ORX_R1_PATH_BLOCK_FLAGS) != 0;
    if (!flagged) {
        if (events != nullptr) r1t::k_step<false, true><<<grid, r1t::kThreads, 0, s>>>(P, mv, result, nullptr, evp, max_events);
        else r1t::k_step<false, false><<<grid, r1t::kThreads, 0, s>>>(P, mv, result, nullptr, nullptr, 0);
        return r1_done();
    }
    cudaLaunchConfig_t lc = {};
    lc.gridDim = dim3(grid); lc.blockDim = dim3(r1t::kThreads); lc.dynamicSmemBytes = 0; lc.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    lc.attrs = at; lc.numAttrs = 1;
    const cudaError_t e = events != nullptr ? cudaLaunchKernelEx(&lc, r1t::k_step<true, true>, P, mv, result, st->sched, evp, max_events)
                                            : cudaLaunchKernelEx(&lc, r1t::k_step<true, false>, P, mv, result, st->sched, (uint2*)nullptr, 0);
    return e == cudaSuccess ? r1_done() : ORX_ERR_CUDA_BASE - (int)e;
}

int orx_r1_step(const OrxR1Config* cfg, const OrxR1State* st, const uint8_t* moves, uint8_t* result,
                int64_t n, uint64_t game_id_base, void* cuda_stream)
{
    return r1_step_impl(cfg, st, moves, result, nullptr, 0, n, game_id_base, cuda_stream);
}

int orx_r1_step_events(const OrxR1Config* cfg, const OrxR1State* st, const uint8_t* moves, uint8_t* result,
                       OrxEvent* events, int max_events, int64_t n, uint64_t game_id_base, void* cuda_stream)
{
    if (events == nullptr) return ORX_ERR_BAD_ARG;
    return r1_step_impl(cfg, st, moves, result, events, max_events, n, game_id_base, cuda_stream);
}

int orx_r1_step_host_sync(const OrxR1Config* cfg, const OrxR1State* st, const uint8_t* host_moves, uint8_t* host_result,
                          int64_t n, uint64_t game_id_base, void* cuda_stream)
{
    // The buffers must be device-accessible host memory (pinned): the kernel reads / writes them across PCIe itself.
    for (const void* p : {static_cast<const void*>(host_moves), static_cast<const void*>(host_result)}) {
        if (p == nullptr) return ORX_ERR_BAD_ARG;
        cudaPointerAttributes a;
        if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return ORX_ERR_BAD_ARG; }
        if (a.type != cudaMemoryTypeHost && a.type != cudaMemoryTypeManaged) return ORX_ERR_BAD_ARG;
    }
    const int rc = r1_step_impl(cfg, st, host_moves, host_result, nullptr, 0, n, game_id_base, cuda_stream);
    if (rc != ORX_OK) return rc;
    const cudaError_t e = cudaStreamSynchronize(static_cast<cudaStream_t>(cuda_stream));
    return e == cudaSuccess ? ORX_OK : ORX_ERR_CUDA_BASE - (int)e;
}

int orx_r1_bot_moves(const OrxR1Config* cfg, const OrxR1State* st, int bot1, int bot2, uint8_t* moves,
                     int64_t n, uint64_t game_id_base, void* cuda_stream)
{
    const int rc = r1_check(cfg, st, n);
    if (rc != ORX_OK) return rc;
    if ((game_id_base >> 54) != 0 || ((game_id_base + (uint64_t)n) >> 54) != 0) return ORX_ERR_BAD_ARG;
    if (moves == nullptr || bot1 < ORX_BOT_NONE || bot1 > ORX_BOT_STAIRCASE || bot2 < ORX_BOT_NONE || bot2 > ORX_BOT_STAIRCASE) return ORX_ERR_BAD_ARG;
    if (n == 0 || (bot1 == ORX_BOT_NONE && bot2 == ORX_BOT_NONE)) return ORX_OK;
    r1t::k_bot_moves<<<(unsigned)((n + 255) / 256), 256, 0, static_cast<cudaStream_t>(cuda_stream)>>>(r1_params(cfg, st, n, game_id_base), bot1, bot2, moves);
    return r1_done();
}

int orx_r1_replay(const OrxR1Config* cfg, const OrxR1State* st, const uint8_t* moves, uint8_t* results, int n_ticks,
                  int64_t n, uint64_t game_id_base, void* cuda_stream)
{
    const int rc = r1_check(cfg, st, n);
    if (rc != ORX_OK) return rc;
    if ((game_id_base >> 54) != 0 || ((game_id_base + (uint64_t)n) >> 54) != 0) return ORX_ERR_BAD_ARG;
    if (n_ticks < 0 || moves == nullptr || results == nullptr || (reinterpret_cast<uintptr_t>(moves) & 1)) return ORX_ERR_BAD_ARG;
    if (n == 0 || n_ticks == 0) return ORX_OK;
    r1t::k_replay<<<(unsigned)((n + r1t::kThreads - 1) / r1t::kThreads), r1t::kThreads, 0, static_cast<cudaStream_t>(cuda_stream)>>>(
        r1_params(cfg, st, n, game_id_base), reinterpret_cast<const uint16_t*>(moves), results, n_ticks);
    return r1_done();
}

int orx_r1_rollout(const OrxR1Config* cfg, const OrxR1State* st, int n_ticks, unsigned long long* stats,
                   int64_t n, uint64_t game_id_base, void* cuda_stream)
{
    const int rc = r1_check(cfg, st, n);
    if (rc != ORX_OK) return rc;
    if ((game_id_base >> 54) != 0 || ((game_id_base + (uint64_t)n) >> 54) != 0) return ORX_ERR_BAD_ARG;
    if (n_ticks < 0) return ORX_ERR_BAD_ARG;
    if (n == 0 || n_ticks == 0) return ORX_OK;
    if (cfg->path_flags & ORX_R1_PATH_HALFWARP)
        k_r1_rollout<<<r1_grid(n), kThreadsR1, 0, static_cast<cudaStream_t>(cuda_stream)>>>(r1_params(cfg, st, n, game_id_base), n_ticks, stats);
    else
        r1t::k_rollout<<<(unsigned)((n + r1t::kThreads - 1) / r1t::kThreads), r1t::kThreads, 0, static_cast<cudaStream_t>(cuda_stream)>>>(
            r1_params(cfg, st, n, game_id_base), n_ticks, stats);
    return r1_done();
}

int orx_r1_observe(const OrxR1Config* cfg, const OrxR1State* st, int16_t* obs, int stairs_radius,
                   int64_t n, void* cuda_stream)
{
    const int rc = r1_check(cfg, st, n);
    if (rc != ORX_OK) return rc;
    if (obs == nullptr) return ORX_ERR_BAD_ARG;
    if (n == 0) return ORX_OK;
    k_r1_observe<<<r1_grid(n), kThreadsR1, 0, static_cast<cudaStream_t>(cuda_stream)>>>(r1_params(cfg, st, n, 0), obs, stairs_radius);
    return r1_done();
}

}  // extern "C"
