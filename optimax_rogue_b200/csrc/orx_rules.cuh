// Data-parallel formulation of one Optimax Rogue tick: one thread owns one game, the whole
// game state lives in registers, and the two movers are resolved in initiative order by
// swapping roles with selects (no divergent "who goes first" branch, no local-memory arrays).
//
// Reference behaviour restated here (paths under the reference root):
//   Updater.update            optimax_rogue/logic/updater.py:76-162
//   Updater.handle_move       optimax_rogue/logic/updater.py:180-243
//   Updater.handle_descend    optimax_rogue/logic/updater.py:259-296
//   Updater.handle_combat     optimax_rogue/logic/updater.py:298-338
//   Dungeon.is_blocked        optimax_rogue/game/world.py:41-46
//   Dungeon.get_random_unblocked  optimax_rogue/game/world.py:57-66
//   EmptyDungeonGenerator     optimax_rogue/logic/worldgen.py:33-43
//   setup_game                optimax_rogue/logic/worldgen.py:77-87, 124-135
#pragma once
#include <stdint.h>

#include "../../include/orx.h"
#include "orx_rng.cuh"

namespace orx {

// Kernel parameter block: config scalars + state plane pointers, passed __grid_constant__.
struct Params {
    int W, H;
    int start_kind, sd0, sd1;
    int despawn, max_ticks;
    int hp0, hp1;
    int dmg0, dmg1;            // damage - armor of the ATTACKER (updater.py:313)
    int auto_reset, n_npc;
    uint32_t k0, k1;
    const uint8_t* tiles;      // DGEN_FIXED: uint8[W*H] x-major
    const uint16_t* ground;    // DGEN_FIXED: Ground tile list
    int n_ground, fsx, fsy;
    uint32_t* pos;             // x1 | y1<<8 | x2<<16 | y2<<24
    uint32_t* hp;              // int16 hp1 | int16 hp2 << 16
    int2* depth;
    uint32_t* stairs;          // sx1 | sy1<<8 | sx2<<16 | sy2<<24
    int* tick;
    uint32_t* episode;
    uint8_t* status;
    uint8_t* npc_pos;
    int16_t* npc_hp;
    int* npc_depth;
    long long n;
    unsigned long long gid_base;
};

struct Counters {               // per-thread, reduced at the end of orx_rollout
    unsigned int ticks, p1, p2, ties, events, descents, hits;
};

struct Mover { int x, y, hp, depth, sx, sy, mv, id, dmg; };

template <bool EV>
struct EvSink {
    uint2* base;   // this game's slots
    int n, cap;
    __device__ __forceinline__ void emit(int kind, int iden, int a, int b, int depth)
    {
        if (EV) {
            if (n < cap) base[n] = make_uint2((uint32_t)kind | ((uint32_t)iden << 8) | ((uint32_t)(a & 255) << 16) | ((uint32_t)(b & 255) << 24), (uint32_t)depth);
        }
        ++n;
    }
    __device__ __forceinline__ void finish()
    {
        if (EV) for (int k = n; k < cap; ++k) base[k] = make_uint2(0u, 0u);
    }
};

// ---------------------------------------------------------------- level model
template <int DGEN>
__device__ __forceinline__ bool is_blocked(const Params& P, const uint8_t* tiles, int x, int y)
{
    if (DGEN == ORX_DGEN_EMPTY)   // border walls; out of bounds is covered by the same compare
        return (x <= 0) | (y <= 0) | (x >= P.W - 1) | (y >= P.H - 1);
    if ((unsigned)x >= (unsigned)P.W || (unsigned)y >= (unsigned)P.H) return true;
    return tiles[x * P.H + y] == ORX_TILE_WALL;
}

template <int DGEN>
__device__ __forceinline__ bool is_stairs(const Params& P, const uint8_t* tiles, int x, int y, int sx, int sy)
{
    if (DGEN == ORX_DGEN_EMPTY) return (x == sx) & (y == sy);
    return tiles[x * P.H + y] == ORX_TILE_STAIRCASE_DOWN;
}

// worldgen.py:39-40 (numpy randint is high-exclusive: W-3 / H-3 values starting at 1)
template <int DGEN>
__device__ __forceinline__ void level_stairs(const Params& P, const Stream& s, int depth, int& sx, int& sy)
{
    if (DGEN == ORX_DGEN_EMPTY) {
        const uint4 b = draw_block(s, DOM_LEVEL, 0, (uint32_t)depth);
        sx = 1 + (int)bounded(b.x, (uint32_t)(P.W - 3));
        sy = 1 + (int)bounded(b.y, (uint32_t)(P.H - 3));
    } else {
        sx = P.fsx; sy = P.fsy;
    }
}

template <int DGEN>
__device__ __forceinline__ int n_ground(const Params& P)
{
    return DGEN == ORX_DGEN_EMPTY ? (P.W - 2) * (P.H - 2) - 1 : P.n_ground;
}

// world.py:59-66: k-th Ground tile in x-major order. Empty room: closed form (interior
// index skipping the staircase); fixed map: precomputed rank table.
template <int DGEN>
__device__ __forceinline__ void kth_ground(const Params& P, int sx, int sy, int k, int& x, int& y)
{
    if (DGEN == ORX_DGEN_EMPTY) {
        const int hm = P.H - 2;
        const int s = (sx - 1) * hm + (sy - 1);
        const int i = k + (k >= s);
        x = 1 + i / hm;
        y = 1 + i - (x - 1) * hm;
    } else {
        const int flat = P.ground[k];
        x = flat / P.H;
        y = flat - x * P.H;
    }
}

__device__ __forceinline__ int npc_at(const Params& P, long long lane, int depth, int x, int y)
{
    for (int k = 0; k < P.n_npc; ++k) {
        const long long j = lane * P.n_npc + k;
        if (P.npc_depth[j] == depth && P.npc_pos[2 * j] == x && P.npc_pos[2 * j + 1] == y) return k;
    }
    return -1;
}

// Is level `depth` present in World.dungeons when player `pid` descends into it? Only feeds the
// DungeonCreated event (updater.py:274-280): levels are re-derived from Philox, never stored.
__device__ __forceinline__ bool level_exists(const Params& P, int pid, int depth, int other_depth)
{
    const int o_start = P.start_kind == ORX_START_SEPARATED ? (pid == 0 ? P.sd1 : P.sd0) : P.sd0;
    if (P.despawn == ORX_DESPAWN_UNUSED) return other_depth == depth;
    return (o_start <= depth) & (depth <= other_depth);
}

// ---------------------------------------------------------------- rare paths (kept out of line)
// handle_descend draws: stairs of the new level, then spawn tries until the tile is free
// (updater.py:282-285). Returns sx | sy<<8 | x<<16 | y<<24.
template <int DGEN, bool NPC>
__device__ __noinline__ uint32_t descend_draw(const Params& P, Stream s, int tick, int pid, int new_depth,
                                              int ox, int oy, int odepth, long long lane)
{
    int sx, sy, x, y;
    level_stairs<DGEN>(P, s, new_depth, sx, sy);
    const int ng = n_ground<DGEN>(P);
    int q = 0;
    bool taken;
    do {
        const int k = (int)seq_bounded(s, DOM_TICK, SUB_DESCEND + 64u * (uint32_t)pid, (uint32_t)tick, q++, (uint32_t)ng);
        kth_ground<DGEN>(P, sx, sy, k, x, y);
        taken = (odepth == new_depth) & (ox == x) & (oy == y);
        if (NPC) taken = taken || npc_at(P, lane, new_depth, x, y) >= 0;
    } while (taken);
    return (uint32_t)sx | ((uint32_t)sy << 8) | ((uint32_t)x << 16) | ((uint32_t)y << 24);
}

// setup_game draws (worldgen.py:77-87 / :124-135). Returns (pos word, stairs word).
template <int DGEN>
__device__ __noinline__ uint2 reset_draw(const Params& P, Stream s)
{
    const int ng = n_ground<DGEN>(P);
    int sx1, sy1, sx2, sy2, x1, y1, x2, y2;
    if (P.start_kind == ORX_START_TOGETHER) {
        level_stairs<DGEN>(P, s, P.sd0, sx1, sy1);
        sx2 = sx1; sy2 = sy1;
        int q = 0;
        kth_ground<DGEN>(P, sx1, sy1, (int)seq_bounded(s, DOM_RESET, 0, 0, q++, (uint32_t)ng), x1, y1);
        do {
            kth_ground<DGEN>(P, sx1, sy1, (int)seq_bounded(s, DOM_RESET, 0, 0, q++, (uint32_t)ng), x2, y2);
        } while ((x2 == x1) & (y2 == y1));
    } else {
        level_stairs<DGEN>(P, s, P.sd0, sx1, sy1);
        level_stairs<DGEN>(P, s, P.sd1, sx2, sy2);
        kth_ground<DGEN>(P, sx1, sy1, (int)seq_bounded(s, DOM_RESET, 0, 0, 0, (uint32_t)ng), x1, y1);
        kth_ground<DGEN>(P, sx2, sy2, (int)seq_bounded(s, DOM_RESET, 0, 0, 1, (uint32_t)ng), x2, y2);
    }
    return make_uint2((uint32_t)x1 | ((uint32_t)y1 << 8) | ((uint32_t)x2 << 16) | ((uint32_t)y2 << 24),
                      (uint32_t)sx1 | ((uint32_t)sy1 << 8) | ((uint32_t)sx2 << 16) | ((uint32_t)sy2 << 24));
}

// ---------------------------------------------------------------- one mover (updater.py:180-243)
template <int DGEN, bool NPC, bool EV, int IND>
__device__ __forceinline__ void do_move(const Params& P, const uint8_t* tiles, Mover& me, Mover& ot,
                                        const Stream& s, int tick, long long lane, EvSink<EV>& ev, Counters& cnt)
{
    if (me.mv == ORX_MOVE_STAY) return;
    const int nx = me.x + (me.mv == ORX_MOVE_RIGHT) - (me.mv == ORX_MOVE_LEFT);
    const int ny = me.y + (me.mv == ORX_MOVE_DOWN) - (me.mv == ORX_MOVE_UP);
    if ((ot.depth == me.depth) & (ot.x == nx) & (ot.y == ny)) {
        // Block if the occupant's (clamped) move is Stay; otherwise Ambush when the occupant acted
        // earlier (it just arrived) or Flee when it acts later. Parry (updater.py:229-234) needs
        // occupant.pos + delta == occupant.pos with a non-Stay move: unreachable.
        const int flag = ot.mv == ORX_MOVE_STAY ? ORX_FLAG_BLOCK : (IND == 1 ? ORX_FLAG_AMBUSH : ORX_FLAG_FLEE);
        if (me.dmg > 0) { ot.hp -= me.dmg; ++cnt.hits; }
        ev.emit(ORX_EV_COMBAT, me.id + 1, ot.id + 1, flag, me.dmg);
        return;                                   // the attacker never advances (updater.py:222-243)
    }
    if (NPC) {
        const int k = npc_at(P, lane, me.depth, nx, ny);
        if (k >= 0) {                             // NPC moves are always Stay (updater.py:165-178)
            if (me.dmg > 0) { P.npc_hp[lane * P.n_npc + k] -= (int16_t)me.dmg; ++cnt.hits; }
            ev.emit(ORX_EV_COMBAT, me.id + 1, 3 + k, ORX_FLAG_BLOCK, me.dmg);
            return;
        }
    }
    if (is_stairs<DGEN>(P, tiles, nx, ny, me.sx, me.sy)) {
        const int nd = me.depth + 1;
        const uint32_t r = descend_draw<DGEN, NPC>(P, s, tick, me.id, nd, ot.x, ot.y, ot.depth, lane);
        const int nsx = r & 255, nsy = (r >> 8) & 255, sxp = (r >> 16) & 255, syp = r >> 24;
        if (!level_exists(P, me.id, nd, ot.depth)) ev.emit(ORX_EV_DUNGEON, 0, nsx, nsy, nd);
        ev.emit(ORX_EV_DESCEND, me.id + 1, sxp, syp, nd);
        me.depth = nd; me.x = sxp; me.y = syp; me.sx = nsx; me.sy = nsy;
        ++cnt.descents;
        return;
    }
    ev.emit(ORX_EV_MOVE, me.id + 1, nx, ny, me.depth);
    me.x = nx; me.y = ny;
}

// Game state of one lane, unpacked into registers.
struct Lane {
    Mover p1, p2;       // .mv/.id/.dmg filled by tick_lane
    int tick;
    uint32_t episode;
};

__device__ __forceinline__ Mover pick(bool c, const Mover& a, const Mover& b)
{
    Mover r;
    r.x = c ? a.x : b.x; r.y = c ? a.y : b.y; r.hp = c ? a.hp : b.hp; r.depth = c ? a.depth : b.depth;
    r.sx = c ? a.sx : b.sx; r.sy = c ? a.sy : b.sy; r.mv = c ? a.mv : b.mv; r.id = c ? a.id : b.id;
    r.dmg = c ? a.dmg : b.dmg;
    return r;
}

// One Updater.update. w_init is word 2 of the tick's main block. Returns the UpdateResult.
template <int DGEN, bool NPC, bool EV>
__device__ __forceinline__ int tick_lane(const Params& P, const uint8_t* tiles, Lane& L, int m1, int m2,
                                         uint32_t w_init, const Stream& s, long long lane,
                                         EvSink<EV>& ev, Counters& cnt)
{
    // commands outside Move (logic/moves.py:6-12) are Stay
    m1 = (m1 >= ORX_MOVE_UP && m1 <= ORX_MOVE_LEFT) ? m1 : ORX_MOVE_STAY;
    m2 = (m2 >= ORX_MOVE_UP && m2 <= ORX_MOVE_LEFT) ? m2 : ORX_MOVE_STAY;
    // wall clamp from pre-tick positions, before the shuffle (updater.py:90-98)
    {
        const int nx = L.p1.x + (m1 == ORX_MOVE_RIGHT) - (m1 == ORX_MOVE_LEFT);
        const int ny = L.p1.y + (m1 == ORX_MOVE_DOWN) - (m1 == ORX_MOVE_UP);
        if (is_blocked<DGEN>(P, tiles, nx, ny)) m1 = ORX_MOVE_STAY;
    }
    {
        const int nx = L.p2.x + (m2 == ORX_MOVE_RIGHT) - (m2 == ORX_MOVE_LEFT);
        const int ny = L.p2.y + (m2 == ORX_MOVE_DOWN) - (m2 == ORX_MOVE_UP);
        if (is_blocked<DGEN>(P, tiles, nx, ny)) m2 = ORX_MOVE_STAY;
    }
    L.p1.mv = m1; L.p1.id = 0; L.p1.dmg = P.dmg0;
    L.p2.mv = m2; L.p2.id = 1; L.p2.dmg = P.dmg1;
    // random.shuffle([p1, p2]) (updater.py:114): j = randbelow(2); j == 0 swaps => p2 first
    const bool p2_first = bounded(w_init, 2u) == 0u;
    Mover A = pick(p2_first, L.p2, L.p1);
    Mover B = pick(p2_first, L.p1, L.p2);
    do_move<DGEN, NPC, EV, 0>(P, tiles, A, B, s, L.tick, lane, ev, cnt);
    do_move<DGEN, NPC, EV, 1>(P, tiles, B, A, s, L.tick, lane, ev, cnt);
    L.p1 = pick(p2_first, B, A);
    L.p2 = pick(p2_first, A, B);
    if (NPC) {   // dead NPCs leave in reverse entity order (updater.py:137-145)
        for (int k = P.n_npc - 1; k >= 0; --k) {
            const long long j = lane * P.n_npc + k;
            if (P.npc_depth[j] >= 0 && P.npc_hp[j] <= 0) {
                ev.emit(ORX_EV_DEATH, 3 + k, 0, 0, 0);
                P.npc_depth[j] = -1;
            }
        }
    }
    L.tick += 1;                                                         // updater.py:148
    int res = ORX_RESULT_IN_PROGRESS;
    if (P.max_ticks != 0 && L.tick >= P.max_ticks) res = ORX_RESULT_TIE; // :158
    if (L.p2.hp <= 0) res = ORX_RESULT_PLAYER1_WIN;                      // :155-157
    if (L.p1.hp <= 0) res = L.p2.hp <= 0 ? ORX_RESULT_TIE : ORX_RESULT_PLAYER2_WIN;  // :151-154
    return res;
}

// Re-initialise a lane for the episode already stored in s.episode.
template <int DGEN, bool NPC>
__device__ __forceinline__ void reset_lane(const Params& P, Lane& L, const Stream& s, long long lane)
{
    const uint2 r = reset_draw<DGEN>(P, s);
    L.p1.x = r.x & 255; L.p1.y = (r.x >> 8) & 255; L.p2.x = (r.x >> 16) & 255; L.p2.y = r.x >> 24;
    L.p1.sx = r.y & 255; L.p1.sy = (r.y >> 8) & 255; L.p2.sx = (r.y >> 16) & 255; L.p2.sy = r.y >> 24;
    L.p1.hp = P.hp0; L.p2.hp = P.hp1;
    L.p1.depth = P.sd0;
    L.p2.depth = P.start_kind == ORX_START_SEPARATED ? P.sd1 : P.sd0;
    L.tick = 1;                                                          // worldgen.py:87
    L.episode = s.episode;
    if (NPC) for (int k = 0; k < P.n_npc; ++k) P.npc_depth[lane * P.n_npc + k] = -1;
}

__device__ __forceinline__ void load_lane(const Params& P, long long i, Lane& L)
{
    const uint32_t pos = P.pos[i], hp = P.hp[i], st = P.stairs[i];
    const int2 d = P.depth[i];
    L.p1.x = pos & 255; L.p1.y = (pos >> 8) & 255; L.p2.x = (pos >> 16) & 255; L.p2.y = pos >> 24;
    L.p1.hp = (int)(int16_t)(hp & 0xFFFF); L.p2.hp = (int)(int16_t)(hp >> 16);
    L.p1.depth = d.x; L.p2.depth = d.y;
    L.p1.sx = st & 255; L.p1.sy = (st >> 8) & 255; L.p2.sx = (st >> 16) & 255; L.p2.sy = st >> 24;
    L.tick = P.tick[i];
    L.episode = P.episode[i];
}

__device__ __forceinline__ void store_lane(const Params& P, long long i, const Lane& L, int status)
{
    P.pos[i] = (uint32_t)L.p1.x | ((uint32_t)L.p1.y << 8) | ((uint32_t)L.p2.x << 16) | ((uint32_t)L.p2.y << 24);
    P.hp[i] = ((uint32_t)L.p1.hp & 0xFFFFu) | ((uint32_t)L.p2.hp << 16);
    P.depth[i] = make_int2(L.p1.depth, L.p2.depth);
    P.stairs[i] = (uint32_t)L.p1.sx | ((uint32_t)L.p1.sy << 8) | ((uint32_t)L.p2.sx << 16) | ((uint32_t)L.p2.sy << 24);
    P.tick[i] = L.tick;
    P.episode[i] = L.episode;
    P.status[i] = (uint8_t)status;
}

__device__ __forceinline__ Stream make_stream(const Params& P, long long i, uint32_t episode)
{
    const unsigned long long gid = P.gid_base + (unsigned long long)i;
    Stream s;
    s.k0 = P.k0; s.k1 = P.k1; s.g0 = (uint32_t)gid; s.g1 = (uint32_t)(gid >> 32); s.episode = episode;
    return s;
}

// randombot.py:20-21 / staircasebot.py:9-20. w is this player's word of the tick's main block.
__device__ __forceinline__ int bot_move(int kind, const Mover& m, uint32_t w)
{
    if (kind == ORX_BOT_RANDOM) return 1 + (int)bounded(w, 5u);
    if (kind == ORX_BOT_STAIRCASE) {
        const int dx = m.sx - m.x, dy = m.sy - m.y;
        if (abs(dx) > abs(dy)) return dx > 0 ? ORX_MOVE_RIGHT : ORX_MOVE_LEFT;
        return dy > 0 ? ORX_MOVE_DOWN : ORX_MOVE_UP;
    }
    return ORX_MOVE_STAY;
}

}  // namespace orx
