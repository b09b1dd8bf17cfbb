// Data-parallel formulation of one Optimax Rogue tick: one thread owns one game and the whole
// game state lives in a handful of registers, in the same packed form the HBM planes use:
//
//   pos  = x1 | y1<<8 | x2<<16 | y2<<24      (one 16-bit "xy" per player)
//   st   = staircase xy of each player's level, same packing
//   a command is reduced to a signed xy delta (Up -256, Right +1, Down +256, Left -1, Stay 0)
//
// so that "move" is one add on the packed word, "is that tile occupied / the staircase" is one
// 16-bit compare, and "who goes first" is a byte permute (PRMT) of the two halves instead of a
// divergent branch or an array in local memory.
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
    RoundKeys rk;              // Philox round keys of the seed
    const uint8_t* tiles;      // DGEN_FIXED: uint8[W*H] x-major
    const uint16_t* ground;    // DGEN_FIXED: Ground tile list
    int n_ground, fsx, fsy;
    uint32_t* pos;
    uint32_t* hp;              // int16 hp1 | int16 hp2 << 16
    int2* depth;
    uint32_t* stairs;
    int* tick;
    uint32_t* episode;
    uint8_t* status;
    uint8_t* npc_pos;
    int16_t* npc_hp;
    int* npc_depth;
    const int8_t* flat;        // Modifier seam: int8[n][2][3] {flat_damage, flat_armor, flat_max_health} per player, or NULL
    unsigned int n;            // games in this launch (host chunks batches above 2^30)
    unsigned long long gid_base;
};

struct Counters {               // per-thread, reduced at the end of orx_rollout
    unsigned int ticks, p1, p2, ties, events, descents, hits;
};

// Game state of one lane in registers.
struct Lane {
    uint32_t pos, st;
    int hp1, hp2, d1, d2, tick;
    uint32_t episode;
};

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

// The NPC slots of ONE game, behind a small accessor interface so that the tick does not care where they live:
//   NoNpc        the configuration has no NPC slots (every loop over slots disappears)
//   NpcView      slot planes in HBM, slot count known at run time (one-thread-per-game kernels)
//   NpcStage<N>  (orx_pipe.cuh) slot planes of a tile in the shared-memory stage, N known at compile time: the slot
//                loops unroll onto constant shared-window offsets
struct NoNpc {
    static constexpr bool kPresent = false;
    __device__ __forceinline__ static constexpr int count() { return 0; }
    __device__ __forceinline__ int depth_at(int) const { return -1; }
    __device__ __forceinline__ uint32_t xy_at(int) const { return 0u; }
    __device__ __forceinline__ int hp_at(int) const { return 0; }
    __device__ __forceinline__ void set_hp(int, int) const {}
    __device__ __forceinline__ void set_depth(int, int) const {}
};

struct NpcView {
    static constexpr bool kPresent = true;
    uint8_t* pos;     // [n_npc][2]
    int16_t* hp;      // [n_npc]
    int* depth;       // [n_npc], -1 = empty slot
    int n;
    __device__ __forceinline__ int count() const { return n; }
    __device__ __forceinline__ int depth_at(int k) const { return depth[k]; }
    __device__ __forceinline__ uint32_t xy_at(int k) const { return (uint32_t)pos[2 * k] | ((uint32_t)pos[2 * k + 1] << 8); }
    __device__ __forceinline__ int hp_at(int k) const { return hp[k]; }
    __device__ __forceinline__ void set_hp(int k, int v) const { hp[k] = (int16_t)v; }
    __device__ __forceinline__ void set_depth(int k, int v) const { depth[k] = v; }
};

__device__ __forceinline__ NpcView npc_view(const Params& P, unsigned int lane)
{
    const size_t j = (size_t)lane * P.n_npc;
    return NpcView{P.npc_pos + 2 * j, P.npc_hp + j, P.npc_depth + j, P.n_npc};
}

// A hit on an NPC slot: the plane is int16, the reference's health a Python int -- saturate instead of wrapping, so
// that an NPC hit twice in one tick by hits near 32767 cannot come back to life.
template <class NV>
__device__ __forceinline__ void npc_hit(const NV& nv, int npc, int dmg)
{
    const int h = nv.hp_at(npc) - dmg;
    nv.set_hp(npc, h < -32768 ? -32768 : h);
}

// First slot (lowest index, the reference's entity order) standing on (depth, xy), xy = x | y << 8; -1 if none.
template <class NV>
__device__ __forceinline__ int npc_at(const NV& nv, int depth, uint32_t xy)
{
    int found = -1;
#pragma unroll
    for (int k = nv.count() - 1; k >= 0; --k)
        if (nv.depth_at(k) == depth && nv.xy_at(k) == xy) found = k;
    return found;
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
// (updater.py:282-285). Returns sx | sy<<8 | x<<16 | y<<24  (= new st half | new pos half << 16).
template <int DGEN, class NV>
__device__ __noinline__ uint32_t descend_draw(const Params& P, Stream s, int tick, int pid, int new_depth,
                                              uint32_t oxy, int odepth, const NV nv)
{
    int sx, sy, x, y;
    level_stairs<DGEN>(P, s, new_depth, sx, sy);
    const int ng = n_ground<DGEN>(P);
    int q = 0;
    bool taken;
    do {
        const int k = (int)seq_bounded(s, DOM_TICK, SUB_DESCEND + 64u * (uint32_t)pid, (uint32_t)tick, q++, (uint32_t)ng);
        kth_ground<DGEN>(P, sx, sy, k, x, y);
        taken = (odepth == new_depth) & (oxy == ((uint32_t)x | ((uint32_t)y << 8)));
        if (NV::kPresent) taken = taken || npc_at(nv, new_depth, (uint32_t)x | ((uint32_t)y << 8)) >= 0;
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

// ---------------------------------------------------------------- command -> clamped xy delta
// A 256-entry shared-memory table turns a raw command byte into its xy delta (0 for Stay and for
// codes outside Move, logic/moves.py:6-12) and, for empty rooms, into a (mask, value) pair such
// that the move is blocked iff (xy & mask) == value. Inside an empty room only the border blocks
// (updater.py:90-98, world.py:41-46): Up is blocked iff y == 1, Right iff x == W-2, Down iff
// y == H-2, Left iff x == 1; Stay/invalid use mask = value = 0, i.e. "always blocked" => delta 0.
struct CmdEntry { uint32_t maskval; int delta; };

__device__ __forceinline__ void build_cmd_lut(const Params& P, CmdEntry* lut, unsigned int tid, unsigned int nthreads)
{
    for (unsigned int m = tid; m < 256u; m += nthreads) {
        CmdEntry e{0u, 0};
        if (m == ORX_MOVE_UP) e = {0xFF00u | (0x0100u << 16), -256};
        else if (m == ORX_MOVE_RIGHT) e = {0x00FFu | ((uint32_t)(P.W - 2) << 16), 1};
        else if (m == ORX_MOVE_DOWN) e = {0xFF00u | ((uint32_t)(P.H - 2) << 24), 256};
        else if (m == ORX_MOVE_LEFT) e = {0x00FFu | (0x0001u << 16), -1};
        lut[m] = e;
    }
}

// Returns 0 for Stay, for invalid codes and for a move into a Wall or off the map, evaluated from
// the pre-tick position.
template <int DGEN>
__device__ __forceinline__ int clamped_delta(const Params& P, const uint8_t* tiles, const CmdEntry* lut, uint32_t m, uint32_t xy)
{
    const CmdEntry e = lut[m];
    if (DGEN == ORX_DGEN_EMPTY) {
        return ((xy & e.maskval & 0xFFFFu) == (e.maskval >> 16)) ? 0 : e.delta;
    } else {
        if (e.delta == 0) return 0;
        const bool odd = (m & 1u) != 0;
        const int x = (int)(xy & 255u) + (odd ? 0 : (int)(3u - m));
        const int y = (int)(xy >> 8) + (odd ? (int)(m - 2u) : 0);
        if ((unsigned)x >= (unsigned)P.W || (unsigned)y >= (unsigned)P.H) return 0;
        return tiles[x * P.H + y] != ORX_TILE_WALL ? e.delta : 0;
    }
}

template <int DGEN>
__device__ __forceinline__ bool is_stairs(const Params& P, const uint8_t* tiles, uint32_t txy, uint32_t sxy)
{
    if (DGEN == ORX_DGEN_EMPTY) return txy == sxy;
    return tiles[(txy & 255u) * P.H + (txy >> 8)] == ORX_TILE_STAIRCASE_DOWN;
}

// What a hit by player `pid` takes off its victim (updater.py:313: attacker.damage.value - attacker.armor.value, the
// attribles being base + the modifiers' flat bonuses, attribles.py:29-43). Only called from the combat branches.
// FLAT: whether this instantiation looks at the bonus plane at all. The tile-pipeline kernels do not (ptxas folds the
// combat branches into the straight-line tick, so even a never-taken look-up cost 16 predicated instructions per
// game-tick, 8 % of the kernel): a state that carries a bonus plane is ticked by the one-thread-per-game kernels.
template <bool FLAT>
__device__ __forceinline__ int hit_of(const Params& P, const Stream& s, int pid)
{
    int d = pid == 0 ? P.dmg0 : P.dmg1;
    if (FLAT && P.flat != nullptr) {
        const size_t i = (size_t)(s.g0 - (uint32_t)P.gid_base);      // lane of this launch (< 2^30)
        d += (int)P.flat[i * 6 + 3 * pid] - (int)P.flat[i * 6 + 3 * pid + 1];
    }
    return d;
}

// One Updater.update. mv = p1 command | p2 command << 8; w_init is word 2 of the tick's main
// block. Returns the UpdateResult.
template <int DGEN, class NV, bool EV, bool FLAT = false>
__device__ __forceinline__ int tick_lane(const Params& P, const uint8_t* tiles, const CmdEntry* lut, Lane& L, uint32_t mv,
                                         uint32_t w_init, const Stream& s, const NV& nv,
                                         EvSink<EV>& ev, Counters& cnt)
{
    constexpr bool NPC = NV::kPresent;
    const int dl1 = clamped_delta<DGEN>(P, tiles, lut, mv & 255u, L.pos & 0xFFFFu);
    const int dl2 = clamped_delta<DGEN>(P, tiles, lut, (mv >> 8) & 255u, L.pos >> 16);
    // random.shuffle([p1, p2]) (updater.py:114): j = randbelow(2) = w >> 31; j == 0 swaps => p2 first.
    // Roles: A acts first and lives in the LOW half of pos/st, B acts second in the HIGH half.
    const bool p2_first = (int)w_init >= 0;
    const uint32_t sel = p2_first ? 0x1032u : 0x3210u;
    uint32_t pos = __byte_perm(L.pos, 0u, sel);
    uint32_t st = __byte_perm(L.st, 0u, sel);
    const int dA = p2_first ? dl2 : dl1, dB = p2_first ? dl1 : dl2;
    // Depths stay in player order: the hot path only needs "both on the same level", and the role of
    // each depth is looked up where it is really used (descents, events).
    int d1 = L.d1, d2 = L.d2;
    bool same = d1 == d2;
    auto depA = [&]() { return p2_first ? d2 : d1; };
    auto depB = [&]() { return p2_first ? d1 : d2; };
    // Health likewise: a hit is subtracted from the victim's own field inside the (rare) combat branch.
    int hp1 = L.hp1, hp2 = L.hp2;
    const int idA = p2_first ? 1 : 0, idB = idA ^ 1;

    // ---- first mover (handle_move, updater.py:180-243)
    if (dA != 0) {
        const uint32_t tA = (pos + (uint32_t)dA) & 0xFFFFu;
        int npc = -1;
        if (same & (tA == (pos >> 16))) {
            // Occupied by B, who acts later: Block if B stays, else Flee. (Parry, updater.py:229-234,
            // would need B.pos + delta == B.pos with a non-Stay move: unreachable.)
            const int dmgA = hit_of<FLAT>(P, s, idA);
            if (dmgA > 0) { if (p2_first) hp1 -= dmgA; else hp2 -= dmgA; ++cnt.hits; }      // the victim is B
            ev.emit(ORX_EV_COMBAT, idA + 1, idB + 1, dB == 0 ? ORX_FLAG_BLOCK : ORX_FLAG_FLEE, dmgA);
        } else if (NPC && (npc = npc_at(nv, depA(), tA)) >= 0) {
            const int dmgA = hit_of<FLAT>(P, s, idA);
            if (dmgA > 0) { npc_hit(nv, npc, dmgA); ++cnt.hits; }
            ev.emit(ORX_EV_COMBAT, idA + 1, 3 + npc, ORX_FLAG_BLOCK, dmgA);
        } else if (is_stairs<DGEN>(P, tiles, tA, st & 0xFFFFu)) {
            const int nd = depA() + 1;
            const uint32_t r = descend_draw<DGEN, NV>(P, s, L.tick, idA, nd, pos >> 16, depB(), nv);
            if (!level_exists(P, idA, nd, depB())) ev.emit(ORX_EV_DUNGEON, 0, r & 255u, (r >> 8) & 255u, nd);
            ev.emit(ORX_EV_DESCEND, idA + 1, (r >> 16) & 255u, r >> 24, nd);
            if (p2_first) d2 = nd; else d1 = nd;
            same = d1 == d2;
            pos = (pos & 0xFFFF0000u) | (r >> 16);
            st = (st & 0xFFFF0000u) | (r & 0xFFFFu);
            ++cnt.descents;
        } else {
            ev.emit(ORX_EV_MOVE, idA + 1, tA & 255u, tA >> 8, depA());
            pos += (uint32_t)dA;          // the clamp keeps the target on the map: no carry between halves
        }
    }
    // ---- second mover: sees A's updated position and depth
    if (dB != 0) {
        const uint32_t tB = ((pos >> 16) + (uint32_t)dB) & 0xFFFFu;
        int npc = -1;
        if (same & (tB == (pos & 0xFFFFu))) {
            const int dmgB = hit_of<FLAT>(P, s, idB);
            if (dmgB > 0) { if (p2_first) hp2 -= dmgB; else hp1 -= dmgB; ++cnt.hits; }      // the victim is A
            ev.emit(ORX_EV_COMBAT, idB + 1, idA + 1, dA == 0 ? ORX_FLAG_BLOCK : ORX_FLAG_AMBUSH, dmgB);
        } else if (NPC && (npc = npc_at(nv, depB(), tB)) >= 0) {
            const int dmgB = hit_of<FLAT>(P, s, idB);
            if (dmgB > 0) { npc_hit(nv, npc, dmgB); ++cnt.hits; }
            ev.emit(ORX_EV_COMBAT, idB + 1, 3 + npc, ORX_FLAG_BLOCK, dmgB);
        } else if (is_stairs<DGEN>(P, tiles, tB, st >> 16)) {
            const int nd = depB() + 1;
            const uint32_t r = descend_draw<DGEN, NV>(P, s, L.tick, idB, nd, pos & 0xFFFFu, depA(), nv);
            if (!level_exists(P, idB, nd, depA())) ev.emit(ORX_EV_DUNGEON, 0, r & 255u, (r >> 8) & 255u, nd);
            ev.emit(ORX_EV_DESCEND, idB + 1, (r >> 16) & 255u, r >> 24, nd);
            if (p2_first) d1 = nd; else d2 = nd;
            pos = (pos & 0xFFFFu) | (r & 0xFFFF0000u);
            st = (st & 0xFFFFu) | (r << 16);
            ++cnt.descents;
        } else {
            ev.emit(ORX_EV_MOVE, idB + 1, tB & 255u, tB >> 8, depB());
            pos += (uint32_t)dB << 16;
        }
    }
    // ---- back to player order
    L.pos = __byte_perm(pos, 0u, sel);
    L.st = __byte_perm(st, 0u, sel);
    L.d1 = d1; L.d2 = d2;
    L.hp1 = hp1; L.hp2 = hp2;
    if (NPC) {   // dead NPCs leave in reverse entity order (updater.py:137-145)
#pragma unroll
        for (int k = nv.count() - 1; k >= 0; --k) {
            if (nv.depth_at(k) >= 0 && nv.hp_at(k) <= 0) {
                ev.emit(ORX_EV_DEATH, 3 + k, 0, 0, 0);
                nv.set_depth(k, -1);
            }
        }
    }
    L.tick += 1;                                                         // updater.py:148
    int res = ORX_RESULT_IN_PROGRESS;
    if (P.max_ticks != 0 && L.tick >= P.max_ticks) res = ORX_RESULT_TIE; // :158
    if (L.hp2 <= 0) res = ORX_RESULT_PLAYER1_WIN;                        // :155-157
    if (L.hp1 <= 0) res = L.hp2 <= 0 ? ORX_RESULT_TIE : ORX_RESULT_PLAYER2_WIN;  // :151-154
    return res;
}

// Re-initialise a lane for the episode already stored in s.episode.
template <int DGEN, class NV>
__device__ __forceinline__ void reset_lane(const Params& P, Lane& L, const Stream& s, const NV& nv)
{
    const uint2 r = reset_draw<DGEN>(P, s);
    L.pos = r.x; L.st = r.y;
    L.hp1 = P.hp0; L.hp2 = P.hp1;
    L.d1 = P.sd0;
    L.d2 = P.start_kind == ORX_START_SEPARATED ? P.sd1 : P.sd0;
    L.tick = 1;                                                          // worldgen.py:87
    L.episode = s.episode;
    if (NV::kPresent) {
#pragma unroll
        for (int k = 0; k < nv.count(); ++k) nv.set_depth(k, -1);
    }
}

__device__ __forceinline__ void unpack_lane(Lane& L, uint32_t pos, uint32_t hp, int2 d, uint32_t st, int tick, uint32_t ep)
{
    L.pos = pos; L.st = st;
    L.hp1 = (int)(int16_t)(hp & 0xFFFFu); L.hp2 = (int)hp >> 16;
    L.d1 = d.x; L.d2 = d.y;
    L.tick = tick; L.episode = ep;
}

__device__ __forceinline__ void load_lane(const Params& P, unsigned int i, Lane& L)
{
    unpack_lane(L, P.pos[i], P.hp[i], P.depth[i], P.stairs[i], P.tick[i], P.episode[i]);
}

__device__ __forceinline__ void store_lane(const Params& P, unsigned int i, const Lane& L, int status)
{
    P.pos[i] = L.pos;
    P.hp[i] = ((uint32_t)L.hp1 & 0xFFFFu) | ((uint32_t)L.hp2 << 16);
    P.depth[i] = make_int2(L.d1, L.d2);
    P.stairs[i] = L.st;
    P.tick[i] = L.tick;
    P.episode[i] = L.episode;
    P.status[i] = (uint8_t)status;
}

// Per-player observation (GameState.view_for, state.py:53-58): 12 int16 per player, both players of a
// game = 12 words. { x, y, depth, hp, opp_visible, opp_x, opp_y, opp_hp, stairs_visible, stairs_x,
// stairs_y, tick }, depth and tick saturated at 32767; radius < 0 = the staircase is always visible.
__device__ __forceinline__ uint32_t pack_i16x2(int lo, int hi) { return ((uint32_t)lo & 0xFFFFu) | ((uint32_t)hi << 16); }

__device__ __forceinline__ void pack_obs(const Lane& L, int radius, uint32_t (&w)[12])
{
    const bool same = L.d1 == L.d2;   // view_for keeps entities on the viewer's depth
    const int tk = min(L.tick, 32767);
#pragma unroll
    for (int p = 0; p < 2; ++p) {
        const uint32_t me = p == 0 ? (L.pos & 0xFFFFu) : (L.pos >> 16), ot = p == 0 ? (L.pos >> 16) : (L.pos & 0xFFFFu);
        const uint32_t sxy = p == 0 ? (L.st & 0xFFFFu) : (L.st >> 16);
        const int mx = me & 255u, my = me >> 8, sx = sxy & 255u, sy = sxy >> 8;
        const int md = p == 0 ? L.d1 : L.d2, mh = p == 0 ? L.hp1 : L.hp2, oh = p == 0 ? L.hp2 : L.hp1;
        const bool has_st = sx != ORX_NO_STAIRS;
        const bool st_vis = has_st && (radius < 0 || max(abs(sx - mx), abs(sy - my)) <= radius);
        w[6 * p + 0] = pack_i16x2(mx, my);
        w[6 * p + 1] = pack_i16x2(min(md, 32767), mh);
        w[6 * p + 2] = pack_i16x2(same, same ? (int)(ot & 255u) : -1);
        w[6 * p + 3] = pack_i16x2(same ? (int)(ot >> 8) : -1, same ? oh : 0);
        w[6 * p + 4] = pack_i16x2(st_vis, st_vis ? sx : -1);
        w[6 * p + 5] = pack_i16x2(st_vis ? sy : -1, tk);
    }
}

__device__ __forceinline__ Stream make_stream(const Params& P, unsigned int i, uint32_t episode)
{
    const unsigned long long gid = P.gid_base + (unsigned long long)i;
    Stream s;
    s.rk = &P.rk; s.g0 = (uint32_t)gid; s.g1 = (uint32_t)(gid >> 32); s.episode = episode;
    return s;
}

// randombot.py:20-21 / staircasebot.py:9-20. w is this player's word of the tick's main block;
// xy / sxy are the player's packed position and its level's staircase.
__device__ __forceinline__ uint32_t bot_move(int kind, uint32_t xy, uint32_t sxy, uint32_t w)
{
    if (kind == ORX_BOT_RANDOM) return 1u + bounded(w, 5u);
    if (kind == ORX_BOT_STAIRCASE) {
        const int dx = (int)(sxy & 255u) - (int)(xy & 255u), dy = (int)(sxy >> 8) - (int)(xy >> 8);
        if (abs(dx) > abs(dy)) return dx > 0 ? ORX_MOVE_RIGHT : ORX_MOVE_LEFT;
        return dy > 0 ? ORX_MOVE_DOWN : ORX_MOVE_UP;
    }
    return ORX_MOVE_STAY;
}

}  // namespace orx
