// Ruleset R1, thread-per-game formulation (included by orx_r1.cu).
//
// The half-warp kernels in orx_r1.cu spend most of their issue slots on work that is uniform across
// the sixteen lanes of a game. Here one thread owns a whole game: the ten movers are ten packed
// (depth<<16 | y<<8 | x) keys plus ten health registers, the four items four keys, and every loop
// over entities is fully unrolled so that all indices are compile-time constants and the state
// never leaves the register file. "Who stands on my target" / "who else wants it" are unrolled
// compare chains on the packed keys; damage is accumulated straight into the victim's register
// inside those chains, so there is no gather pass. Same planes, same draws, same results as
// oracle/orx_r1_oracle.c.
#pragma once

namespace r1t {

constexpr uint32_t DEAD = 0xFFFFFFFFu;

struct Game {
    uint32_t key[NM];     // movers: depth<<16 | y<<8 | x, DEAD when the slot is empty
    int hp[NM];
    uint32_t ikey[ORX_R1_ITEMS];   // ground items, same packing
    int ikind[ORX_R1_ITEMS];
    int mana[2], max_hp[2], max_mana[2], xp[2], level[2], n_items[2], cd[2], damage[2], armor[2];
    int sx[2], sy[2];
    uint32_t lkey[2];
    int sep, tick;
    uint32_t episode;
};

__device__ __forceinline__ int kx(uint32_t k) { return (int)(k & 255u); }
__device__ __forceinline__ int ky(uint32_t k) { return (int)((k >> 8) & 255u); }
__device__ __forceinline__ int kd(uint32_t k) { return (int)(k >> 16); }
__device__ __forceinline__ uint32_t mk(int depth, int x, int y) { return ((uint32_t)depth << 16) | ((uint32_t)y << 8) | (uint32_t)x; }

__device__ __forceinline__ bool occupied(const Game& G, uint32_t k, int skip)
{
    bool o = false;
#pragma unroll
    for (int j = 0; j < NM; ++j) o |= (j != skip) & (G.key[j] == k);
    return o;
}

// Random free tile on a level (docs/RULESET_R1.md): two words per try, then the x-major scan.
// `skip` is a mover slot that does not count as an obstacle (the descending player itself).
__device__ __forceinline__ uint32_t free_tile(const R1Params& P, const Game& G, const Stream& s, uint32_t domain, uint32_t sub_base,
                                              uint32_t index, int depth, uint32_t lkey, int sx, int sy, int skip)
{
    for (uint32_t r = 0; r < R1_MAX_TRIES; ++r) {
        const uint4 b = draw_block(s, domain, sub_base + (r >> 1), index);
        const int x = 1 + (int)bounded((r & 1) ? b.z : b.x, (uint32_t)(P.W - 2));
        const int y = 1 + (int)bounded((r & 1) ? b.w : b.y, (uint32_t)(P.H - 2));
        if (!is_wall(P, lkey, sx, sy, x, y) && !(x == sx && y == sy) && !occupied(G, mk(depth, x, y), skip)) return mk(depth, x, y);
    }
    for (int x = 1; x < P.W - 1; ++x)
        for (int y = 1; y < P.H - 1; ++y)
            if (!is_wall(P, lkey, sx, sy, x, y) && !(x == sx && y == sy) && !occupied(G, mk(depth, x, y), skip)) return mk(depth, x, y);
    return mk(depth, 1, 1);
}

__device__ __forceinline__ void setup_game(const R1Params& P, Game& G, const Stream& s)
{
#pragma unroll
    for (int m = 0; m < NM; ++m) { G.key[m] = DEAD; G.hp[m] = 0; }
#pragma unroll
    for (int i = 0; i < ORX_R1_ITEMS; ++i) { G.ikey[i] = DEAD; G.ikind[i] = 0; }
    level_init(P, s, 0, G.sx[0], G.sy[0], G.lkey[0]);
    G.sx[1] = G.sx[0]; G.sy[1] = G.sy[0]; G.lkey[1] = G.lkey[0];
    // ONE copy of the free-tile search for both players (a rolled loop; the slot is written through selects): the rare
    // paths are what made the tick's body 6k instructions, and a warp runs a rare path whenever one of its 32 games does
#pragma unroll 1
    for (int p = 0; p < 2; ++p) {
        const uint32_t k = free_tile(P, G, s, DOM_RESET, 32u * (uint32_t)p, 0u, 0, G.lkey[0], G.sx[0], G.sy[0], -1);
        if (p == 0) G.key[0] = k; else G.key[1] = k;
    }
#pragma unroll
    for (int p = 0; p < 2; ++p) {
        G.hp[p] = 10; G.mana[p] = 9; G.max_hp[p] = 10; G.max_mana[p] = 9; G.xp[p] = 0; G.level[p] = 1;
        G.n_items[p] = 0; G.cd[p] = 0; G.damage[p] = 2; G.armor[p] = 1;
    }
    G.sep = 0; G.tick = 1;
}

// element m of a register array through a select chain (m is a run-time value, the array stays in registers)
template <int N>
__device__ __forceinline__ uint32_t pick(const uint32_t (&a)[N], int m)
{
    uint32_t v = a[0];
#pragma unroll
    for (int j = 1; j < N; ++j) v = m == j ? a[j] : v;
    return v;
}

// Replication log of one game-tick (include/orx.h "Replication log of an R1 tick"): records go straight to the game's
// slots in global memory, in emission order; without EV every call folds away.
template <bool EV>
struct Sink {
    uint2* base;
    int n, cap;
    __device__ __forceinline__ void emit(int kind, int iden, int a, int b, int value)
    {
        if (EV) {
            if (n < cap) base[n] = make_uint2((uint32_t)kind | ((uint32_t)iden << 8) | ((uint32_t)(a & 255) << 16) | ((uint32_t)(b & 255) << 24), (uint32_t)value);
            ++n;
        }
    }
    __device__ __forceinline__ void finish() { if (EV) if (n < cap) base[n] = make_uint2(0u, 0u); }
};

// One tick (docs/RULESET_R1.md, oracle/orx_r1_oracle.c:r1_tick). Returns the UpdateResult.
template <bool EV = false>
__device__ __forceinline__ int tick(const R1Params& P, Game& G, const Stream& s, int c1, int c2, R1Counters& cnt, Sink<EV>& ev)
{
    int dl[NM];                                    // packed delta on the key: dx + 256*dy, 0 = stays
    const int cd_pre0 = G.cd[0], cd_pre1 = G.cd[1];
    // 1. heal, 2. intents
#pragma unroll
    for (int p = 0; p < 2; ++p) {
        int c = p == 0 ? c1 : c2;
        if (c == ORX_MOVE_HEAL) {
            const int h = min(G.mana[p], G.max_mana[p] / 3), before = G.hp[p];
            G.hp[p] = min(G.max_hp[p], G.hp[p] + h);
            G.mana[p] -= h;
            if (h > 0) ev.emit(ORX_EV_HEALTH, p + 1, p + 1, ORX_R1_HEALTH_HEAL, G.hp[p] - before);
            c = ORX_MOVE_STAY;
        }
        const int ddx = (c == ORX_MOVE_RIGHT) - (c == ORX_MOVE_LEFT), ddy = (c == ORX_MOVE_DOWN) - (c == ORX_MOVE_UP);
        dl[p] = 0;
        if ((ddx | ddy) != 0 && !is_wall(P, G.lkey[p], G.sx[p], G.sy[p], kx(G.key[p]) + ddx, ky(G.key[p]) + ddy)) dl[p] = ddx + 256 * ddy;
    }
    const int p0x = kx(G.key[0]), p0y = ky(G.key[0]), p0d = kd(G.key[0]);
    const int p1x = kx(G.key[1]), p1y = ky(G.key[1]), p1d = kd(G.key[1]);
#pragma unroll
    for (int m = 2; m < NM; ++m) {               // decide_npc_move: chase the nearest player on this depth
        dl[m] = 0;
        if (G.key[m] == DEAD) continue;
        const int x = kx(G.key[m]), y = ky(G.key[m]), d = kd(G.key[m]);
        const int m0 = p0d == d ? abs(p0x - x) + abs(p0y - y) : (1 << 30);
        const int m1 = p1d == d ? abs(p1x - x) + abs(p1y - y) : (1 << 30);
        if (min(m0, m1) == (1 << 30)) continue;
        const bool t1 = m1 < m0;
        const int ex = (t1 ? p1x : p0x) - x, ey = (t1 ? p1y : p0y) - y;
        if (max(abs(ex), abs(ey)) > 6) continue;
        int ddx = 0, ddy = 0;
        if (abs(ex) > abs(ey)) ddx = ex > 0 ? 1 : -1; else ddy = ey > 0 ? 1 : -1;
        const bool l0 = p0d == d;
        const int sx = l0 ? G.sx[0] : G.sx[1], sy = l0 ? G.sy[0] : G.sy[1];
        const int tx = x + ddx, ty = y + ddy;
        if (!is_wall(P, l0 ? G.lkey[0] : G.lkey[1], sx, sy, tx, ty) && !(tx == sx && ty == sy)) dl[m] = ddx + 256 * ddy;
    }
    // 3. cooldown conversion: an attack attempted while on cooldown is measured as Stay
#pragma unroll
    for (int p = 0; p < 2; ++p)
        if ((p == 0 ? cd_pre0 : cd_pre1) > 0 && dl[p] != 0 && occupied(G, G.key[p] + (uint32_t)dl[p], p)) dl[p] = 0;
    // 4. attacks, all from the start-of-tick keys; damage lands in taken[] of the victim
    int taken[NM], credit[NM];
    bool moves[NM];
    int newcd0 = 0, newcd1 = 0;
    bool spend0 = false, spend1 = false;
    uint32_t tgt[NM];      // target key of an acting mover; a per-slot sentinel no key can equal otherwise
#pragma unroll
    for (int j = 0; j < NM; ++j) {
        taken[j] = 0; credit[j] = 0; moves[j] = false;
        tgt[j] = (G.key[j] != DEAD && dl[j] != 0) ? G.key[j] + (uint32_t)dl[j] : (0xFFFF0000u | (uint32_t)j);
    }
#pragma unroll
    for (int m = 0; m < NM; ++m) {
        const uint32_t t = tgt[m];
        if (t >= 0xFFFF0000u) continue;
        const int my_dmg = m < 2 ? G.damage[m] + min(G.mana[m], G.max_mana[m] / 3) : 2 + kd(G.key[m]) / 4;
        const bool m_on_cd = m < 2 && (m == 0 ? cd_pre0 : cd_pre1) > 0;
        bool has_occ = false;
#pragma unroll
        for (int o = 0; o < NM; ++o) {
            if (o == m) continue;
            if (G.key[o] != t) continue;
            has_occ = true;
            if (m >= 2 && o >= 2) continue;                         // enemies do not fight each other
            const int amount = max(0, my_dmg - (o < 2 ? G.armor[o] : 0));
            const int o_cd = o < 2 ? (o == 0 ? cd_pre0 : cd_pre1) : 1;   // enemies never negate
            if (tgt[o] >= 0xFFFF0000u) {                            // the occupant stays
                if (o < 2 && o_cd == 0) {                           // negated; the attacker is stunned
                    if (m == 0) { newcd0 = max(newcd0, 1); spend0 = true; }
                    if (m == 1) { newcd1 = max(newcd1, 1); spend1 = true; }
                    ev.emit(ORX_EV_COMBAT, m + 1, o + 1, ORX_R1_HIT_NEGATED, 0);
                } else {
                    taken[o] += amount;
                    if (m == 0) { spend0 = true; if (amount > 0) credit[o] |= 1; }
                    if (m == 1) { spend1 = true; if (amount > 0) credit[o] |= 2; }
                    ev.emit(ORX_EV_COMBAT, m + 1, o + 1, ORX_R1_HIT_FULL, amount);
                }
            } else if (tgt[o] == G.key[m]) {                        // mutual attack: half damage, cooldown
                taken[o] += amount / 2;
                if (m == 0) { newcd0 = 3; spend0 = true; if (amount / 2 > 0) credit[o] |= 1; }
                if (m == 1) { newcd1 = 3; spend1 = true; if (amount / 2 > 0) credit[o] |= 2; }
                ev.emit(ORX_EV_COMBAT, m + 1, o + 1, ORX_R1_HIT_HALF, amount / 2);
            }
        }
        if (has_occ) continue;
        bool contested = false, hit = false;
#pragma unroll
        for (int c = 0; c < NM; ++c) {
            if (c == m) continue;
            if (tgt[c] != t) continue;
            contested = true;
            if (!hit && (m < 2 || c < 2) && !m_on_cd) {           // full damage on the contender's new location
                hit = true;
                const int amount = max(0, my_dmg - (c < 2 ? G.armor[c] : 0));
                taken[c] += amount;
                if (m == 0) { spend0 = true; if (amount > 0) credit[c] |= 1; }
                if (m == 1) { spend1 = true; if (amount > 0) credit[c] |= 2; }
                ev.emit(ORX_EV_COMBAT, m + 1, c + 1, ORX_R1_HIT_CONTEST, amount);
            }
        }
        moves[m] = !contested;
    }
    // 5. apply
    if (spend0) G.mana[0] -= min(G.mana[0], G.max_mana[0] / 3);
    if (spend1) G.mana[1] -= min(G.mana[1], G.max_mana[1] / 3);
#pragma unroll
    for (int m = 0; m < NM; ++m) {
        if (G.key[m] == DEAD) continue;
        G.hp[m] -= taken[m];
        cnt.hits += taken[m] > 0;
        if (moves[m]) { G.key[m] = tgt[m]; ev.emit(ORX_EV_MOVE, m + 1, kx(G.key[m]), ky(G.key[m]), kd(G.key[m])); }
    }
    uint32_t descending = 0;                     // players that reached their staircase this tick
#pragma unroll
    for (int p = 0; p < 2; ++p) {
        if (!moves[p]) continue;
        if (kx(G.key[p]) == G.sx[p] && ky(G.key[p]) == G.sy[p]) { descending |= 1u << p; continue; }
#pragma unroll
        for (int i = 0; i < ORX_R1_ITEMS; ++i)                                     // pickup
            if (G.ikey[i] == G.key[p] && G.n_items[p] < 4) {
                if (G.ikind[i] == 0) G.damage[p] += 1; else if (G.ikind[i] == 1) G.armor[p] += 1; else { G.max_hp[p] += 2; G.hp[p] += 2; }
                G.n_items[p] += 1;
                G.ikey[i] = DEAD;
                ev.emit(ORX_EV_PICKUP, p + 1, NM + i + 1, G.ikind[i], G.ikind[i] == 2 ? 2 : 0);
            }
    }
    // Descents, player 1 before player 2 (the second search sees the first player's new tile), as ONE copy of the
    // level draw and the free-tile search: games in which player 1 descends and games in which player 2 does run
    // the same instructions together. Pickups above touch neither the movers' keys nor the levels, so taking them
    // first changes nothing.
    while (descending != 0u) {
        const int p = __ffs((int)descending) - 1;
        descending &= descending - 1u;
        const int nd = kd(p == 0 ? G.key[0] : G.key[1]) + 1;
        int nsx, nsy; uint32_t nlk;
        level_init(P, s, nd, nsx, nsy, nlk);
        const uint32_t nk = free_tile(P, G, s, DOM_TICK, SUB_DESCEND + 64u * (uint32_t)p, (uint32_t)G.tick, nd, nlk, nsx, nsy, p);
        if (kd(p == 0 ? G.key[1] : G.key[0]) != nd) ev.emit(ORX_EV_DUNGEON, 0, nsx, nsy, nd);      // nobody stood there: the level is new
        ev.emit(ORX_EV_DESCEND, p + 1, kx(nk), ky(nk), nd);
        if (p == 0) { G.sx[0] = nsx; G.sy[0] = nsy; G.lkey[0] = nlk; G.key[0] = nk; }
        else { G.sx[1] = nsx; G.sy[1] = nsy; G.lkey[1] = nlk; G.key[1] = nk; }
        ++cnt.descents;
    }
    // 6. enemy deaths in slot order: xp, level-ups, drops
    uint32_t died = 0;                            // enemy slots that died this tick
#pragma unroll
    for (int m = 2; m < NM; ++m) {
        if (G.key[m] == DEAD || G.hp[m] > 0) continue;
        died |= 1u << m;
        ev.emit(ORX_EV_DEATH, m + 1, 0, 0, 0);
#pragma unroll
        for (int p = 0; p < 2; ++p)
            if ((credit[m] >> p) & 1) {
                int gained = 0;
                G.xp[p] += 1;
                while (G.xp[p] >= 3) { G.xp[p] -= 3; G.level[p] += 1; G.hp[p] = G.max_hp[p]; G.mana[p] = G.max_mana[p]; ++gained; }
                ev.emit(ORX_EV_XP, p + 1, m + 1, gained, G.xp[p]);
            }
    }
    // Drops in slot order, ONE copy of the draw: a warp in which enemy 2 of one game and enemy 5 of another die runs it
    // once for both (unrolled per slot it was eight Philox blocks in the instruction stream, each run for one lane).
    // xp and level-ups above read nothing a drop writes, and a drop reads nothing they write.
    for (uint32_t left = died; left != 0u; left &= left - 1u) {
        const int m = __ffs((int)left) - 1;
        const uint32_t where = pick(G.key, m);
        const uint4 b = draw_block(s, DOM_TICK, SUB_DROP + (uint32_t)((m - 2) >> 1), (uint32_t)G.tick);
        const uint32_t chance = ((m - 2) & 1) ? b.z : b.x, kind = (((m - 2) & 1) ? b.w : b.y) % 3u;
        if (chance < (1u << 30)) {
            bool placed = false;
#pragma unroll
            for (int i = 0; i < ORX_R1_ITEMS; ++i)
                if (!placed && G.ikey[i] == DEAD) {
                    placed = true; G.ikey[i] = where; G.ikind[i] = (int)kind;
                    ev.emit(ORX_EV_SPAWN, NM + i + 1, kx(where), ky(where), kd(where) | ((int)kind << 16));
                }
        }
    }
#pragma unroll
    for (int m = 2; m < NM; ++m) if ((died >> m) & 1u) G.key[m] = DEAD;
    // 7. levels without a player vanish; spawn roll every fourth tick
    const uint32_t d0 = G.key[0] >> 16, d1 = G.key[1] >> 16;
#pragma unroll
    for (int m = 2; m < NM; ++m) if (G.key[m] != DEAD && (G.key[m] >> 16) != d0 && (G.key[m] >> 16) != d1) { G.key[m] = DEAD; ev.emit(ORX_EV_DEATH, m + 1, 1, 0, 0); }
#pragma unroll
    for (int i = 0; i < ORX_R1_ITEMS; ++i) if (G.ikey[i] != DEAD && (G.ikey[i] >> 16) != d0 && (G.ikey[i] >> 16) != d1) { G.ikey[i] = DEAD; ev.emit(ORX_EV_DEATH, NM + i + 1, 1, 0, 0); }
    if ((G.tick & 3) == 0) {
        const uint4 sb = draw_block(s, DOM_TICK, SUB_SPAWN, (uint32_t)G.tick);
#pragma unroll 1                  // one copy of the search for both levels
        for (int p = 0; p < 2; ++p) {
            if (p == 1 && d1 == d0) continue;
            if ((p == 0 ? sb.x : sb.y) >= (1u << 30)) continue;
            bool any_free = false;
#pragma unroll
            for (int m = 2; m < NM; ++m) any_free |= G.key[m] == DEAD;
            if (!any_free) continue;
            const int d = (int)(p == 0 ? d0 : d1);
            const uint32_t t = free_tile(P, G, s, DOM_TICK, SUB_SPAWN_TRY + 32u * (uint32_t)p, (uint32_t)G.tick, d, p == 0 ? G.lkey[0] : G.lkey[1],
                                         p == 0 ? G.sx[0] : G.sx[1], p == 0 ? G.sy[0] : G.sy[1], -1);
            bool placed = false;
#pragma unroll
            for (int m = 2; m < NM; ++m)
                if (!placed && G.key[m] == DEAD) {
                    placed = true; G.key[m] = t; G.hp[m] = min(20, 2 + d / 2);
                    ev.emit(ORX_EV_SPAWN, m + 1, kx(t), ky(t), d | (G.hp[m] << 16));
                }
        }
    }
    // 8. mana, 9. separation, 10. cooldowns
    if ((G.tick & 3) == 0) { G.mana[0] = min(G.max_mana[0], G.mana[0] + 1); G.mana[1] = min(G.max_mana[1], G.mana[1] + 1); }
    if (d0 != d1) {
        G.sep += 1;
        if (d0 < d1) G.hp[0] -= G.sep / 16; else G.hp[1] -= G.sep / 16;
        if (G.sep / 16 > 0) ev.emit(ORX_EV_HEALTH, d0 < d1 ? 1 : 2, d0 < d1 ? 1 : 2, ORX_R1_HEALTH_SEPARATION, -(G.sep / 16));
    } else G.sep = 0;
    G.cd[0] = cd_pre0 > 0 ? cd_pre0 - 1 : newcd0;
    G.cd[1] = cd_pre1 > 0 ? cd_pre1 - 1 : newcd1;
    const int tick_pre = G.tick;
    G.tick += 1;
    const bool dead0 = G.hp[0] <= 0, dead1 = G.hp[1] <= 0;
    if (dead0) ev.emit(ORX_EV_DEATH, 1, 0, 0, 0);
    if (dead1) ev.emit(ORX_EV_DEATH, 2, 0, 0, 0);
    int res = ORX_RESULT_IN_PROGRESS;
    if (P.max_ticks != 0 && G.tick >= P.max_ticks) res = ORX_RESULT_TIE;
    if (dead1) res = ORX_RESULT_PLAYER1_WIN;
    if (dead0) res = ORX_RESULT_PLAYER2_WIN;
    if (dead0 && dead1) {
        const uint4 b = draw_block(s, DOM_TICK, SUB_MAIN, (uint32_t)tick_pre);
        res = (b.z >> 31) ? ORX_RESULT_PLAYER1_WIN : ORX_RESULT_PLAYER2_WIN;
    }
    return res;
}

__device__ __forceinline__ uint4 ldg128(const void* p)
{
    uint4 v;
    asm volatile("ld.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}
__device__ __forceinline__ uint2 ldg64(const void* p)
{
    uint2 v;
    asm volatile("ld.global.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p));
    return v;
}
__device__ __forceinline__ uint32_t w_of(const uint4 (&q)[4], int l) { const uint4& v = q[l >> 2]; return (l & 3) == 0 ? v.x : (l & 3) == 1 ? v.y : (l & 3) == 2 ? v.z : v.w; }

// A game's 16-lane rows are 64 contiguous bytes per plane: four 16-byte loads each.
__device__ __forceinline__ void load_game(const R1Params& P, unsigned int game, Game& G)
{
    uint4 loc[4], dep[4], stat[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        loc[q] = ldg128(P.ent_loc + (size_t)game * 16 + 4 * q);
        dep[q] = ldg128(P.ent_depth + (size_t)game * 16 + 4 * q);
        stat[q] = ldg128(P.ent_stat + (size_t)game * 16 + 4 * q);
    }
    const uint2 a = ldg64(P.pl_a + (size_t)game * 2), b = ldg64(P.pl_b + (size_t)game * 2), c = ldg64(P.pl_c + (size_t)game * 2);
    const uint2 lk = ldg64(P.lvl_key + (size_t)game * 2);
    const uint32_t sw = ldg32(P.lvl_stairs + game), sep = ldg32(P.sep + game), tick = ldg32(P.tick + game), ep = ldg32(P.episode + game);
#pragma unroll
    for (int m = 0; m < NM; ++m) {
        const uint32_t l = w_of(loc, m), st = w_of(stat, m);
        const bool alive = (l >> 16) & 1;
        G.key[m] = alive ? ((w_of(dep, m) << 16) | (l & 0xFFFFu)) : DEAD;
        G.hp[m] = (int)(int16_t)(st & 0xFFFF);
        if (m < 2) G.mana[m] = (int)(int16_t)(st >> 16);
    }
#pragma unroll
    for (int i = 0; i < ORX_R1_ITEMS; ++i) {
        const uint32_t l = w_of(loc, NM + i);
        G.ikey[i] = ((l >> 16) & 1) ? ((w_of(dep, NM + i) << 16) | (l & 0xFFFFu)) : DEAD;
        G.ikind[i] = (int)((l >> 17) & 3);
    }
#pragma unroll
    for (int p = 0; p < 2; ++p) {
        const uint32_t av = p == 0 ? a.x : a.y, bv = p == 0 ? b.x : b.y, cv = p == 0 ? c.x : c.y;
        G.max_hp[p] = (int)(int16_t)(av & 0xFFFF); G.max_mana[p] = (int)(int16_t)(av >> 16);
        G.xp[p] = bv & 255; G.level[p] = (bv >> 8) & 255; G.n_items[p] = (bv >> 16) & 255; G.cd[p] = bv >> 24;
        G.damage[p] = cv & 255; G.armor[p] = (cv >> 8) & 255;
        G.sx[p] = (sw >> (16 * p)) & 255; G.sy[p] = (sw >> (16 * p + 8)) & 255;
    }
    G.lkey[0] = lk.x; G.lkey[1] = lk.y;
    G.sep = (int)sep; G.tick = (int)tick; G.episode = ep;
}

__device__ __forceinline__ void store_game(const R1Params& P, unsigned int game, const Game& G, int status)
{
    uint32_t loc[16], dep[16], stat[16];
#pragma unroll
    for (int m = 0; m < NM; ++m) {
        const bool alive = G.key[m] != DEAD;
        loc[m] = alive ? ((G.key[m] & 0xFFFFu) | (1u << 16)) : 0u;
        dep[m] = alive ? (G.key[m] >> 16) : 0u;
        stat[m] = alive ? (((uint32_t)G.hp[m] & 0xFFFFu) | (m < 2 ? ((uint32_t)G.mana[m] << 16) : 0u)) : 0u;
    }
#pragma unroll
    for (int i = 0; i < ORX_R1_ITEMS; ++i) {
        const bool alive = G.ikey[i] != DEAD;
        loc[NM + i] = alive ? ((G.ikey[i] & 0xFFFFu) | (1u << 16) | ((uint32_t)(G.ikind[i] & 3) << 17)) : 0u;
        dep[NM + i] = alive ? (G.ikey[i] >> 16) : 0u;
        stat[NM + i] = 0u;
    }
    loc[14] = loc[15] = dep[14] = dep[15] = stat[14] = stat[15] = 0u;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        reinterpret_cast<uint4*>(P.ent_loc + (size_t)game * 16)[q] = make_uint4(loc[4 * q], loc[4 * q + 1], loc[4 * q + 2], loc[4 * q + 3]);
        reinterpret_cast<uint4*>(P.ent_depth + (size_t)game * 16)[q] = make_uint4(dep[4 * q], dep[4 * q + 1], dep[4 * q + 2], dep[4 * q + 3]);
        reinterpret_cast<uint4*>(P.ent_stat + (size_t)game * 16)[q] = make_uint4(stat[4 * q], stat[4 * q + 1], stat[4 * q + 2], stat[4 * q + 3]);
    }
    uint32_t a[2], b[2], c[2];
#pragma unroll
    for (int p = 0; p < 2; ++p) {
        a[p] = ((uint32_t)G.max_hp[p] & 0xFFFFu) | ((uint32_t)G.max_mana[p] << 16);
        b[p] = (uint32_t)(G.xp[p] & 255) | ((uint32_t)(G.level[p] & 255) << 8) | ((uint32_t)(G.n_items[p] & 255) << 16) | ((uint32_t)G.cd[p] << 24);
        c[p] = (uint32_t)(G.damage[p] & 255) | ((uint32_t)(G.armor[p] & 255) << 8);
    }
    reinterpret_cast<uint2*>(P.pl_a)[game] = make_uint2(a[0], a[1]);
    reinterpret_cast<uint2*>(P.pl_b)[game] = make_uint2(b[0], b[1]);
    reinterpret_cast<uint2*>(P.pl_c)[game] = make_uint2(c[0], c[1]);
    reinterpret_cast<uint2*>(P.lvl_key)[game] = make_uint2(G.lkey[0], G.lkey[1]);
    P.lvl_stairs[game] = (uint32_t)G.sx[0] | ((uint32_t)G.sy[0] << 8) | ((uint32_t)G.sx[1] << 16) | ((uint32_t)G.sy[1] << 24);
    P.sep[game] = (uint32_t)G.sep; P.tick[game] = G.tick; P.episode[game] = G.episode; P.status[game] = (uint8_t)status;
}

#ifndef ORX_R1T_THREADS
#define ORX_R1T_THREADS 128
#endif
// CTAs per SM the register allocation is held to. Five (96 registers, 100 bytes of spills in the rare paths) for the
// kernels whose launches overlap or whose threads live long -- 20 instead of 16 warps per SM hide more of the
// fixed-latency dependencies that bound this tick: 10.7 against 11.0 us per 65,536-game step in throughput mode, the
// fused rollout 5.7e9 against 4.3e9 game-ticks/s -- four (as many registers as it likes, up to 128) for a tick launched
// alone, which was 4 % slower with five (profiles/r02_r1_rare_paths.log).
#ifndef ORX_R1T_MINBLOCKS
#define ORX_R1T_MINBLOCKS 5
#endif
#ifndef ORX_R1T_MINBLOCKS_ALONE
#define ORX_R1T_MINBLOCKS_ALONE 4
#endif
constexpr int kThreads = ORX_R1T_THREADS;

// FLAGGED: consecutive launches on one state are ordered block by block (a block = the kThreads games of a CTA; the
// protocol of orx_pipe.cuh's flag mode with a CTA's games as the only chunk): the CTA draws a ticket for its block
// before it lets dependents launch (programmatic dependent launch), touches the state once the block's serving word
// equals the ticket, releases serving = ticket + 1 at gpu scope after the barrier that follows its last store, and
// waits for the grid dependency last, so that the completion of this grid implies that of every earlier one.
// EV: also write the tick's replication log, max_events record slots per game (orx_r1_step_events).
#ifdef ORX_PIPE_JITTER      // schedule fuzzing (tuning build, see orx_pipe.cuh): random pauses around the block hand-over
__device__ __forceinline__ void jitter(unsigned int salt)
{
    unsigned int t;
    asm volatile("mov.u32 %0, %%clock;" : "=r"(t));
    const unsigned int h = (t ^ (blockIdx.x * 2654435761u) ^ (salt * 40503u) ^ (threadIdx.x >> 5)) * 2246822519u;
    if ((h >> 29) == 0u) __nanosleep((h >> 8) & 2047u);
}
#define ORX_R1_JITTER(salt) jitter(salt)
#else
#define ORX_R1_JITTER(salt) do { } while (0)
#endif

template <bool FLAGGED, bool EV = false>
__global__ void __launch_bounds__(kThreads, FLAGGED ? ORX_R1T_MINBLOCKS : ORX_R1T_MINBLOCKS_ALONE)
k_step(const __grid_constant__ R1Params P, const uint16_t* __restrict__ moves, uint8_t* __restrict__ result, unsigned int* __restrict__ flags,
       uint2* __restrict__ events, int max_events)
{
    __shared__ unsigned int s_ticket;
    if (FLAGGED) {
        if (threadIdx.x == 0) {
            ORX_R1_JITTER(1u);
            unsigned int* f = flags + 2 * (size_t)blockIdx.x;
            const unsigned int ticket = atomicAdd(f, 1u);
            unsigned int serving;
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(serving) : "l"(f + 1) : "memory");
            asm volatile("griddepcontrol.launch_dependents;" ::"r"(ticket) : "memory");      // the ticket has been drawn
            while (serving != ticket) asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(serving) : "l"(f + 1) : "memory");
            s_ticket = ticket;
        }
        __syncthreads();
    }
    const unsigned int game = blockIdx.x * kThreads + threadIdx.x;
    ORX_R1_JITTER(2u);
    if (game < P.n) {
        Game G;
        load_game(P, game, G);
        uint32_t mvw, status;
        asm volatile("ld.global.u16 %0, [%1];" : "=r"(mvw) : "l"(moves + game));
        asm volatile("ld.global.u8 %0, [%1];" : "=r"(status) : "l"(P.status + game));
        Sink<EV> ev{EV ? events + (size_t)game * (size_t)max_events : nullptr, 0, max_events};
        if (status != ORX_RESULT_IN_PROGRESS) {
            result[game] = (uint8_t)status;
            ev.finish();                                   // a frozen game: the terminator only
        } else {
            Stream s = make_stream(P, game, G.episode);
            R1Counters cnt{};
            int res = tick<EV>(P, G, s, (int)(mvw & 255u), (int)(mvw >> 8), cnt, ev);
            ev.finish();
            result[game] = (uint8_t)res;
            if (res != ORX_RESULT_IN_PROGRESS && P.auto_reset) {
                G.episode += 1;
                s.episode = G.episode;
                setup_game(P, G, s);
                res = ORX_RESULT_IN_PROGRESS;
            }
            store_game(P, game, G, res);
        }
    }
    ORX_R1_JITTER(3u);
    if (FLAGGED) {
        __syncthreads();                       // every store of the block has been issued
        if (threadIdx.x == 0) {
            asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(flags + 2 * (size_t)blockIdx.x + 1), "r"(s_ticket + 1u) : "memory");
            asm volatile("griddepcontrol.wait;" ::: "memory");
        }
    }
}

__global__ void __launch_bounds__(kThreads, ORX_R1T_MINBLOCKS)
k_rollout(const __grid_constant__ R1Params P, int n_ticks, unsigned long long* __restrict__ stats)
{
    __shared__ unsigned int s_cnt[ORX_STAT_COUNT];
    if (threadIdx.x < ORX_STAT_COUNT) s_cnt[threadIdx.x] = 0;
    __syncthreads();
    const unsigned int game = blockIdx.x * kThreads + threadIdx.x;
    R1Counters cnt{};
    if (game < P.n) {
        int status = P.status[game];
        if (status == ORX_RESULT_IN_PROGRESS) {
            Game G;
            load_game(P, game, G);
            Stream s = make_stream(P, game, G.episode);
            for (int t = 0; t < n_ticks; ++t) {
                const uint4 b = draw_block(s, DOM_TICK, SUB_MAIN, (uint32_t)G.tick);
                Sink<false> ev{nullptr, 0, 0};
                const int res = tick<false>(P, G, s, 1 + (int)bounded(b.x, 6u), 1 + (int)bounded(b.y, 6u), cnt, ev);
                ++cnt.ticks;
                cnt.p1 += res == ORX_RESULT_PLAYER1_WIN; cnt.p2 += res == ORX_RESULT_PLAYER2_WIN; cnt.ties += res == ORX_RESULT_TIE;
                if (res != ORX_RESULT_IN_PROGRESS) {
                    if (P.auto_reset) { G.episode += 1; s.episode = G.episode; setup_game(P, G, s); }
                    else { status = res; break; }
                }
            }
            store_game(P, game, G, status);
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

// orx_r1_replay: n_ticks ticks with both players' commands queued in advance (uint16[n_ticks][n] = p1 | p2 << 8),
// state in registers in between; results uint8[n_ticks][n]. Same outcome as n_ticks k_step launches.
__global__ void __launch_bounds__(kThreads, ORX_R1T_MINBLOCKS)
k_replay(const __grid_constant__ R1Params P, const uint16_t* __restrict__ moves, uint8_t* __restrict__ results, int n_ticks)
{
    const unsigned int game = blockIdx.x * kThreads + threadIdx.x;
    if (game >= P.n) return;
    int status = P.status[game];
    Game G;
    load_game(P, game, G);
    Stream s = make_stream(P, game, G.episode);
    R1Counters cnt{};
    Sink<false> ev{nullptr, 0, 0};
    bool touched = false;
    for (int t = 0; t < n_ticks; ++t) {
        const size_t at = (size_t)t * P.n + game;
        if (status != ORX_RESULT_IN_PROGRESS) { results[at] = (uint8_t)status; continue; }       // frozen until reset
        const uint32_t mvw = moves[at];
        const int res = tick<false>(P, G, s, (int)(mvw & 255u), (int)(mvw >> 8), cnt, ev);
        results[at] = (uint8_t)res;
        touched = true;
        if (res != ORX_RESULT_IN_PROGRESS) {
            if (P.auto_reset) { G.episode += 1; s.episode = G.episode; setup_game(P, G, s); }
            else status = res;
        }
    }
    if (touched) store_game(P, game, G, status);
}

// orx_r1_bot_moves: RandomBot over the six R1 commands (words 0 / 1 of the tick's main block, the draws of the fused
// rollout), StaircaseBot towards the staircase of the player's own level (staircasebot.py:9-20).
__global__ void __launch_bounds__(256)
k_bot_moves(const __grid_constant__ R1Params P, int bot1, int bot2, uint8_t* __restrict__ moves)
{
    const unsigned int game = blockIdx.x * 256u + threadIdx.x;
    if (game >= P.n) return;
    const uint2 loc = ldg64(P.ent_loc + (size_t)game * 16);          // lanes 0, 1: x | y << 8 | ...
    const uint32_t sw = ldg32(P.lvl_stairs + game);
    Stream s = make_stream(P, game, P.episode[game]);
    const uint4 b = draw_block(s, DOM_TICK, SUB_MAIN, (uint32_t)P.tick[game]);
#pragma unroll
    for (int p = 0; p < 2; ++p) {
        const int kind = p == 0 ? bot1 : bot2;
        const uint32_t l = p == 0 ? loc.x : loc.y;
        if (kind == ORX_BOT_RANDOM) moves[2 * (size_t)game + p] = (uint8_t)(1u + bounded(p == 0 ? b.x : b.y, 6u));
        else if (kind == ORX_BOT_STAIRCASE) {
            const int ddx = (int)((sw >> (16 * p)) & 255u) - (int)(l & 255u), ddy = (int)((sw >> (16 * p + 8)) & 255u) - (int)((l >> 8) & 255u);
            moves[2 * (size_t)game + p] = (uint8_t)(abs(ddx) > abs(ddy) ? (ddx > 0 ? ORX_MOVE_RIGHT : ORX_MOVE_LEFT) : (ddy > 0 ? ORX_MOVE_DOWN : ORX_MOVE_UP));
        }
    }
}

}  // namespace r1t
