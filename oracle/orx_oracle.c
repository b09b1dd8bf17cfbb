/*
 * TEST INFRASTRUCTURE (oracle) -- not part of the product path. Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load this.
 *
 * Plain-C, one-game-at-a-time restatement of the reference's turn dynamics, written to follow
 * the reference's control flow (entity list, position lookup by scan, sequential movers) rather
 * than the CUDA kernels' data-parallel formulation. It is pinned against the live reference
 * (oracle/ref_harness.py) through tests/golden/ and tests/test_oracle_vs_reference.py.
 *
 * Reference files restated (paths under the reference root):
 *   optimax_rogue/logic/updater.py:76-162   Updater.update            -> oro_tick()
 *   optimax_rogue/logic/updater.py:180-243  Updater.handle_move       -> handle_move()
 *   optimax_rogue/logic/updater.py:259-296  Updater.handle_descend    -> handle_descend()
 *   optimax_rogue/logic/updater.py:298-338  Updater.handle_combat     -> handle_combat()
 *   optimax_rogue/logic/updater.py:245-257  Updater.should_despawn    -> level_exists()
 *   optimax_rogue/logic/updater.py:340-351  calculate_pos             -> calculate_pos()
 *   optimax_rogue/game/world.py:41-46       Dungeon.is_blocked        -> is_blocked()
 *   optimax_rogue/game/world.py:57-66       Dungeon.get_random_unblocked -> kth_ground()
 *   optimax_rogue/logic/worldgen.py:33-43   EmptyDungeonGenerator.spawn_dungeon -> level_stairs()
 *   optimax_rogue/logic/worldgen.py:77-87,124-135  setup_game         -> oro_setup_game()
 *   optimax_rogue/game/attribles.py:21-43   derived stats (no modifiers exist) -> cfg->damage/armor
 *   optimax_rogue_bots/randombot.py:20-21, staircasebot.py:9-20      -> bot_move()
 * Random draws follow the schedule in oracle/philox.py.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../include/orx.h"
#ifdef _OPENMP
#include <omp.h>
#endif

/* threads used by oro_rollout (the CPU baseline); torchrun exports OMP_NUM_THREADS=1 */
int oro_set_threads(int n)
{
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
    return omp_get_max_threads();
#else
    (void)n;
    return 1;
#endif
}

/* ------------------------------------------------------------------ Philox4x32-10 */
static void philox(uint32_t c[4], uint32_t k0, uint32_t k1)
{
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c[0];
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c[2];
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
}

void oro_philox(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4])
{
    memcpy(out, ctr, 16);
    philox(out, key[0], key[1]);
}

enum { DOM_TICK = 0, DOM_LEVEL = 1, DOM_RESET = 2 };
enum { SUB_MAIN = 0, SUB_NPC = 1, SUB_DESCEND = 64, MAX_TRIES = 256 };

typedef struct { uint64_t seed, gid; uint32_t episode; } Stream;

static void draw_block(const Stream* s, int domain, int sub, uint32_t index, uint32_t w[4])
{
    w[0] = (uint32_t)s->gid;
    w[1] = ((uint32_t)(s->gid >> 32) & 0x3FFFFFu) | ((uint32_t)(sub & 0xFF) << 22) | ((uint32_t)domain << 30);
    w[2] = s->episode;
    w[3] = index;
    philox(w, (uint32_t)s->seed, (uint32_t)(s->seed >> 32));
}

static uint32_t bounded(uint32_t w, uint32_t n) { return (uint32_t)(((uint64_t)w * n) >> 32); }

/* bounded value of draw q of a sequence starting at block sub_base */
static uint32_t seq_bounded(const Stream* s, int domain, int sub_base, uint32_t index, int q, uint32_t n)
{
    uint32_t w[4];
    if (q < MAX_TRIES) {
        draw_block(s, domain, sub_base + (q >> 2), index, w);
        return bounded(w[q & 3], n);
    }
    draw_block(s, domain, sub_base + ((MAX_TRIES - 1) >> 2), index, w);
    return (bounded(w[(MAX_TRIES - 1) & 3], n) + (uint32_t)(q - (MAX_TRIES - 1))) % n;
}

/* ------------------------------------------------------------------ one game, unpacked */
#define MAX_ENT (2 + ORX_MAX_NPC)
typedef struct { int iden, depth, x, y, health, present; } Entity;

typedef struct {
    const OrxConfig* cfg;
    const uint8_t* tiles;       /* host copy of the fixed map (ORX_DGEN_FIXED) */
    const uint16_t* ground;
    Stream rng;
    int tick;
    int status;
    Entity ent[MAX_ENT];        /* entities list: players then NPC slots */
    int n_ent;
    int stairs[2][2];           /* staircase of each player's level */
    int flat[2][3];             /* per player: flat damage / armor / max-health bonus of its modifiers (OrxState.flat) */
    OrxEvent* ev;               /* nullable */
    int n_ev, max_ev;
    unsigned long long* stats;  /* nullable */
} Game;

static void emit(Game* g, int kind, int iden, int a, int b, int depth)
{
    if (g->stats) g->stats[ORX_STAT_EVENTS]++;
    if (!g->ev || g->n_ev >= g->max_ev) return;
    OrxEvent* e = &g->ev[g->n_ev++];
    e->kind = (uint8_t)kind; e->iden = (uint8_t)iden; e->a = (uint8_t)a; e->b = (uint8_t)b; e->depth = depth;
}

/* updater.py:340-351 */
static void calculate_pos(int x, int y, int move, int* nx, int* ny)
{
    *nx = x; *ny = y;
    if (move == ORX_MOVE_UP) *ny = y - 1;
    else if (move == ORX_MOVE_DOWN) *ny = y + 1;
    else if (move == ORX_MOVE_RIGHT) *nx = x + 1;
    else if (move == ORX_MOVE_LEFT) *nx = x - 1;
}

/* worldgen.py:39-40: rx = randint(1, W-2), ry = randint(1, H-2), numpy high-exclusive */
static void level_stairs(const Game* g, int depth, int* sx, int* sy)
{
    const OrxConfig* c = g->cfg;
    if (c->dgen_kind == ORX_DGEN_FIXED) { *sx = c->fixed_stairs[0]; *sy = c->fixed_stairs[1]; return; }
    uint32_t w[4];
    draw_block(&g->rng, DOM_LEVEL, 0, (uint32_t)depth, w);
    *sx = 1 + (int)bounded(w[0], (uint32_t)(c->width - 3));
    *sy = 1 + (int)bounded(w[1], (uint32_t)(c->height - 3));
}

/* tile code of a level whose staircase is (sx, sy) */
static int tile_at(const Game* g, int sx, int sy, int x, int y)
{
    const OrxConfig* c = g->cfg;
    if (c->dgen_kind == ORX_DGEN_FIXED) return g->tiles[x * c->height + y];
    if (x == 0 || y == 0 || x == c->width - 1 || y == c->height - 1) return ORX_TILE_WALL;
    if (x == sx && y == sy) return ORX_TILE_STAIRCASE_DOWN;
    return ORX_TILE_GROUND;
}

/* world.py:41-46 */
static int is_blocked(const Game* g, int sx, int sy, int x, int y)
{
    if (x < 0 || x >= g->cfg->width || y < 0 || y >= g->cfg->height) return 1;
    return tile_at(g, sx, sy, x, y) == ORX_TILE_WALL;
}

static int n_ground(const Game* g)
{
    const OrxConfig* c = g->cfg;
    if (c->dgen_kind == ORX_DGEN_FIXED) return c->fixed_n_ground;
    return (c->width - 2) * (c->height - 2) - 1;
}

/* world.py:59-66: the k-th Ground tile in x-major flat order; done by scanning, as numpy does */
static void kth_ground(const Game* g, int sx, int sy, int k, int* ox, int* oy)
{
    const OrxConfig* c = g->cfg;
    if (c->dgen_kind == ORX_DGEN_FIXED) {
        int flat = g->ground[k];
        *ox = flat / c->height; *oy = flat - *ox * c->height;
        return;
    }
    int seen = 0;
    for (int x = 0; x < c->width; ++x)
        for (int y = 0; y < c->height; ++y)
            if (tile_at(g, sx, sy, x, y) == ORX_TILE_GROUND) {
                if (seen == k) { *ox = x; *oy = y; return; }
                ++seen;
            }
    *ox = *oy = -1; /* unreachable */
}

/* state.py:33 pos_lookup, by scan over present entities */
static int entity_at(const Game* g, int depth, int x, int y)
{
    for (int i = 0; i < g->n_ent; ++i)
        if (g->ent[i].present && g->ent[i].depth == depth && g->ent[i].x == x && g->ent[i].y == y)
            return i;
    return -1;
}

/* Does the level `depth` exist in World.dungeons when player p is about to descend into it?
 * Follows from updater.py:245-257,274-280,295-296 given that players only ever descend one
 * level at a time from start_depth: under Unreachable a level is alive while some player is at
 * or above it, under Unused only while a player stands on it. */
static int level_exists(const Game* g, int p, int depth)
{
    const Entity* o = &g->ent[1 - p];
    int o_start = g->cfg->start_kind == ORX_START_SEPARATED ? g->cfg->start_depth[1 - p] : g->cfg->start_depth[0];
    if (g->cfg->despawn_strat == ORX_DESPAWN_UNUSED) return o->depth == depth;
    return o_start <= depth && depth <= o->depth;
}

/* updater.py:298-338: damage = attacker.damage.value - attacker.armor.value (:313), the attribles being the base
 * stat plus the flat bonuses of the entity's modifiers (attribles.py:29-43; OrxState.flat, zero when absent).
 * Modifier event hooks are not modelled (no concrete modifier exists upstream). */
static void handle_combat(Game* g, int att, int def, int flag)
{
    int p = att; /* only players ever attack: NPC moves are always Stay (updater.py:165-178) */
    int og_dmg = (g->cfg->damage[p] + g->flat[p][0]) - (g->cfg->armor[p] + g->flat[p][1]);
    if (og_dmg > 0) {
        g->ent[def].health -= og_dmg;
        if (g->stats) g->stats[ORX_STAT_HITS]++;
    }
    emit(g, ORX_EV_COMBAT, g->ent[att].iden, g->ent[def].iden, flag, og_dmg);
}

/* updater.py:259-296 (player branch; an NPC never moves so :263-270 is unreachable) */
static void handle_descend(Game* g, int p)
{
    Entity* e = &g->ent[p];
    int old_depth = e->depth, new_depth = old_depth + 1;
    int sx, sy;
    level_stairs(g, new_depth, &sx, &sy);
    if (!level_exists(g, p, new_depth))
        emit(g, ORX_EV_DUNGEON, 0, sx, sy, new_depth);
    int ng = n_ground(g), x, y, q = 0;
    do {
        int k = (int)seq_bounded(&g->rng, DOM_TICK, SUB_DESCEND + 64 * p, (uint32_t)g->tick, q++, (uint32_t)ng);
        kth_ground(g, sx, sy, k, &x, &y);
    } while (entity_at(g, new_depth, x, y) >= 0);
    emit(g, ORX_EV_DESCEND, e->iden, x, y, new_depth);
    e->depth = new_depth; e->x = x; e->y = y;
    g->stairs[p][0] = sx; g->stairs[p][1] = sy;
    if (g->stats) g->stats[ORX_STAT_DESCENTS]++;
}

/* updater.py:180-243. order[] lists entity indices by initiative; ind is the mover's slot */
static void handle_move(Game* g, int ind, const int* order, const int* moves /* by entity */, int n_upd)
{
    int me = order[ind];
    Entity* e = &g->ent[me];
    if (moves[me] == ORX_MOVE_STAY) return;
    int nx, ny;
    calculate_pos(e->x, e->y, moves[me], &nx, &ny);
    int occ = entity_at(g, e->depth, nx, ny);
    if (occ < 0) {
        int tile = tile_at(g, g->stairs[me][0], g->stairs[me][1], nx, ny);
        if (tile == ORX_TILE_STAIRCASE_DOWN) { handle_descend(g, me); return; }
        emit(g, ORX_EV_MOVE, e->iden, nx, ny, e->depth);
        e->x = nx; e->y = ny;
        return;
    }
    int occ_ind = -1;
    for (int i = 0; i < n_upd; ++i) if (order[i] == occ) occ_ind = i;
    if (moves[occ] == ORX_MOVE_STAY) { handle_combat(g, me, occ, ORX_FLAG_BLOCK); return; }
    int ox, oy;
    calculate_pos(g->ent[occ].x, g->ent[occ].y, moves[occ], &ox, &oy);
    if (ox == nx && oy == ny) { handle_combat(g, me, occ, ORX_FLAG_PARRY); return; }
    if (occ_ind < ind) { handle_combat(g, me, occ, ORX_FLAG_AMBUSH); return; }
    handle_combat(g, me, occ, ORX_FLAG_FLEE);
}

/* updater.py:76-162 */
static int oro_tick(Game* g, int m1, int m2)
{
    const OrxConfig* c = g->cfg;
    int moves[MAX_ENT];
    int in_moves[2] = { m1, m2 };
    /* commands outside the Move enum are not representable in the reference; the ABI maps them to Stay */
    for (int p = 0; p < 2; ++p) {
        int m = in_moves[p];
        if (m < ORX_MOVE_UP || m > ORX_MOVE_STAY) m = ORX_MOVE_STAY;
        int nx, ny;
        calculate_pos(g->ent[p].x, g->ent[p].y, m, &nx, &ny);
        if (is_blocked(g, g->stairs[p][0], g->stairs[p][1], nx, ny)) m = ORX_MOVE_STAY; /* :90-98 */
        moves[p] = m;
    }
    /* :101-114 random.shuffle([p1, p2]): j = randbelow(2); j == 0 swaps */
    uint32_t w[4];
    draw_block(&g->rng, DOM_TICK, SUB_MAIN, (uint32_t)g->tick, w);
    int order[MAX_ENT], n_upd = 2;
    if (bounded(w[2], 2) == 0) { order[0] = 1; order[1] = 0; } else { order[0] = 0; order[1] = 1; }
    /* :116-128 NPCs: move Stay, shuffled among themselves, appended after the players */
    int npcs[ORX_MAX_NPC], n_npcs = 0;
    for (int i = 2; i < g->n_ent; ++i) if (g->ent[i].present) { npcs[n_npcs++] = i; moves[i] = ORX_MOVE_STAY; }
    for (int i = n_npcs - 1, q = 0; i >= 1; --i, ++q) {
        int j = (int)seq_bounded(&g->rng, DOM_TICK, SUB_NPC, (uint32_t)g->tick, q, (uint32_t)(i + 1));
        int t = npcs[i]; npcs[i] = npcs[j]; npcs[j] = t;
    }
    for (int i = 0; i < n_npcs; ++i) order[n_upd++] = npcs[i];

    for (int ind = 0; ind < n_upd; ++ind) handle_move(g, ind, order, moves, n_upd); /* :133-134 */

    for (int i = g->n_ent - 1; i >= 2; --i)                                     /* :137-145 */
        if (g->ent[i].present && g->ent[i].health <= 0) {
            emit(g, ORX_EV_DEATH, g->ent[i].iden, 0, 0, 0);
            g->ent[i].present = 0;
        }
    g->tick += 1;                                                                /* :148 */
    if (g->ent[0].health <= 0) return g->ent[1].health <= 0 ? ORX_RESULT_TIE : ORX_RESULT_PLAYER2_WIN;
    if (g->ent[1].health <= 0) return ORX_RESULT_PLAYER1_WIN;
    if (c->max_ticks && g->tick >= c->max_ticks) return ORX_RESULT_TIE;
    return ORX_RESULT_IN_PROGRESS;
}

/* worldgen.py:77-87 / :124-135 */
static void oro_setup_game(Game* g)
{
    const OrxConfig* c = g->cfg;
    int ng = n_ground(g);
    for (int i = 0; i < MAX_ENT; ++i) g->ent[i].present = 0;
    for (int p = 0; p < 2; ++p) {
        g->ent[p].iden = p + 1; g->ent[p].present = 1; g->ent[p].health = c->hp[p];
    }
    if (c->start_kind == ORX_START_TOGETHER) {
        int d = c->start_depth[0];
        level_stairs(g, d, &g->stairs[0][0], &g->stairs[0][1]);
        g->stairs[1][0] = g->stairs[0][0]; g->stairs[1][1] = g->stairs[0][1];
        int q = 0, x, y;
        kth_ground(g, g->stairs[0][0], g->stairs[0][1],
                   (int)seq_bounded(&g->rng, DOM_RESET, 0, 0, q++, (uint32_t)ng), &x, &y);
        g->ent[0].x = x; g->ent[0].y = y; g->ent[0].depth = d;
        do {
            kth_ground(g, g->stairs[0][0], g->stairs[0][1],
                       (int)seq_bounded(&g->rng, DOM_RESET, 0, 0, q++, (uint32_t)ng), &x, &y);
        } while (x == g->ent[0].x && y == g->ent[0].y);
        g->ent[1].x = x; g->ent[1].y = y; g->ent[1].depth = d;
    } else {
        for (int p = 0; p < 2; ++p) {
            int d = c->start_depth[p], x, y;
            level_stairs(g, d, &g->stairs[p][0], &g->stairs[p][1]);
            kth_ground(g, g->stairs[p][0], g->stairs[p][1],
                       (int)seq_bounded(&g->rng, DOM_RESET, 0, 0, p, (uint32_t)ng), &x, &y);
            g->ent[p].x = x; g->ent[p].y = y; g->ent[p].depth = d;
        }
    }
    g->n_ent = 2 + c->n_npc;
    g->tick = 1;
    g->status = ORX_RESULT_IN_PROGRESS;
}

/* randombot.py:20-21 / staircasebot.py:9-20 */
static int bot_move(const Game* g, int p, int kind, const uint32_t w[4])
{
    if (kind == ORX_BOT_NONE) return ORX_MOVE_STAY;
    if (kind == ORX_BOT_RANDOM) return 1 + (int)bounded(w[p], 5);
    int dx = g->stairs[p][0] - g->ent[p].x, dy = g->stairs[p][1] - g->ent[p].y;
    if (abs(dx) > abs(dy)) return dx > 0 ? ORX_MOVE_RIGHT : ORX_MOVE_LEFT;
    return dy > 0 ? ORX_MOVE_DOWN : ORX_MOVE_UP;
}

/* ------------------------------------------------------------------ SoA load / store */
static void load_game(Game* g, const OrxConfig* cfg, const OrxState* st, int64_t i, uint64_t gid)
{
    memset(g, 0, sizeof(*g));
    g->cfg = cfg; g->tiles = cfg->fixed_tiles; g->ground = cfg->fixed_ground;
    g->rng.seed = cfg->seed; g->rng.gid = gid; g->rng.episode = st->episode[i];
    g->tick = st->tick[i]; g->status = st->status[i];
    for (int p = 0; p < 2; ++p) {
        Entity* e = &g->ent[p];
        e->iden = p + 1; e->present = 1;
        e->x = st->pos[4 * i + 2 * p]; e->y = st->pos[4 * i + 2 * p + 1];
        e->depth = st->depth[2 * i + p]; e->health = st->hp[2 * i + p];
        g->stairs[p][0] = st->stairs[4 * i + 2 * p]; g->stairs[p][1] = st->stairs[4 * i + 2 * p + 1];
    }
    if (st->flat != NULL)
        for (int p = 0; p < 2; ++p)
            for (int k = 0; k < 3; ++k) g->flat[p][k] = st->flat[6 * i + 3 * p + k];
    g->n_ent = 2 + cfg->n_npc;
    for (int k = 0; k < cfg->n_npc; ++k) {
        Entity* e = &g->ent[2 + k];
        int64_t j = i * cfg->n_npc + k;
        e->iden = 3 + k; e->depth = st->npc_depth[j]; e->present = e->depth >= 0;
        e->x = st->npc_pos[2 * j]; e->y = st->npc_pos[2 * j + 1]; e->health = st->npc_hp[j];
    }
}

/* health is a Python int in the reference (entities.py:38) and an int16 plane here: a value below the plane's
 * range (only reachable with hits near 32767) is stored as the lowest one, so that a dead entity stays dead */
static int16_t health_plane(int h) { return (int16_t)(h < -32768 ? -32768 : h > 32767 ? 32767 : h); }

static void store_game(const Game* g, const OrxState* st, int64_t i)
{
    st->tick[i] = g->tick; st->status[i] = (uint8_t)g->status; st->episode[i] = g->rng.episode;
    for (int p = 0; p < 2; ++p) {
        const Entity* e = &g->ent[p];
        st->pos[4 * i + 2 * p] = (uint8_t)e->x; st->pos[4 * i + 2 * p + 1] = (uint8_t)e->y;
        st->depth[2 * i + p] = e->depth; st->hp[2 * i + p] = health_plane(e->health);
        st->stairs[4 * i + 2 * p] = (uint8_t)g->stairs[p][0]; st->stairs[4 * i + 2 * p + 1] = (uint8_t)g->stairs[p][1];
    }
    for (int k = 0; k < g->cfg->n_npc; ++k) {
        const Entity* e = &g->ent[2 + k];
        int64_t j = i * g->cfg->n_npc + k;
        st->npc_depth[j] = e->present ? e->depth : -1;
        st->npc_pos[2 * j] = (uint8_t)e->x; st->npc_pos[2 * j + 1] = (uint8_t)e->y;
        st->npc_hp[j] = health_plane(e->health);
    }
}

static void finish_tick(Game* g, int res, uint8_t* result_out)
{
    *result_out = (uint8_t)res;
    if (g->stats && res != ORX_RESULT_IN_PROGRESS) {
        g->stats[res == ORX_RESULT_PLAYER1_WIN ? ORX_STAT_P1_WINS : res == ORX_RESULT_PLAYER2_WIN ? ORX_STAT_P2_WINS : ORX_STAT_TIES]++;
    }
    if (res != ORX_RESULT_IN_PROGRESS && g->cfg->auto_reset) {
        g->rng.episode += 1;
        oro_setup_game(g);
    } else {
        g->status = res;
    }
}

/* ------------------------------------------------------------------ batched entry points (host pointers) */
int oro_reset(const OrxConfig* cfg, const OrxState* st, const uint8_t* mask, int bump_episode,
              int64_t n, uint64_t game_id_base)
{
    for (int64_t i = 0; i < n; ++i) {
        if (mask && !mask[i]) continue;
        Game g;
        load_game(&g, cfg, st, i, game_id_base + (uint64_t)i);
        if (bump_episode) g.rng.episode += 1;
        oro_setup_game(&g);
        store_game(&g, st, i);
    }
    return 0;
}

int oro_step(const OrxConfig* cfg, const OrxState* st, const uint8_t* moves, uint8_t* result,
             OrxEvent* events, int64_t n, uint64_t game_id_base)
{
    int max_ev = ORX_MAX_EVENTS_BASE + cfg->n_npc;
    for (int64_t i = 0; i < n; ++i) {
        Game g;
        load_game(&g, cfg, st, i, game_id_base + (uint64_t)i);
        if (events) {
            g.ev = events + i * max_ev; g.max_ev = max_ev;
            memset(g.ev, 0, sizeof(OrxEvent) * (size_t)max_ev);
        }
        if (g.status != ORX_RESULT_IN_PROGRESS) { result[i] = (uint8_t)g.status; continue; }
        int res = oro_tick(&g, moves[2 * i], moves[2 * i + 1]);
        finish_tick(&g, res, &result[i]);
        store_game(&g, st, i);
    }
    return 0;
}

int oro_bot_moves(const OrxConfig* cfg, const OrxState* st, int bot_p1, int bot_p2,
                  uint8_t* moves, int64_t n, uint64_t game_id_base)
{
    int kinds[2] = { bot_p1, bot_p2 };
    for (int64_t i = 0; i < n; ++i) {
        Game g;
        load_game(&g, cfg, st, i, game_id_base + (uint64_t)i);
        uint32_t w[4];
        draw_block(&g.rng, DOM_TICK, SUB_MAIN, (uint32_t)g.tick, w);
        for (int p = 0; p < 2; ++p)
            if (kinds[p] != ORX_BOT_NONE) moves[2 * i + p] = (uint8_t)bot_move(&g, p, kinds[p], w);
    }
    return 0;
}

/* n_ticks ticks with both bots inlined; OpenMP over games when built with -fopenmp
 * (this is the all-host-cores CPU baseline of bench.py). stats accumulates. */
int oro_rollout(const OrxConfig* cfg, const OrxState* st, int bot_p1, int bot_p2, int n_ticks,
                unsigned long long* stats, int64_t n, uint64_t game_id_base)
{
    int kinds[2] = { bot_p1, bot_p2 };
    unsigned long long total[ORX_STAT_COUNT] = { 0 };
#pragma omp parallel
    {
        unsigned long long local[ORX_STAT_COUNT] = { 0 };
#pragma omp for schedule(static)
        for (int64_t i = 0; i < n; ++i) {
            Game g;
            load_game(&g, cfg, st, i, game_id_base + (uint64_t)i);
            g.stats = local;
            for (int t = 0; t < n_ticks; ++t) {
                if (g.status != ORX_RESULT_IN_PROGRESS) break;
                uint32_t w[4];
                draw_block(&g.rng, DOM_TICK, SUB_MAIN, (uint32_t)g.tick, w);
                int m1 = bot_move(&g, 0, kinds[0], w), m2 = bot_move(&g, 1, kinds[1], w);
                uint8_t r;
                int res = oro_tick(&g, m1, m2);
                local[ORX_STAT_TICKS]++;
                finish_tick(&g, res, &r);
            }
            store_game(&g, st, i);
        }
#pragma omp critical
        for (int k = 0; k < ORX_STAT_COUNT; ++k) total[k] += local[k];
    }
    if (stats) for (int k = 0; k < ORX_STAT_COUNT; ++k) stats[k] += total[k];
    return 0;
}
