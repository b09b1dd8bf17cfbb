/*
 * TEST INFRASTRUCTURE (oracle) -- not part of the product path.
 *
 * Plain-C, one-game-at-a-time restatement of ruleset R1 (docs/RULESET_R1.md): the README-only
 * rules of optimax_rogue (readme.md:44-48,69-74). PARITY UNPINNED: the reference has no code for
 * these rules, so this oracle pins the CUDA kernel to the written spec, not to the reference.
 * Hooks it mirrors: enemies are the entities after the players (logic/updater.py:116-128), their
 * policy is decide_npc_move (:165-178), dead ones leave like dead NPCs (:137-145).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../include/orx.h"
#ifdef _OPENMP
#include <omp.h>
#endif

/* ---- Philox4x32-10 + schedule (same stream as R0, see oracle/philox.py) */
static void philox(uint32_t c[4], uint32_t k0, uint32_t k1)
{
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c[0], p1 = (uint64_t)0xCD9E8D57u * c[2];
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k0, n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k1, n3 = (uint32_t)p0;
        c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
}
enum { DOM_TICK = 0, DOM_LEVEL = 1, DOM_RESET = 2 };
enum { SUB_MAIN = 0, SUB_SPAWN = 8, SUB_DROP = 10, SUB_SPAWN_TRY = 16, SUB_DESCEND = 64, MAX_TRIES = 64 };
typedef struct { uint64_t seed, gid; uint32_t episode; } Stream;
static void draw_block(const Stream* s, int domain, int sub, uint32_t index, uint32_t w[4])
{
    w[0] = (uint32_t)s->gid;
    w[1] = ((uint32_t)(s->gid >> 32) & 0x3FFFFFu) | ((uint32_t)(sub & 0xFF) << 22) | ((uint32_t)domain << 30);
    w[2] = s->episode; w[3] = index;
    philox(w, (uint32_t)s->seed, (uint32_t)(s->seed >> 32));
}
static uint32_t bounded(uint32_t w, uint32_t n) { return (uint32_t)(((uint64_t)w * n) >> 32); }

/* ---- one game */
#define NM 10   /* movers: 2 players + 8 enemies */
typedef struct { int x, y, depth, alive, hp, aux; } Ent;   /* aux: mana (players) / kind (items) */
typedef struct { int max_hp, max_mana, xp, level, n_items, cd, damage, armor; } Pl;
typedef struct {
    const OrxR1Config* cfg;
    Stream rng;
    Ent e[ORX_R1_LANES];
    Pl p[2];
    int sx[2], sy[2];
    uint32_t key[2];
    int sep, tick, status;
    unsigned long long* stats;
    OrxEvent* ev;            /* this game's record slots for the tick (docs/RULESET_R1.md "Replication log"), or NULL */
    int n_ev, cap_ev;
} G;

/* one record; beyond the capacity records are dropped (ORX_R1_MAX_EVENTS slots never overflow) */
static void emit(G* g, int kind, int iden, int a, int b, int value)
{
    if (!g->ev) return;
    if (g->n_ev < g->cap_ev) {
        OrxEvent* e = &g->ev[g->n_ev];
        e->kind = (uint8_t)kind; e->iden = (uint8_t)iden; e->a = (uint8_t)a; e->b = (uint8_t)b; e->depth = value;
    }
    g->n_ev++;
}

static uint32_t mix(uint32_t x, uint32_t y, uint32_t key)
{
    uint32_t h = x * 0x9E3779B1u ^ y * 0x85EBCA77u ^ key;
    h ^= h >> 15; h *= 0x2C1B3C6Du; h ^= h >> 12; h *= 0x297A2D39u; h ^= h >> 15;
    return h;
}
static int is_wall(const G* g, uint32_t key, int sx, int sy, int x, int y)
{
    const OrxR1Config* c = g->cfg;
    if (x <= 0 || y <= 0 || x >= c->width - 1 || y >= c->height - 1) return 1;
    if (x == sx && y == sy) return 0;
    return (int)(mix((uint32_t)x, (uint32_t)y, key) & 255u) < c->wall_density;
}
static void level_init(const G* g, int depth, int* sx, int* sy, uint32_t* key)
{
    uint32_t w[4];
    draw_block(&g->rng, DOM_LEVEL, 0, (uint32_t)depth, w);
    *sx = 1 + (int)bounded(w[0], (uint32_t)(g->cfg->width - 3));
    *sy = 1 + (int)bounded(w[1], (uint32_t)(g->cfg->height - 3));
    *key = w[2];
}
static int occupant(const G* g, int depth, int x, int y)
{
    for (int m = 0; m < NM; ++m)
        if (g->e[m].alive && g->e[m].depth == depth && g->e[m].x == x && g->e[m].y == y) return m;
    return -1;
}
static int tile_ok(const G* g, int depth, uint32_t key, int sx, int sy, int x, int y)
{
    return !is_wall(g, key, sx, sy, x, y) && !(x == sx && y == sy) && occupant(g, depth, x, y) < 0;
}
/* random free tile: two words per try, then the first acceptable tile in x-major order */
static void free_tile(const G* g, int domain, int sub_base, uint32_t index, int depth, uint32_t key,
                      int sx, int sy, int* ox, int* oy)
{
    const OrxR1Config* c = g->cfg;
    for (int r = 0; r < MAX_TRIES; ++r) {
        uint32_t w[4];
        draw_block(&g->rng, domain, sub_base + (r >> 1), index, w);
        int x = 1 + (int)bounded(w[2 * (r & 1)], (uint32_t)(c->width - 2));
        int y = 1 + (int)bounded(w[2 * (r & 1) + 1], (uint32_t)(c->height - 2));
        if (tile_ok(g, depth, key, sx, sy, x, y)) { *ox = x; *oy = y; return; }
    }
    for (int x = 1; x < c->width - 1; ++x)
        for (int y = 1; y < c->height - 1; ++y)
            if (tile_ok(g, depth, key, sx, sy, x, y)) { *ox = x; *oy = y; return; }
    *ox = 1; *oy = 1;
}

static void setup_game(G* g)
{
    memset(g->e, 0, sizeof(g->e));
    level_init(g, 0, &g->sx[0], &g->sy[0], &g->key[0]);
    g->sx[1] = g->sx[0]; g->sy[1] = g->sy[0]; g->key[1] = g->key[0];
    for (int p = 0; p < 2; ++p) {
        int x, y;
        free_tile(g, DOM_RESET, 32 * p, 0, 0, g->key[0], g->sx[0], g->sy[0], &x, &y);
        Ent* e = &g->e[p];
        e->x = x; e->y = y; e->depth = 0; e->alive = 1; e->hp = 10; e->aux = 9;
        Pl* q = &g->p[p];
        q->max_hp = 10; q->max_mana = 9; q->xp = 0; q->level = 1; q->n_items = 0; q->cd = 0; q->damage = 2; q->armor = 1;
    }
    g->sep = 0; g->tick = 1; g->status = ORX_RESULT_IN_PROGRESS;
}

static int isign(int v) { return v > 0 ? 1 : -1; }
static int imin(int a, int b) { return a < b ? a : b; }
static int imax(int a, int b) { return a > b ? a : b; }

static int attack_amount(const G* g, int m, int d)
{
    int dmg = m < 2 ? g->p[m].damage + imin(g->e[m].aux, g->p[m].max_mana / 3) : 2 + g->e[m].depth / 4;
    int arm = d < 2 ? g->p[d].armor : 0;
    return imax(0, dmg - arm);
}

static int r1_tick(G* g, int c1, int c2)
{
    const OrxR1Config* cfg = g->cfg;
    int cmd[2] = { c1, c2 };
    int dx[NM] = { 0 }, dy[NM] = { 0 };
    int cd_pre[2] = { g->p[0].cd, g->p[1].cd };
    /* 1. heal, 2. intents */
    for (int p = 0; p < 2; ++p) {
        Ent* e = &g->e[p];
        int c = cmd[p];
        if (c == ORX_MOVE_HEAL) {
            int h = imin(e->aux, g->p[p].max_mana / 3), before = e->hp;
            e->hp = imin(g->p[p].max_hp, e->hp + h);
            e->aux -= h;
            if (h > 0) emit(g, ORX_EV_HEALTH, p + 1, p + 1, ORX_R1_HEALTH_HEAL, e->hp - before);
            c = ORX_MOVE_STAY;
        }
        int ddx = (c == ORX_MOVE_RIGHT) - (c == ORX_MOVE_LEFT), ddy = (c == ORX_MOVE_DOWN) - (c == ORX_MOVE_UP);
        if ((ddx || ddy) && !is_wall(g, g->key[p], g->sx[p], g->sy[p], e->x + ddx, e->y + ddy)) { dx[p] = ddx; dy[p] = ddy; }
    }
    for (int m = 2; m < NM; ++m) {
        const Ent* e = &g->e[m];
        if (!e->alive) continue;
        int best = -1, bestd = 1 << 30;
        for (int p = 0; p < 2; ++p)
            if (g->e[p].depth == e->depth) {
                int d = abs(g->e[p].x - e->x) + abs(g->e[p].y - e->y);
                if (d < bestd) { bestd = d; best = p; }
            }
        if (best < 0) continue;
        int ex = g->e[best].x - e->x, ey = g->e[best].y - e->y;
        if (imax(abs(ex), abs(ey)) > 6) continue;
        int ddx = 0, ddy = 0;
        if (abs(ex) > abs(ey)) ddx = isign(ex); else ddy = isign(ey);
        int lp = g->e[0].depth == e->depth ? 0 : 1;
        int tx = e->x + ddx, ty = e->y + ddy;
        if (is_wall(g, g->key[lp], g->sx[lp], g->sy[lp], tx, ty) || (tx == g->sx[lp] && ty == g->sy[lp])) continue;
        dx[m] = ddx; dy[m] = ddy;
    }
    /* 3. cooldown conversion */
    for (int p = 0; p < 2; ++p)
        if (cd_pre[p] > 0 && (dx[p] || dy[p])) {
            int o = occupant(g, g->e[p].depth, g->e[p].x + dx[p], g->e[p].y + dy[p]);
            if (o >= 0) { dx[p] = 0; dy[p] = 0; }
        }
    /* 4. attacks, from the start-of-tick state */
    int taken[NM] = { 0 }, credit[NM][2], moves[NM] = { 0 }, newcd[2] = { 0, 0 }, spend[2] = { 0, 0 };
    memset(credit, 0, sizeof(credit));
    for (int m = 0; m < NM; ++m) {
        const Ent* e = &g->e[m];
        if (!e->alive || !(dx[m] || dy[m])) continue;
        int tx = e->x + dx[m], ty = e->y + dy[m];
        int o = occupant(g, e->depth, tx, ty);
        if (o >= 0) {
            if (m >= 2 && o >= 2) continue;
            int amount = attack_amount(g, m, o);
            if (!(dx[o] || dy[o])) {
                if (o < 2 && cd_pre[o] == 0) { if (m < 2) { newcd[m] = imax(newcd[m], 1); spend[m] = 1; } emit(g, ORX_EV_COMBAT, m + 1, o + 1, ORX_R1_HIT_NEGATED, 0); }
                else { taken[o] += amount; if (m < 2) { spend[m] = 1; if (o >= 2 && amount > 0) credit[o][m] = 1; } emit(g, ORX_EV_COMBAT, m + 1, o + 1, ORX_R1_HIT_FULL, amount); }
            } else if (g->e[o].x + dx[o] == e->x && g->e[o].y + dy[o] == e->y) {
                taken[o] += amount / 2;
                if (m < 2) { newcd[m] = 3; spend[m] = 1; if (o >= 2 && amount / 2 > 0) credit[o][m] = 1; }
                emit(g, ORX_EV_COMBAT, m + 1, o + 1, ORX_R1_HIT_HALF, amount / 2);
            }
            continue;
        }
        int victim = -1, contested = 0;
        for (int c = 0; c < NM; ++c) {
            const Ent* f = &g->e[c];
            if (c == m || !f->alive || !(dx[c] || dy[c]) || f->depth != e->depth) continue;
            if (f->x + dx[c] == tx && f->y + dy[c] == ty) {
                contested = 1;
                if (victim < 0 && (m < 2 || c < 2)) victim = c;
            }
        }
        if (!contested) { moves[m] = 1; continue; }
        if (victim >= 0 && !(m < 2 && cd_pre[m] > 0)) {
            int amount = attack_amount(g, m, victim);
            taken[victim] += amount;
            if (m < 2) { spend[m] = 1; if (victim >= 2 && amount > 0) credit[victim][m] = 1; }
            emit(g, ORX_EV_COMBAT, m + 1, victim + 1, ORX_R1_HIT_CONTEST, amount);
        }
    }
    /* 5. apply */
    for (int p = 0; p < 2; ++p) if (spend[p]) g->e[p].aux -= imin(g->e[p].aux, g->p[p].max_mana / 3);
    for (int m = 0; m < NM; ++m) if (g->e[m].alive) {
        g->e[m].hp -= taken[m];
        if (taken[m] > 0 && g->stats) g->stats[ORX_STAT_HITS]++;
        if (moves[m]) { g->e[m].x += dx[m]; g->e[m].y += dy[m]; emit(g, ORX_EV_MOVE, m + 1, g->e[m].x, g->e[m].y, g->e[m].depth); }
    }
    /* pickups of both players, then descents of both players: neither reads what the other writes, and the
     * replication log lists them in this order */
    int descends[2] = { 0, 0 };
    for (int p = 0; p < 2; ++p) {
        Ent* e = &g->e[p];
        if (!moves[p]) continue;
        if (e->x == g->sx[p] && e->y == g->sy[p]) { descends[p] = 1; continue; }
        for (int i = NM; i < NM + ORX_R1_ITEMS; ++i) {       /* pickup */
            Ent* it = &g->e[i];
            if (it->alive && it->depth == e->depth && it->x == e->x && it->y == e->y && g->p[p].n_items < 4) {
                if (it->aux == 0) g->p[p].damage += 1;
                else if (it->aux == 1) g->p[p].armor += 1;
                else { g->p[p].max_hp += 2; e->hp += 2; }
                g->p[p].n_items += 1;
                it->alive = 0;
                emit(g, ORX_EV_PICKUP, p + 1, i + 1, it->aux, it->aux == 2 ? 2 : 0);
            }
        }
    }
    for (int p = 0; p < 2; ++p) {
        Ent* e = &g->e[p];
        if (!descends[p]) continue;
        int nd = e->depth + 1, x, y;
        e->alive = 0;                                       /* not an obstacle for its own spawn search */
        level_init(g, nd, &g->sx[p], &g->sy[p], &g->key[p]);
        free_tile(g, DOM_TICK, SUB_DESCEND + 64 * p, (uint32_t)g->tick, nd, g->key[p], g->sx[p], g->sy[p], &x, &y);
        e->alive = 1; e->depth = nd; e->x = x; e->y = y;
        if (g->e[1 - p].depth != nd) emit(g, ORX_EV_DUNGEON, 0, g->sx[p], g->sy[p], nd);   /* nobody stood there: the level is new */
        emit(g, ORX_EV_DESCEND, p + 1, x, y, nd);
        if (g->stats) g->stats[ORX_STAT_DESCENTS]++;
    }
    /* 6. enemy deaths, xp, drops */
    int died[NM] = { 0 };
    for (int m = 2; m < NM; ++m) {
        Ent* e = &g->e[m];
        if (!e->alive || e->hp > 0) continue;
        died[m] = 1;
        e->alive = 0;
        emit(g, ORX_EV_DEATH, m + 1, 0, 0, 0);
        for (int p = 0; p < 2; ++p) if (credit[m][p]) {
            int gained = 0;
            g->p[p].xp += 1;
            while (g->p[p].xp >= 3) { g->p[p].xp -= 3; g->p[p].level += 1; g->e[p].hp = g->p[p].max_hp; g->e[p].aux = g->p[p].max_mana; ++gained; }
            emit(g, ORX_EV_XP, p + 1, m + 1, gained, g->p[p].xp);
        }
    }
    /* drops in slot order, after every death of the tick has been recorded */
    for (int m = 2; m < NM; ++m) {
        Ent* e = &g->e[m];
        if (!died[m]) continue;
        uint32_t w[4];
        draw_block(&g->rng, DOM_TICK, SUB_DROP + (m - 2) / 2, (uint32_t)g->tick, w);
        uint32_t chance = w[2 * ((m - 2) & 1)], kind = w[2 * ((m - 2) & 1) + 1] % 3u;
        if (chance < (1u << 30))
            for (int i = NM; i < NM + ORX_R1_ITEMS; ++i)
                if (!g->e[i].alive) {
                    Ent* it = &g->e[i]; it->alive = 1; it->depth = e->depth; it->x = e->x; it->y = e->y; it->aux = (int)kind; it->hp = 0;
                    emit(g, ORX_EV_SPAWN, i + 1, it->x, it->y, it->depth | ((int)kind << 16));
                    break;
                }
    }
    /* 7. vanish + spawn */
    for (int l = 2; l < NM + ORX_R1_ITEMS; ++l)
        if (g->e[l].alive && g->e[l].depth != g->e[0].depth && g->e[l].depth != g->e[1].depth) { g->e[l].alive = 0; emit(g, ORX_EV_DEATH, l + 1, 1, 0, 0); }
    uint32_t sw[4] = { 0xFFFFFFFFu, 0xFFFFFFFFu, 0, 0 };
    if (g->tick % 4 == 0) draw_block(&g->rng, DOM_TICK, SUB_SPAWN, (uint32_t)g->tick, sw);
    for (int p = 0; p < 2; ++p) {
        if (p == 1 && g->e[1].depth == g->e[0].depth) continue;
        if (sw[p] >= (1u << 30)) continue;
        int slot = -1;
        for (int m = 2; m < NM; ++m) if (!g->e[m].alive) { slot = m; break; }
        if (slot < 0) continue;
        int x, y, d = g->e[p].depth;
        free_tile(g, DOM_TICK, SUB_SPAWN_TRY + 32 * p, (uint32_t)g->tick, d, g->key[p], g->sx[p], g->sy[p], &x, &y);
        Ent* e = &g->e[slot];
        e->alive = 1; e->depth = d; e->x = x; e->y = y; e->hp = imin(20, 2 + d / 2); e->aux = 0;
        emit(g, ORX_EV_SPAWN, slot + 1, x, y, d | (e->hp << 16));
    }
    /* 8. mana, 9. separation, 10. cooldowns */
    if (g->tick % 4 == 0) for (int p = 0; p < 2; ++p) g->e[p].aux = imin(g->p[p].max_mana, g->e[p].aux + 1);
    if (g->e[0].depth != g->e[1].depth) {
        g->sep += 1;
        int behind = g->e[0].depth < g->e[1].depth ? 0 : 1;
        g->e[behind].hp -= g->sep / 16;
        if (g->sep / 16 > 0) emit(g, ORX_EV_HEALTH, behind + 1, behind + 1, ORX_R1_HEALTH_SEPARATION, -(g->sep / 16));
    } else g->sep = 0;
    for (int p = 0; p < 2; ++p) g->p[p].cd = cd_pre[p] > 0 ? cd_pre[p] - 1 : newcd[p];
    uint32_t w[4];
    draw_block(&g->rng, DOM_TICK, SUB_MAIN, (uint32_t)g->tick, w);
    g->tick += 1;
    int d0 = g->e[0].hp <= 0, d1 = g->e[1].hp <= 0;
    if (d0) emit(g, ORX_EV_DEATH, 1, 0, 0, 0);
    if (d1) emit(g, ORX_EV_DEATH, 2, 0, 0, 0);
    if (d0 && d1) return (w[2] >> 31) ? ORX_RESULT_PLAYER1_WIN : ORX_RESULT_PLAYER2_WIN;
    if (d0) return ORX_RESULT_PLAYER2_WIN;
    if (d1) return ORX_RESULT_PLAYER1_WIN;
    if (cfg->max_ticks && g->tick >= cfg->max_ticks) return ORX_RESULT_TIE;
    return ORX_RESULT_IN_PROGRESS;
}

/* ---- SoA load / store */
static void load_game(G* g, const OrxR1Config* cfg, const OrxR1State* st, int64_t i, uint64_t gid)
{
    memset(g, 0, sizeof(*g));
    g->cfg = cfg; g->rng.seed = cfg->seed; g->rng.gid = gid; g->rng.episode = st->episode[i];
    for (int l = 0; l < ORX_R1_LANES; ++l) {
        uint32_t loc = st->ent_loc[16 * i + l], stat = st->ent_stat[16 * i + l];
        Ent* e = &g->e[l];
        e->x = loc & 255; e->y = (loc >> 8) & 255; e->alive = (loc >> 16) & 1; e->depth = st->ent_depth[16 * i + l];
        e->hp = (int16_t)(stat & 0xFFFF);
        e->aux = l >= NM ? (int)((loc >> 17) & 3) : (int16_t)(stat >> 16);
    }
    for (int p = 0; p < 2; ++p) {
        uint32_t a = st->pl_a[2 * i + p], b = st->pl_b[2 * i + p], c = st->pl_c[2 * i + p];
        Pl* q = &g->p[p];
        q->max_hp = (int16_t)(a & 0xFFFF); q->max_mana = (int16_t)(a >> 16);
        q->xp = b & 255; q->level = (b >> 8) & 255; q->n_items = (b >> 16) & 255; q->cd = b >> 24;
        q->damage = c & 255; q->armor = (c >> 8) & 255;
        g->sx[p] = (st->lvl_stairs[i] >> (16 * p)) & 255; g->sy[p] = (st->lvl_stairs[i] >> (16 * p + 8)) & 255;
        g->key[p] = st->lvl_key[2 * i + p];
    }
    g->sep = (int)st->sep[i]; g->tick = st->tick[i]; g->status = st->status[i];
}
static void store_game(const G* g, const OrxR1State* st, int64_t i)
{
    for (int l = 0; l < ORX_R1_LANES; ++l) {
        const Ent* e = &g->e[l];
        uint32_t loc = (uint32_t)(e->x & 255) | ((uint32_t)(e->y & 255) << 8) | ((uint32_t)(e->alive & 1) << 16);
        uint32_t stat = (uint32_t)e->hp & 0xFFFFu;
        if (l >= NM) loc |= (uint32_t)(e->aux & 3) << 17; else stat |= (uint32_t)e->aux << 16;
        if (!e->alive && l >= 2) { loc = 0; stat = 0; }
        st->ent_loc[16 * i + l] = loc; st->ent_stat[16 * i + l] = stat;
        st->ent_depth[16 * i + l] = (e->alive || l < 2) ? e->depth : 0;
    }
    for (int p = 0; p < 2; ++p) {
        const Pl* q = &g->p[p];
        st->pl_a[2 * i + p] = ((uint32_t)q->max_hp & 0xFFFFu) | ((uint32_t)q->max_mana << 16);
        st->pl_b[2 * i + p] = (uint32_t)(q->xp & 255) | ((uint32_t)(q->level & 255) << 8) | ((uint32_t)(q->n_items & 255) << 16) | ((uint32_t)q->cd << 24);
        st->pl_c[2 * i + p] = (uint32_t)(q->damage & 255) | ((uint32_t)(q->armor & 255) << 8);
        st->lvl_key[2 * i + p] = g->key[p];
    }
    st->lvl_stairs[i] = (uint32_t)g->sx[0] | ((uint32_t)g->sy[0] << 8) | ((uint32_t)g->sx[1] << 16) | ((uint32_t)g->sy[1] << 24);
    st->sep[i] = (uint32_t)g->sep; st->tick[i] = g->tick; st->episode[i] = g->rng.episode; st->status[i] = (uint8_t)g->status;
}
static void finish(G* g, int res, uint8_t* out)
{
    *out = (uint8_t)res;
    if (g->stats && res != ORX_RESULT_IN_PROGRESS)
        g->stats[res == ORX_RESULT_PLAYER1_WIN ? ORX_STAT_P1_WINS : res == ORX_RESULT_PLAYER2_WIN ? ORX_STAT_P2_WINS : ORX_STAT_TIES]++;
    if (res != ORX_RESULT_IN_PROGRESS && g->cfg->auto_reset) { g->rng.episode += 1; setup_game(g); }
    else g->status = res;
}

int oro_r1_reset(const OrxR1Config* cfg, const OrxR1State* st, const uint8_t* mask, int bump, int64_t n, uint64_t base)
{
    for (int64_t i = 0; i < n; ++i) {
        if (mask && !mask[i]) continue;
        G g; load_game(&g, cfg, st, i, base + (uint64_t)i);
        if (bump) g.rng.episode += 1;
        setup_game(&g); store_game(&g, st, i);
    }
    return 0;
}
int oro_r1_step(const OrxR1Config* cfg, const OrxR1State* st, const uint8_t* moves, uint8_t* result, int64_t n, uint64_t base)
{
    for (int64_t i = 0; i < n; ++i) {
        G g; load_game(&g, cfg, st, i, base + (uint64_t)i);
        if (g.status != ORX_RESULT_IN_PROGRESS) { result[i] = (uint8_t)g.status; continue; }
        int res = r1_tick(&g, moves[2 * i], moves[2 * i + 1]);
        finish(&g, res, &result[i]);
        store_game(&g, st, i);
    }
    return 0;
}
/* The tick with its replication log: events[i * max_events + k], terminated by kind == ORX_EV_NONE when fewer than
 * max_events records were written (slots behind the terminator are left alone); a frozen game writes the terminator only. */
int oro_r1_step_events(const OrxR1Config* cfg, const OrxR1State* st, const uint8_t* moves, uint8_t* result, OrxEvent* events,
                       int max_events, int64_t n, uint64_t base)
{
    for (int64_t i = 0; i < n; ++i) {
        G g; load_game(&g, cfg, st, i, base + (uint64_t)i);
        g.ev = events + i * max_events; g.cap_ev = max_events; g.n_ev = 0;
        if (g.status != ORX_RESULT_IN_PROGRESS) { result[i] = (uint8_t)g.status; }
        else {
            int res = r1_tick(&g, moves[2 * i], moves[2 * i + 1]);
            finish(&g, res, &result[i]);
            store_game(&g, st, i);
        }
        if (g.n_ev < max_events) memset(&g.ev[g.n_ev], 0, sizeof(OrxEvent));
    }
    return 0;
}
/* Scripted players for R1 (docs/RULESET_R1.md "Bots"): RandomBot draws uniformly from the six commands with the words the
 * fused rollout uses (TICK sub 0, word p), StaircaseBot walks towards the staircase of its own level
 * (optimax_rogue_bots/staircasebot.py:9-20). ORX_BOT_NONE leaves the player's byte untouched. */
int oro_r1_bot_moves(const OrxR1Config* cfg, const OrxR1State* st, int bot1, int bot2, uint8_t* moves, int64_t n, uint64_t base)
{
    for (int64_t i = 0; i < n; ++i) {
        G g; load_game(&g, cfg, st, i, base + (uint64_t)i);
        uint32_t w[4];
        draw_block(&g.rng, DOM_TICK, SUB_MAIN, (uint32_t)g.tick, w);
        for (int p = 0; p < 2; ++p) {
            int kind = p == 0 ? bot1 : bot2;
            if (kind == ORX_BOT_RANDOM) moves[2 * i + p] = (uint8_t)(1 + bounded(w[p], 6));
            else if (kind == ORX_BOT_STAIRCASE) {
                int ddx = g.sx[p] - g.e[p].x, ddy = g.sy[p] - g.e[p].y;
                moves[2 * i + p] = (uint8_t)(abs(ddx) > abs(ddy) ? (ddx > 0 ? ORX_MOVE_RIGHT : ORX_MOVE_LEFT) : (ddy > 0 ? ORX_MOVE_DOWN : ORX_MOVE_UP));
            }
        }
    }
    return 0;
}
int oro_r1_rollout(const OrxR1Config* cfg, const OrxR1State* st, int n_ticks, unsigned long long* stats, int64_t n, uint64_t base)
{
    unsigned long long total[ORX_STAT_COUNT] = { 0 };
#pragma omp parallel
    {
        unsigned long long local[ORX_STAT_COUNT] = { 0 };
#pragma omp for schedule(static)
        for (int64_t i = 0; i < n; ++i) {
            G g; load_game(&g, cfg, st, i, base + (uint64_t)i);
            g.stats = local;
            for (int t = 0; t < n_ticks && g.status == ORX_RESULT_IN_PROGRESS; ++t) {
                uint32_t w[4];
                draw_block(&g.rng, DOM_TICK, SUB_MAIN, (uint32_t)g.tick, w);
                uint8_t r;
                int res = r1_tick(&g, 1 + (int)bounded(w[0], 6), 1 + (int)bounded(w[1], 6));
                local[ORX_STAT_TICKS]++;
                finish(&g, res, &r);
            }
            store_game(&g, st, i);
        }
#pragma omp critical
        for (int k = 0; k < ORX_STAT_COUNT; ++k) total[k] += local[k];
    }
    if (stats) for (int k = 0; k < ORX_STAT_COUNT; ++k) stats[k] += total[k];
    return 0;
}

/* docs/RULESET_R1.md "Observation"; layout in include/orx.h (orx_r1_observe) */
int oro_r1_observe(const OrxR1Config* cfg, const OrxR1State* st, int16_t* obs, int radius, int64_t n)
{
    for (int64_t i = 0; i < n; ++i) {
        G g; load_game(&g, cfg, st, i, 0);
        for (int p = 0; p < 2; ++p) {
            int16_t* o = obs + (i * 2 + p) * ORX_R1_OBS_LEN;
            const Ent* me = &g.e[p]; const Ent* ot = &g.e[1 - p]; const Pl* q = &g.p[p];
            memset(o, 0, sizeof(int16_t) * ORX_R1_OBS_LEN);
            o[0] = (int16_t)me->x; o[1] = (int16_t)me->y; o[2] = (int16_t)imin(me->depth, 32767); o[3] = (int16_t)me->hp;
            o[4] = (int16_t)me->aux; o[5] = (int16_t)q->cd; o[6] = (int16_t)q->damage; o[7] = (int16_t)q->armor;
            o[8] = (int16_t)q->max_hp; o[9] = (int16_t)q->max_mana; o[10] = (int16_t)q->level; o[11] = (int16_t)q->xp;
            o[12] = (int16_t)q->n_items; o[13] = (int16_t)imin(g.sep, 32767); o[14] = (int16_t)imin(g.tick, 32767); o[15] = (int16_t)g.status;
            int same = ot->depth == me->depth;
            o[16] = (int16_t)same; o[17] = (int16_t)(same ? ot->x : -1); o[18] = (int16_t)(same ? ot->y : -1); o[19] = (int16_t)(same ? ot->hp : 0);
            int vis = radius < 0 || imax(abs(g.sx[p] - me->x), abs(g.sy[p] - me->y)) <= radius;
            o[20] = (int16_t)vis; o[21] = (int16_t)(vis ? g.sx[p] : -1); o[22] = (int16_t)(vis ? g.sy[p] : -1);
            for (int k = 0; k < ORX_R1_ENEMIES; ++k) {
                const Ent* e = &g.e[2 + k];
                int here = e->alive && e->depth == me->depth;
                o[23 + 3 * k] = (int16_t)(here ? e->x : -1); o[24 + 3 * k] = (int16_t)(here ? e->y : -1); o[25 + 3 * k] = (int16_t)(here ? e->hp : 0);
            }
            for (int k = 0; k < ORX_R1_ITEMS; ++k) {
                const Ent* e = &g.e[NM + k];
                int here = e->alive && e->depth == me->depth;
                o[47 + 3 * k] = (int16_t)(here ? e->x : -1); o[48 + 3 * k] = (int16_t)(here ? e->y : -1); o[49 + 3 * k] = (int16_t)(here ? e->aux : -1);
            }
            for (int t = 0; t < 49; ++t) {
                int x = me->x + t % 7 - 3, y = me->y + t / 7 - 3;
                if (is_wall(&g, g.key[p], g.sx[p], g.sy[p], x, y)) o[59 + t / 16] = (int16_t)((uint16_t)o[59 + t / 16] | (1u << (t % 16)));
            }
        }
    }
    return 0;
}
