/*
 * orx.h -- C ABI of liborx.so, the B200 batched replacement for Optimax Rogue's turn
 * dynamics. Plain pointers and sizes only; every device buffer is owned by the caller
 * (PyTorch tensors on the host side); the library allocates nothing persistent, keeps
 * no game state of its own (only a mutex-guarded cache of per-device launch geometry),
 * may be called from several host threads, and enqueues all work on the caller's CUDA stream.
 *
 * The reference (Tjstretchalot/optimax_rogue) is pure Python and has no FFI; its seams
 * are duck-typed objects. Each entry point below names the reference interface it
 * replaces (file:line under the reference root):
 *
 *   orx_step        Updater.update(game_state, player1_move, player2_move)
 *                       optimax_rogue/logic/updater.py:76-162  (+ handle_move :180-243,
 *                       handle_combat :298-338, handle_descend :259-296, calculate_pos :340-351)
 *                   GameState.on_tick  optimax_rogue/game/state.py:46-51 (derived stats)
 *                   called from Server.update  optimax_rogue/networking/server.py:126
 *   orx_reset       TogetherGameStartGenerator.setup_game   optimax_rogue/logic/worldgen.py:77-87
 *                   SeparatedGameStartGenerator.setup_game  optimax_rogue/logic/worldgen.py:124-135
 *                   EmptyDungeonGenerator.spawn_dungeon     optimax_rogue/logic/worldgen.py:33-43
 *                   Dungeon.get_random_unblocked            optimax_rogue/game/world.py:57-66
 *   orx_bot_moves   RandomBot.move     optimax_rogue_bots/randombot.py:20-21
 *                   StaircaseBot.move  optimax_rogue_bots/staircasebot.py:9-20
 *   orx_rollout     the tick loop  optimax_rogue/server/main.py:110-113 with both bots inlined
 *   orx_replay      the same loop with both players' commands queued in advance
 *   orx_observe     GameState.view_for  optimax_rogue/game/state.py:53-58
 *   orx_step_observe  Updater.update followed by view_for for both players (the self-play tick)
 *   orx_step_packed the same tick with both players' commands of a game in one byte (p1 | p2 << 4)
 *   orx_step_host   orx_step with host command/result buffers (what a remote caller holds);
 *   orx_step_host_packed(_sync)  the host-buffer tick with nibble-packed commands (half the PCIe bytes)
 *   orx_step_host_sync  the same plus a stream synchronisation (Server.update returns the result, server.py:132-138)
 *   orx_step_bits / orx_step_host_bits(_sync)  the same tick told and answered in the fewest bytes: 5 bits of command
 *                       pair in, 2 bits of result out per game
 *
 * Integer codes are the reference's enum values and must not change.
 */
#ifndef ORX_H_
#define ORX_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORX_ABI_VERSION 5

/* logic/moves.py:6-12 */
enum { ORX_MOVE_UP = 1, ORX_MOVE_RIGHT = 2, ORX_MOVE_DOWN = 3, ORX_MOVE_LEFT = 4, ORX_MOVE_STAY = 5 };
/* logic/updater.py:16-21 */
enum { ORX_RESULT_IN_PROGRESS = 1, ORX_RESULT_PLAYER1_WIN = 2, ORX_RESULT_PLAYER2_WIN = 3, ORX_RESULT_TIE = 4 };
/* game/world.py:10-17 */
enum { ORX_TILE_GROUND = 1, ORX_TILE_WALL = 2, ORX_TILE_STAIRCASE_DOWN = 3 };
/* game/modifiers.py:7-12 */
enum { ORX_FLAG_BLOCK = 1, ORX_FLAG_AMBUSH = 2, ORX_FLAG_FLEE = 3, ORX_FLAG_PARRY = 4 };
/* logic/updater.py:47-50 */
enum { ORX_DESPAWN_UNREACHABLE = 1, ORX_DESPAWN_UNUSED = 2 };

/* DungeonGenerator kinds (logic/worldgen.py:9-43). FIXED is a plugin generator that returns
 * the same tile grid at every depth (the reference loads generators by dotted path). */
enum { ORX_DGEN_EMPTY = 0, ORX_DGEN_FIXED = 1 };
/* GameStartGenerator kinds (logic/worldgen.py:60-137) */
enum { ORX_START_TOGETHER = 0, ORX_START_SEPARATED = 1 };
/* Bot policies (optimax_rogue_bots/) */
enum { ORX_BOT_NONE = 0, ORX_BOT_RANDOM = 1, ORX_BOT_STAIRCASE = 2 };

/* GameStateUpdate kinds emitted by the updater (logic/updates.py). MOVE and DESCEND are both
 * EntityPositionUpdate (:186); DESCEND has old_depth == depth - 1. */
enum { ORX_EV_NONE = 0, ORX_EV_MOVE = 1, ORX_EV_COMBAT = 2, ORX_EV_DUNGEON = 3, ORX_EV_DEATH = 4, ORX_EV_DESCEND = 5 };

#define ORX_NO_STAIRS 255      /* stairs plane value for a level without a staircase */
#define ORX_MAX_NPC 8          /* static NPC slots per game (updater.py:116-128) */
#define ORX_MAX_EVENTS_BASE 4  /* two movers x (DungeonCreated + Position) */
#define ORX_MAX_DIM 255        /* coordinates are stored as uint8 */

/* error codes: 0 ok, negative failure; -(100 + cudaError_t) wraps a CUDA error */
#define ORX_OK 0
#define ORX_ERR_BAD_ARG (-1)
#define ORX_ERR_UNSUPPORTED (-2)
#define ORX_ERR_CUDA_BASE (-100)

typedef struct OrxConfig {
    uint32_t struct_size;       /* = sizeof(OrxConfig), checked */
    int32_t width, height;      /* server/main.py:27-28 (--width 60 --height 10) */
    int32_t dgen_kind;          /* ORX_DGEN_* */
    int32_t start_kind;         /* ORX_START_* */
    int32_t start_depth[2];     /* worldgen.py:78 (0,0) / :110-111 (0,1000) */
    int32_t despawn_strat;      /* ORX_DESPAWN_*; affects DungeonCreated events only, because
                                   levels are pure functions of (seed, game, episode, depth) */
    int32_t max_ticks;          /* updater.py:158; 0 = unlimited */
    int32_t hp[2];              /* worldgen.py:85-86: health = base_max_health = 10 */
    int32_t damage[2];          /* base_damage = 2 */
    int32_t armor[2];           /* base_armor = 1 (damage dealt = attacker.damage - attacker.armor, updater.py:313) */
    int32_t auto_reset;         /* 1: a finished game is re-initialised (episode+1) inside the step */
    int32_t n_npc;              /* NPC slots in use, 0..ORX_MAX_NPC */
    uint64_t seed;              /* Philox key */
    /* ORX_DGEN_FIXED only (device pointers, shared by all games): */
    const uint8_t* fixed_tiles;   /* uint8[width*height], x-major (tiles[x*height+y]), ORX_TILE_*. Immutable while a
                                     tick that uses it is enqueued or running: a CTA stages the map when it starts,
                                     which may be before earlier work in the stream has finished */
    const uint16_t* fixed_ground; /* flat indices of the Ground tiles in ascending (x-major) order */
    int32_t fixed_n_ground;
    int32_t fixed_stairs[2];      /* first StaircaseDown in x-major order, or ORX_NO_STAIRS */
    uint32_t path_flags;          /* ORX_PATH_* bits, 0 = the default (fastest) kernels. Results never depend on them;
                                     they exist so that tests and A/B measurements can pin a code path without
                                     process-wide switches (the library reads no environment variables). */
} OrxConfig;

/* OrxConfig.path_flags. TILE_FLAGS and STATIC_TILES choose how launches on ONE state are ordered; keep them
 * constant for a state between two operations that serialise the stream (orx_reset, any non-step kernel, a sync). */
#define ORX_PATH_NO_TENSOR_MAP 1u   /* move the five 4-byte planes as five 1-D bulk copies */
#define ORX_PATH_NO_NPC_PIPE 2u     /* NPC slots: one-thread-per-game kernel instead of the tile pipeline */
#define ORX_PATH_STATIC_TILES 4u    /* grid-wait mode: static tile striding instead of the dynamic counter */
#define ORX_PATH_NO_EVENT_PIPE 8u   /* event log: one-thread-per-game kernel */
#define ORX_PATH_HOST_STAGED 16u    /* host buffers: staged cudaMemcpyAsync instead of in-kernel PCIe access */
#define ORX_PATH_TILE_FLAGS 32u     /* THROUGHPUT MODE (opt-in): order consecutive tick launches on a state chunk by
                                       chunk through OrxState.sched instead of grid by grid, so that ticks enqueued back
                                       to back overlap (see OrxState.sched). For queued command streams and for several
                                       states in flight; a loop that runs other kernels between two ticks (a policy
                                       network) is faster without it. Needs sched_words >= orx_sched_words(n); batches
                                       above 2^22 games are ticked grid by grid regardless. */
#define ORX_PATH_TILES_PER_CTA_SHIFT 8  /* bits 8..15: tiles per CTA in tile-flag mode (0 = built-in default) */

/* Structure-of-arrays game state; game i of the batch is element i of every plane.
 * Layout hint (optional, no effect on results): when the five 4-byte planes are carved out of one
 * allocation at a common pitch, i.e. hp = pos + pitch, stairs = pos + 2 pitch, tick = pos + 3 pitch,
 * episode = pos + 4 pitch with pitch a multiple of 16 bytes >= 4 n and pos 16-byte aligned, the tick /
 * observe kernels address them as one u32[5][n] array through a TMA tensor map and move a 256-game
 * tile of all five with a single copy instruction each way instead of five. Any other placement works
 * plane by plane. */
typedef struct OrxState {
    uint8_t* pos;       /* [n][4]  x1 y1 x2 y2                       Entity.x/.y   entities.py:35-37 */
    int16_t* hp;        /* [n][2]  health                            Entity.health entities.py:38 */
    int32_t* depth;     /* [n][2]  dungeon depth                     Entity.depth  entities.py:34 */
    uint8_t* stairs;    /* [n][4]  staircase (x,y) of each player's current level  world.py:52-55 */
    int32_t* tick;      /* [n]     GameState.tick (starts at 1)      state.py:29, worldgen.py:87 */
    uint32_t* episode;  /* [n]     resets seen by this lane (Philox counter word) */
    uint8_t* status;    /* [n]     ORX_RESULT_* of the lane */
    /* NPC slots, nullable when n_npc == 0; slot k of game i at [i*n_npc + k] */
    uint8_t* npc_pos;   /* [n][n_npc][2] */
    int16_t* npc_hp;    /* [n][n_npc]    */
    int32_t* npc_depth; /* [n][n_npc]    -1 = empty slot */
    /* Scratch of the tick kernel for THIS state: device uint32[sched_words], 16-byte aligned, zero-initialised
     * by the caller (orx_reset zeroes it again), or NULL. Never shared between states, and tied to the batch
     * size the state is ticked with. Two uses (csrc/orx_pipe.cuh):
     *   sched_words >= ORX_SCHED_HEADER_WORDS: word 0 is a tile counter, 256-game tiles are handed out
     *     dynamically (CTAs that run slower take fewer); zero again when a launch completes.
     *   sched_words >= orx_sched_words(n) and ORX_PATH_TILE_FLAGS set: words 4.. hold {tickets, completed passes} per
     *     run of tiles (a CTA's run: n_tiles / (SMs / 4) tiles, at most 32), and consecutive tick launches on the state
     *     are ordered run by run instead of grid by grid: CTA b of tick k+1 starts as soon as CTA b of tick k has
     *     written its run, and ticks of different states in one stream overlap freely. Measured per step with ticks
     *     enqueued back to back (rotating states), this mode against the default: 1.8 against 4.2 us at 2^17 games,
     *     2.9 / 5.5 at 2^18, 5.2 / 7.9 at 2^19, 9.9 / 12.2 at 2^20, 19.6 / 21.2 at 2^21. The protocol costs latency per
     *     launch (ticket, acquire, completion of the run's stores, release), so a tick that runs alone, with an
     *     ordinary kernel between two ticks or back to back on ONE small state is slower in this mode (9.4 against
     *     3.6 us at 2^17 games on one state): opt in only where ticks of several states really follow each other.
     *     Stream order towards everything else is kept (a tick completes only after all earlier work has).
     *     In this mode two consecutive tick calls on DIFFERENT states must not share a result / observation /
     *     event buffer unless something else in the stream consumes it in between. */
    uint32_t* sched;
    uint32_t sched_words;
    uint32_t reserved;
    /* Modifier seam (game/modifiers.py:92-108, game/attribles.py:21-43), nullable: int8[n][2][3], per player
     * { sum of flat_damage, sum of flat_armor, sum of flat_max_health } over the modifiers the entity carries --
     * what Entity.on_tick folds into damage.value / armor.value / max_health.value. A hit then deals
     * (base_damage + flat_damage) - (base_armor + flat_armor) of the ATTACKER (updater.py:313); flat_max_health is
     * carried for the host view only (nothing on the tick path reads max_health). Read-only for the tick. A state
     * that carries it is ticked by the one-thread-per-game kernels (the tile pipeline is compiled without the
     * look-up, so a state without modifiers pays nothing); it belongs to the lane, so a caller that models per-episode
     * modifiers rewrites it when a lane's result says the episode ended. Modifier EVENT hooks (pre/on/post_event)
     * are arbitrary Python upstream and have no counterpart here. */
    const int8_t* flat;
} OrxState;

/* One replication-log record (logic/updates.py); slots of game i at events[i*max_events + k],
 * in emission order, kind == ORX_EV_NONE terminates. */
typedef struct OrxEvent {
    uint8_t kind;   /* ORX_EV_* */
    uint8_t iden;   /* MOVE/DESCEND/DEATH: entity iden; COMBAT: attacker iden; DUNGEON: 0 */
    uint8_t a;      /* MOVE/DESCEND: posx; COMBAT: defender iden; DUNGEON: stair x */
    uint8_t b;      /* MOVE/DESCEND: posy; COMBAT: CombatFlag;    DUNGEON: stair y */
    int32_t depth;  /* MOVE/DESCEND/DUNGEON: (new) depth; COMBAT: og_damage */
} OrxEvent;

/* Aggregate counters written by orx_rollout (uint64 each). */
enum {
    ORX_STAT_TICKS = 0, ORX_STAT_P1_WINS = 1, ORX_STAT_P2_WINS = 2, ORX_STAT_TIES = 3,
    ORX_STAT_EVENTS = 4, ORX_STAT_DESCENTS = 5, ORX_STAT_HITS = 6, ORX_STAT_RESERVED = 7,
    ORX_STAT_COUNT = 8
};

int orx_abi_version(void);
const char* orx_strerror(int code);

/* Words of OrxState.sched that enable tile-by-tile ordering for a batch of n games. */
#define ORX_SCHED_HEADER_WORDS 4
size_t orx_sched_words(int64_t n);

/* Bytes of per-game mutable state (S_rw of the roofline formula) for this config. */
size_t orx_state_bytes(const OrxConfig* cfg);
/* Event slots per game per tick for this config (ORX_MAX_EVENTS_BASE + n_npc). */
int orx_max_events(const OrxConfig* cfg);

/* Updater.current_update_order over a batch (updater.py:71-74: every emitted GameStateUpdate takes the next
 * order number): order[i] += number of records game i emitted this tick (slots with kind != ORX_EV_NONE).
 * events: device OrxEvent[n][max_events] as written by orx_step; order: device uint64[n]. One pass. */
int orx_event_count_add(const OrxEvent* events, int max_events, unsigned long long* order, int64_t n,
                        void* cuda_stream);

/* Episode reset (worldgen.py:77-87 / :124-135). mask: device uint8[n], nullable = all lanes.
 * bump_episode != 0 increments episode[i] before drawing (use 0 for the first initialisation
 * or when the caller wrote the episode plane itself). */
int orx_reset(const OrxConfig* cfg, const OrxState* st, const uint8_t* mask, int bump_episode,
              int64_t n, uint64_t game_id_base, void* cuda_stream);

/* One tick for n games (updater.py:76-162). moves: device uint8[n][2] (p1, p2), codes outside
 * 1..5 are treated as Stay; result: device uint8[n]; events: device OrxEvent[n][orx_max_events]
 * or NULL. Lanes whose status is not IN_PROGRESS are left untouched (result = status). */
int orx_step(const OrxConfig* cfg, const OrxState* st, const uint8_t* moves, uint8_t* result,
             OrxEvent* events, int64_t n, uint64_t game_id_base, void* cuda_stream);

/* orx_step with HOST command/result buffers, all on cuda_stream; results are valid in result_host
 * once the stream has been synchronised. Pinned (page-locked) buffers are read/written directly by
 * the tick kernel over PCIe, tile by tile; pageable buffers go through H2D/D2H copies into the
 * caller-owned device staging buffers moves_dev/result_dev (same shapes). */
int orx_step_host(const OrxConfig* cfg, const OrxState* st, const uint8_t* moves_host,
                  uint8_t* result_host, uint8_t* moves_dev, uint8_t* result_dev, int64_t n,
                  uint64_t game_id_base, void* cuda_stream);

/* orx_step_host followed by a synchronisation of cuda_stream: on return result_host holds this
 * tick's results. One call per tick for a host-side loop that needs the results before it can
 * choose the next commands. */
int orx_step_host_sync(const OrxConfig* cfg, const OrxState* st, const uint8_t* moves_host,
                       uint8_t* result_host, uint8_t* moves_dev, uint8_t* result_dev, int64_t n,
                       uint64_t game_id_base, void* cuda_stream);

/* Nibble-packed commands: cmds uint8[n], game i's byte = p1_move | (p2_move << 4); nibbles outside
 * 1..5 are Stay. Everything else is orx_step / orx_step_host / orx_step_host_sync: same kernels,
 * same results, half the command bytes (the command stream is what bounds the host-buffer path).
 * cmds_dev/result_dev are only used for pageable host buffers and may be NULL for pinned ones. */
int orx_step_packed(const OrxConfig* cfg, const OrxState* st, const uint8_t* cmds, uint8_t* result,
                    OrxEvent* events, int64_t n, uint64_t game_id_base, void* cuda_stream);
int orx_step_host_packed(const OrxConfig* cfg, const OrxState* st, const uint8_t* cmds_host,
                         uint8_t* result_host, uint8_t* cmds_dev, uint8_t* result_dev, int64_t n,
                         uint64_t game_id_base, void* cuda_stream);
int orx_step_host_packed_sync(const OrxConfig* cfg, const OrxState* st, const uint8_t* cmds_host,
                              uint8_t* result_host, uint8_t* cmds_dev, uint8_t* result_dev, int64_t n,
                              uint64_t game_id_base, void* cuda_stream);

/* Bit-packed command / result streams: the fewest bytes a tick can be told and can answer with, for callers whose
 * buffers live in HOST memory (a game's command pair is one of 25 states, its result one of 4).
 *   cmd5: ORX_CMD5_BYTES(n) bytes; game i owns bits [5i, 5i+5) of the little-endian bit stream (bit b of the stream
 *         = bit b&7 of byte b>>3), value (p1_move - 1) * 5 + (p2_move - 1) with moves 1..5; 25..31 = both Stay.
 *   res2: ORX_RES2_BYTES(n) bytes; game i owns bits [2i, 2i+2), value ORX_RESULT_* - 1; padding bits are 0.
 * Both 16-byte aligned. Same tick as orx_step (logic/moves.py:6-12 and updater.py:16-21 are the codes packed here).
 * orx_step_bits takes device pointers (or device-mapped pinned host pointers); orx_step_host_bits[_sync] take host
 * buffers like orx_step_host[_sync]: pinned ones are read and written by the tick kernel itself, ONE PCIe
 * transaction per CTA each way (the commands of all its tiles when it starts, their results when it ends), pageable
 * ones go through the caller's device staging buffers cmd5_dev / res2_dev (same sizes, else nullable). No NPC slots. */
#define ORX_FMT_BYTES 0
#define ORX_FMT_NIBBLES 1
#define ORX_FMT_BITS 2
#define ORX_CMD5_BYTES(n) ((size_t)((5 * (uint64_t)(n) + 7) / 8))
#define ORX_RES2_BYTES(n) ((size_t)((2 * (uint64_t)(n) + 7) / 8))
int orx_step_bits(const OrxConfig* cfg, const OrxState* st, const uint8_t* cmd5, uint8_t* res2, int64_t n,
                  uint64_t game_id_base, void* cuda_stream);
int orx_step_host_bits(const OrxConfig* cfg, const OrxState* st, const uint8_t* cmd5_host, uint8_t* res2_host,
                       uint8_t* cmd5_dev, uint8_t* res2_dev, int64_t n, uint64_t game_id_base, void* cuda_stream);
int orx_step_host_bits_sync(const OrxConfig* cfg, const OrxState* st, const uint8_t* cmd5_host, uint8_t* res2_host,
                            uint8_t* cmd5_dev, uint8_t* res2_dev, int64_t n, uint64_t game_id_base, void* cuda_stream);

/* Command generation for scripted bots; ORX_BOT_NONE leaves that player's byte untouched. */
int orx_bot_moves(const OrxConfig* cfg, const OrxState* st, int bot_p1, int bot_p2,
                  uint8_t* moves, int64_t n, uint64_t game_id_base, void* cuda_stream);

/* n_ticks fused ticks with both bots on device; state stays in registers between ticks.
 * stats: device uint64[ORX_STAT_COUNT], accumulated (not cleared). */
int orx_rollout(const OrxConfig* cfg, const OrxState* st, int bot_p1, int bot_p2, int n_ticks,
                unsigned long long* stats, int64_t n, uint64_t game_id_base, void* cuda_stream);

/* n_ticks ticks of QUEUED commands in one launch: moves device uint8[n_ticks][n][2], results device
 * uint8[n_ticks][n]; identical to n_ticks calls of orx_step, with the state held in registers between
 * ticks (replaying recorded command streams, open-loop evaluation). */
int orx_replay(const OrxConfig* cfg, const OrxState* st, const uint8_t* moves, uint8_t* results, int n_ticks,
               int64_t n, uint64_t game_id_base, void* cuda_stream);

/* Per-player observation (state.py:53-58): obs int16[n][2][ORX_OBS_LEN]. */
#define ORX_OBS_LEN 12
int orx_observe(const OrxConfig* cfg, const OrxState* st, int16_t* obs, int stairs_radius,
                int64_t n, void* cuda_stream);
/* The rest of GameState.view_for (state.py:53-58) for states with NPC slots: the entities besides the players that
 * stand on the viewer's depth. npc_obs: device int16[n][2][n_npc][4] = per player, per slot { on_my_depth, x, y, health }
 * ({0, -1, -1, 0} for an empty slot or one on another depth), 8-byte aligned. n_npc == 0: nothing to do. */
int orx_observe_npc(const OrxConfig* cfg, const OrxState* st, int16_t* npc_obs, int64_t n, void* cuda_stream);

/* orx_step with scripted players (optimax_rogue_bots/randombot.py:20-21, staircasebot.py:9-20): a player whose
 * bot kind is ORX_BOT_RANDOM / ORX_BOT_STAIRCASE gets the command orx_bot_moves would compute for this tick, inside
 * the tick kernel (no second launch, no round trip of the commands through HBM); ORX_BOT_NONE takes the player's
 * command from moves (device uint8[n][2], always required). events and obs as in orx_step / orx_step_observe, both
 * nullable. For training a policy against a scripted opponent. */
int orx_step_bots(const OrxConfig* cfg, const OrxState* st, const uint8_t* moves, int bot_p1, int bot_p2,
                  uint8_t* result, OrxEvent* events, int16_t* obs, int stairs_radius, int64_t n,
                  uint64_t game_id_base, void* cuda_stream);

/* orx_step (moves_packed == 0: uint8[n][2]) or orx_step_packed (moves_packed != 0: uint8[n]) fused with
 * orx_observe of the resulting state: what a self-play loop needs per tick, in one pass over the
 * planes (61 + 48 bytes per game instead of 61 + 77). obs: device int16[n][2][ORX_OBS_LEN], 16-byte
 * aligned. Same results and observations as the two calls in sequence. */
int orx_step_observe(const OrxConfig* cfg, const OrxState* st, const uint8_t* moves, int moves_packed,
                     uint8_t* result, int16_t* obs, int stairs_radius, int64_t n, uint64_t game_id_base,
                     void* cuda_stream);

/* ---------------------------------------------------------------------------------------------
 * Ruleset R1: the README-only rules (readme.md:44-48,69-74), specified in docs/RULESET_R1.md.
 * PARITY UNPINNED -- the reference has no code for them. Separate state and entry points; R0 above
 * is unaffected. 16 entity lanes per game: 0-1 players, 2-9 enemies, 10-13 ground items.
 * ------------------------------------------------------------------------------------------- */
#define ORX_R1_LANES 16
#define ORX_R1_ENEMIES 8
#define ORX_R1_ITEMS 4
#define ORX_MOVE_HEAL 6            /* readme.md:74 */
#define ORX_R1_STATE_BYTES 241

typedef struct OrxR1Config {
    uint32_t struct_size;
    int32_t width, height;
    int32_t max_ticks;              /* 0 = unlimited */
    int32_t auto_reset;
    int32_t wall_density;           /* interior tile is a wall iff mix(x, y, key) & 255 < wall_density */
    uint64_t seed;
    uint32_t path_flags;            /* ORX_R1_PATH_*; 0 = default kernels. Results never depend on it */
    uint32_t reserved;
} OrxR1Config;

#define ORX_R1_PATH_HALFWARP 1u     /* the sixteen-lanes-per-game kernels instead of one thread per game */
#define ORX_R1_PATH_BLOCK_FLAGS 2u  /* throughput mode (opt-in): order consecutive orx_r1_step launches block by block through OrxR1State.sched */
#ifndef ORX_R1_BLOCK
#define ORX_R1_BLOCK 128            /* games per hand-over block of orx_r1_step (= threads per CTA of its kernel) */
#endif
#define ORX_R1_SCHED_WORDS(n) (2 * (((size_t)(n) + ORX_R1_BLOCK - 1) / ORX_R1_BLOCK))

typedef struct OrxR1State {
    uint32_t* ent_loc;     /* [n][16]  x | y<<8 | alive<<16 | item kind<<17 */
    int32_t* ent_depth;    /* [n][16] */
    uint32_t* ent_stat;    /* [n][16]  int16 hp | int16 mana (players) << 16 */
    uint32_t* pl_a;        /* [n][2]   max_hp | max_mana<<16 */
    uint32_t* pl_b;        /* [n][2]   xp | level<<8 | n_items<<16 | cd<<24 */
    uint32_t* pl_c;        /* [n][2]   damage | armor<<8 */
    uint32_t* lvl_stairs;  /* [n]      staircase xy of player 1's level | player 2's << 16 */
    uint32_t* lvl_key;     /* [n][2]   wall key of each player's level */
    uint32_t* sep;         /* [n]      ticks since the players separated */
    int32_t* tick;         /* [n] */
    uint32_t* episode;     /* [n] */
    uint8_t* status;       /* [n]      ORX_RESULT_* */
    /* Nullable scratch of orx_r1_step for this state, device uint32[sched_words >= ORX_R1_SCHED_WORDS(n)], zeroed by
     * the caller (orx_r1_reset zeroes it again): {tickets, completed passes} per block of ORX_R1_BLOCK games. With it
     * and ORX_R1_PATH_BLOCK_FLAGS, consecutive orx_r1_step launches are ordered block by block instead of grid by
     * grid (same protocol, same trade-off and same caveats as ORX_PATH_TILE_FLAGS / OrxState.sched). */
    uint32_t* sched;
    uint32_t sched_words;
    uint32_t reserved;
} OrxR1State;

int orx_r1_reset(const OrxR1Config* cfg, const OrxR1State* st, const uint8_t* mask, int bump_episode,
                 int64_t n, uint64_t game_id_base, void* cuda_stream);
/* moves: device uint8[n][2], codes 1..6; result: device uint8[n] */
int orx_r1_step(const OrxR1Config* cfg, const OrxR1State* st, const uint8_t* moves, uint8_t* result,
                int64_t n, uint64_t game_id_base, void* cuda_stream);
/* n_ticks fused ticks, both players uniform random over the six commands; stats as orx_rollout */
int orx_r1_rollout(const OrxR1Config* cfg, const OrxR1State* st, int n_ticks, unsigned long long* stats,
                   int64_t n, uint64_t game_id_base, void* cuda_stream);

/* Replication log of an R1 tick (logic/updates.py: the reference's GameStateUpdate classes, which its Updater appends to
 * in tick order, updater.py:71-74). Same 8-byte OrxEvent record as R0; idens are lane + 1 (players 1, 2; enemy slot e
 * 3 + e; item slot i 11 + i). Kinds beyond R0's, with their (iden, a, b, depth) fields:
 *   ORX_EV_SPAWN   EntitySpawnUpdate (updates.py:141): the new entity's iden, x, y, depth | aux << 16 (aux = hp of an
 *                  enemy, kind of an item)
 *   ORX_EV_HEALTH  EntityHealthUpdate (updates.py:222), health changed outside combat: entity iden, source iden,
 *                  ORX_R1_HEALTH_* tag, signed amount
 *   ORX_EV_PICKUP  a player took an item (the item's flat bonus is a Modifier, EntityModifierAddedUpdate updates.py:255):
 *                  player iden, item iden, item kind, health gained (2 for the max-health item)
 *   ORX_EV_XP      EntityEventUpdate (updates.py:31) 'xp': player iden, the enemy's iden, levels gained (a level-up
 *                  refills health and mana), xp after
 * and of R0's kinds: MOVE / DESCEND / DUNGEON as in R0; COMBAT: attacker iden, defender iden, ORX_R1_HIT_* tag, damage dealt;
 * DEATH: iden, a = 1 when the entity vanished with its level instead of dying.
 * Order within a tick (docs/RULESET_R1.md "Replication log"): heals; attacks by attacker lane; moves by lane; pickups by
 * player; descents by player (DUNGEON only when the other player is not on that depth already, then DESCEND); enemy
 * deaths by slot, each followed by its XP records; drops by slot; vanished enemies / items by lane; spawns by level;
 * separation damage; player deaths. At most 51 records per tick can occur; ORX_R1_MAX_EVENTS slots never overflow,
 * records beyond a smaller capacity are dropped. kind == ORX_EV_NONE terminates a game's list when it is shorter than
 * the capacity; slots behind the terminator are not written. */
enum { ORX_EV_SPAWN = 6, ORX_EV_HEALTH = 7, ORX_EV_PICKUP = 8, ORX_EV_XP = 9 };
enum { ORX_R1_HIT_FULL = 1, ORX_R1_HIT_HALF = 2, ORX_R1_HIT_NEGATED = 3, ORX_R1_HIT_CONTEST = 4 };
enum { ORX_R1_HEALTH_HEAL = 1, ORX_R1_HEALTH_SEPARATION = 2 };
#define ORX_R1_MAX_EVENTS 64
/* orx_r1_step + the tick's records: events device OrxEvent[n][max_events], 8-byte aligned, 1 <= max_events <= 255.
 * One-thread-per-game kernel (not with ORX_R1_PATH_HALFWARP: ORX_ERR_UNSUPPORTED). */
int orx_r1_step_events(const OrxR1Config* cfg, const OrxR1State* st, const uint8_t* moves, uint8_t* result,
                       OrxEvent* events, int max_events, int64_t n, uint64_t game_id_base, void* cuda_stream);
/* Scripted players (optimax_rogue_bots/randombot.py:20-21 widened to the six R1 commands -- the policy of
 * orx_r1_rollout, same draws --, staircasebot.py:9-20 on the player's own level). ORX_BOT_NONE leaves that player's
 * byte of moves (device uint8[n][2]) untouched. */
int orx_r1_bot_moves(const OrxR1Config* cfg, const OrxR1State* st, int bot1, int bot2, uint8_t* moves,
                     int64_t n, uint64_t game_id_base, void* cuda_stream);
/* n_ticks ticks with both players' commands queued in advance, the state in registers in between (the loop of
 * server/main.py:110-113 fed from a buffer): moves device uint8[n_ticks][n][2], results device uint8[n_ticks][n].
 * Identical to n_ticks orx_r1_step calls; a game that ends without auto_reset repeats its status. */
int orx_r1_replay(const OrxR1Config* cfg, const OrxR1State* st, const uint8_t* moves, uint8_t* results, int n_ticks,
                  int64_t n, uint64_t game_id_base, void* cuda_stream);
/* orx_r1_step with the caller's PINNED HOST buffers (cudaHostAlloc / cudaHostRegister; host_moves uint8[n][2],
 * host_result uint8[n]): the kernel reads the commands and writes the results across PCIe itself, then the call waits
 * for the stream -- on return host_result holds this tick's results. */
int orx_r1_step_host_sync(const OrxR1Config* cfg, const OrxR1State* st, const uint8_t* host_moves, uint8_t* host_result,
                          int64_t n, uint64_t game_id_base, void* cuda_stream);

/* Per-player observation for ruleset R1: obs int16[n][2][ORX_R1_OBS_LEN] (GameState.view_for,
 * game/state.py:53-58, widened to the R1 state; "a ladder ... becomes visible when an agent gets
 * near it", readme.md:44 -> stairs_radius, Chebyshev, < 0 = always visible).
 *   0..7   x y depth hp mana cd damage armor        8..15  max_hp max_mana level xp n_items sep tick status
 *   16..19 opponent: on_my_depth x y hp (else 0 -1 -1 0)        20..22 stairs: visible x y (else 0 -1 -1)
 *   23..46 enemies 8 x (x y hp) on my depth (else -1 -1 0)      47..58 items 4 x (x y kind) on my depth (else -1)
 *   59..62 walls of the 7x7 window centred on the player, row-major (dy, dx = -3..3), bit k of word k/16;
 *          tiles off the map count as walls                       63 reserved (0)
 * depth, sep and tick saturate at 32767. */
#define ORX_R1_OBS_LEN 64
int orx_r1_observe(const OrxR1Config* cfg, const OrxR1State* st, int16_t* obs, int stairs_radius,
                   int64_t n, void* cuda_stream);

#ifdef __cplusplus
}
#endif
#endif /* ORX_H_ */
