// liborx.so: kernels + the C ABI declared in include/orx.h. sm_100a only.
#include <cuda_runtime.h>
#include <stdlib.h>
#include <string.h>

#include <mutex>
#include <type_traits>

#include "../../include/orx.h"
#include "orx_rules.cuh"
#include "orx_pipe.cuh"

using namespace orx;

namespace {

constexpr int kThreads = 256;
constexpr int kMaxFixedTiles = 16384;   // shared-memory staging of the fixed map
#ifndef ORX_FLAG_MODE_MAX_TILES
#define ORX_FLAG_MODE_MAX_TILES 16384
#endif
constexpr unsigned int kFlagModeMaxTiles = ORX_FLAG_MODE_MAX_TILES;   // 2^22 games: the largest batch ticked in flag mode (tile_ctl)
constexpr int64_t kMaxGamesPerCall = 1ll << 30;   // 32-bit lane index inside the kernels; larger batches: call per chunk

__device__ __forceinline__ uint32_t ldg_u32(const uint32_t* p)
{
    uint32_t v;
    asm volatile("ld.global.u32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ int2 ldg_s32x2(const int2* p)
{
    int2 v;
    asm volatile("ld.global.v2.s32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p));
    return v;
}
__device__ __forceinline__ uint32_t ldg_u8(const uint8_t* p)
{
    uint32_t v;
    asm volatile("ld.global.u8 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ uint32_t ldg_u16(const uint16_t* p)
{
    uint32_t v;
    asm volatile("ld.global.u16 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}

__device__ __forceinline__ const uint8_t* stage_tiles(const Params& P, uint8_t* smem)
{
    // The fixed wall map is shared by every game of the batch: stage it once per CTA.
    const int nt = P.W * P.H;
    for (int t = threadIdx.x; t < nt; t += blockDim.x) smem[t] = P.tiles[t];
    __syncthreads();
    return smem;
}

// ------------------------------------------------------------------ K1: one tick
// One thread per game. All eight plane loads are issued before anything depends on them; the
// per-game traffic is 29 B of planes in, 29 B out, 2 B of commands in, 1 B of result out.
template <int DGEN, bool NPC, bool EV>
__global__ void __launch_bounds__(kThreads, 4)
k_step(const __grid_constant__ Params P, const void* __restrict__ moves, uint8_t* __restrict__ result,
       uint2* __restrict__ events, int max_ev, int fmt, int bots = 0)
{
    extern __shared__ uint8_t smem[];
    __shared__ CmdEntry lut[256];
    build_cmd_lut(P, lut, threadIdx.x, kThreads);
    const uint8_t* tiles = nullptr;
    if (DGEN == ORX_DGEN_FIXED) tiles = stage_tiles(P, smem);
    else __syncthreads();
    const unsigned int i = blockIdx.x * kThreads + threadIdx.x;
    const bool in = i < P.n;
    int res = 0;
    if (in) {
        // Eight independent loads in flight per thread before the first use (asm volatile keeps ptxas
        // from sinking them below the frozen-lane test, which would serialise two DRAM round trips).
        const uint32_t pos = ldg_u32(P.pos + i), hpw = ldg_u32(P.hp + i), stw = ldg_u32(P.stairs + i);
        const uint32_t ep = ldg_u32(P.episode + i);
        const int tick = (int)ldg_u32(reinterpret_cast<const uint32_t*>(P.tick) + i);
        const int2 dep = ldg_s32x2(P.depth + i);
        const int status = (int)ldg_u8(P.status + i);
        uint32_t mv;
        if (fmt == ORX_FMT_NIBBLES) {     // one byte per game: p1 in the low nibble, p2 in the high nibble
            const uint32_t c = ldg_u8(static_cast<const uint8_t*>(moves) + i);
            mv = (c & 15u) | ((c >> 4) << 8);
        } else if (fmt == ORX_FMT_BITS) { // five bits per game: (p1 - 1) * 5 + (p2 - 1); 25..31 = both Stay
            const uint8_t* c = static_cast<const uint8_t*>(moves) + ((size_t)i * 5u >> 3);
            const uint32_t sh = (i * 5u) & 7u;
            uint32_t w = ldg_u8(c);
            if (sh > 3u) w |= ldg_u8(c + 1) << 8;
            const uint32_t v = (w >> sh) & 31u, p1 = v / 5u;
            mv = v < 25u ? ((p1 + 1u) | ((v - 5u * p1 + 1u) << 8)) : 0u;
        } else {
            mv = ldg_u16(static_cast<const uint16_t*>(moves) + i);
        }
        EvSink<EV> ev{EV ? events + (size_t)i * max_ev : nullptr, 0, max_ev};
        if (status != ORX_RESULT_IN_PROGRESS) {   // finished lanes are frozen until reset
            res = status;
            ev.finish();
        } else {
            Lane L;
            unpack_lane(L, pos, hpw, dep, stw, tick, ep);
            Stream s = make_stream(P, i, ep);
            const uint4 blk = draw_block(s, DOM_TICK, SUB_MAIN, (uint32_t)tick);
            if (bots != 0) {                      // scripted players (orx_step_bots): kind of p1 | kind of p2 << 8
                if ((bots & 255) != ORX_BOT_NONE) mv = (mv & 0xFF00u) | bot_move(bots & 255, L.pos & 0xFFFFu, L.st & 0xFFFFu, blk.x);
                if ((bots >> 8) != ORX_BOT_NONE) mv = (mv & 0x00FFu) | (bot_move(bots >> 8, L.pos >> 16, L.st >> 16, blk.y) << 8);
            }
            Counters cnt{};
            using NV = std::conditional_t<NPC, NpcView, NoNpc>;
            NV nv;
            if constexpr (NPC) nv = npc_view(P, i);
            res = tick_lane<DGEN, NV, EV, true>(P, tiles, lut, L, mv, blk.z, s, nv, ev, cnt);
            ev.finish();
            int new_status = res;
            if (res != ORX_RESULT_IN_PROGRESS && P.auto_reset) {
                s.episode += 1;
                reset_lane<DGEN, NV>(P, L, s, nv);
                new_status = ORX_RESULT_IN_PROGRESS;
            }
            store_lane(P, i, L, new_status);
        }
    }
    if (fmt == ORX_FMT_BITS) {
        // two bits per game (result - 1), four games per byte: the lane of a byte's first game gathers the other three
        const uint32_t r = in ? (uint32_t)(res - 1) & 3u : 0u;
        const uint32_t lo = __ballot_sync(0xffffffffu, (r & 1u) != 0u), hi = __ballot_sync(0xffffffffu, (r & 2u) != 0u);
        const uint32_t q = threadIdx.x & 28u;
        auto spread4 = [](uint32_t x) { return (x & 1u) | ((x & 2u) << 1) | ((x & 4u) << 2) | ((x & 8u) << 3); };
        if (in && (threadIdx.x & 3u) == 0u) result[i >> 2] = (uint8_t)(spread4((lo >> q) & 15u) | (spread4((hi >> q) & 15u) << 1));
    } else if (in) {
        result[i] = (uint8_t)res;
    }
}

// ------------------------------------------------------------------ K2: masked episode reset
template <int DGEN, bool NPC>
__global__ void __launch_bounds__(kThreads)
k_reset(const __grid_constant__ Params P, const uint8_t* __restrict__ mask, int bump)
{
    const unsigned int i = blockIdx.x * kThreads + threadIdx.x;
    if (i >= P.n) return;
    if (mask != nullptr && mask[i] == 0) return;
    Lane L;
    Stream s = make_stream(P, i, P.episode[i] + (bump ? 1u : 0u));
    using NV = std::conditional_t<NPC, NpcView, NoNpc>;
    NV nv;
    if constexpr (NPC) nv = npc_view(P, i);
    reset_lane<DGEN, NV>(P, L, s, nv);
    store_lane(P, i, L, ORX_RESULT_IN_PROGRESS);
}

// ------------------------------------------------------------------ K3: scripted bots
__global__ void __launch_bounds__(kThreads)
k_bot_moves(const __grid_constant__ Params P, int bot1, int bot2, uint8_t* __restrict__ moves)
{
    const unsigned int i = blockIdx.x * kThreads + threadIdx.x;
    if (i >= P.n) return;
    const bool need_rng = bot1 == ORX_BOT_RANDOM || bot2 == ORX_BOT_RANDOM;
    const uint32_t pos = P.pos[i], st = P.stairs[i];
    uint4 blk = make_uint4(0, 0, 0, 0);
    if (need_rng) blk = draw_block(make_stream(P, i, P.episode[i]), DOM_TICK, SUB_MAIN, (uint32_t)P.tick[i]);
    if (bot1 != ORX_BOT_NONE && bot2 != ORX_BOT_NONE) {
        const uint32_t m1 = bot_move(bot1, pos & 0xFFFFu, st & 0xFFFFu, blk.x);
        const uint32_t m2 = bot_move(bot2, pos >> 16, st >> 16, blk.y);
        reinterpret_cast<uint16_t*>(moves)[i] = (uint16_t)(m1 | (m2 << 8));
    } else if (bot1 != ORX_BOT_NONE) {
        moves[2 * (size_t)i] = (uint8_t)bot_move(bot1, pos & 0xFFFFu, st & 0xFFFFu, blk.x);
    } else if (bot2 != ORX_BOT_NONE) {
        moves[2 * (size_t)i + 1] = (uint8_t)bot_move(bot2, pos >> 16, st >> 16, blk.y);
    }
}

// ------------------------------------------------------------------ fused rollout (K1+K3, T ticks)
__device__ __forceinline__ unsigned int warp_sum(unsigned int v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

#ifndef ORX_ROLLOUT_MINBLOCKS
#define ORX_ROLLOUT_MINBLOCKS 3
#endif
template <int DGEN, bool NPC>
__global__ void __launch_bounds__(kThreads, ORX_ROLLOUT_MINBLOCKS)
k_rollout(const __grid_constant__ Params P, int bot1, int bot2, int n_ticks, unsigned long long* __restrict__ stats)
{
    extern __shared__ uint8_t smem[];
    const uint8_t* tiles = nullptr;
    if (DGEN == ORX_DGEN_FIXED) tiles = stage_tiles(P, smem);
    __shared__ unsigned int s_cnt[7];
    __shared__ CmdEntry lut[256];
    build_cmd_lut(P, lut, threadIdx.x, kThreads);
    if (threadIdx.x < 7) s_cnt[threadIdx.x] = 0;
    __syncthreads();
    Counters cnt{};
    const unsigned int i = blockIdx.x * kThreads + threadIdx.x;
    if (i < P.n) {
        int status = P.status[i];
        if (status == ORX_RESULT_IN_PROGRESS) {
            Lane L;
            load_lane(P, i, L);
            Stream s = make_stream(P, i, L.episode);
            EvSink<false> ev{nullptr, 0, 0};
            using NV = std::conditional_t<NPC, NpcView, NoNpc>;
            NV nv;
            if constexpr (NPC) nv = npc_view(P, i);
            for (int t = 0; t < n_ticks; ++t) {
                const uint4 blk = draw_block(s, DOM_TICK, SUB_MAIN, (uint32_t)L.tick);
                const uint32_t m1 = bot_move(bot1, L.pos & 0xFFFFu, L.st & 0xFFFFu, blk.x);
                const uint32_t m2 = bot_move(bot2, L.pos >> 16, L.st >> 16, blk.y);
                const int res = tick_lane<DGEN, NV, false, true>(P, tiles, lut, L, m1 | (m2 << 8), blk.z, s, nv, ev, cnt);
                ++cnt.ticks;
                if (res != ORX_RESULT_IN_PROGRESS) {
                    cnt.p1 += res == ORX_RESULT_PLAYER1_WIN;
                    cnt.p2 += res == ORX_RESULT_PLAYER2_WIN;
                    cnt.ties += res == ORX_RESULT_TIE;
                    if (P.auto_reset) {
                        s.episode += 1;
                        reset_lane<DGEN, NV>(P, L, s, nv);
                    } else {
                        status = res;
                        break;
                    }
                }
            }
            cnt.events += (unsigned int)ev.n;
            store_lane(P, i, L, status);
        }
    }
    if (stats != nullptr) {
        unsigned int v[7] = {cnt.ticks, cnt.p1, cnt.p2, cnt.ties, cnt.events, cnt.descents, cnt.hits};
#pragma unroll
        for (int k = 0; k < 7; ++k) {
            const unsigned int w = warp_sum(v[k]);
            if ((threadIdx.x & 31) == 0 && w != 0) atomicAdd(&s_cnt[k], w);
        }
        __syncthreads();
        if (threadIdx.x < 7 && s_cnt[threadIdx.x] != 0)
            atomicAdd(&stats[threadIdx.x], (unsigned long long)s_cnt[threadIdx.x]);
    }
}

// ------------------------------------------------------------------ replay: T ticks of queued commands
template <int DGEN, bool NPC>
__global__ void __launch_bounds__(kThreads, ORX_ROLLOUT_MINBLOCKS)
k_replay(const __grid_constant__ Params P, const uint16_t* __restrict__ moves, uint8_t* __restrict__ results, int n_ticks)
{
    extern __shared__ uint8_t smem[];
    __shared__ CmdEntry lut[256];
    build_cmd_lut(P, lut, threadIdx.x, kThreads);
    const uint8_t* tiles = nullptr;
    if (DGEN == ORX_DGEN_FIXED) tiles = stage_tiles(P, smem);
    else __syncthreads();
    const unsigned int i = blockIdx.x * kThreads + threadIdx.x;
    if (i >= P.n) return;
    int status = P.status[i];
    Lane L;
    load_lane(P, i, L);
    Stream s = make_stream(P, i, L.episode);
    EvSink<false> ev{nullptr, 0, 0};
    Counters cnt{};
    using NV = std::conditional_t<NPC, NpcView, NoNpc>;
    NV nv;
    if constexpr (NPC) nv = npc_view(P, i);
    for (int t = 0; t < n_ticks; ++t) {
        const size_t at = (size_t)t * P.n + i;
        if (status != ORX_RESULT_IN_PROGRESS) { results[at] = (uint8_t)status; continue; }   // frozen lane
        const uint32_t mv = moves[at];
        const uint4 blk = draw_block(s, DOM_TICK, SUB_MAIN, (uint32_t)L.tick);
        const int res = tick_lane<DGEN, NV, false, true>(P, tiles, lut, L, mv, blk.z, s, nv, ev, cnt);
        results[at] = (uint8_t)res;
        if (res != ORX_RESULT_IN_PROGRESS) {
            if (P.auto_reset) { s.episode += 1; reset_lane<DGEN, NV>(P, L, s, nv); }
            else status = res;
        }
    }
    store_lane(P, i, L, status);
}

// ------------------------------------------------------------------ observations (state.py:53-58)
// obs[i][p] = { x, y, depth (saturated), hp, other_visible, other_x, other_y, other_hp,
//               stairs_visible, stairs_x, stairs_y, tick (saturated) }
__global__ void __launch_bounds__(kThreads)
k_observe(const __grid_constant__ Params P, int16_t* __restrict__ obs, int radius)
{
    const unsigned int i = blockIdx.x * kThreads + threadIdx.x;
    if (i >= P.n) return;
    Lane L;
    load_lane(P, i, L);
    uint32_t w[12];
    pack_obs(L, radius, w);
    uint4* o = reinterpret_cast<uint4*>(obs + (size_t)i * 2 * ORX_OBS_LEN);       // 48 B per game, 16-byte aligned
    o[0] = make_uint4(w[0], w[1], w[2], w[3]);
    o[1] = make_uint4(w[4], w[5], w[6], w[7]);
    o[2] = make_uint4(w[8], w[9], w[10], w[11]);
}

// orx_observe_npc: what view_for (state.py:53-58) keeps of the entities besides the players -- the NPC slots on the
// viewer's depth. One thread per game, 8 bytes per (player, slot).
__global__ void __launch_bounds__(kThreads)
k_observe_npc(const __grid_constant__ Params P, uint2* __restrict__ out)
{
    const unsigned int i = blockIdx.x * kThreads + threadIdx.x;
    if (i >= P.n) return;
    const int2 d = P.depth[i];
    const int e = P.n_npc;
    for (int k = 0; k < e; ++k) {
        const size_t at = (size_t)i * e + k;
        const int nd = P.npc_depth[at], hp = P.npc_hp[at];
        const int x = P.npc_pos[2 * at], y = P.npc_pos[2 * at + 1];
#pragma unroll
        for (int p = 0; p < 2; ++p) {
            const bool here = nd >= 0 && nd == (p == 0 ? d.x : d.y);
            out[((size_t)i * 2 + p) * e + k] = make_uint2(pack_i16x2(here, here ? x : -1), pack_i16x2(here ? y : -1, here ? hp : 0));
        }
    }
}

// ------------------------------------------------------------------ Updater.current_update_order (updater.py:71-74)
__global__ void __launch_bounds__(kThreads)
k_event_count_add(const uint2* __restrict__ events, int max_ev, unsigned long long* __restrict__ order, unsigned int n)
{
    const unsigned int i = blockIdx.x * kThreads + threadIdx.x;
    if (i >= n) return;
    const uint2* e = events + (size_t)i * max_ev;
    unsigned int c = 0;
    if (max_ev == ORX_MAX_EVENTS_BASE) {              // 32 bytes per game: two 16-byte loads
        const uint4 a = reinterpret_cast<const uint4*>(e)[0], b = reinterpret_cast<const uint4*>(e)[1];
        c = ((a.x & 255u) != 0) + ((a.z & 255u) != 0) + ((b.x & 255u) != 0) + ((b.z & 255u) != 0);
    } else {
        for (int k = 0; k < max_ev; ++k) c += (e[k].x & 255u) != 0;
    }
    if (c != 0) order[i] += c;
}

// ------------------------------------------------------------------ host side
int cuda_fail(cudaError_t e) { return ORX_ERR_CUDA_BASE - (int)e; }

int launch_done() {
    const cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? ORX_OK : cuda_fail(e);
}

bool aligned(const void* p, size_t a) { return (reinterpret_cast<uintptr_t>(p) & (a - 1)) == 0; }

bool ids_ok(uint64_t base, int64_t n) { return (base >> 54) == 0 && ((base + (uint64_t)(n > 0 ? n : 0)) >> 54) == 0; }   // Philox counter holds 54 id bits

int check_common(const OrxConfig* cfg, const OrxState* st, int64_t n)
{
    if (cfg == nullptr || st == nullptr || n < 0) return ORX_ERR_BAD_ARG;
    if (n > kMaxGamesPerCall) return ORX_ERR_UNSUPPORTED;
    if (cfg->struct_size != sizeof(OrxConfig)) return ORX_ERR_BAD_ARG;
    if (cfg->width < 4 || cfg->height < 4 || cfg->width > ORX_MAX_DIM || cfg->height > ORX_MAX_DIM) return ORX_ERR_BAD_ARG;
    if (cfg->n_npc < 0 || cfg->n_npc > ORX_MAX_NPC) return ORX_ERR_BAD_ARG;
    if (cfg->dgen_kind != ORX_DGEN_EMPTY && cfg->dgen_kind != ORX_DGEN_FIXED) return ORX_ERR_UNSUPPORTED;
    if (cfg->start_kind != ORX_START_TOGETHER && cfg->start_kind != ORX_START_SEPARATED) return ORX_ERR_UNSUPPORTED;
    if (cfg->despawn_strat != ORX_DESPAWN_UNREACHABLE && cfg->despawn_strat != ORX_DESPAWN_UNUSED) return ORX_ERR_BAD_ARG;
    if (cfg->start_kind == ORX_START_SEPARATED && cfg->start_depth[0] == cfg->start_depth[1]) return ORX_ERR_BAD_ARG;
    if (cfg->max_ticks < 0) return ORX_ERR_BAD_ARG;
    for (int k = 0; k < 2; ++k) {   // health lives in int16 planes; a hit is subtracted from them
        if (cfg->hp[k] < 1 || cfg->hp[k] > 32767) return ORX_ERR_BAD_ARG;
        const int64_t hit = (int64_t)cfg->damage[k] - (int64_t)cfg->armor[k];
        if (hit < -32767 + 254 || hit > 32767 - 254) return ORX_ERR_BAD_ARG;      // room for the int8 flat bonuses (OrxState.flat)
    }
    if (cfg->dgen_kind == ORX_DGEN_FIXED) {
        if (cfg->fixed_tiles == nullptr || cfg->fixed_ground == nullptr) return ORX_ERR_BAD_ARG;
        if (cfg->fixed_n_ground < 2 + cfg->n_npc) return ORX_ERR_BAD_ARG;
        if (cfg->width * cfg->height > kMaxFixedTiles) return ORX_ERR_UNSUPPORTED;
        if (!aligned(cfg->fixed_ground, 2)) return ORX_ERR_BAD_ARG;
    }
    if (!st->pos || !st->hp || !st->depth || !st->stairs || !st->tick || !st->episode || !st->status) return ORX_ERR_BAD_ARG;
    if (!aligned(st->pos, 4) || !aligned(st->hp, 4) || !aligned(st->depth, 8) || !aligned(st->stairs, 4) ||
        !aligned(st->tick, 4) || !aligned(st->episode, 4)) return ORX_ERR_BAD_ARG;
    if (cfg->n_npc > 0 && (!st->npc_pos || !st->npc_hp || !st->npc_depth || !aligned(st->npc_hp, 2) || !aligned(st->npc_depth, 4)))
        return ORX_ERR_BAD_ARG;
    return ORX_OK;
}

Params make_params(const OrxConfig* c, const OrxState* st, int64_t n, uint64_t gid_base)
{
    Params P;
    memset(&P, 0, sizeof(P));
    P.W = c->width; P.H = c->height;
    P.start_kind = c->start_kind; P.sd0 = c->start_depth[0]; P.sd1 = c->start_depth[1];
    P.despawn = c->despawn_strat; P.max_ticks = c->max_ticks;
    P.hp0 = c->hp[0]; P.hp1 = c->hp[1];
    P.dmg0 = c->damage[0] - c->armor[0]; P.dmg1 = c->damage[1] - c->armor[1];
    P.auto_reset = c->auto_reset; P.n_npc = c->n_npc;
    make_round_keys(P.rk, (uint32_t)c->seed, (uint32_t)(c->seed >> 32));
    P.tiles = c->fixed_tiles; P.ground = c->fixed_ground; P.n_ground = c->fixed_n_ground;
    P.fsx = c->fixed_stairs[0]; P.fsy = c->fixed_stairs[1];
    P.pos = reinterpret_cast<uint32_t*>(st->pos); P.hp = reinterpret_cast<uint32_t*>(st->hp);
    P.depth = reinterpret_cast<int2*>(st->depth); P.stairs = reinterpret_cast<uint32_t*>(st->stairs);
    P.tick = st->tick; P.episode = st->episode; P.status = st->status;
    P.npc_pos = st->npc_pos; P.npc_hp = st->npc_hp; P.npc_depth = st->npc_depth;
    P.flat = st->flat;
    P.n = (unsigned int)n; P.gid_base = gid_base;
    return P;
}

int grid_for(int64_t n) { return (int)((n + kThreads - 1) / kThreads); }


size_t tiles_smem(const OrxConfig* c) { return c->dgen_kind == ORX_DGEN_FIXED ? (size_t)c->width * c->height : 0; }

template <typename F>
int dispatch_dgen_npc(const OrxConfig* c, F&& f)
{
    const bool npc = c->n_npc > 0;
    if (c->dgen_kind == ORX_DGEN_EMPTY) return npc ? f.template operator()<ORX_DGEN_EMPTY, true>() : f.template operator()<ORX_DGEN_EMPTY, false>();
    return npc ? f.template operator()<ORX_DGEN_FIXED, true>() : f.template operator()<ORX_DGEN_FIXED, false>();
}


bool host_mapped(const void* p, void** dev)
{
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
    if (a.type != cudaMemoryTypeHost || a.devicePointer == nullptr) return false;
    *dev = a.devicePointer;
    return true;
}

bool pipe_aligned(const OrxState* st, const void* moves, const void* result)
{
    return aligned(st->pos, 16) && aligned(st->hp, 16) && aligned(st->depth, 16) && aligned(st->stairs, 16) &&
           aligned(st->tick, 16) && aligned(st->episode, 16) && aligned(st->status, 16) && aligned(moves, 16) &&
           aligned(result, 16);
}

Params offset_params(const Params& P, int64_t off, int64_t n)
{
    Params T = P;
    T.pos += off; T.hp += off; T.depth += off; T.stairs += off; T.tick += off; T.episode += off; T.status += off;
    if (P.n_npc > 0) { T.npc_pos += 2 * off * P.n_npc; T.npc_hp += off * P.n_npc; T.npc_depth += off * P.n_npc; }
    if (P.flat != nullptr) T.flat += 6 * off;
    T.n = (unsigned int)n; T.gid_base = P.gid_base + (unsigned long long)off;
    return T;
}

// cuTensorMapEncodeTiled through the runtime's driver entry point table: liborx.so does not link libcuda, so
// it still loads (for the ABI checks) on a machine without a driver.
typedef CUresult (*TensorMapEncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                      const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                      CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
TensorMapEncodeFn tensor_map_encoder()
{
    static const TensorMapEncodeFn fn = [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q = cudaDriverEntryPointSymbolNotFound;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess) { cudaGetLastError(); p = nullptr; }
        return reinterpret_cast<TensorMapEncodeFn>(p);
    }();
    return fn;
}

// The planes pos, hp, stairs, tick, episode as one u32[5][games] array, if the caller laid them out that way
// (one allocation, common pitch, see orx.h); false: the kernel moves each plane slice on its own.
// Encoding a descriptor costs the host a driver call, so the last few (base, pitch, games) are kept: a state is
// ticked over and over with the same planes.
bool planes5_map(const Params& P, unsigned int n_tiles, CUtensorMap* map)
{
    const char* p0 = reinterpret_cast<const char*>(P.pos);
    const ptrdiff_t pitch = reinterpret_cast<const char*>(P.hp) - p0;
    const int64_t games = (int64_t)n_tiles * kTile;
    if (pitch < games * 4 || (pitch & 15) != 0 || !aligned(p0, 16)) return false;
    if (reinterpret_cast<const char*>(P.stairs) - p0 != 2 * pitch || reinterpret_cast<const char*>(P.tick) - p0 != 3 * pitch ||
        reinterpret_cast<const char*>(P.episode) - p0 != 4 * pitch) return false;
    struct Entry { const char* base; ptrdiff_t pitch; int64_t games; int dev; CUtensorMap map; };
    constexpr int kEntries = 64;
    static std::mutex mu;
    static Entry cache[kEntries];
    static unsigned int used = 0, next = 0;
    int dev = 0;
    cudaGetDevice(&dev);
    {
        std::lock_guard<std::mutex> lk(mu);
        for (unsigned int k = 0; k < used; ++k)
            if (cache[k].base == p0 && cache[k].pitch == pitch && cache[k].games == games && cache[k].dev == dev) { *map = cache[k].map; return true; }
    }
    const TensorMapEncodeFn encode = tensor_map_encoder();
    if (encode == nullptr) return false;
    const cuuint64_t dims[2] = {(cuuint64_t)games, 5};
    const cuuint64_t strides[1] = {(cuuint64_t)pitch};
    const cuuint32_t box[2] = {(cuuint32_t)kTile, 5}, estr[2] = {1, 1};
    if (encode(map, CU_TENSOR_MAP_DATA_TYPE_UINT32, 2, const_cast<char*>(p0), dims, strides, box, estr,
               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) return false;
    std::lock_guard<std::mutex> lk(mu);
    Entry& e = cache[next];
    e.base = p0; e.pitch = pitch; e.games = games; e.dev = dev; e.map = *map;
    next = (next + 1) % kEntries;
    if (used < kEntries) ++used;
    return true;
}

// How the tiles of one launch are handed out and how consecutive launches on a state are ordered (orx_pipe.cuh).
struct TileCtl {
    unsigned int* counter;   // grid-wait mode: dynamic tile counter (sched[0]) or NULL = static stride
    unsigned int* flags;     // flag mode: {next, serving} per tile (sched + ORX_SCHED_HEADER_WORDS) or NULL
    int tiles_per_cta;       // flag mode: target number of tiles per CTA (smaller grids let consecutive launches share the SMs)
    int no_tensor_map;
};

int device_sms(int dev)
{
    static std::mutex mu;
    static int sms[64] = {0};
    std::lock_guard<std::mutex> lk(mu);
    int& v = sms[dev & 63];
    if (v == 0 && cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) { cudaGetLastError(); v = 148; }
    return v;
}

// Tiles per CTA in flag mode: a function of the state's tile count, the device and the caller's override ONLY (never
// of the kernel variant), because CTA b must own the same run of tiles in every launch on the state. Default: about a
// QUARTER of a CTA per SM and launch -- in this mode a launch never has the machine to itself (several launches are in
// flight, that is the point), the next launch starts when EVERY CTA of this one has drawn its ticket, long runs keep
// each CTA's pipeline full and amortise the hand-over protocol (measured at 2^17 games = 512 tiles, us per step with
// 2 / 3 / 4 / 5 / 7 / 10 / 14 tiles per CTA: 3.45 / 2.70 / 2.33 / 2.11 / 1.94 / 1.84 / 1.82; at 2^20 with 10 / 14 / 20 /
// 28 / 32: 11.3 / 10.7 / 10.2 / 9.89 / 9.87, profiles/r02_ab_ticket_first.log) -- and at most kBitsMaxTiles tiles (a CTA
// holds the bit-packed commands of its whole run, and the run is one hand-over chunk).
unsigned int flag_tiles_per_cta(unsigned int n_tiles, int sms, int override_tiles)
{
    const unsigned int target_ctas = (unsigned int)(sms > 3 ? sms / 4 : 1);
    unsigned int m = override_tiles > 0 ? (unsigned int)override_tiles : (n_tiles + target_ctas - 1) / target_ctas;
    if (m < 1) m = 1;
    if (m > kBitsMaxTiles) m = kBitsMaxTiles;
    return m;
}

// Flag mode is a property of the STATE (its scratch size and its number of tiles), never of the kernel variant:
// every pipelined launch on a state must make the same choice, because a flag-mode launch does not wait for an
// earlier grid-wait-mode launch on the same planes.
TileCtl tile_ctl(const OrxConfig* cfg, const OrxState* st, unsigned int n_tiles)
{
    TileCtl c{nullptr, nullptr, 0, (cfg->path_flags & ORX_PATH_NO_TENSOR_MAP) != 0};
    if (st->sched == nullptr || !aligned(st->sched, 16)) return c;
    int dev = 0;
    cudaGetDevice(&dev);
    const int sms = device_sms(dev);
    const uint64_t need = (uint64_t)ORX_SCHED_HEADER_WORDS + 2ull * n_tiles;
    // Flag mode is the caller's choice (ORX_PATH_TILE_FLAGS): it pays about 2.5 us of latency per launch (tickets,
    // acquire, completion + release fence per chunk) for not having a grid-wide boundary between launches, which is a
    // gain for ticks enqueued back to back and a loss for a tick that runs alone (profiles/r02_batch_sweep.json).
    // Measured up to kFlagModeMaxTiles tiles (2^22 games: 19.6 against 21.2 us per step at 2^21, 38.9 against 39.8 at
    // 2^22, profiles/r02_ab_ticket_first.log); above that a launch is long enough for the boundary not to matter and the
    // batch is ticked grid by grid.
    if ((cfg->path_flags & ORX_PATH_TILE_FLAGS) && st->sched_words >= need && n_tiles <= kFlagModeMaxTiles) {
        c.flags = st->sched + ORX_SCHED_HEADER_WORDS;
        c.tiles_per_cta = (int)flag_tiles_per_cta(n_tiles, sms, (int)((cfg->path_flags >> ORX_PATH_TILES_PER_CTA_SHIFT) & 255u));
        return c;
    }
    if (!(cfg->path_flags & ORX_PATH_STATIC_TILES) && st->sched_words >= ORX_SCHED_HEADER_WORDS) c.counter = st->sched;
    return c;
}

template <int DGEN, int CMD, bool OBS, bool TICK, bool EV, int NPC, bool FLAGGED>
int launch_pipe_impl(const Params& P, const void* mv, uint8_t* result, unsigned int n_tiles, size_t tiles_bytes, const TileCtl& ctl,
                     int16_t* obs, int obs_radius, cudaStream_t s, uint2* events, int bots)
{
    constexpr bool BITS = CMD == CMD_BITS;
    const size_t smem = pipe_smem_bytes<OBS, EV, NPC, BITS>((int)tiles_bytes);
    auto kernel = k_step_pipe<DGEN, CMD, OBS, TICK, EV, NPC, FLAGGED>;
    // Launch geometry depends only on (device, kernel, smem): looked up once per process, the occupancy query
    // costs more than the launch itself. (A cache of device properties, not state; one per kernel instantiation.)
    // The dynamic shared-memory limit of the kernel is raised ONCE per device to the most any configuration
    // needs (stages + the largest fixed map), so concurrent host threads with different maps never lower it
    // under each other. Guarded by a mutex: host threads driving different GPUs may call in concurrently.
    struct Geom { int dev; size_t smem; int per_sm; };
    static std::mutex cache_mu;
    static Geom cache[16];
    static unsigned int cache_used = 0, cache_next = 0;
    static bool attr_set[64] = {false};
    int dev = 0;
    cudaGetDevice(&dev);
    int per_sm = 0;
    {
        std::lock_guard<std::mutex> lk(cache_mu);
        if (!attr_set[dev & 63]) {
            const size_t most = pipe_smem_bytes<OBS, EV, NPC, BITS>(kMaxFixedTiles);
            if (most > 48 * 1024) {
                const cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)most);
                if (e != cudaSuccess) return cuda_fail(e);
            }
            attr_set[dev & 63] = true;
        }
        for (unsigned int k = 0; k < cache_used; ++k)
            if (cache[k].dev == dev && cache[k].smem == smem) per_sm = cache[k].per_sm;
        if (per_sm == 0) {
            if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, kPipeThreads, smem) != cudaSuccess || per_sm < 1) { cudaGetLastError(); per_sm = 1; }
            cache[cache_next] = Geom{dev, smem, per_sm};
            cache_next = (cache_next + 1) % 16;
            if (cache_used < 16) ++cache_used;
        }
    }
    const unsigned int sms = (unsigned int)device_sms(dev);
    unsigned int grid = sms * (unsigned int)per_sm;       // persistent: every CTA resident
    if (grid > n_tiles) grid = n_tiles;
    unsigned int tiles_per_cta = 0;
    if (ctl.flags != nullptr) {
        // Flag mode: contiguous runs of tile_ctl's length, whatever this variant's occupancy (CTAs beyond the
        // resident ones simply start later; nothing in a launch waits for a CTA of the same launch)
        tiles_per_cta = (unsigned int)ctl.tiles_per_cta;
        grid = (n_tiles + tiles_per_cta - 1) / tiles_per_cta;
    } else if (BITS) {      // contiguous runs, every CTA at least one tile; the caller keeps n_tiles <= kBitsMaxTiles * grid
        tiles_per_cta = (n_tiles + grid - 1) / grid;
        grid = (n_tiles + tiles_per_cta - 1) / tiles_per_cta;
    }
    if (BITS && tiles_per_cta > kBitsMaxTiles) return ORX_ERR_UNSUPPORTED;
    alignas(64) CUtensorMap planes5;
    int use_map = 0;
    if (!ctl.no_tensor_map) use_map = planes5_map(P, n_tiles, &planes5) ? 1 : 0;
    if (!use_map) memset(&planes5, 0, sizeof(planes5));
    cudaLaunchConfig_t lc = {};
    lc.gridDim = dim3(grid); lc.blockDim = dim3(kPipeThreads); lc.dynamicSmemBytes = smem; lc.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;     // PDL, see k_step_pipe
    at[0].val.programmaticStreamSerializationAllowed = 1;
    lc.attrs = at; lc.numAttrs = ORX_PIPE_PDL ? 1 : 0;
#ifdef ORX_PIPE_TRACE
    static unsigned int trace_slot = 0;
    const cudaError_t e = cudaLaunchKernelEx(&lc, kernel, P, planes5, use_map, mv, result, n_tiles, BITS ? nullptr : ctl.counter, ctl.flags, obs, obs_radius, events, bots, tiles_per_cta, trace_slot++);
#else
    const cudaError_t e = cudaLaunchKernelEx(&lc, kernel, P, planes5, use_map, mv, result, n_tiles, BITS ? nullptr : ctl.counter, ctl.flags, obs, obs_radius, events, bots, tiles_per_cta);
#endif
    return e == cudaSuccess ? launch_done() : cuda_fail(e);
}

// Flag mode and grid-wait mode are separate kernel instantiations (see k_step_pipe).
template <int DGEN, int CMD, bool OBS, bool TICK, bool EV = false, int NPC = 0>
int launch_pipe(const Params& P, const void* mv, uint8_t* result, unsigned int n_tiles, size_t tiles_bytes, const TileCtl& ctl,
                int16_t* obs, int obs_radius, cudaStream_t s, uint2* events = nullptr, int bots = 0)
{
    return ctl.flags != nullptr ? launch_pipe_impl<DGEN, CMD, OBS, TICK, EV, NPC, true>(P, mv, result, n_tiles, tiles_bytes, ctl, obs, obs_radius, s, events, bots)
                                : launch_pipe_impl<DGEN, CMD, OBS, TICK, EV, NPC, false>(P, mv, result, n_tiles, tiles_bytes, ctl, obs, obs_radius, s, events, bots);
}

// The tick with NPC slots: one instantiation per slot count (the slot loops unroll, the stage holds exactly the slots
// in use, and with few slots more stages fit).
template <int DGEN, int CMD>
int launch_npc_pipe(int n_npc, const Params& P, const void* mv, uint8_t* result, unsigned int n_tiles, size_t tiles_bytes,
                    const TileCtl& ctl, cudaStream_t s)
{
    switch (n_npc) {
#define ORX_NPC_CASE(N) case N: return launch_pipe<DGEN, CMD, false, true, false, N>(P, mv, result, n_tiles, tiles_bytes, ctl, nullptr, -1, s);
        ORX_NPC_CASE(1) ORX_NPC_CASE(2) ORX_NPC_CASE(3) ORX_NPC_CASE(4) ORX_NPC_CASE(5) ORX_NPC_CASE(6) ORX_NPC_CASE(7) ORX_NPC_CASE(8)
#undef ORX_NPC_CASE
    }
    return ORX_ERR_BAD_ARG;
}

template <bool OBS, bool EV = false>
int launch_tick_pipe(bool empty, int packed, const Params& P, const void* mv, uint8_t* result, unsigned int n_tiles, size_t smem,
                     const TileCtl& sched, int16_t* obs, int obs_radius, cudaStream_t s, uint2* events = nullptr, int bots = 0)
{
    if constexpr (!EV) {
        if (bots != 0 && !packed)
            return empty ? launch_pipe<ORX_DGEN_EMPTY, CMD_BYTES_BOTS, OBS, true>(P, mv, result, n_tiles, 0, sched, obs, obs_radius, s, nullptr, bots)
                         : launch_pipe<ORX_DGEN_FIXED, CMD_BYTES_BOTS, OBS, true>(P, mv, result, n_tiles, smem, sched, obs, obs_radius, s, nullptr, bots);
    }
    if (packed) return empty ? launch_pipe<ORX_DGEN_EMPTY, CMD_NIBBLES, OBS, true, EV>(P, mv, result, n_tiles, 0, sched, obs, obs_radius, s, events)
                             : launch_pipe<ORX_DGEN_FIXED, CMD_NIBBLES, OBS, true, EV>(P, mv, result, n_tiles, smem, sched, obs, obs_radius, s, events);
    return empty ? launch_pipe<ORX_DGEN_EMPTY, CMD_BYTES, OBS, true, EV>(P, mv, result, n_tiles, 0, sched, obs, obs_radius, s, events)
                 : launch_pipe<ORX_DGEN_FIXED, CMD_BYTES, OBS, true, EV>(P, mv, result, n_tiles, smem, sched, obs, obs_radius, s, events);
}

// One tick; packed = 0: moves uint8[n][2], packed = 1: uint8[n] with p1 | p2 << 4.
// obs != NULL: also writes the observations of the resulting state (orx_step_observe).
int step_impl(const OrxConfig* cfg, const OrxState* st, const uint8_t* moves, uint8_t* result,
              OrxEvent* events, int64_t n, uint64_t game_id_base, void* cuda_stream, int packed,
              int16_t* obs = nullptr, int obs_radius = -1, int bots = 0)
{
    const int rc = check_common(cfg, st, n);
    if (rc != ORX_OK) return rc;
    if (!ids_ok(game_id_base, n)) return ORX_ERR_BAD_ARG;
    if (moves == nullptr || result == nullptr || (!packed && !aligned(moves, 2)) || (events && !aligned(events, 8))) return ORX_ERR_BAD_ARG;
    if (n == 0) return ORX_OK;
    const Params P = make_params(cfg, st, n, game_id_base);
    cudaStream_t s = static_cast<cudaStream_t>(cuda_stream);
    const size_t smem = tiles_smem(cfg);
    const size_t mv_stride = packed ? 1 : 2;        // command bytes per game
    uint2* ev = reinterpret_cast<uint2*>(events);
    const int max_ev = orx_max_events(cfg);
    // Hot variants (plain, with observations, with the event log, or with NPC slots): persistent
    // TMA-pipelined kernel over the full 256-game tiles, the simple kernel for a ragged tail (< 256 games).
    const bool no_flat = st->flat == nullptr;      // a bonus plane (Modifier seam) is read by the one-thread-per-game kernels only
    const bool npc_pipe = no_flat && cfg->n_npc > 0 && ev == nullptr && obs == nullptr && bots == 0 && aligned(st->npc_pos, 16) &&
                          aligned(st->npc_hp, 16) && aligned(st->npc_depth, 16) && !(cfg->path_flags & ORX_PATH_NO_NPC_PIPE);
    if (npc_pipe && n >= kTile && pipe_aligned(st, moves, result)) {
        const unsigned int n_tiles = (unsigned int)(n / kTile);
        const int64_t n_body = (int64_t)n_tiles * kTile;
        const bool empty = cfg->dgen_kind == ORX_DGEN_EMPTY;
        const TileCtl sched = tile_ctl(cfg, st, n_tiles);
        int rc2;
        if (packed) rc2 = empty ? launch_npc_pipe<ORX_DGEN_EMPTY, CMD_NIBBLES>(cfg->n_npc, P, moves, result, n_tiles, 0, sched, s)
                                : launch_npc_pipe<ORX_DGEN_FIXED, CMD_NIBBLES>(cfg->n_npc, P, moves, result, n_tiles, smem, sched, s);
        else rc2 = empty ? launch_npc_pipe<ORX_DGEN_EMPTY, CMD_BYTES>(cfg->n_npc, P, moves, result, n_tiles, 0, sched, s)
                         : launch_npc_pipe<ORX_DGEN_FIXED, CMD_BYTES>(cfg->n_npc, P, moves, result, n_tiles, smem, sched, s);
        if (rc2 != ORX_OK || n_body == n) return rc2;
        const Params T = offset_params(P, n_body, n - n_body);
        const int tgrid = grid_for(n - n_body);
        const uint8_t* tail = moves + (size_t)n_body * mv_stride;
        if (empty) k_step<ORX_DGEN_EMPTY, true, false><<<tgrid, kThreads, 0, s>>>(T, tail, result + n_body, nullptr, max_ev, packed);
        else k_step<ORX_DGEN_FIXED, true, false><<<tgrid, kThreads, smem, s>>>(T, tail, result + n_body, nullptr, max_ev, packed);
        return launch_done();
    }
    const bool ev_pipe = ev != nullptr && obs == nullptr && bots == 0 && aligned(ev, 16) && !(cfg->path_flags & ORX_PATH_NO_EVENT_PIPE);
    if (no_flat && (ev == nullptr || ev_pipe) && cfg->n_npc == 0 && n >= kTile && pipe_aligned(st, moves, result)) {
        const unsigned int n_tiles = (unsigned int)(n / kTile);
        const int64_t n_body = (int64_t)n_tiles * kTile;
        const bool empty = cfg->dgen_kind == ORX_DGEN_EMPTY;
        const TileCtl sched = tile_ctl(cfg, st, n_tiles);
        const int rc2 = obs != nullptr ? launch_tick_pipe<true>(empty, packed, P, moves, result, n_tiles, smem, sched, obs, obs_radius, s, nullptr, bots)
                        : ev_pipe      ? launch_tick_pipe<false, true>(empty, packed, P, moves, result, n_tiles, smem, sched, nullptr, -1, s, ev)
                                       : launch_tick_pipe<false>(empty, packed, P, moves, result, n_tiles, smem, sched, nullptr, -1, s, nullptr, bots);
        if (rc2 != ORX_OK || n_body == n) return rc2;
        const Params T = offset_params(P, n_body, n - n_body);
        const int tgrid = grid_for(n - n_body);
        const uint8_t* tail = moves + (size_t)n_body * mv_stride;
        if (ev_pipe) {
            uint2* tev = ev + (size_t)n_body * max_ev;
            if (empty) k_step<ORX_DGEN_EMPTY, false, true><<<tgrid, kThreads, 0, s>>>(T, tail, result + n_body, tev, max_ev, packed);
            else k_step<ORX_DGEN_FIXED, false, true><<<tgrid, kThreads, smem, s>>>(T, tail, result + n_body, tev, max_ev, packed);
        } else if (empty) k_step<ORX_DGEN_EMPTY, false, false><<<tgrid, kThreads, 0, s>>>(T, tail, result + n_body, nullptr, max_ev, packed, bots);
        else k_step<ORX_DGEN_FIXED, false, false><<<tgrid, kThreads, smem, s>>>(T, tail, result + n_body, nullptr, max_ev, packed, bots);
        if (obs != nullptr) k_observe<<<tgrid, kThreads, 0, s>>>(T, obs + (size_t)n_body * 2 * ORX_OBS_LEN, obs_radius);
        return launch_done();
    }
    const int grid = grid_for(n);
    return dispatch_dgen_npc(cfg, [&]<int DGEN, bool NPC>() {
        if (ev != nullptr) k_step<DGEN, NPC, true><<<grid, kThreads, smem, s>>>(P, moves, result, ev, max_ev, packed, bots);
        else k_step<DGEN, NPC, false><<<grid, kThreads, smem, s>>>(P, moves, result, nullptr, max_ev, packed, bots);
        if (obs != nullptr) k_observe<<<grid, kThreads, 0, s>>>(P, obs, obs_radius);
        return launch_done();
    });
}

// One tick with the bit-packed streams: cmd5 = 5 bits per game in, res2 = 2 bits per game out (orx.h). Device
// pointers, or device-mapped pinned host memory (the kernel then reads / writes it over PCIe, one transaction per
// CTA each way).
int step_bits_impl(const OrxConfig* cfg, const OrxState* st, const uint8_t* cmd5, uint8_t* res2, int64_t n,
                   uint64_t game_id_base, void* cuda_stream)
{
    const int rc = check_common(cfg, st, n);
    if (rc != ORX_OK) return rc;
    if (!ids_ok(game_id_base, n)) return ORX_ERR_BAD_ARG;
    if (cmd5 == nullptr || res2 == nullptr || !aligned(cmd5, 16) || !aligned(res2, 16)) return ORX_ERR_BAD_ARG;
    if (cfg->n_npc != 0) return ORX_ERR_UNSUPPORTED;
    if (n == 0) return ORX_OK;
    const Params P = make_params(cfg, st, n, game_id_base);
    cudaStream_t s = static_cast<cudaStream_t>(cuda_stream);
    const size_t smem = tiles_smem(cfg);
    const bool empty = cfg->dgen_kind == ORX_DGEN_EMPTY;
    const unsigned int n_tiles = (unsigned int)(n / kTile);
    const int64_t n_body = (int64_t)n_tiles * kTile;
    if (n_tiles > 0 && st->flat == nullptr && pipe_aligned(st, cmd5, res2)) {
        const TileCtl ctl = tile_ctl(cfg, st, n_tiles);
        int dev = 0;
        cudaGetDevice(&dev);
        // a CTA holds the commands / results of at most kBitsMaxTiles tiles: batches beyond that many tiles per
        // resident CTA (only possible in grid-wait mode; flag mode is bounded the same way) go chunk by chunk
        const unsigned int max_tiles = ctl.flags != nullptr ? n_tiles : (unsigned int)kBitsMaxTiles * (unsigned int)device_sms(dev);
        for (unsigned int t0 = 0; t0 < n_tiles; t0 += max_tiles) {
            const unsigned int nt = n_tiles - t0 < max_tiles ? n_tiles - t0 : max_tiles;
            const Params C = t0 == 0 && nt == n_tiles ? P : offset_params(P, (int64_t)t0 * kTile, (int64_t)nt * kTile);
            TileCtl c = ctl;
            if (c.flags != nullptr) c.flags += 2 * (size_t)(t0 / kChunk);
            const uint8_t* cm = cmd5 + (size_t)t0 * kCmdBitsTile;
            uint8_t* rs = res2 + (size_t)t0 * kResBitsTile;
            const int rc2 = empty ? launch_pipe<ORX_DGEN_EMPTY, CMD_BITS, false, true>(C, cm, rs, nt, 0, c, nullptr, -1, s)
                                  : launch_pipe<ORX_DGEN_FIXED, CMD_BITS, false, true>(C, cm, rs, nt, smem, c, nullptr, -1, s);
            if (rc2 != ORX_OK) return rc2;
        }
        if (n_body == n) return ORX_OK;
        const Params T = offset_params(P, n_body, n - n_body);
        const int tgrid = grid_for(n - n_body);
        const uint8_t* cm = cmd5 + (size_t)n_tiles * kCmdBitsTile;
        uint8_t* rs = res2 + (size_t)n_tiles * kResBitsTile;
        if (empty) k_step<ORX_DGEN_EMPTY, false, false><<<tgrid, kThreads, 0, s>>>(T, cm, rs, nullptr, ORX_MAX_EVENTS_BASE, ORX_FMT_BITS);
        else k_step<ORX_DGEN_FIXED, false, false><<<tgrid, kThreads, smem, s>>>(T, cm, rs, nullptr, ORX_MAX_EVENTS_BASE, ORX_FMT_BITS);
        return launch_done();
    }
    const int grid = grid_for(n);
    if (empty) k_step<ORX_DGEN_EMPTY, false, false><<<grid, kThreads, 0, s>>>(P, cmd5, res2, nullptr, ORX_MAX_EVENTS_BASE, ORX_FMT_BITS);
    else k_step<ORX_DGEN_FIXED, false, false><<<grid, kThreads, smem, s>>>(P, cmd5, res2, nullptr, ORX_MAX_EVENTS_BASE, ORX_FMT_BITS);
    return launch_done();
}

int step_host_bits_impl(const OrxConfig* cfg, const OrxState* st, const uint8_t* cmd5_host, uint8_t* res2_host,
                        uint8_t* cmd5_dev, uint8_t* res2_dev, int64_t n, uint64_t game_id_base, void* cuda_stream)
{
    if (cmd5_host == nullptr || res2_host == nullptr || cfg == nullptr) return ORX_ERR_BAD_ARG;
    cudaStream_t s = static_cast<cudaStream_t>(cuda_stream);
    void *cm = nullptr, *rs = nullptr;
    if (!(cfg->path_flags & ORX_PATH_HOST_STAGED) && host_mapped(cmd5_host, &cm) && host_mapped(res2_host, &rs))
        return step_bits_impl(cfg, st, static_cast<const uint8_t*>(cm), static_cast<uint8_t*>(rs), n, game_id_base, cuda_stream);
    if (cmd5_dev == nullptr || res2_dev == nullptr || n < 0) return ORX_ERR_BAD_ARG;
    cudaError_t e = cudaMemcpyAsync(cmd5_dev, cmd5_host, ORX_CMD5_BYTES(n), cudaMemcpyHostToDevice, s);
    if (e != cudaSuccess) return cuda_fail(e);
    const int rc = step_bits_impl(cfg, st, cmd5_dev, res2_dev, n, game_id_base, cuda_stream);
    if (rc != ORX_OK) return rc;
    e = cudaMemcpyAsync(res2_host, res2_dev, ORX_RES2_BYTES(n), cudaMemcpyDeviceToHost, s);
    return e == cudaSuccess ? ORX_OK : cuda_fail(e);
}

int step_host_impl(const OrxConfig* cfg, const OrxState* st, const uint8_t* moves_host,
                   uint8_t* result_host, uint8_t* moves_dev, uint8_t* result_dev, int64_t n,
                   uint64_t game_id_base, void* cuda_stream, int packed)
{
    if (moves_host == nullptr || result_host == nullptr) return ORX_ERR_BAD_ARG;
    const int rc = check_common(cfg, st, n);
    if (rc != ORX_OK) return rc;
    if (n == 0) return ORX_OK;
    cudaStream_t s = static_cast<cudaStream_t>(cuda_stream);
    // Pinned (page-locked, UVA-mapped) host buffers are read and written by the tick kernel itself:
    // the TMA producer pulls each tile's commands over PCIe next to its HBM planes and the results
    // stream back the same way, so there is no separate copy launch and the transfer overlaps the
    // compute tile by tile. Pageable buffers fall back to staged cudaMemcpyAsync copies.
    void *mv_map = nullptr, *res_map = nullptr;
    if (!(cfg->path_flags & ORX_PATH_HOST_STAGED) && host_mapped(moves_host, &mv_map) && host_mapped(result_host, &res_map))
        return step_impl(cfg, st, static_cast<const uint8_t*>(mv_map), static_cast<uint8_t*>(res_map), nullptr, n,
                         game_id_base, cuda_stream, packed);
    if (moves_dev == nullptr || result_dev == nullptr) return ORX_ERR_BAD_ARG;
    cudaError_t e = cudaMemcpyAsync(moves_dev, moves_host, (size_t)n * (packed ? 1 : 2), cudaMemcpyHostToDevice, s);
    if (e != cudaSuccess) return cuda_fail(e);
    const int rs = step_impl(cfg, st, moves_dev, result_dev, nullptr, n, game_id_base, cuda_stream, packed);
    if (rs != ORX_OK) return rs;
    e = cudaMemcpyAsync(result_host, result_dev, (size_t)n, cudaMemcpyDeviceToHost, s);
    return e == cudaSuccess ? ORX_OK : cuda_fail(e);
}

// The *_sync entry points tick in grid-wait mode whatever the state's scratch allows: the call returns only when the
// stream has drained, so no later launch can overlap this one, and the hand-over words would be a microsecond or two
// of overhead for nothing. Safe next to flag-mode launches on the same state: a grid-wait launch waits for everything
// before it, and everything after it is enqueued after it has completed.
OrxConfig without_tile_flags(const OrxConfig* cfg)
{
    OrxConfig c = *cfg;
    c.path_flags &= ~ORX_PATH_TILE_FLAGS;
    return c;
}

}  // namespace

extern "C" {

int orx_abi_version(void) { return ORX_ABI_VERSION; }

#ifdef ORX_PIPE_TRACE
// Tuning builds only (not declared in orx.h): copies the per-CTA time stamps of the last 16 tick launches.
int orx_debug_trace(unsigned long long* host_out)
{
    const cudaError_t e = cudaMemcpyFromSymbol(host_out, orx::g_trace, sizeof(orx::g_trace));
    return e == cudaSuccess ? ORX_OK : ORX_ERR_CUDA_BASE - (int)e;
}
#endif

const char* orx_strerror(int code)
{
    if (code == ORX_OK) return "ok";
    if (code == ORX_ERR_BAD_ARG) return "bad argument (null/misaligned pointer, config out of range, or struct_size mismatch)";
    if (code == ORX_ERR_UNSUPPORTED) return "unsupported configuration";
    if (code <= ORX_ERR_CUDA_BASE) return cudaGetErrorString((cudaError_t)(ORX_ERR_CUDA_BASE - code));
    return "unknown error";
}

size_t orx_sched_words(int64_t n)
{
    return (size_t)ORX_SCHED_HEADER_WORDS + 2 * (size_t)((n > 0 ? n : 0) / kTile);
}

size_t orx_state_bytes(const OrxConfig* cfg)
{
    if (cfg == nullptr) return 0;
    // pos 4 + hp 4 + depth 8 + stairs 4 + tick 4 + episode 4 + status 1; NPC slot: pos 2 + hp 2 + depth 4
    return 29 + (size_t)(cfg->n_npc > 0 ? cfg->n_npc : 0) * 8;
}

int orx_max_events(const OrxConfig* cfg)
{
    return ORX_MAX_EVENTS_BASE + (cfg != nullptr && cfg->n_npc > 0 ? cfg->n_npc : 0);
}

int orx_event_count_add(const OrxEvent* events, int max_events, unsigned long long* order, int64_t n, void* cuda_stream)
{
    if (events == nullptr || order == nullptr || n < 0 || n > kMaxGamesPerCall || max_events < 1 ||
        max_events > ORX_MAX_EVENTS_BASE + ORX_MAX_NPC) return ORX_ERR_BAD_ARG;
    if (!aligned(order, 8) || !aligned(events, max_events == ORX_MAX_EVENTS_BASE ? 16 : 8)) return ORX_ERR_BAD_ARG;
    if (n == 0) return ORX_OK;
    k_event_count_add<<<grid_for(n), kThreads, 0, static_cast<cudaStream_t>(cuda_stream)>>>(
        reinterpret_cast<const uint2*>(events), max_events, order, (unsigned int)n);
    return launch_done();
}

int orx_reset(const OrxConfig* cfg, const OrxState* st, const uint8_t* mask, int bump_episode,
              int64_t n, uint64_t game_id_base, void* cuda_stream)
{
    const int rc = check_common(cfg, st, n);
    if (rc != ORX_OK) return rc;
    if (!ids_ok(game_id_base, n)) return ORX_ERR_BAD_ARG;
    if (n == 0) return ORX_OK;
    const Params P = make_params(cfg, st, n, game_id_base);
    cudaStream_t s = static_cast<cudaStream_t>(cuda_stream);
    const int grid = grid_for(n);
    // The scheduler words are zero between launches; a reset re-establishes that after an aborted run.
    if (st->sched != nullptr && aligned(st->sched, 4) && st->sched_words > 0) {
        const cudaError_t e = cudaMemsetAsync(st->sched, 0, (size_t)st->sched_words * sizeof(unsigned int), s);
        if (e != cudaSuccess) return cuda_fail(e);
    }
    return dispatch_dgen_npc(cfg, [&]<int DGEN, bool NPC>() {
        k_reset<DGEN, NPC><<<grid, kThreads, 0, s>>>(P, mask, bump_episode);
        return launch_done();
    });
}

int orx_step(const OrxConfig* cfg, const OrxState* st, const uint8_t* moves, uint8_t* result,
             OrxEvent* events, int64_t n, uint64_t game_id_base, void* cuda_stream)
{
    return step_impl(cfg, st, moves, result, events, n, game_id_base, cuda_stream, 0);
}

int orx_step_packed(const OrxConfig* cfg, const OrxState* st, const uint8_t* cmds, uint8_t* result,
                    OrxEvent* events, int64_t n, uint64_t game_id_base, void* cuda_stream)
{
    return step_impl(cfg, st, cmds, result, events, n, game_id_base, cuda_stream, 1);
}

int orx_step_host(const OrxConfig* cfg, const OrxState* st, const uint8_t* moves_host,
                  uint8_t* result_host, uint8_t* moves_dev, uint8_t* result_dev, int64_t n,
                  uint64_t game_id_base, void* cuda_stream)
{
    return step_host_impl(cfg, st, moves_host, result_host, moves_dev, result_dev, n, game_id_base, cuda_stream, 0);
}

int orx_step_host_sync(const OrxConfig* cfg_in, const OrxState* st, const uint8_t* moves_host,
                       uint8_t* result_host, uint8_t* moves_dev, uint8_t* result_dev, int64_t n,
                       uint64_t game_id_base, void* cuda_stream)
{
    if (cfg_in == nullptr) return ORX_ERR_BAD_ARG;
    const OrxConfig c = without_tile_flags(cfg_in);
    const OrxConfig* cfg = &c;
    const int rc = step_host_impl(cfg, st, moves_host, result_host, moves_dev, result_dev, n, game_id_base, cuda_stream, 0);
    if (rc != ORX_OK) return rc;
    const cudaError_t e = cudaStreamSynchronize(static_cast<cudaStream_t>(cuda_stream));
    return e == cudaSuccess ? ORX_OK : cuda_fail(e);
}

int orx_step_host_packed(const OrxConfig* cfg, const OrxState* st, const uint8_t* cmds_host,
                         uint8_t* result_host, uint8_t* cmds_dev, uint8_t* result_dev, int64_t n,
                         uint64_t game_id_base, void* cuda_stream)
{
    return step_host_impl(cfg, st, cmds_host, result_host, cmds_dev, result_dev, n, game_id_base, cuda_stream, 1);
}

int orx_step_host_packed_sync(const OrxConfig* cfg_in, const OrxState* st, const uint8_t* cmds_host,
                              uint8_t* result_host, uint8_t* cmds_dev, uint8_t* result_dev, int64_t n,
                              uint64_t game_id_base, void* cuda_stream)
{
    if (cfg_in == nullptr) return ORX_ERR_BAD_ARG;
    const OrxConfig c = without_tile_flags(cfg_in);
    const OrxConfig* cfg = &c;
    const int rc = step_host_impl(cfg, st, cmds_host, result_host, cmds_dev, result_dev, n, game_id_base, cuda_stream, 1);
    if (rc != ORX_OK) return rc;
    const cudaError_t e = cudaStreamSynchronize(static_cast<cudaStream_t>(cuda_stream));
    return e == cudaSuccess ? ORX_OK : cuda_fail(e);
}

int orx_step_bits(const OrxConfig* cfg, const OrxState* st, const uint8_t* cmd5, uint8_t* res2, int64_t n,
                  uint64_t game_id_base, void* cuda_stream)
{
    return step_bits_impl(cfg, st, cmd5, res2, n, game_id_base, cuda_stream);
}

int orx_step_host_bits(const OrxConfig* cfg, const OrxState* st, const uint8_t* cmd5_host, uint8_t* res2_host,
                       uint8_t* cmd5_dev, uint8_t* res2_dev, int64_t n, uint64_t game_id_base, void* cuda_stream)
{
    return step_host_bits_impl(cfg, st, cmd5_host, res2_host, cmd5_dev, res2_dev, n, game_id_base, cuda_stream);
}

int orx_step_host_bits_sync(const OrxConfig* cfg_in, const OrxState* st, const uint8_t* cmd5_host, uint8_t* res2_host,
                            uint8_t* cmd5_dev, uint8_t* res2_dev, int64_t n, uint64_t game_id_base, void* cuda_stream)
{
    if (cfg_in == nullptr) return ORX_ERR_BAD_ARG;
    const OrxConfig c = without_tile_flags(cfg_in);
    const OrxConfig* cfg = &c;
    const int rc = step_host_bits_impl(cfg, st, cmd5_host, res2_host, cmd5_dev, res2_dev, n, game_id_base, cuda_stream);
    if (rc != ORX_OK) return rc;
    const cudaError_t e = cudaStreamSynchronize(static_cast<cudaStream_t>(cuda_stream));
    return e == cudaSuccess ? ORX_OK : cuda_fail(e);
}

int orx_bot_moves(const OrxConfig* cfg, const OrxState* st, int bot_p1, int bot_p2,
                  uint8_t* moves, int64_t n, uint64_t game_id_base, void* cuda_stream)
{
    const int rc = check_common(cfg, st, n);
    if (rc != ORX_OK) return rc;
    if (!ids_ok(game_id_base, n)) return ORX_ERR_BAD_ARG;
    if (moves == nullptr || !aligned(moves, 2)) return ORX_ERR_BAD_ARG;
    if (bot_p1 < ORX_BOT_NONE || bot_p1 > ORX_BOT_STAIRCASE || bot_p2 < ORX_BOT_NONE || bot_p2 > ORX_BOT_STAIRCASE) return ORX_ERR_BAD_ARG;
    if (n == 0) return ORX_OK;
    const Params P = make_params(cfg, st, n, game_id_base);
    k_bot_moves<<<grid_for(n), kThreads, 0, static_cast<cudaStream_t>(cuda_stream)>>>(P, bot_p1, bot_p2, moves);
    return launch_done();
}

int orx_rollout(const OrxConfig* cfg, const OrxState* st, int bot_p1, int bot_p2, int n_ticks,
                unsigned long long* stats, int64_t n, uint64_t game_id_base, void* cuda_stream)
{
    const int rc = check_common(cfg, st, n);
    if (rc != ORX_OK) return rc;
    if (!ids_ok(game_id_base, n)) return ORX_ERR_BAD_ARG;
    if (n_ticks < 0 || (stats && !aligned(stats, 8))) return ORX_ERR_BAD_ARG;
    if (bot_p1 < ORX_BOT_NONE || bot_p1 > ORX_BOT_STAIRCASE || bot_p2 < ORX_BOT_NONE || bot_p2 > ORX_BOT_STAIRCASE) return ORX_ERR_BAD_ARG;
    if (n == 0 || n_ticks == 0) return ORX_OK;
    const Params P = make_params(cfg, st, n, game_id_base);
    cudaStream_t s = static_cast<cudaStream_t>(cuda_stream);
    const int grid = grid_for(n);
    const size_t smem = tiles_smem(cfg);
    return dispatch_dgen_npc(cfg, [&]<int DGEN, bool NPC>() {
        k_rollout<DGEN, NPC><<<grid, kThreads, smem, s>>>(P, bot_p1, bot_p2, n_ticks, stats);
        return launch_done();
    });
}

int orx_replay(const OrxConfig* cfg, const OrxState* st, const uint8_t* moves, uint8_t* results, int n_ticks,
               int64_t n, uint64_t game_id_base, void* cuda_stream)
{
    const int rc = check_common(cfg, st, n);
    if (rc != ORX_OK) return rc;
    if (!ids_ok(game_id_base, n)) return ORX_ERR_BAD_ARG;
    if (moves == nullptr || results == nullptr || !aligned(moves, 2) || n_ticks < 0) return ORX_ERR_BAD_ARG;
    if (n == 0 || n_ticks == 0) return ORX_OK;
    const Params P = make_params(cfg, st, n, game_id_base);
    cudaStream_t s = static_cast<cudaStream_t>(cuda_stream);
    const int grid = grid_for(n);
    const size_t smem = tiles_smem(cfg);
    const uint16_t* mv = reinterpret_cast<const uint16_t*>(moves);
    return dispatch_dgen_npc(cfg, [&]<int DGEN, bool NPC>() {
        k_replay<DGEN, NPC><<<grid, kThreads, smem, s>>>(P, mv, results, n_ticks);
        return launch_done();
    });
}

int orx_observe_npc(const OrxConfig* cfg, const OrxState* st, int16_t* npc_obs, int64_t n, void* cuda_stream)
{
    const int rc = check_common(cfg, st, n);
    if (rc != ORX_OK) return rc;
    if (cfg->n_npc == 0 || n == 0) return ORX_OK;
    if (npc_obs == nullptr || !aligned(npc_obs, 8)) return ORX_ERR_BAD_ARG;
    k_observe_npc<<<grid_for(n), kThreads, 0, static_cast<cudaStream_t>(cuda_stream)>>>(make_params(cfg, st, n, 0), reinterpret_cast<uint2*>(npc_obs));
    return launch_done();
}

int orx_observe(const OrxConfig* cfg, const OrxState* st, int16_t* obs, int stairs_radius,
                int64_t n, void* cuda_stream)
{
    const int rc = check_common(cfg, st, n);
    if (rc != ORX_OK) return rc;
    if (obs == nullptr || !aligned(obs, 16)) return ORX_ERR_BAD_ARG;
    if (n == 0) return ORX_OK;
    const Params P = make_params(cfg, st, n, 0);
    cudaStream_t s = static_cast<cudaStream_t>(cuda_stream);
    // Full 256-game tiles stream through the TMA pipeline (planes in, 48 B per game out as bulk stores);
    // a ragged tail and small batches use the simple kernel.
    if (n >= kTile && pipe_aligned(st, obs, obs)) {
        const unsigned int n_tiles = (unsigned int)(n / kTile);
        const int64_t n_body = (int64_t)n_tiles * kTile;
        const TileCtl sched = tile_ctl(cfg, st, n_tiles);
        const int rc2 = launch_pipe<ORX_DGEN_EMPTY, CMD_BYTES, true, false>(P, nullptr, nullptr, n_tiles, 0, sched, obs, stairs_radius, s);
        if (rc2 != ORX_OK || n_body == n) return rc2;
        const Params T = offset_params(P, n_body, n - n_body);
        k_observe<<<grid_for(n - n_body), kThreads, 0, s>>>(T, obs + (size_t)n_body * 2 * ORX_OBS_LEN, stairs_radius);
        return launch_done();
    }
    k_observe<<<grid_for(n), kThreads, 0, s>>>(P, obs, stairs_radius);
    return launch_done();
}

int orx_step_bots(const OrxConfig* cfg, const OrxState* st, const uint8_t* moves, int bot_p1, int bot_p2, uint8_t* result,
                  OrxEvent* events, int16_t* obs, int stairs_radius, int64_t n, uint64_t game_id_base, void* cuda_stream)
{
    if (bot_p1 < ORX_BOT_NONE || bot_p1 > ORX_BOT_STAIRCASE || bot_p2 < ORX_BOT_NONE || bot_p2 > ORX_BOT_STAIRCASE) return ORX_ERR_BAD_ARG;
    if (obs != nullptr && !aligned(obs, 16)) return ORX_ERR_BAD_ARG;
    return step_impl(cfg, st, moves, result, events, n, game_id_base, cuda_stream, 0, obs, stairs_radius, bot_p1 | (bot_p2 << 8));
}

int orx_step_observe(const OrxConfig* cfg, const OrxState* st, const uint8_t* moves, int moves_packed,
                     uint8_t* result, int16_t* obs, int stairs_radius, int64_t n, uint64_t game_id_base,
                     void* cuda_stream)
{
    if (obs == nullptr || !aligned(obs, 16)) return ORX_ERR_BAD_ARG;
    return step_impl(cfg, st, moves, result, nullptr, n, game_id_base, cuda_stream, moves_packed != 0, obs, stairs_radius);
}

}  // extern "C"
