// liborx.so: kernels + the C ABI declared in include/orx.h. sm_100a only.
#include <cuda_runtime.h>
#include <stdio.h>
#include <string.h>

#include "../../include/orx.h"
#include "orx_rules.cuh"

using namespace orx;

namespace {

constexpr int kThreads = 256;
constexpr int kMaxFixedTiles = 16384;   // shared-memory staging of the fixed map

__device__ __forceinline__ const uint8_t* stage_tiles(const Params& P, uint8_t* smem)
{
    // The fixed wall map is shared by every game of the batch: stage it once per CTA.
    const int nt = P.W * P.H;
    for (int t = threadIdx.x; t < nt; t += blockDim.x) smem[t] = P.tiles[t];
    __syncthreads();
    return smem;
}

// ------------------------------------------------------------------ K1: one tick
template <int DGEN, bool NPC, bool EV>
__global__ void __launch_bounds__(kThreads)
k_step(const __grid_constant__ Params P, const uint16_t* __restrict__ moves, uint8_t* __restrict__ result,
       uint2* __restrict__ events, int max_ev)
{
    extern __shared__ uint8_t smem[];
    const uint8_t* tiles = nullptr;
    if (DGEN == ORX_DGEN_FIXED) tiles = stage_tiles(P, smem);
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < P.n; i += stride) {
        const int status = P.status[i];
        EvSink<EV> ev{EV ? events + i * max_ev : nullptr, 0, max_ev};
        if (status != ORX_RESULT_IN_PROGRESS) {   // finished lanes are frozen until reset
            result[i] = (uint8_t)status;
            ev.finish();
            continue;
        }
        Lane L;
        load_lane(P, i, L);
        const uint16_t mv = moves[i];
        Stream s = make_stream(P, i, L.episode);
        const uint4 blk = draw_block(s, DOM_TICK, SUB_MAIN, (uint32_t)L.tick);
        Counters cnt{};
        int res = tick_lane<DGEN, NPC, EV>(P, tiles, L, mv & 255, mv >> 8, blk.z, s, i, ev, cnt);
        ev.finish();
        result[i] = (uint8_t)res;
        if (res != ORX_RESULT_IN_PROGRESS && P.auto_reset) {
            s.episode += 1;
            reset_lane<DGEN, NPC>(P, L, s, i);
            res = ORX_RESULT_IN_PROGRESS;
        }
        store_lane(P, i, L, res);
    }
}

// ------------------------------------------------------------------ K2: masked episode reset
template <int DGEN, bool NPC>
__global__ void __launch_bounds__(kThreads)
k_reset(const __grid_constant__ Params P, const uint8_t* __restrict__ mask, int bump)
{
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < P.n; i += stride) {
        if (mask != nullptr && mask[i] == 0) continue;
        Lane L;
        Stream s = make_stream(P, i, P.episode[i] + (bump ? 1u : 0u));
        reset_lane<DGEN, NPC>(P, L, s, i);
        store_lane(P, i, L, ORX_RESULT_IN_PROGRESS);
    }
}

// ------------------------------------------------------------------ K3: scripted bots
__global__ void __launch_bounds__(kThreads)
k_bot_moves(const __grid_constant__ Params P, int bot1, int bot2, uint8_t* __restrict__ moves)
{
    const long long stride = (long long)gridDim.x * blockDim.x;
    const bool need_rng = bot1 == ORX_BOT_RANDOM || bot2 == ORX_BOT_RANDOM;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < P.n; i += stride) {
        Lane L;
        load_lane(P, i, L);
        uint4 blk = make_uint4(0, 0, 0, 0);
        if (need_rng) blk = draw_block(make_stream(P, i, L.episode), DOM_TICK, SUB_MAIN, (uint32_t)L.tick);
        if (bot1 != ORX_BOT_NONE && bot2 != ORX_BOT_NONE) {
            const int m1 = bot_move(bot1, L.p1, blk.x), m2 = bot_move(bot2, L.p2, blk.y);
            reinterpret_cast<uint16_t*>(moves)[i] = (uint16_t)(m1 | (m2 << 8));
        } else if (bot1 != ORX_BOT_NONE) {
            moves[2 * i] = (uint8_t)bot_move(bot1, L.p1, blk.x);
        } else if (bot2 != ORX_BOT_NONE) {
            moves[2 * i + 1] = (uint8_t)bot_move(bot2, L.p2, blk.y);
        }
    }
}

// ------------------------------------------------------------------ fused rollout (K1+K3, T ticks)
__device__ __forceinline__ unsigned int warp_sum(unsigned int v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

template <int DGEN, bool NPC>
__global__ void __launch_bounds__(kThreads)
k_rollout(const __grid_constant__ Params P, int bot1, int bot2, int n_ticks, unsigned long long* __restrict__ stats)
{
    extern __shared__ uint8_t smem[];
    const uint8_t* tiles = nullptr;
    if (DGEN == ORX_DGEN_FIXED) tiles = stage_tiles(P, smem);
    __shared__ unsigned int s_cnt[7];
    if (threadIdx.x < 7) s_cnt[threadIdx.x] = 0;
    __syncthreads();
    Counters cnt{};
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < P.n; i += stride) {
        int status = P.status[i];
        if (status != ORX_RESULT_IN_PROGRESS) continue;
        Lane L;
        load_lane(P, i, L);
        Stream s = make_stream(P, i, L.episode);
        EvSink<false> ev{nullptr, 0, 0};
        for (int t = 0; t < n_ticks; ++t) {
            const uint4 blk = draw_block(s, DOM_TICK, SUB_MAIN, (uint32_t)L.tick);
            const int m1 = bot_move(bot1, L.p1, blk.x), m2 = bot_move(bot2, L.p2, blk.y);
            const int res = tick_lane<DGEN, NPC, false>(P, tiles, L, m1, m2, blk.z, s, i, ev, cnt);
            ++cnt.ticks;
            if (res != ORX_RESULT_IN_PROGRESS) {
                cnt.p1 += res == ORX_RESULT_PLAYER1_WIN;
                cnt.p2 += res == ORX_RESULT_PLAYER2_WIN;
                cnt.ties += res == ORX_RESULT_TIE;
                if (P.auto_reset) {
                    s.episode += 1;
                    reset_lane<DGEN, NPC>(P, L, s, i);
                } else {
                    status = res;
                    break;
                }
            }
        }
        cnt.events += (unsigned int)ev.n;
        store_lane(P, i, L, status);
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

// ------------------------------------------------------------------ observations (state.py:53-58)
// obs[i][p] = { x, y, depth (saturated), hp, other_visible, other_x, other_y, other_hp,
//               stairs_visible, stairs_x, stairs_y, tick (saturated) }
__global__ void __launch_bounds__(kThreads)
k_observe(const __grid_constant__ Params P, int16_t* __restrict__ obs, int radius)
{
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < P.n; i += stride) {
        Lane L;
        load_lane(P, i, L);
        const bool same = L.p1.depth == L.p2.depth;   // view_for keeps entities on the viewer's depth
        const int tk = min(L.tick, 32767);
#pragma unroll
        for (int p = 0; p < 2; ++p) {
            const Mover& me = p == 0 ? L.p1 : L.p2;
            const Mover& ot = p == 0 ? L.p2 : L.p1;
            const bool has_st = me.sx != ORX_NO_STAIRS;
            const bool st_vis = has_st && (radius < 0 || max(abs(me.sx - me.x), abs(me.sy - me.y)) <= radius);
            // packed as 3 x 8-byte stores of int16 quads
            const short q0[4] = {(short)me.x, (short)me.y, (short)min(me.depth, 32767), (short)me.hp};
            const short q1[4] = {(short)same, (short)(same ? ot.x : -1), (short)(same ? ot.y : -1), (short)(same ? ot.hp : 0)};
            const short q2[4] = {(short)st_vis, (short)(st_vis ? me.sx : -1), (short)(st_vis ? me.sy : -1), (short)tk};
            short* o = obs + (i * 2 + p) * ORX_OBS_LEN;
            *reinterpret_cast<uint2*>(o) = *reinterpret_cast<const uint2*>(q0);
            *reinterpret_cast<uint2*>(o + 4) = *reinterpret_cast<const uint2*>(q1);
            *reinterpret_cast<uint2*>(o + 8) = *reinterpret_cast<const uint2*>(q2);
        }
    }
}

// ------------------------------------------------------------------ host side
int cuda_fail(cudaError_t e) { return ORX_ERR_CUDA_BASE - (int)e; }

bool aligned(const void* p, size_t a) { return (reinterpret_cast<uintptr_t>(p) & (a - 1)) == 0; }

int check_common(const OrxConfig* cfg, const OrxState* st, int64_t n)
{
    if (cfg == nullptr || st == nullptr || n < 0) return ORX_ERR_BAD_ARG;
    if (cfg->struct_size != sizeof(OrxConfig)) return ORX_ERR_BAD_ARG;
    if (cfg->width < 4 || cfg->height < 4 || cfg->width > ORX_MAX_DIM || cfg->height > ORX_MAX_DIM) return ORX_ERR_BAD_ARG;
    if (cfg->n_npc < 0 || cfg->n_npc > ORX_MAX_NPC) return ORX_ERR_BAD_ARG;
    if (cfg->dgen_kind != ORX_DGEN_EMPTY && cfg->dgen_kind != ORX_DGEN_FIXED) return ORX_ERR_UNSUPPORTED;
    if (cfg->start_kind != ORX_START_TOGETHER && cfg->start_kind != ORX_START_SEPARATED) return ORX_ERR_UNSUPPORTED;
    if (cfg->despawn_strat != ORX_DESPAWN_UNREACHABLE && cfg->despawn_strat != ORX_DESPAWN_UNUSED) return ORX_ERR_BAD_ARG;
    if (cfg->start_kind == ORX_START_SEPARATED && cfg->start_depth[0] == cfg->start_depth[1]) return ORX_ERR_BAD_ARG;
    if (cfg->max_ticks < 0) return ORX_ERR_BAD_ARG;
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
    P.k0 = (uint32_t)c->seed; P.k1 = (uint32_t)(c->seed >> 32);
    P.tiles = c->fixed_tiles; P.ground = c->fixed_ground; P.n_ground = c->fixed_n_ground;
    P.fsx = c->fixed_stairs[0]; P.fsy = c->fixed_stairs[1];
    P.pos = reinterpret_cast<uint32_t*>(st->pos); P.hp = reinterpret_cast<uint32_t*>(st->hp);
    P.depth = reinterpret_cast<int2*>(st->depth); P.stairs = reinterpret_cast<uint32_t*>(st->stairs);
    P.tick = st->tick; P.episode = st->episode; P.status = st->status;
    P.npc_pos = st->npc_pos; P.npc_hp = st->npc_hp; P.npc_depth = st->npc_depth;
    P.n = n; P.gid_base = gid_base;
    return P;
}

int grid_for(int64_t n, int ctas_per_sm)
{
    int dev = 0, sms = 148;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int64_t need = (n + kThreads - 1) / kThreads;
    const int64_t cap = (int64_t)sms * ctas_per_sm;      // whole waves of resident CTAs
    return (int)(need < cap ? need : cap);
}

size_t tiles_smem(const OrxConfig* c) { return c->dgen_kind == ORX_DGEN_FIXED ? (size_t)c->width * c->height : 0; }

template <typename F>
int dispatch_dgen_npc(const OrxConfig* c, F&& f)
{
    const bool npc = c->n_npc > 0;
    if (c->dgen_kind == ORX_DGEN_EMPTY) return npc ? f.template operator()<ORX_DGEN_EMPTY, true>() : f.template operator()<ORX_DGEN_EMPTY, false>();
    return npc ? f.template operator()<ORX_DGEN_FIXED, true>() : f.template operator()<ORX_DGEN_FIXED, false>();
}

int launch_done() {
    const cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? ORX_OK : cuda_fail(e);
}

}  // namespace

extern "C" {

int orx_abi_version(void) { return ORX_ABI_VERSION; }

const char* orx_strerror(int code)
{
    if (code == ORX_OK) return "ok";
    if (code == ORX_ERR_BAD_ARG) return "bad argument (null/misaligned pointer, config out of range, or struct_size mismatch)";
    if (code == ORX_ERR_UNSUPPORTED) return "unsupported configuration";
    if (code <= ORX_ERR_CUDA_BASE) return cudaGetErrorString((cudaError_t)(ORX_ERR_CUDA_BASE - code));
    return "unknown error";
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

int orx_reset(const OrxConfig* cfg, const OrxState* st, const uint8_t* mask, int bump_episode,
              int64_t n, uint64_t game_id_base, void* cuda_stream)
{
    const int rc = check_common(cfg, st, n);
    if (rc != ORX_OK) return rc;
    if (n == 0) return ORX_OK;
    const Params P = make_params(cfg, st, n, game_id_base);
    cudaStream_t s = static_cast<cudaStream_t>(cuda_stream);
    const int grid = grid_for(n, 8);
    return dispatch_dgen_npc(cfg, [&]<int DGEN, bool NPC>() {
        k_reset<DGEN, NPC><<<grid, kThreads, 0, s>>>(P, mask, bump_episode);
        return launch_done();
    });
}

int orx_step(const OrxConfig* cfg, const OrxState* st, const uint8_t* moves, uint8_t* result,
             OrxEvent* events, int64_t n, uint64_t game_id_base, void* cuda_stream)
{
    const int rc = check_common(cfg, st, n);
    if (rc != ORX_OK) return rc;
    if (moves == nullptr || result == nullptr || !aligned(moves, 2) || (events && !aligned(events, 8))) return ORX_ERR_BAD_ARG;
    if (n == 0) return ORX_OK;
    const Params P = make_params(cfg, st, n, game_id_base);
    cudaStream_t s = static_cast<cudaStream_t>(cuda_stream);
    const int grid = grid_for(n, 8);
    const size_t smem = tiles_smem(cfg);
    const uint16_t* mv = reinterpret_cast<const uint16_t*>(moves);
    uint2* ev = reinterpret_cast<uint2*>(events);
    const int max_ev = orx_max_events(cfg);
    return dispatch_dgen_npc(cfg, [&]<int DGEN, bool NPC>() {
        if (ev != nullptr) k_step<DGEN, NPC, true><<<grid, kThreads, smem, s>>>(P, mv, result, ev, max_ev);
        else k_step<DGEN, NPC, false><<<grid, kThreads, smem, s>>>(P, mv, result, nullptr, max_ev);
        return launch_done();
    });
}

int orx_step_host(const OrxConfig* cfg, const OrxState* st, const uint8_t* moves_host,
                  uint8_t* result_host, uint8_t* moves_dev, uint8_t* result_dev, int64_t n,
                  uint64_t game_id_base, void* cuda_stream)
{
    if (moves_host == nullptr || result_host == nullptr || moves_dev == nullptr || result_dev == nullptr) return ORX_ERR_BAD_ARG;
    const int rc = check_common(cfg, st, n);
    if (rc != ORX_OK) return rc;
    if (n == 0) return ORX_OK;
    cudaStream_t s = static_cast<cudaStream_t>(cuda_stream);
    cudaError_t e = cudaMemcpyAsync(moves_dev, moves_host, (size_t)n * 2, cudaMemcpyHostToDevice, s);
    if (e != cudaSuccess) return cuda_fail(e);
    const int rs = orx_step(cfg, st, moves_dev, result_dev, nullptr, n, game_id_base, cuda_stream);
    if (rs != ORX_OK) return rs;
    e = cudaMemcpyAsync(result_host, result_dev, (size_t)n, cudaMemcpyDeviceToHost, s);
    return e == cudaSuccess ? ORX_OK : cuda_fail(e);
}

int orx_bot_moves(const OrxConfig* cfg, const OrxState* st, int bot_p1, int bot_p2,
                  uint8_t* moves, int64_t n, uint64_t game_id_base, void* cuda_stream)
{
    const int rc = check_common(cfg, st, n);
    if (rc != ORX_OK) return rc;
    if (moves == nullptr || !aligned(moves, 2)) return ORX_ERR_BAD_ARG;
    if (bot_p1 < ORX_BOT_NONE || bot_p1 > ORX_BOT_STAIRCASE || bot_p2 < ORX_BOT_NONE || bot_p2 > ORX_BOT_STAIRCASE) return ORX_ERR_BAD_ARG;
    if (n == 0) return ORX_OK;
    const Params P = make_params(cfg, st, n, game_id_base);
    k_bot_moves<<<grid_for(n, 8), kThreads, 0, static_cast<cudaStream_t>(cuda_stream)>>>(P, bot_p1, bot_p2, moves);
    return launch_done();
}

int orx_rollout(const OrxConfig* cfg, const OrxState* st, int bot_p1, int bot_p2, int n_ticks,
                unsigned long long* stats, int64_t n, uint64_t game_id_base, void* cuda_stream)
{
    const int rc = check_common(cfg, st, n);
    if (rc != ORX_OK) return rc;
    if (n_ticks < 0 || (stats && !aligned(stats, 8))) return ORX_ERR_BAD_ARG;
    if (bot_p1 < ORX_BOT_NONE || bot_p1 > ORX_BOT_STAIRCASE || bot_p2 < ORX_BOT_NONE || bot_p2 > ORX_BOT_STAIRCASE) return ORX_ERR_BAD_ARG;
    if (n == 0 || n_ticks == 0) return ORX_OK;
    const Params P = make_params(cfg, st, n, game_id_base);
    cudaStream_t s = static_cast<cudaStream_t>(cuda_stream);
    const int grid = grid_for(n, 8);
    const size_t smem = tiles_smem(cfg);
    return dispatch_dgen_npc(cfg, [&]<int DGEN, bool NPC>() {
        k_rollout<DGEN, NPC><<<grid, kThreads, smem, s>>>(P, bot_p1, bot_p2, n_ticks, stats);
        return launch_done();
    });
}

int orx_observe(const OrxConfig* cfg, const OrxState* st, int16_t* obs, int stairs_radius,
                int64_t n, void* cuda_stream)
{
    const int rc = check_common(cfg, st, n);
    if (rc != ORX_OK) return rc;
    if (obs == nullptr || !aligned(obs, 8)) return ORX_ERR_BAD_ARG;
    if (n == 0) return ORX_OK;
    const Params P = make_params(cfg, st, n, 0);
    k_observe<<<grid_for(n, 8), kThreads, 0, static_cast<cudaStream_t>(cuda_stream)>>>(P, obs, stairs_radius);
    return launch_done();
}

}  // extern "C"
