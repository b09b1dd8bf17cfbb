// Persistent, TMA-fed version of the tick kernel (plain, with the observations, with the event log, with
// scripted players, or with NPC slots).
//
// Every CTA works through 256-game tiles: its first kStages tiles are fixed (blockIdx.x + i * grid),
// the rest are claimed from a per-state counter (OrxState.sched) as stages free up, so that CTAs
// which happen to run slower -- the spread of finishing times under static assignment was 6 us on
// a 14 us launch -- take fewer tiles. Without a counter (sched == NULL) the assignment is the static
// stride. One producer thread streams each tile's plane slices HBM -> shared memory (completion on
// an mbarrier), kStages tiles ahead of the eight compute warps: the five 4-byte planes as ONE 2-D
// tensor-map box when the caller stacked them at a common pitch (else five 1-D bulk copies), depth,
// status and commands as 1-D bulk copies. The compute warps pull their game into registers, run
// tick_lane, write the new planes back into the same stage, and the producer streams the stage
// shared memory -> HBM with bulk stores before refilling it. Loads therefore stay in flight for the
// whole life of the CTA instead of only at the start of each thread, and no thread does per-plane
// 64-bit address arithmetic.
//
// Stage life cycle (stage s, tiles s, s+kStages, ...):
//   producer: wait_group.read (earlier bulk stores have drained the stage) -> expect_tx(full[s])
//             -> loads (4 copies with the tensor map, 8 without)
//   consumer: wait full[s] -> smem -> registers -> tick -> registers -> smem (in place)
//             -> fence.proxy.async -> one arrive per warp on done[s]
//   producer: wait done[s] -> stores + commit_group; with >= 5 stages the refill of a stage happens
//             one iteration after its stores were committed (wait_group.read 1), so the producer never
//             sits out the read-out of the stage it has just handed to the store engine
//
// Between launches (two modes, chosen by the host per state, see launch_pipe):
//   grid-wait mode  PDL lets the next grid's CTAs become resident while this grid drains; until
//                   griddepcontrol.wait releases them they prefetch their first two tiles into L2.
//   flag mode       (flags != NULL) consecutive launches overlap. A CTA owns a CONTIGUOUS run of tiles_per_cta
//                   tiles (the same run in every launch on the state), handed over in chunks of kChunk tiles -- as
//                   shipped ONE chunk, the whole run (runs are at most 32 tiles). Each chunk of the state has two
//                   words in OrxState.sched: next (tickets handed out) and serving (passes completed). A CTA draws
//                   the tickets of its chunks before it lets dependents launch, so tickets are in launch order; it
//                   loads a chunk once serving equals its ticket and releases serving = ticket + 1 (gpu scope) once
//                   the chunk's bulk stores have completed. There is no grid-wide wait before the first load: CTA b
//                   of tick k+1 starts as soon as CTA b of tick k is done, whatever the rest of tick k is doing, and
//                   launches on different states do not wait for each other at all. The producer runs
//                   griddepcontrol.wait last, before it exits, so that "this grid is complete" still implies "every
//                   earlier grid in the stream is complete" for whatever the caller enqueues next.
//                   Why not finer: the release is MEMBAR.ALL.GPU in the one thread that moves data (0.65 us). Per
//                   tile it cost more than the overlap gained (19.0 against 12.5 us per 2^20-game step), per 4 / 8 /
//                   16 / 32 tiles a step took 11.7 / 10.8 / 10.3 / 10.1 us; and without the fence the completion of a
//                   bulk store is only visible to its own thread -- relaxed flags gave wrong planes in long
//                   unsynchronised runs (profiles/r02_ab_flag_fences.log, r02_ab_chunk_size.log).
#pragma once
#include <cuda.h>             // CUtensorMap (type only; the encoder is looked up at run time)
#include <cuda_runtime.h>
#include <stdint.h>

#include <type_traits>

#include "orx_rules.cuh"

namespace orx {

#ifndef ORX_PIPE_TILE
#define ORX_PIPE_TILE 256
#endif
constexpr int kTile = ORX_PIPE_TILE;  // games per tile = compute threads per CTA
#ifndef ORX_PIPE_STAGES
#define ORX_PIPE_STAGES 6
#endif
#ifndef ORX_PIPE_MINBLOCKS
#define ORX_PIPE_MINBLOCKS 3
#endif
#ifndef ORX_PIPE_PDL
#define ORX_PIPE_PDL 1
#endif
#ifndef ORX_PIPE_LAZY_REFILL
#define ORX_PIPE_LAZY_REFILL 1
#endif
#ifndef ORX_PIPE_PREFETCH
#define ORX_PIPE_PREFETCH 2          // fixed first tiles per CTA whose planes are prefetched into L2 ahead of the grid dependency
#endif
constexpr int kStages = ORX_PIPE_STAGES;
#ifndef ORX_PIPE_HELPER_WARPS
#define ORX_PIPE_HELPER_WARPS 1
#endif
constexpr int kPipeThreads = kTile + 32 * ORX_PIPE_HELPER_WARPS;   // + the producer warp (tuning builds: more helper warps, idle, to see what the register cap costs)
constexpr uint32_t kTileIdxBytes = ((uint32_t)kStages * 4u + 15u) & ~15u;   // per-stage tile index words
#ifndef ORX_PIPE_OBS_STAGES
#define ORX_PIPE_OBS_STAGES 3
#endif
// With observations each stage also carries the tile's 48 B/game observation block (12 KB), so three
// stages keep three CTAs per SM resident.
#ifndef ORX_PIPE_EV_STAGES
#define ORX_PIPE_EV_STAGES 4
#endif
// ... and with the event log its 32 B/game of records (8 KB): four stages.
// ... and with N NPC slots per game their three slot planes (8 N bytes per game): as many stages as fit the ~72 KB that
// keep three CTAs per SM resident, between 3 and 6 (one slot: 6 stages of 10 KB, eight slots: 3 of 24 KB).
constexpr int npc_stages(int npc) { const int fit = 72 * 1024 / (int)(29u * ORX_PIPE_TILE + 3u * ORX_PIPE_TILE + 8u * (unsigned)npc * ORX_PIPE_TILE); return fit > 6 ? 6 : fit < 3 ? 3 : fit; }
template <bool OBS, bool EV = false, int NPC = 0> constexpr int kPipeStages = NPC ? npc_stages(NPC) : OBS ? ORX_PIPE_OBS_STAGES : EV ? ORX_PIPE_EV_STAGES : kStages;

// byte offsets of the plane slices inside a stage (all multiples of 16)
constexpr uint32_t T4 = 4u * kTile, T8 = 8u * kTile, T2 = 2u * kTile, T1 = kTile;   // slice sizes in bytes
constexpr uint32_t OFF_POS = 0, OFF_HP = T4, OFF_ST = 2 * T4, OFF_TICK = 3 * T4, OFF_EP = 4 * T4,
                   OFF_DEPTH = 5 * T4, OFF_STATUS = 5 * T4 + T8, OFF_MOVES = OFF_STATUS + T1, OFF_RESULT = OFF_MOVES + T2,
                   STAGE_BYTES = OFF_RESULT + T1;
constexpr uint32_t PLANE_LOAD_BYTES = 5 * T4 + T8 + T1;         // per tile, HBM -> smem, without the commands
constexpr uint32_t OBS_GAME_BYTES = 2u * ORX_OBS_LEN * 2u, OBS_BYTES = OBS_GAME_BYTES * kTile, OFF_OBS = STAGE_BYTES;
static_assert(OBS_GAME_BYTES == 48 && (STAGE_BYTES % 128) == 0, "observation block: 12 words per game behind the planes");
constexpr uint32_t EV_GAME_BYTES = 8u * ORX_MAX_EVENTS_BASE, EV_BYTES = EV_GAME_BYTES * kTile, OFF_EV = STAGE_BYTES;   // event records share the slot behind the planes
static_assert(EV_GAME_BYTES == 32, "four 8-byte records per game without NPC slots");
// NPC slot planes of a tile with N slots per game (compile time): pos 2 B, hp 2 B, depth 4 B per slot, behind the planes
constexpr uint32_t OFF_NPOS = STAGE_BYTES;
template <int NPC> constexpr uint32_t kNpcPosBytes = 2u * (uint32_t)NPC * kTile;          // = hp slice; the depth slice is twice that
template <bool OBS, bool EV = false, int NPC = 0> constexpr uint32_t kPipeStageBytes = STAGE_BYTES + (OBS ? OBS_BYTES : 0u) + (EV ? EV_BYTES : 0u) + 4u * kNpcPosBytes<NPC>;
// Command formats: CMD_BYTES = uint8[n][2] (p1, p2); CMD_NIBBLES = uint8[n], p1 in the low nibble,
// p2 in the high nibble (halves the command traffic when the commands come over PCIe).
constexpr int CMD_BYTES = 0, CMD_NIBBLES = 1, CMD_BYTES_BOTS = 2;   // _BOTS: uint8[n][2], scripted players' commands computed in the kernel
// CMD_BITS: the bit-packed streams of orx_step_bits (orx.h): 5 bits of command pair per game in, 2 bits of result
// per game out -- what crosses PCIe when the caller's buffers live in host memory. A CTA then owns a CONTIGUOUS
// run of tiles (tiles_per_cta of them). The commands arrive four tiles (640 bytes) per bulk copy, each copy with its
// own mbarrier, requested together with the planes of the chunk's first tile, i.e. kStages tiles before they are
// needed and in the order they are needed (asking for the whole run at once made a CTA wait for its whole share of
// the link's 16 us per 2^20 games before its first tile: 27 instead of ~19 us per tick). The results of the run go
// back with ONE bulk copy when the CTA ends. No PCIe transaction sits on a tile's critical path.
constexpr int CMD_BITS = 3;
constexpr uint32_t kCmdBitsTile = 5u * kTile / 8u, kResBitsTile = 2u * kTile / 8u;      // 160 and 64 bytes per tile
constexpr uint32_t kBitsMaxTiles = 32;                                                    // tiles per CTA in CMD_BITS mode
constexpr uint32_t kBitsCmdBytes = kBitsMaxTiles * kCmdBitsTile + 16u, kBitsResBytes = kBitsMaxTiles * kResBitsTile;
constexpr uint32_t kBitsCmdChunk = 4;                                                     // tiles per command fetch
constexpr uint32_t kBitsBarBytes = 8u * (kBitsMaxTiles / kBitsCmdChunk);                  // one mbarrier per command fetch
constexpr uint32_t kBitsSmemBytes = kBitsBarBytes + kBitsCmdBytes + kBitsResBytes;       // mbarriers + command block + result block
static_assert(kCmdBitsTile % 16 == 0 && kResBitsTile % 16 == 0, "bit-packed tiles move as bulk copies");
static_assert(kTile % 32 == 0 && (T1 % 16) == 0, "bulk copies move multiples of 16 bytes");
static_assert(OFF_HP == OFF_POS + T4 && OFF_ST == OFF_HP + T4 && OFF_TICK == OFF_ST + T4 && OFF_EP == OFF_TICK + T4 && kTile <= 256,
              "the five 4-byte slices are the rows of the tensor-map box, in plane order");

__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "ORX_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra ORX_DONE;\n"
        "bra ORX_WAIT;\n"
        "ORX_DONE:\n"
        "}\n" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_load(uint32_t dst_smem, const void* src, uint32_t bytes, uint32_t bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst_smem), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void bulk_store(void* dst, uint32_t src_smem, uint32_t bytes)
{
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                 ::"l"(dst), "r"(src_smem), "r"(bytes) : "memory");
}
// 2-D tiled copies through a tensor map: when the five 4-byte planes (pos, hp, stairs, tick, episode) sit
// in one allocation at a common pitch they form a u32[5][n] array, and the {256 games x 5 planes} box of a
// tile moves with ONE instruction instead of five; its rows land back to back in the stage, which is the
// order the stage already uses. Every bulk instruction costs the producer thread ~40 ns, so this is what
// bounds how fast a CTA can get its first stages going.
__device__ __forceinline__ void tensor_load_2d(uint32_t dst_smem, const CUtensorMap* map, uint32_t c0, uint32_t c1, uint32_t bar)
{
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(dst_smem), "l"(map), "r"(c0), "r"(c1), "r"(bar) : "memory");
}
__device__ __forceinline__ void tensor_store_2d(const CUtensorMap* map, uint32_t c0, uint32_t c1, uint32_t src_smem)
{
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];"
                 ::"l"(map), "r"(c0), "r"(c1), "r"(src_smem) : "memory");
}
__device__ __forceinline__ void tensor_prefetch_l2_2d(const CUtensorMap* map, uint32_t c0, uint32_t c1)
{
    asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global [%0, {%1, %2}];" ::"l"(map), "r"(c0), "r"(c1) : "memory");
}
// Hint: pull [src, src + bytes) into L2. No architectural effect (L2 is the coherence point of the GPU, a line
// written later by an earlier grid is updated in place), so it may run ahead of griddepcontrol.wait.
__device__ __forceinline__ void bulk_prefetch_l2(const void* src, uint32_t bytes)
{
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read_all() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read_but_last() { asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ uint32_t lds_u32(uint32_t a) { uint32_t v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
__device__ __forceinline__ uint32_t lds_u16(uint32_t a) { uint32_t v; asm volatile("ld.shared.u16 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
__device__ __forceinline__ uint32_t lds_u8(uint32_t a) { uint32_t v; asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
__device__ __forceinline__ int2 lds_s32x2(uint32_t a) { int2 v; asm volatile("ld.shared.v2.s32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(a)); return v; }
__device__ __forceinline__ void sts_u32(uint32_t a, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ void sts_u8(uint32_t a, uint32_t v) { asm volatile("st.shared.u8 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ void sts_s32x2(uint32_t a, int x, int y) { asm volatile("st.shared.v2.s32 [%0], {%1, %2};" ::"r"(a), "r"(x), "r"(y) : "memory"); }
__device__ __forceinline__ void sts_u32x4(uint32_t a, uint32_t x, uint32_t y, uint32_t z, uint32_t w) { asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(x), "r"(y), "r"(z), "r"(w) : "memory"); }
__device__ __forceinline__ int lds_s16(uint32_t a) { int v; asm volatile("ld.shared.s16 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
__device__ __forceinline__ void sts_u16(uint32_t a, uint32_t v) { asm volatile("st.shared.u16 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
// The NPC slots of one game inside the shared-memory stage (see NpcView in orx_rules.cuh): N is a compile-time
// constant, so the tick's slot loops unroll onto constant offsets from three shared-window addresses.
template <int N>
struct NpcStage {
    static constexpr bool kPresent = true;
    uint32_t pos, hp, depth;      // this game's slot 0 in the stage's npc_pos / npc_hp / npc_depth slices
    __device__ __forceinline__ static constexpr int count() { return N; }
    __device__ __forceinline__ int depth_at(int k) const { return (int)lds_u32(depth + 4u * (uint32_t)k); }
    __device__ __forceinline__ uint32_t xy_at(int k) const { return lds_u16(pos + 2u * (uint32_t)k); }
    __device__ __forceinline__ int hp_at(int k) const { return lds_s16(hp + 2u * (uint32_t)k); }
    __device__ __forceinline__ void set_hp(int k, int v) const { sts_u16(hp + 2u * (uint32_t)k, (uint32_t)v & 0xFFFFu); }
    __device__ __forceinline__ void set_depth(int k, int v) const { sts_u32(depth + 4u * (uint32_t)k, (uint32_t)v); }
};
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
// flag mode: per-tile hand-over between consecutive launches on one state
// flag mode hand-over: acquire / release at gpu scope (LDG.STRONG.GPU + CCTL.IVALL, MEMBAR.ALL.GPU + STG.STRONG.GPU).
// cp.async.bulk.wait_group makes the bulk stores visible to the WAITING thread only; the release is what makes them
// visible to the CTA of the next launch (relaxed accesses here gave wrong planes in long unsynchronised runs).
__device__ __forceinline__ uint32_t ld_acquire_gpu(const unsigned int* p) { uint32_t v; asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ void st_release_gpu(unsigned int* p, uint32_t v) { asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
// generic-proxy flag access <-> async-proxy bulk copies (CCTL.IVALL + FENCE.VIEW.ASYNC.G: no measurable cost)
__device__ __forceinline__ void fence_proxy_async_global() { asm volatile("fence.proxy.async.global;" ::: "memory"); }
template <int N> __device__ __forceinline__ void bulk_wait_group() { asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory"); }
#ifndef ORX_PIPE_CHUNK
#define ORX_PIPE_CHUNK 32            // tiles per hand-over chunk in flag mode: with runs of at most 32 tiles, ONE chunk = the CTA's whole run
#endif
constexpr int kChunk = ORX_PIPE_CHUNK;
constexpr int kMaxChunksPerCta = 32;          // one ticket per lane of the producer warp
constexpr uint32_t kTicketBytes = 4u * kMaxChunksPerCta;

#ifdef ORX_PIPE_JITTER
// Schedule fuzzing (tuning build only, tests/test_gpu_tile_flags.py under ORX_LIB=liborx_jitter.so): pseudo-random pauses
// of up to 2 us at the points where the roles of the pipeline and consecutive launches hand data to each other. A
// missing fence or a wrong barrier phase that the natural timing happens to hide shows up as a parity failure.
__device__ __forceinline__ void orx_jitter(unsigned int salt)
{
    unsigned int t;
    asm volatile("mov.u32 %0, %%clock;" : "=r"(t));
    const unsigned int h = (t ^ (blockIdx.x * 2654435761u) ^ (salt * 40503u) ^ (threadIdx.x >> 5)) * 2246822519u;
    if ((h >> 29) == 0u) __nanosleep((h >> 8) & 2047u);
}
#define ORX_JITTER(salt) orx_jitter(salt)
#else
#define ORX_JITTER(salt) do { } while (0)
#endif

#ifdef ORX_PIPE_TRACE
// Tuning aid (tools/pipetrace.py, separate build): per-CTA %globaltimer stamps of the last 16 launches.
__device__ unsigned long long g_trace[16][512][24];     // 0-7: see tools/pipetrace.py; 8+k: producer saw tile k of this CTA done
__device__ __forceinline__ unsigned long long gtime() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
#define ORX_TRACE(slot, idx) do { if (blockIdx.x < 512) g_trace[(slot) & 15][blockIdx.x][idx] = gtime(); } while (0)
#define ORX_TRACE_PARAM , unsigned int trace_slot
#else
#define ORX_TRACE(slot, idx) do { } while (0)
#define ORX_TRACE_PARAM
#endif

// n_tiles full tiles of kTile games; the caller handles a ragged tail with the simple kernel.
// TICK: play one tick and write the planes and results back. OBS: also (or, without TICK, only) write
// the per-player observations of the resulting state, 48 B per game, staged behind the planes of the
// stage and streamed out with the same bulk stores.
// EV: also write the tick's replication-log records (OrxEvent[4] per game, no NPC slots), staged like the
// observations and streamed out with one bulk store per tile.
// NPC: the game's NPC slot planes (n_npc <= ORX_MAX_NPC slots) travel with the tile as three more slices.
// FLAGGED: flag mode (flags != NULL) or grid-wait mode, as separate instantiations: the hand-over code in the same
// kernel cost grid-wait mode 9 % at 2^20 and 20 % at 2^22 games per launch (1632 -> 2912 SASS instructions, the
// producer's loops around them; profiles/r02_ab_flag_code_size.log).
template <int DGEN, int CMD, bool OBS, bool TICK, bool EV = false, int NPC = 0, bool FLAGGED = false>
__global__ void __launch_bounds__(kPipeThreads, ORX_PIPE_MINBLOCKS)
k_step_pipe(const __grid_constant__ Params P, const __grid_constant__ CUtensorMap planes5, const int use_map,
            const void* __restrict__ moves_v, uint8_t* __restrict__ result,
            unsigned int n_tiles, unsigned int* __restrict__ sched, unsigned int* __restrict__ flags,
            int16_t* __restrict__ obs, int obs_radius, uint2* __restrict__ events, int bots, unsigned int tiles_per_cta ORX_TRACE_PARAM)
{
    static_assert(OBS || TICK, "nothing to do");
    static_assert(!EV || (TICK && !OBS), "the event log rides with the plain tick");
    static_assert(!NPC || (TICK && !OBS && !EV), "NPC slots ride with the plain tick");
    if (threadIdx.x == 0) ORX_TRACE(trace_slot, 0);
    constexpr int kStages = kPipeStages<OBS, EV, NPC>;             // shadows the namespace constant on purpose
    constexpr uint32_t STAGE_BYTES = kPipeStageBytes<OBS, EV, NPC>;
    constexpr uint32_t npc2 = kNpcPosBytes<NPC>;                   // bytes of a tile's npc_pos / npc_hp slice; npc_depth: twice that
    constexpr uint32_t OFF_NHP = OFF_NPOS + npc2, OFF_NDEPTH = OFF_NHP + npc2;
    constexpr uint32_t MV_BYTES = CMD == CMD_BITS ? 0u : CMD == CMD_NIBBLES ? T1 : T2, LOAD_BYTES = PLANE_LOAD_BYTES + (TICK ? MV_BYTES : 0u);
    static_assert(CMD != CMD_BITS || (TICK && !OBS && !EV && !NPC), "bit-packed streams ride with the plain tick");
    const uint8_t* moves = static_cast<const uint8_t*>(moves_v);
    extern __shared__ __align__(128) uint8_t smem[];
    uint8_t* stages = smem;                                             // kStages * STAGE_BYTES
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kStages * STAGE_BYTES);   // full[kStages], done[kStages]
    constexpr uint32_t kBitsBytes = CMD == CMD_BITS ? kBitsSmemBytes : 0u;
    const uint32_t bits0 = smem_addr(smem + kStages * STAGE_BYTES + 2 * kStages * 8 + kTileIdxBytes + kTicketBytes);
    const uint32_t cmdbar = bits0, cmd0 = bits0 + kBitsBarBytes, res0 = cmd0 + kBitsCmdBytes;      // CMD_BITS only
    uint8_t* tiles_sm = smem + kStages * STAGE_BYTES + 2 * kStages * 8 + kTileIdxBytes + kTicketBytes + kBitsBytes;
    const uint32_t full0 = smem_addr(bars), done0 = smem_addr(bars + kStages);
    const uint32_t tidx0 = smem_addr(bars + 2 * kStages);      // tile index published with each stage
    const uint32_t tk0 = tidx0 + kTileIdxBytes;                // flag mode: this CTA's ticket of each of its chunks
    const uint32_t stage0 = smem_addr(stages);
    const unsigned int tid = threadIdx.x;
    constexpr bool flagged = FLAGGED;
    // Programmatic dependent launch: the next kernel in the stream may begin while this grid is still
    // running. Grid-wait mode: at once; its producer waits for this grid to complete
    // (griddepcontrol.wait) before it touches any plane. Flag mode: only after this CTA holds the tickets
    // of all its tiles (below), which is what keeps the tickets of consecutive launches in launch order.
    if (ORX_PIPE_PDL && !flagged) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    // Static assignment: tile of this CTA's it-th iteration. Strided over the grid, or (flag mode, CMD_BITS) a contiguous run.
    constexpr uint32_t NONE = 0xFFFFFFFFu;    // published instead of a tile index: no more work for this CTA
    auto tile_at = [&](unsigned int it) -> uint32_t {
        if (CMD == CMD_BITS || flagged) {
            const uint64_t t = (uint64_t)blockIdx.x * tiles_per_cta + it;
            return it < tiles_per_cta && t < n_tiles ? (uint32_t)t : NONE;
        }
        const uint64_t t = (uint64_t)blockIdx.x + (uint64_t)it * gridDim.x;
        return t < n_tiles ? (uint32_t)t : NONE;
    };
    // flag mode: chunk c of this CTA = its tiles [c kChunk, (c+1) kChunk) (the last one may be shorter); chunks are
    // numbered CTA by CTA, so a chunk is the same tiles in every launch on the state
    const size_t chunk0 = (size_t)blockIdx.x * ((tiles_per_cta + (unsigned)kChunk - 1u) / (unsigned)kChunk);
#ifndef ORX_PIPE_LATE_TICKET
    if (flagged && tid >= kTile) {
        // FIRST thing the producer warp does, ahead of the barrier set-up and the command table: lane c draws the
        // ticket of this CTA's chunk c (one atomic instruction, one round trip for all of them) and the CTA lets the
        // dependents go. The value has arrived when the shared-memory store that depends on it has been issued, i.e.
        // the atomic has been performed at the L2 before the next launch can start -- which keeps the tickets of
        // consecutive launches in launch order. The next launch starts once EVERY CTA of this one is past this
        // point, so everything ahead of it is on the launch-to-launch critical path of back-to-back ticks.
        const unsigned int c = tid - kTile;
        ORX_JITTER(1u);
#ifdef ORX_EXPERIMENT_TRIGGER_BEFORE_TICKET     // tuning experiment only (tickets may come out of launch order): what the ticket's round trip costs the launch chain
        if (ORX_PIPE_PDL && tid == kTile) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
#endif
        if (tile_at(c * (unsigned)kChunk) != NONE) sts_u32(tk0 + 4u * c, atomicAdd(flags + 2 * (chunk0 + c), 1u));
        __syncwarp();
#ifndef ORX_EXPERIMENT_TRIGGER_BEFORE_TICKET
        if (ORX_PIPE_PDL && tid == kTile) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
#endif
    }
#endif

    if (tid == 0) {
        for (int s = 0; s < kStages; ++s) {
            mbar_init(full0 + 8 * s, 1);
            mbar_init(done0 + 8 * s, kTile / 32);
        }
        if (CMD == CMD_BITS) for (uint32_t c = 0; c < kBitsMaxTiles / kBitsCmdChunk; ++c) mbar_init(cmdbar + 8u * c, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __shared__ CmdEntry lut[256];
    if (TICK) build_cmd_lut(P, lut, tid, kPipeThreads);
    const uint8_t* tiles = nullptr;
    if (TICK && DGEN == ORX_DGEN_FIXED) {
        const int nt = P.W * P.H;
        for (int t = tid; t < nt; t += kPipeThreads) tiles_sm[t] = P.tiles[t];
        tiles = tiles_sm;
    }
    __syncthreads();

    if (tid >= kTile) {
        // ------------------------------------------------------------ producer (one thread)
#ifdef ORX_PIPE_LATE_TICKET       // tuning build: tickets and hand-off to the dependents after the set-up (the round-2 order until call P)
        if (flagged) {
            const unsigned int c = tid - kTile;
            if (tile_at(c * (unsigned)kChunk) != NONE) sts_u32(tk0 + 4u * c, atomicAdd(flags + 2 * (chunk0 + c), 1u));
            __syncwarp();
        }
        if (tid != kTile) return;
        if (ORX_PIPE_PDL && flagged) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
#else
        if (tid != kTile) return;
#endif
        if (use_map) asm volatile("prefetch.tensormap [%0];" ::"l"(&planes5) : "memory");     // descriptor fetch off the first copy's path
        auto peek = [&](unsigned int c) -> uint32_t {        // serving word of this CTA's chunk c (0 beyond its run)
            return tile_at(c * (unsigned)kChunk) != NONE ? ld_acquire_gpu(flags + 2 * (chunk0 + c) + 1) : 0u;
        };
        auto await = [&](unsigned int c, uint32_t seen) {    // seen: an earlier peek at the chunk's serving word
            const uint32_t want = lds_u32(tk0 + 4u * c);
            ORX_JITTER(2u + c);
            while (seen != want) seen = ld_acquire_gpu(flags + 2 * (chunk0 + c) + 1);
            fence_proxy_async_global();         // the acquire (generic proxy) before the bulk loads (async proxy)
        };
        auto publish = [&](unsigned int c) {                 // this CTA's bulk stores of its chunk c have completed
            ORX_JITTER(40u + c);
            fence_proxy_async_global();
            st_release_gpu(flags + 2 * (chunk0 + c) + 1, lds_u32(tk0 + 4u * c) + 1u);
        };
        // Publishes tile `tile` (or NONE) in the stage of iteration `it` and starts its loads.
        auto issue = [&](unsigned int it, uint32_t tile) {
            const unsigned int s = it % kStages;
            const uint32_t bar = full0 + 8 * s, base = stage0 + s * STAGE_BYTES;
            ORX_JITTER(100u + it);
            sts_u32(tidx0 + 4 * s, tile);
            if (tile == NONE) { mbar_arrive(bar); return; }
            const size_t g = (size_t)tile * kTile;     // first game of the tile
            if (CMD == CMD_BITS && it % kBitsCmdChunk == 0u) {
                // The commands of this tile and the three after it, one 640-byte copy with its own mbarrier, requested
                // when the tile's planes are (kStages tiles ahead of its tick). Every CTA asks for its first chunk
                // first and for the later ones as it gets to them, so a link that serves requests in order delivers
                // them in the order they are needed.
                const uint32_t left = tiles_per_cta - it, n_here = min(min(left, kBitsCmdChunk), n_tiles - tile);
                const uint32_t cbar = cmdbar + 8u * (it / kBitsCmdChunk);
                mbar_expect_tx(cbar, n_here * kCmdBitsTile);
                bulk_load(cmd0 + it * kCmdBitsTile, moves + (size_t)tile * kCmdBitsTile, n_here * kCmdBitsTile, cbar);
            }
            mbar_expect_tx(bar, LOAD_BYTES + 4u * npc2);
            if (NPC) {
                bulk_load(base + OFF_NPOS, P.npc_pos + g * 2u * NPC, npc2, bar);
                bulk_load(base + OFF_NHP, P.npc_hp + g * NPC, npc2, bar);
                bulk_load(base + OFF_NDEPTH, P.npc_depth + g * NPC, 2u * npc2, bar);
            }
            if (use_map) {
                tensor_load_2d(base + OFF_POS, &planes5, (uint32_t)g, 0u, bar);
            } else {
                bulk_load(base + OFF_POS, P.pos + g, T4, bar);
                bulk_load(base + OFF_HP, P.hp + g, T4, bar);
                bulk_load(base + OFF_ST, P.stairs + g, T4, bar);
                bulk_load(base + OFF_TICK, P.tick + g, T4, bar);
                bulk_load(base + OFF_EP, P.episode + g, T4, bar);
            }
            bulk_load(base + OFF_DEPTH, P.depth + g, T8, bar);
            bulk_load(base + OFF_STATUS, P.status + g, T1, bar);
            if (TICK) bulk_load(base + OFF_MOVES, moves + g * (MV_BYTES / kTile), MV_BYTES, bar);
        };
        bool ended = false;
        uint32_t held = 0;      // grid-wait mode: a ticket of the tile counter; flag mode: a peek at the next chunk's serving word
        if (flagged) {
            // Prologue, flag mode: a look at the serving words of the chunks the first kStages tiles belong to, then
            // each tile's loads as soon as the launch before this one on the state has handed its chunk over.
            constexpr unsigned int kPeek = (kStages + kChunk - 1) / kChunk;
            uint32_t seen[kPeek + 1];
#pragma unroll
            for (unsigned int c = 0; c <= kPeek; ++c) seen[c] = peek(c);
#pragma unroll
            for (unsigned int it = 0; it < (unsigned)kStages; ++it) {
                if (!ended) {
                    const uint32_t tile = tile_at(it);
                    if (tile != NONE && it % (unsigned)kChunk == 0u) await(it / (unsigned)kChunk, seen[it / (unsigned)kChunk]);
                    issue(it, tile);
                    ended = tile == NONE;
                }
            }
            held = seen[kPeek];           // chunk kPeek holds tile kPeek * kChunk >= kStages: awaited in the refill loop
        } else {
#if ORX_PIPE_PREFETCH
        // This CTA became resident when a CTA of the previous grid left, i.e. while that grid is still
        // draining its last tiles, and it now has to sit out the rest of that grid; HBM has little to do
        // meanwhile. Pull the planes of the first fixed tiles into L2 so that the first stages fill at L2
        // latency once the dependency resolves. (Measured at 2^20 games per launch: 0 tiles 13.3 us,
        // 1: 12.7, 2: 12.4, 3: 12.6, 4: 12.9 -- more than two compete with the previous grid's last loads.)
        if (ORX_PIPE_PDL) {
            for (unsigned int it = 0; it < (unsigned)ORX_PIPE_PREFETCH && it < (unsigned)kStages; ++it) {
                const uint32_t t = tile_at(it);
                if (t == NONE) break;
                const size_t g = (size_t)t * kTile;
                if (use_map) {
                    tensor_prefetch_l2_2d(&planes5, (uint32_t)g, 0u);
                } else {
                    bulk_prefetch_l2(P.pos + g, T4);
                    bulk_prefetch_l2(P.hp + g, T4);
                    bulk_prefetch_l2(P.stairs + g, T4);
                    bulk_prefetch_l2(P.tick + g, T4);
                    bulk_prefetch_l2(P.episode + g, T4);
                }
                bulk_prefetch_l2(P.depth + g, T8);
                bulk_prefetch_l2(P.status + g, T1);
            }
        }
#endif
#ifndef ORX_EXPERIMENT_NO_GRID_WAIT     // tuning experiment only: measures what the grid dependency costs when consecutive launches touch different states
        if (ORX_PIPE_PDL) asm volatile("griddepcontrol.wait;" ::: "memory");     // all earlier work in the stream is complete and visible
#endif
        ORX_TRACE(trace_slot, 1);
        // Prologue: the first kStages tiles of a CTA are fixed, so its loads start without a round trip
        // to the counter.
        for (unsigned int it = 0; it < (unsigned)kStages && !ended; ++it) {
            const uint32_t tile = tile_at(it);
            issue(it, tile);
            ended = tile == NONE;
        }
        }
        // Tiles beyond the fixed prefix come from the counter: ticket t is tile dyn_base + t. Every CTA
        // that got kStages fixed tiles claims until its first miss, so a launch draws exactly
        // `claims` tickets; whoever draws the last one puts the counter back to zero for the next
        // launch on this state (which reads it only after griddepcontrol.wait).
        const uint32_t dyn_base = (uint32_t)kStages * gridDim.x;
        const uint32_t last_fixed = (uint32_t)(kStages - 1) * gridDim.x;
        const uint32_t claimers = n_tiles > last_fixed ? (n_tiles - last_fixed < gridDim.x ? n_tiles - last_fixed : gridDim.x) : 0u;
        const uint32_t claims = (n_tiles > dyn_base ? n_tiles - dyn_base : 0u) + claimers;
        // The atomic's result is not looked at where it is issued (draw) but one refill later (settle), so
        // its round trip to L2 overlaps the wait for the next stage instead of stalling this thread, the
        // only one that moves data for the CTA (measured: 12.4 -> 12.15 us at 2^20 games per launch,
        // 45.5 -> 42.1 us at 2^22).
        auto draw = [&]() -> uint32_t { return atomicAdd(sched, 1u); };
        auto settle = [&](uint32_t ticket) -> uint32_t {
            if (ticket == claims - 1u) atomicExch(sched, 0u);
            return ticket;
        };
        if (!ended && !flagged && sched != nullptr) held = draw();
#ifdef ORX_PIPE_TRACE
        unsigned int trace_tiles = 0;
        ORX_TRACE(trace_slot, 23);          // prologue loads issued
#endif
        unsigned int n_done = 0;      // tiles this CTA has stored
        for (unsigned int it = 0;; ++it) {
            const unsigned int s = it % kStages;
            const uint32_t tile = lds_u32(tidx0 + 4 * s);
            if (tile == NONE) break;
            mbar_wait(done0 + 8 * s, (it / kStages) & 1u);
            ORX_JITTER(200u + it);
#ifdef ORX_PIPE_TRACE
            if (trace_tiles < 14) ORX_TRACE(trace_slot, 8 + trace_tiles);
            ++trace_tiles;
#endif
            const size_t g = (size_t)tile * kTile;
            const uint32_t base = stage0 + s * STAGE_BYTES;
            if (TICK) {
                if (use_map) {
                    tensor_store_2d(&planes5, (uint32_t)g, 0u, base + OFF_POS);
                } else {
                    bulk_store(P.pos + g, base + OFF_POS, T4);
                    bulk_store(P.hp + g, base + OFF_HP, T4);
                    bulk_store(P.stairs + g, base + OFF_ST, T4);
                    bulk_store(P.tick + g, base + OFF_TICK, T4);
                    bulk_store(P.episode + g, base + OFF_EP, T4);
                }
                bulk_store(P.depth + g, base + OFF_DEPTH, T8);
                bulk_store(P.status + g, base + OFF_STATUS, T1);
                if (CMD != CMD_BITS) bulk_store(result + g, base + OFF_RESULT, T1);
            }
            if (OBS) bulk_store(obs + g * (2 * ORX_OBS_LEN), base + OFF_OBS, OBS_BYTES);
            if (EV) bulk_store(events + g * ORX_MAX_EVENTS_BASE, base + OFF_EV, EV_BYTES);
            if (NPC) {
                bulk_store(P.npc_pos + g * 2u * NPC, base + OFF_NPOS, npc2);
                bulk_store(P.npc_hp + g * NPC, base + OFF_NHP, npc2);
                bulk_store(P.npc_depth + g * NPC, base + OFF_NDEPTH, 2u * npc2);
            }
            bulk_commit();
            if (flagged && (it + 1u) % (unsigned)kChunk == 0u && it + 1u > (unsigned)kChunk) {
                // a chunk's stores have just been committed: hand the chunk BEFORE it to the next launch (its kChunk
                // store groups are the oldest pending ones; waiting for them does not wait for the current chunk)
                bulk_wait_group<kChunk>();
                publish((it + 1u) / (unsigned)kChunk - 2u);
            }
            // Refill. With a deep pipeline (>= 5 stages) one iteration late, i.e. the stage whose stores were
            // committed in the PREVIOUS iteration: waiting for the group just committed parks this thread
            // until the bulk-store engine has read the whole stage out of shared memory, and nothing else
            // issues loads or stores for the CTA meanwhile. Shallow pipelines (the 20 KB stages that also
            // carry observations) cannot spare the stage and refill at once.
            constexpr bool kLazy = ORX_PIPE_LAZY_REFILL && kStages >= 5;
            constexpr unsigned int kLag = kLazy ? 1u : 0u;
            if (!ended && (!kLazy || it >= 1u)) {
                const unsigned int nit = it - kLag + kStages;
                uint32_t nt;
                if (sched != nullptr) {
                    const uint64_t next = (uint64_t)dyn_base + settle(held);
                    nt = next < n_tiles ? (uint32_t)next : NONE;
                } else {
                    nt = tile_at(nit);
                }
                if (kLazy) bulk_wait_read_but_last();      // every group but the one just committed has been read out
                else bulk_wait_read_all();                 // the stage has been read out: safe to overwrite
                if (flagged && nt != NONE && nit % (unsigned)kChunk == 0u) {
                    await(nit / (unsigned)kChunk, held);
                    held = peek(nit / (unsigned)kChunk + 1u);
                }
                issue(nit, nt);
                ended = nt == NONE;
                if (!ended && !flagged && sched != nullptr) held = draw();
            }
            ++n_done;
        }
        if (CMD == CMD_BITS && n_done != 0u) {       // every consumer has arrived on the last tile's barrier: the result block is complete
            bulk_store(result + (size_t)blockIdx.x * tiles_per_cta * kResBitsTile, res0, n_done * kResBitsTile);
            bulk_commit();
        }
        if (flagged) {
            // The chunks not handed over yet (the last one or two): their stores must have completed first. Then the
            // grid dependency, so that the completion of this grid implies the completion of every earlier one.
            bulk_wait_all();
            const unsigned int n_chunks = (n_done + (unsigned)kChunk - 1u) / (unsigned)kChunk;
            const unsigned int handed = n_done / (unsigned)kChunk > 0u ? n_done / (unsigned)kChunk - 1u : 0u;    // chunks published inside the loop
            for (unsigned int c = handed; c < n_chunks; ++c) publish(c);
            if (ORX_PIPE_PDL) asm volatile("griddepcontrol.wait;" ::: "memory");
        } else {
            bulk_wait_read_all(); // shared memory may be released once the last stores have read it; the
                                  // writes themselves complete with the grid (as CUTLASS epilogues do)
        }
        ORX_TRACE(trace_slot, 3);
#ifdef ORX_PIPE_TRACE
        bulk_wait_all();
        ORX_TRACE(trace_slot, 4);
        {   // slot 7: SM id | tiles this CTA handled << 16
            unsigned int smid;
            asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
            if (blockIdx.x < 512) g_trace[trace_slot & 15][blockIdx.x][7] = (unsigned long long)smid | ((unsigned long long)trace_tiles << 16);
        }
#endif
        return;
    }

    // ---------------------------------------------------------------- consumers (8 warps)
    // Shared-memory addresses are formed once, as 32-bit shared-window offsets, and pinned with an
    // empty asm so ptxas keeps them in registers instead of re-deriving them (S2R + LEA chains)
    // every tile.
    uint32_t a4 = stage0 + tid * 4u;            // this thread's slot in the 4-byte slices of stage 0
    asm volatile("" : "+r"(a4));
    for (unsigned int it = 0;; ++it) {
        const unsigned int s = it % kStages;
        const uint32_t b4 = a4 + s * STAGE_BYTES, b8 = b4 + tid * 4u, b1 = b4 - tid * 3u, b2 = b4 - tid * 2u;
        ORX_JITTER(300u + it);
        mbar_wait(full0 + 8 * s, (it / kStages) & 1u);
        const uint32_t tile = lds_u32(tidx0 + 4 * s);
        if (tid == 0 && it == 0) ORX_TRACE(trace_slot, 2);
        if (tile == NONE) break;
        if (tid == 0) ORX_TRACE(trace_slot, 5);
        const uint32_t lane = tile * kTile + tid;      // game index inside the launch
        const uint32_t pos = lds_u32(b4 + OFF_POS), hpw = lds_u32(b4 + OFF_HP), stw = lds_u32(b4 + OFF_ST);
        const int tick = (int)lds_u32(b4 + OFF_TICK);
        const uint32_t ep = lds_u32(b4 + OFF_EP);
        const int2 dep = lds_s32x2(b8 + OFF_DEPTH);
        const int status = (int)lds_u8(b1 + OFF_STATUS);
        Lane L;
        unpack_lane(L, pos, hpw, dep, stw, tick, ep);
        if (TICK) {
            uint32_t mv;
            if (CMD == CMD_BITS) {
                mbar_wait(cmdbar + 8u * (it / kBitsCmdChunk), 0u);         // this tile's commands have arrived (no wait once its fetch is complete)
                // game g of the tile: bits [5g, 5g + 5) of the tile's 160 bytes; v = (p1 - 1) * 5 + (p2 - 1), 25..31 = both Stay
                const uint32_t a = cmd0 + it * kCmdBitsTile + ((tid * 5u) >> 3);
                const uint32_t v = ((lds_u8(a) | (lds_u8(a + 1u) << 8)) >> ((tid * 5u) & 7u)) & 31u;
                const uint32_t p1 = v / 5u;
                mv = v < 25u ? ((p1 + 1u) | ((v - 5u * p1 + 1u) << 8)) : 0u;
            } else if (CMD == CMD_NIBBLES) {
                const uint32_t c = lds_u8(b1 + OFF_MOVES);
                mv = (c & 15u) | ((c >> 4) << 8);
            } else {
                mv = lds_u16(b2 + OFF_MOVES);
            }
            int res = status;
            // this game's four record slots in the stage (generic pointer into shared memory)
            EvSink<EV> ev{EV ? reinterpret_cast<uint2*>(smem + s * STAGE_BYTES + OFF_EV) + tid * ORX_MAX_EVENTS_BASE : nullptr, 0,
                          ORX_MAX_EVENTS_BASE};
            // this game's NPC slots in the stage
            using NV = std::conditional_t<NPC != 0, NpcStage<NPC>, NoNpc>;
            NV nv;
            if constexpr (NPC != 0) {
                const uint32_t sb = stage0 + s * STAGE_BYTES;
                nv.pos = sb + OFF_NPOS + tid * 2u * NPC; nv.hp = sb + OFF_NHP + tid * 2u * NPC; nv.depth = sb + OFF_NDEPTH + tid * 4u * NPC;
            }
            if (status == ORX_RESULT_IN_PROGRESS) {          // finished lanes are frozen until reset
                Stream rs = make_stream(P, lane, ep);
                const uint4 blk = draw_block(rs, DOM_TICK, SUB_MAIN, (uint32_t)tick);
                if (CMD == CMD_BYTES_BOTS) {      // bots = kind of p1 | kind of p2 << 8; words 0 / 1 of the block are the RandomBots' draws
                    if ((bots & 255) != ORX_BOT_NONE) mv = (mv & 0xFF00u) | bot_move(bots & 255, L.pos & 0xFFFFu, L.st & 0xFFFFu, blk.x);
                    if ((bots >> 8) != ORX_BOT_NONE) mv = (mv & 0x00FFu) | (bot_move(bots >> 8, L.pos >> 16, L.st >> 16, blk.y) << 8);
                }
                Counters cnt{};
                res = tick_lane<DGEN, NV, EV>(P, tiles, lut, L, mv, blk.z, rs, nv, ev, cnt);
                int new_status = res;
                if (res != ORX_RESULT_IN_PROGRESS && P.auto_reset) {
                    rs.episode += 1;
                    reset_lane<DGEN, NV>(P, L, rs, nv);
                    new_status = ORX_RESULT_IN_PROGRESS;
                }
                sts_u32(b4 + OFF_POS, L.pos);
                sts_u32(b4 + OFF_HP, ((uint32_t)L.hp1 & 0xFFFFu) | ((uint32_t)L.hp2 << 16));
                sts_u32(b4 + OFF_ST, L.st);
                sts_u32(b4 + OFF_TICK, (uint32_t)L.tick);
                sts_u32(b4 + OFF_EP, L.episode);
                sts_s32x2(b8 + OFF_DEPTH, L.d1, L.d2);
                sts_u8(b1 + OFF_STATUS, (uint32_t)new_status);
            }
            if (CMD == CMD_BITS) {
                // 2 bits per game (result - 1): two ballots give the low and the high bit of the warp's 32 games,
                // lanes 0 / 1 interleave one 16-game half each into a word of the CTA's result block
                const uint32_t r = (uint32_t)(res - 1) & 3u;
                const uint32_t lo = __ballot_sync(0xffffffffu, (r & 1u) != 0u), hi = __ballot_sync(0xffffffffu, (r & 2u) != 0u);
                const uint32_t half = tid & 1u;
                auto spread = [](uint32_t x) {
                    x = (x | (x << 8)) & 0x00FF00FFu; x = (x | (x << 4)) & 0x0F0F0F0Fu;
                    x = (x | (x << 2)) & 0x33333333u; return (x | (x << 1)) & 0x55555555u;
                };
                const uint32_t word = spread((lo >> (16u * half)) & 0xFFFFu) | (spread((hi >> (16u * half)) & 0xFFFFu) << 1);
                if ((tid & 31u) < 2u) sts_u32(res0 + it * kResBitsTile + (tid >> 5) * 8u + half * 4u, word);
            } else {
                sts_u8(b1 + OFF_RESULT, (uint32_t)res);
            }
            ev.finish();                             // unused slots (all four of a frozen lane) read ORX_EV_NONE
        }
        if (OBS) {                                   // what each player sees of the state as it now is
            uint32_t w[12];
            pack_obs(L, obs_radius, w);
            const uint32_t bo = b4 + tid * 44u + OFF_OBS;      // 48 bytes per game
            sts_u32x4(bo, w[0], w[1], w[2], w[3]);
            sts_u32x4(bo + 16u, w[4], w[5], w[6], w[7]);
            sts_u32x4(bo + 32u, w[8], w[9], w[10], w[11]);
        }
        if (tid == 0 && it == 0) ORX_TRACE(trace_slot, 22);          // first tile ticked (warp 0)
        ORX_JITTER(400u + it);
        fence_proxy_async();                 // generic-proxy writes -> visible to the bulk-store engine
        __syncwarp();
        if ((tid & 31u) == 0) mbar_arrive(done0 + 8 * s);
    }
    if (tid == 0) ORX_TRACE(trace_slot, 6);
}

template <bool OBS, bool EV = false, int NPC = 0, bool BITS = false>
constexpr size_t pipe_smem_bytes(int fixed_tiles)
{
    return (size_t)kPipeStages<OBS, EV, NPC> * kPipeStageBytes<OBS, EV, NPC> + 2 * kPipeStages<OBS, EV, NPC> * 8 + kTileIdxBytes + kTicketBytes +
           (BITS ? kBitsSmemBytes : 0u) + (size_t)fixed_tiles;
}

}  // namespace orx
