/* The drop-in boundary without Python or PyTorch: a plain C host program against include/orx.h and the CUDA runtime.
 * It plays the reference's server loop (optimax_rogue/server/main.py:110-113: on_tick + Updater.update, fed by
 * StaircaseBot vs RandomBot, optimax_rogue_bots/) for N games on one GPU -- orx_reset, then orx_bot_moves + orx_step per
 * tick, then a fused orx_rollout -- and prints counters a caller can check (tests/test_gpu_c_abi_demo.py compares them
 * with the Python host's for the same seed).
 *
 *   gcc -O2 -I include -I /usr/local/cuda/include examples/c_abi_demo.c -o c_abi_demo \
 *       -L optimax_rogue_b200 -l:liborx.so -L /usr/local/cuda/lib64 -lcudart -Wl,-rpath,$PWD/optimax_rogue_b200
 *   ./c_abi_demo [games] [ticks] [seed] [max_ticks]
 */
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "orx.h"

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_)); return 2; } } while (0)
#define OK(x) do { int rc_ = (x); if (rc_ != ORX_OK) { fprintf(stderr, "%s: %s\n", #x, orx_strerror(rc_)); return 3; } } while (0)

static void* dalloc(size_t bytes)
{
    void* p = NULL;
    if (cudaMalloc(&p, bytes) != cudaSuccess || cudaMemset(p, 0, bytes) != cudaSuccess) { fprintf(stderr, "cudaMalloc(%zu) failed\n", bytes); exit(2); }
    return p;
}

int main(int argc, char** argv)
{
    const int64_t n = argc > 1 ? atoll(argv[1]) : 4096;
    const int ticks = argc > 2 ? atoi(argv[2]) : 100;
    const uint64_t seed = argc > 3 ? strtoull(argv[3], NULL, 0) : 7;
    const int max_ticks = argc > 4 ? atoi(argv[4]) : 1000;
    if (orx_abi_version() != ORX_ABI_VERSION) { fprintf(stderr, "liborx ABI %d != header %d\n", orx_abi_version(), ORX_ABI_VERSION); return 1; }

    OrxConfig cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.struct_size = sizeof(cfg);
    cfg.width = 60; cfg.height = 10;                       /* server/main.py:27-28 */
    cfg.dgen_kind = ORX_DGEN_EMPTY; cfg.start_kind = ORX_START_TOGETHER;
    cfg.despawn_strat = ORX_DESPAWN_UNREACHABLE; cfg.max_ticks = max_ticks;
    cfg.hp[0] = cfg.hp[1] = 10; cfg.damage[0] = cfg.damage[1] = 2; cfg.armor[0] = cfg.armor[1] = 1;   /* worldgen.py:85-86 */
    cfg.auto_reset = 1; cfg.seed = seed;
    cfg.fixed_stairs[0] = cfg.fixed_stairs[1] = ORX_NO_STAIRS;

    /* structure of arrays, caller-owned device memory; the five 4-byte planes at one pitch (the layout hint of orx.h) */
    OrxState st;
    memset(&st, 0, sizeof(st));
    const size_t pitch = ((size_t)(4 * n) + 127) / 128 * 128;
    uint8_t* words = (uint8_t*)dalloc(5 * pitch);
    st.pos = (void*)words; st.hp = (void*)(words + pitch); st.stairs = (void*)(words + 2 * pitch);
    st.tick = (void*)(words + 3 * pitch); st.episode = (void*)(words + 4 * pitch);
    st.depth = dalloc(8 * (size_t)n); st.status = dalloc((size_t)n);
    st.npc_pos = dalloc(2 * (size_t)n); st.npc_hp = dalloc(2 * (size_t)n); st.npc_depth = dalloc(4 * (size_t)n);
    st.sched_words = (uint32_t)orx_sched_words(n);
    st.sched = dalloc(4 * (size_t)st.sched_words);
    uint8_t* moves = (uint8_t*)dalloc(2 * (size_t)n);
    uint8_t* result = (uint8_t*)dalloc((size_t)n);
    unsigned long long* stats = (unsigned long long*)dalloc(8 * ORX_STAT_COUNT);
    uint8_t* h_result = (uint8_t*)malloc((size_t)n);
    cudaStream_t s;
    CK(cudaStreamCreate(&s));

    OK(orx_reset(&cfg, &st, NULL, 0, n, 0, s));
    unsigned long long finished[5] = {0, 0, 0, 0, 0}, checksum = 0;
    for (int t = 0; t < ticks; ++t) {
        OK(orx_bot_moves(&cfg, &st, ORX_BOT_STAIRCASE, ORX_BOT_RANDOM, moves, n, 0, s));
        OK(orx_step(&cfg, &st, moves, result, NULL, n, 0, s));
        CK(cudaMemcpyAsync(h_result, result, (size_t)n, cudaMemcpyDeviceToHost, s));
        CK(cudaStreamSynchronize(s));
        for (int64_t i = 0; i < n; ++i) {
            finished[h_result[i] <= 4 ? h_result[i] : 0]++;
            checksum = checksum * 1099511628211ULL + h_result[i];
        }
    }
    OK(orx_rollout(&cfg, &st, ORX_BOT_STAIRCASE, ORX_BOT_RANDOM, ticks, stats, n, 0, s));
    unsigned long long h_stats[ORX_STAT_COUNT];
    CK(cudaMemcpyAsync(h_stats, stats, sizeof(h_stats), cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    printf("games=%lld ticks=%d seed=%llu\n", (long long)n, ticks, (unsigned long long)seed);
    printf("step: in_progress=%llu p1_wins=%llu p2_wins=%llu ties=%llu checksum=%llu\n", finished[ORX_RESULT_IN_PROGRESS], finished[ORX_RESULT_PLAYER1_WIN],
           finished[ORX_RESULT_PLAYER2_WIN], finished[ORX_RESULT_TIE], checksum);
    printf("rollout:");
    for (int k = 0; k < ORX_STAT_COUNT; ++k) printf(" %llu", h_stats[k]);
    printf("\n");
    return 0;
}
