// Tuning aid (not product): what does the k_step ACCESS PATTERN sustain with no game logic?
// Same planes, same widths, one thread per game (v1) or 4 games per thread with 16-byte accesses (v4).
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
struct Planes { uint32_t *pos, *hp, *st, *tick, *ep; int2* depth; uint8_t *status, *result; uint16_t* moves; };
__global__ void __launch_bounds__(256) copy_v1(Planes p, unsigned n)
{
    unsigned i = blockIdx.x * 256 + threadIdx.x;
    if (i >= n) return;
    uint32_t a = p.pos[i], b = p.hp[i], c = p.st[i], d = p.tick[i], e = p.ep[i];
    int2 f = p.depth[i];
    uint32_t s = p.status[i], m = p.moves[i];
    p.pos[i] = a + m; p.hp[i] = b; p.st[i] = c; p.tick[i] = d + 1; p.ep[i] = e; p.depth[i] = f;
    p.status[i] = (uint8_t)s; p.result[i] = (uint8_t)(s + m);
}
__global__ void __launch_bounds__(256) copy_v4(Planes p, unsigned n4)
{
    unsigned i = blockIdx.x * 256 + threadIdx.x;
    if (i >= n4) return;
    uint4 a = ((uint4*)p.pos)[i], b = ((uint4*)p.hp)[i], c = ((uint4*)p.st)[i], d = ((uint4*)p.tick)[i], e = ((uint4*)p.ep)[i];
    uint4 f0 = ((uint4*)p.depth)[2 * i], f1 = ((uint4*)p.depth)[2 * i + 1];
    uint32_t s = ((uint32_t*)p.status)[i];
    uint2 m = ((uint2*)p.moves)[i];
    a.x += m.x; d.x += 1;
    ((uint4*)p.pos)[i] = a; ((uint4*)p.hp)[i] = b; ((uint4*)p.st)[i] = c; ((uint4*)p.tick)[i] = d; ((uint4*)p.ep)[i] = e;
    ((uint4*)p.depth)[2 * i] = f0; ((uint4*)p.depth)[2 * i + 1] = f1;
    ((uint32_t*)p.status)[i] = s; ((uint32_t*)p.result)[i] = s + m.y;
}
int main(int argc, char** argv)
{
    unsigned G = argc > 1 ? atoi(argv[1]) : (1u << 20);
    int nb = 10, steps = 200;
    Planes P[16];
    for (int b = 0; b < nb; ++b) {
        cudaMalloc(&P[b].pos, 4ull * G); cudaMalloc(&P[b].hp, 4ull * G); cudaMalloc(&P[b].st, 4ull * G);
        cudaMalloc(&P[b].tick, 4ull * G); cudaMalloc(&P[b].ep, 4ull * G); cudaMalloc(&P[b].depth, 8ull * G);
        cudaMalloc(&P[b].status, G); cudaMalloc(&P[b].result, G); cudaMalloc(&P[b].moves, 2ull * G);
        cudaMemset(P[b].pos, 1, 4ull * G); cudaMemset(P[b].moves, 1, 2ull * G); cudaMemset(P[b].status, 1, G);
    }
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int v = 0; v < 2; ++v) {
        for (int rep = 0; rep < 3; ++rep) {
            cudaEventRecord(e0);
            for (int k = 0; k < steps; ++k) {
                if (v == 0) copy_v1<<<(G + 255) / 256, 256>>>(P[k % nb], G);
                else copy_v4<<<(G / 4 + 255) / 256, 256>>>(P[k % nb], G / 4);
            }
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            printf("copy_v%d G=%u: %.2f us/step, %.0f GB/s (61 B/game)\n", v == 0 ? 1 : 4, G, ms * 1e3 / steps, 61.0 * G / (ms / steps) / 1e6);
        }
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
