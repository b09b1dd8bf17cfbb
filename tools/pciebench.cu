// Tuning aid (not product): how fast can a kernel pull a small command stream out of PINNED HOST memory, as a
// function of how it asks? total bytes = what one 2^20-game tick needs (655,360 B of cmd5) and 1/8 of it.
//   bulk   cp.async.bulk (TMA) of `chunk` bytes per request, `inflight` requests per CTA, `ctas` CTAs
//   ldg    plain 16-byte loads, one per thread (coalesced 512 B per warp)
// and the copy engine (cudaMemcpyAsync) for the same bytes, both directions.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/pciebench tools/pciebench.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>
#include <algorithm>

__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void __launch_bounds__(32) k_bulk(const uint8_t* __restrict__ src, uint32_t total, uint32_t chunk, unsigned long long* sink)
{
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ uint64_t bar;
    if (threadIdx.x != 0) return;
    const uint32_t b = smem_addr(&bar), dst = smem_addr(smem);
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    // this CTA's contiguous share, requested chunk by chunk, all requests in flight at once
    const uint32_t per = ((total / gridDim.x) + 15u) & ~15u;
    const uint32_t lo = blockIdx.x * per, hi = min(total, lo + per);
    if (lo >= hi) return;
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(hi - lo) : "memory");
    for (uint32_t o = lo; o < hi; o += chunk) {
        const uint32_t n = min(chunk, hi - o);
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(dst + (o - lo)), "l"(src + o), "r"(n), "r"(b) : "memory");
    }
    asm volatile("{\n.reg .pred p;\nW:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n@p bra D;\nbra W;\nD:\n}\n" ::"r"(b) : "memory");
    if (smem[0] == 123 && sink) *sink = 1;
}

__global__ void __launch_bounds__(256) k_ldg(const uint4* __restrict__ src, uint32_t n16, unsigned long long* sink)
{
    const uint32_t i = blockIdx.x * 256 + threadIdx.x;
    if (i >= n16) return;
    const uint4 v = src[i];
    if (v.x == 0x12345678u && v.y == 1 && sink) *sink = v.z;
}

static float time_ms(cudaStream_t s, int reps, auto&& fn)
{
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    fn(); cudaStreamSynchronize(s);
    std::vector<float> t;
    for (int r = 0; r < reps; ++r) {
        cudaEventRecord(a, s); fn(); cudaEventRecord(b, s); cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b); t.push_back(ms);
    }
    std::sort(t.begin(), t.end());
    return t[t.size() / 2];
}

int main()
{
    cudaStream_t s; cudaStreamCreate(&s);
    const uint32_t full = 655360;
    uint8_t *h, *hd, *d; unsigned long long* sink;
    cudaHostAlloc(&h, 4 << 20, cudaHostAllocMapped); cudaHostGetDevicePointer(&hd, h, 0);
    cudaMalloc(&d, 4 << 20); cudaMalloc(&sink, 8);
    for (int i = 0; i < (4 << 20); ++i) h[i] = (uint8_t)(i * 7);
    cudaFuncSetAttribute(k_bulk, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    for (uint32_t total : {full, full / 8}) {
        printf("== %u bytes\n", total);
        for (uint32_t ctas : {37u, 74u, 148u, 296u, 444u, 888u}) {
            const uint32_t per = ((total / ctas) + 15u) & ~15u;
            if (per > 60 * 1024) continue;
            for (uint32_t chunk : {128u, 256u, 512u, 1024u, 4096u, 16384u, 65536u}) {
                if (chunk > per && chunk != 65536u) continue;
                const float ms = time_ms(s, 21, [&] { k_bulk<<<ctas, 32, per + 128, s>>>(hd, total, chunk, sink); });
                printf("bulk  ctas=%4u per_cta=%6u chunk=%6u : %7.2f us  %6.1f GB/s\n", ctas, per, chunk == 65536u ? per : chunk, ms * 1e3, total / ms / 1e6);
            }
        }
        {
            const uint32_t n16 = total / 16;
            const float ms = time_ms(s, 21, [&] { k_ldg<<<(n16 + 255) / 256, 256, 0, s>>>((const uint4*)hd, n16, sink); });
            printf("ldg   16 B per thread                  : %7.2f us  %6.1f GB/s\n", ms * 1e3, total / ms / 1e6);
        }
        {
            const float ms = time_ms(s, 21, [&] { cudaMemcpyAsync(d, h, total, cudaMemcpyHostToDevice, s); });
            printf("copy engine H2D                        : %7.2f us  %6.1f GB/s\n", ms * 1e3, total / ms / 1e6);
            const float ms2 = time_ms(s, 21, [&] { cudaMemcpyAsync(h, d, total * 2 / 5, cudaMemcpyDeviceToHost, s); });
            printf("copy engine D2H (%u B)             : %7.2f us  %6.1f GB/s\n", total * 2 / 5, ms2 * 1e3, total * 0.4 / ms2 / 1e6);
            for (int parts : {2, 4, 8}) {
                const float ms3 = time_ms(s, 21, [&] { for (int p = 0; p < parts; ++p) cudaMemcpyAsync(d + (size_t)p * (total / parts), h + (size_t)p * (total / parts), total / parts, cudaMemcpyHostToDevice, s); });
                printf("copy engine H2D in %d parts             : %7.2f us\n", parts, ms3 * 1e3);
            }
        }
        {
            const float ms = time_ms(s, 21, [&] { k_ldg<<<1, 32, 0, s>>>((const uint4*)d, 1, sink); });
            printf("empty kernel launch+event               : %7.2f us\n", ms * 1e3);
        }
    }
    // host-side latency of launch + stream sync (what a synchronous host loop pays per tick)
    {
        cudaStreamSynchronize(s);
        timespec t0, t1;
        clock_gettime(CLOCK_MONOTONIC, &t0);
        for (int i = 0; i < 2000; ++i) { k_ldg<<<1, 32, 0, s>>>((const uint4*)d, 1, sink); cudaStreamSynchronize(s); }
        clock_gettime(CLOCK_MONOTONIC, &t1);
        printf("launch(empty kernel) + cudaStreamSynchronize: %.2f us per iteration (host clock)\n", ((t1.tv_sec - t0.tv_sec) * 1e9 + (t1.tv_nsec - t0.tv_nsec)) / 2000 / 1e3);
        // other ways to wait for the same kernel: spinning on cudaStreamQuery, on an event, and on a word the kernel writes
        clock_gettime(CLOCK_MONOTONIC, &t0);
        for (int i = 0; i < 2000; ++i) { k_ldg<<<1, 32, 0, s>>>((const uint4*)d, 1, sink); while (cudaStreamQuery(s) == cudaErrorNotReady) { } }
        clock_gettime(CLOCK_MONOTONIC, &t1);
        printf("launch(empty kernel) + spin on cudaStreamQuery: %.2f us per iteration\n", ((t1.tv_sec - t0.tv_sec) * 1e9 + (t1.tv_nsec - t0.tv_nsec)) / 2000 / 1e3);
        cudaEvent_t ev;
        cudaEventCreateWithFlags(&ev, cudaEventDisableTiming);
        clock_gettime(CLOCK_MONOTONIC, &t0);
        for (int i = 0; i < 2000; ++i) { k_ldg<<<1, 32, 0, s>>>((const uint4*)d, 1, sink); cudaEventRecord(ev, s); while (cudaEventQuery(ev) == cudaErrorNotReady) { } }
        clock_gettime(CLOCK_MONOTONIC, &t1);
        printf("launch(empty kernel) + event record + spin on cudaEventQuery: %.2f us per iteration\n", ((t1.tv_sec - t0.tv_sec) * 1e9 + (t1.tv_nsec - t0.tv_nsec)) / 2000 / 1e3);
        clock_gettime(CLOCK_MONOTONIC, &t0);
        for (int i = 0; i < 2000; ++i) { cudaStreamSynchronize(s); }
        clock_gettime(CLOCK_MONOTONIC, &t1);
        printf("cudaStreamSynchronize on an idle stream: %.2f us per call\n", ((t1.tv_sec - t0.tv_sec) * 1e9 + (t1.tv_nsec - t0.tv_nsec)) / 2000 / 1e3);
        clock_gettime(CLOCK_MONOTONIC, &t0);
        for (int i = 0; i < 2000; ++i) { k_ldg<<<1, 32, 0, s>>>((const uint4*)d, 1, sink); }
        clock_gettime(CLOCK_MONOTONIC, &t1);
        cudaStreamSynchronize(s);
        printf("launch(empty kernel) alone, host time per call: %.2f us\n", ((t1.tv_sec - t0.tv_sec) * 1e9 + (t1.tv_nsec - t0.tv_nsec)) / 2000 / 1e3);
    }
    return 0;
}
