// tools/probe_scatter_write.cu — what HBM write rate does the coalesced gather's access pattern allow?
// A persistent grid writes 32 KiB tiles of non-constant data to M destination buffers (tile t of every
// buffer by the same CTA, like a coalesced group of M GETs), with plain 16 B stores and with bulk
// (TMA) stores from shared memory.  M = 1 is a plain sequential fill.  No reads.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o nezha_b200/_build/probe_scatter_write tools/probe_scatter_write.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

constexpr uint32_t TILE = 32768;

__global__ void __launch_bounds__(256) k_st(uint8_t* const* dst, int M, uint64_t tiles) {
    for (uint64_t t = blockIdx.x; t < tiles; t += gridDim.x) {
        uint4 v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) v[u] = make_uint4((uint32_t)t, threadIdx.x, u, (uint32_t)(t >> 7) * 2654435761u);
        for (int m = 0; m < M; ++m) {
            uint4* d = reinterpret_cast<uint4*>(dst[m] + t * TILE);
#pragma unroll
            for (int u = 0; u < 8; ++u)
                asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};" :: "l"(d + u * 256 + threadIdx.x), "r"(v[u].x), "r"(v[u].y), "r"(v[u].z), "r"(v[u].w) : "memory");
        }
    }
}

__global__ void __launch_bounds__(128) k_bulk(uint8_t* const* dst, int M, uint64_t tiles) {
    extern __shared__ __align__(128) uint8_t sb[];
    for (uint32_t i = threadIdx.x; i < TILE / 16; i += 128) reinterpret_cast<uint4*>(sb)[i] = make_uint4(i, blockIdx.x, i * 7, i ^ 0x5555);
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        const uint32_t buf = (uint32_t)__cvta_generic_to_shared(sb);
        for (uint64_t t = blockIdx.x; t < tiles; t += gridDim.x) {
            for (int m = 0; m < M; ++m)
                asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" :: "l"(dst[m] + t * TILE), "r"(buf), "r"(TILE) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            asm volatile("cp.async.bulk.wait_group.read 4;" ::: "memory");
        }
        asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
}

// read one tile (bulk load, 2 alternating stages), write it to M buffers: the coalesced gather in miniature
// mode 0: source tile t from DRAM; 1: source always in L2 (the same 2 MiB over and over: is it the DRAM READS that cost?);
// 2: DRAM source, but pulled into L2 in 2 MiB bulk prefetches 32 MiB ahead of use (long read bursts instead of one tile at a time)
__global__ void __launch_bounds__(128) k_copy(const uint8_t* src, uint8_t* const* dst, int M, uint64_t tiles, int mode) {
    extern __shared__ __align__(128) uint8_t sb[];
    __shared__ __align__(8) uint64_t bar[2];
    if (threadIdx.x == 0) {
        for (int i = 0; i < 2; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"((uint32_t)__cvta_generic_to_shared(&bar[i])));
        asm volatile("fence.mbarrier_init.release.cluster;");
        uint32_t phase = 0, slot = 0;
        for (uint64_t t = blockIdx.x; t < tiles; t += gridDim.x) {
            const uint32_t buf = (uint32_t)__cvta_generic_to_shared(sb + slot * TILE), b = (uint32_t)__cvta_generic_to_shared(&bar[slot]);
            asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
            const uint8_t* sp = src + (mode == 1 ? (t & 63) : t) * TILE;
            if (mode == 2 && (t & 63) == 0) {                      // this CTA drew the first tile of a 2 MiB chunk: prefetch the chunk 16 chunks ahead
                const uint64_t pt = t + 16 * 64;
                if (pt < tiles)
                    for (int k = 0; k < 64; ++k)
                        asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" :: "l"(src + (pt + k) * TILE), "r"(TILE) : "memory");
            }
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(b), "r"(TILE) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" :: "r"(buf), "l"(sp), "r"(TILE), "r"(b) : "memory");
            uint32_t done = 0;
            while (!done) asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(done) : "r"(b), "r"((phase >> slot) & 1u) : "memory");
            phase ^= 1u << slot;
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            for (int m = 0; m < M; ++m)
                asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" :: "l"(dst[m] + t * TILE), "r"(buf), "r"(TILE) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            slot ^= 1;
        }
        asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
}

// three stages of TB bytes, the NEXT tile's load issued before the current tile's stores: in the SM's TMA queue the load sits
// ahead of 32 stores instead of behind them, so it has landed by the time the stores have been issued
template <uint32_t TB>
__global__ void __launch_bounds__(128) k_copy_ahead(const uint8_t* src, uint8_t* const* dst, int M, uint64_t tiles, int l2src) {
    extern __shared__ __align__(128) uint8_t sb[];
    __shared__ __align__(8) uint64_t bar[3];
    if (threadIdx.x == 0) {
        for (int i = 0; i < 3; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"((uint32_t)__cvta_generic_to_shared(&bar[i])));
        asm volatile("fence.mbarrier_init.release.cluster;");
        uint32_t phase = 0, slot = 0;
        auto issue = [&](uint64_t t, uint32_t sl) {
            const uint32_t buf = (uint32_t)__cvta_generic_to_shared(sb + sl * TB), b = (uint32_t)__cvta_generic_to_shared(&bar[sl]);
            const uint8_t* sp = src + (l2src ? (t & 63) : t) * (uint64_t)TB;
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(b), "r"(TB) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" :: "r"(buf), "l"(sp), "r"(TB), "r"(b) : "memory");
        };
        uint64_t t = blockIdx.x;
        if (t < tiles) issue(t, 0);
        for (; t < tiles; t += gridDim.x) {
            const uint32_t nslot = slot == 2 ? 0 : slot + 1;
            if (t + gridDim.x < tiles) {                       // stage nslot was read by the stores two tiles back
                asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
                issue(t + gridDim.x, nslot);
            }
            const uint32_t buf = (uint32_t)__cvta_generic_to_shared(sb + slot * TB), b = (uint32_t)__cvta_generic_to_shared(&bar[slot]);
            uint32_t done = 0;
            while (!done) asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(done) : "r"(b), "r"((phase >> slot) & 1u) : "memory");
            phase ^= 1u << slot;
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            for (int m = 0; m < M; ++m)
                asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" :: "l"(dst[m] + t * (uint64_t)TB), "r"(buf), "r"(TB) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            slot = nslot;
        }
        asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
}

int main(int argc, char** argv) {
    const uint64_t total = (argc > 1 ? strtoull(argv[1], nullptr, 10) : 32ull) << 30;
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    uint8_t* pool = nullptr;
    if (cudaMalloc(&pool, total) != cudaSuccess) { fprintf(stderr, "cudaMalloc failed\n"); return 1; }
    cudaFuncSetAttribute(k_bulk, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TILE);
    cudaFuncSetAttribute(k_copy, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(2 * TILE));
    uint8_t* srcbuf = nullptr;
    cudaMalloc(&srcbuf, 1ull << 30);
    cudaMemset(srcbuf, 0x5a, 1ull << 30);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    const bool only32 = argc > 2;
    for (int M : {1, 8, 32, 64}) {
        if (only32 && M != 32) continue;
        const uint64_t per = total / M > (1ull << 30) ? (1ull << 30) : total / M, tiles = per / TILE;
        std::vector<uint8_t*> h(M);
        for (int m = 0; m < M; ++m) h[m] = pool + (uint64_t)m * per;
        uint8_t** d = nullptr;
        cudaMalloc(&d, M * sizeof(uint8_t*));
        cudaMemcpy(d, h.data(), M * sizeof(uint8_t*), cudaMemcpyHostToDevice);
        for (int kind = 0; kind < 5; ++kind) {
            for (int cps : {2, 3}) {
                float best = 1e30f;
                for (int rep = 0; rep < 4; ++rep) {
                    cudaEventRecord(e0);
                    if (kind == 0) k_st<<<sms * cps, 256>>>(d, M, tiles);
                    else if (kind == 1) k_bulk<<<sms * cps, 128, TILE>>>(d, M, tiles);
                    else k_copy<<<sms * cps, 128, 2 * TILE>>>(srcbuf, d, M, tiles, kind - 2);
                    cudaEventRecord(e1);
                    cudaEventSynchronize(e1);
                    float ms = 0;
                    cudaEventElapsedTime(&ms, e0, e1);
                    if (rep && ms < best) best = ms;
                }
                printf("M=%2d %-5s %d CTAs/SM: %7.3f ms  %7.1f GB/s written%s\n", M, kind == 0 ? "st.v4" : kind == 1 ? "bulk" : kind == 2 ? "copy" : kind == 3 ? "copyL2" : "copyPF", cps, best,
                       (double)tiles * TILE * M / (best * 1e-3) / 1e9, kind >= 2 ? " (+ 1 tile read per M written)" : "");
            }
        }
        if (M == 32) {
            auto run_ahead = [&](auto kern, uint32_t tb, const char* name, int l2src) {
                cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(3 * tb));
                const uint64_t tl = per / tb;
                for (int cps : {2, 3}) {
                    float best = 1e30f;
                    for (int rep = 0; rep < 4; ++rep) {
                        cudaEventRecord(e0);
                        kern<<<sms * cps, 128, 3 * tb>>>(srcbuf, d, M, tl, l2src);
                        cudaEventRecord(e1);
                        cudaEventSynchronize(e1);
                        float ms = 0;
                        cudaEventElapsedTime(&ms, e0, e1);
                        if (rep && ms < best) best = ms;
                    }
                    printf("M=%2d %-12s %d CTAs/SM: %7.3f ms  %7.1f GB/s written (+ 1 tile read per M written)\n", M, name, cps, best, (double)tl * tb * M / (best * 1e-3) / 1e9);
                }
            };
            run_ahead(k_copy_ahead<24576>, 24576, "ahead3x24K", 0);
            run_ahead(k_copy_ahead<32768>, 32768, "ahead3x32K", 0);
            run_ahead(k_copy_ahead<16384>, 16384, "ahead3x16K", 0);
            run_ahead(k_copy_ahead<24576>, 24576, "ahead3x24K-L2", 1);
        }
        cudaFree(d);
    }
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { fprintf(stderr, "error: %s\n", cudaGetErrorString(e)); return 1; }
    return 0;
}
