// nezha_b200/csrc/nz_crc.cuh — CRC-32C (Castagnoli) / CRC-32 (IEEE) device code
// shared by nz_kernels.cu (K3, verify) and nz_inflate.cu (gzip trailer check).
//
// GF(2) helpers are zlib's crc32_combine arithmetic (multmodp / x2nmodp)
// re-derived for both polynomials, reflected bit order: x^0 = 0x80000000.
// POLY index 0 = Castagnoli 0x82F63B78, 1 = IEEE 0xEDB88320.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "nz_internal.h"

#define NZ_POLY_C    0x82F63B78u
#define NZ_POLY_IEEE 0xEDB88320u

// x^(8 * 2^k) mod P for k = 0..47; one copy per translation unit, filled by
// nz_crc_upload_constants() of that unit.
static __constant__ uint32_t c_x8pow[2][48];

template <int POLY>
__host__ __device__ __forceinline__ uint32_t gf_mul(uint32_t a, uint32_t b) {
    constexpr uint32_t P = POLY ? NZ_POLY_IEEE : NZ_POLY_C;
    uint32_t p = 0;
    for (int i = 0; i < 32; ++i) {
        p ^= b & (0u - ((a >> (31 - i)) & 1u));
        b = (b >> 1) ^ (P & (0u - (b & 1u)));
    }
    return p;
}

template <int POLY>
__device__ __forceinline__ uint32_t gf_xpow8(uint64_t nbytes) {   // x^(8*nbytes)
    uint32_t p = 0x80000000u;
    int k = 0;
    while (nbytes) {
        if (nbytes & 1) p = gf_mul<POLY>(c_x8pow[POLY][k], p);
        nbytes >>= 1;
        ++k;
    }
    return p;
}

__device__ __forceinline__ uint4 crc_ldg16(const void* p) {
    uint4 r;
    asm volatile("ld.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}

// CRC of one chunk (<= 64 KiB) by a 256-thread CTA: thread t owns bytes
// [256 t, 256 t + 256), slice-by-4 from shared tables (sT: 4 x 256 u32 of the
// chosen polynomial), then the per-thread CRCs are shifted to the chunk end
// (multiply by x^(8*bytes_after)) and XOR-reduced.  sred: 8 u32.  All 256
// threads must call it.
template <int POLY>
__device__ uint32_t block_crc_chunk(const uint8_t* __restrict__ p, uint32_t m, const uint32_t* sT, uint32_t* sred) {
    const uint32_t tid = threadIdx.x;
    const uint32_t beg = tid * 256u;
    uint32_t contrib = 0;
    if (beg < m) {
        const uint32_t cnt = min(256u, m - beg);
        const uint8_t* q = p + beg;
        uint32_t crc = 0xFFFFFFFFu;
        uint32_t i = 0;
        if ((((uintptr_t)q) & 15u) == 0) {
            for (; i + 16 <= cnt; i += 16) {
                const uint4 v = crc_ldg16(q + i);
                const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    crc ^= w[j];
                    crc = sT[3 * 256 + (crc & 0xff)] ^ sT[2 * 256 + ((crc >> 8) & 0xff)] ^
                          sT[1 * 256 + ((crc >> 16) & 0xff)] ^ sT[crc >> 24];
                }
            }
        }
        for (; i < cnt; ++i) crc = (crc >> 8) ^ sT[(crc ^ q[i]) & 0xff];
        crc = ~crc;
        const uint32_t after = m - (beg + cnt);
        contrib = after ? gf_mul<POLY>(gf_xpow8<POLY>(after), crc) : crc;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) contrib ^= __shfl_xor_sync(0xffffffffu, contrib, o);
    __syncthreads();   // sred reuse across calls
    if ((tid & 31u) == 0) sred[tid >> 5] = contrib;
    __syncthreads();
    uint32_t r = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) r ^= sred[i];
    return r;
}

// fills this translation unit's c_x8pow
static inline cudaError_t nz_crc_upload_constants() {
    uint32_t h[2][48];
    uint32_t p = 0x80000000u >> 8;   // x^8
    for (int k = 0; k < 48; ++k) { h[0][k] = p; p = gf_mul<0>(p, p); }
    p = 0x80000000u >> 8;
    for (int k = 0; k < 48; ++k) { h[1][k] = p; p = gf_mul<1>(p, p); }
    return cudaMemcpyToSymbol(c_x8pow, h, sizeof h);
}
