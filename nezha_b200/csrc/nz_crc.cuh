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

__device__ __forceinline__ uint4 crc_ldg16_stream(const void* p) {   // read-once input: keep it out of L1, where the lookup tables live
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
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

// ---------------------------------------------------------------------------
// Fast path for a FULL 64 KiB chunk on a 16 B aligned pointer: coalesced 16 B loads.
// Thread t reads vector t of each 4 KiB row (16 rows); each of its four 32-bit lanes keeps its own
// running remainder with the row stride folded in:  acc = Z_4096(acc) ^ word   (Z_k = "append k zero
// bytes", a linear map on 32 bits = 4 lookups in the stride table).  At the end the four lanes are
// merged with the ordinary slice tables (Z_4), the thread's value is moved to the chunk end with one
// GF(2) multiply by x^(8*(4096-16t-12)) and the CTA XOR-reduces.  The stride table is replicated
// 8x in shared memory ([entry][copy], copy = lane & 7) so that random indices spread over the banks.
//   fast tables in global memory (per polynomial): ZS[4*256] | K[256] | L[5][4*256] (warp-tree levels)
// ---------------------------------------------------------------------------
constexpr int NZ_CRC_LVL_OFF = 4 * 256 + 256;            // warp-tree level tables start here (5 x 4 x 256 words)
constexpr int NZ_CRC_FAST_WORDS = NZ_CRC_LVL_OFF + 5 * 4 * 256;
static_assert(NZ_CRC_FAST_WORDS == NZ_CRC_FAST_WORDS_DECL, "table block layout (nz_internal.h)");

// Stride table in shared memory, 8 copies, laid out [byte value][byte position k][copy]: word b*32 + k*8 + cp.
// A lane uses copy cp = lane & 7 and takes the four byte positions of a word in the rotated order k = (j + g) & 3,
// g = (lane >> 3) & 3 (XOR does not care about the order), so in every one of the four lookups the 32 lanes of a
// warp hit 32 different banks (bank = k*8 + cp) whatever the data: no bank conflicts.  (The first layout,
// [k*256 + b][copy], measured 2.3 wavefronts per lookup in the fused-verify gather: ncu, round 1.)
// One lookup = rotate the byte into bits 7..14 (SHF.R.W), mask and OR in the lane's constant part of the byte offset
// (LOP3), load (LDS [R + UR]): 3 instructions.  (The first form - shift, x32, mask, + lane offset, x4 + base - was 6
// and made the kernel issue bound at 3.4 TB/s: ncu, profiles/r02_k_crc_chunks_before.txt.)
#define NZ_ZS_SETUP(tid)                                                                                   \
    const uint32_t zs_cp = (tid) & 7u, zs_g = ((tid) >> 3) & 3u;                                           \
    const uint32_t zs_k0 = zs_g, zs_k1 = (zs_g + 1u) & 3u, zs_k2 = (zs_g + 2u) & 3u, zs_k3 = (zs_g + 3u) & 3u; \
    const uint32_t zs_r0 = (8u * zs_k0 + 25u) & 31u, zs_r1 = (8u * zs_k1 + 25u) & 31u, zs_r2 = (8u * zs_k2 + 25u) & 31u, zs_r3 = (8u * zs_k3 + 25u) & 31u; \
    const uint32_t zs_l0 = zs_k0 * 32u + zs_cp * 4u, zs_l1 = zs_k1 * 32u + zs_cp * 4u, zs_l2 = zs_k2 * 32u + zs_cp * 4u, zs_l3 = zs_k3 * 32u + zs_cp * 4u; \
    const char* const zs_b = reinterpret_cast<const char*>(sZS8);
#define NZ_ZS1(a, r, l) (*reinterpret_cast<const uint32_t*>(zs_b + ((__funnelshift_r((a), (a), (r)) & 0x7f80u) | (l))))
#define NZ_ZS(a) (NZ_ZS1(a, zs_r0, zs_l0) ^ NZ_ZS1(a, zs_r1, zs_l1) ^ NZ_ZS1(a, zs_r2, zs_l2) ^ NZ_ZS1(a, zs_r3, zs_l3))

// Sum over the lanes of a warp of v_l * x^(8*16*(31-l)), returned in lane 0: five shuffle levels, each "append
// 16 * 2^s zero bytes" on the left operand = 4 lookups in that level's table (global memory, L1 resident: 20 loads
// per thread and chunk against 256 in the row loop).  Replaces one 32-step GF(2) multiply PER THREAD (~160
// instructions, 0.6 per byte checksummed) by one per WARP.  gL: 5 x [4][256] words.
__device__ __forceinline__ uint32_t crc_warp_tree(uint32_t v, const uint32_t* __restrict__ gL) {
    const uint32_t lane = threadIdx.x & 31u;
#pragma unroll
    for (int s = 0; s < 5; ++s) {
        const uint32_t u = __shfl_down_sync(0xffffffffu, v, 1u << s);
        // only the lanes that carry a partial sum at this level look anything up: 16 + 8 + 4 + 2 + 1 lanes x 4 scattered
        // loads per warp, not 5 x 32 x 4 (each scattered load is a wavefront of its own in the L1 data stage - with every
        // lane loading, the tree cost more L1 bandwidth than the whole row loop: profiles/r02_k_crc_chunks_v2.txt)
        if ((lane & ((2u << s) - 1u)) == 0) {
            const uint32_t* T = gL + s * 1024;
            v = __ldg(T + (v & 0xffu)) ^ __ldg(T + 256 + ((v >> 8) & 0xffu)) ^ __ldg(T + 512 + ((v >> 16) & 0xffu)) ^ __ldg(T + 768 + (v >> 24)) ^ u;
        }
    }
    return v;
}

template <int POLY, bool TREE = true>
__device__ uint32_t block_crc_chunk64k(const uint8_t* __restrict__ p, const uint32_t* sZS8, const uint32_t* sT,
                                       const uint32_t* __restrict__ gK, uint32_t* sred) {
    const uint32_t tid = threadIdx.x;
    uint32_t a0 = 0, a1 = 0, a2 = 0, a3 = 0;
    const uint4* v = reinterpret_cast<const uint4*>(p) + tid;
    NZ_ZS_SETUP(tid)
    // rows in groups of four, the next group's loads issued before the current group is folded: eight 16 B loads per
    // thread in flight (the compiler kept two when left to itself and the kernel sat on long-scoreboard stalls)
    uint4 w[2][4];
#pragma unroll
    for (int j = 0; j < 4; ++j) w[0][j] = crc_ldg16_stream(v + j * 256);
    if (tid == 0) w[0][0].x ^= 0xFFFFFFFFu;                // the CRC's initial value, folded into the first word
#pragma unroll
    for (int g = 0; g < 4; ++g) {
        if (g < 3) {
#pragma unroll
            for (int j = 0; j < 4; ++j) w[(g + 1) & 1][j] = crc_ldg16_stream(v + ((g + 1) * 4 + j) * 256);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const uint4 x = w[g & 1][j];
            a0 = NZ_ZS(a0) ^ x.x;
            a1 = NZ_ZS(a1) ^ x.y;
            a2 = NZ_ZS(a2) ^ x.z;
            a3 = NZ_ZS(a3) ^ x.w;
        }
    }
#define NZ_Z4(a) (sT[3 * 256 + ((a) & 0xffu)] ^ sT[2 * 256 + (((a) >> 8) & 0xffu)] ^ sT[256 + (((a) >> 16) & 0xffu)] ^ sT[(a) >> 24])
    uint32_t vv = NZ_Z4(a0) ^ a1;
    vv = NZ_Z4(vv) ^ a2;
    vv = NZ_Z4(vv) ^ a3;
#undef NZ_Z4
    if constexpr (!TREE) {
        // A/B partner: every thread multiplies (32 steps of shift-and-add, no memory traffic), then XOR-reduce
        uint32_t contrib = gf_mul<POLY>(__ldg(gK + tid), vv);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) contrib ^= __shfl_xor_sync(0xffffffffu, contrib, o);
        __syncthreads();
        if ((tid & 31u) == 0) sred[tid >> 5] = contrib;
        __syncthreads();
        uint32_t r = 0;
#pragma unroll
        for (int i = 0; i < 8; ++i) r ^= sred[i];
        return ~r;
    }
    // lanes of a warp first (tree), then ONE multiply per warp moves its 512 B to the chunk end: K[t + 31] is lane 31's factor
    const uint32_t wsum = crc_warp_tree(vv, gK + (NZ_CRC_LVL_OFF - 1024));
    __syncthreads();                                       // sred is free again
    if ((tid & 31u) == 0) sred[tid >> 5] = wsum;
    __syncthreads();
    // the eight warp sums meet in warp 0: ONE warp runs the 32-step multiply (a predicated-off lane costs the same issue
    // slots as an active one, so "lane 0 of every warp" would still be eight multiplies' worth of instructions)
    uint32_t r = 0;
    if (tid < 32u) {
        if (tid < 8u) r = gf_mul<POLY>(__ldg(gK + 32u * tid + 31u), sred[tid]);
        r ^= __shfl_xor_sync(0xffffffffu, r, 4);
        r ^= __shfl_xor_sync(0xffffffffu, r, 2);
        r ^= __shfl_xor_sync(0xffffffffu, r, 1);
    }
    return ~r;                                             // valid in thread 0 (the only reader)
}

// loads the replicated stride table of one polynomial into shared memory (256 threads)
__device__ __forceinline__ void crc_load_fast_tables(uint32_t* sZS8, const uint32_t* __restrict__ gfast) {
    for (uint32_t i = threadIdx.x; i < 1024u * 8u; i += blockDim.x) sZS8[i] = gfast[((i >> 3) & 3u) * 256u + (i >> 5)];   // [b][k][cp] <- ZS[k][b]
}

// host: fills ZS | K for both polynomials (2 * NZ_CRC_FAST_WORDS words)
static inline void nz_crc_fill_fast_tables(uint32_t* out) {
    for (int poly = 0; poly < 2; ++poly) {
        uint32_t* T = out + poly * NZ_CRC_FAST_WORDS;
        auto mul = [&](uint32_t a, uint32_t b) { return poly ? gf_mul<1>(a, b) : gf_mul<0>(a, b); };
        auto xpow8 = [&](uint64_t nbytes) {
            uint32_t p = 0x80000000u, sq = 0x80000000u >> 8;
            while (nbytes) {
                if (nbytes & 1) p = mul(sq, p);
                sq = mul(sq, sq);
                nbytes >>= 1;
            }
            return p;
        };
        const uint32_t xs = xpow8(4096);
        for (int k = 0; k < 4; ++k)
            for (uint32_t b = 0; b < 256; ++b) T[k * 256 + b] = mul(xs, b << (8 * k));
        for (uint32_t t = 0; t < 256; ++t) T[1024 + t] = xpow8(4096 - 16 * t - 12);
        for (int lv = 0; lv < 5; ++lv) {                       // warp-tree levels: append 16 << lv zero bytes
            const uint32_t xl = xpow8(16u << lv);
            for (int k = 0; k < 4; ++k)
                for (uint32_t b = 0; b < 256; ++b) T[NZ_CRC_LVL_OFF + lv * 1024 + k * 256 + b] = mul(xl, b << (8 * k));
        }
    }
}

// ---------------------------------------------------------------------------
// Fused verify (K2 + K3): the gather kernel checksums a tile while it sits in shared memory.
// A piece = the part of a tile inside one 64 KiB chunk; 16 B aligned, length a multiple of 16.
// NT threads read it as rows of NT vectors (thread t takes vector t of every row; Z_{16 NT}
// between rows), then the few vectors of the last partial row.  Each thread moves its value to
// the end of the full-row region / of the piece with one GF(2) multiply by a table power; the CTA
// XOR-reduces and ONE thread applies the factors every thread has in common.
//   NT = 256: rows of 4 KiB, so the stride table is the fast path's Z_4096 (crc_tables + 16*256).
//   powers (Castagnoli, global memory at crc_tables + NZ_CRC_FUSE_OFF):
//     P16[k] = x^(8*16*k), k < 256 | P2K[m] = x^(8*2048*m), m < 32
// Values are raw register states (initial all-ones folded into the first word of a chunk, no final
// inversion), so pieces of one chunk combine by XOR after shifting each to the chunk end.
// ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t crc_z4g(uint32_t a, const uint32_t* __restrict__ gT) {
    return __ldg(gT + 3 * 256 + (a & 0xffu)) ^ __ldg(gT + 2 * 256 + ((a >> 8) & 0xffu)) ^ __ldg(gT + 256 + ((a >> 16) & 0xffu)) ^ __ldg(gT + (a >> 24));
}
// x^(8 n), n < 64 KiB
__device__ __forceinline__ uint32_t crc_xpow8_small(uint32_t n, const uint32_t* __restrict__ gF) {
    uint32_t r = __ldg(gF + ((n >> 4) & 127u));
    const uint32_t m = (n >> 11) & 31u;
    if (m) r = gf_mul<0>(r, __ldg(gF + 256 + m));
    if (n & 15u) r = gf_mul<0>(r, gf_xpow8<0>(n & 15u));
    return r;
}
template <int NT>
__device__ __forceinline__ void crc_piece_partial(const uint8_t* p, uint32_t P, bool ones, const uint32_t* sZS8, const uint32_t* __restrict__ gT,
                                                  const uint32_t* __restrict__ gF, const uint32_t* __restrict__ gL, uint32_t& main, uint32_t& part) {
    static_assert(NT == 256, "stride table is Z_4096");
    const uint32_t tid = threadIdx.x;
    const uint32_t nvec = P >> 4, R = nvec / NT, rem = nvec % NT;
    uint32_t a0 = 0, a1 = 0, a2 = 0, a3 = 0;
    const uint4* v = reinterpret_cast<const uint4*>(p) + tid;
    NZ_ZS_SETUP(tid)
#pragma unroll 4
    for (uint32_t r = 0; r < R; ++r) {
        uint4 w = v[r * NT];
        if (r == 0 && tid == 0 && ones) w.x ^= 0xFFFFFFFFu;
        a0 = NZ_ZS(a0) ^ w.x;
        a1 = NZ_ZS(a1) ^ w.y;
        a2 = NZ_ZS(a2) ^ w.z;
        a3 = NZ_ZS(a3) ^ w.w;
    }
    main = 0;
    if (R) {
        uint32_t vv = crc_z4g(a0, gT) ^ a1;
        vv = crc_z4g(vv, gT) ^ a2;
        vv = crc_z4g(vv, gT) ^ a3;
        vv = crc_z4g(vv, gT);
        main = crc_warp_tree(vv, gL);                          // R is uniform over the CTA: every lane is here.  Lane 0: the warp's 512 B,
                                                               // still to be moved to the end of the full-row region (factor gF[NT - 32 - tid])
        if ((tid & 31u) != 0) main = 0;
    }
    part = 0;
    if (tid < rem) {
        uint4 w = v[R * NT];
        if (R == 0 && tid == 0 && ones) w.x ^= 0xFFFFFFFFu;
        uint32_t vv = crc_z4g(w.x, gT) ^ w.y;
        vv = crc_z4g(vv, gT) ^ w.z;
        vv = crc_z4g(vv, gT) ^ w.w;
        vv = crc_z4g(vv, gT);
        part = gf_mul<0>(__ldg(gF + (rem - 1 - tid)), vv);     // to the end of the piece
    }
}

// host: P16 | P2K
static inline void nz_crc_fill_fuse_tables(uint32_t* T) {
    auto xpow8 = [](uint64_t nbytes) {
        uint32_t p = 0x80000000u, sq = 0x80000000u >> 8;
        while (nbytes) {
            if (nbytes & 1) p = gf_mul<0>(sq, p);
            sq = gf_mul<0>(sq, sq);
            nbytes >>= 1;
        }
        return p;
    };
    for (uint32_t k = 0; k < 256; ++k) T[k] = xpow8(16 * k);
    for (uint32_t m = 0; m < 32; ++m) T[256 + m] = xpow8(2048 * m);
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
