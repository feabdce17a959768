// nezha_b200/csrc/nz_inflate.cu — K4: RFC 1952 gzip / RFC 1951 DEFLATE inflate.
//
// NEW behaviour named by BASELINE.json:north_star (SPEC.md G1-G4); the reference's
// nginx.conf has no `gunzip on`.  Output must be bit-exact with
// zlib.decompress(member, 31) and the member's CRC-32 / ISIZE must check.
//
// Shape of the work: DEFLATE is bit-serial within one stream, so parallelism comes
// from many streams in flight (config C3: thousands of 1 MiB objects per wave).
//   * one WARP owns one stream; the Huffman/bit-reader state is warp-uniform (every
//     lane holds the same registers, so no shuffles or divergence in the decode loop);
//   * the 32 KiB LZ77 window is the OUTPUT ITSELF (L1/L2-resident recent bytes), which keeps
//     shared memory down to ~6 KiB of decode tables per warp and lets 32 warps =
//     32 streams run per SM; match copies are done by all 32 lanes at once;
//   * table entries carry base + extra-bit counts, the bit reader keeps the next input
//     word in a register one refill ahead: ~35 instructions per symbol instead of ~100
//     in the first version (ncu: it was issue bound, 43 G warp instructions per 2 GiB);
//   * the gzip trailer CRC-32 is checked by a second kernel (k_inflate_crc) that runs
//     the same chunked CRC code as K3 with the IEEE tables.
// This kernel is latency/ALU bound, not HBM bound: its roofline fraction is reported
// as compressed-bytes-in + raw-bytes-out over HBM peak and is expected to be small.
#include <cuda_runtime.h>
#include <stdint.h>

#include "nz_internal.h"
#include "nz_crc.cuh"

constexpr int INF_WARPS = 4;          // streams per CTA
constexpr int LL_BITS = 10;           // primary literal/length table: 1024 x u32
constexpr int DS_BITS = 8;            // primary distance table: 256 x u32

// Table entries carry everything the hot loop needs (no side tables):
//   both:           [7:0] bits to drop = code length + extra bits (0 = not in the primary table ->
//                   canonical slow path), [11:8] code length (the extra bits sit right above it)
//   literal/length: [13:12] kind 0 literal, 1 length, 2 end-of-block, 3 invalid symbol,
//                   [24:16] literal byte or length base
//   distance:       [30:16] base, [31] invalid symbol
struct __align__(16) WarpTabs {
    uint32_t ll[1 << LL_BITS];
    uint32_t ds[1 << DS_BITS];
    uint16_t ll_sym[288];             // symbols in canonical order (for codes longer than the table)
    uint16_t ds_sym[32];
    uint32_t ll_cnt[16];              // codes per length
    uint32_t ds_cnt[16];
    uint8_t  lens[320];               // code lengths: [0,288) literal/length, [288,320) distance
};

__constant__ uint16_t c_lbase[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
__constant__ uint8_t  c_lext[29]  = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
__constant__ uint16_t c_dbase[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577};
__constant__ uint8_t  c_dext[30]  = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};
__constant__ uint8_t  c_clorder[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};

__device__ __forceinline__ uint32_t ll_entry(int sym, int len) {
    const uint32_t l = (uint32_t)len;
    if (sym < 256) return ((uint32_t)sym << 16) | (l << 8) | l;
    if (sym == 256) return (2u << 12) | (l << 8) | l;
    if (sym >= 286) return (3u << 12) | (l << 8) | l;
    return ((uint32_t)c_lbase[sym - 257] << 16) | (1u << 12) | (l << 8) | (l + c_lext[sym - 257]);
}
__device__ __forceinline__ uint32_t ds_entry(int sym, int len) {
    const uint32_t l = (uint32_t)len;
    if (sym >= 30) return 0x80000000u | (l << 8) | l;
    return ((uint32_t)c_dbase[sym] << 16) | (l << 8) | (l + c_dext[sym]);
}
__device__ __forceinline__ uint32_t bfe32(uint32_t v, uint32_t pos, uint32_t len) {
    uint32_t r;
    asm("bfe.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(v), "r"(pos), "r"(len));
    return r;
}

// ---------------------------------------------------------------------------
// warp-uniform bit reader over global memory (LSB-first, RFC 1951 3.1.1).
// 64-bit buffer refilled 32 aligned bits at a time; the next word is always
// already in a register (loaded one refill ahead, so its latency is hidden).
// Reads may run a few words past the member's end (inside the extent's slack);
// the position checks at block boundaries and at the trailer catch that.
// ---------------------------------------------------------------------------
struct BitReader {
    const uint8_t* base;   // member start
    const uint32_t* wp;    // next word to load
    const uint32_t* wlim;  // loads beyond this mean a runaway stream
    uint64_t buf;
    uint32_t cnt;          // valid bits in buf
    uint32_t nextw;        // word at wp[-1], not yet in buf
};
__device__ __forceinline__ void br_init(BitReader& b, const uint8_t* base, const uint8_t* at, const uint8_t* end) {
    b.base = base;
    b.buf = 0;
    b.cnt = 0;
    while ((((uintptr_t)at) & 3u) != 0) {          // up to 3 bytes to reach word alignment
        b.buf |= (uint64_t)__ldg(at) << b.cnt;
        b.cnt += 8;
        ++at;
    }
    b.wp = reinterpret_cast<const uint32_t*>(at);
    b.wlim = reinterpret_cast<const uint32_t*>(end + 64);
    b.nextw = __ldg(b.wp);
    ++b.wp;
}
__device__ __forceinline__ bool br_refill(BitReader& b) {   // afterwards cnt >= 33
    if (b.cnt <= 32) {
        b.buf |= (uint64_t)b.nextw << b.cnt;
        b.cnt += 32;
        if (b.wp > b.wlim) return false;
        b.nextw = __ldg(b.wp);
        ++b.wp;
    }
    return true;
}
__device__ __forceinline__ uint32_t br_peek(const BitReader& b, uint32_t n) { return (uint32_t)b.buf & ((1u << n) - 1u); }
__device__ __forceinline__ void br_drop(BitReader& b, uint32_t n) { b.buf >>= n; b.cnt -= n; }
__device__ __forceinline__ uint32_t br_get(BitReader& b, uint32_t n) {   // n <= 16, caller refilled
    const uint32_t v = br_peek(b, n);
    br_drop(b, n);
    return v;
}
// bits consumed since the member start
__device__ __forceinline__ uint64_t br_bitpos(const BitReader& b) {
    return 8ull * (uint64_t)(reinterpret_cast<const uint8_t*>(b.wp - 1) - b.base) - b.cnt;
}

// canonical decode one bit at a time (code-length code, and codes longer than the primary table)
__device__ __forceinline__ int slow_decode(BitReader& b, const uint32_t* cnt, const uint16_t* sym, int* len_out) {
    int code = 0, first = 0, index = 0;
    for (int len = 1; len <= 15; ++len) {
        code |= (int)((b.buf >> (len - 1)) & 1u);
        const int count = (int)cnt[len];
        if (code - count < first) {
            *len_out = len;
            return sym[index + (code - first)];
        }
        index += count;
        first += count;
        first <<= 1;
        code <<= 1;
    }
    return -1;
}

// Build canonical arrays (+ optional primary table) for `n` symbols whose lengths are
// in lens[].  Warp-cooperative; returns 0 ok / -1 invalid set, following zlib's
// inflate_table acceptance rule: over-subscribed sets are rejected, incomplete sets
// only allowed when the longest code is 1 bit (or the set is empty).
// kind: 0 = code-length code (no table), 1 = literal/length, 2 = distance
__device__ int build_tables(const uint8_t* lens, int n, uint32_t* cnt, uint16_t* sym, uint32_t* tab, int tab_bits, int kind) {
    const int lane = threadIdx.x & 31;
    __syncwarp();
    if (lane < 16) cnt[lane] = 0;
    if (tab) for (int i = lane; i < (1 << tab_bits); i += 32) tab[i] = 0;
    __syncwarp();
    for (int s = lane; s < n; s += 32) atomicAdd(&cnt[lens[s]], 1u);
    __syncwarp();
    int left = 1, maxlen = 0;
    for (int l = 1; l <= 15; ++l) {
        const int c = (int)cnt[l];
        if (c) maxlen = l;
        left <<= 1;
        left -= c;
        if (left < 0) return -1;                         // over-subscribed
    }
    if (left > 0 && (kind == 0 || maxlen != 1)) {
        if (!(maxlen == 0 && kind != 0)) return -1;      // incomplete (an empty set is accepted: no codes used)
    }
    // lane l owns code length l: its first canonical index and first code
    int my_index = 0, my_code = 0;
    {
        int idx = 0, code = 0;
        for (int l = 1; l <= 15; ++l) {
            code <<= 1;
            if (l == lane) { my_index = idx; my_code = code; }
            idx += (int)cnt[l];
            code += (int)cnt[l];
        }
    }
    // every lane walks the symbols in order and handles the ones of its own length
    for (int s = 0; s < n; ++s) {
        const int l = lens[s];                           // broadcast read
        if (l == lane && l != 0) {
            sym[my_index++] = (uint16_t)s;
            if (tab && l <= tab_bits) {
                const uint32_t rev = __brev((uint32_t)my_code) >> (32 - l);
                const uint32_t e = kind == 1 ? ll_entry(s, l) : ds_entry(s, l);
                for (uint32_t i = rev; i < (1u << tab_bits); i += (1u << l)) tab[i] = e;
            }
            ++my_code;
        }
    }
    __syncwarp();
    return 0;
}

// ---------------------------------------------------------------------------
// one gzip member -> raw bytes.  Returns produced length, or -1.
// crc_expect receives the trailer CRC-32.  Members inflating to >= 4 GiB are
// declined (32-bit output cursor; ISIZE is mod 2^32 anyway).
// ---------------------------------------------------------------------------
__device__ int64_t inflate_member(const uint8_t* __restrict__ src, uint64_t n, uint8_t* __restrict__ out, uint64_t cap64,
                                  WarpTabs& T, uint32_t* crc_expect) {
    const int lane = threadIdx.x & 31;
    if (n < 18 || cap64 >= 0xFFFFFF00ull) return -1;
    const uint32_t cap = (uint32_t)cap64;
    // ---- RFC 1952 header (warp-uniform scalar reads) ----
    if (src[0] != 0x1f || src[1] != 0x8b || src[2] != 8) return -1;
    const uint32_t flg = src[3];
    if (flg & 0xE0u) return -1;
    uint64_t pos = 10;
    if (flg & 4u) {
        if (pos + 2 > n) return -1;
        const uint32_t xlen = src[pos] | ((uint32_t)src[pos + 1] << 8);
        pos += 2 + xlen;
    }
    if (flg & 8u) {
        while (pos < n && src[pos]) ++pos;
        ++pos;
    }
    if (flg & 16u) {
        while (pos < n && src[pos]) ++pos;
        ++pos;
    }
    if (flg & 2u) {
        if (pos + 2 > n) return -1;
        uint32_t c = 0xFFFFFFFFu;                        // CRC-32 of the header so far, low 16 bits stored
        for (uint64_t i = 0; i < pos; ++i) {
            c ^= src[i];
            for (int k = 0; k < 8; ++k) c = (c >> 1) ^ (NZ_POLY_IEEE & (0u - (c & 1u)));
        }
        c = ~c;
        const uint32_t h = src[pos] | ((uint32_t)src[pos + 1] << 8);
        if ((c & 0xFFFFu) != h) return -1;
        pos += 2;
    }
    if (pos + 8 > n) return -1;

    BitReader b;
    br_init(b, src, src + pos, src + n);
    const uint64_t bit_end = 8ull * (n - 8);             // deflate data may not reach into the trailer
    uint32_t p = 0;                                      // bytes produced
    bool last = false;
    bool pend = false;                                   // a deferred short copy (see the symbol loop)
    uint8_t pend_val = 0;
    uint8_t* pend_dst = nullptr;
    uint32_t pend_len = 0;
    while (!last) {
        if (!br_refill(b)) return -1;
        if (br_bitpos(b) + 3 > bit_end) return -1;
        const uint32_t hdr = br_get(b, 3);
        last = hdr & 1u;
        const uint32_t type = hdr >> 1;
        if (type == 0) {
            // ---- stored block: byte-granular, so leave the bit reader and re-enter after it ----
            const uint64_t bp = (br_bitpos(b) + 7) >> 3;
            if (bp + 4 > n - 8) return -1;
            const uint8_t* q = src + bp;
            const uint32_t len = q[0] | ((uint32_t)q[1] << 8), nlen = q[2] | ((uint32_t)q[3] << 8);
            if ((len ^ 0xFFFFu) != nlen) return -1;
            if (bp + 4 + len > n - 8) return -1;
            if (len > cap - p) return -1;
            __syncwarp();
            for (uint32_t k = lane; k < len; k += 32) out[p + k] = __ldg(q + 4 + k);
            p += len;
            br_init(b, src, q + 4 + len, src + n);
            continue;
        }
        if (type == 3) return -1;
        if (type == 1) {
            // ---- fixed codes (RFC 1951 3.2.6) ----
            __syncwarp();
            for (int s = lane; s < 288; s += 32) T.lens[s] = (uint8_t)(s < 144 ? 8 : (s < 256 ? 9 : (s < 280 ? 7 : 8)));
            T.lens[288 + lane] = 5;                       // 30 used + 2 never-valid, as zlib's fixed table
            __syncwarp();
            if (build_tables(T.lens, 288, T.ll_cnt, T.ll_sym, T.ll, LL_BITS, 1) < 0) return -1;
            if (build_tables(T.lens + 288, 32, T.ds_cnt, T.ds_sym, T.ds, DS_BITS, 2) < 0) return -1;
        } else {
            // ---- dynamic codes (3.2.7) ----
            if (!br_refill(b)) return -1;
            uint32_t hlit = br_get(b, 5) + 257, hdist = br_get(b, 5) + 1, hclen = br_get(b, 4) + 4;
            if (hlit > 286 || hdist > 30) return -1;
            __syncwarp();
            if (lane < 19) T.lens[lane] = 0;
            __syncwarp();
            for (uint32_t i = 0; i < hclen; ++i) {
                if (!br_refill(b)) return -1;
                const uint32_t v = br_get(b, 3);
                if (lane == 0) T.lens[c_clorder[i]] = (uint8_t)v;
            }
            __syncwarp();
            // code-length code: canonical arrays only (reuse the distance slots)
            if (build_tables(T.lens, 19, T.ds_cnt, T.ds_sym, nullptr, 0, 0) < 0) return -1;
            __syncwarp();
            uint32_t idx = 0;
            int prev = 0;
            const uint32_t total = hlit + hdist;
            while (idx < total) {
                if (!br_refill(b)) return -1;
                int l;
                const int sym = slow_decode(b, T.ds_cnt, T.ds_sym, &l);
                if (sym < 0) return -1;
                br_drop(b, (uint32_t)l);
                if (sym < 16) {
                    if (lane == 0) T.lens[idx] = (uint8_t)sym;
                    prev = sym;
                    ++idx;
                } else {
                    uint32_t rep, val = 0;
                    if (sym == 16) {
                        if (idx == 0) return -1;
                        rep = br_get(b, 2) + 3;
                        val = (uint32_t)prev;
                    } else if (sym == 17) {
                        rep = br_get(b, 3) + 3;
                    } else {
                        rep = br_get(b, 7) + 11;
                    }
                    if (idx + rep > total) return -1;
                    if (lane == 0) for (uint32_t k = 0; k < rep; ++k) T.lens[idx + k] = (uint8_t)val;
                    idx += rep;
                    if (sym != 16) prev = 0;
                }
            }
            if (br_bitpos(b) > bit_end) return -1;
            __syncwarp();
            // split into the two alphabets: distance lengths move to [288, 288+hdist)
            // (read all first, then write: the ranges can overlap)
            uint8_t dl = 0;
            if (lane < (int)hdist) dl = T.lens[hlit + lane];
            __syncwarp();
            for (int s = (int)hlit + lane; s < 288; s += 32) T.lens[s] = 0;
            __syncwarp();
            T.lens[288 + lane] = (lane < (int)hdist) ? dl : 0;
            __syncwarp();
            if (T.lens[256] == 0) return -1;              // zlib: "invalid code -- missing end-of-block"
            if (build_tables(T.lens, 288, T.ll_cnt, T.ll_sym, T.ll, LL_BITS, 1) < 0) return -1;
            if (build_tables(T.lens + 288, 32, T.ds_cnt, T.ds_sym, T.ds, DS_BITS, 2) < 0) return -1;
        }
        // ---- symbol loop ----
        for (;;) {
            if (!br_refill(b)) return -1;
            uint32_t w = (uint32_t)b.buf;
            uint32_t e = T.ll[w & ((1u << LL_BITS) - 1u)];
            if (e == 0) {
                int l;
                const int sym = slow_decode(b, T.ll_cnt, T.ll_sym, &l);
                if (sym < 0) return -1;
                e = ll_entry(sym, l);
            }
            const uint32_t tot = e & 0xffu;
            b.buf >>= tot;
            b.cnt -= tot;
            const uint32_t kind = (e >> 12) & 3u;
            if (kind == 0) {                              // literal
                if (p >= cap) return -1;
                if (lane == 0) out[p] = (uint8_t)(e >> 16);
                ++p;
                continue;
            }
            if (kind != 1) {
                if (kind == 2) break;                     // end of block
                return -1;                                // symbols 286/287
            }
            const uint32_t cl = (e >> 8) & 15u;
            const uint32_t mlen = (e >> 16) + bfe32(w, cl, tot - cl);
            if (!br_refill(b)) return -1;
            w = (uint32_t)b.buf;
            e = T.ds[w & ((1u << DS_BITS) - 1u)];
            if (e == 0) {
                int l;
                const int dsym = slow_decode(b, T.ds_cnt, T.ds_sym, &l);
                if (dsym < 0) return -1;
                e = ds_entry(dsym, l);
            }
            const uint32_t dtot = e & 0xffu, dcl = (e >> 8) & 15u;
            const uint32_t dist = ((e >> 16) & 0x7FFFu) + bfe32(w, dcl, dtot - dcl);
            b.buf >>= dtot;
            b.cnt -= dtot;
            // distance symbols 30/31, "invalid distance too far back", output overrun: one cold exit
            if ((e >> 31) | (uint32_t)(dist > p) | (uint32_t)(mlen > cap - p)) return -1;
            // ---- LZ77 copy: the window is the output itself (L1/L2) ----
            // A window read costs an L2 round trip.  For the common short, non-overlapping match the load is
            // issued now and its store is deferred until the next match (or the block end): the next symbols
            // decode while the bytes are in flight.  Literals never need the pending bytes; a later copy might,
            // so the pending store is retired before any new window read.
            if (pend) { if ((uint32_t)lane < pend_len) *pend_dst = pend_val; pend = false; }
            __syncwarp();                                 // orders lane 0's literal stores / earlier copies
            uint8_t* dl = out + (p + (uint32_t)lane);
            const uint8_t* sl = dl - dist;
            if (dist >= mlen && mlen <= 32) {
                pend_val = (uint32_t)lane < mlen ? *sl : (uint8_t)0;
                pend_dst = dl;
                pend_len = mlen;
                pend = true;
            } else if (dist >= mlen) {                    // no overlap: lanes independent
                for (uint32_t k = (uint32_t)lane; k < mlen; k += 32) dl[k - lane] = sl[k - lane];
            } else if (dist >= 32) {                      // overlap, period >= warp: 32-byte steps in order
                for (uint32_t k0 = 0; k0 < mlen; k0 += 32) {
                    if (k0 + (uint32_t)lane < mlen) dl[k0] = sl[k0];
                    __syncwarp();
                }
            } else {                                      // short period: every source byte precedes p
                const uint8_t* s0 = out + (p - dist);
                for (uint32_t k = lane; k < mlen; k += 32) out[p + k] = s0[k % dist];
            }
            p += mlen;
        }
        if (pend) { if ((uint32_t)lane < pend_len) *pend_dst = pend_val; pend = false; }
        __syncwarp();
        if (br_bitpos(b) > bit_end) return -1;
    }
    // ---- RFC 1952 trailer ----
    const uint64_t tpos = (br_bitpos(b) + 7) >> 3;
    if (tpos != n - 8) return -1;                         // SPEC G4: trailing bytes / second member -> declined
    const uint8_t* tp = src + tpos;
    const uint32_t crc = tp[0] | ((uint32_t)tp[1] << 8) | ((uint32_t)tp[2] << 16) | ((uint32_t)tp[3] << 24);
    const uint32_t isize = tp[4] | ((uint32_t)tp[5] << 8) | ((uint32_t)tp[6] << 16) | ((uint32_t)tp[7] << 24);
    if (isize != p) return -1;
    *crc_expect = crc;
    return (int64_t)p;
}

__global__ void __launch_bounds__(INF_WARPS * 32, 8)
k_inflate(const NzInflWork* __restrict__ infl, uint64_t* __restrict__ meta, uint32_t* counters, NzDevRes* res) {
    __shared__ WarpTabs tabs[INF_WARPS];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t n = counters[1];
    for (;;) {
        uint32_t i = 0;
        if (lane == 0) i = atomicAdd(&counters[2], 1u);
        i = __shfl_sync(0xffffffffu, i, 0);
        if (i >= n) break;
        const NzInflWork w = infl[i];
        uint32_t crc = 0;
        const int64_t got = inflate_member(reinterpret_cast<const uint8_t*>(w.src), w.src_len,
                                           reinterpret_cast<uint8_t*>(w.dst), w.dst_cap, tabs[warp], &crc);
        __syncwarp();
        if (lane == 0) {
            const bool ok = got >= 0 && (uint64_t)got == res[w.req].length;
            if (!ok) {
                res[w.req].status = NZ_ECORRUPT;
                res[w.req].bytes = 0;
                meta[i] = 0xFFFFFFFF00000000ull;
            } else {
                meta[i] = ((uint64_t)(uint32_t)got << 32) | crc;
            }
        }
    }
}

// gzip trailer CRC-32 (IEEE) of every inflated stream, chunked like K3
__global__ void __launch_bounds__(256)
k_inflate_crc(const NzInflWork* __restrict__ infl, const uint64_t* __restrict__ meta, const uint32_t* __restrict__ counters,
              NzDevRes* res, const uint32_t* __restrict__ tables) {
    __shared__ uint32_t sT[4 * 256];
    __shared__ uint32_t sZS8[1024 * 8];
    __shared__ uint32_t sred[8];
    const uint32_t* fast = tables + 16 * 256 + NZ_CRC_FAST_WORDS;                         // IEEE: ZS | K
    for (uint32_t i = threadIdx.x; i < 1024; i += 256) sT[i] = tables[8 * 256 + i];   // IEEE slice tables
    crc_load_fast_tables(sZS8, fast);
    __syncthreads();
    const uint32_t n = counters[1];
    for (uint32_t i = blockIdx.x; i < n; i += gridDim.x) {
        const uint64_t m = meta[i];
        const uint32_t len = (uint32_t)(m >> 32), expect = (uint32_t)m;
        if (len == 0xFFFFFFFFu) continue;
        const uint8_t* p = reinterpret_cast<const uint8_t*>(infl[i].dst);
        uint32_t acc = 0;
        for (uint64_t off = 0; off < len; off += NZ_CRC_CHUNK) {
            const uint32_t cl = (uint32_t)min((uint64_t)NZ_CRC_CHUNK, (uint64_t)len - off);
            const uint32_t c = (cl == NZ_CRC_CHUNK && (((uintptr_t)p) & 15u) == 0) ? block_crc_chunk64k<1>(p + off, sZS8, sT, fast + 1024, sred)
                                                                                    : block_crc_chunk<1>(p + off, cl, sT, sred);
            if (off == 0) acc = c;
            else acc = gf_mul<1>(cl == NZ_CRC_CHUNK ? c_x8pow[1][16] : gf_xpow8<1>(cl), acc) ^ c;
        }
        if (threadIdx.x == 0 && acc != expect) {
            res[infl[i].req].status = NZ_ECORRUPT;
            res[infl[i].req].bytes = 0;
        }
    }
}

cudaError_t nz_launch_inflate(const NzInflWork* infl, uint64_t* meta, uint32_t* counters, NzDevRes* res,
                              const uint32_t* crc_tables, int grid, int sm_count, cudaStream_t s) {
    k_inflate<<<grid, INF_WARPS * 32, 0, s>>>(infl, meta, counters, res);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    k_inflate_crc<<<sm_count * 4, 256, 0, s>>>(infl, meta, counters, res, crc_tables);
    return cudaGetLastError();
}

int nz_inflate_blocks_per_sm() {
    int nb = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_inflate, INF_WARPS * 32, 0) != cudaSuccess || nb < 1) nb = 8;
    return nb;
}

cudaError_t nz_inflate_upload_constants() { return nz_crc_upload_constants(); }
