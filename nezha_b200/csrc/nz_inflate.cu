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
//   * the 32 KiB LZ77 window is the OUTPUT ITSELF in HBM/L2 (ld.global.cg), which keeps
//     shared memory down to ~2.5 KiB of decode tables per warp and lets 48 warps =
//     48 streams run per SM; match copies are done by all 32 lanes at once
//     (out[p+k] = out[p-dist + k mod dist], every source byte precedes p);
//   * the gzip trailer CRC-32 is checked by a second kernel (k_inflate_crc) that runs
//     the same chunked CRC code as K3 with the IEEE tables.
// This kernel is latency/ALU bound, not HBM bound: its roofline fraction is reported
// as compressed-bytes-in + raw-bytes-out over HBM peak and is expected to be small.
#include <cuda_runtime.h>
#include <stdint.h>

#include "nz_internal.h"
#include "nz_crc.cuh"

constexpr int INF_WARPS = 4;          // streams per CTA
constexpr int LL_BITS = 9;            // primary literal/length table: 512 entries
constexpr int DS_BITS = 8;            // primary distance table: 256 entries

struct __align__(16) WarpTabs {
    uint16_t ll[1 << LL_BITS];        // (symbol << 4) | code length ; 0 = take the slow path
    uint16_t ds[1 << DS_BITS];
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

// ---------------------------------------------------------------------------
// warp-uniform bit reader over global memory (LSB-first, RFC 1951 §3.1.1)
// ---------------------------------------------------------------------------
struct BitReader {
    const uint8_t* p;      // next unread byte
    const uint8_t* end;
    uint64_t buf;
    int cnt;               // valid bits in buf
};

__device__ __forceinline__ void br_refill(BitReader& b) {
    while (b.cnt <= 32) {
        if ((((uintptr_t)b.p) & 3u) == 0 && b.p + 4 <= b.end) {
            const uint32_t w = __ldg(reinterpret_cast<const uint32_t*>(b.p));
            b.buf |= (uint64_t)w << b.cnt;
            b.cnt += 32;
            b.p += 4;
        } else if (b.p < b.end) {
            b.buf |= (uint64_t)__ldg(b.p) << b.cnt;
            b.cnt += 8;
            b.p += 1;
        } else {
            break;
        }
    }
}
__device__ __forceinline__ bool br_take(BitReader& b, int n, uint32_t& v) {   // n <= 32
    if (b.cnt < n) return false;
    v = (uint32_t)(b.buf & ((1ull << n) - 1ull));
    b.buf >>= n;
    b.cnt -= n;
    return true;
}

// canonical decode one bit at a time (used for the code-length code and for codes
// longer than the primary table).  cnt[len] = number of codes of that length,
// sym = symbols in canonical order.
__device__ __forceinline__ int slow_decode(BitReader& b, const uint32_t* cnt, const uint16_t* sym) {
    int code = 0, first = 0, index = 0;
    for (int len = 1; len <= 15; ++len) {
        if (b.cnt < len) return -1;
        code |= (int)((b.buf >> (len - 1)) & 1u);
        const int count = (int)cnt[len];
        if (code - count < first) {
            b.buf >>= len;
            b.cnt -= len;
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
// only allowed when the longest code is 1 bit.
__device__ int build_tables(const uint8_t* lens, int n, uint32_t* cnt, uint16_t* sym, uint16_t* tab, int tab_bits, bool is_codes) {
    const int lane = threadIdx.x & 31;
    __syncwarp();
    if (lane < 16) cnt[lane] = 0;
    if (tab) for (int i = lane; i < (1 << tab_bits); i += 32) tab[i] = 0;
    __syncwarp();
    for (int s = lane; s < n; s += 32) atomicAdd(&cnt[lens[s]], 1u);
    __syncwarp();
    // validity (uniform)
    int left = 1, maxlen = 0;
    for (int l = 1; l <= 15; ++l) {
        const int c = (int)cnt[l];
        if (c) maxlen = l;
        left <<= 1;
        left -= c;
        if (left < 0) return -1;                         // over-subscribed
    }
    if (left > 0 && (is_codes || maxlen != 1)) {
        if (!(maxlen == 0 && !is_codes)) return -1;      // incomplete (an empty set is accepted: no codes used)
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
                const uint16_t e = (uint16_t)((s << 4) | l);
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
// crc_expect receives the trailer CRC-32.
// ---------------------------------------------------------------------------
__device__ int64_t inflate_member(const uint8_t* __restrict__ src, uint64_t n, uint8_t* __restrict__ out, uint64_t cap,
                                  WarpTabs& T, uint32_t* crc_expect) {
    const int lane = threadIdx.x & 31;
    if (n < 18) return -1;
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
    b.p = src + pos;
    b.end = src + n;
    b.buf = 0;
    b.cnt = 0;
    uint64_t p = 0;                                      // bytes produced
    uint32_t v;
    bool last = false;
    while (!last) {
        br_refill(b);
        if (!br_take(b, 3, v)) return -1;
        last = v & 1u;
        const uint32_t type = v >> 1;
        if (type == 0) {
            // ---- stored block ----
            b.buf >>= (b.cnt & 7);
            b.cnt &= ~7;
            br_refill(b);
            uint32_t len, nlen;
            if (!br_take(b, 16, len) || !br_take(b, 16, nlen)) return -1;
            if ((len ^ 0xFFFFu) != nlen) return -1;
            // give whole buffered bytes back to the byte stream
            b.p -= (b.cnt >> 3);
            b.buf = 0;
            b.cnt = 0;
            if ((uint64_t)(b.end - b.p) < len) return -1;
            if (p + len > cap) return -1;
            __syncwarp();
            for (uint32_t k = lane; k < len; k += 32) out[p + k] = __ldg(b.p + k);
            p += len;
            b.p += len;
            continue;
        }
        if (type == 3) return -1;
        if (type == 1) {
            // ---- fixed codes (RFC 1951 §3.2.6) ----
            __syncwarp();
            for (int s = lane; s < 288; s += 32) T.lens[s] = (uint8_t)(s < 144 ? 8 : (s < 256 ? 9 : (s < 280 ? 7 : 8)));
            if (lane < 32) T.lens[288 + lane] = 5;        // 30 used + 2 never-occurring, as zlib's fixed table
            __syncwarp();
            if (build_tables(T.lens, 288, T.ll_cnt, T.ll_sym, T.ll, LL_BITS, false) < 0) return -1;
            if (build_tables(T.lens + 288, 32, T.ds_cnt, T.ds_sym, T.ds, DS_BITS, false) < 0) return -1;
        } else {
            // ---- dynamic codes (§3.2.7) ----
            br_refill(b);
            uint32_t hlit, hdist, hclen;
            if (!br_take(b, 5, hlit) || !br_take(b, 5, hdist) || !br_take(b, 4, hclen)) return -1;
            hlit += 257; hdist += 1; hclen += 4;
            if (hlit > 286 || hdist > 30) return -1;
            __syncwarp();
            if (lane < 19) T.lens[lane] = 0;
            __syncwarp();
            for (uint32_t i = 0; i < hclen; ++i) {
                br_refill(b);
                if (!br_take(b, 3, v)) return -1;
                if (lane == 0) T.lens[c_clorder[i]] = (uint8_t)v;
            }
            __syncwarp();
            // code-length code: canonical arrays only (reuse the distance slots)
            if (build_tables(T.lens, 19, T.ds_cnt, T.ds_sym, nullptr, 0, true) < 0) return -1;
            // the 19 lengths were read; now decode hlit + hdist code lengths into lens[]
            __syncwarp();
            uint32_t idx = 0;
            int prev = 0;
            const uint32_t total = hlit + hdist;
            while (idx < total) {
                br_refill(b);
                const int sym = slow_decode(b, T.ds_cnt, T.ds_sym);
                if (sym < 0) return -1;
                if (sym < 16) {
                    if (lane == 0) T.lens[idx] = (uint8_t)sym;
                    prev = sym;
                    ++idx;
                } else {
                    uint32_t rep, val = 0;
                    if (sym == 16) {
                        if (idx == 0) return -1;
                        if (!br_take(b, 2, rep)) return -1;
                        rep += 3;
                        val = (uint32_t)prev;
                    } else if (sym == 17) {
                        if (!br_take(b, 3, rep)) return -1;
                        rep += 3;
                    } else {
                        if (!br_take(b, 7, rep)) return -1;
                        rep += 11;
                    }
                    if (idx + rep > total) return -1;
                    if (lane == 0) for (uint32_t k = 0; k < rep; ++k) T.lens[idx + k] = (uint8_t)val;
                    idx += rep;
                    if (sym != 16) prev = 0;
                }
            }
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
            if (build_tables(T.lens, 288, T.ll_cnt, T.ll_sym, T.ll, LL_BITS, false) < 0) return -1;
            if (build_tables(T.lens + 288, 32, T.ds_cnt, T.ds_sym, T.ds, DS_BITS, false) < 0) return -1;
        }
        // ---- symbol loop ----
        for (;;) {
            br_refill(b);
            uint32_t e = T.ll[b.buf & ((1u << LL_BITS) - 1u)];
            int sym;
            if (e) {
                const int l = (int)(e & 15u);
                if (b.cnt < l) return -1;
                b.buf >>= l;
                b.cnt -= l;
                sym = (int)(e >> 4);
            } else {
                sym = slow_decode(b, T.ll_cnt, T.ll_sym);
                if (sym < 0) return -1;
            }
            if (sym < 256) {
                if (p >= cap) return -1;
                if (lane == 0) out[p] = (uint8_t)sym;
                ++p;
                continue;
            }
            if (sym == 256) break;
            sym -= 257;
            if (sym >= 29) return -1;
            uint32_t ext;
            if (!br_take(b, c_lext[sym], ext)) return -1;
            const uint32_t mlen = c_lbase[sym] + ext;
            br_refill(b);
            e = T.ds[b.buf & ((1u << DS_BITS) - 1u)];
            int dsym;
            if (e) {
                const int l = (int)(e & 15u);
                if (b.cnt < l) return -1;
                b.buf >>= l;
                b.cnt -= l;
                dsym = (int)(e >> 4);
            } else {
                dsym = slow_decode(b, T.ds_cnt, T.ds_sym);
                if (dsym < 0) return -1;
            }
            if (dsym >= 30) return -1;
            if (!br_take(b, c_dext[dsym], ext)) return -1;
            const uint32_t dist = c_dbase[dsym] + ext;
            if (dist > p) return -1;                      // "invalid distance too far back"
            if (p + mlen > cap) return -1;
            // ---- LZ77 copy: all source bytes precede p, so the lanes are independent ----
            __syncwarp();                                 // orders lane 0's literal stores / earlier copies
            const uint8_t* s0 = out + (p - dist);
            uint8_t* d0 = out + p;
            if (dist >= mlen) {
                for (uint32_t k = lane; k < mlen; k += 32) d0[k] = __ldcg(s0 + k);
            } else {
                for (uint32_t k = lane; k < mlen; k += 32) d0[k] = __ldcg(s0 + (k % dist));
            }
            p += mlen;
        }
    }
    // ---- RFC 1952 trailer ----
    b.buf >>= (b.cnt & 7);
    b.cnt &= ~7;
    const uint8_t* tp = b.p - (b.cnt >> 3);
    if ((uint64_t)(b.end - tp) != 8) return -1;           // SPEC G4: trailing bytes / second member -> declined
    const uint32_t crc = tp[0] | ((uint32_t)tp[1] << 8) | ((uint32_t)tp[2] << 16) | ((uint32_t)tp[3] << 24);
    const uint32_t isize = tp[4] | ((uint32_t)tp[5] << 8) | ((uint32_t)tp[6] << 16) | ((uint32_t)tp[7] << 24);
    if (isize != (uint32_t)p) return -1;
    *crc_expect = crc;
    return (int64_t)p;
}

__global__ void __launch_bounds__(INF_WARPS * 32, 12)
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
    __shared__ uint32_t sred[8];
    for (uint32_t i = threadIdx.x; i < 1024; i += 256) sT[i] = tables[8 * 256 + i];   // IEEE slice tables
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
            const uint32_t c = block_crc_chunk<1>(p + off, cl, sT, sred);
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
