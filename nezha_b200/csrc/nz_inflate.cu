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
// That whole-member kernel (k_inflate) is what runs for a member WITHOUT a checkpoint index.
// Members are indexed once at nz_put_commit (k_inflate_index, nz_internal.h "Checkpoint index");
// a GET then runs k_inflate_seg:
//   * one warp per SEGMENT (~128 KiB of output), segments of one member decode concurrently;
//   * inside a segment the warp walks GROUPS of <= 32 sub-segments of <= 32 tokens each:
//     phase 1, every LANE Huffman-decodes its own sub-segment from its recorded bit offset (tables
//     shared by the warp), stores its literals straight to their final place and queues its matches
//     in shared memory; phase 2, the warp resolves the queued LZ77 copies in stream order, eight
//     matches per step (four lanes each) with readiness rounds for the few that depend on an earlier
//     match of the same step;
//   * bytes a segment's matches copy from before the segment's start come from the index ("side bytes").
// All inflate kernels are latency/ALU bound, not HBM bound: the roofline fraction is reported
// as compressed-bytes-in + raw-bytes-out over HBM peak and is expected to be small.
#include <cuda_runtime.h>
#include <stdint.h>

#include <stdlib.h>

#include <algorithm>

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

// position the warp-uniform reader at an arbitrary bit offset of the member
__device__ __forceinline__ bool br_init_at(BitReader& b, const uint8_t* base, uint64_t bit, const uint8_t* end) {
    br_init(b, base, base + (bit >> 3), end);
    if (!br_refill(b)) return false;
    br_drop(b, (uint32_t)(bit & 7u));
    return true;
}

// RFC 1952 member header (warp-uniform scalar reads): offset of the first deflate byte, or -1
__device__ int64_t gzip_header(const uint8_t* __restrict__ src, uint64_t n) {
    if (n < 18) return -1;
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
    return (int64_t)pos;
}

// After a block's 3-bit header: build its decode tables (type 1 fixed codes, RFC 1951 3.2.6; type 2 dynamic
// codes, 3.2.7).  Warp-cooperative, reader warp-uniform.  0 ok / -1 malformed.
__device__ int block_tables(BitReader& b, uint32_t type, WarpTabs& T, uint64_t bit_end) {
    const int lane = threadIdx.x & 31;
    if (type == 1) {
        __syncwarp();
        for (int s = lane; s < 288; s += 32) T.lens[s] = (uint8_t)(s < 144 ? 8 : (s < 256 ? 9 : (s < 280 ? 7 : 8)));
        T.lens[288 + lane] = 5;                       // 30 used + 2 never-valid, as zlib's fixed table
        __syncwarp();
        if (build_tables(T.lens, 288, T.ll_cnt, T.ll_sym, T.ll, LL_BITS, 1) < 0) return -1;
        if (build_tables(T.lens + 288, 32, T.ds_cnt, T.ds_sym, T.ds, DS_BITS, 2) < 0) return -1;
        return 0;
    }
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
    return 0;
}

// LZ77 copy by the whole warp: the window is the output itself (L1/L2).  p = bytes produced so far.
__device__ __forceinline__ void warp_copy_match(uint8_t* __restrict__ out, uint32_t p, uint32_t mlen, uint32_t dist, int lane) {
    uint8_t* dl = out + (p + (uint32_t)lane);
    const uint8_t* sl = dl - dist;
    if (dist >= mlen) {                           // no overlap: lanes independent
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
}

// ---------------------------------------------------------------------------
// Index builder state (k_inflate_index).  Everything is warp-uniform; lane 0 writes the records.
// ---------------------------------------------------------------------------
struct IdxBuild {
    NzIdxSeg* seg; NzIdxGrp* grp; uint32_t* delta; uint8_t* side;
    uint64_t side_cap, side_cur;
    uint32_t seg_cap, grp_cap, seg_bytes;
    uint32_t n_seg, n_grp;
    bool on;                       // false: not indexing (plain GET decode, or the caps were exceeded)
    uint32_t tok_in_sub, sub_in_grp;
    uint64_t sub_bit0, grp_bit0, blk_hdr_bit;
    uint32_t sub_out0, grp_out0;
    uint32_t seg_out0, seg_first_grp;
    uint64_t seg_side0;
};
__device__ __forceinline__ void ib_close_seg(IdxBuild& x, uint32_t p, int lane) {
    if (!x.on) return;
    if (p == x.seg_out0 && x.n_grp == x.seg_first_grp) return;      // nothing in it
    if (x.n_seg >= x.seg_cap) { x.on = false; return; }
    if (lane == 0) {
        NzIdxSeg s;
        s.out_start = x.seg_out0; s.out_len = p - x.seg_out0; s.side_off = x.seg_side0;
        s.first_grp = x.seg_first_grp; s.n_grp = x.n_grp - x.seg_first_grp;
        x.seg[x.n_seg] = s;
    }
    ++x.n_seg;
    x.seg_out0 = p; x.seg_first_grp = x.n_grp; x.seg_side0 = x.side_cur;
}
__device__ __forceinline__ void ib_close_grp(IdxBuild& x, uint32_t p, int lane) {
    if (!x.on || x.sub_in_grp == 0) return;
    if (x.n_grp >= x.grp_cap) { x.on = false; return; }
    if (lane == 0) {
        NzIdxGrp g;
        g.tok_bit = x.grp_bit0; g.out_off = x.grp_out0; g.hdr_bit = x.blk_hdr_bit; g.n_sub = x.sub_in_grp; g.flags = 0;
        x.grp[x.n_grp] = g;
    }
    if (lane >= (int)x.sub_in_grp) x.delta[(size_t)x.n_grp * NZ_IDX_SUBS + lane] = 0;
    ++x.n_grp;
    x.sub_in_grp = 0;
    if (p - x.seg_out0 >= x.seg_bytes) ib_close_seg(x, p, lane);
}
__device__ __forceinline__ void ib_close_sub(IdxBuild& x, uint64_t bit_now, uint32_t p, int lane) {
    if (!x.on || x.tok_in_sub == 0) return;
    if (x.n_grp >= x.grp_cap) { x.on = false; return; }
    if (lane == 0) x.delta[(size_t)x.n_grp * NZ_IDX_SUBS + x.sub_in_grp] = (uint32_t)(bit_now - x.sub_bit0) | ((p - x.sub_out0) << 16);
    x.tok_in_sub = 0;
    if (++x.sub_in_grp == NZ_IDX_SUBS) ib_close_grp(x, p, lane);
}
__device__ __forceinline__ void ib_stored(IdxBuild& x, uint64_t data_byte, uint32_t len, uint32_t p_before, int lane) {
    if (!x.on || len == 0) return;
    if (x.n_grp >= x.grp_cap) { x.on = false; return; }
    if (lane == 0) {
        NzIdxGrp g;
        g.tok_bit = data_byte * 8; g.out_off = p_before; g.hdr_bit = len; g.n_sub = 0; g.flags = NZ_IDX_GRP_STORED;
        x.grp[x.n_grp] = g;
    }
    x.delta[(size_t)x.n_grp * NZ_IDX_SUBS + lane] = 0;
    ++x.n_grp;
    if (p_before + len - x.seg_out0 >= x.seg_bytes) ib_close_seg(x, p_before + len, lane);
}

// ---------------------------------------------------------------------------
// one gzip member -> raw bytes, whole member by one warp.  Returns produced length, or -1.
// crc_expect receives the trailer CRC-32.  Members inflating to >= 4 GiB are
// declined (32-bit output cursor; ISIZE is mod 2^32 anyway).
// INDEX: also record the checkpoint index (k_inflate_index).
// ---------------------------------------------------------------------------
template <bool INDEX>
__device__ int64_t inflate_member(const uint8_t* __restrict__ src, uint64_t n, uint8_t* __restrict__ out, uint64_t cap64,
                                  WarpTabs& T, uint32_t* crc_expect, IdxBuild* ib) {
    const int lane = threadIdx.x & 31;
    if (n < 18 || cap64 >= 0xFFFFFF00ull) return -1;
    const uint32_t cap = (uint32_t)cap64;
    const int64_t hpos = gzip_header(src, n);
    if (hpos < 0) return -1;
    const uint64_t pos = (uint64_t)hpos;

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
        if constexpr (INDEX) ib->blk_hdr_bit = br_bitpos(b);
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
            if constexpr (INDEX) ib_stored(*ib, bp + 4, len, p, lane);
            p += len;
            br_init(b, src, q + 4 + len, src + n);
            continue;
        }
        if (type == 3) return -1;
        if (block_tables(b, type, T, bit_end) < 0) return -1;
        // ---- symbol loop ----
        for (;;) {
            if (!br_refill(b)) return -1;
            if constexpr (INDEX) {
                if (ib->on && ib->tok_in_sub == 0) {
                    ib->sub_bit0 = br_bitpos(b); ib->sub_out0 = p;
                    if (ib->sub_in_grp == 0) { ib->grp_bit0 = ib->sub_bit0; ib->grp_out0 = p; }
                }
            }
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
                if constexpr (INDEX) { if (ib->on && ++ib->tok_in_sub == NZ_IDX_T) ib_close_sub(*ib, br_bitpos(b), p, lane); }
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
            if constexpr (INDEX) {
                __syncwarp();                             // orders lane 0's literal stores / earlier copies
                // bytes this match takes from before the current segment's start go into the index, in stream order
                if (ib->on && p - dist < ib->seg_out0) {
                    const uint32_t nside = min(mlen, ib->seg_out0 - (p - dist));
                    if (ib->side_cur + nside > ib->side_cap) {
                        ib->on = false;
                    } else {
                        for (uint32_t k = lane; k < nside; k += 32) ib->side[ib->side_cur + k] = out[p - dist + k];
                        ib->side_cur += nside;
                    }
                }
                warp_copy_match(out, p, mlen, dist, lane);
                p += mlen;
                if (ib->on && ++ib->tok_in_sub == NZ_IDX_T) ib_close_sub(*ib, br_bitpos(b), p, lane);
                continue;
            }
            // ---- LZ77 copy: the window is the output itself (L1/L2) ----
            // A window read costs an L2 round trip.  For the common short, non-overlapping match the load is
            // issued now and its store is deferred until the next match (or the block end): the next symbols
            // decode while the bytes are in flight.  Literals never need the pending bytes; a later copy might,
            // so the pending store is retired before any new window read.
            if (pend) { if ((uint32_t)lane < pend_len) *pend_dst = pend_val; pend = false; }
            __syncwarp();                                 // orders lane 0's literal stores / earlier copies
            if (dist >= mlen && mlen <= 32) {
                uint8_t* dl = out + (p + (uint32_t)lane);
                pend_val = (uint32_t)lane < mlen ? *(dl - dist) : (uint8_t)0;
                pend_dst = dl;
                pend_len = mlen;
                pend = true;
            } else {
                warp_copy_match(out, p, mlen, dist, lane);
            }
            p += mlen;
        }
        if (pend) { if ((uint32_t)lane < pend_len) *pend_dst = pend_val; pend = false; }
        __syncwarp();
        if (br_bitpos(b) > bit_end) return -1;
        if constexpr (INDEX) {                            // the block's last (short) sub-segment includes the end-of-block code
            ib_close_sub(*ib, br_bitpos(b), p, lane);
            ib_close_grp(*ib, p, lane);
        }
    }
    // ---- RFC 1952 trailer ----
    const uint64_t tpos = (br_bitpos(b) + 7) >> 3;
    if (tpos != n - 8) return -1;                         // SPEC G4: trailing bytes / second member -> declined
    const uint8_t* tp = src + tpos;
    const uint32_t crc = tp[0] | ((uint32_t)tp[1] << 8) | ((uint32_t)tp[2] << 16) | ((uint32_t)tp[3] << 24);
    const uint32_t isize = tp[4] | ((uint32_t)tp[5] << 8) | ((uint32_t)tp[6] << 16) | ((uint32_t)tp[7] << 24);
    if (isize != p) return -1;
    if constexpr (INDEX) ib_close_seg(*ib, p, lane);
    *crc_expect = crc;
    return (int64_t)p;
}

// every request of the wave that is served from this member's inflated bytes fails with it
__device__ void infl_fail(const NzInflWork& w, NzDevRes* res, const NzDevReq* __restrict__ reqs, uint32_t n_req) {
    res[w.req].status = NZ_ECORRUPT;
    res[w.req].bytes = 0;
    for (uint32_t r = 0; r < n_req; ++r)
        if (reqs[r].infl_leader == w.req && r != w.req && (reqs[r].flags & NZ_GET_INFLATE) && res[r].status == NZ_OK) {
            res[r].status = NZ_ECORRUPT;
            res[r].bytes = 0;
        }
}

__global__ void __launch_bounds__(INF_WARPS * 32, 8)
k_inflate(const NzInflWork* __restrict__ infl, uint64_t* __restrict__ meta, uint32_t* counters, NzDevRes* res,
          const NzDevReq* __restrict__ reqs, uint32_t n_req) {
    __shared__ WarpTabs tabs[INF_WARPS];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (counters[5] == 0) return;                         // every member of the wave has an index
    const uint32_t n = counters[1];
    for (;;) {
        uint32_t i = 0;
        if (lane == 0) i = atomicAdd(&counters[2], 1u);
        i = __shfl_sync(0xffffffffu, i, 0);
        if (i >= n) break;
        const NzInflWork w = infl[i];
        if (w.idx) continue;                              // indexed member: k_inflate_seg's work
        uint32_t crc = 0;
        const int64_t got = inflate_member<false>(reinterpret_cast<const uint8_t*>(w.src), w.src_len,
                                                  reinterpret_cast<uint8_t*>(w.dst), w.dst_cap, tabs[warp], &crc, nullptr);
        __syncwarp();
        if (lane == 0) {
            const bool ok = got >= 0 && (uint64_t)got == w.dst_cap;
            if (!ok) {
                infl_fail(w, res, reqs, n_req);
                meta[i] = 0xFFFFFFFF00000000ull;
            } else {
                meta[i] = ((uint64_t)(uint32_t)got << 32) | crc;
            }
        }
    }
}

// ---------------------------------------------------------------------------
// K4 with the checkpoint index: one warp per segment.
// ---------------------------------------------------------------------------
constexpr int SEG_WARPS = 2;                              // 2 x (6.2 KiB tables + 8.1 KiB match queue) of static shared memory
constexpr int LIST_STRIDE = NZ_IDX_T + 1;                 // queue of lane s: list[s * 65 + k] (odd stride: a step's 8 reads hit 8 banks)
struct SegShared {
    WarpTabs T;
    uint32_t list[NZ_IDX_SUBS * LIST_STRIDE];
};

// lane-private bit reader (phase 1: every lane walks its own sub-segment).  Two words are always in flight ahead of
// the buffer, so a refill never waits for memory (a sub-segment's compressed bytes are touched once: they mostly miss L1)
struct LaneBits {
    const uint32_t* wp;     // next word to request
    uint64_t buf;
    uint32_t cnt;
    uint32_t w1, w2;        // the next two words, already requested
};
__device__ __forceinline__ void lb_init(LaneBits& b, const uint8_t* base, uint64_t bit) {
    const uint8_t* at = base + (bit >> 3);
    const uint32_t mis = (uint32_t)((uintptr_t)at & 3u);
    const uint32_t* wa = reinterpret_cast<const uint32_t*>(at - mis);
    const uint32_t skip = mis * 8u + (uint32_t)(bit & 7u);          // 0..31
    b.buf = ((uint64_t)__ldg(wa) | ((uint64_t)__ldg(wa + 1) << 32)) >> skip;
    b.cnt = 64u - skip;
    b.w1 = __ldg(wa + 2);
    b.w2 = __ldg(wa + 3);
    b.wp = wa + 4;
}
__device__ __forceinline__ void lb_refill(LaneBits& b) {             // afterwards cnt >= 33
    if (b.cnt <= 32) {
        b.buf |= (uint64_t)b.w1 << b.cnt;
        b.cnt += 32;
        b.w1 = b.w2;
        b.w2 = __ldg(b.wp);
        ++b.wp;
    }
}
// canonical decode from a lane-private buffer (codes longer than the primary table)
__device__ __forceinline__ int lane_slow_decode(uint64_t buf, const uint32_t* cnt, const uint16_t* sym, int* len_out) {
    int code = 0, first = 0, index = 0;
    for (int len = 1; len <= 15; ++len) {
        code |= (int)((buf >> (len - 1)) & 1u);
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

// One segment.  Returns true if every recorded length was reproduced exactly.
__device__ bool inflate_segment(const NzInflWork& w, uint32_t si, SegShared& S) {
    const int lane = threadIdx.x & 31;
    const uint8_t* __restrict__ src = reinterpret_cast<const uint8_t*>(w.src);
    uint8_t* out = reinterpret_cast<uint8_t*>(w.dst);
    const uint8_t* idx = reinterpret_cast<const uint8_t*>(w.idx);
    const NzIdxHeader* H = reinterpret_cast<const NzIdxHeader*>(idx);
    if (H->magic != NZ_IDX_MAGIC || si >= H->n_seg || H->raw_len != w.dst_cap) return false;
    const NzIdxSeg sg = reinterpret_cast<const NzIdxSeg*>(idx + H->off_seg)[si];
    const uint4* grps = reinterpret_cast<const uint4*>(idx + H->off_grp);
    const uint32_t* deltas = reinterpret_cast<const uint32_t*>(idx + H->off_delta);
    const uint8_t* __restrict__ side = idx + H->off_side + sg.side_off;
    const uint32_t seg_start = (uint32_t)sg.out_start;
    const uint64_t bit_end = 8ull * (w.src_len - 8);
    uint32_t side_cur = 0;
    uint64_t tables_of = ~0ull;                            // header bit of the block whose tables are in shared memory
    bool bad = false;
    const uint32_t lm_le = 0xFFFFFFFFu >> (31 - lane);      // bits 0..lane
    const uint32_t g_end = sg.first_grp + sg.n_grp;
    // the next group's record and deltas are requested one group ahead
    uint4 nga = make_uint4(0, 0, 0, 0), ngb = nga;
    uint32_t nd = 0;
    if (sg.n_grp) { nga = __ldg(grps + 2 * (size_t)sg.first_grp); ngb = __ldg(grps + 2 * (size_t)sg.first_grp + 1); nd = __ldg(deltas + (size_t)sg.first_grp * NZ_IDX_SUBS + lane); }
    for (uint32_t g = sg.first_grp; g < g_end; ++g) {
        NzIdxGrp G;
        G.tok_bit = (uint64_t)nga.x | ((uint64_t)nga.y << 32); G.out_off = (uint64_t)nga.z | ((uint64_t)nga.w << 32);
        G.hdr_bit = (uint64_t)ngb.x | ((uint64_t)ngb.y << 32); G.n_sub = ngb.z; G.flags = ngb.w;
        const uint32_t d = nd;
        if (g + 1 < g_end) { nga = __ldg(grps + 2 * (size_t)(g + 1)); ngb = __ldg(grps + 2 * (size_t)(g + 1) + 1); nd = __ldg(deltas + (size_t)(g + 1) * NZ_IDX_SUBS + lane); }
        if (G.flags & NZ_IDX_GRP_STORED) {
            const uint8_t* q = src + (G.tok_bit >> 3);
            const uint32_t len = (uint32_t)G.hdr_bit;
            if ((G.tok_bit >> 3) + len > w.src_len || G.out_off + len > w.dst_cap) return false;
            for (uint32_t k = lane; k < len; k += 32) out[G.out_off + k] = __ldg(q + k);
            continue;
        }
        if (G.n_sub == 0 || G.n_sub > NZ_IDX_SUBS) return false;
        if (G.hdr_bit != tables_of) {
            BitReader b;
            __syncwarp();
            if (!br_init_at(b, src, G.hdr_bit, src + w.src_len)) return false;
            if (!br_refill(b)) return false;
            const uint32_t type = br_get(b, 3) >> 1;
            if (type != 1 && type != 2) return false;
            if (block_tables(b, type, S.T, bit_end) < 0) return false;
            tables_of = G.hdr_bit;
        }
        // ---- this lane's sub-segment ----
        const uint32_t my_bits = (uint32_t)lane < G.n_sub ? d & 0xffffu : 0u, my_out = (uint32_t)lane < G.n_sub ? d >> 16 : 0u;
        uint32_t bit_incl = my_bits, out_incl = my_out;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t tb = __shfl_up_sync(0xffffffffu, bit_incl, o), to = __shfl_up_sync(0xffffffffu, out_incl, o);
            if (lane >= o) { bit_incl += tb; out_incl += to; }
        }
        const uint32_t out0 = (uint32_t)G.out_off + out_incl - my_out;         // first output byte of this lane's sub-segment
        const uint32_t grp_out_end = (uint32_t)G.out_off + __shfl_sync(0xffffffffu, out_incl, 31);
        if (grp_out_end > w.dst_cap) return false;
        // ---- phase 1: Huffman decode, literals to their place, matches queued ----
        uint32_t nm = 0;
        if (my_bits) {
            LaneBits b;
            lb_init(b, src, G.tok_bit + (bit_incl - my_bits));
            uint32_t rem = my_bits, nlit = 0, ntok = 0;
            uint8_t* o = out + out0;
            uint32_t* myq = S.list + lane * LIST_STRIDE;
            while (rem) {
                lb_refill(b);
                uint32_t wv = (uint32_t)b.buf;
                uint32_t e = S.T.ll[wv & ((1u << LL_BITS) - 1u)];
                if (e == 0) {
                    int l;
                    const int sym = lane_slow_decode(b.buf, S.T.ll_cnt, S.T.ll_sym, &l);
                    if (sym < 0) { bad = true; break; }
                    e = ll_entry(sym, l);
                }
                const uint32_t tot = e & 0xffu;
                if (tot > rem || ++ntok > NZ_IDX_T + 1) { bad = true; break; }
                b.buf >>= tot; b.cnt -= tot; rem -= tot;
                const uint32_t kind = (e >> 12) & 3u;
                if (kind == 0) {
                    *o++ = (uint8_t)(e >> 16);
                    ++nlit;
                    continue;
                }
                if (kind != 1) {                          // end of block closes the block's last sub-segment
                    if (kind != 2 || rem != 0) bad = true;
                    break;
                }
                const uint32_t cl = (e >> 8) & 15u;
                const uint32_t mlen = (e >> 16) + bfe32(wv, cl, tot - cl);
                lb_refill(b);
                wv = (uint32_t)b.buf;
                e = S.T.ds[wv & ((1u << DS_BITS) - 1u)];
                if (e == 0) {
                    int l;
                    const int dsym = lane_slow_decode(b.buf, S.T.ds_cnt, S.T.ds_sym, &l);
                    if (dsym < 0) { bad = true; break; }
                    e = ds_entry(dsym, l);
                }
                const uint32_t dtot = e & 0xffu, dcl = (e >> 8) & 15u;
                const uint32_t dist = ((e >> 16) & 0x7FFFu) + bfe32(wv, dcl, dtot - dcl);
                if (dtot > rem || (e >> 31) || nm >= NZ_IDX_T || nlit >= NZ_IDX_T) { bad = true; break; }
                b.buf >>= dtot; b.cnt -= dtot; rem -= dtot;
                myq[nm++] = nlit | ((mlen - 3u) << 6) | ((dist - 1u) << 14);
                o += mlen;
                nlit = 0;
            }
            if ((uint32_t)(o - (out + out0)) != my_out) bad = true;
        }
        if (__any_sync(0xffffffffu, bad)) return false;
        __syncwarp();
        // ---- phase 2: LZ77 copies, OUTPUT-centric.  One sub-segment's queue (<= 32 matches, one per lane) at a time; its
        // output range is walked 128 bytes per iteration, lane b taking bytes p+b, p+32+b, p+64+b, p+96+b: each finds the
        // match that covers its byte (start bits of the window + population count, then three shuffles), every byte whose
        // source lies below p - or on a byte of this iteration that is already final - is loaded, then all are stored
        // (coalesced), and the few that depend on a byte of the same iteration take another round.
        for (uint32_t s = 0; s < G.n_sub; ++s) {
            const uint32_t c = __shfl_sync(0xffffffffu, nm, s);
            if (c == 0) continue;
            const uint32_t base = __shfl_sync(0xffffffffu, out0, s);
            const bool have = (uint32_t)lane < c;
            const uint32_t ent = have ? S.list[s * LIST_STRIDE + lane] : 0u;
            const uint32_t ml = have ? ((ent >> 6) & 255u) + 3u : 0u;
            const uint32_t ds = (ent >> 14) + 1u;
            uint32_t incl = (ent & 63u) + ml;              // literals before the match + the match
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += t;
            }
            const uint32_t mpos = have ? base + incl - ml : 0xFFFFFFFFu;      // absent lanes never start inside a window
            const uint32_t mend = have ? mpos + ml : 0xFFFFFFFFu;
            if (__any_sync(0xffffffffu, have && ds > mpos)) return false;    // would reach before the member's first byte
            // bytes from before the segment come from the index's side bytes, consumed in stream order
            uint32_t ns = 0, sb = 0;
            bool list_side = false;
            if (seg_start != 0) {
                ns = (have && mpos - ds < seg_start) ? min(ml, seg_start - (mpos - ds)) : 0u;
                list_side = __any_sync(0xffffffffu, ns != 0);
                if (list_side) {
                    uint32_t sincl = ns;
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1) {
                        const uint32_t t = __shfl_up_sync(0xffffffffu, sincl, o);
                        if (lane >= o) sincl += t;
                    }
                    sb = side_cur + sincl - ns;
                    side_cur += __shfl_sync(0xffffffffu, sincl, 31);
                }
            }
            const uint32_t p_first = __shfl_sync(0xffffffffu, mpos, 0), p_last = __shfl_sync(0xffffffffu, mend, (int)c - 1);
            uint32_t nbefore = 0;                           // matches that start below p
            for (uint32_t p = p_first; p < p_last; p += 128u) {
                const uint32_t rel = mpos - p;              // < 128: this match starts in the iteration's range
                const uint32_t bit = 1u << (rel & 31u);
                uint32_t sm[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) sm[u] = __reduce_or_sync(0xffffffffu, (rel >> 5) == (uint32_t)u ? bit : 0u);
                uint32_t sp[4], so[4], vv[4], mpv[4], mdv[4], mnsv[4], msbv[4];
                bool act[4], fs[4];
                uint32_t nb = nbefore;
                bool periodic = false;
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const uint32_t cnt = nb + __popc(sm[u] & lm_le);          // matches that start at or below this lane's byte
                    nb += __popc(sm[u]);
                    const int mi = (int)(cnt - 1u) & 31;
                    const uint32_t mp = __shfl_sync(0xffffffffu, mpos, mi), me = __shfl_sync(0xffffffffu, mend, mi);
                    const uint32_t md = __shfl_sync(0xffffffffu, ds, mi);
                    const uint32_t x = p + 32u * u + (uint32_t)lane;
                    act[u] = cnt != 0 && x < me;                              // covered by that match (else a literal, already in place)
                    sp[u] = x - md;                                           // source position (first period)
                    mpv[u] = mp; mdv[u] = md;
                    periodic |= act[u] && x - mp >= md;                       // a match longer than its distance repeats its first `dist` bytes
                    fs[u] = false;
                    so[u] = 0; mnsv[u] = 0; msbv[u] = 0;
                    if (list_side) {
                        mnsv[u] = __shfl_sync(0xffffffffu, ns, mi);
                        msbv[u] = __shfl_sync(0xffffffffu, sb, mi);
                        fs[u] = act[u] && x - mp < mnsv[u];
                        so[u] = msbv[u] + (x - mp);
                    }
                }
                if (__any_sync(0xffffffffu, periodic)) {                      // rare on text: take the byte from the first period, which lies below the match
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const uint32_t i = p + 32u * u + (uint32_t)lane - mpv[u];
                        if (act[u] && i >= mdv[u]) {
                            const uint32_t j = i % mdv[u];
                            sp[u] = mpv[u] - mdv[u] + j;
                            fs[u] = j < mnsv[u];
                            so[u] = msbv[u] + j;
                        }
                    }
                }
                nbefore = nb;
                // round 1: every byte whose source lies below p (or in the side bytes) - on text nearly all of them
                {
                    bool rdy[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        rdy[u] = act[u] && (fs[u] || (int32_t)(sp[u] - p) < 0);
                        if (rdy[u]) vv[u] = fs[u] ? (uint32_t)__ldg(side + so[u]) : (uint32_t)out[sp[u]];
                    }
#pragma unroll
                    for (int u = 0; u < 4; ++u)
                        if (rdy[u]) { out[p + 32u * u + (uint32_t)lane] = (uint8_t)vv[u]; act[u] = false; }
                }
                __syncwarp();                                                 // the stores above are visible to every lane's later loads
                if (!__any_sync(0xffffffffu, act[0] | act[1] | act[2] | act[3])) continue;
                // further rounds: a byte is ready once its source byte (inside this iteration's range) is no longer pending
                uint32_t pend[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) pend[u] = __ballot_sync(0xffffffffu, act[u]);
                while (pend[0] | pend[1] | pend[2] | pend[3]) {
                    bool rdy[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const uint32_t dd = sp[u] - p;
                        const uint32_t wsel = (dd >> 5) & 3u;
                        const uint32_t pw = wsel == 0 ? pend[0] : wsel == 1 ? pend[1] : wsel == 2 ? pend[2] : pend[3];
                        rdy[u] = act[u] && !((pw >> (dd & 31u)) & 1u);
                        if (rdy[u]) vv[u] = (uint32_t)out[sp[u]];
                    }
#pragma unroll
                    for (int u = 0; u < 4; ++u)
                        if (rdy[u]) { out[p + 32u * u + (uint32_t)lane] = (uint8_t)vv[u]; act[u] = false; }
                    __syncwarp();
#pragma unroll
                    for (int u = 0; u < 4; ++u) pend[u] = __ballot_sync(0xffffffffu, act[u]);
                }
            }
        }
        __syncwarp();
    }
    return true;
}

// seg_begin[0..n_infl]: exclusive prefix of the members' segment counts (0 for members without an index), by k_plan
template <int MINB>
__global__ void __launch_bounds__(SEG_WARPS * 32, MINB)
k_inflate_seg(const NzInflWork* __restrict__ infl, const uint32_t* __restrict__ seg_begin, uint32_t* counters, NzDevRes* res,
              const NzDevReq* __restrict__ reqs, uint32_t n_req) {
    __shared__ SegShared sh[SEG_WARPS];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t n = counters[1];
    const uint32_t total = seg_begin[n];
    for (;;) {
        uint32_t gid = 0;
        if (lane == 0) gid = atomicAdd(&counters[4], 1u);
        gid = __shfl_sync(0xffffffffu, gid, 0);
        if (gid >= total) break;
        uint32_t a = 0, b = n;
        while (b - a > 1) {                                // last member with seg_begin[m] <= gid
            const uint32_t m = (a + b) >> 1;
            if (__ldg(seg_begin + m) <= gid) a = m; else b = m;
        }
        const NzInflWork w = infl[a];
        const bool ok = inflate_segment(w, gid - __ldg(seg_begin + a), sh[warp]);
        __syncwarp();
        if (!ok && lane == 0) infl_fail(w, res, reqs, n_req);
    }
}

// ---------------------------------------------------------------------------
// Index build at commit: one warp per member, whole-member decode into a temporary buffer.
// Fills infl/meta so that k_inflate_crc can check the trailer CRC-32 of the temporary output afterwards.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(INF_WARPS * 32, 6)
k_inflate_index(const NzIdxJob* __restrict__ jobs, NzIdxResult* __restrict__ results, uint32_t n, uint32_t* counters,
                NzInflWork* infl, uint64_t* meta) {
    __shared__ WarpTabs tabs[INF_WARPS];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (;;) {
        uint32_t i = 0;
        if (lane == 0) i = atomicAdd(&counters[2], 1u);
        i = __shfl_sync(0xffffffffu, i, 0);
        if (i >= n) break;
        const NzIdxJob j = jobs[i];
        IdxBuild x;
        x.seg = reinterpret_cast<NzIdxSeg*>(j.seg); x.grp = reinterpret_cast<NzIdxGrp*>(j.grp);
        x.delta = reinterpret_cast<uint32_t*>(j.delta); x.side = reinterpret_cast<uint8_t*>(j.side);
        x.side_cap = j.side_cap; x.side_cur = 0; x.seg_cap = j.seg_cap; x.grp_cap = j.grp_cap; x.seg_bytes = j.seg_bytes;
        x.n_seg = 0; x.n_grp = 0; x.on = true; x.tok_in_sub = 0; x.sub_in_grp = 0;
        x.sub_bit0 = x.grp_bit0 = x.blk_hdr_bit = 0; x.sub_out0 = x.grp_out0 = 0;
        x.seg_out0 = 0; x.seg_first_grp = 0; x.seg_side0 = 0;
        uint32_t crc = 0;
        const int64_t got = inflate_member<true>(reinterpret_cast<const uint8_t*>(j.src), j.src_len,
                                                 reinterpret_cast<uint8_t*>(j.out), j.out_cap, tabs[warp], &crc, &x);
        __syncwarp();
        if (lane == 0) {
            NzIdxResult r;
            r.status = got < 0 ? -1 : (x.on ? 0 : 1);
            r.n_seg = x.n_seg; r.n_grp = x.n_grp; r.crc32 = crc;
            r.raw_len = got < 0 ? 0 : (uint64_t)got; r.side_bytes = x.side_cur;
            results[i] = r;
            NzInflWork w;
            w.src = j.src; w.src_len = j.src_len; w.dst = j.out; w.dst_cap = j.out_cap; w.req = i; w.pad = 0; w.idx = 0;
            infl[i] = w;
            meta[i] = got < 0 ? 0xFFFFFFFF00000000ull : (((uint64_t)(uint32_t)got << 32) | crc);
        }
    }
}

// Move a freshly built index from its temporary arrays into its blob in the arena (one CTA per member).
__global__ void __launch_bounds__(256)
k_index_pack(const NzIdxJob* __restrict__ jobs, const NzIdxResult* __restrict__ results, uint32_t n) {
    const uint32_t i = blockIdx.x;
    if (i >= n) return;
    const NzIdxJob j = jobs[i];
    const NzIdxResult r = results[i];
    if (j.dest == 0 || r.status != 0) return;
    uint8_t* d = reinterpret_cast<uint8_t*>(j.dest);
    NzIdxHeader h;
    h.magic = NZ_IDX_MAGIC; h.version = 1; h.n_seg = r.n_seg; h.n_grp = r.n_grp; h.raw_len = r.raw_len; h.side_bytes = r.side_bytes;
    h.off_seg = sizeof(NzIdxHeader);
    h.off_grp = h.off_seg + (uint64_t)r.n_seg * sizeof(NzIdxSeg);
    h.off_delta = h.off_grp + (uint64_t)r.n_grp * sizeof(NzIdxGrp);
    h.off_side = h.off_delta + (uint64_t)r.n_grp * NZ_IDX_SUBS * 4;
    if (threadIdx.x == 0) *reinterpret_cast<NzIdxHeader*>(d) = h;
    auto cp16 = [&](uint64_t off, uint64_t from, uint64_t bytes) {
        const uint4* s4 = reinterpret_cast<const uint4*>(from);
        uint4* d4 = reinterpret_cast<uint4*>(d + off);
        for (uint64_t k = threadIdx.x; k < (bytes + 15) / 16; k += blockDim.x) d4[k] = s4[k];
    };
    cp16(h.off_seg, j.seg, (uint64_t)r.n_seg * sizeof(NzIdxSeg));
    cp16(h.off_grp, j.grp, (uint64_t)r.n_grp * sizeof(NzIdxGrp));
    cp16(h.off_delta, j.delta, (uint64_t)r.n_grp * NZ_IDX_SUBS * 4);
    cp16(h.off_side, j.side, r.side_bytes);
}

// gzip trailer CRC-32 (IEEE) of every inflated stream, chunked like K3
__global__ void __launch_bounds__(256)
k_inflate_crc(const NzInflWork* __restrict__ infl, const uint64_t* __restrict__ meta, const uint32_t* __restrict__ counters,
              NzDevRes* res, const uint32_t* __restrict__ tables, const NzDevReq* __restrict__ reqs, uint32_t n_req) {
    __shared__ uint32_t sT[4 * 256];
    __shared__ uint32_t sZS8[1024 * 8];
    __shared__ uint32_t sred[8];
    if (counters[5] == 0) return;                         // nothing was decoded by the whole-member kernel
    const uint32_t* fast = tables + 16 * 256 + NZ_CRC_FAST_WORDS;                         // IEEE: ZS | K
    for (uint32_t i = threadIdx.x; i < 1024; i += 256) sT[i] = tables[8 * 256 + i];   // IEEE slice tables
    crc_load_fast_tables(sZS8, fast);
    __syncthreads();
    const uint32_t n = counters[1];
    for (uint32_t i = blockIdx.x; i < n; i += gridDim.x) {
        if (infl[i].idx) continue;                        // indexed member: its output was checked when the index was built
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
            if (reqs) infl_fail(infl[i], res, reqs, n_req);
            else { res[infl[i].req].status = NZ_ECORRUPT; res[infl[i].req].bytes = 0; }
        }
    }
}

// k_inflate_seg is built for 10 CTAs/SM (96 registers, a few spilled words; default: measured +13 % on the C3 wave) and for 8
// (115 registers, no spills); env NZ_INFL_OCC=8 picks the latter
static int seg_variant() {
    static const int v = [] { const char* e = getenv("NZ_INFL_OCC"); return e && atoi(e) == 8 ? 8 : 10; }();
    return v;
}
cudaError_t nz_launch_inflate(const NzInflWork* infl, uint64_t* meta, uint32_t* counters, NzDevRes* res,
                              const uint32_t* crc_tables, int grid, int sm_count, cudaStream_t s,
                              const uint32_t* seg_begin, const NzDevReq* reqs, uint32_t n_req, int seg_grid, bool legacy) {
    cudaError_t e;
    if (seg_grid > 0) {
        if (seg_variant() == 10) k_inflate_seg<10><<<seg_grid, SEG_WARPS * 32, 0, s>>>(infl, seg_begin, counters, res, reqs, n_req);
        else k_inflate_seg<8><<<seg_grid, SEG_WARPS * 32, 0, s>>>(infl, seg_begin, counters, res, reqs, n_req);
        e = cudaGetLastError();
        if (e != cudaSuccess) return e;
    }
    if (!legacy) return cudaSuccess;
    k_inflate<<<grid, INF_WARPS * 32, 0, s>>>(infl, meta, counters, res, reqs, n_req);
    e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    k_inflate_crc<<<sm_count * 4, 256, 0, s>>>(infl, meta, counters, res, crc_tables, reqs, n_req);
    return cudaGetLastError();
}

// counters: >= 8 zeroed words; res: n entries with status NZ_OK (k_inflate_crc flags CRC mismatches there)
cudaError_t nz_launch_index(const NzIdxJob* jobs, NzIdxResult* results, uint32_t n, uint32_t* counters, NzInflWork* infl,
                            uint64_t* meta, NzDevRes* res, const uint32_t* crc_tables, int sm_count, cudaStream_t s) {
    if (!n) return cudaSuccess;
    const uint32_t init[8] = {0, n, 0, 0, 0, n, 0, 0};                           // [1] = members, [2] = work cursor, [5] = members for the CRC pass
    cudaError_t e = cudaMemcpyAsync(counters, init, sizeof init, cudaMemcpyHostToDevice, s);
    if (e != cudaSuccess) return e;
    const int grid = (int)std::min<uint32_t>((n + INF_WARPS - 1) / INF_WARPS, (uint32_t)sm_count * 6u);
    k_inflate_index<<<grid, INF_WARPS * 32, 0, s>>>(jobs, results, n, counters, infl, meta);
    e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    k_inflate_crc<<<(int)std::min<uint32_t>(n, (uint32_t)sm_count * 4u), 256, 0, s>>>(infl, meta, counters, res, crc_tables, nullptr, 0);
    return cudaGetLastError();
}
cudaError_t nz_launch_index_pack(const NzIdxJob* jobs, const NzIdxResult* results, uint32_t n, cudaStream_t s) {
    if (!n) return cudaSuccess;
    k_index_pack<<<n, 256, 0, s>>>(jobs, results, n);
    return cudaGetLastError();
}

int nz_inflate_blocks_per_sm() {
    int nb = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_inflate, INF_WARPS * 32, 0) != cudaSuccess || nb < 1) nb = 8;
    return nb;
}
int nz_inflate_seg_blocks_per_sm() {
    int nb = 0;
    const cudaError_t e = seg_variant() == 10 ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_inflate_seg<10>, SEG_WARPS * 32, 0)
                                              : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_inflate_seg<8>, SEG_WARPS * 32, 0);
    if (e != cudaSuccess || nb < 1) nb = 4;
    return nb;
}

cudaError_t nz_inflate_upload_constants() { return nz_crc_upload_constants(); }
