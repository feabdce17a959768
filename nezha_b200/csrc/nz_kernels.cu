// nezha_b200/csrc/nz_kernels.cu — sm_100a kernels of the cache-hit path.
//
//   K1 k_lookup / probe_group4   key128 -> extent descriptor (warp-ballot over 128 B lines)
//   K6 k_insert / k_erase        claim (CAS) + publish / tombstone
//      k_plan                    one wave of GETs: probe, validate ranges, build copy work
//   K2 k_gather                  persistent range-slice gather into the response buffers
//   K3 k_crc_chunks / k_crc_fold CRC-32C chunk table + whole-object combine
//      k_verify                  re-checksum the chunks a GET touched
//
// Byte/integer work bounded by HBM bandwidth: no tensor cores.  What these
// replace in the reference: nginx's keys_zone rbtree lookup
// (/root/reference/examples/demo/nginx.conf:38) and its cache-file read loop +
// range filter.
#include <cuda_runtime.h>
#include <algorithm>
#include <stdint.h>
#include <stdlib.h>

#include "nz_internal.h"
#include "nz_crc.cuh"

// ---------------------------------------------------------------------------
// small PTX helpers
// ---------------------------------------------------------------------------
__device__ __forceinline__ uint4 ldg_stream(const void* p) {  // read-once data: do not pollute L1
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ uint4 ldg_cached(const void* p) {  // neighbours re-read this vector: keep in L1
    uint4 r;
    asm volatile("ld.global.nc.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ uint4 ldcg_ordered(const uint4* p) {   // L2 load the compiler may neither merge nor reorder
    uint4 r;
    asm volatile("ld.global.cg.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p) : "memory");
    return r;
}
__device__ __forceinline__ void stg_stream(void* p, const uint4& v) {
    asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};"
                 :: "l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

// ---------------------------------------------------------------------------
// K1: probe.  A 4-lane group owns one key; each lane reads one 32 B slot of the
// current 128 B line (one coalesced transaction per group), the group matches
// with a warp ballot.  Must be called by all 32 lanes of the warp.
// ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t probe_group4(const NzSlot* __restrict__ table, uint64_t line_mask,
                                                 uint64_t k0, uint64_t k1, bool active) {
    const unsigned lane = threadIdx.x & 31u;
    const unsigned gshift = lane & ~3u;
    const unsigned q = lane & 3u;
    uint64_t line = k1 & line_mask;
    uint32_t found = 0xFFFFFFFFu;
    bool done = !active;
    for (uint64_t probe = 0; probe <= line_mask; ++probe) {
        uint64_t s_k0 = 0, s_k1 = 0;
        uint32_t s_state = NZ_SLOT_BUSY, s_desc = 0;
        if (!done) {
            const uint4* sp = reinterpret_cast<const uint4*>(table + (line * 4 + q));
            const uint4 a = ldcg_ordered(sp), b = ldcg_ordered(sp + 1);
            // key and state/descriptor are two loads: if the slot was erased and re-used in between, the key read
            // again afterwards differs and the slot is taken as busy (no match, not empty: the probe moves on)
            const uint4 a2 = ldcg_ordered(sp);
            s_k0 = (uint64_t)a.x | ((uint64_t)a.y << 32);
            s_k1 = (uint64_t)a.z | ((uint64_t)a.w << 32);
            s_state = (a.x == a2.x && a.y == a2.y && a.z == a2.z && a.w == a2.w) ? b.x : (uint32_t)NZ_SLOT_BUSY;
            s_desc = b.y;
        }
        const bool match = !done && s_state == NZ_SLOT_LIVE && s_k0 == k0 && s_k1 == k1;
        const bool empty = !done && s_state == NZ_SLOT_EMPTY;
        const unsigned bm = __ballot_sync(0xffffffffu, match);
        const unsigned be = __ballot_sync(0xffffffffu, empty);
        const unsigned gm = (bm >> gshift) & 0xFu, ge = (be >> gshift) & 0xFu;
        const uint32_t d = __shfl_sync(0xffffffffu, s_desc, gshift + (gm ? (__ffs(gm) - 1) : 0));
        if (!done) {
            if (gm) { found = d; done = true; }
            else if (ge) { done = true; }
        }
        if (__all_sync(0xffffffffu, done)) break;
        line = (line + 1) & line_mask;
    }
    return found;
}

__global__ void __launch_bounds__(256) k_lookup(const NzSlot* __restrict__ table, uint64_t line_mask,
                                                const NzDesc* __restrict__ descs, const uint64_t* __restrict__ keys,
                                                uint32_t n, nz_extent* __restrict__ out) {
    const uint32_t groups_per_block = blockDim.x / 4;
    const uint32_t g = threadIdx.x / 4;
    const uint32_t stride = gridDim.x * groups_per_block;
    const uint32_t rounds = (n + stride - 1) / stride;
    for (uint32_t it = 0; it < rounds; ++it) {
        const uint32_t r = it * stride + blockIdx.x * groups_per_block + g;
        const bool active = r < n;
        uint64_t k0 = 0, k1 = 0;
        if (active) { k0 = keys[2 * r]; k1 = keys[2 * r + 1]; }
        const uint32_t d = probe_group4(table, line_mask, k0, k1, active);
        if (active && (threadIdx.x & 3u) == 0) {
            nz_extent e;
            if (d == 0xFFFFFFFFu) {
                e.status = NZ_ENOTFOUND; e.flags = 0; e.stored_length = 0; e.raw_length = 0;
                e.crc32c = 0; e.desc_index = 0xFFFFFFFFu; e.user_meta = 0;
            } else {
                const NzDesc ds = descs[d];
                e.status = NZ_OK; e.flags = ds.flags; e.stored_length = ds.stored_len; e.raw_length = ds.raw_len;
                e.crc32c = ds.crc32c; e.desc_index = d; e.user_meta = ds.user_meta;
            }
            out[r] = e;
        }
    }
}

// ---------------------------------------------------------------------------
// K6: insert (claim with CAS, publish with a fence) / erase (tombstone)
// ---------------------------------------------------------------------------
__global__ void k_insert(NzSlot* table, uint64_t line_mask, uint64_t k0, uint64_t k1, uint32_t desc, uint64_t gen,
                         int32_t* status) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    uint64_t line = k1 & line_mask;
    for (uint64_t probe = 0; probe <= line_mask; ++probe) {
        for (int q = 0; q < 4; ++q) {
            NzSlot* s = table + line * 4 + q;
            uint32_t st = *(volatile uint32_t*)&s->state;
            if (st == NZ_SLOT_LIVE) {
                if (s->k0 == k0 && s->k1 == k1) { *status = NZ_EEXIST; return; }
                continue;
            }
            if (st == NZ_SLOT_EMPTY || st == NZ_SLOT_TOMB) {
                if (atomicCAS(&s->state, st, (uint32_t)NZ_SLOT_BUSY) == st) {
                    s->k0 = k0; s->k1 = k1; s->desc = desc; s->gen = gen;
                    __threadfence();
                    *(volatile uint32_t*)&s->state = NZ_SLOT_LIVE;
                    *status = NZ_OK;
                    return;
                }
            }
        }
        line = (line + 1) & line_mask;
    }
    *status = NZ_EFULL;
}

__global__ void k_erase(NzSlot* table, uint64_t line_mask, uint64_t k0, uint64_t k1, int32_t* status) {
    const uint32_t d = probe_group4(table, line_mask, k0, k1, threadIdx.x < 4);
    (void)d;
    // re-walk to find the slot address (cheap; erase is rare)
    if (threadIdx.x != 0) return;
    uint64_t line = k1 & line_mask;
    for (uint64_t probe = 0; probe <= line_mask; ++probe) {
        bool saw_empty = false;
        for (int q = 0; q < 4; ++q) {
            NzSlot* s = table + line * 4 + q;
            uint32_t st = *(volatile uint32_t*)&s->state;
            if (st == NZ_SLOT_LIVE && s->k0 == k0 && s->k1 == k1) {
                *(volatile uint32_t*)&s->state = NZ_SLOT_TOMB;
                __threadfence();
                *status = NZ_OK;
                return;
            }
            if (st == NZ_SLOT_EMPTY) saw_empty = true;
        }
        if (saw_empty) break;
        line = (line + 1) & line_mask;
    }
    *status = NZ_ENOTFOUND;
}

// Batched publish: one thread per object writes its descriptor, fences, then claims and publishes a slot exactly
// like k_insert (the host reserved every key of the batch in its shadow, so no two threads insert the same key).
__global__ void __launch_bounds__(128)
k_publish(NzSlot* table, uint64_t line_mask, NzDesc* descs, const NzPubItem* __restrict__ items, uint32_t n, int32_t* status) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const NzPubItem it = items[i];
    descs[it.desc_idx] = it.d;
    __threadfence();
    uint64_t line = it.k1 & line_mask;
    for (uint64_t probe = 0; probe <= line_mask; ++probe) {
        for (int q = 0; q < 4; ++q) {
            NzSlot* s = table + line * 4 + q;
            const uint32_t st = *(volatile uint32_t*)&s->state;
            if (st == NZ_SLOT_LIVE) {
                if (s->k0 == it.k0 && s->k1 == it.k1) { status[i] = NZ_EEXIST; return; }
                continue;
            }
            if (st == NZ_SLOT_EMPTY || st == NZ_SLOT_TOMB) {
                if (atomicCAS(&s->state, st, (uint32_t)NZ_SLOT_BUSY) == st) {
                    s->k0 = it.k0; s->k1 = it.k1; s->desc = it.desc_idx; s->gen = it.gen;
                    __threadfence();
                    *(volatile uint32_t*)&s->state = NZ_SLOT_LIVE;
                    status[i] = NZ_OK;
                    return;
                }
            }
            // NZ_SLOT_BUSY: another insert is writing this slot; it is not ours, move on
        }
        line = (line + 1) & line_mask;
    }
    status[i] = NZ_EFULL;
}

// ---------------------------------------------------------------------------
// k_plan: one CTA turns a wave of GETs into copy / verify / inflate work.
// ---------------------------------------------------------------------------
struct PlanScratch {
    uint64_t* rq_src;   // n: device address of byte 0 of the served representation
    uint64_t* rq_len;   // n: its length
};

// small waves: the planning CTA copies the few tiles itself (defined after copy_tile)
__device__ void plan_fused_copy(const NzPlanArgs& a);

// exclusive scan over the 1024 threads of the planning CTA (warp shuffles + one pass over the 32 warp totals);
// *total (shared memory) receives the sum.  s_warp: 32 entries of shared memory.
template <typename Tv>
__device__ __forceinline__ Tv block_scan_excl_1024(Tv v, Tv* s_warp, Tv* total) {
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    Tv incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const Tv t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= (uint32_t)o) incl += t;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        const Tv w = s_warp[lane];
        Tv wi = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const Tv t = __shfl_up_sync(0xffffffffu, wi, o);
            if (lane >= (uint32_t)o) wi += t;
        }
        s_warp[lane] = wi - w;
        if (lane == 31) *total = wi;
    }
    __syncthreads();
    const Tv r = s_warp[warp] + incl - v;
    __syncthreads();                                  // s_warp may be reused right away
    return r;
}

template <bool SMALL, bool FUSED>
__global__ void __launch_bounds__(1024, 1)
k_plan(NzPlanArgs a, uint64_t* rq_src, uint64_t* rq_len, uint64_t* rq_tab, uint64_t* rq_ssz, const __grid_constant__ NzSmallWave small) {
    __shared__ uint64_t s_warp64[32];
    __shared__ uint64_t s_total64;
    __shared__ uint64_t s_carry;
    const uint32_t tid = threadIdx.x;
    const uint32_t g = tid >> 2;

    // ---- stage the wave's descriptors from mapped pinned memory into HBM: one coalesced PCIe
    // read instead of a round trip every time a later phase looks at a request or a slice -----
    {
        // (a small wave carries them as kernel parameters instead: no PCIe read at all)
        const uint4* s4 = SMALL ? reinterpret_cast<const uint4*>(small.reqs) : reinterpret_cast<const uint4*>(a.reqs_host);
        uint4* d4 = reinterpret_cast<uint4*>(a.reqs_dev);
        for (uint32_t i = tid; i < a.n * (uint32_t)(sizeof(NzDevReq) / 16); i += 1024) d4[i] = s4[i];
        s4 = SMALL ? reinterpret_cast<const uint4*>(small.segs) : reinterpret_cast<const uint4*>(a.segs_host);
        d4 = reinterpret_cast<uint4*>(a.segs_dev);
        for (uint32_t i = tid; i < a.nseg * (uint32_t)(sizeof(NzDevSeg) / 16); i += 1024) d4[i] = s4[i];
        if (tid < 16) a.counters[tid] = 0;
    }
    __threadfence_block();
    __syncthreads();
    const NzDevReq* __restrict__ reqs = a.reqs_dev;
    const NzDevSeg* __restrict__ segs = a.segs_dev;

    // ---- phase A: probe + per-request validation --------------------------
    for (uint32_t base = 0; base < a.n; base += 256) {
        const uint32_t r = base + g;
        const bool active = r < a.n;
        NzDevReq rq;
        if (active) rq = reqs[r];
        else { rq.k0 = rq.k1 = 0; }
        const uint32_t d = probe_group4(a.table, a.line_mask, rq.k0, rq.k1, active);
        if (!active || (tid & 3u) != 0) continue;
        NzDevRes out;
        out.status = NZ_OK; out.flags = 0; out.bytes = 0; out.length = 0; out.crc32c = 0; out.reserved = 0;
        uint64_t src = 0, len = 0, tab = 0, ssz = 0;
        if (d == 0xFFFFFFFFu) {
            out.status = NZ_ENOTFOUND;
        } else {
            const NzDesc ds = a.descs[d];
            out.flags = ds.flags;
            out.crc32c = ds.crc32c;
            const bool inflate = (rq.flags & NZ_GET_INFLATE) && (ds.flags & NZ_ENC_GZIP);
            len = inflate ? ds.raw_len : ds.stored_len;
            out.length = len;
            src = (uint64_t)(a.arena + ds.arena_off);
            if (ds.flags & NZ_ENC_STRIPED) {             // bytes live on several GPUs: arena_off is the stripe map
                tab = src;
                ssz = ds.stripe_bytes;
                src = 0;
                if (rq.whole || ssz == 0) out.status = NZ_EINVAL;   // the host always splits striped GETs per stripe
            }
            // validate slices against the representation and the destination
            uint64_t bytes = 0;
            if (rq.whole) {
                bytes = len;
                if (len > rq.dst_cap) out.status = NZ_ERANGE;
            } else {
                for (uint32_t i = 0; i < rq.nseg; ++i) {
                    const NzDevSeg sg = segs[rq.seg_begin + i];
                    if (sg.src_off > len || sg.len > len - sg.src_off) { out.status = NZ_EINVAL; break; }
                    if (sg.dst_off > rq.dst_cap || sg.len > rq.dst_cap - sg.dst_off) { out.status = NZ_ERANGE; break; }
                    bytes += sg.len;
                }
            }
            if (out.status == NZ_OK) out.bytes = bytes;
            if (inflate) {
                // Inflate dedupe: the host names one LEADER per distinct key of the wave.  The leader's member is inflated
                // once (into the leader's scratch when the key has several requesters or the request is not a whole-object
                // GET into HBM, else straight into its destination); every requester is then a plain copy out of it.
                const uint32_t L = rq.infl_leader;
                if (L != r) {
                    if (out.status == NZ_OK) {
                        bool okl = L < a.n;
                        NzDevReq lq;
                        if (okl) lq = reqs[L];
                        okl = okl && lq.scratch != 0 && lq.scratch_cap >= len && lq.k0 == rq.k0 && lq.k1 == rq.k1 && (lq.flags & NZ_GET_INFLATE);
                        if (!okl) { out.status = NZ_ENOMEM; out.bytes = 0; }
                        else src = lq.scratch;
                    }
                } else {
                    const bool direct = rq.whole && rq.scratch == 0;
                    const bool have_out = direct ? out.status == NZ_OK : (rq.scratch != 0 && rq.scratch_cap >= len);
                    if (!have_out) {
                        if (out.status == NZ_OK) { out.status = NZ_ENOMEM; out.bytes = 0; }
                    } else {
                        const uint32_t slot = atomicAdd(&a.counters[1], 1u);
                        NzInflWork w;
                        w.src = src; w.src_len = ds.stored_len;
                        w.dst = direct ? rq.dst : rq.scratch;
                        w.dst_cap = len;             // exactly the inflated length (validated against dst/scratch above)
                        w.req = r; w.pad = direct ? 1u : 0u;
                        w.idx = ds.stripe_bytes ? (uint64_t)(a.arena + ds.stripe_bytes) : 0;     // gzip objects: arena offset of the checkpoint index
                        a.infl[slot] = w;
                        a.infl_seg_begin[slot] = w.idx ? reinterpret_cast<const NzIdxHeader*>(w.idx)->n_seg : 0u;
                        if (!w.idx) atomicAdd(&a.counters[5], 1u);      // members without an index: the whole-member kernel runs
                        if (out.status == NZ_OK) {
                            src = direct ? 0 : rq.scratch;   // direct: nothing left to gather
                            if (direct) len = 0;
                        }
                    }
                }
            }
            // verify: inflate reads the whole member, so check every stored chunk
            if (out.status == NZ_OK && (rq.flags & NZ_GET_VERIFY) && inflate) {
                const uint32_t* tab = reinterpret_cast<const uint32_t*>(a.arena + ds.crc_tab_off);
                const uint32_t first = atomicAdd(&a.counters[0], ds.n_chunks);
                if ((uint64_t)first + ds.n_chunks > a.ver_cap) { out.status = NZ_ENOMEM; out.bytes = 0; }   // never return unverified bytes as verified
                for (uint32_t c = 0; c < ds.n_chunks && first + c < a.ver_cap; ++c) {
                    NzVerWork v;
                    const uint64_t off = (uint64_t)c * NZ_CRC_CHUNK;
                    v.ptr = (uint64_t)(a.arena + ds.arena_off + off);
                    v.len = (uint32_t)min((uint64_t)NZ_CRC_CHUNK, ds.stored_len - off);
                    v.expect = tab[c]; v.req = r; v.seg = NZ_VER_NOSEG; v.acc = 0; v.got = 0;
                    a.ver[first + c] = v;
                }
            }
        }
        if (out.status != NZ_OK) { len = 0; src = 0; tab = 0; }
        rq_src[r] = src;
        rq_len[r] = len;
        rq_tab[r] = tab;
        rq_ssz[r] = ssz;
        a.rq_stat[r] = out.status;
        a.rq_desc[r] = d;
        a.res[r] = out;
    }
    __syncthreads();

    // ---- phase A2: segment work list of the indexed inflate members (exclusive prefix of their segment counts) ----
    {
        const uint32_t ni = a.counters[1];
        if (ni) {
            __shared__ uint32_t s_warp32[32];
            __shared__ uint32_t s_total32;
            uint32_t carry = 0;
            for (uint32_t base = 0; base < ni; base += 1024) {
                const uint32_t i = base + tid;
                const uint32_t v = i < ni ? a.infl_seg_begin[i] : 0u;
                const uint32_t ex = block_scan_excl_1024<uint32_t>(v, s_warp32, &s_total32);
                if (i < ni) a.infl_seg_begin[i] = carry + ex;
                carry += s_total32;
                __syncthreads();
            }
            if (tid == 0) {
                a.infl_seg_begin[ni] = carry;
                a.ver_stat[3] = carry;                 // statistics: index segments / members without an index in this wave
                a.ver_stat[4] = a.counters[5];
            }
        } else if (a.infl_seg_begin && tid == 0) {
            // a wave that carries NZ_GET_INFLATE but names no gzip object: the segment kernel still runs and reads its
            // total from seg_begin[0] - in a recycled ticket that word is whatever the previous wave left there
            a.infl_seg_begin[0] = 0;
        }
    }

    // ---- phase B1: per-segment copy work and tile counts ----------------------
    uint64_t* tile_cnt = a.tile_begin;           // counts first, scanned in place in B3
    for (uint32_t s = tid; s < a.nseg; s += 1024) {
        uint64_t tiles = 0;
        const NzDevSeg sg = segs[s];
        const NzDevReq& rq = reqs[sg.req];
        const int32_t st = a.rq_stat[sg.req];
        NzSegWork w;
        w.src = 0; w.dst = 0; w.len = 0; w.vstart = 0;
        const uint64_t rlen = rq_len[sg.req];
        if (st == NZ_OK && rlen > 0) {
            const uint64_t off = rq.whole ? 0 : sg.src_off;
            const uint64_t len = rq.whole ? rlen : sg.len;
            const uint64_t doff = rq.whole ? 0 : sg.dst_off;
            w.src = rq_src[sg.req] + off;
            const uint64_t tab = rq_tab[sg.req];
            uint64_t ok_len = len;
            if (tab) {                               // K5: the slice lies inside one stripe, maybe on a peer GPU
                const uint64_t ssz = rq_ssz[sg.req];
                const uint64_t k = off / ssz, within = off - k * ssz;
                const NzDevStripe stp = reinterpret_cast<const NzDevStripe*>(tab)[k];
                const uint64_t pbase = a.peer_base ? a.peer_base[stp.shard] : 0;
                if (within + len > stp.len || pbase == 0) {
                    ok_len = 0;
                    a.res[sg.req].status = pbase == 0 ? NZ_ENOSYS : NZ_EINVAL;
                    a.res[sg.req].bytes = 0;
                }
                w.src = pbase + stp.arena_off + within;
            }
            w.dst = rq.dst + doff;
            w.len = ok_len;
            w.vstart = -(int64_t)(w.dst & 127u);
            tiles = ok_len ? (ok_len + (w.dst & 127u) + a.tile_bytes - 1) / a.tile_bytes : 0;
        }
        a.work[s] = w;
        tile_cnt[s] = tiles;
        a.chain_head[s] = 0xFFFFFFFFu;
        a.chain_next[s] = 0xFFFFFFFFu;
    }
    __threadfence_block();
    __syncthreads();

    // ---- phase B2: coalesce identical slices (read once, write many) -----------
    // Concurrent GETs of the same bytes are the common case for a dataset cache (every
    // worker pulls the same shard).  Slices with the same source, length and destination
    // phase are chained behind one leader: the gather loads each tile once and stores it to
    // every member, so HBM reads the source once instead of once per request.
    if (a.coalesce && a.nseg > 1 && (a.nseg <= 2048 || a.coal_tab)) {
        __shared__ int s_tab[4096];
        // waves of up to 2048 slices hash in shared memory, larger ones in a table the host put in HBM
        int* tabp = a.nseg <= 2048 ? s_tab : a.coal_tab;
        const uint32_t tmask = (a.nseg <= 2048 ? 4096u : a.coal_cap) - 1u;
        for (uint32_t i = tid; i <= tmask; i += 1024) tabp[i] = -1;
        __threadfence_block();
        __syncthreads();
        for (uint32_t s = tid; s < a.nseg; s += 1024) {
            if (tile_cnt[s] == 0) continue;
            const NzSegWork w = a.work[s];
            uint64_t hsh = w.src * 0x9E3779B97F4A7C15ull ^ (w.len + (w.dst & 127u)) * 0xC2B2AE3D27D4EB4Full;
            uint32_t h = (uint32_t)(hsh >> 40) & tmask;
            for (;;) {
                int cur = *(volatile int*)&tabp[h];
                if (cur == -1) {
                    cur = atomicCAS(&tabp[h], -1, (int)s);
                    if (cur == -1) break;                       // this slice leads its group
                }
                const NzSegWork o = a.work[cur];
                if (o.src == w.src && o.len == w.len && ((o.dst ^ w.dst) & 127u) == 0) {
                    tile_cnt[s] = 0;                            // follower: its bytes are written by the leader's tiles
                    a.chain_next[s] = atomicExch(&a.chain_head[cur], s);
                    break;
                }
                h = (h + 1) & tmask;
            }
        }
        __syncthreads();
    }

    // ---- phase B3: exclusive scan of tile counts -> tile_begin -------------------
    if (tid == 0) s_carry = 0;
    __syncthreads();
    for (uint32_t base = 0; base < a.nseg; base += 1024) {
        const uint32_t s = base + tid;
        const uint64_t tiles = s < a.nseg ? tile_cnt[s] : 0;
        const uint64_t ex = block_scan_excl_1024<uint64_t>(tiles, s_warp64, &s_total64);
        const uint64_t carry = s_carry;
        if (s < a.nseg) a.tile_begin[s] = carry + ex;
        __syncthreads();
        if (tid == 0) s_carry = carry + s_total64;
        __syncthreads();
    }
    if (tid == 0) a.tile_begin[a.nseg] = s_carry;
    if constexpr (FUSED) {
        __syncthreads();                         // work[] and tile_begin[] of this CTA are visible to all its threads
        plan_fused_copy(a);
        // completion flag in pinned host memory: results and payload first (system-scope fence by every writer), then
        // the flag - the host polls a cache line instead of querying an event
        __threadfence_system();
        __syncthreads();
        if (tid == 0) *reinterpret_cast<volatile uint32_t*>(a.ver_stat + 2) = 1u;
    }
}

// Verify work for stored (not inflated) bytes, emitted after coalescing is known: one warp per
// slice.  Only the LEADER of a coalesced group gets entries (its followers deliver the same source
// bytes; k_verify flags every member), and it gets them if any member asked for NZ_GET_VERIFY.
// A chunk wholly inside a slice that the bulk-copy engine moves is marked NZ_VER_FUSED: the gather
// checksums it from shared memory and k_verify only compares.
__global__ void __launch_bounds__(1024)
k_plan_verify(NzPlanArgs a) {
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    uint32_t n_fused = 0;
    for (uint32_t s = warp; s < a.nseg; s += 32) {
        a.seg_ver[s] = 0xFFFFFFFFu;
        const NzSegWork w = a.work[s];
        if (w.len == 0 || a.tile_begin[s + 1] == a.tile_begin[s]) continue;          // failed / follower
        const NzDevSeg sg = a.segs_dev[s];
        bool want = (a.reqs_dev[sg.req].flags & NZ_GET_VERIFY) != 0;
        for (uint32_t m = a.chain_head[s]; !want && m != 0xFFFFFFFFu; m = a.chain_next[m])
            want = (a.reqs_dev[a.segs_dev[m].req].flags & NZ_GET_VERIFY) != 0;
        if (!want) continue;
        const NzDevReq rq = a.reqs_dev[sg.req];
        const uint32_t d = a.rq_desc[sg.req];
        if (d == 0xFFFFFFFFu) continue;
        const NzDesc ds = a.descs[d];
        if ((rq.flags & NZ_GET_INFLATE) && (ds.flags & NZ_ENC_GZIP)) continue;       // done in k_plan (whole member)
        // where the checked bytes and their chunk table live: the object's extent, or - for a striped object, whose GET the
        // host split at stripe boundaries - the one stripe this slice lies in, on whichever GPU holds it (its chunk
        // table sits next to it in that GPU's arena and is read over NVLink like the bytes)
        const uint8_t* vbase_ptr = a.arena + ds.arena_off;
        const uint32_t* tab = reinterpret_cast<const uint32_t*>(a.arena + ds.crc_tab_off);
        uint64_t vlen = ds.stored_len;
        uint64_t off = rq.whole ? 0 : sg.src_off;
        const uint64_t len = w.len;
        if (ds.flags & NZ_ENC_STRIPED) {
            const uint64_t k = off / ds.stripe_bytes;
            const NzDevStripe stp = reinterpret_cast<const NzDevStripe*>(a.arena + ds.arena_off)[k];
            const uint64_t pbase = a.peer_base ? a.peer_base[stp.shard] : 0;
            if (pbase == 0) continue;                    // k_plan already failed this request (NZ_ENOSYS)
            off -= k * ds.stripe_bytes;
            vbase_ptr = reinterpret_cast<const uint8_t*>(pbase + stp.arena_off);
            tab = reinterpret_cast<const uint32_t*>(pbase + stp.crc_tab_off);
            vlen = stp.len;
        }
        const uint32_t c0 = (uint32_t)(off / NZ_CRC_CHUNK), c1 = (uint32_t)((off + len - 1) / NZ_CRC_CHUNK);
        const uint32_t cnt = c1 - c0 + 1;
        uint32_t first = 0;
        if (lane == 0) first = atomicAdd(&a.counters[0], cnt);
        first = __shfl_sync(0xffffffffu, first, 0);
        const bool fits = (uint64_t)first + cnt <= a.ver_cap;
        if (!fits && lane == 0) {                 // SPEC X4: a verify that cannot run is an error, never a silent NZ_OK
            if (a.reqs_dev[sg.req].flags & NZ_GET_VERIFY) { a.res[sg.req].status = NZ_ENOMEM; a.res[sg.req].bytes = 0; }
            for (uint32_t m = a.chain_head[s]; m != 0xFFFFFFFFu; m = a.chain_next[m]) {
                const uint32_t r2 = a.segs_dev[m].req;
                if (a.reqs_dev[r2].flags & NZ_GET_VERIFY) { a.res[r2].status = NZ_ENOMEM; a.res[r2].bytes = 0; }
            }
        }
        const bool fusable = a.fuse_verify && fits && ((w.src | w.dst) & 15u) == 0;
        for (uint32_t c = lane; c < cnt && first + c < a.ver_cap; c += 32) {
            NzVerWork v;
            const uint64_t coff = (uint64_t)(c0 + c) * NZ_CRC_CHUNK;
            v.ptr = (uint64_t)(vbase_ptr + coff);
            v.len = (uint32_t)min((uint64_t)NZ_CRC_CHUNK, vlen - coff);
            const bool inside = coff >= off && coff + v.len <= off + len;
            v.expect = tab[c0 + c]; v.req = sg.req;
            v.seg = s | ((fusable && inside) ? NZ_VER_FUSED : 0u);
            v.acc = 0; v.got = 0;
            a.ver[first + c] = v;
            n_fused += (fusable && inside) ? 1u : 0u;
        }
        if (fusable && lane == 0) a.seg_ver[s] = first;
    }
    if (n_fused) atomicAdd(&a.counters[3], n_fused);
    __syncthreads();
    if (threadIdx.x == 0) { a.ver_stat[0] = min(a.counters[0], a.ver_cap); a.ver_stat[1] = a.counters[3]; }
}

// ---------------------------------------------------------------------------
// K2: range-slice gather.  Persistent grid; a tile is NZ_TILE bytes of one
// segment, cut so that every interior tile boundary is 128 B aligned in the
// DESTINATION (full-line writes).  The source may sit at any byte offset
// (Range: bytes=a-b), so vectors are rebuilt from two aligned 16 B loads with a
// funnel shift; every global access is a 16 B vector on a 16 B boundary.
// ---------------------------------------------------------------------------
template <int WS>   // WS = (source misalignment) / 4, 0..3 ; bs = 8 * (misalignment % 4)
__device__ __forceinline__ uint4 realign(const uint4 a, const uint4 b, unsigned bs) {
    uint4 o;
    if constexpr (WS == 0) {
        o.x = __funnelshift_r(a.x, a.y, bs); o.y = __funnelshift_r(a.y, a.z, bs);
        o.z = __funnelshift_r(a.z, a.w, bs); o.w = __funnelshift_r(a.w, b.x, bs);
    } else if constexpr (WS == 1) {
        o.x = __funnelshift_r(a.y, a.z, bs); o.y = __funnelshift_r(a.z, a.w, bs);
        o.z = __funnelshift_r(a.w, b.x, bs); o.w = __funnelshift_r(b.x, b.y, bs);
    } else if constexpr (WS == 2) {
        o.x = __funnelshift_r(a.z, a.w, bs); o.y = __funnelshift_r(a.w, b.x, bs);
        o.z = __funnelshift_r(b.x, b.y, bs); o.w = __funnelshift_r(b.y, b.z, bs);
    } else {
        o.x = __funnelshift_r(a.w, b.x, bs); o.y = __funnelshift_r(b.x, b.y, bs);
        o.z = __funnelshift_r(b.y, b.z, bs); o.w = __funnelshift_r(b.z, b.w, bs);
    }
    return o;
}

// destinations of one tile: the leader's and, through the chain, every coalesced follower's
struct DstChain {
    uint8_t* d0;                                   // leader destination of this tile's first byte
    int64_t lo;                                    // offset of the tile inside the slice
    const NzSegWork* __restrict__ work;
    const uint32_t* __restrict__ next;
    uint32_t head;                                 // first follower or 0xFFFFFFFF
};

template <int WS, int NT, int U2>
__device__ __forceinline__ void copy_vectors_shifted(const uint4* __restrict__ sa /*aligned base*/, const DstChain& c, uint32_t dskip,
                                                     uint32_t nvec, unsigned bs) {
    constexpr int U = U2 > 1 ? U2 / 2 : 1;   // two loads per output vector
    constexpr uint32_t STEP = NT * U;
    const uint32_t tid = threadIdx.x;
    uint32_t base = 0;
    for (; base + STEP <= nvec; base += STEP) {   // full steps: no predicates, all loads issued before any store
        uint4 v[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const uint32_t i = base + u * NT + tid;
            v[u] = realign<WS>(ldg_cached(sa + i), ldg_cached(sa + i + 1), bs);
        }
        uint4* d = reinterpret_cast<uint4*>(c.d0 + dskip);
#pragma unroll
        for (int u = 0; u < U; ++u) stg_stream(d + base + u * NT + tid, v[u]);
        for (uint32_t m = c.head; m != 0xFFFFFFFFu; m = __ldg(c.next + m)) {
            uint4* dm = reinterpret_cast<uint4*>(reinterpret_cast<uint8_t*>(__ldg(&c.work[m].dst)) + c.lo + dskip);
#pragma unroll
            for (int u = 0; u < U; ++u) stg_stream(dm + base + u * NT + tid, v[u]);
        }
    }
    for (uint32_t i = base + tid; i < nvec; i += NT) {
        const uint4 v = realign<WS>(ldg_cached(sa + i), ldg_cached(sa + i + 1), bs);
        stg_stream(reinterpret_cast<uint4*>(c.d0 + dskip) + i, v);
        for (uint32_t m = c.head; m != 0xFFFFFFFFu; m = __ldg(c.next + m))
            stg_stream(reinterpret_cast<uint4*>(reinterpret_cast<uint8_t*>(__ldg(&c.work[m].dst)) + c.lo + dskip) + i, v);
    }
}

template <int NT, int U>
__device__ __forceinline__ void copy_vectors_aligned(const uint4* __restrict__ sv, const DstChain& c, uint32_t dskip, uint32_t nvec) {
    constexpr uint32_t STEP = NT * U;
    const uint32_t tid = threadIdx.x;
    uint32_t base = 0;
    for (; base + STEP <= nvec; base += STEP) {
        uint4 v[U];
#pragma unroll
        for (int u = 0; u < U; ++u) v[u] = ldg_stream(sv + base + u * NT + tid);
        uint4* d = reinterpret_cast<uint4*>(c.d0 + dskip);
#pragma unroll
        for (int u = 0; u < U; ++u) stg_stream(d + base + u * NT + tid, v[u]);
        for (uint32_t m = c.head; m != 0xFFFFFFFFu; m = __ldg(c.next + m)) {
            uint4* dm = reinterpret_cast<uint4*>(reinterpret_cast<uint8_t*>(__ldg(&c.work[m].dst)) + c.lo + dskip);
#pragma unroll
            for (int u = 0; u < U; ++u) stg_stream(dm + base + u * NT + tid, v[u]);
        }
    }
    for (uint32_t i = base + tid; i < nvec; i += NT) {
        const uint4 v = ldg_stream(sv + i);
        stg_stream(reinterpret_cast<uint4*>(c.d0 + dskip) + i, v);
        for (uint32_t m = c.head; m != 0xFFFFFFFFu; m = __ldg(c.next + m))
            stg_stream(reinterpret_cast<uint4*>(reinterpret_cast<uint8_t*>(__ldg(&c.work[m].dst)) + c.lo + dskip) + i, v);
    }
}

template <int NT, int U>
__device__ __forceinline__ void copy_tile(const uint8_t* __restrict__ s, const DstChain& c, uint32_t n) {
    const uint32_t tid = threadIdx.x;
    // head: bring the destination to a 16 B boundary (only the first tile of a slice can need it;
    // coalesced members share the destination phase, so the same split applies to all of them)
    const uint32_t head = min(n, (uint32_t)((16u - ((uintptr_t)c.d0 & 15u)) & 15u));
    const uint32_t nvec = (n - head) >> 4;
    const uint32_t tail = (n - head) & 15u;
    if (tid < head || tid < tail) {
        const uint32_t toff = head + (nvec << 4);
        const uint8_t hb = tid < head ? s[tid] : 0, tb = tid < tail ? s[toff + tid] : 0;
        if (tid < head) c.d0[tid] = hb;
        if (tid < tail) c.d0[toff + tid] = tb;
        for (uint32_t m = c.head; m != 0xFFFFFFFFu; m = __ldg(c.next + m)) {
            uint8_t* dm = reinterpret_cast<uint8_t*>(__ldg(&c.work[m].dst)) + c.lo;
            if (tid < head) dm[tid] = hb;
            if (tid < tail) dm[toff + tid] = tb;
        }
    }
    if (nvec == 0) return;
    s += head;
    const unsigned mis = (unsigned)((uintptr_t)s & 15u);
    if (mis == 0) {
        copy_vectors_aligned<NT, U>(reinterpret_cast<const uint4*>(s), c, head, nvec);
    } else {
        const uint4* sa = reinterpret_cast<const uint4*>(s - mis);
        const unsigned bs = (mis & 3u) * 8u;
        switch (mis >> 2) {
        case 0: copy_vectors_shifted<0, NT, U>(sa, c, head, nvec, bs); break;
        case 1: copy_vectors_shifted<1, NT, U>(sa, c, head, nvec, bs); break;
        case 2: copy_vectors_shifted<2, NT, U>(sa, c, head, nvec, bs); break;
        default: copy_vectors_shifted<3, NT, U>(sa, c, head, nvec, bs); break;
        }
    }
}

// tile -> segment lookup shared by the gather variants (uniform across the CTA)
// A wave of a few small slices (a response's first window): one launch instead of two.  The planning CTA's
// 1024 threads move the tiles (32 KiB = 1024 x 2 x 16 B) right after planning them; no coalescing chains (the
// host turns coalescing off for such waves), plain loads of what this kernel wrote itself.
__device__ void plan_fused_copy(const NzPlanArgs& a) {
    DstChain dc;
    dc.work = a.work;
    dc.next = a.chain_next;
    dc.head = 0xFFFFFFFFu;
    for (uint32_t s = 0; s < a.nseg; ++s) {
        const NzSegWork w = a.work[s];
        const uint64_t t0 = a.tile_begin[s], t1 = a.tile_begin[s + 1];
        for (uint64_t t = t0; t < t1; ++t) {
            int64_t lo = w.vstart + (int64_t)((t - t0) * a.tile_bytes), hi = lo + a.tile_bytes;
            if (lo < 0) lo = 0;
            if (hi > (int64_t)w.len) hi = (int64_t)w.len;
            dc.d0 = reinterpret_cast<uint8_t*>(w.dst) + lo;
            dc.lo = lo;
            copy_tile<1024, 2>(reinterpret_cast<const uint8_t*>(w.src) + lo, dc, (uint32_t)(hi - lo));
        }
    }
}

struct TileCursor {
    uint64_t lo_t = 0, hi_t = 0, src = 0, dst = 0, len = 0;
    int64_t vstart = 0;
    uint32_t seg = 0;
    __device__ __forceinline__ void seek(uint64_t tile, const NzSegWork* __restrict__ work, const uint64_t* __restrict__ tile_begin,
                                         uint32_t nseg) {
        if (tile >= lo_t && tile < hi_t) return;
        uint32_t a = 0, b = nseg;
        while (b - a > 1) {                    // last segment with tile_begin[seg] <= tile
            const uint32_t m = (a + b) >> 1;
            if (__ldg(tile_begin + m) <= tile) a = m; else b = m;
        }
        seg = a;
        lo_t = __ldg(tile_begin + a);
        hi_t = __ldg(tile_begin + a + 1);
        const ulonglong2 p0 = __ldg(reinterpret_cast<const ulonglong2*>(work + a));
        const ulonglong2 p1 = __ldg(reinterpret_cast<const ulonglong2*>(work + a) + 1);
        src = p0.x; dst = p0.y; len = p1.x; vstart = (int64_t)p1.y;
    }
    __device__ __forceinline__ void span(uint64_t tile, uint32_t tile_bytes, int64_t& lo, int64_t& hi) const {
        lo = vstart + (int64_t)((tile - lo_t) * tile_bytes);
        hi = lo + tile_bytes;
        if (lo < 0) lo = 0;
        if (hi > (int64_t)len) hi = (int64_t)len;
    }
};

// Tile order.  Normally CTA b takes tiles b, b + grid, ...: the whole grid works on one neighbourhood of the wave,
// which keeps DRAM pages of the destinations open.  A wave with striped sources (K5) instead walks the tile list as a
// `spread` x C matrix column by column: the CTAs running at any moment sit at `spread` evenly spaced positions of the
// wave, i.e. in stripes of different peer GPUs, so the NVLink pulls of one GPU are spread over all its peers instead
// of hammering one peer's egress at a time.
struct TileOrder {
    uint64_t total, cols;
    uint32_t spread;
    __device__ __forceinline__ TileOrder(uint64_t total_, uint32_t spread_) : total(total_), cols(0), spread(spread_) {
        if (spread > 1 && total > 4ull * spread) cols = (total + spread - 1) / spread; else spread = 0;
    }
    __device__ __forceinline__ uint64_t steps() const { return spread ? cols * spread : total; }
    // tile of step i, or ~0 when the step falls off the ragged last column
    __device__ __forceinline__ uint64_t tile(uint64_t i) const {
        if (!spread) return i;
        const uint64_t t = (i % spread) * cols + i / spread;
        return t < total ? t : ~0ull;
    }
};

// Which step a CTA takes next.  Static: b, b + grid, ... (HBM-bound waves: every tile costs about the same).  Dynamic
// (`dyn` = a zeroed counter): the next free step, drawn one iteration ahead so the atomic's latency hides behind the
// current tile.  Waves with striped sources use it: a tile pulled over NVLink takes several times as long as a local
// one, and with a static split the CTAs that happened to get more remote tiles finish last while NVLink idles.
struct StepIter {
    uint64_t i;
    uint32_t* dyn;
    uint64_t* slot;          // shared memory, one word
    uint32_t ahead;          // thread 0: the step after this one (dynamic only)
    __device__ __forceinline__ void begin(uint32_t* dyn_, uint64_t* slot_) {
        dyn = dyn_; slot = slot_; ahead = 0;
        if (!dyn) { i = blockIdx.x; return; }
        if (threadIdx.x == 0) { *slot = atomicAdd(dyn, 1u); ahead = atomicAdd(dyn, 1u); }
        __syncthreads();
        i = *slot;
    }
    __device__ __forceinline__ void next() {
        if (!dyn) { i += gridDim.x; return; }
        __syncthreads();                                   // everyone has read the current step
        if (threadIdx.x == 0) { *slot = ahead; ahead = atomicAdd(dyn, 1u); }
        __syncthreads();
        i = *slot;
    }
    // thread 0 only: the step this CTA takes after the current one (the fused-verify kernel loads it ahead)
    __device__ __forceinline__ uint64_t peek() const { return dyn ? (uint64_t)ahead : i + gridDim.x; }
};

// Register path on its own (NZ_GATHER_VARIANT=0): the fallback and the A/B partner of the bulk-copy kernel.
template <int NT, int U, int MINB>
__global__ void __launch_bounds__(NT, MINB)
k_gather(const NzSegWork* __restrict__ work, const uint64_t* __restrict__ tile_begin, const uint32_t* __restrict__ chain_head,
         const uint32_t* __restrict__ chain_next, uint32_t nseg, uint32_t tile_bytes, uint32_t spread, uint32_t* dyn) {
    __shared__ uint64_t s_step;
    const TileOrder ord(tile_begin[nseg], spread);
    TileCursor c;
    DstChain dc;
    dc.work = work;
    dc.next = chain_next;
    dc.head = 0xFFFFFFFFu;
    StepIter it;
    for (it.begin(dyn, &s_step); it.i < ord.steps(); it.next()) {
        const uint64_t tile = ord.tile(it.i);
        if (tile == ~0ull) continue;
        const uint32_t before = c.seg;
        const bool fresh = !(tile >= c.lo_t && tile < c.hi_t);
        c.seek(tile, work, tile_begin, nseg);
        if (fresh || before != c.seg) dc.head = __ldg(chain_head + c.seg);
        int64_t lo, hi;
        c.span(tile, tile_bytes, lo, hi);
        dc.d0 = reinterpret_cast<uint8_t*>(c.dst) + lo;
        dc.lo = lo;
        copy_tile<NT, U>(reinterpret_cast<const uint8_t*>(c.src) + lo, dc, (uint32_t)(hi - lo));
    }
}

// ---------------------------------------------------------------------------
// K3: CRC-32C chunk table + whole-object fold (device code in nz_crc.cuh)
// ---------------------------------------------------------------------------
template <bool TREE>
__global__ void __launch_bounds__(256, 4)
k_crc_chunks(const uint8_t* __restrict__ p, uint64_t n, uint32_t* __restrict__ chunk_out, const uint32_t* __restrict__ tables) {
    __shared__ uint32_t sT[4 * 256];
    __shared__ uint32_t sZS8[1024 * 8];
    __shared__ uint32_t sred[8];
    const uint32_t* fast = tables + 16 * 256;                 // Castagnoli: ZS | K
    for (uint32_t i = threadIdx.x; i < 1024; i += 256) sT[i] = tables[i];
    crc_load_fast_tables(sZS8, fast);
    __syncthreads();
    const uint64_t nchunks = (n + NZ_CRC_CHUNK - 1) / NZ_CRC_CHUNK;
    const bool aligned = (((uintptr_t)p) & 15u) == 0;
    for (uint64_t c = blockIdx.x; c < nchunks; c += gridDim.x) {
        const uint64_t off = c * NZ_CRC_CHUNK;
        const uint32_t m = (uint32_t)min((uint64_t)NZ_CRC_CHUNK, n - off);
        const uint32_t crc = (aligned && m == NZ_CRC_CHUNK) ? block_crc_chunk64k<0, TREE>(p + off, sZS8, sT, fast + 1024, sred)
                                                            : block_crc_chunk<0>(p + off, m, sT, sred);
        if (threadIdx.x == 0) chunk_out[c] = crc;
    }
}

// Batched form for bulk loads of small objects (nz_put_commit_many): one launch checksums every chunk of every object of
// the batch (a flat job list), one more folds each object's chunks (one thread per object).  Per-object launches cost
// ~2 us each on the host: at 10^6 objects they were half of the fill time.
__global__ void __launch_bounds__(256, 4)
k_crc_jobs(const NzCrcJob* __restrict__ jobs, uint32_t n, const uint32_t* __restrict__ tables) {
    __shared__ uint32_t sT[4 * 256];
    __shared__ uint32_t sZS8[1024 * 8];
    __shared__ uint32_t sred[8];
    const uint32_t* fast = tables + 16 * 256;
    for (uint32_t i = threadIdx.x; i < 1024; i += 256) sT[i] = tables[i];
    crc_load_fast_tables(sZS8, fast);
    __syncthreads();
    for (uint32_t j = blockIdx.x; j < n; j += gridDim.x) {
        const NzCrcJob jb = jobs[j];
        const uint8_t* p = reinterpret_cast<const uint8_t*>(jb.ptr);
        const uint32_t crc = (jb.len == NZ_CRC_CHUNK && (jb.ptr & 15u) == 0) ? block_crc_chunk64k<0>(p, sZS8, sT, fast + 1024, sred)
                                                                              : block_crc_chunk<0>(p, jb.len, sT, sred);
        if (threadIdx.x == 0) *reinterpret_cast<uint32_t*>(jb.out) = crc;
    }
}
__global__ void __launch_bounds__(128)
k_crc_fold_jobs(const NzFoldJob* __restrict__ jobs, uint32_t n) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const NzFoldJob jb = jobs[i];
    const uint32_t* tab = reinterpret_cast<const uint32_t*>(jb.tab);
    const uint64_t nch = (jb.len + NZ_CRC_CHUNK - 1) / NZ_CRC_CHUNK;
    uint32_t acc = 0;                                              // CRC-32C of the empty string
    const uint32_t x64k = c_x8pow[0][16];
    for (uint64_t c = 0; c < nch; ++c) {
        const uint64_t cl = min((uint64_t)NZ_CRC_CHUNK, jb.len - c * NZ_CRC_CHUNK);
        acc = c == 0 ? tab[0] : (gf_mul<0>(cl == NZ_CRC_CHUNK ? x64k : gf_xpow8<0>(cl), acc) ^ tab[c]);
    }
    *reinterpret_cast<uint32_t*>(jb.whole_out) = acc;
}

// whole-object CRC = sum_i chunk_i * x^(8 * bytes_after_i): each thread folds a
// contiguous run of chunks Horner-style, then one shift to the object end.
__global__ void __launch_bounds__(1024)
k_crc_fold(const uint32_t* __restrict__ chunks, uint64_t n, uint32_t* __restrict__ whole_out) {
    __shared__ uint32_t sred[32];
    const uint64_t nchunks = (n + NZ_CRC_CHUNK - 1) / NZ_CRC_CHUNK;
    const uint32_t tid = threadIdx.x;
    const uint64_t per = (nchunks + 1023) / 1024;
    const uint64_t c0 = (uint64_t)tid * per;
    const uint64_t c1 = min(nchunks, c0 + per);
    uint32_t acc = 0;
    if (c0 < c1) {
        const uint32_t x64k = c_x8pow[0][16];                       // x^(8*65536)
        for (uint64_t c = c0; c < c1; ++c) {
            const uint64_t len = min((uint64_t)NZ_CRC_CHUNK, n - c * NZ_CRC_CHUNK);
            const uint32_t mul = (len == NZ_CRC_CHUNK) ? x64k : gf_xpow8<0>(len);
            acc = (c == c0) ? chunks[c] : (gf_mul<0>(mul, acc) ^ chunks[c]);
        }
        const uint64_t end = min(n, c1 * (uint64_t)NZ_CRC_CHUNK);
        const uint64_t after = n - end;
        if (after) acc = gf_mul<0>(gf_xpow8<0>(after), acc);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc ^= __shfl_xor_sync(0xffffffffu, acc, o);
    if ((tid & 31u) == 0) sred[tid >> 5] = acc;
    __syncthreads();
    if (tid == 0) {
        uint32_t r = 0;
        for (int i = 0; i < 32; ++i) r ^= sred[i];
        *whole_out = r;
    }
}

__global__ void __launch_bounds__(256)
k_verify(NzPlanArgs a, const uint32_t* __restrict__ tables) {
    __shared__ uint32_t sT[4 * 256];
    __shared__ uint32_t sZS8[1024 * 8];
    __shared__ uint32_t sred[8];
    const uint32_t* fast = tables + 16 * 256;
    for (uint32_t i = threadIdx.x; i < 1024; i += 256) sT[i] = tables[i];
    crc_load_fast_tables(sZS8, fast);
    __syncthreads();
    const uint32_t n = min(a.counters[0], a.ver_cap);
    for (uint32_t c = blockIdx.x; c < n; c += gridDim.x) {
        const NzVerWork v = a.ver[c];
        bool bad;
        if (v.seg & NZ_VER_FUSED) {                       // the gather already checksummed these bytes from shared memory
            bad = v.got != v.len || ~v.acc != v.expect;
        } else {
            const uint8_t* vp = reinterpret_cast<const uint8_t*>(v.ptr);
            const uint32_t crc = (v.len == NZ_CRC_CHUNK && (v.ptr & 15u) == 0) ? block_crc_chunk64k<0>(vp, sZS8, sT, fast + 1024, sred)
                                                                               : block_crc_chunk<0>(vp, v.len, sT, sred);
            bad = crc != v.expect;
        }
        if (threadIdx.x == 0 && bad) {
            const uint32_t sg = v.seg & ~NZ_VER_FUSED;
            if (sg == NZ_VER_NOSEG) {
                a.res[v.req].status = NZ_ECORRUPT;
            } else {                                      // the slice's request and every coalesced follower that asked
                if (a.reqs_dev[v.req].flags & NZ_GET_VERIFY) a.res[v.req].status = NZ_ECORRUPT;
                for (uint32_t m = a.chain_head[sg]; m != 0xFFFFFFFFu; m = a.chain_next[m]) {
                    const uint32_t r = a.segs_dev[m].req;
                    if (a.reqs_dev[r].flags & NZ_GET_VERIFY) a.res[r].status = NZ_ECORRUPT;
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------
// launchers
// ---------------------------------------------------------------------------
cudaError_t nz_launch_lookup(const NzSlot* table, uint64_t line_mask, const NzDesc* descs, const uint64_t* keys,
                             uint32_t n, nz_extent* out, cudaStream_t s) {
    if (n == 0) return cudaSuccess;
    const uint32_t per_block = 64;
    uint32_t grid = (n + per_block - 1) / per_block;
    if (grid > 1184) grid = 1184;
    k_lookup<<<grid, 256, 0, s>>>(table, line_mask, descs, keys, n, out);
    return cudaGetLastError();
}
cudaError_t nz_launch_insert(NzSlot* table, uint64_t line_mask, uint64_t k0, uint64_t k1, uint32_t desc, uint64_t gen,
                             int32_t* status, cudaStream_t s) {
    k_insert<<<1, 32, 0, s>>>(table, line_mask, k0, k1, desc, gen, status);
    return cudaGetLastError();
}
cudaError_t nz_launch_publish(NzSlot* table, uint64_t line_mask, NzDesc* descs, const NzPubItem* items, uint32_t n,
                              int32_t* status, cudaStream_t s) {
    if (!n) return cudaSuccess;
    k_publish<<<(n + 127) / 128, 128, 0, s>>>(table, line_mask, descs, items, n, status);
    return cudaGetLastError();
}
cudaError_t nz_launch_erase(NzSlot* table, uint64_t line_mask, uint64_t k0, uint64_t k1, int32_t* status, cudaStream_t s) {
    k_erase<<<1, 32, 0, s>>>(table, line_mask, k0, k1, status);
    return cudaGetLastError();
}
cudaError_t nz_launch_plan(const NzPlanArgs& a, cudaStream_t s, const NzSmallWave* small, bool fused_copy) {
    // rq_src / rq_len live right after the inflate work array (see nz_api.cu: device block layout)
    uint64_t* rq_src = reinterpret_cast<uint64_t*>(a.infl + a.n);
    uint64_t* rq_len = rq_src + a.n;
    uint64_t* rq_tab = rq_src + 3 * (size_t)a.n;     // [rq_src | rq_len | infl_meta | rq_tab | rq_ssz]
    uint64_t* rq_ssz = rq_src + 4 * (size_t)a.n;
    if (small && fused_copy) {
        k_plan<true, true><<<1, 1024, 0, s>>>(a, rq_src, rq_len, rq_tab, rq_ssz, *small);
    } else if (small) {
        k_plan<true, false><<<1, 1024, 0, s>>>(a, rq_src, rq_len, rq_tab, rq_ssz, *small);
    } else {
        static const NzSmallWave none = {};
        k_plan<false, false><<<1, 1024, 0, s>>>(a, rq_src, rq_len, rq_tab, rq_ssz, none);
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    if (a.ver_cap) {
        k_plan_verify<<<1, 1024, 0, s>>>(a);
        e = cudaGetLastError();
    }
    return e;
}
// ---------------------------------------------------------------------------
// K2 with TMA staging (the default kernel).  A tile whose source, destinations and length are all
// 16 B aligned is moved by the bulk-copy engine instead of through registers:
//   cp.async.bulk  global -> shared   (one 32 KiB tile, completion on an mbarrier)
//   cp.async.bulk  shared -> global   once per destination of the coalesced group
// so a tile read once from HBM is written to every member with large contiguous bulk
// stores and almost no instructions.  STAGES shared buffers per CTA alternate: the load of
// tile i+1 is issued while the stores of tile i drain.  Tiles that do not qualify (byte
// ranges, tails) take the register path of copy_tile.  One elected thread drives the engine.
// ---------------------------------------------------------------------------
constexpr int NZ_TMA_THREADS = 128;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int N>
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" :: "n"(N) : "memory"); }

// Fused verify (VERIFY = true): the CRC-32C of every tile piece is taken from the copy that sits in
// shared memory (or, for the few short tiles of such a slice, from the source lines just read) and
// XOR-accumulated per 64 KiB chunk in the wave's verify entries; k_verify compares afterwards.
// HBM traffic stays read S + write P: the separate checksum pass over the source disappears.
struct VerifyCtx {
    const uint32_t* seg_ver;
    NzVerWork* ver;
    const uint32_t* gT;      // Castagnoli slice tables
    const uint32_t* gF;      // fused-verify powers (P16 | P2K)
    const uint32_t* gL;      // warp-tree level tables (Castagnoli)
    const uint32_t* sZS8;    // Z_4096 stride table, 8 copies, shared memory
    uint32_t* sred;          // 16 words
};

// data: the tile's bytes (generic address: shared or global), gsrc: their address in the arena
constexpr int NZ_TMA_VTHREADS = 256;      // the VERIFY instantiation runs more warps: the checksum is latency bound

__device__ __forceinline__ void tile_verify(const VerifyCtx& vc, const uint8_t* data, const uint8_t* gsrc, uint32_t n, uint32_t vbase, uint64_t c0ptr) {
    const uint32_t tid = threadIdx.x;
    const uint64_t q0 = (uint64_t)(uintptr_t)gsrc - c0ptr;
    const uint32_t j0 = (uint32_t)(q0 >> 16), j1 = (uint32_t)((q0 + n - 1) >> 16);
    for (uint32_t j = j0; j <= j1; ++j) {
        NzVerWork* e = vc.ver + vbase + j;
        const uint32_t eseg = *reinterpret_cast<volatile uint32_t*>(&e->seg);
        if (!(eseg & NZ_VER_FUSED)) continue;                  // chunk only partly inside the slice: k_verify reads it
        const uint32_t elen = *reinterpret_cast<volatile uint32_t*>(&e->len);
        const uint64_t cb = (uint64_t)j << 16;
        const uint64_t pb = q0 > cb ? q0 : cb;
        const uint64_t pe = (q0 + n) < (cb + NZ_CRC_CHUNK) ? (q0 + n) : (cb + NZ_CRC_CHUNK);
        const uint32_t plen = (uint32_t)(pe - pb), off = (uint32_t)(pb - q0);
        const uint32_t P = plen & ~15u, tail = plen & 15u;
        const bool ones = pb == cb;
        uint32_t m, pt;
        crc_piece_partial<NZ_TMA_VTHREADS>(data + off, P, ones, vc.sZS8, vc.gT, vc.gF, vc.gL, m, pt);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            m ^= __shfl_xor_sync(0xffffffffu, m, o);
            pt ^= __shfl_xor_sync(0xffffffffu, pt, o);
        }
        __syncthreads();                                       // sred reuse
        if ((tid & 31u) == 0) { vc.sred[tid >> 5] = m; vc.sred[8 + (tid >> 5)] = pt; }
        __syncthreads();                                       // also: every thread is done reading the tile
        if (tid >= 32 && tid < 64) {                           // warp 1 finishes the piece
            // every warp left the sum of its 512 B column block (warp tree); lane l moves warp l's to the end of the
            // full-row region with that warp's factor - one 32-step multiply for the whole CTA instead of one per thread
            const uint32_t l = tid - 32;
            uint32_t M = 0;
            if (l < NZ_TMA_VTHREADS / 32) M = gf_mul<0>(__ldg(vc.gF + (NZ_TMA_VTHREADS - 32 - 32 * l)), vc.sred[l]);
            M ^= __shfl_xor_sync(0xffffffffu, M, 4);
            M ^= __shfl_xor_sync(0xffffffffu, M, 2);
            M ^= __shfl_xor_sync(0xffffffffu, M, 1);
        if (l == 0) {
            uint32_t st = 0;
#pragma unroll
            for (int i = 0; i < NZ_TMA_VTHREADS / 32; ++i) st ^= vc.sred[8 + i];
            if (M) st ^= gf_mul<0>(__ldg(vc.gF + ((P >> 4) % NZ_TMA_VTHREADS)), M);   // full rows -> end of the piece
            if (tail) {
                if (ones && P == 0) st = 0xFFFFFFFFu;
                const uint8_t* tb = gsrc + off + P;
                for (uint32_t i = 0; i < tail; ++i) st = (st >> 8) ^ __ldg(vc.gT + ((st ^ tb[i]) & 0xffu));
            }
            const uint32_t after = elen - (uint32_t)(pe - cb);
            if (after) st = gf_mul<0>(crc_xpow8_small(after, vc.gF), st);
            atomicXor(&e->acc, st);
            atomicAdd(&e->got, plen);
        }
        }
    }
}

// L2 policy of the engine's stores (NZ_TMA_L2HINT): 0 = none, 1 = evict_first (the response bytes are not read again
// by this GPU: let them leave L2 ahead of everything else), 2 = evict_first stores + evict_last loads.
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ void bulk_store(void* dst, uint32_t smem, uint32_t bytes, uint64_t pol) {
    if (pol)
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;" :: "l"(dst), "r"(smem), "r"(bytes), "l"(pol) : "memory");
    else
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" :: "l"(dst), "r"(smem), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_load(uint32_t smem, const void* src, uint32_t bytes, uint32_t bar, uint64_t pol) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(bytes) : "memory");
    if (pol)
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
                     :: "r"(smem), "l"(src), "r"(bytes), "r"(bar), "l"(pol) : "memory");
    else
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     :: "r"(smem), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done = 0;
    while (!done) {
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                     : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    }
}

// l2h: 0 no hints, 1 evict_first stores, 2 evict_first stores + evict_last loads (a createpolicy result is never 0)
template <int STAGES, bool VERIFY>
__global__ void __launch_bounds__(VERIFY ? NZ_TMA_VTHREADS : NZ_TMA_THREADS)
k_gather_tma(const NzSegWork* __restrict__ work, const uint64_t* __restrict__ tile_begin, const uint32_t* __restrict__ chain_head,
             const uint32_t* __restrict__ chain_next, uint32_t nseg, uint32_t tile_bytes, const uint32_t* __restrict__ seg_ver,
             NzVerWork* ver, const uint32_t* __restrict__ crc_tables, uint32_t pace_ns, uint32_t spread, uint32_t l2h, uint32_t* dyn) {
    extern __shared__ __align__(128) uint8_t s_buf[];          // STAGES x tile_bytes [ | 32 KiB stride table when VERIFY ]
    __shared__ __align__(8) uint64_t s_bar[STAGES];
    __shared__ uint64_t s_step;
    __shared__ uint32_t s_red[16];
    constexpr int NT = VERIFY ? NZ_TMA_VTHREADS : NZ_TMA_THREADS;
    const uint32_t tid = threadIdx.x;
    VerifyCtx vc;
    uint32_t vbase = 0xFFFFFFFFu;
    uint64_t c0ptr = 0;
    if constexpr (VERIFY) {
        uint32_t* sz = reinterpret_cast<uint32_t*>(s_buf + (size_t)STAGES * tile_bytes);
        crc_load_fast_tables(sz, crc_tables + 16 * 256);       // Castagnoli Z_4096
        vc.seg_ver = seg_ver; vc.ver = ver; vc.gT = crc_tables; vc.gF = crc_tables + NZ_CRC_FUSE_OFF; vc.gL = crc_tables + 16 * 256 + NZ_CRC_LVL_OFF; vc.sZS8 = sz; vc.sred = s_red;
    }
    if (tid == 0) {
#pragma unroll
        for (int i = 0; i < STAGES; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(&s_bar[i])));
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    __syncthreads();
    const uint64_t pol_st = l2h >= 1 ? l2_policy_evict_first() : 0, pol_ld = l2h >= 2 ? l2_policy_evict_last() : 0;
    const TileOrder ord(tile_begin[nseg], spread);
    TileCursor c;
    DstChain dc;
    dc.work = work;
    dc.next = chain_next;
    dc.head = 0xFFFFFFFFu;
    uint32_t phase_bits = 0;                                   // bit s = parity to wait for on stage s
    uint32_t slot = 0;
    // pacing (waves that write pinned host memory): this CTA issues at most one tile per pace_ns, so the wave
    // offers the PCIe link a bounded rate and the posted-write queue ahead of latency-critical waves stays short
    bool pre_valid = false;                                    // VERIFY look-ahead: load of `pre_tile` already in flight
    uint64_t pre_tile = 0;
    uint64_t pace_next = 0;
    if (pace_ns && tid == 0) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(pace_next));
    StepIter it;
    for (it.begin(dyn, &s_step); it.i < ord.steps(); it.next()) {
        const uint64_t tile = ord.tile(it.i);
        if (tile == ~0ull) continue;
        const uint32_t before = c.seg;
        const bool fresh = !(tile >= c.lo_t && tile < c.hi_t);
        c.seek(tile, work, tile_begin, nseg);
        if (fresh || before != c.seg) {
            dc.head = __ldg(chain_head + c.seg);
            if constexpr (VERIFY) {
                vbase = __ldg(seg_ver + c.seg);
                if (vbase != 0xFFFFFFFFu) c0ptr = *reinterpret_cast<volatile uint64_t*>(&ver[vbase].ptr);
            }
        }
        int64_t lo, hi;
        c.span(tile, tile_bytes, lo, hi);
        const uint8_t* s = reinterpret_cast<const uint8_t*>(c.src) + lo;
        uint8_t* d = reinterpret_cast<uint8_t*>(c.dst) + lo;
        const uint32_t n = (uint32_t)(hi - lo);
        const uint32_t bulk = n & ~15u;                        // bytes the engine moves; the (<16 B) tail goes by hand
        const bool engine = bulk >= 2048 && ((((uintptr_t)s) | ((uintptr_t)d)) & 15u) == 0;
        if (!engine) {
            dc.d0 = d;
            dc.lo = lo;
            copy_tile<NT, 8>(s, dc, n);
            if constexpr (VERIFY) {
                if (vbase != 0xFFFFFFFFu) tile_verify(vc, s, s, n, vbase, c0ptr);   // short tile of an engine slice: lines are L2-hot
            }
            continue;
        }
        if (pace_ns && tid == 0) {
            uint64_t now;
            for (;;) {
                asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
                if (now >= pace_next) break;
                __nanosleep(100);
            }
            pace_next = (now - pace_next > 8ull * pace_ns ? now : pace_next) + pace_ns;     // no burst after a stall
        }
        const uint32_t buf = smem_u32(s_buf + (size_t)slot * tile_bytes), bar = smem_u32(&s_bar[slot]);
        if constexpr (VERIFY) {
            // every thread follows the barrier phases: load -> (all wait) -> stores -> checksum from shared memory
            if (tid == 0 && !(pre_valid && pre_tile == tile)) {   // not already in flight from the previous iteration's look-ahead
                bulk_wait_read<STAGES - 1>();
                bulk_load(buf, s, bulk, bar, pol_ld);
            }
            pre_valid = false;
            const bool vt = vbase != 0xFFFFFFFFu;
            if (vt || tid == 0) mbar_wait(bar, (phase_bits >> slot) & 1u);
            phase_bits ^= 1u << slot;
            if (tid == 0) {
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                bulk_store(d, buf, bulk, pol_st);
                for (uint32_t m = dc.head; m != 0xFFFFFFFFu; m = __ldg(chain_next + m))
                    bulk_store(reinterpret_cast<uint8_t*>(__ldg(&work[m].dst)) + lo, buf, bulk, pol_st);
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
            // look-ahead: the checksum below takes longer than a tile's load, so the next tile of this CTA is requested
            // first (into the other stage; its previous stores are done once at most this tile's group is pending)
            // (with dynamic steps only thread 0 knows the next one - and only thread 0 acts on it)
            const uint64_t step2 = it.peek();
            const uint64_t tile2 = (tid == 0 || !dyn) && step2 < ord.steps() ? ord.tile(step2) : ~0ull;
            if (vt && tile2 != ~0ull) {
                TileCursor c2 = c;
                c2.seek(tile2, work, tile_begin, nseg);
                int64_t lo2, hi2;
                c2.span(tile2, tile_bytes, lo2, hi2);
                const uint8_t* s2 = reinterpret_cast<const uint8_t*>(c2.src) + lo2;
                const uint8_t* d2 = reinterpret_cast<const uint8_t*>(c2.dst) + lo2;
                const uint32_t bulk2 = (uint32_t)(hi2 - lo2) & ~15u;
                if (bulk2 >= 2048 && ((((uintptr_t)s2) | ((uintptr_t)d2)) & 15u) == 0) {
                    const uint32_t slot2 = slot + 1 == STAGES ? 0 : slot + 1;
                    if (tid == 0) {
                        bulk_wait_read<STAGES - 1>();
                        bulk_load(smem_u32(s_buf + (size_t)slot2 * tile_bytes), s2, bulk2, smem_u32(&s_bar[slot2]), pol_ld);
                    }
                    pre_valid = true;
                    pre_tile = tile2;
                }
            }
            if (vt) tile_verify(vc, s_buf + (size_t)slot * tile_bytes, s, n, vbase, c0ptr);
        } else if (tid == 0) {
            // one elected thread drives the engine.  Stage `slot` was last read by the stores of the engine tile STAGES
            // back: at most STAGES-1 younger store groups may still be reading shared memory.  The other CTAs resident
            // on the SM cover this tile's load latency (a look-ahead inside the CTA was measured slower: it costs a stage).
            bulk_wait_read<STAGES - 1>();
            bulk_load(buf, s, bulk, bar, pol_ld);
            mbar_wait(bar, (phase_bits >> slot) & 1u);
            phase_bits ^= 1u << slot;
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            bulk_store(d, buf, bulk, pol_st);
            for (uint32_t m = dc.head; m != 0xFFFFFFFFu; m = __ldg(chain_next + m))
                bulk_store(reinterpret_cast<uint8_t*>(__ldg(&work[m].dst)) + lo, buf, bulk, pol_st);
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
        slot = slot + 1 == STAGES ? 0 : slot + 1;
        // tail bytes of a slice's last tile (every member)
        const uint32_t tail = n - bulk;
        if (tid < tail) {
            const uint8_t b8 = s[bulk + tid];
            d[bulk + tid] = b8;
            for (uint32_t m = dc.head; m != 0xFFFFFFFFu; m = __ldg(chain_next + m))
                (reinterpret_cast<uint8_t*>(__ldg(&work[m].dst)) + lo)[bulk + tid] = b8;
        }
    }
    if (tid == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

struct TmaCfg { int stages; uint32_t tile; int l2hint; bool stages_forced; };
static TmaCfg tma_cfg() {
    static TmaCfg c = [] {
        TmaCfg t{2, 32768, 2, false};      // L2 policy: evict_first stores + evict_last loads (+0.4 % coalesced, +1.3 % uncoalesced: profiles/r02_sweep_tma_l2hint.txt)
        if (const char* e = getenv("NZ_TMA_STAGES")) { const int v = atoi(e); if (v >= 2 && v <= 4) { t.stages = v; t.stages_forced = true; } }
        if (const char* e = getenv("NZ_TMA_TILE_KB")) { const int v = atoi(e); if (v == 16 || v == 32 || v == 64) t.tile = (uint32_t)v << 10; }
        if (const char* e = getenv("NZ_TMA_L2HINT")) { const int v = atoi(e); if (v >= 0 && v <= 2) t.l2hint = v; }
        while ((size_t)t.stages * t.tile > (200u << 10)) --t.stages;     // stay inside one SM's shared memory
        return t;
    }();
    return c;
}
struct TmaArgs {
    const NzSegWork* work; const uint64_t* tile_begin; const uint32_t* chain_head; const uint32_t* chain_next;
    uint32_t nseg; int grid; cudaStream_t s; int* occ;
    const uint32_t* seg_ver; NzVerWork* ver; const uint32_t* crc_tables;
    uint32_t pace_ns, spread;
    int stages_hint;         // 2 or 3 (0 = no opinion): see nz_launch_gather
    uint32_t* dyn;           // zeroed step counter for dynamic tile scheduling, or nullptr
};
template <int STAGES, bool VERIFY>
static cudaError_t launch_tma(const TmaArgs& a) {
    const TmaCfg t = tma_cfg();
    const size_t smem = (size_t)STAGES * t.tile + (VERIFY ? 32768 : 0);
    constexpr int NTH = VERIFY ? NZ_TMA_VTHREADS : NZ_TMA_THREADS;
    static bool attr_set[64] = {};          // function attributes are per device
    static int max_grid[64] = {};           // VERIFY: resident CTAs (fewer fit per SM than the plain kernel's persistent grid)
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64 || !attr_set[dev]) {
        cudaFuncSetAttribute(k_gather_tma<STAGES, VERIFY>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (dev >= 0 && dev < 64) {
            attr_set[dev] = true;
            int nb = 0, sms = 0;
            cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
            if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_gather_tma<STAGES, VERIFY>, NTH, smem) != cudaSuccess || nb < 1) nb = 1;
            max_grid[dev] = nb * (sms > 0 ? sms : 1);
        }
    }
    if (a.occ) {
        int nb = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_gather_tma<STAGES, VERIFY>, NTH, smem) != cudaSuccess || nb < 1) nb = 1;
        *a.occ = nb;
        return cudaSuccess;
    }
    int grid = a.grid;
    if (VERIFY && dev >= 0 && dev < 64 && max_grid[dev] > 0 && grid > max_grid[dev]) grid = max_grid[dev];
    k_gather_tma<STAGES, VERIFY><<<grid, NTH, smem, a.s>>>(a.work, a.tile_begin, a.chain_head, a.chain_next, a.nseg, t.tile,
                                                                            a.seg_ver, a.ver, a.crc_tables, a.pace_ns, a.spread, (uint32_t)t.l2hint, a.dyn);
    return cudaGetLastError();
}
static cudaError_t dispatch_tma(const TmaArgs& a) {
    TmaCfg t = tma_cfg();
    // three stages win when every tile is read once and written once (0.98 -> 1.00 of the copy peak), two when tiles fan
    // out to coalesced followers (a third buffer costs a resident CTA): the host says which kind of wave this is
    // (the fused-verify instantiation keeps two: its 32 KiB stride table already costs it a resident CTA)
    if (!t.stages_forced && !a.seg_ver && a.stages_hint >= 2 && a.stages_hint <= 3 && (size_t)a.stages_hint * t.tile <= (200u << 10)) t.stages = a.stages_hint;
    if (a.seg_ver) {                          // fused verify
        switch (t.stages) {
        case 3: return launch_tma<3, true>(a);
        case 4: return launch_tma<4, true>(a);
        default: return launch_tma<2, true>(a);
        }
    }
    switch (t.stages) {
    case 3: return launch_tma<3, false>(a);
    case 4: return launch_tma<4, false>(a);
    default: return launch_tma<2, false>(a);
    }
}

// Two gather kernels ship: NZ_GATHER_TMA (default; bulk-copy engine staging, register path for byte-phase tiles) and
// NZ_GATHER_REG (register path only: 256 threads x 8 vectors in flight, 2 CTAs/SM, 32 KiB tiles) as fallback and A/B partner.
uint32_t nz_gather_tile_bytes(int v) { return v == NZ_GATHER_TMA ? tma_cfg().tile : 256u * 8u * 16u; }
bool nz_gather_fuses_verify(int variant) { return variant == NZ_GATHER_TMA; }
cudaError_t nz_launch_gather(int variant, const NzSegWork* work, const uint64_t* tile_begin, const uint32_t* chain_head,
                             const uint32_t* chain_next, uint32_t nseg, int grid, cudaStream_t s,
                             const uint32_t* seg_ver, NzVerWork* ver, const uint32_t* crc_tables, uint32_t pace_ns, uint32_t spread,
                             int stages_hint, uint32_t* dyn) {
    if (variant == NZ_GATHER_TMA)
        return dispatch_tma(TmaArgs{work, tile_begin, chain_head, chain_next, nseg, grid, s, nullptr, seg_ver, ver, crc_tables, pace_ns, spread, stages_hint, dyn});
    k_gather<256, 8, 2><<<grid, 256, 0, s>>>(work, tile_begin, chain_head, chain_next, nseg, 256u * 8u * 16u, spread, dyn);
    return cudaGetLastError();
}
int nz_gather_blocks_per_sm(int variant) {
    int nb = 1;
    if (variant == NZ_GATHER_TMA) {
        dispatch_tma(TmaArgs{nullptr, nullptr, nullptr, nullptr, 0, 0, nullptr, &nb, nullptr, nullptr, nullptr, 0, 0, 0, nullptr});
        return nb;
    }
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_gather<256, 8, 2>, 256, 0) != cudaSuccess || nb < 1) nb = 2;
    return nb;
}
cudaError_t nz_launch_verify(const NzPlanArgs& a, const uint32_t* crc_tables, int grid, cudaStream_t s) {
    k_verify<<<grid, 256, 0, s>>>(a, crc_tables);
    return cudaGetLastError();
}
cudaError_t nz_launch_crc_chunks(const uint8_t* p, uint64_t n, uint32_t* chunk_out, uint32_t* whole_out,
                                 const uint32_t* crc_tables, int sm_count, cudaStream_t s) {
    const uint64_t nchunks = (n + NZ_CRC_CHUNK - 1) / NZ_CRC_CHUNK;
    if (nchunks) {
        uint64_t grid = nchunks;
        // one full wave of resident CTAs, each striding over its share of the chunks: a grid of 8 per SM with 6 resident
        // (shared memory) ran a second, one-third-occupied wave that cost about a third of the kernel's time
        static int per_sm = 0;
        static const bool tree = [] { const char* e = getenv("NZ_CRC_TREE"); return !e || atoi(e) != 0; }();
        if (!per_sm && (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_crc_chunks<true>, 256, 0) != cudaSuccess || per_sm < 1)) per_sm = 4;
        const uint64_t cap = (uint64_t)sm_count * per_sm;
        if (grid > cap) grid = cap;
        if (tree) k_crc_chunks<true><<<(unsigned)grid, 256, 0, s>>>(p, n, chunk_out, crc_tables);
        else k_crc_chunks<false><<<(unsigned)grid, 256, 0, s>>>(p, n, chunk_out, crc_tables);
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return e;
    }
    k_crc_fold<<<1, 1024, 0, s>>>(chunk_out, n, whole_out);
    return cudaGetLastError();
}

cudaError_t nz_launch_crc_jobs(const NzCrcJob* jobs, uint32_t n_chunks, const NzFoldJob* folds, uint32_t n_objects, const uint32_t* crc_tables,
                               int sm_count, cudaStream_t s) {
    if (n_chunks) {
        k_crc_jobs<<<(unsigned)std::min<uint64_t>(n_chunks, (uint64_t)sm_count * 4), 256, 0, s>>>(jobs, n_chunks, crc_tables);
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return e;
    }
    if (n_objects) k_crc_fold_jobs<<<(n_objects + 127) / 128, 128, 0, s>>>(folds, n_objects);
    return cudaGetLastError();
}

cudaError_t nz_upload_constants(uint32_t* crc_tables_dev) {
    cudaError_t e = nz_crc_upload_constants();
    if (e != cudaSuccess) return e;
    e = nz_inflate_upload_constants();
    if (e != cudaSuccess) return e;
    // slice tables: [0..7] Castagnoli, [8..15] IEEE (gzip trailer CRC-32)
    static uint32_t t[NZ_CRC_TABLE_WORDS];
    static_assert(NZ_CRC_FUSE_OFF == 16 * 256 + 2 * NZ_CRC_FAST_WORDS, "table block layout");
    nz_crc_fill_fast_tables(t + 16 * 256);
    nz_crc_fill_fuse_tables(t + NZ_CRC_FUSE_OFF);
    const uint32_t polys[2] = {NZ_POLY_C, NZ_POLY_IEEE};
    for (int pi = 0; pi < 2; ++pi) {
        uint32_t* T = t + pi * 8 * 256;
        for (uint32_t i = 0; i < 256; ++i) {
            uint32_t c = i;
            for (int k = 0; k < 8; ++k) c = (c & 1) ? (c >> 1) ^ polys[pi] : c >> 1;
            T[i] = c;
        }
        for (int s = 1; s < 8; ++s)
            for (uint32_t i = 0; i < 256; ++i) T[s * 256 + i] = (T[(s - 1) * 256 + i] >> 8) ^ T[T[(s - 1) * 256 + i] & 0xff];
    }
    return cudaMemcpy(crc_tables_dev, t, sizeof t, cudaMemcpyHostToDevice);
}
