// nezha_b200/csrc/nz_internal.h — device data layout and host context of
// libnezha_gpu.so.  Names follow the reference's domain (cache zone, extents,
// keys — /root/reference/examples/demo/nginx.conf:38,56), not ML vocabulary.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <condition_variable>
#include <map>
#include <mutex>
#include <set>
#include <unordered_map>
#include <unordered_set>
#include <vector>

#include "../../include/nezha_gpu.h"

// ---------------------------------------------------------------------------
// HBM layout
// ---------------------------------------------------------------------------
// Hash table ("zone"): open addressing, linear probing by 128-byte lines.
// One slot = 32 B, four slots per line, so one probe step is exactly one
// coalesced 128 B transaction read by a 4-lane group (warp-ballot match).
struct __align__(32) NzSlot {
    uint64_t k0, k1;    // key128 (MD5 of "$host/$proxy_host$uri")
    uint32_t state;     // NZ_SLOT_*
    uint32_t desc;      // index into the descriptor array
    uint64_t gen;       // insertion generation (diagnostic / ABA guard)
};
enum : uint32_t { NZ_SLOT_EMPTY = 0, NZ_SLOT_TOMB = 1, NZ_SLOT_LIVE = 2, NZ_SLOT_BUSY = 3 };

// Extent descriptor: where an object's bytes live in the arena.
struct __align__(64) NzDesc {
    uint64_t arena_off;      // byte offset of the extent in the arena (4 KiB aligned)
    uint64_t stored_len;     // bytes at rest
    uint64_t raw_len;        // inflated length (== stored_len for identity)
    uint64_t user_meta;      // host cookie
    uint64_t crc_tab_off;    // arena offset of the u32 chunk-CRC table
    uint32_t flags;          // NZ_ENC_*
    uint32_t crc32c;         // whole-object CRC-32C of the stored bytes
    uint32_t n_chunks;
    uint32_t gen;
    uint64_t stripe_bytes;   // striped objects: stripe size; arena_off then points at the NzDevStripe map, n_chunks = stripes
                             // gzip objects (never striped): arena offset of the checkpoint index (NzIdxHeader), 0 = none
};
static_assert(sizeof(NzSlot) == 32, "slot");
static_assert(sizeof(NzDesc) == 64, "desc");

// One GET as the device sees it (lives in mapped pinned memory, read once).
struct __align__(16) NzDevReq {
    uint64_t k0, k1;
    uint64_t dst;          // device-visible address of the destination buffer + dst_base
    uint64_t dst_cap;      // bytes available from dst
    uint32_t flags;        // NZ_GET_*
    uint32_t seg_begin;    // first entry in the segment array
    uint32_t nseg;         // >= 1 (nseg==0 requests get one whole-object pseudo segment)
    uint32_t whole;        // 1 = pseudo segment: serve [0, length)
    uint64_t scratch;      // device address of the inflate scratch (0 = inflate straight to dst)
    uint64_t scratch_cap;
    uint32_t infl_leader;  // inflate dedupe: index of the request in this wave that inflates this key (== own index for a
                           // leader); followers are served by a copy from the leader's scratch
    uint32_t pad0;
    uint64_t pad1;
};
static_assert(sizeof(NzDevReq) == 80, "req layout");
struct __align__(8) NzDevSeg {
    uint64_t src_off, len, dst_off;
    uint32_t req;
    uint32_t pad;
};
// Descriptors of a small wave travel as kernel parameters (host -> device direction), so the plan
// kernel does not start with a PCIe read that queues behind bulk waves' posted writes.
constexpr uint32_t NZ_SMALL_WAVE = 4;       // 4 x (64 + 32) B of kernel parameters (32 entries measured +1.3 us on every single-GET wave)
struct NzSmallWave {
    NzDevReq reqs[NZ_SMALL_WAVE];
    NzDevSeg segs[NZ_SMALL_WAVE];
};
// Resolved copy work for the gather kernel.
struct __align__(16) NzSegWork {
    uint64_t src;          // device address of first byte
    uint64_t dst;          // device-visible address of first byte
    uint64_t len;
    int64_t  vstart;       // = -(dst & 127): tile k covers [vstart + k*TILE, +TILE) ∩ [0,len)
};
// Per-request inflate work
struct __align__(16) NzInflWork {
    uint64_t src;          // gzip member
    uint64_t src_len;
    uint64_t dst;          // raw output
    uint64_t dst_cap;
    uint32_t req;
    uint32_t pad;
    uint64_t idx;          // device address of the member's checkpoint index (NzIdxHeader), 0 = none: legacy whole-member decode
};
static_assert(sizeof(NzInflWork) == 48, "inflate work layout");

// ---------------------------------------------------------------------------
// Checkpoint index of a gzip member (built once at nz_put_commit by k_inflate_index, kept in the arena next to the
// member).  DEFLATE is bit-serial, so a GET gets its parallelism from the index:
//   member -> SEGMENTS  (~128 KiB of output, one warp each, independent of each other: the bytes a segment's matches
//                        copy from before its own start are kept verbatim in the index = "side bytes", in stream order)
//          -> GROUPS    (<= 32 sub-segments of one deflate block = one warp iteration; tables are shared)
//          -> SUB-SEGMENTS (<= NZ_IDX_T = 32 tokens, one lane each: bit length and output length as u16 deltas)
// Layout of the blob: NzIdxHeader | NzIdxSeg[n_seg] | NzIdxGrp[n_grp] | u32 delta[n_grp][32] | side bytes.
// ---------------------------------------------------------------------------
constexpr uint32_t NZ_IDX_MAGIC = 0x58495A4Eu;      // "NZIX"
constexpr uint32_t NZ_IDX_T = 32;                   // tokens (literal or match) per sub-segment
constexpr uint32_t NZ_IDX_SUBS = 32;                // sub-segments per group
constexpr uint32_t NZ_IDX_GRP_STORED = 1u;          // group = one stored block (type 0): tok_bit = 8 * byte offset of its data, hdr_bit = its length
struct __align__(16) NzIdxHeader {
    uint32_t magic, version;
    uint32_t n_seg, n_grp;
    uint64_t raw_len;
    uint64_t side_bytes;
    uint64_t off_seg, off_grp, off_delta, off_side;  // byte offsets from the header
};
struct __align__(16) NzIdxSeg {
    uint64_t out_start, out_len;     // the segment's slice of the inflated output
    uint64_t side_off;               // its side bytes start here (consumed in stream order)
    uint32_t first_grp, n_grp;
};
struct __align__(16) NzIdxGrp {
    uint64_t tok_bit;                // bit offset (from the member's first byte) of the group's first token
    uint64_t out_off;                // output offset of the group's first byte
    uint64_t hdr_bit;                // bit offset of the 3-bit header of the deflate block the group belongs to
    uint32_t n_sub;                  // 1..32 (0 for a stored block)
    uint32_t flags;                  // NZ_IDX_GRP_*
};
static_assert(sizeof(NzIdxHeader) == 64 && sizeof(NzIdxSeg) == 32 && sizeof(NzIdxGrp) == 32, "index layout");

// one member to index (k_inflate_index: one warp per job) and what came out
struct __align__(16) NzIdxJob {
    uint64_t src, src_len;           // the stored member
    uint64_t out, out_cap;           // temporary inflated output (cap = the trailer's ISIZE)
    uint64_t seg, grp, delta, side;  // temporary index arrays
    uint64_t side_cap;
    uint32_t seg_cap, grp_cap;
    uint32_t seg_bytes, pad;
    uint64_t dest;                   // filled by the host before k_index_pack: device address of the final blob (0 = none)
};
struct __align__(16) NzIdxResult {
    int32_t  status;                 // 0 = member valid and indexed, 1 = member valid but the index outgrew its caps, -1 = malformed
    uint32_t n_seg, n_grp;
    uint32_t crc32;                  // trailer CRC-32 (checked by k_inflate_crc against the temporary output)
    uint64_t raw_len, side_bytes;
};
// batched publish (K6): descriptor + key of one object
struct NzPubItem {
    uint64_t k0, k1, gen;
    uint32_t desc_idx, pad;
    NzDesc d;
};

struct __align__(8) NzDevRes {   // mirrors nz_get_res
    int32_t  status;
    uint32_t flags;
    uint64_t bytes;
    uint64_t length;
    uint32_t crc32c;
    uint32_t reserved;
};
static_assert(sizeof(NzDevRes) == sizeof(nz_get_res), "res layout");

// Verify work: one 64 KiB chunk of a stored extent whose CRC-32C must equal the table entry.
// Two ways to get there.  Fused: the gather kernel checksums the tile while it sits in shared
// memory and XORs the piece's contribution (shifted to the chunk end) into `acc`, counting bytes
// in `got`; k_verify then only compares.  Separate: k_verify reads the chunk itself (chunks a
// slice covers only partly, slices the bulk-copy engine cannot take).
// device CRC table block (c->crc_tables): slice tables 16 x 256 | fast-path tables 2 x 6400 | fused-verify powers 288
constexpr int NZ_CRC_FAST_WORDS_DECL = 4 * 256 + 256 + 5 * 4 * 256;   // per polynomial: stride table | row factors | warp-tree levels (nz_crc.cuh)
constexpr int NZ_CRC_FUSE_OFF = 16 * 256 + 2 * NZ_CRC_FAST_WORDS_DECL;
constexpr int NZ_CRC_FUSE_WORDS = 256 + 32;
constexpr int NZ_CRC_TABLE_WORDS = NZ_CRC_FUSE_OFF + NZ_CRC_FUSE_WORDS;
constexpr uint32_t NZ_VER_FUSED = 0x80000000u;   // in seg
constexpr uint32_t NZ_VER_NOSEG = 0x7FFFFFFFu;   // seg & ~FUSED: no copy slice behind this entry (inflate input)
// batched commit of small objects: one chunk to checksum / one object to fold
struct NzCrcJob { uint64_t ptr; uint32_t len; uint32_t pad; uint64_t out; };     // out: device address of the chunk-table entry
struct NzFoldJob { uint64_t tab; uint64_t len; uint64_t whole_out; };              // tab: the object's chunk table; whole_out: device address of a u32
struct __align__(16) NzVerWork {
    uint64_t ptr;
    uint32_t len;
    uint32_t expect;
    uint32_t req;
    uint32_t seg;          // leader slice (work index) | NZ_VER_FUSED
    uint32_t acc;          // fused: XOR of piece remainders (raw register state, before the final inversion)
    uint32_t got;          // fused: bytes accumulated
};
static_assert(sizeof(NzVerWork) == 32, "ver layout");


// ---------------------------------------------------------------------------
// host context
// ---------------------------------------------------------------------------
struct nz_buf {
    nz_ctx* ctx;
    void* host;        // pinned host address or nullptr
    void* dev;         // device-visible address
    uint64_t size;
    int32_t kind;
    bool owned;
};

struct __align__(8) NzDevStripe {   // mirrors nz_stripe
    uint32_t shard, reserved;
    uint64_t arena_off, len, crc_tab_off;
};
static_assert(sizeof(NzDevStripe) == sizeof(nz_stripe), "stripe layout");

struct NzHostObj {     // host shadow of a live object (arena bookkeeping, EEXIST, inflate sizing)
    uint64_t stripe_bytes = 0;   // != 0: striped object, arena extent holds the stripe map
    uint32_t crc_whole = 0;
    uint64_t crc_tab_off = 0;
    uint32_t desc;
    uint64_t arena_off, arena_len;
    uint64_t stored_len, raw_len;
    uint32_t flags;
    uint64_t idx_off = 0, idx_len = 0;   // gzip objects: extent of the checkpoint index (0 = none)
};

struct KeyHash {
    size_t operator()(const std::pair<uint64_t, uint64_t>& k) const { return (size_t)(k.second ^ (k.first * 0x9E3779B97F4A7C15ull)); }
};

struct nz_ticket {
    nz_ctx* ctx;
    uint32_t n = 0, nseg = 0;
    // pinned+mapped block: reqs | segs | res
    void* pin = nullptr; void* pin_dev = nullptr; size_t pin_cap = 0;
    // device block: segwork | tile prefix | verify work | inflate work | counters
    void* dev = nullptr; size_t dev_cap = 0;
    void* scratch = nullptr; size_t scratch_cap = 0;   // inflate scratch (grow-only, reused across waves)
    cudaStream_t stream = nullptr;
    bool profiled = false;             // kernel-level events recorded for this wave
    cudaEvent_t ev_start = nullptr, ev_done = nullptr, ev_g0 = nullptr, ev_g1 = nullptr, ev_i0 = nullptr, ev_i1 = nullptr;
    bool ran_inflate = false;
    bool own_stream = false;
    uint64_t req_bytes_hint = 0;
    uint64_t seq = 0;                  // submit order (ctx->inflight)
    bool flag_wait = false;            // fused small wave: completion is the pinned flag ver_stat[2], no event query needed
    bool paced = false;                // counted in ctx->paced_inflight
    uint32_t* ver_stat = nullptr;      // pinned: {verify chunks emitted, of which checksummed inside the gather}
};

struct nz_put {
    nz_ctx* ctx;
    uint64_t k0, k1;
    uint64_t stored_len, raw_len, user_meta;
    uint32_t flags;
    uint64_t arena_off, arena_len;     // extent + chunk-crc table
    uint64_t crc_tab_off;
    uint32_t n_chunks;
    cudaStream_t stream;
};

constexpr int NZ_GATHER_REG = 0, NZ_GATHER_TMA = 1;     // the two K2 kernels (nz_kernels.cu)

struct nz_ctx {
    nz_config cfg;
    int device;
    int sm_count;
    int numa_node = -1;                 // host memory node next to this GPU (sysfs), -1 unknown / single node
    bool numa_place = false;            // pinned buffers of this context are placed on numa_node (NZ_NUMA=0 turns it off)
    int gather_grid;       // persistent grid size for k_gather
    int gather_grid3 = 0;  // ... when the bulk-copy kernel runs with three stages (waves without repeated keys)
    bool profile = false;   // record per-kernel CUDA events in every wave (nz_set_profiling / env NZ_PROFILE=1)
    bool fuse_small = true;  // env NZ_FUSE_SMALL=0: small waves launch plan and gather separately
    bool fuse_verify = true; // env NZ_FUSE_VERIFY=0: NZ_GET_VERIFY re-reads every chunk in k_verify (A/B measurements)
    bool coalesce = true;   // env NZ_COALESCE=0 turns request coalescing off (A/B measurements)
    bool span_dynamic = true;  // K5: such waves hand out tiles dynamically (env NZ_SPAN_DYN=0: static split)
    uint32_t span_spread = 256; // K5: rows of the spread tile order of waves with striped sources (env NZ_SPAN_SPREAD, 0 = plain order)
    int gather_variant = NZ_GATHER_TMA; // K2 kernel: bulk-copy engine staging (default) or the register path (env NZ_GATHER_VARIANT=0|reg)
    int inflate_grid;
    int inflate_seg_grid;
    uint8_t* arena = nullptr;
    uint64_t arena_bytes = 0;
    NzSlot* table = nullptr;
    uint64_t n_lines = 0;           // table lines (4 slots each), power of two
    NzDesc* descs = nullptr;
    uint32_t* crc_tables = nullptr; // device: 8 x 256 u32 slice tables (Castagnoli), then 8 x 256 (IEEE)
    std::mutex mu;                  // arena + shadow + desc free list
    std::map<uint64_t, uint64_t> free_by_off;   // arena free list: off -> len (coalescing)
    std::set<std::pair<uint64_t, uint64_t>> free_by_len;   // the same extents as (len, off): best fit in O(log n)
    std::vector<uint32_t> free_desc;
    std::unordered_map<std::pair<uint64_t, uint64_t>, NzHostObj, KeyHash> shadow;
    uint64_t gen = 1;
    // stream pool
    std::vector<cudaStream_t> streams;
    std::atomic<uint32_t> rr{0};
    cudaStream_t put_stream = nullptr;
    std::vector<cudaStream_t> hi_streams;   // highest priority: NZ_GET_URGENT waves
    std::atomic<double> host_rate_gbs{0};   // > 0: pace such waves to this many GB/s in total (nz_set_host_rate / env NZ_HOST_RATE_GBS; 0 = unpaced)
    // staged D2H: large waves to pinned host memory gather into an HBM staging slot at HBM speed, then the copy engine
    // moves every slice out (256 B TLPs: ~57 GB/s where SM stores reach ~52).  Ring of slots, reuse ordered by events.
    struct StageSlot { void* buf = nullptr; size_t cap = 0; cudaEvent_t free_ev = nullptr; bool used = false; };
    StageSlot stage[4];
    uint32_t stage_next = 0;
    std::mutex stage_mu;
    bool host_dma = true;                   // env NZ_HOST_DMA=0: always store from the gather kernel
    std::atomic<uint64_t> dma_waves{0};
    std::atomic<int> paced_inflight{0};     // paced waves submitted and not yet seen complete
    std::atomic<int> host_grid{32};                     // CTAs for waves that only write pinned host memory (PCIe bound: leave the SMs free)
    // ticket recycling + the waves in flight (nz_evict frees an extent only after every wave submitted before the
    // erase has finished: such a wave may have probed the slot while it was still live)
    std::mutex tmu;
    std::vector<nz_ticket*> ticket_pool;
    uint64_t submit_seq = 0;
    std::map<uint64_t, nz_ticket*> inflight;
    uint32_t idx_seg_bytes = 128u << 10;    // env NZ_INFL_SEG_KB: output bytes per index segment (one warp each at GET time)
    bool infl_dedupe = true;                // env NZ_INFL_DEDUPE=0: every inflate request decodes its member itself (A/B measurements)
    bool idx_build = true;                  // env NZ_INFL_INDEX=0: commit builds no index (every inflate GET decodes the whole member in one warp)
    // stats
    std::atomic<uint64_t> gets{0}, get_bytes{0}, hits{0}, misses{0}, puts{0}, put_bytes{0}, evictions{0}, corrupt{0}, launches{0}, file_cufile_bytes{0}, file_bounce_bytes{0}, verify_chunks{0}, verify_fused{0}, infl_segs{0}, infl_whole{0}, index_bytes{0};
    // bounce lanes for nz_put_file: one pinned allocation cut into lanes of 2 x 8 MiB; a call takes as many free
    // lanes as its length warrants (at least one, waiting if none is free), so files load concurrently
    void* bounce[2] = {nullptr, nullptr};
    size_t bounce_bytes = 0;
    cudaEvent_t bounce_ev[2] = {nullptr, nullptr};
    std::vector<char> lane_busy;
    std::condition_variable lane_cv;
    std::mutex put_mu;              // serialises fill-side device work
    uint64_t* peer_base_dev = nullptr;      // device array [shard_count]
    std::vector<void*> peer_ptrs;           // host copy; nullptr = not attached
    std::atomic<uint32_t> n_striped{0};     // live striped objects (GET submit consults the shadow only if > 0)
};

// ---------------------------------------------------------------------------
// kernel launchers (nz_kernels.cu / nz_inflate.cu)
// ---------------------------------------------------------------------------
struct NzPlanArgs {
    const NzSlot* table; uint64_t line_mask;
    const NzDesc* descs;
    const uint8_t* arena;
    const NzDevReq* reqs_host; uint32_t n;      // mapped pinned (written by the host at submit)
    const NzDevSeg* segs_host; uint32_t nseg;
    NzDevReq* reqs_dev;                         // staged copies in HBM (filled by k_plan)
    NzDevSeg* segs_dev;
    int32_t* rq_stat;                           // n: per-request status as of phase A
    NzDevRes* res;
    NzSegWork* work;            // nseg entries
    uint64_t* tile_begin;       // nseg + 1
    NzVerWork* ver;             // capacity ver_cap
    uint32_t ver_cap;
    NzInflWork* infl;           // capacity n
    uint32_t* counters;         // [0]=n_ver [1]=n_infl [2]=inflate work-steal cursor ... [6]=gather step counter (dynamic tile scheduling)
    uint64_t* infl_meta;        // n: (produced length << 32) | trailer CRC-32, written by k_inflate
    uint32_t tile_bytes;        // gather tile size of the selected k_gather variant
    uint32_t coalesce;          // 1 = chain identical slices behind one leader (read once, write many)
    uint32_t* chain_head;       // nseg: first follower of a leader slice, 0xFFFFFFFF = none
    uint32_t* chain_next;       // nseg: next follower in the chain
    const uint64_t* peer_base;  // device array [shard_count]: base address of every shard's arena (own + IPC-mapped peers)
    uint32_t* rq_desc;          // n: descriptor index of every request (0xFFFFFFFF = miss), written in phase A
    uint32_t* seg_ver;          // nseg: first verify entry of a slice whose tiles are checksummed by the gather, else 0xFFFFFFFF
    uint32_t fuse_verify;       // 1 = the gather variant in use checksums tiles in shared memory
    uint32_t* infl_seg_begin;   // n + 1: exclusive prefix of the inflate members' segment counts (k_inflate_seg's work list)
    int32_t* coal_tab;          // coal_cap ints: hash table for request coalescing of waves too large for shared memory (or nullptr)
    uint32_t coal_cap;          // power of two
    uint32_t* ver_stat;         // mapped pinned: [0] = verify chunks of this wave, [1] = of which fused (written by k_plan_verify),
                                // [2] = completion flag of a fused small wave (written last by k_plan, polled by nz_get_wait)
};
cudaError_t nz_launch_lookup(const NzSlot* table, uint64_t line_mask, const NzDesc* descs, const uint64_t* keys,
                             uint32_t n, nz_extent* out, cudaStream_t s);
cudaError_t nz_launch_insert(NzSlot* table, uint64_t line_mask, uint64_t k0, uint64_t k1, uint32_t desc, uint64_t gen,
                             int32_t* status, cudaStream_t s);
cudaError_t nz_launch_erase(NzSlot* table, uint64_t line_mask, uint64_t k0, uint64_t k1, int32_t* status, cudaStream_t s);
cudaError_t nz_launch_plan(const NzPlanArgs& a, cudaStream_t s, const NzSmallWave* small = nullptr, bool fused_copy = false);
cudaError_t nz_launch_gather(int variant, const NzSegWork* work, const uint64_t* tile_begin, const uint32_t* chain_head,
                             const uint32_t* chain_next, uint32_t nseg, int grid, cudaStream_t s,
                             const uint32_t* seg_ver = nullptr, NzVerWork* ver = nullptr, const uint32_t* crc_tables = nullptr,
                             uint32_t pace_ns = 0, uint32_t spread = 0, int stages_hint = 0, uint32_t* dyn = nullptr);
bool nz_gather_fuses_verify(int variant);
uint32_t nz_gather_tile_bytes(int variant);
cudaError_t nz_launch_verify(const NzPlanArgs& a, const uint32_t* crc_tables, int grid, cudaStream_t s);
cudaError_t nz_launch_crc_jobs(const NzCrcJob* jobs, uint32_t n_chunks, const NzFoldJob* folds, uint32_t n_objects, const uint32_t* crc_tables,
                               int sm_count, cudaStream_t s);
cudaError_t nz_launch_crc_chunks(const uint8_t* p, uint64_t n, uint32_t* chunk_out, uint32_t* whole_out,
                                 const uint32_t* crc_tables, int sm_count, cudaStream_t s);
cudaError_t nz_launch_inflate(const NzInflWork* infl, uint64_t* meta, uint32_t* counters, NzDevRes* res,
                              const uint32_t* crc_tables, int grid, int sm_count, cudaStream_t s,
                              const uint32_t* seg_begin, const NzDevReq* reqs, uint32_t n_req, int seg_grid, bool legacy);
cudaError_t nz_launch_index(const NzIdxJob* jobs, NzIdxResult* results, uint32_t n, uint32_t* counters, NzInflWork* infl,
                            uint64_t* meta, NzDevRes* res, const uint32_t* crc_tables, int sm_count, cudaStream_t s);
cudaError_t nz_launch_index_pack(const NzIdxJob* jobs, const NzIdxResult* results, uint32_t n, cudaStream_t s);
cudaError_t nz_launch_publish(NzSlot* table, uint64_t line_mask, NzDesc* descs, const NzPubItem* items, uint32_t n,
                              int32_t* status, cudaStream_t s);
int nz_inflate_seg_blocks_per_sm();
int64_t nz_cufile_read(int fd, uint64_t file_off, void* dev_base, uint64_t dev_off, uint64_t len);
void nz_cufile_why(char* out, size_t cap);
cudaError_t nz_upload_constants(uint32_t* crc_tables_dev);
cudaError_t nz_inflate_upload_constants();
int nz_gather_blocks_per_sm(int variant);
int nz_inflate_blocks_per_sm();
