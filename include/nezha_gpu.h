/*
 * include/nezha_gpu.h — C ABI of libnezha_gpu.so, the B200 cache-hit data path.
 *
 * What this boundary replaces.  The reference (fast-ml/nezha) has NO plugin /
 * FFI surface: it is built CGO_ENABLED=0 (/root/reference/Makefile:9,13) and its
 * data plane is a separate stock nginx process configured by
 * /root/reference/examples/demo/nginx.conf.  BASELINE.json:north_star therefore
 * prescribes a NEW thin cgo shim; each entry point below cites the reference
 * behaviour (nginx directive) whose work it takes over.  INTEGRATION.md shows the
 * cgo stub a maintainer adds on the Go side.
 *
 * Rules of the ABI
 *  - C linkage, plain pointers and sizes, no C++ or torch types, no exceptions.
 *  - every function returns int32_t: 0 = NZ_OK, negative = NZ_E*; nz_strerror().
 *  - every entry point may be called concurrently from any OS thread (cgo moves
 *    goroutines between threads); the library sets the CUDA device itself.
 *  - all I/O memory is C-owned (nz_buf); Go wraps it with unsafe.Slice and never
 *    passes Go-heap pointers that must outlive a call.
 *  - there is NO CPU fallback: without a CUDA device nz_init fails NZ_ECUDA.
 */
#ifndef NEZHA_GPU_H
#define NEZHA_GPU_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NZ_ABI_VERSION 2

/* ---- status codes ------------------------------------------------------ */
#define NZ_OK          0
#define NZ_EINVAL     (-1)   /* bad argument / range outside the object          */
#define NZ_ENOMEM     (-2)   /* arena, descriptor table or host memory exhausted  */
#define NZ_ECUDA      (-3)   /* CUDA runtime error (no device, launch failure)    */
#define NZ_ENOTFOUND  (-4)   /* key not in the table (a MISS; per-request status) */
#define NZ_ECORRUPT   (-5)   /* CRC-32C mismatch / malformed gzip member          */
#define NZ_ETIMEOUT   (-6)   /* nz_get_wait timed out; ticket still valid         */
#define NZ_EEXIST     (-7)   /* put of a key that is already live                 */
#define NZ_EFULL      (-8)   /* hash table load limit reached                     */
#define NZ_ENOSYS     (-9)   /* feature unavailable in this build / on this box   */
#define NZ_ERANGE     (-10)  /* destination buffer too small                      */

/* ---- object flags (stored) --------------------------------------------- */
#define NZ_ENC_IDENTITY 0u
#define NZ_ENC_GZIP     1u    /* stored bytes are one RFC 1952 member             */
#define NZ_ENC_STRIPE   4u    /* put: this object is one stripe of a larger object */
                              /* (placed by the stripe plan, not by its own key hash) */
#define NZ_ENC_STRIPED  8u    /* the object's bytes live in stripes on several GPUs */
#define NZ_ENC_INFLATABLE 16u /* set by nz_put_commit on an NZ_ENC_GZIP object whose member inflated cleanly there
                                 (trailer CRC-32 and ISIZE checked): raw_length is exact.  Never passed to nz_put_begin. */

/* ---- get flags ---------------------------------------------------------- */
#define NZ_GET_INFLATE  1u    /* serve the inflated representation (SPEC G)       */
#define NZ_GET_VERIFY   2u    /* recompute CRC-32C of every 64 KiB chunk touched  */
#define NZ_GET_URGENT   4u    /* latency-critical (first window of a response): the wave runs on a
                                 high-priority stream so it is not queued behind bulk waves      */

/* ---- buffer kinds -------------------------------------------------------- */
#define NZ_BUF_PINNED_HOST 0  /* cudaHostAlloc, mapped: kernels write it directly */
#define NZ_BUF_DEVICE      1  /* HBM response buffer                              */

#define NZ_CRC_CHUNK 65536u   /* SPEC X2 */

typedef struct nz_ctx    nz_ctx;
typedef struct nz_buf    nz_buf;
typedef struct nz_put    nz_put;
typedef struct nz_ticket nz_ticket;

typedef struct nz_config {
    uint32_t struct_size;    /* sizeof(nz_config), for forward compatibility     */
    int32_t  device;         /* CUDA ordinal                                      */
    uint64_t arena_bytes;    /* HBM reserved for object extents                   */
    uint64_t table_slots;    /* hash-table slots (rounded up to a power of two);  */
                             /* replaces keys_zone=my_cache:10m (nginx.conf:38)   */
    uint32_t max_objects;    /* extent descriptors                                */
    uint32_t shard_index;    /* this GPU's shard (key-hash sharding)   */
    uint32_t shard_count;    /* 0 or 1 = unsharded                                */
    uint32_t flags;          /* reserved, 0                                       */
} nz_config;

typedef struct nz_extent {   /* result of nz_lookup for one key                   */
    int32_t  status;         /* NZ_OK or NZ_ENOTFOUND                             */
    uint32_t flags;          /* NZ_ENC_*                                          */
    uint64_t stored_length;  /* bytes at rest                                     */
    uint64_t raw_length;     /* inflated length (== stored_length for identity)   */
    uint32_t crc32c;         /* CRC-32C of the stored bytes (x-goog-hash form)    */
    uint32_t desc_index;     /* descriptor slot (diagnostic)                      */
    uint64_t user_meta;      /* opaque host cookie given at put (headers blob id) */
} nz_extent;

typedef struct nz_seg {      /* one slice of the served representation            */
    uint64_t src_off;        /* offset in the representation (stored or inflated) */
    uint64_t len;
    uint64_t dst_off;        /* offset in the destination buffer                  */
} nz_seg;

typedef struct nz_get_req {
    uint8_t       key[16];   /* MD5 of "$host/$proxy_host$uri" (nginx.conf:56,38) */
    uint32_t      flags;     /* NZ_GET_*                                          */
    uint32_t      nseg;      /* 0 = whole object to dst_off 0                     */
    const nz_seg* segs;      /* nseg entries (range filter output, SPEC R8/R9)    */
    nz_buf*       dst;       /* response buffer; multipart part headers are       */
                             /* written by the host into the gaps between segs    */
    uint64_t      dst_base;  /* byte offset added to every dst_off                */
} nz_get_req;

typedef struct nz_get_res {
    int32_t  status;         /* NZ_OK, NZ_ENOTFOUND, NZ_EINVAL, NZ_ECORRUPT, NZ_ERANGE */
    uint32_t flags;          /* NZ_ENC_* of the stored object                     */
    uint64_t bytes;          /* payload bytes written                             */
    uint64_t length;         /* length of the representation served               */
    uint32_t crc32c;         /* stored whole-object CRC-32C                       */
    uint32_t reserved;
} nz_get_res;

typedef struct nz_stats {
    uint64_t objects, arena_bytes, arena_used, table_slots, table_live;
    uint64_t gets, get_bytes, hits, misses, puts, put_bytes, evictions, corrupt;
    uint64_t kernel_launches;   /* kernels launched by this library since init     */
    uint64_t file_cufile_bytes; /* nz_put_file bytes that arrived through cuFile/GDS */
    uint64_t file_bounce_bytes; /* ... and through pread + pinned bounce buffers    */
    uint64_t verify_chunks;       /* 64 KiB chunks checked for NZ_GET_VERIFY (stored-byte GETs) */
    uint64_t verify_fused_chunks; /* ... of which the gather checksummed from its shared-memory copy (no second HBM read) */
    uint64_t host_dma_waves;      /* waves to pinned host memory that went HBM staging slot -> copy engine (8 MiB .. 1 GiB, explicit slices) */
    uint64_t inflate_segments;    /* index segments decoded by GETs (one warp each; K4 with the checkpoint index) */
    uint64_t inflate_whole_members; /* members GETs decoded with one warp each because they carry no index */
    uint64_t index_bytes;         /* arena bytes held by checkpoint indexes of live gzip objects */
} nz_stats;

/* ---- lifecycle ----------------------------------------------------------- */
int32_t     nz_abi_version(void);
const char* nz_strerror(int32_t code);
/* last CUDA error string seen by this thread's most recent failing call */
const char* nz_last_error(void);
int32_t nz_init(const nz_config* cfg, nz_ctx** out);
int32_t nz_shutdown(nz_ctx* ctx);

/* ---- C-owned I/O memory -------------------------------------------------- */
/* NZ_BUF_PINNED_HOST pages are placed on the host memory node next to ctx's GPU whatever thread calls (on a
 * two-socket node a window on the far socket moves at about a third of the near one); NZ_NUMA=0 turns that off. */
int32_t  nz_buf_alloc(nz_ctx* ctx, uint64_t bytes, int32_t kind, nz_buf** out);
/* host memory node next to ctx's GPU (-1: unknown): callers that pool windows keep one pool per node */
int32_t  nz_numa_node(nz_ctx* ctx);
/* wrap device memory owned by the caller (e.g. a framework tensor); not freed */
int32_t  nz_buf_wrap_device(nz_ctx* ctx, void* devptr, uint64_t bytes, nz_buf** out);
void*    nz_buf_ptr(nz_buf* b);       /* host address (NULL for device buffers)   */
void*    nz_buf_devptr(nz_buf* b);    /* device address (mapped alias for pinned) */
uint64_t nz_buf_size(nz_buf* b);
int32_t  nz_buf_free(nz_buf* b);
/* synchronous copies between a buffer and host memory (tools/tests; device buffers too) */
int32_t  nz_buf_read(nz_buf* b, uint64_t off, void* host_dst, uint64_t n);
int32_t  nz_buf_write(nz_buf* b, uint64_t off, const void* host_src, uint64_t n);

/* ---- fill side: replaces nginx temp-file -> rename (SPEC H) - */
int32_t nz_put_begin(nz_ctx* ctx, const uint8_t key[16], uint64_t stored_len,
                     uint64_t raw_len, uint32_t enc_flags, uint64_t user_meta,
                     nz_put** out);
/* copy len bytes from host memory (pinned nz_buf memory goes async) */
int32_t nz_put_write(nz_put* p, const void* src_host, uint64_t dst_off, uint64_t len);
/* wait until every copy queued on this put has left its source buffer (to reuse a pinned window) */
int32_t nz_put_sync(nz_put* p);
/* copy from device memory (synthetic fills, peer fills) */
int32_t nz_put_write_device(nz_put* p, const void* src_dev, uint64_t dst_off, uint64_t len);
/* read from a file: cuFile/GDS when available, else pread into pinned bounce buffers */
int32_t nz_put_file(nz_put* p, int fd, uint64_t file_off, uint64_t dst_off, uint64_t len);
/* which stage of the cuFile/GDS path last decided ("ok", or the call that refused and its error codes) */
const char* nz_cufile_status(void);
/* computes the 64 KiB chunk CRC-32C table and the whole-object CRC on the device;
 * if has_expected and it differs -> NZ_ECORRUPT and nothing is published; else the
 * slot is published (visible to GETs).  The handle is consumed either way.
 * NZ_ENC_GZIP objects are inflated once here (trailer CRC-32 and ISIZE checked): raw_length becomes the member's
 * real inflated length whatever nz_put_begin was told (pass 0 when it is not known), and a checkpoint index is
 * stored next to the member so that a GET with NZ_GET_INFLATE decodes it with one warp per ~128 KiB of output
 * instead of one warp per member.  A malformed member is still stored (its bytes are served verbatim, SPEC G4);
 * inflate GETs on it report NZ_ECORRUPT. */
int32_t nz_put_commit(nz_put* p, uint32_t expected_crc32c, int32_t has_expected,
                      uint32_t* out_crc32c);
/* Commit n objects of one context in one go (bulk loads: a cold start from the PV, a bench fill): the per-object
 * kernels are batched and the table is published with one launch.  expected_crc32c may be NULL (nothing to compare);
 * has_expected may be NULL (every entry of expected_crc32c counts).  out_crc32c / out_status (n entries each) may be
 * NULL.  Every handle is consumed.  Returns NZ_OK when the call ran; per-object outcomes are in out_status. */
int32_t nz_put_commit_many(nz_put** puts, uint32_t n, const uint32_t* expected_crc32c, const int32_t* has_expected,
                           uint32_t* out_crc32c, int32_t* out_status);
/* Bulk fill of n identity objects that lie in ONE pinned window (nz_buf_alloc ... NZ_BUF_PINNED_HOST) at window+src_off[i],
 * len[i] bytes each: a single host->device copy of the span they cover, one scatter launch into their extents, batched
 * CRC-32C, one publish launch.  For cold starts with very many small objects (10^6 per GPU): nz_put_begin + nz_put_write per
 * object cost a cudaMemcpyAsync each.  user_meta, out_crc32c, out_status (n entries each) may be NULL.  All or nothing on
 * argument / allocation errors (NZ_EEXIST, NZ_ENOMEM ...: nothing is stored); per-object outcomes otherwise in out_status. */
int32_t nz_put_packed(nz_ctx* ctx, uint32_t n, const uint8_t (*keys)[16], const uint64_t* src_off, const uint64_t* len,
                      nz_buf* window, const uint64_t* user_meta, uint32_t* out_crc32c, int32_t* out_status);
int32_t nz_put_abort(nz_put* p);
int32_t nz_evict(nz_ctx* ctx, const uint8_t key[16]);

/* ---- hit side ------------------------------------------------------------ */
/* K1 probe: replaces the shm rbtree lookup of keys_zone (nginx.conf:38) */
int32_t nz_lookup(nz_ctx* ctx, const uint8_t (*keys)[16], uint32_t n, nz_extent* out);

/* One wave of GETs: probe + plan + range gather (+inflate, +verify) on the device.
 * Replaces the cache-file read loop + range filter.  `cuda_stream`
 * is a cudaStream_t to launch on, or NULL for the library's own stream pool.
 * Returns immediately; results via the ticket.  A request whose status is not NZ_OK
 * leaves its destination unspecified (large host-destination waves are staged in HBM
 * and moved out by the copy engine whatever the per-request outcome). */
int32_t nz_get_submit(nz_ctx* ctx, const nz_get_req* reqs, uint32_t n,
                      void* cuda_stream, nz_ticket** out);
/* timeout_us < 0 waits forever; res (n entries) may be NULL.  Short waves are polled (a wave of a few small
 * slices completes through a flag its kernel writes into pinned memory), long ones sleep between polls. */
int32_t nz_get_wait(nz_ticket* t, int64_t timeout_us, nz_get_res* res);
int32_t nz_get_poll(nz_ticket* t);                 /* NZ_OK done, NZ_ETIMEOUT not yet */
int32_t nz_ticket_elapsed_ms(nz_ticket* t, float* ms);  /* device time of the wave */
int32_t nz_ticket_gather_ms(nz_ticket* t, float* ms);   /* device time of its gather kernel (K2) alone */
int32_t nz_ticket_inflate_ms(nz_ticket* t, float* ms);  /* device time of its inflate kernels (K4); 0 if none ran */
int32_t nz_ticket_free(nz_ticket* t);

int32_t nz_get_stats(nz_ctx* ctx, nz_stats* out);
/* record per-kernel CUDA events in every wave so nz_ticket_*_ms work (off by default: each
 * event record costs host time on the latency path) */
int32_t nz_set_profiling(nz_ctx* ctx, int32_t on);
/* Waves whose destinations are all pinned host memory (a front end streaming bodies to sockets):
 * bound what they offer the PCIe link.  total_gbs > 0 paces the non-urgent ones so that together they
 * stay near that rate (shared between the waves in flight); ctas > 0 sets how many CTAs such a wave
 * runs on.  A saturated link keeps ~150-300 us of posted writes queued ahead of every later wave; kept
 * just below saturation, a first window (NZ_GET_URGENT) still completes in ~40-60 us.
 * 0 / 0 restores the defaults (unpaced, 32 CTAs): maximum throughput, the bench's e2e mode.
 * Nothing in the reference corresponds to it (examples/demo/nginx.conf sets no limit_rate /
 * sendfile_max_chunk; its one worker never fills a link). */
int32_t nz_set_host_rate(nz_ctx* ctx, double total_gbs, int32_t ctas);

/* ---- cross-shard objects (the one exchange step) ---------------- */
/* An object larger than one stripe is striped round-robin over the shards; the owner of
 * its key holds the stripe map.  A GET on the owner gathers every stripe with direct peer
 * loads over NVLink inside the same K2 kernel (no separate collective): the other GPUs'
 * arenas are mapped through CUDA IPC. */
#define NZ_IPC_HANDLE_BYTES 64
typedef struct nz_stripe {
    uint32_t shard;          /* GPU that holds this stripe                         */
    uint32_t reserved;
    uint64_t arena_off;      /* where it sits in that GPU's arena (nz_extent_location) */
    uint64_t len;
    uint64_t crc_tab_off;    /* its 64 KiB chunk-CRC table in that arena           */
} nz_stripe;
int32_t nz_arena_ipc_export(nz_ctx* ctx, uint8_t handle[NZ_IPC_HANDLE_BYTES], uint64_t* arena_bytes);
int32_t nz_peer_attach(nz_ctx* ctx, uint32_t shard, const uint8_t handle[NZ_IPC_HANDLE_BYTES]);
int32_t nz_extent_location(nz_ctx* ctx, const uint8_t key[16], uint64_t* arena_off,
                           uint64_t* stored_len, uint64_t* crc_tab_off, uint32_t* crc32c);
/* publish the stripe map on the owner; every stripe but the last must be stripe_bytes long */
int32_t nz_put_striped(nz_ctx* ctx, const uint8_t key[16], uint64_t total_len, uint64_t stripe_bytes,
                       uint32_t n_stripes, const nz_stripe* stripes, uint32_t crc32c, uint64_t user_meta);
/* CRC-32C of A||B from crc(A), crc(B), len(B) (host arithmetic; whole-object CRC of a striped object) */
uint32_t nz_crc32c_combine(uint32_t crc_a, uint32_t crc_b, uint64_t len_b);

/* ---- standalone kernels exposed for parity tests and profiling ------------ */
/* CRC-32C of n device bytes (K3); chunk table (may be NULL) gets ceil(n/64Ki) u32 */
int32_t nz_crc32c_device(nz_ctx* ctx, const void* dev, uint64_t n, uint32_t* out_crc,
                         uint32_t* out_chunks_host);
/* fault injection for tests: XOR one stored byte of a live object in HBM */
int32_t nz_debug_poke(nz_ctx* ctx, const uint8_t key[16], uint64_t off, uint8_t xor_mask);
/* shard that owns a key: u64le(key[0:8]) mod shard_count */
uint32_t nz_shard_of(const uint8_t key[16], uint32_t shard_count);

#ifdef __cplusplus
}
#endif
#endif /* NEZHA_GPU_H */
