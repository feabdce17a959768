// nezha_b200/csrc/nz_api.cu — the C ABI of include/nezha_gpu.h.
//
// Host side of the cache zone: arena allocator, descriptor slots, shadow index,
// stream pool and GET waves.  All data-path work is CUDA (nz_kernels.cu,
// nz_inflate.cu); there is deliberately no CPU fallback — if CUDA is missing every
// entry point fails with NZ_ECUDA.
#include <cuda_runtime.h>
#include <errno.h>
#include <sched.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <sys/syscall.h>
#include <unistd.h>

#include <algorithm>
#include <chrono>
#include <functional>
#include <string>
#include <thread>

#include "nz_internal.h"

static thread_local char tl_err[512] = "";

static int32_t cuda_fail(cudaError_t e, const char* what) {
    snprintf(tl_err, sizeof tl_err, "%s: %s (%s)", what, cudaGetErrorName(e), cudaGetErrorString(e));
    return NZ_ECUDA;
}
#define CK(call)                                              \
    do {                                                      \
        cudaError_t _e = (call);                              \
        if (_e != cudaSuccess) return cuda_fail(_e, #call);   \
    } while (0)

static inline uint64_t rd64(const uint8_t* p) {
    uint64_t v;
    memcpy(&v, p, 8);
    return v;   // little-endian host
}
static inline uint64_t round_up(uint64_t v, uint64_t a) { return (v + a - 1) / a * a; }
static uint32_t nz_gf_mul_c(uint32_t a, uint32_t b) {   // GF(2)[x] mod the reflected Castagnoli polynomial
    uint32_t p = 0;
    for (int i = 0; i < 32; ++i) {
        p ^= b & (0u - ((a >> (31 - i)) & 1u));
        b = (b >> 1) ^ (0x82F63B78u & (0u - (b & 1u)));
    }
    return p;
}

extern "C" {

int32_t nz_abi_version(void) { return NZ_ABI_VERSION; }

const char* nz_strerror(int32_t code) {
    switch (code) {
    case NZ_OK: return "ok";
    case NZ_EINVAL: return "invalid argument or range outside the object";
    case NZ_ENOMEM: return "out of arena / descriptor / host memory";
    case NZ_ECUDA: return "CUDA error (no device or launch failure)";
    case NZ_ENOTFOUND: return "key not found (miss)";
    case NZ_ECORRUPT: return "CRC-32C mismatch or malformed gzip member";
    case NZ_ETIMEOUT: return "timed out";
    case NZ_EEXIST: return "key already present";
    case NZ_EFULL: return "hash table full";
    case NZ_ENOSYS: return "not available";
    case NZ_ERANGE: return "destination buffer too small";
    default: return "unknown error";
    }
}
const char* nz_last_error(void) { return tl_err; }

uint32_t nz_shard_of(const uint8_t key[16], uint32_t shard_count) {
    if (shard_count <= 1) return 0;
    return (uint32_t)(rd64(key) % shard_count);
}


// ---------------------------------------------------------------------------
// host memory placement
// ---------------------------------------------------------------------------
// A pinned window is written by the GPU's copy engine through the PCIe root of ONE socket.  Pages on the other
// socket cost a trip over the inter-socket link for every byte (measured on the 2-socket B200 node: a window
// placed on the far node moves at ~0.37x the near one, profiles/r02_d2h_probe_n8.txt).  Callers of the C ABI are
// goroutines / connection threads that run anywhere, so the library places the pages itself: while it allocates
// and pins, the calling thread prefers the GPU's node (set_mempolicy; where a container's seccomp profile refuses
// that, the thread is moved onto that node's CPUs for the duration, so first touch lands there).
static int read_int_file(const char* path, int dflt) {
    FILE* f = fopen(path, "r");
    if (!f) return dflt;
    int v = dflt;
    if (fscanf(f, "%d", &v) != 1) v = dflt;
    fclose(f);
    return v;
}
static bool parse_cpulist(const char* path, cpu_set_t* out) {
    CPU_ZERO(out);
    FILE* f = fopen(path, "r");
    if (!f) return false;
    char buf[4096];
    const bool got = fgets(buf, sizeof buf, f) != nullptr;
    fclose(f);
    if (!got) return false;
    int n = 0;
    for (char* p = buf; *p && *p != '\n';) {
        char* e;
        long a = strtol(p, &e, 10), b = a;
        if (e == p) break;
        if (*e == '-') { p = e + 1; b = strtol(p, &e, 10); }
        for (long i = a; i <= b && i < CPU_SETSIZE; ++i) { CPU_SET((int)i, out); ++n; }
        p = (*e == ',') ? e + 1 : e;
    }
    return n > 0;
}
static void detect_numa(nz_ctx* c) {
    if (const char* ev = getenv("NZ_NUMA")) if (atoi(ev) == 0) return;
    char bus[32] = {0};
    if (cudaDeviceGetPCIBusId(bus, sizeof bus, c->device) != cudaSuccess) { cudaGetLastError(); return; }
    for (char* p = bus; *p; ++p) if (*p >= 'A' && *p <= 'Z') *p = (char)(*p - 'A' + 'a');
    char path[128];
    snprintf(path, sizeof path, "/sys/bus/pci/devices/%s/numa_node", bus);
    const int node = read_int_file(path, -1);
    if (node < 0 || node >= 64) return;
    c->numa_node = node;
    snprintf(path, sizeof path, "/sys/devices/system/node/node%d", node ? 0 : 1);
    c->numa_place = access(path, F_OK) == 0;             // a second node exists: placement matters
    if (const char* ev = getenv("NZ_NUMA")) if (atoi(ev) == 2) c->numa_place = false;   // node reported, placement left to the caller
}
struct NumaScope {
    bool policy = false, moved = false;
    cpu_set_t saved;
    explicit NumaScope(const nz_ctx* c) {
        if (!c->numa_place) return;
#ifdef SYS_set_mempolicy
        unsigned long mask = 1ul << c->numa_node;
        if (syscall(SYS_set_mempolicy, 1 /* MPOL_PREFERRED */, &mask, sizeof(mask) * 8 + 1) == 0) { policy = true; return; }
#endif
        char path[96];
        cpu_set_t want, both;
        snprintf(path, sizeof path, "/sys/devices/system/node/node%d/cpulist", c->numa_node);
        if (!parse_cpulist(path, &want) || sched_getaffinity(0, sizeof saved, &saved) != 0) return;
        CPU_AND(&both, &want, &saved);
        if (CPU_COUNT(&both) == 0 || CPU_EQUAL(&both, &saved)) return;      // not allowed on that node, or already only there
        moved = sched_setaffinity(0, sizeof both, &both) == 0;
    }
    ~NumaScope() {
#ifdef SYS_set_mempolicy
        if (policy) syscall(SYS_set_mempolicy, 0 /* MPOL_DEFAULT */, nullptr, 0);
#endif
        if (moved) sched_setaffinity(0, sizeof saved, &saved);
    }
};
int32_t nz_numa_node(nz_ctx* c) { return c ? c->numa_node : -1; }

// ---------------------------------------------------------------------------
// lifecycle
// ---------------------------------------------------------------------------
int32_t nz_init(const nz_config* cfg, nz_ctx** out) {
    if (!cfg || !out || cfg->struct_size < sizeof(nz_config)) return NZ_EINVAL;
    *out = nullptr;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess) return cuda_fail(e, "cudaGetDeviceCount");
    if (ndev <= 0 || cfg->device < 0 || cfg->device >= ndev) {
        snprintf(tl_err, sizeof tl_err, "no CUDA device %d (count %d)", cfg->device, ndev);
        return NZ_ECUDA;
    }
    CK(cudaSetDevice(cfg->device));
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, cfg->device));
    nz_ctx* c = new (std::nothrow) nz_ctx();
    if (!c) return NZ_ENOMEM;
    c->cfg = *cfg;
    c->device = cfg->device;
    c->sm_count = prop.multiProcessorCount;
    detect_numa(c);
    if (c->cfg.shard_count == 0) c->cfg.shard_count = 1;
    if (c->cfg.shard_index >= c->cfg.shard_count) { delete c; return NZ_EINVAL; }
    if (c->cfg.max_objects == 0) c->cfg.max_objects = 1u << 16;
    uint64_t slots = c->cfg.table_slots ? c->cfg.table_slots : 4ull * c->cfg.max_objects;
    if (slots < 2ull * c->cfg.max_objects) slots = 2ull * c->cfg.max_objects;
    uint64_t lines = 1;
    while (lines * 4 < slots) lines <<= 1;
    c->n_lines = lines;
    c->arena_bytes = round_up(c->cfg.arena_bytes ? c->cfg.arena_bytes : (1ull << 30), 4096);

#define CKI(call)                                                         \
    do {                                                                  \
        cudaError_t _e = (call);                                          \
        if (_e != cudaSuccess) { int32_t rc = cuda_fail(_e, #call); nz_shutdown(c); return rc; } \
    } while (0)
    CKI(cudaMalloc(&c->arena, c->arena_bytes + 4096));
    CKI(cudaMalloc(&c->table, lines * 4 * sizeof(NzSlot)));
    CKI(cudaMemset(c->table, 0, lines * 4 * sizeof(NzSlot)));
    CKI(cudaMalloc(&c->descs, (size_t)c->cfg.max_objects * sizeof(NzDesc)));
    CKI(cudaMemset(c->descs, 0, (size_t)c->cfg.max_objects * sizeof(NzDesc)));
    CKI(cudaMalloc(&c->crc_tables, NZ_CRC_TABLE_WORDS * sizeof(uint32_t)));   // slice tables + fast-path tables + fused-verify tables (nz_crc.cuh)
    CKI(nz_upload_constants(c->crc_tables));
    for (int i = 0; i < 8; ++i) {
        cudaStream_t s;
        CKI(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
        c->streams.push_back(s);
    }
    CKI(cudaStreamCreateWithFlags(&c->put_stream, cudaStreamNonBlocking));
    {
        int lo = 0, hi = 0;
        CKI(cudaDeviceGetStreamPriorityRange(&lo, &hi));
        for (int i = 0; i < 4; ++i) {
            cudaStream_t s;
            CKI(cudaStreamCreateWithPriority(&s, cudaStreamNonBlocking, hi));
            c->hi_streams.push_back(s);
        }
        if (const char* ev = getenv("NZ_HOST_GRID")) c->host_grid = atoi(ev) > 0 ? atoi(ev) : c->host_grid.load();
        if (const char* ev = getenv("NZ_HOST_RATE_GBS")) c->host_rate_gbs = atof(ev);
        if (const char* ev = getenv("NZ_HOST_DMA")) c->host_dma = atoi(ev) != 0;
    }
    CKI(cudaMalloc(&c->peer_base_dev, sizeof(uint64_t) * c->cfg.shard_count));
    CKI(cudaMemset(c->peer_base_dev, 0, sizeof(uint64_t) * c->cfg.shard_count));
    c->peer_ptrs.assign(c->cfg.shard_count, nullptr);
    c->peer_ptrs[c->cfg.shard_index] = c->arena;
    {
        const uint64_t self = (uint64_t)(uintptr_t)c->arena;
        CKI(cudaMemcpy(c->peer_base_dev + c->cfg.shard_index, &self, 8, cudaMemcpyHostToDevice));
    }
#undef CKI
    c->free_by_off[0] = c->arena_bytes;
    c->free_by_len.emplace(c->arena_bytes, 0);
    c->free_desc.reserve(c->cfg.max_objects);
    for (uint32_t i = c->cfg.max_objects; i-- > 0;) c->free_desc.push_back(i);
    if (const char* ev = getenv("NZ_GATHER_VARIANT"))      // "0" / "reg": register path only; anything else: the default bulk-copy kernel
        c->gather_variant = (!strcmp(ev, "0") || !strcmp(ev, "reg")) ? NZ_GATHER_REG : NZ_GATHER_TMA;
    if (const char* ev = getenv("NZ_COALESCE")) c->coalesce = atoi(ev) != 0;
    if (const char* ev = getenv("NZ_SPAN_DYN")) c->span_dynamic = atoi(ev) != 0;
    if (const char* ev = getenv("NZ_SPAN_SPREAD")) { const int v = atoi(ev); if (v >= 0 && v <= 4096) c->span_spread = (uint32_t)v; }
    if (const char* ev = getenv("NZ_PROFILE")) c->profile = atoi(ev) != 0;
    if (const char* ev = getenv("NZ_FUSE_VERIFY")) c->fuse_verify = atoi(ev) != 0;
    if (const char* ev = getenv("NZ_FUSE_SMALL")) c->fuse_small = atoi(ev) != 0;
    c->gather_grid = c->sm_count * nz_gather_blocks_per_sm(c->gather_variant);
    c->gather_grid3 = c->sm_count * 2;              // three 32 KiB stages: two CTAs fit an SM
    if (const char* ev = getenv("NZ_GATHER_CTAS_PER_SM")) { const int v = atoi(ev); if (v >= 1 && v <= 16) c->gather_grid = c->sm_count * v; }
    c->inflate_grid = c->sm_count * nz_inflate_blocks_per_sm();
    c->inflate_seg_grid = c->sm_count * nz_inflate_seg_blocks_per_sm();
    if (const char* ev = getenv("NZ_INFL_SEG_KB")) { const int v = atoi(ev); if (v >= 16 && v <= 4096) c->idx_seg_bytes = (uint32_t)v << 10; }
    if (const char* ev = getenv("NZ_INFL_INDEX")) c->idx_build = atoi(ev) != 0;
    if (const char* ev = getenv("NZ_INFL_DEDUPE")) c->infl_dedupe = atoi(ev) != 0;
    {   // temporary buffers of the index build come from the stream-ordered pool: keep what it has freed
        cudaMemPool_t pool;
        if (cudaDeviceGetDefaultMemPool(&pool, c->device) == cudaSuccess) {
            uint64_t keep = 1ull << 30;
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
        }
        cudaGetLastError();
    }
    *out = c;
    return NZ_OK;
}

static void ticket_destroy(nz_ticket* t) {
    if (!t) return;
    if (t->scratch) cudaFree(t->scratch);
    if (t->pin) cudaFreeHost(t->pin);
    if (t->dev) cudaFree(t->dev);
    if (t->ev_start) cudaEventDestroy(t->ev_start);
    if (t->ev_done) cudaEventDestroy(t->ev_done);
    if (t->ev_g0) cudaEventDestroy(t->ev_g0);
    if (t->ev_g1) cudaEventDestroy(t->ev_g1);
    if (t->ev_i0) cudaEventDestroy(t->ev_i0);
    if (t->ev_i1) cudaEventDestroy(t->ev_i1);
    delete t;
}

int32_t nz_shutdown(nz_ctx* c) {
    if (!c) return NZ_EINVAL;
    cudaSetDevice(c->device);
    cudaDeviceSynchronize();
    for (nz_ticket* t : c->ticket_pool) ticket_destroy(t);
    for (cudaStream_t s : c->streams) cudaStreamDestroy(s);
    for (cudaStream_t s : c->hi_streams) cudaStreamDestroy(s);
    if (c->put_stream) cudaStreamDestroy(c->put_stream);
    if (c->bounce[0]) cudaFreeHost(c->bounce[0]);
    for (size_t i = 0; i < c->peer_ptrs.size(); ++i)
        if (c->peer_ptrs[i] && i != c->cfg.shard_index) cudaIpcCloseMemHandle(c->peer_ptrs[i]);
    if (c->peer_base_dev) cudaFree(c->peer_base_dev);
    if (c->arena) cudaFree(c->arena);
    if (c->table) cudaFree(c->table);
    if (c->descs) cudaFree(c->descs);
    if (c->crc_tables) cudaFree(c->crc_tables);
    for (auto& sl : c->stage) { if (sl.buf) cudaFree(sl.buf); if (sl.free_ev) cudaEventDestroy(sl.free_ev); }
    delete c;
    return NZ_OK;
}

// ---------------------------------------------------------------------------
// buffers
// ---------------------------------------------------------------------------
int32_t nz_buf_alloc(nz_ctx* c, uint64_t bytes, int32_t kind, nz_buf** out) {
    if (!c || !out || bytes == 0) return NZ_EINVAL;
    CK(cudaSetDevice(c->device));
    nz_buf* b = new (std::nothrow) nz_buf();
    if (!b) return NZ_ENOMEM;
    b->ctx = c; b->size = bytes; b->kind = kind; b->owned = true; b->host = nullptr; b->dev = nullptr;
    cudaError_t e;
    if (kind == NZ_BUF_PINNED_HOST) {
        NumaScope near(c);
        e = cudaHostAlloc(&b->host, bytes + 256, cudaHostAllocMapped | cudaHostAllocPortable);
        if (e == cudaSuccess) e = cudaHostGetDevicePointer(&b->dev, b->host, 0);
    } else if (kind == NZ_BUF_DEVICE) {
        e = cudaMalloc(&b->dev, bytes + 256);
    } else {
        delete b;
        return NZ_EINVAL;
    }
    if (e != cudaSuccess) {
        if (b->host) cudaFreeHost(b->host);
        delete b;
        cudaGetLastError();
        cuda_fail(e, "nz_buf_alloc");
        return e == cudaErrorMemoryAllocation ? NZ_ENOMEM : NZ_ECUDA;
    }
    *out = b;
    return NZ_OK;
}
int32_t nz_buf_wrap_device(nz_ctx* c, void* devptr, uint64_t bytes, nz_buf** out) {
    if (!c || !out || !devptr || bytes == 0) return NZ_EINVAL;
    nz_buf* b = new (std::nothrow) nz_buf();
    if (!b) return NZ_ENOMEM;
    b->ctx = c; b->size = bytes; b->kind = NZ_BUF_DEVICE; b->owned = false; b->host = nullptr; b->dev = devptr;
    *out = b;
    return NZ_OK;
}
void* nz_buf_ptr(nz_buf* b) { return b ? b->host : nullptr; }
void* nz_buf_devptr(nz_buf* b) { return b ? b->dev : nullptr; }
uint64_t nz_buf_size(nz_buf* b) { return b ? b->size : 0; }
int32_t nz_buf_free(nz_buf* b) {
    if (!b) return NZ_EINVAL;
    if (b->owned) {
        cudaSetDevice(b->ctx->device);
        if (b->kind == NZ_BUF_PINNED_HOST) cudaFreeHost(b->host);
        else cudaFree(b->dev);
    }
    delete b;
    return NZ_OK;
}

int32_t nz_buf_read(nz_buf* b, uint64_t off, void* host_dst, uint64_t n) {
    if (!b || (!host_dst && n) || off > b->size || n > b->size - off) return NZ_EINVAL;
    if (!n) return NZ_OK;
    CK(cudaSetDevice(b->ctx->device));
    if (b->host) { memcpy(host_dst, (char*)b->host + off, n); return NZ_OK; }
    CK(cudaMemcpy(host_dst, (char*)b->dev + off, n, cudaMemcpyDeviceToHost));
    return NZ_OK;
}
int32_t nz_buf_write(nz_buf* b, uint64_t off, const void* host_src, uint64_t n) {
    if (!b || (!host_src && n) || off > b->size || n > b->size - off) return NZ_EINVAL;
    if (!n) return NZ_OK;
    CK(cudaSetDevice(b->ctx->device));
    if (b->host) { memcpy((char*)b->host + off, host_src, n); return NZ_OK; }
    CK(cudaMemcpy((char*)b->dev + off, host_src, n, cudaMemcpyHostToDevice));
    return NZ_OK;
}

// ---------------------------------------------------------------------------
// arena: free extents indexed twice - by offset (coalescing on free) and by (length, offset) (best fit in
// O(log n): the smallest extent that holds the request, lowest address among equals) - so a zone with 10^6
// objects and a churned free list allocates as fast as an empty one.  Extents are 4 KiB aligned so the same
// offsets work for cuFile/GDS reads from the PV.
// ---------------------------------------------------------------------------
static void free_insert(nz_ctx* c, uint64_t off, uint64_t len) { c->free_by_off[off] = len; c->free_by_len.emplace(len, off); }
static void free_remove(nz_ctx* c, std::map<uint64_t, uint64_t>::iterator it) {
    c->free_by_len.erase({it->second, it->first});
    c->free_by_off.erase(it);
}
static bool arena_alloc(nz_ctx* c, uint64_t len, uint64_t* off) {
    len = round_up(len ? len : 1, 4096);
    auto it = c->free_by_len.lower_bound({len, 0});
    if (it == c->free_by_len.end()) return false;
    const uint64_t have = it->first, o = it->second;
    c->free_by_len.erase(it);
    c->free_by_off.erase(o);
    if (have > len) free_insert(c, o + len, have - len);
    *off = o;
    return true;
}
static void arena_free(nz_ctx* c, uint64_t off, uint64_t len) {
    len = round_up(len ? len : 1, 4096);
    auto nx = c->free_by_off.lower_bound(off);
    if (nx != c->free_by_off.end() && off + len == nx->first) {          // merge with the extent behind
        len += nx->second;
        auto dead = nx++;
        free_remove(c, dead);
    }
    if (nx != c->free_by_off.begin()) {                                  // ... and with the one in front
        auto pv = std::prev(nx);
        if (pv->first + pv->second == off) {
            off = pv->first;
            len += pv->second;
            free_remove(c, pv);
        }
    }
    free_insert(c, off, len);
}

// small pinned mailboxes for kernel -> host scalars (status words, CRCs)
struct Mailbox { uint32_t crc; int32_t status; uint32_t pad[2]; };
static Mailbox* mailbox_get() {
    Mailbox* m = nullptr;
    if (cudaHostAlloc((void**)&m, sizeof(Mailbox), cudaHostAllocMapped | cudaHostAllocPortable) != cudaSuccess) return nullptr;
    m->crc = 0; m->status = NZ_ECUDA;
    return m;
}
struct MailboxPool {
    std::mutex mu;
    std::vector<Mailbox*> free_;
    Mailbox* get() {
        {
            std::lock_guard<std::mutex> g(mu);
            if (!free_.empty()) { Mailbox* m = free_.back(); free_.pop_back(); m->crc = 0; m->status = NZ_ECUDA; return m; }
        }
        return mailbox_get();
    }
    void put(Mailbox* m) { std::lock_guard<std::mutex> g(mu); free_.push_back(m); }
};
static MailboxPool g_mail;

// ---------------------------------------------------------------------------
// fill side
// ---------------------------------------------------------------------------
int32_t nz_put_begin(nz_ctx* c, const uint8_t key[16], uint64_t stored_len, uint64_t raw_len, uint32_t enc_flags,
                     uint64_t user_meta, nz_put** out) {
    if (!c || !key || !out) return NZ_EINVAL;
    if (enc_flags & ~(NZ_ENC_GZIP | NZ_ENC_STRIPE)) return NZ_EINVAL;
    if (!(enc_flags & NZ_ENC_GZIP)) raw_len = stored_len;
    // a stripe is placed by its parent's stripe plan, every other object by its own key hash
    if (!(enc_flags & NZ_ENC_STRIPE) && nz_shard_of(key, c->cfg.shard_count) != c->cfg.shard_index) return NZ_EINVAL;
    const uint64_t k0 = rd64(key), k1 = rd64(key + 8);
    const uint32_t n_chunks = (uint32_t)((stored_len + NZ_CRC_CHUNK - 1) / NZ_CRC_CHUNK);
    const uint64_t data_len = round_up(stored_len ? stored_len : 1, 4096);
    const uint64_t tab_len = round_up((uint64_t)(n_chunks ? n_chunks : 1) * 4, 256);
    nz_put* p = new (std::nothrow) nz_put();
    if (!p) return NZ_ENOMEM;
    {
        std::lock_guard<std::mutex> g(c->mu);
        if (c->shadow.count({k0, k1})) { delete p; return NZ_EEXIST; }
        if (c->shadow.size() * 2 >= c->n_lines * 4) { delete p; return NZ_EFULL; }
        if (c->free_desc.empty()) { delete p; return NZ_ENOMEM; }
        uint64_t off;
        if (!arena_alloc(c, data_len + tab_len, &off)) { delete p; return NZ_ENOMEM; }
        p->arena_off = off;
        p->arena_len = data_len + tab_len;
        // reserve the key so a concurrent put of the same key fails EEXIST
        NzHostObj ho;
        ho.desc = 0xFFFFFFFFu; ho.arena_off = off; ho.arena_len = p->arena_len;
        ho.stored_len = stored_len; ho.raw_len = raw_len; ho.flags = enc_flags;
        c->shadow[{k0, k1}] = ho;
    }
    p->ctx = c; p->k0 = k0; p->k1 = k1; p->stored_len = stored_len; p->raw_len = raw_len; p->user_meta = user_meta;
    p->flags = enc_flags; p->crc_tab_off = p->arena_off + data_len; p->n_chunks = n_chunks;
    p->stream = c->streams[c->rr.fetch_add(1) % c->streams.size()];
    *out = p;
    return NZ_OK;
}

static int32_t bounce_copy(nz_put* p, uint64_t dst_off, uint64_t len, const std::function<const char*(char*, uint64_t, size_t)>& reader);

int32_t nz_put_write(nz_put* p, const void* src_host, uint64_t dst_off, uint64_t len) {
    if (!p || (!src_host && len) || dst_off > p->stored_len || len > p->stored_len - dst_off) return NZ_EINVAL;
    if (!len) return NZ_OK;
    CK(cudaSetDevice(p->ctx->device));
    if (len >= (32ull << 20)) {          // big pageable source: the driver's own staging copy runs at ~10 GB/s; the bounce lanes at ~40
        cudaPointerAttributes at;
        const bool pageable = cudaPointerGetAttributes(&at, src_host) != cudaSuccess || at.type == cudaMemoryTypeUnregistered;
        cudaGetLastError();
        if (pageable)
            return bounce_copy(p, dst_off, len, [&](char* dst, uint64_t off, size_t n) -> const char* {
                memcpy(dst, static_cast<const char*>(src_host) + off, n);
                return nullptr;
            });
    }
    CK(cudaMemcpyAsync(p->ctx->arena + p->arena_off + dst_off, src_host, len, cudaMemcpyHostToDevice, p->stream));
    return NZ_OK;
}
int32_t nz_put_sync(nz_put* p) {
    if (!p) return NZ_EINVAL;
    CK(cudaSetDevice(p->ctx->device));
    CK(cudaStreamSynchronize(p->stream));
    return NZ_OK;
}
int32_t nz_put_write_device(nz_put* p, const void* src_dev, uint64_t dst_off, uint64_t len) {
    if (!p || (!src_dev && len) || dst_off > p->stored_len || len > p->stored_len - dst_off) return NZ_EINVAL;
    if (!len) return NZ_OK;
    CK(cudaSetDevice(p->ctx->device));
    CK(cudaMemcpyAsync(p->ctx->arena + p->arena_off + dst_off, src_dev, len, cudaMemcpyDeviceToDevice, p->stream));
    return NZ_OK;
}

// host bytes -> pinned bounce lanes -> cudaMemcpyAsync into the extent.  `reader(dst, off, n)` fills n bytes of the
// source starting at byte `off` (a pread from a file, a memcpy from pageable memory) and returns nullptr or what failed.
static int32_t bounce_copy(nz_put* p, uint64_t dst_off, uint64_t len, const std::function<const char*(char*, uint64_t, size_t)>& reader) {
    nz_ctx* c = p->ctx;
    // pread -> pinned bounce -> cudaMemcpyAsync.  One reader tops out at the page-cache copy rate
    // (~5-10 GB/s), so large reads are split over up to NZ_BOUNCE_LANES (12; env) reader threads, each with two pinned
    // 8 MiB buffers and its own stream: reads and H2D copies of all lanes overlap.  Lanes come from a pool shared by
    // concurrent calls (a cold start loads many files at once).
    constexpr int NZ_BOUNCE_LANES_MAX = 16;
    static const int NZ_BOUNCE_LANES = [] {
        int v = 12;
        if (const char* e = getenv("NZ_BOUNCE_LANES")) v = atoi(e);
        return v < 1 ? 1 : v > NZ_BOUNCE_LANES_MAX ? NZ_BOUNCE_LANES_MAX : v;
    }();
    constexpr size_t BUF = 8u << 20;
    const int want = (int)std::max<uint64_t>(1, std::min<uint64_t>((uint64_t)NZ_BOUNCE_LANES, len / (2 * BUF)));   // >= 16 MiB per lane
    std::vector<int> mine;
    {
        std::unique_lock<std::mutex> g(c->put_mu);
        if (!c->bounce[0]) {
            c->bounce_bytes = BUF * 2 * NZ_BOUNCE_LANES;
            NumaScope near(c);
            CK(cudaHostAlloc(&c->bounce[0], c->bounce_bytes, cudaHostAllocPortable));
            c->bounce[1] = c->bounce[0];
            c->lane_busy.assign(NZ_BOUNCE_LANES, 0);
        }
        c->lane_cv.wait(g, [&] { for (char b : c->lane_busy) if (!b) return true; return false; });
        for (int li = 0; li < NZ_BOUNCE_LANES && (int)mine.size() < want; ++li)
            if (!c->lane_busy[li]) { c->lane_busy[li] = 1; mine.push_back(li); }
    }
    struct LaneRelease {
        nz_ctx* c; const std::vector<int>& mine;
        ~LaneRelease() { { std::lock_guard<std::mutex> g(c->put_mu); for (int li : mine) c->lane_busy[li] = 0; } c->lane_cv.notify_all(); }
    } lane_release{c, mine};
    const int lanes = (int)mine.size();
    const uint64_t per = round_up((len + lanes - 1) / lanes, 4096);
    std::vector<int32_t> rcs(lanes, NZ_OK);
    std::vector<std::string> errs(lanes);
    auto lane_fn = [&](int li) {
        cudaSetDevice(c->device);
        const uint64_t beg = std::min<uint64_t>(len, per * li), end = std::min<uint64_t>(len, beg + per);
        if (beg >= end) return;
        cudaStream_t s = li == 0 ? p->stream : c->streams[(c->rr.fetch_add(1)) % c->streams.size()];
        cudaEvent_t ev[2];
        cudaEventCreateWithFlags(&ev[0], cudaEventDisableTiming);
        cudaEventCreateWithFlags(&ev[1], cudaEventDisableTiming);
        char* base = (char*)c->bounce[0] + (size_t)mine[li] * 2 * BUF;
        uint64_t done = beg;
        int which = 0;
        bool used[2] = {false, false};
        while (done < end && rcs[li] == NZ_OK) {
            const size_t want = (size_t)std::min<uint64_t>(BUF, end - done);
            if (used[which] && cudaEventSynchronize(ev[which]) != cudaSuccess) { rcs[li] = NZ_ECUDA; break; }
            const char* why = reader(base + which * BUF, done, want);
            if (why) { rcs[li] = NZ_EINVAL; errs[li] = why; break; }
            if (cudaMemcpyAsync(c->arena + p->arena_off + dst_off + done, base + which * BUF, want, cudaMemcpyHostToDevice, s) != cudaSuccess ||
                cudaEventRecord(ev[which], s) != cudaSuccess) { rcs[li] = NZ_ECUDA; break; }
            used[which] = true;
            done += want;
            which ^= 1;
        }
        if (cudaStreamSynchronize(s) != cudaSuccess && rcs[li] == NZ_OK) rcs[li] = NZ_ECUDA;
        cudaEventDestroy(ev[0]);
        cudaEventDestroy(ev[1]);
    };
    if (lanes == 1) {
        lane_fn(0);
    } else {
        std::vector<std::thread> th;
        for (int li = 1; li < lanes; ++li) th.emplace_back(lane_fn, li);
        lane_fn(0);
        for (auto& t : th) t.join();
    }
    for (int li = 0; li < lanes; ++li)
        if (rcs[li] != NZ_OK) {
            snprintf(tl_err, sizeof tl_err, "bounce lane %d: %s", li, errs[li].empty() ? "CUDA error" : errs[li].c_str());
            cudaGetLastError();
            return rcs[li];
        }
    return NZ_OK;
}

int32_t nz_put_file(nz_put* p, int fd, uint64_t file_off, uint64_t dst_off, uint64_t len) {
    if (!p || fd < 0 || dst_off > p->stored_len || len > p->stored_len - dst_off) return NZ_EINVAL;
    nz_ctx* c = p->ctx;
    CK(cudaSetDevice(c->device));
    if (len) {   // GPUDirect Storage first: the DMA lands straight in the extent
        CK(cudaStreamSynchronize(p->stream));
        const int64_t got = nz_cufile_read(fd, file_off, c->arena, p->arena_off + dst_off, len);
        if (got == (int64_t)len) { c->file_cufile_bytes += len; return NZ_OK; }
    }
    c->file_bounce_bytes += len;
    return bounce_copy(p, dst_off, len, [&](char* dst, uint64_t off, size_t n) -> const char* {
        size_t got = 0;
        while (got < n) {
            ssize_t r = pread(fd, dst + got, n - got, (off_t)(file_off + off + got));
            if (r < 0) { if (errno == EINTR) continue; return "pread failed"; }
            if (r == 0) return "pread: short file";
            got += (size_t)r;
        }
        return nullptr;
    });
}

const char* nz_cufile_status(void) {
    static thread_local char buf[256];
    nz_cufile_why(buf, sizeof buf);
    return buf;
}

static void put_release(nz_put* p, bool drop_reservation) {
    nz_ctx* c = p->ctx;
    if (drop_reservation) {
        std::lock_guard<std::mutex> g(c->mu);
        c->shadow.erase({p->k0, p->k1});
        arena_free(c, p->arena_off, p->arena_len);
    }
    delete p;
}

int32_t nz_put_abort(nz_put* p) {
    if (!p) return NZ_EINVAL;
    cudaSetDevice(p->ctx->device);
    cudaStreamSynchronize(p->stream);
    put_release(p, true);
    return NZ_OK;
}

// ---------------------------------------------------------------------------
// commit: CRC-32C chunk table + whole CRC (K3), checkpoint index of gzip members (k_inflate_index), batched publish (K6)
// ---------------------------------------------------------------------------
// Everything the commit kernels read or write on the host's behalf lives in one mapped pinned block: no memcpy calls,
// one stream synchronisation per stage.  A single commit takes a recycled block; a batch allocates its own.
struct CommitBlock { void* host = nullptr; void* dev = nullptr; size_t cap = 0; };
struct CommitBlockPool {
    std::mutex mu;
    std::vector<CommitBlock> free_;
    int large_ = 0;
    bool get(size_t need, CommitBlock* out) {
        {
            std::lock_guard<std::mutex> g(mu);
            for (size_t i = 0; i < free_.size(); ++i)
                if (free_[i].cap >= need) { *out = free_[i]; if (free_[i].cap > 8192) --large_; free_.erase(free_.begin() + i); return true; }
        }
        CommitBlock b;
        b.cap = std::max<size_t>(round_up(need, 4096), 4096);
        if (cudaHostAlloc(&b.host, b.cap, cudaHostAllocMapped | cudaHostAllocPortable) != cudaSuccess) { cudaGetLastError(); return false; }
        if (cudaHostGetDevicePointer(&b.dev, b.host, 0) != cudaSuccess) { cudaFreeHost(b.host); cudaGetLastError(); return false; }
        *out = b;
        return true;
    }
    void put(CommitBlock b) {
        // small blocks (single commits) are kept in numbers, large ones (bulk commits) a few: pinning memory costs milliseconds
        { std::lock_guard<std::mutex> g(mu); if (b.cap <= 8192 ? free_.size() < 64 : (b.cap <= (64u << 20) && large_++ < 4)) { free_.push_back(b); return; } }
        cudaFreeHost(b.host);
    }
};
static CommitBlockPool g_commit_blocks;

static uint64_t idx_blob_bytes(const NzIdxResult& r) {
    return sizeof(NzIdxHeader) + (uint64_t)r.n_seg * sizeof(NzIdxSeg) + (uint64_t)r.n_grp * (sizeof(NzIdxGrp) + NZ_IDX_SUBS * 4) +
           round_up(r.side_bytes, 16);
}

int32_t nz_put_commit_many(nz_put** puts, uint32_t n, const uint32_t* expected, const int32_t* has_expected,
                           uint32_t* out_crc, int32_t* out_status) {
    if (!puts || !n) return NZ_EINVAL;
    for (uint32_t i = 0; i < n; ++i) if (!puts[i] || puts[i]->ctx != puts[0]->ctx) return NZ_EINVAL;
    nz_ctx* c = puts[0]->ctx;
    std::vector<int32_t> st(n, NZ_OK);
    std::vector<uint32_t> desc_of(n, 0xFFFFFFFFu);
    std::vector<uint64_t> idx_off(n, 0), idx_len(n, 0), raw_final(n, 0);
    std::vector<uint8_t> inflatable(n, 0);
    for (uint32_t i = 0; i < n; ++i) raw_final[i] = puts[i]->raw_len;
    CommitBlock blk;
    int32_t rc = NZ_OK;
    cudaError_t e = cudaSetDevice(c->device);
    cudaStream_t cs = puts[0]->stream;
    // block layout
    const size_t o_items = 0;
    const size_t o_jobs = o_items + (size_t)n * sizeof(NzPubItem);
    const size_t o_results = o_jobs + (size_t)n * sizeof(NzIdxJob);
    const size_t o_ires = o_results + (size_t)n * sizeof(NzIdxResult);
    const size_t o_crc = o_ires + (size_t)n * sizeof(NzDevRes);
    const size_t o_status = o_crc + round_up((size_t)n * 4, 16);
    const size_t o_trailer = o_status + round_up((size_t)n * 4, 16);
    // batched CRC of the small objects of the batch (<= 64 chunks each): chunk jobs + one fold job per object
    size_t n_cjobs = 0, n_fjobs = 0;
    if (n > 1)
        for (uint32_t i = 0; i < n; ++i) if (puts[i]->n_chunks <= 64) { n_cjobs += puts[i]->n_chunks; ++n_fjobs; }
    const size_t o_cjobs = round_up(o_trailer + (size_t)n * 8, 16);
    const size_t o_fjobs = o_cjobs + n_cjobs * sizeof(NzCrcJob);
    const size_t need = o_fjobs + n_fjobs * sizeof(NzFoldJob);
    NzPubItem* items = nullptr; NzIdxJob* jobs = nullptr; NzIdxResult* results = nullptr; NzDevRes* ires = nullptr;
    uint32_t* crcs = nullptr; int32_t* pst = nullptr; uint32_t* trailers = nullptr;
    auto dev_of = [&](const void* hostp) { return (void*)((char*)blk.dev + ((const char*)hostp - (const char*)blk.host)); };
    if (e != cudaSuccess) { rc = cuda_fail(e, "cudaSetDevice"); goto done; }
    if (!g_commit_blocks.get(need, &blk)) { rc = NZ_ENOMEM; goto done; }
    items = (NzPubItem*)((char*)blk.host + o_items); jobs = (NzIdxJob*)((char*)blk.host + o_jobs);
    results = (NzIdxResult*)((char*)blk.host + o_results); ires = (NzDevRes*)((char*)blk.host + o_ires);
    crcs = (uint32_t*)((char*)blk.host + o_crc); pst = (int32_t*)((char*)blk.host + o_status);
    trailers = (uint32_t*)((char*)blk.host + o_trailer);
    // ---- 1. chunk table + whole-object CRC-32C of every object, each on the stream that carries its fill copies ----
    {
        NzCrcJob* cj = (NzCrcJob*)((char*)blk.host + o_cjobs);
        NzFoldJob* fj = (NzFoldJob*)((char*)blk.host + o_fjobs);
        size_t nc = 0, nf = 0;
        // small objects (<= 64 chunks): two launches per fill STREAM of the batch instead of two per object (the checksum must
        // run behind that object's own copies; nz_put_begin spreads puts over the stream pool)
        std::vector<cudaStream_t> streams;
        if (n_fjobs)
            for (uint32_t i = 0; i < n; ++i)
                if (puts[i]->n_chunks <= 64 && std::find(streams.begin(), streams.end(), puts[i]->stream) == streams.end()) streams.push_back(puts[i]->stream);
        for (cudaStream_t sx : streams) {
            const size_t c0 = nc, f0 = nf;
            for (uint32_t i = 0; i < n; ++i) {
                nz_put* p = puts[i];
                if (p->n_chunks > 64 || p->stream != sx) continue;
                uint32_t* tab = reinterpret_cast<uint32_t*>(c->arena + p->crc_tab_off);
                for (uint32_t k = 0; k < p->n_chunks; ++k) {
                    const uint64_t off = (uint64_t)k * NZ_CRC_CHUNK;
                    cj[nc++] = NzCrcJob{(uint64_t)(c->arena + p->arena_off + off), (uint32_t)std::min<uint64_t>(NZ_CRC_CHUNK, p->stored_len - off), 0, (uint64_t)(tab + k)};
                }
                crcs[i] = 0;
                fj[nf++] = NzFoldJob{(uint64_t)tab, p->stored_len, (uint64_t)dev_of(&crcs[i])};
            }
            e = nz_launch_crc_jobs((const NzCrcJob*)dev_of(cj + c0), (uint32_t)(nc - c0), (const NzFoldJob*)dev_of(fj + f0), (uint32_t)(nf - f0), c->crc_tables, c->sm_count, sx);
            c->launches += 2;
            if (e != cudaSuccess) { rc = cuda_fail(e, "batched crc kernels"); goto done; }
        }
        for (uint32_t i = 0; i < n; ++i) {
            nz_put* p = puts[i];
            if (n_fjobs && p->n_chunks <= 64) continue;
            crcs[i] = 0;
            uint32_t* tab = reinterpret_cast<uint32_t*>(c->arena + p->crc_tab_off);
            e = nz_launch_crc_chunks(c->arena + p->arena_off, p->stored_len, tab, (uint32_t*)dev_of(&crcs[i]), c->crc_tables, c->sm_count, p->stream);
            c->launches += 2;
            if (e != cudaSuccess) { rc = cuda_fail(e, "crc kernels"); goto done; }
        }
    }
    {
        std::vector<cudaStream_t> seen;
        for (uint32_t i = 0; i < n; ++i) {
            if (std::find(seen.begin(), seen.end(), puts[i]->stream) != seen.end()) continue;
            seen.push_back(puts[i]->stream);
            e = cudaStreamSynchronize(puts[i]->stream);
            if (e != cudaSuccess) { rc = cuda_fail(e, "crc sync"); goto done; }
        }
    }
    for (uint32_t i = 0; i < n; ++i) {
        if (out_crc) out_crc[i] = crcs[i];
        const bool has = expected && (!has_expected || has_expected[i]);
        if (has && crcs[i] != expected[i]) { st[i] = NZ_ECORRUPT; c->corrupt++; }
    }
    // ---- 2. gzip members: inflate once, check trailer CRC-32 and ISIZE, record the checkpoint index ----
    if (c->idx_build) {
        std::vector<uint32_t> cand;
        for (uint32_t i = 0; i < n; ++i) {
            nz_put* p = puts[i];
            if (st[i] != NZ_OK || !(p->flags & NZ_ENC_GZIP) || (p->flags & NZ_ENC_STRIPE) || p->stored_len < 18 || p->stored_len >= (1ull << 32)) continue;
            e = cudaMemcpyAsync(&trailers[2 * i], c->arena + p->arena_off + p->stored_len - 8, 8, cudaMemcpyDeviceToHost, cs);
            if (e != cudaSuccess) { rc = cuda_fail(e, "trailer read"); goto done; }
            cand.push_back(i);
        }
        if (!cand.empty()) {
            e = cudaStreamSynchronize(cs);
            if (e != cudaSuccess) { rc = cuda_fail(e, "trailer sync"); goto done; }
        }
        static const uint64_t kBudget = [] { const char* ev = getenv("NZ_INDEX_TMP_MIB"); return (ev && atoll(ev) > 0 ? (uint64_t)atoll(ev) : 3072ull) << 20; }();
        size_t ci = 0;
        while (ci < cand.size()) {
            // one round: as many members as fit the temporary budget
            struct Carve { uint64_t out, seg, grp, delta, side; };
            std::vector<uint32_t> round;
            std::vector<Carve> carve;
            uint64_t total = 4096;      // counters | then per round: infl | meta
            for (; ci < cand.size(); ++ci) {
                const uint32_t i = cand[ci];
                nz_put* p = puts[i];
                const uint64_t isize = trailers[2 * i + 1];
                if (isize > p->stored_len * 1032ull + 1024 || isize >= 0xFFFFFF00ull) continue;      // not a deflate ratio: no index (GETs decline it)
                const uint64_t seg_cap = isize / c->idx_seg_bytes + 2, grp_cap = isize / 1024 + p->stored_len / 1024 + 16;
                const uint64_t side_cap = isize / 4 + 4096;
                const uint64_t bytes = round_up(isize + 256, 256) + round_up(seg_cap * sizeof(NzIdxSeg), 256) + round_up(grp_cap * sizeof(NzIdxGrp), 256) +
                                       round_up(grp_cap * NZ_IDX_SUBS * 4, 256) + round_up(side_cap, 256) + 256;
                if (bytes > kBudget) continue;                                   // too large to index in one piece: legacy decode
                if (!round.empty() && total + bytes > kBudget) break;
                Carve cv;
                cv.out = total; total += round_up(isize + 256, 256);
                cv.seg = total; total += round_up(seg_cap * sizeof(NzIdxSeg), 256);
                cv.grp = total; total += round_up(grp_cap * sizeof(NzIdxGrp), 256);
                cv.delta = total; total += round_up(grp_cap * NZ_IDX_SUBS * 4, 256);
                cv.side = total; total += round_up(side_cap, 256);
                round.push_back(i);
                carve.push_back(cv);
            }
            const uint32_t m = (uint32_t)round.size();
            if (!m) continue;
            const uint64_t o_infl = total; total += round_up((uint64_t)m * sizeof(NzInflWork), 256);
            const uint64_t o_meta = total; total += round_up((uint64_t)m * 8, 256);
            char* tmp = nullptr;
            if (cudaMallocAsync((void**)&tmp, total, cs) != cudaSuccess) { cudaGetLastError(); continue; }    // no room: these members go unindexed
            for (uint32_t k = 0; k < m; ++k) {
                const uint32_t i = round[k];
                nz_put* p = puts[i];
                const uint64_t isize = trailers[2 * i + 1];
                NzIdxJob& j = jobs[k];
                j.src = (uint64_t)(c->arena + p->arena_off); j.src_len = p->stored_len;
                j.out = (uint64_t)(tmp + carve[k].out); j.out_cap = isize;
                j.seg = (uint64_t)(tmp + carve[k].seg); j.grp = (uint64_t)(tmp + carve[k].grp);
                j.delta = (uint64_t)(tmp + carve[k].delta); j.side = (uint64_t)(tmp + carve[k].side);
                j.side_cap = isize / 4 + 4096;
                j.seg_cap = (uint32_t)(isize / c->idx_seg_bytes + 2); j.grp_cap = (uint32_t)(isize / 1024 + p->stored_len / 1024 + 16);
                j.seg_bytes = c->idx_seg_bytes; j.pad = 0; j.dest = 0;
                memset(&results[k], 0, sizeof(NzIdxResult));
                results[k].status = -1;
                memset(&ires[k], 0, sizeof(NzDevRes));
            }
            e = nz_launch_index((const NzIdxJob*)dev_of(jobs), (NzIdxResult*)dev_of(results), m, (uint32_t*)tmp, (NzInflWork*)(tmp + o_infl),
                                (uint64_t*)(tmp + o_meta), (NzDevRes*)dev_of(ires), c->crc_tables, c->sm_count, cs);
            c->launches += 2;
            if (e == cudaSuccess) e = cudaStreamSynchronize(cs);
            if (e != cudaSuccess) { rc = cuda_fail(e, "index build"); cudaFreeAsync(tmp, cs); goto done; }
            for (uint32_t k = 0; k < m; ++k) {
                const uint32_t i = round[k];
                const NzIdxResult& r = results[k];
                if (r.status < 0 || ires[k].status != NZ_OK) continue;          // malformed member or CRC-32 mismatch: stored as is, inflate GETs decline it (SPEC G4)
                raw_final[i] = r.raw_len;                                       // the member's real inflated length
                inflatable[i] = 1;
                if (r.status != 0) continue;                                    // valid, but too fragmented to index within the caps
                const uint64_t bytes = idx_blob_bytes(r);
                uint64_t off = 0;
                {
                    std::lock_guard<std::mutex> g(c->mu);
                    if (!arena_alloc(c, bytes, &off)) continue;                 // arena full: serve it through the whole-member decode
                }
                idx_off[i] = off; idx_len[i] = bytes;
                jobs[k].dest = (uint64_t)(c->arena + off);
            }
            e = nz_launch_index_pack((const NzIdxJob*)dev_of(jobs), (const NzIdxResult*)dev_of(results), m, cs);
            c->launches += 1;
            if (e == cudaSuccess) e = cudaFreeAsync(tmp, cs);
            if (e == cudaSuccess) e = cudaStreamSynchronize(cs);                // the jobs array (pinned) is reused by the next round
            if (e != cudaSuccess) { rc = cuda_fail(e, "index pack"); goto done; }
        }
    }
    // ---- 3. publish: descriptors + table slots in one launch ----
    {
        uint32_t m = 0;
        std::vector<uint32_t> who;
        {
            std::lock_guard<std::mutex> g(c->mu);
            for (uint32_t i = 0; i < n; ++i) {
                if (st[i] != NZ_OK) continue;
                if (c->free_desc.empty()) { st[i] = NZ_ENOMEM; continue; }
                desc_of[i] = c->free_desc.back();
                c->free_desc.pop_back();
                nz_put* p = puts[i];
                NzPubItem& it = items[m];
                memset(&it, 0, sizeof it);
                it.k0 = p->k0; it.k1 = p->k1; it.gen = c->gen++; it.desc_idx = desc_of[i];
                it.d.arena_off = p->arena_off; it.d.stored_len = p->stored_len; it.d.raw_len = raw_final[i]; it.d.user_meta = p->user_meta;
                it.d.crc_tab_off = p->crc_tab_off; it.d.flags = p->flags | (inflatable[i] ? NZ_ENC_INFLATABLE : 0u); it.d.crc32c = crcs[i]; it.d.n_chunks = p->n_chunks;
                it.d.gen = (uint32_t)it.gen; it.d.stripe_bytes = idx_off[i];
                pst[m] = NZ_ECUDA;
                who.push_back(i);
                ++m;
            }
        }
        if (m) {
            e = nz_launch_publish(c->table, c->n_lines - 1, c->descs, (const NzPubItem*)dev_of(items), m, (int32_t*)dev_of(pst), cs);
            c->launches += 1;
            if (e == cudaSuccess) e = cudaStreamSynchronize(cs);
            if (e != cudaSuccess) { rc = cuda_fail(e, "publish"); goto done; }
            for (uint32_t k = 0; k < m; ++k) st[who[k]] = pst[k];
        }
    }
done:
    if (rc != NZ_OK) for (uint32_t i = 0; i < n; ++i) if (st[i] == NZ_OK) st[i] = rc;
    if (rc != NZ_OK) for (uint32_t i = 0; i < n; ++i) cudaStreamSynchronize(puts[i]->stream);
    {
        std::lock_guard<std::mutex> g(c->mu);
        for (uint32_t i = 0; i < n; ++i) {
            nz_put* p = puts[i];
            if (st[i] == NZ_OK) {
                NzHostObj& ho = c->shadow[{p->k0, p->k1}];
                ho.desc = desc_of[i]; ho.crc_whole = crcs ? crcs[i] : 0; ho.crc_tab_off = p->crc_tab_off;
                ho.raw_len = raw_final[i]; ho.idx_off = idx_off[i]; ho.idx_len = idx_len[i];
                if (inflatable[i]) ho.flags |= NZ_ENC_INFLATABLE;
                c->index_bytes += idx_len[i];
            } else {
                if (desc_of[i] != 0xFFFFFFFFu) c->free_desc.push_back(desc_of[i]);
                if (idx_len[i]) arena_free(c, idx_off[i], idx_len[i]);
            }
        }
    }
    for (uint32_t i = 0; i < n; ++i) {
        if (st[i] == NZ_OK) { c->puts++; c->put_bytes += puts[i]->stored_len; }
        if (out_status) out_status[i] = st[i];
        put_release(puts[i], st[i] != NZ_OK);
    }
    if (blk.host) g_commit_blocks.put(blk);
    return rc;
}

// Bulk fill of many (small) identity objects that sit in ONE pinned window: one host->device copy of the span they
// cover, one scatter launch (the K2 gather kernel, staging -> extents), then nz_put_commit_many (batched CRC-32C, one
// publish launch).  The per-object cudaMemcpyAsync of nz_put_write (~2-3 us each) is what bounded 10^6-object loads.
int32_t nz_put_packed(nz_ctx* c, uint32_t n, const uint8_t (*keys)[16], const uint64_t* src_off, const uint64_t* len, nz_buf* window,
                      const uint64_t* user_meta, uint32_t* out_crc, int32_t* out_status) {
    if (!c || !n || !keys || !src_off || !len || !window || !window->host) return NZ_EINVAL;
    CK(cudaSetDevice(c->device));
    uint64_t lo = ~0ull, hi = 0;
    for (uint32_t i = 0; i < n; ++i) {
        if (src_off[i] > window->size || len[i] > window->size - src_off[i]) return NZ_EINVAL;
        if (!len[i]) continue;
        lo = std::min(lo, src_off[i]);
        hi = std::max(hi, src_off[i] + len[i]);
    }
    if (lo == ~0ull) { lo = 0; hi = 0; }
    lo &= ~255ull;                                    // keep the 16 B phase of every object: staging offset == window offset - lo
    std::vector<nz_put*> puts(n, nullptr);
    std::vector<int32_t> st(n, NZ_OK);
    uint32_t made = 0;
    int32_t rc = NZ_OK;
    for (; made < n; ++made) {
        rc = nz_put_begin(c, keys[made], len[made], len[made], NZ_ENC_IDENTITY, user_meta ? user_meta[made] : 0, &puts[made]);
        if (rc != NZ_OK) break;
    }
    if (rc != NZ_OK) {
        for (uint32_t i = 0; i < made; ++i) nz_put_abort(puts[i]);
        return rc;
    }
    cudaStream_t sx = puts[0]->stream;
    for (uint32_t i = 0; i < n; ++i) puts[i]->stream = sx;           // one stream carries the copy, the scatter and the checksums
    char* stage = nullptr;
    CommitBlock blk;
    const size_t o_work = 0, o_tiles = round_up((size_t)n * sizeof(NzSegWork), 256), o_head = o_tiles + round_up((size_t)(n + 1) * 8, 256);
    const size_t need = o_head + round_up((size_t)n * 8, 256);
    cudaError_t e = cudaSuccess;
    const uint32_t tb = nz_gather_tile_bytes(c->gather_variant);
    if (hi > lo) {
        if (!g_commit_blocks.get(need, &blk)) rc = NZ_ENOMEM;
        const size_t span = round_up((size_t)(hi - lo), 256);
        if (rc == NZ_OK && (e = cudaMallocAsync((void**)&stage, span + need, sx)) != cudaSuccess) { cudaGetLastError(); rc = NZ_ENOMEM; }
        if (rc == NZ_OK) {
            NzSegWork* w = (NzSegWork*)((char*)blk.host + o_work);
            uint64_t* tbeg = (uint64_t*)((char*)blk.host + o_tiles);
            uint32_t* head = (uint32_t*)((char*)blk.host + o_head);
            uint64_t tiles = 0;
            for (uint32_t i = 0; i < n; ++i) {
                w[i].src = (uint64_t)(stage + (src_off[i] - lo));
                w[i].dst = (uint64_t)(c->arena + puts[i]->arena_off);      // 4 KiB aligned: vstart 0
                w[i].len = len[i];
                w[i].vstart = 0;
                tbeg[i] = tiles;
                tiles += (len[i] + tb - 1) / tb;
                head[i] = 0xFFFFFFFFu;
                head[n + i] = 0xFFFFFFFFu;
            }
            tbeg[n] = tiles;
            char* lists = stage + span;                               // the work lists travel with the data: the gather binary-searches them per tile
            e = cudaMemcpyAsync(stage, (const char*)window->host + lo, hi - lo, cudaMemcpyHostToDevice, sx);
            if (e == cudaSuccess) e = cudaMemcpyAsync(lists, blk.host, need, cudaMemcpyHostToDevice, sx);
            if (e == cudaSuccess && tiles) {
                const int grid = (int)std::min<uint64_t>(tiles, (uint64_t)c->gather_grid);
                e = nz_launch_gather(c->gather_variant, (const NzSegWork*)(lists + o_work), (const uint64_t*)(lists + o_tiles), (const uint32_t*)(lists + o_head),
                                     (const uint32_t*)(lists + o_head) + n, n, grid, sx);
                c->launches += 1;
            }
            if (e == cudaSuccess) e = cudaFreeAsync(stage, sx);
            if (e != cudaSuccess) rc = cuda_fail(e, "nz_put_packed");
        }
    }
    if (rc != NZ_OK) {
        cudaStreamSynchronize(sx);
        for (uint32_t i = 0; i < n; ++i) nz_put_abort(puts[i]);
        if (blk.host) g_commit_blocks.put(blk);
        return rc;
    }
    rc = nz_put_commit_many(puts.data(), n, nullptr, nullptr, out_crc, out_status);      // syncs the stream: the work list may go after it
    if (blk.host) g_commit_blocks.put(blk);
    return rc;
}

int32_t nz_put_commit(nz_put* p, uint32_t expected, int32_t has_expected, uint32_t* out_crc) {
    if (!p) return NZ_EINVAL;
    int32_t st = NZ_OK;
    const int32_t rc = nz_put_commit_many(&p, 1, &expected, &has_expected, out_crc, &st);
    return rc != NZ_OK ? rc : st;
}

int32_t nz_evict(nz_ctx* c, const uint8_t key[16]) {
    if (!c || !key) return NZ_EINVAL;
    CK(cudaSetDevice(c->device));
    const uint64_t k0 = rd64(key), k1 = rd64(key + 8);
    NzHostObj ho;
    {
        std::lock_guard<std::mutex> g(c->mu);
        auto it = c->shadow.find({k0, k1});
        if (it == c->shadow.end() || it->second.desc == 0xFFFFFFFFu) return NZ_ENOTFOUND;
        ho = it->second;
        c->shadow.erase(it);
        if (ho.stripe_bytes && c->n_striped) c->n_striped--;
    }
    Mailbox* mb = g_mail.get();
    if (!mb) return NZ_ENOMEM;
    void* mb_dev = nullptr;
    CK(cudaHostGetDevicePointer(&mb_dev, mb, 0));
    cudaError_t e = nz_launch_erase(c->table, c->n_lines - 1, k0, k1, &((Mailbox*)mb_dev)->status, c->put_stream);
    c->launches += 1;
    if (e != cudaSuccess) { g_mail.put(mb); return cuda_fail(e, "erase"); }
    e = cudaStreamSynchronize(c->put_stream);
    int32_t rc = mb->status;
    g_mail.put(mb);
    if (e != cudaSuccess) return cuda_fail(e, "erase sync");
    // The slot is a tombstone now, so no wave submitted from here on can resolve the key.  Waves submitted BEFORE may have
    // probed it while it was live and may still be reading the extent: the bytes (and the descriptor) are handed back
    // only after every one of those has finished.  Later waves are not waited for.
    {
        uint64_t barrier;
        { std::lock_guard<std::mutex> g(c->tmu); barrier = c->submit_seq; }
        const auto t0 = std::chrono::steady_clock::now();
        for (;;) {
            bool busy = false;
            {
                std::lock_guard<std::mutex> g(c->tmu);
                for (auto it = c->inflight.begin(); it != c->inflight.end() && it->first <= barrier; ++it) {
                    const cudaError_t q = cudaEventQuery(it->second->ev_done);
                    if (q == cudaErrorNotReady) { busy = true; break; }
                    if (q != cudaSuccess) cudaGetLastError();
                }
            }
            if (!busy) break;
            if (std::chrono::steady_clock::now() - t0 > std::chrono::seconds(5)) { cudaDeviceSynchronize(); break; }
            std::this_thread::sleep_for(std::chrono::microseconds(20));
        }
    }
    {
        std::lock_guard<std::mutex> g(c->mu);
        arena_free(c, ho.arena_off, ho.arena_len);
        if (ho.idx_len) { arena_free(c, ho.idx_off, ho.idx_len); c->index_bytes -= ho.idx_len; }
        c->free_desc.push_back(ho.desc);
    }
    c->evictions++;
    return rc;
}

// ---------------------------------------------------------------------------
// K1 lookup
// ---------------------------------------------------------------------------
int32_t nz_lookup(nz_ctx* c, const uint8_t (*keys)[16], uint32_t n, nz_extent* out) {
    if (!c || (n && (!keys || !out))) return NZ_EINVAL;
    if (!n) return NZ_OK;
    CK(cudaSetDevice(c->device));
    const size_t kb = (size_t)n * 16, ob = (size_t)n * sizeof(nz_extent);
    void* pin = nullptr;
    void* pin_dev = nullptr;
    CK(cudaHostAlloc(&pin, round_up(kb, 256) + ob, cudaHostAllocMapped | cudaHostAllocPortable));
    cudaError_t e = cudaHostGetDevicePointer(&pin_dev, pin, 0);
    if (e != cudaSuccess) { cudaFreeHost(pin); return cuda_fail(e, "cudaHostGetDevicePointer"); }
    memcpy(pin, keys, kb);
    cudaStream_t s = c->streams[c->rr.fetch_add(1) % c->streams.size()];
    nz_extent* dout = reinterpret_cast<nz_extent*>((char*)pin_dev + round_up(kb, 256));
    e = nz_launch_lookup(c->table, c->n_lines - 1, c->descs, (const uint64_t*)pin_dev, n, dout, s);
    c->launches += 1;
    if (e == cudaSuccess) e = cudaStreamSynchronize(s);
    if (e != cudaSuccess) { cudaFreeHost(pin); return cuda_fail(e, "lookup"); }
    memcpy(out, (char*)pin + round_up(kb, 256), ob);
    cudaFreeHost(pin);
    return NZ_OK;
}

// ---------------------------------------------------------------------------
// GET waves
// ---------------------------------------------------------------------------
static nz_ticket* ticket_acquire(nz_ctx* c, size_t pin_need, size_t dev_need) {
    nz_ticket* t = nullptr;
    {
        std::lock_guard<std::mutex> g(c->tmu);
        for (size_t i = 0; i < c->ticket_pool.size(); ++i) {
            if (c->ticket_pool[i]->pin_cap >= pin_need && c->ticket_pool[i]->dev_cap >= dev_need) {
                t = c->ticket_pool[i];
                c->ticket_pool.erase(c->ticket_pool.begin() + i);
                break;
            }
        }
        if (!t && !c->ticket_pool.empty()) { t = c->ticket_pool.back(); c->ticket_pool.pop_back(); }
    }
    if (!t) {
        t = new (std::nothrow) nz_ticket();
        if (!t) return nullptr;
        t->ctx = c;
        if (cudaEventCreate(&t->ev_start) != cudaSuccess || cudaEventCreate(&t->ev_done) != cudaSuccess ||
            cudaEventCreate(&t->ev_g0) != cudaSuccess || cudaEventCreate(&t->ev_g1) != cudaSuccess ||
            cudaEventCreate(&t->ev_i0) != cudaSuccess || cudaEventCreate(&t->ev_i1) != cudaSuccess) { ticket_destroy(t); return nullptr; }
    }
    if (t->pin_cap < pin_need) {
        if (t->pin) cudaFreeHost(t->pin);
        t->pin = nullptr; t->pin_cap = 0;
        size_t cap = std::max<size_t>(round_up(pin_need, 4096), 64 << 10);
        if (cudaHostAlloc(&t->pin, cap, cudaHostAllocMapped | cudaHostAllocPortable) != cudaSuccess ||
            cudaHostGetDevicePointer(&t->pin_dev, t->pin, 0) != cudaSuccess) { ticket_destroy(t); return nullptr; }
        t->pin_cap = cap;
    }
    if (t->dev_cap < dev_need) {
        if (t->dev) cudaFree(t->dev);
        t->dev = nullptr; t->dev_cap = 0;
        size_t cap = std::max<size_t>(round_up(dev_need, 4096), 64 << 10);
        if (cudaMalloc(&t->dev, cap) != cudaSuccess) { ticket_destroy(t); return nullptr; }
        t->dev_cap = cap;
    }
    return t;
}

// K5: a GET of a striped object is split at stripe boundaries on the host (the shadow knows the stripe
// size), so on the device every slice lies inside exactly one stripe.  `split` stays empty when the zone
// holds no striped object; split[i] stays empty for requests of ordinary objects.
static void split_striped_gets(nz_ctx* c, const nz_get_req* reqs, uint32_t n, std::vector<std::vector<nz_seg>>& split) {
    if (c->n_striped.load() == 0) return;
    split.resize(n);
    std::lock_guard<std::mutex> g(c->mu);
    for (uint32_t i = 0; i < n; ++i) {
        const nz_get_req& r = reqs[i];
        auto it = c->shadow.find({rd64(r.key), rd64(r.key + 8)});
        if (it == c->shadow.end() || !it->second.stripe_bytes || it->second.desc == 0xFFFFFFFFu) continue;
        const uint64_t S = it->second.stripe_bytes, L = it->second.stored_len;
        nz_seg whole{0, L, 0};
        const nz_seg* src = r.nseg ? r.segs : &whole;
        const uint32_t cnt = r.nseg ? r.nseg : 1;
        for (uint32_t k = 0; k < cnt; ++k) {
            uint64_t off = src[k].src_off, len = src[k].len, dst = src[k].dst_off;
            if (off > L || len > L - off) { split[i].clear(); split[i].push_back(src[k]); break; }   // device reports EINVAL
            while (len) {
                const uint64_t room = S - off % S, take = len < room ? len : room;
                split[i].push_back(nz_seg{off, take, dst});
                off += take; dst += take; len -= take;
            }
        }
        if (split[i].empty()) split[i].push_back(nz_seg{0, 0, 0});   // empty object / zero-length request
    }
}

// Staged D2H: one copy-engine transfer per slice, from its place in the staging slot to its pinned destination.
// Returns false (direct stores from the gather kernel instead) when a request has no explicit slices, a slice does
// not fit its destination (the device path reports NZ_ERANGE), or the wave is outside 8 MiB .. 1 GiB.
struct DmaCopy { void* host; uint64_t off, len; };
static bool plan_staged_copies(const nz_get_req* reqs, uint32_t n, uint64_t nseg, std::vector<DmaCopy>& copies, uint64_t& bytes) {
    copies.reserve(nseg);
    bytes = 0;
    for (uint32_t i = 0; i < n; ++i) {
        const nz_get_req& r = reqs[i];
        if (r.nseg == 0 || !r.dst->host) return false;
        const uint64_t cap = r.dst->size - r.dst_base;
        for (uint32_t k = 0; k < r.nseg; ++k) {
            const nz_seg& sg = r.segs[k];
            if (sg.dst_off > cap || sg.len > cap - sg.dst_off) return false;
            copies.push_back(DmaCopy{(char*)r.dst->host + r.dst_base + sg.dst_off, bytes, sg.len});
            bytes += round_up(sg.len, 256);
        }
    }
    return bytes >= (8ull << 20) && bytes <= (1ull << 30);
}

// Verify work bound + inflate scratch sizing come from the host shadow (metadata only; the data path itself
// resolves keys on the device).  scratch_need is empty when no request asks for inflate.
// Inflate dedupe: the first inflate request of every distinct gzip key of the wave is its LEADER (leader[i] == i);
// the others follow it (leader[i] = the leader's index) and are served by a copy out of the leader's scratch.
static void size_verify_and_scratch(nz_ctx* c, const nz_get_req* reqs, uint32_t n, uint64_t& ver_cap, std::vector<uint64_t>& scratch_need,
                                    std::vector<uint32_t>& leader) {
    std::unordered_map<std::pair<uint64_t, uint64_t>, uint32_t, KeyHash> first;
    std::vector<uint8_t> followed(scratch_need.empty() ? 0 : n, 0);
    std::lock_guard<std::mutex> g(c->mu);
    for (uint32_t i = 0; i < n; ++i) {
        const nz_get_req& r = reqs[i];
        if (!(r.flags & (NZ_GET_INFLATE | NZ_GET_VERIFY))) continue;
        const std::pair<uint64_t, uint64_t> key{rd64(r.key), rd64(r.key + 8)};
        auto it = c->shadow.find(key);
        if (it == c->shadow.end() || it->second.desc == 0xFFFFFFFFu) continue;
        const NzHostObj& ho = it->second;
        const bool inflate = (r.flags & NZ_GET_INFLATE) && (ho.flags & NZ_ENC_GZIP);
        if (r.flags & NZ_GET_VERIFY) {
            const uint64_t all = (ho.stored_len + NZ_CRC_CHUNK - 1) / NZ_CRC_CHUNK;
            if (ho.stripe_bytes) {
                // striped: the GET is split at stripe boundaries and every piece is checked against its stripe's own chunk
                // table (on the GPU that holds the stripe): up to two partial chunks per piece
                auto bound = [&](uint64_t len) { return len / NZ_CRC_CHUNK + 2 + 2 * (len / ho.stripe_bytes + 2); };
                if (r.nseg == 0) ver_cap += bound(ho.stored_len);
                else for (uint32_t k = 0; k < r.nseg; ++k) ver_cap += bound(r.segs[k].len);
            }
            else if (inflate || r.nseg == 0) ver_cap += all;
            else for (uint32_t k = 0; k < r.nseg; ++k) ver_cap += std::min<uint64_t>(all, r.segs[k].len / NZ_CRC_CHUNK + 2);
        }
        if (!inflate || scratch_need.empty()) continue;
        auto ins = c->infl_dedupe ? first.emplace(key, i) : std::make_pair(first.end(), true);
        if (!ins.second) {                                    // a follower: its leader inflates into scratch
            const uint32_t L = ins.first->second;
            leader[i] = L;
            if (!followed[L]) { followed[L] = 1; scratch_need[L] = ho.raw_len; }
            continue;
        }
        // the LZ77 window is the output itself: it must live in HBM, so ranges and host
        // destinations inflate into device scratch first and are then gathered out
        if (r.nseg != 0 || r.dst->kind != NZ_BUF_DEVICE) scratch_need[i] = ho.raw_len;
    }
}

int32_t nz_get_submit(nz_ctx* c, const nz_get_req* reqs, uint32_t n, void* cuda_stream, nz_ticket** out) {
    if (!c || !reqs || !n || !out) return NZ_EINVAL;
    CK(cudaSetDevice(c->device));
    // ---- host-side sizing ----
    uint64_t nseg = 0;
    bool any_inflate = false, any_verify = false, any_urgent = false, all_host = true;
    std::vector<std::vector<nz_seg>> split;
    split_striped_gets(c, reqs, n, split);
    for (uint32_t i = 0; i < n; ++i) {
        const nz_get_req& r = reqs[i];
        if (!r.dst || (r.nseg && !r.segs) || r.dst_base > r.dst->size) return NZ_EINVAL;
        if (!split.empty() && !split[i].empty()) nseg += split[i].size();
        else nseg += r.nseg ? r.nseg : 1;
        any_inflate |= (r.flags & NZ_GET_INFLATE) != 0;
        any_verify |= (r.flags & NZ_GET_VERIFY) != 0;
        any_urgent |= (r.flags & NZ_GET_URGENT) != 0;
        all_host &= r.dst->kind == NZ_BUF_PINNED_HOST;
    }
    if (nseg > 0x7fffffffull) return NZ_EINVAL;
    // ---- staged D2H (see nz_ctx::stage): explicit slices only, unpaced, not latency critical, 8 MiB .. 1 GiB ----
    std::vector<DmaCopy> dma_copies;
    uint64_t dma_bytes = 0;
    const bool dma = all_host && c->host_dma && !any_urgent && !any_inflate && split.empty() && c->host_rate_gbs.load() <= 0 &&
                     plan_staged_copies(reqs, n, nseg, dma_copies, dma_bytes);
    uint64_t ver_cap = 0;
    std::vector<uint64_t> scratch_need(any_inflate ? n : 0, 0);
    std::vector<uint32_t> leader(any_inflate ? n : 0);
    for (uint32_t i = 0; i < (uint32_t)leader.size(); ++i) leader[i] = i;
    if (any_inflate || any_verify) size_verify_and_scratch(c, reqs, n, ver_cap, scratch_need, leader);
    if (ver_cap > 0xfffffff0ull) return NZ_EINVAL;

    const size_t pin_reqs = round_up((size_t)n * sizeof(NzDevReq), 256);
    const size_t pin_segs = round_up((size_t)nseg * sizeof(NzDevSeg), 256);
    const size_t pin_res = round_up((size_t)n * sizeof(NzDevRes) + 32, 256);     // + {verify chunks, of which fused, completion flag}
    const size_t dev_work = round_up((size_t)nseg * sizeof(NzSegWork), 256);
    const size_t dev_tiles = round_up((size_t)(nseg + 1) * 8, 256);
    const size_t dev_chain = round_up((size_t)nseg * 8, 256);      // chain_head | chain_next (u32 each)
    const size_t dev_ver = round_up((size_t)ver_cap * sizeof(NzVerWork), 256);
    const size_t dev_cnt = 256;
    const size_t dev_infl = (size_t)n * sizeof(NzInflWork);      // followed directly by rq_src, rq_len
    const size_t dev_rq = (size_t)n * 48;                        // rq_src, rq_len, infl_meta, rq_tab, rq_ssz, rq_stat(+pad)
    const size_t dev_stage = round_up((size_t)n * sizeof(NzDevReq), 256) + round_up((size_t)nseg * sizeof(NzDevSeg), 256);
    const size_t dev_segver = any_verify ? round_up((size_t)nseg * 4, 256) : 0;
    const size_t dev_iseg = any_inflate ? round_up((size_t)(n + 1) * 4, 256) : 0;          // inflate members' segment prefix
    uint32_t coal_cap = 0;                                                                 // request-coalescing hash table for waves beyond 2048 slices
    if (c->coalesce && nseg > 2048) { coal_cap = 8192; while (coal_cap < 2 * nseg) coal_cap <<= 1; }
    const size_t dev_coal = (size_t)coal_cap * 4;
    nz_ticket* t = ticket_acquire(c, pin_reqs + pin_segs + pin_res, dev_work + dev_tiles + dev_chain + dev_ver + dev_cnt + round_up(dev_infl + dev_rq, 256) + dev_stage + dev_segver + dev_iseg + dev_coal);
    if (!t) { cudaGetLastError(); snprintf(tl_err, sizeof tl_err, "ticket allocation failed"); return NZ_ENOMEM; }
    t->n = n; t->nseg = (uint32_t)nseg;

    NzDevReq* hreq = reinterpret_cast<NzDevReq*>(t->pin);
    NzDevSeg* hseg = reinterpret_cast<NzDevSeg*>((char*)t->pin + pin_reqs);
    NzDevRes* hres = reinterpret_cast<NzDevRes*>((char*)t->pin + pin_reqs + pin_segs);
    if (any_inflate) {
        size_t need = 0;
        for (uint32_t i = 0; i < n; ++i) if (scratch_need[i]) need += round_up(scratch_need[i] + 256, 256);
        if (need > t->scratch_cap) {
            if (t->scratch) cudaFree(t->scratch);
            t->scratch = nullptr; t->scratch_cap = 0;
            // with headroom: the next wave's distinct members differ a little, and growing again means cudaFree (which waits
            // for every wave in flight) + cudaMalloc on the submit path
            size_t cap = round_up(need + need / 8, 1ull << 20);     // an eighth more, MiB granular (a one-member wave keeps a small scratch)
            if (cudaMalloc(&t->scratch, cap) != cudaSuccess) {
                cudaGetLastError();
                cap = need;
                if (cudaMalloc(&t->scratch, cap) != cudaSuccess) {
                    cudaGetLastError();
                    nz_ticket_free(t);
                    snprintf(tl_err, sizeof tl_err, "inflate scratch allocation of %zu bytes failed", need);
                    return NZ_ENOMEM;
                }
            }
            t->scratch_cap = cap;
        }
    }
    size_t scratch_off = 0;
    uint32_t sb = 0;
    for (uint32_t i = 0; i < n; ++i) {
        const nz_get_req& r = reqs[i];
        NzDevReq& d = hreq[i];
        d.k0 = rd64(r.key); d.k1 = rd64(r.key + 8);
        d.dst = (uint64_t)(uintptr_t)r.dst->dev + r.dst_base;
        d.dst_cap = r.dst->size - r.dst_base;
        if (dma) { d.dst = 0; d.dst_cap = dma_bytes; }          // + staging base, patched once the slot is known
        d.flags = r.flags;
        d.seg_begin = sb;
        const bool striped = !split.empty() && !split[i].empty();
        d.whole = r.nseg == 0 && !striped;
        d.nseg = striped ? (uint32_t)split[i].size() : (r.nseg ? r.nseg : 1);
        d.scratch = 0; d.scratch_cap = 0;
        d.infl_leader = any_inflate ? leader[i] : i; d.pad0 = 0; d.pad1 = 0;
        if (any_inflate && scratch_need[i]) {
            d.scratch = (uint64_t)(uintptr_t)((char*)t->scratch + scratch_off);
            d.scratch_cap = scratch_need[i];
            scratch_off += round_up(scratch_need[i] + 256, 256);
        }
        if (striped) {
            for (const nz_seg& sg : split[i]) {
                hseg[sb].src_off = sg.src_off; hseg[sb].len = sg.len; hseg[sb].dst_off = sg.dst_off;
                hseg[sb].req = i; hseg[sb].pad = 0;
                ++sb;
            }
        } else if (r.nseg == 0) {
            hseg[sb].src_off = 0; hseg[sb].len = 0; hseg[sb].dst_off = 0; hseg[sb].req = i; hseg[sb].pad = 0;
            ++sb;
        } else {
            for (uint32_t k = 0; k < r.nseg; ++k, ++sb) {
                hseg[sb].src_off = r.segs[k].src_off; hseg[sb].len = r.segs[k].len; hseg[sb].dst_off = r.segs[k].dst_off;
                if (dma) hseg[sb].dst_off = dma_copies[sb].off;  // slices land back to back in the staging slot
                hseg[sb].req = i; hseg[sb].pad = 0;
            }
        }
        hres[i].status = NZ_ECUDA; hres[i].flags = 0; hres[i].bytes = 0; hres[i].length = 0; hres[i].crc32c = 0; hres[i].reserved = 0;
    }

    // waves that only write pinned host memory may use their own gather variant (write-through stores)
    const bool host_kernel = all_host && !dma;      // the gather kernel itself stores across PCIe
    const int variant = c->gather_variant;
    // upper bound on gather tiles when every slice length is known on the host (0 = unknown)
    uint64_t tile_bound = 0;
    {
        const uint64_t tb = nz_gather_tile_bytes(variant);
        bool known = true;
        for (uint32_t i = 0; i < (uint32_t)nseg && known; ++i) {
            if (hseg[i].len == 0 && hreq[hseg[i].req].whole) known = false;
            else tile_bound += hseg[i].len / tb + 2;
        }
        if (!known) tile_bound = 0;
    }
    char* dv = (char*)t->dev;
    NzPlanArgs a;
    a.table = c->table; a.line_mask = c->n_lines - 1; a.descs = c->descs; a.arena = c->arena;
    a.reqs_host = reinterpret_cast<const NzDevReq*>(t->pin_dev); a.n = n;
    a.segs_host = reinterpret_cast<const NzDevSeg*>((char*)t->pin_dev + pin_reqs); a.nseg = (uint32_t)nseg;
    a.res = reinterpret_cast<NzDevRes*>((char*)t->pin_dev + pin_reqs + pin_segs);
    a.work = reinterpret_cast<NzSegWork*>(dv);
    a.tile_begin = reinterpret_cast<uint64_t*>(dv + dev_work);
    a.chain_head = reinterpret_cast<uint32_t*>(dv + dev_work + dev_tiles);
    a.chain_next = a.chain_head + nseg;
    a.coalesce = c->coalesce ? 1u : 0u;
    a.ver = reinterpret_cast<NzVerWork*>(dv + dev_work + dev_tiles + dev_chain);
    a.ver_cap = (uint32_t)ver_cap;
    a.counters = reinterpret_cast<uint32_t*>(dv + dev_work + dev_tiles + dev_chain + dev_ver);
    a.infl = reinterpret_cast<NzInflWork*>(dv + dev_work + dev_tiles + dev_chain + dev_ver + dev_cnt);
    a.infl_meta = reinterpret_cast<uint64_t*>(a.infl + n) + 2 * (size_t)n;   // after rq_src, rq_len
    a.rq_stat = reinterpret_cast<int32_t*>(reinterpret_cast<uint64_t*>(a.infl + n) + 5 * (size_t)n);
    a.rq_desc = reinterpret_cast<uint32_t*>(a.rq_stat + n);        // second half of the 8 B / request reserved for rq_stat
    {
        char* stg = dv + dev_work + dev_tiles + dev_chain + dev_ver + dev_cnt + round_up(dev_infl + dev_rq, 256);
        a.reqs_dev = reinterpret_cast<NzDevReq*>(stg);
        a.segs_dev = reinterpret_cast<NzDevSeg*>(stg + round_up((size_t)n * sizeof(NzDevReq), 256));
        a.seg_ver = any_verify ? reinterpret_cast<uint32_t*>(stg + dev_stage) : nullptr;
        a.infl_seg_begin = any_inflate ? reinterpret_cast<uint32_t*>(stg + dev_stage + dev_segver) : nullptr;
        a.coal_tab = coal_cap ? reinterpret_cast<int32_t*>(stg + dev_stage + dev_segver + dev_iseg) : nullptr;
        a.coal_cap = coal_cap;
    }
    a.ver_stat = reinterpret_cast<uint32_t*>(a.res + n);
    t->ver_stat = reinterpret_cast<uint32_t*>(hres + n);
    t->ver_stat[0] = t->ver_stat[1] = t->ver_stat[2] = t->ver_stat[3] = t->ver_stat[4] = 0;
    t->flag_wait = false;
    a.fuse_verify = (any_verify && ver_cap && c->fuse_verify && nz_gather_fuses_verify(variant)) ? 1u : 0u;
    a.peer_base = c->peer_base_dev;
    a.tile_bytes = nz_gather_tile_bytes(variant);

    cudaStream_t s = cuda_stream ? (cudaStream_t)cuda_stream
                     : any_urgent ? c->hi_streams[c->rr.fetch_add(1) % c->hi_streams.size()]
                                  : c->streams[c->rr.fetch_add(1) % c->streams.size()];
    t->stream = s;
    cudaError_t e;
#define CKT(call) do { e = (call); if (e != cudaSuccess) { int32_t rc = cuda_fail(e, #call); nz_ticket_free(t); return rc; } } while (0)
    t->profiled = c->profile;
    std::unique_lock<std::mutex> stage_lk;
    nz_ctx::StageSlot* slot = nullptr;
    if (dma) {                                   // staging slot: next in the ring, reuse ordered behind its previous wave's copies
        stage_lk = std::unique_lock<std::mutex>(c->stage_mu);
        slot = &c->stage[c->stage_next++ % 4];
        if (!slot->free_ev) CKT(cudaEventCreateWithFlags(&slot->free_ev, cudaEventDisableTiming));
        if (slot->cap < dma_bytes) {
            if (slot->used) CKT(cudaEventSynchronize(slot->free_ev));
            if (slot->buf) cudaFree(slot->buf);
            slot->buf = nullptr; slot->cap = 0; slot->used = false;
            const size_t cap = round_up((size_t)dma_bytes, (size_t)64 << 20);
            if (cudaMalloc(&slot->buf, cap) != cudaSuccess) {      // HBM is full (arena sized to the device): no staging from now on,
                cudaGetLastError();                                // this wave is rebuilt for the direct path
                stage_lk.unlock();
                nz_ticket_free(t);
                c->host_dma = false;
                return nz_get_submit(c, reqs, n, cuda_stream, out);
            }
            slot->cap = cap;
        }
        if (slot->used) CKT(cudaStreamWaitEvent(s, slot->free_ev, 0));
        for (uint32_t i = 0; i < n; ++i) hreq[i].dst = (uint64_t)(uintptr_t)slot->buf;
    }
    if (t->profiled) CKT(cudaEventRecord(t->ev_start, s));
    // a few small slices (a first window): the planning CTA copies them itself, one launch instead of two
    const bool fused_small = c->fuse_small && n <= NZ_SMALL_WAVE && nseg <= NZ_SMALL_WAVE && !any_inflate && !any_verify && !dma &&
                             tile_bound > 0 && tile_bound <= 64 && a.tile_bytes == 32768;
    if (n <= NZ_SMALL_WAVE && nseg <= NZ_SMALL_WAVE) {
        NzSmallWave sw;
        memset(&sw, 0, sizeof sw);
        memcpy(sw.reqs, hreq, n * sizeof(NzDevReq));
        memcpy(sw.segs, hseg, (size_t)nseg * sizeof(NzDevSeg));
        if (fused_small) { a.coalesce = 0; t->flag_wait = true; }
        CKT(nz_launch_plan(a, s, &sw, fused_small));
    } else {
        CKT(nz_launch_plan(a, s));
    }
    c->launches += a.ver_cap ? 2 : 1;
    t->ran_inflate = any_inflate;
    if (any_inflate) {
        if (t->profiled) CKT(cudaEventRecord(t->ev_i0, s));
        // members with a checkpoint index: one warp per segment; members without: one warp per member + trailer CRC-32
        // (whether any of those exist is only known on the device: the legacy kernels find an empty work list otherwise)
        CKT(nz_launch_inflate(a.infl, a.infl_meta, a.counters, a.res, c->crc_tables, c->inflate_grid, c->sm_count, s,
                              a.infl_seg_begin, a.reqs_dev, n, c->inflate_seg_grid, true));
        if (t->profiled) CKT(cudaEventRecord(t->ev_i1, s));
        c->launches += 3;
    }
    if (t->profiled) CKT(cudaEventRecord(t->ev_g0, s));
    int ggrid = c->gather_grid;
    if (tile_bound && tile_bound < (uint64_t)ggrid) ggrid = (int)tile_bound;   // small waves: no more CTAs than tiles
    if (host_kernel && ggrid > c->host_grid) ggrid = c->host_grid;  // PCIe-bound wave: a few SMs saturate the link
    // paced host waves: each of the wave's CTAs issues one tile per pace_ns (urgent waves are never paced)
    uint32_t pace_ns = 0;
    const double total_rate = c->host_rate_gbs.load();
    if (host_kernel && !any_urgent && total_rate > 0 && ggrid > 0) {
        const int sharing = c->paced_inflight.fetch_add(1) + 1;          // this wave and the paced ones still in flight
        t->paced = true;
        pace_ns = (uint32_t)std::min(1e6, (double)a.tile_bytes * ggrid * sharing / total_rate);      // bytes / (GB/s) = ns
    }
    // K5: a wave with striped sources walks its tiles spread over the whole wave, so the CTAs running at any moment
    // pull from stripes on different peers (see TileOrder in nz_kernels.cu)
    bool striped_wave = false;
    for (const auto& sp : split) if (!sp.empty()) { striped_wave = true; break; }
    const uint32_t spread = striped_wave ? c->span_spread : 0;
    // does the wave name a key twice?  Then request coalescing will very likely chain slices (fan-out tiles: 2 stages);
    // a wave of distinct keys copies every tile once (3 stages).  The grid follows the stage count's occupancy.
    int stages_hint = 3;
    if (c->coalesce && n > 1) {
        std::unordered_set<uint64_t> seen;
        seen.reserve(n * 2);
        for (uint32_t i = 0; i < n; ++i)
            if (!seen.insert(rd64(reqs[i].key) ^ (rd64(reqs[i].key + 8) * 0x9E3779B97F4A7C15ull)).second) { stages_hint = 2; break; }
    }
    if (stages_hint == 3 && !a.fuse_verify && ggrid > c->gather_grid3) ggrid = c->gather_grid3;
    uint32_t* dyn = (striped_wave && c->span_dynamic) ? a.counters + 6 : nullptr;      // zeroed by k_plan
    if (fused_small) { /* copied by the plan kernel */ }
    else if (a.fuse_verify) CKT(nz_launch_gather(variant, a.work, a.tile_begin, a.chain_head, a.chain_next, a.nseg, ggrid, s, a.seg_ver, a.ver, c->crc_tables, pace_ns, spread, stages_hint, dyn));
    else CKT(nz_launch_gather(variant, a.work, a.tile_begin, a.chain_head, a.chain_next, a.nseg, ggrid, s, nullptr, nullptr, nullptr, pace_ns, spread, stages_hint, dyn));
    if (t->profiled) CKT(cudaEventRecord(t->ev_g1, s));
    c->launches += fused_small ? 0 : 1;
    if (any_verify && ver_cap) {
        int grid = (int)std::min<uint64_t>(ver_cap, (uint64_t)c->sm_count * 6);      // k_verify: 6 CTAs of 256 threads fit an SM (36 KiB of tables each)
        CKT(nz_launch_verify(a, c->crc_tables, grid, s));
        c->launches += 1;
    }
    if (dma) {                                   // copy engine: every slice from the staging slot to its pinned destination
        for (const DmaCopy& cp : dma_copies)
            if (cp.len) CKT(cudaMemcpyAsync(cp.host, (char*)slot->buf + cp.off, cp.len, cudaMemcpyDeviceToHost, s));
        CKT(cudaEventRecord(slot->free_ev, s));
        slot->used = true;
        c->dma_waves++;
    }
    CKT(cudaEventRecord(t->ev_done, s));
#undef CKT
    {
        std::lock_guard<std::mutex> g(c->tmu);
        t->seq = ++c->submit_seq;
        c->inflight[t->seq] = t;
    }
    *out = t;
    return NZ_OK;
}

int32_t nz_get_poll(nz_ticket* t) {
    if (!t) return NZ_EINVAL;
    if (t->flag_wait && t->ver_stat && *(volatile uint32_t*)(t->ver_stat + 2)) return NZ_OK;
    cudaError_t e = cudaEventQuery(t->ev_done);
    if (e == cudaSuccess) return NZ_OK;
    if (e == cudaErrorNotReady) return NZ_ETIMEOUT;
    return cuda_fail(e, "cudaEventQuery");
}

int32_t nz_get_wait(nz_ticket* t, int64_t timeout_us, nz_get_res* res) {
    if (!t) return NZ_EINVAL;
    nz_ctx* c = t->ctx;
    cudaSetDevice(c->device);
    if (t->flag_wait && t->ver_stat) {            // fused small wave: the kernel's last store is a flag in pinned memory
        volatile uint32_t* flag = t->ver_stat + 2;
        const auto t0 = std::chrono::steady_clock::now();
        uint32_t spins = 0;
        while (!*flag) {
            __builtin_ia32_pause();
            if ((++spins & 1023u) == 0) {
                const auto us = std::chrono::duration_cast<std::chrono::microseconds>(std::chrono::steady_clock::now() - t0).count();
                if (us > 300 || (timeout_us >= 0 && us >= timeout_us)) break;      // long queue ahead (or an error): the event path decides
            }
        }
        std::atomic_thread_fence(std::memory_order_acquire);
    }
    if (t->flag_wait && t->ver_stat && t->ver_stat[2]) {
        // complete: results are in pinned memory already
    } else if (timeout_us < 0) {
        // short waves finish in microseconds: poll first (a blocking sync costs an interrupt wake-up)
        const auto t0 = std::chrono::steady_clock::now();
        for (;;) {
            cudaError_t e = cudaEventQuery(t->ev_done);
            if (e == cudaSuccess) break;
            if (e != cudaErrorNotReady) return cuda_fail(e, "cudaEventQuery");
            // long wave: sleep between polls (a blocking cudaEventSynchronize was measured to return ~300 us late)
            if (std::chrono::steady_clock::now() - t0 > std::chrono::microseconds(200)) {
                struct timespec ts = {0, 20000};
                nanosleep(&ts, nullptr);
            }
        }
    } else {
        const auto t0 = std::chrono::steady_clock::now();
        for (;;) {
            cudaError_t e = cudaEventQuery(t->ev_done);
            if (e == cudaSuccess) break;
            if (e != cudaErrorNotReady) return cuda_fail(e, "cudaEventQuery");
            const auto us = std::chrono::duration_cast<std::chrono::microseconds>(std::chrono::steady_clock::now() - t0).count();
            if (us >= timeout_us) return NZ_ETIMEOUT;
            if (us > 200) std::this_thread::sleep_for(std::chrono::microseconds(50));
            else std::this_thread::yield();
        }
    }
    const size_t pin_reqs = round_up((size_t)t->n * sizeof(NzDevReq), 256);
    const size_t pin_segs = round_up((size_t)t->nseg * sizeof(NzDevSeg), 256);
    const NzDevRes* hres = reinterpret_cast<const NzDevRes*>((char*)t->pin + pin_reqs + pin_segs);
    uint64_t bytes = 0, hits = 0, misses = 0, corrupt = 0;
    for (uint32_t i = 0; i < t->n; ++i) {
        if (res) memcpy(&res[i], &hres[i], sizeof(nz_get_res));
        if (hres[i].status == NZ_OK) { ++hits; bytes += hres[i].bytes; }
        else if (hres[i].status == NZ_ENOTFOUND) ++misses;
        else if (hres[i].status == NZ_ECORRUPT) ++corrupt;
    }
    if (t->paced) { t->paced = false; c->paced_inflight--; }
    if (!t->req_bytes_hint) {   // count a wave once even if waited twice
        t->req_bytes_hint = 1;
        c->gets += t->n; c->get_bytes += bytes; c->hits += hits; c->misses += misses; c->corrupt += corrupt;
        if (t->ver_stat) { c->verify_chunks += t->ver_stat[0]; c->verify_fused += t->ver_stat[1]; c->infl_segs += t->ver_stat[3]; c->infl_whole += t->ver_stat[4]; }
    }
    return NZ_OK;
}

int32_t nz_ticket_elapsed_ms(nz_ticket* t, float* ms) {
    if (!t || !ms) return NZ_EINVAL;
    if (!t->profiled) return NZ_ENOSYS;      // nz_set_profiling(ctx, 1) first
    CK(cudaEventSynchronize(t->ev_done));    // a flag-completed small wave may return before its last event has been timestamped
    CK(cudaEventElapsedTime(ms, t->ev_start, t->ev_done));
    return NZ_OK;
}

int32_t nz_ticket_gather_ms(nz_ticket* t, float* ms) {
    if (!t || !ms) return NZ_EINVAL;
    if (!t->profiled) return NZ_ENOSYS;
    CK(cudaEventSynchronize(t->ev_done));
    CK(cudaEventElapsedTime(ms, t->ev_g0, t->ev_g1));
    return NZ_OK;
}

int32_t nz_ticket_inflate_ms(nz_ticket* t, float* ms) {
    if (!t || !ms) return NZ_EINVAL;
    *ms = 0.f;
    if (!t->ran_inflate) return NZ_OK;
    if (!t->profiled) return NZ_ENOSYS;
    CK(cudaEventElapsedTime(ms, t->ev_i0, t->ev_i1));
    return NZ_OK;
}

int32_t nz_ticket_free(nz_ticket* t) {
    if (!t) return NZ_EINVAL;
    nz_ctx* c = t->ctx;
    cudaSetDevice(c->device);
    cudaEventSynchronize(t->ev_done);   // never recycle buffers a kernel may still read
    if (t->paced) { t->paced = false; c->paced_inflight--; }
    t->req_bytes_hint = 0;
    t->stream = nullptr;
    std::lock_guard<std::mutex> g(c->tmu);
    if (t->seq) { c->inflight.erase(t->seq); t->seq = 0; }
    if (c->ticket_pool.size() < 64) c->ticket_pool.push_back(t);
    else ticket_destroy(t);
    return NZ_OK;
}

int32_t nz_set_host_rate(nz_ctx* c, double total_gbs, int32_t ctas) {
    if (!c || total_gbs < 0 || ctas < 0) return NZ_EINVAL;
    c->host_rate_gbs = total_gbs;
    c->host_grid = ctas > 0 ? ctas : 32;
    return NZ_OK;
}

int32_t nz_set_profiling(nz_ctx* c, int32_t on) {
    if (!c) return NZ_EINVAL;
    c->profile = on != 0;
    return NZ_OK;
}

int32_t nz_get_stats(nz_ctx* c, nz_stats* out) {
    if (!c || !out) return NZ_EINVAL;
    memset(out, 0, sizeof *out);
    {
        std::lock_guard<std::mutex> g(c->mu);
        out->objects = c->shadow.size();
        uint64_t freeb = 0;
        for (auto& kv : c->free_by_off) freeb += kv.second;
        out->arena_used = c->arena_bytes - freeb;
        out->table_live = c->shadow.size();
    }
    out->arena_bytes = c->arena_bytes;
    out->table_slots = c->n_lines * 4;
    out->gets = c->gets; out->get_bytes = c->get_bytes; out->hits = c->hits; out->misses = c->misses;
    out->puts = c->puts; out->put_bytes = c->put_bytes; out->evictions = c->evictions; out->corrupt = c->corrupt;
    out->kernel_launches = c->launches;
    out->file_cufile_bytes = c->file_cufile_bytes;
    out->file_bounce_bytes = c->file_bounce_bytes;
    out->verify_chunks = c->verify_chunks;
    out->verify_fused_chunks = c->verify_fused;
    out->host_dma_waves = c->dma_waves;
    out->inflate_segments = c->infl_segs; out->inflate_whole_members = c->infl_whole; out->index_bytes = c->index_bytes;
    return NZ_OK;
}


// ---------------------------------------------------------------------------
// cross-shard objects: IPC-mapped peer arenas + stripe maps
// ---------------------------------------------------------------------------
static_assert(sizeof(cudaIpcMemHandle_t) == NZ_IPC_HANDLE_BYTES, "IPC handle size");

int32_t nz_arena_ipc_export(nz_ctx* c, uint8_t handle[NZ_IPC_HANDLE_BYTES], uint64_t* arena_bytes) {
    if (!c || !handle) return NZ_EINVAL;
    CK(cudaSetDevice(c->device));
    cudaIpcMemHandle_t h;
    CK(cudaIpcGetMemHandle(&h, c->arena));
    memcpy(handle, &h, sizeof h);
    if (arena_bytes) *arena_bytes = c->arena_bytes;
    return NZ_OK;
}

int32_t nz_peer_attach(nz_ctx* c, uint32_t shard, const uint8_t handle[NZ_IPC_HANDLE_BYTES]) {
    if (!c || !handle || shard >= c->cfg.shard_count || shard == c->cfg.shard_index) return NZ_EINVAL;
    CK(cudaSetDevice(c->device));
    if (c->peer_ptrs[shard]) return NZ_EEXIST;
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, sizeof h);
    void* p = nullptr;
    CK(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    c->peer_ptrs[shard] = p;
    const uint64_t v = (uint64_t)(uintptr_t)p;
    CK(cudaMemcpy(c->peer_base_dev + shard, &v, 8, cudaMemcpyHostToDevice));
    return NZ_OK;
}

int32_t nz_extent_location(nz_ctx* c, const uint8_t key[16], uint64_t* arena_off, uint64_t* stored_len,
                           uint64_t* crc_tab_off, uint32_t* crc32c) {
    if (!c || !key) return NZ_EINVAL;
    std::lock_guard<std::mutex> g(c->mu);
    auto it = c->shadow.find({rd64(key), rd64(key + 8)});
    if (it == c->shadow.end() || it->second.desc == 0xFFFFFFFFu || it->second.stripe_bytes) return NZ_ENOTFOUND;
    if (arena_off) *arena_off = it->second.arena_off;
    if (stored_len) *stored_len = it->second.stored_len;
    if (crc_tab_off) *crc_tab_off = it->second.crc_tab_off;
    if (crc32c) *crc32c = it->second.crc_whole;
    return NZ_OK;
}

uint32_t nz_crc32c_combine(uint32_t crc_a, uint32_t crc_b, uint64_t len_b) {
    // x^(8*len_b) by square and multiply (no cycle shortcut: the Castagnoli polynomial is not irreducible)
    uint32_t p = 0x80000000u, sq = 0x80000000u >> 8;
    while (len_b) {
        if (len_b & 1) p = nz_gf_mul_c(sq, p);
        sq = nz_gf_mul_c(sq, sq);
        len_b >>= 1;
    }
    return nz_gf_mul_c(p, crc_a) ^ crc_b;
}

int32_t nz_put_striped(nz_ctx* c, const uint8_t key[16], uint64_t total_len, uint64_t stripe_bytes, uint32_t n_stripes,
                       const nz_stripe* stripes, uint32_t crc32c, uint64_t user_meta) {
    if (!c || !key || !stripes || !n_stripes || !stripe_bytes) return NZ_EINVAL;
    if (nz_shard_of(key, c->cfg.shard_count) != c->cfg.shard_index) return NZ_EINVAL;
    uint64_t sum = 0;
    for (uint32_t i = 0; i < n_stripes; ++i) {
        if (stripes[i].shard >= c->cfg.shard_count) return NZ_EINVAL;
        if (i + 1 < n_stripes && stripes[i].len != stripe_bytes) return NZ_EINVAL;
        if (stripes[i].len == 0 || stripes[i].len > stripe_bytes) return NZ_EINVAL;
        sum += stripes[i].len;
    }
    if (sum != total_len) return NZ_EINVAL;
    CK(cudaSetDevice(c->device));
    const uint64_t k0 = rd64(key), k1 = rd64(key + 8);
    const uint64_t map_len = round_up((uint64_t)n_stripes * sizeof(NzDevStripe), 4096);
    uint64_t off = 0;
    uint32_t desc_idx;
    uint64_t gen;
    {
        std::lock_guard<std::mutex> g(c->mu);
        if (c->shadow.count({k0, k1})) return NZ_EEXIST;
        if (c->shadow.size() * 2 >= c->n_lines * 4) return NZ_EFULL;
        if (c->free_desc.empty()) return NZ_ENOMEM;
        if (!arena_alloc(c, map_len, &off)) return NZ_ENOMEM;
        desc_idx = c->free_desc.back();
        c->free_desc.pop_back();
        gen = c->gen++;
        NzHostObj ho;
        ho.desc = 0xFFFFFFFFu; ho.arena_off = off; ho.arena_len = map_len; ho.stored_len = total_len; ho.raw_len = total_len;
        ho.flags = NZ_ENC_STRIPED; ho.stripe_bytes = stripe_bytes; ho.crc_whole = crc32c;
        c->shadow[{k0, k1}] = ho;
    }
    int32_t rc = NZ_OK;
    Mailbox* mb = g_mail.get();
    cudaStream_t s = c->put_stream;
    cudaError_t e = cudaSuccess;
    void* mb_dev = nullptr;
    NzDesc d;
    memset(&d, 0, sizeof d);
    d.arena_off = off; d.stored_len = total_len; d.raw_len = total_len; d.user_meta = user_meta; d.crc_tab_off = 0;
    d.flags = NZ_ENC_STRIPED; d.crc32c = crc32c; d.n_chunks = n_stripes; d.gen = (uint32_t)gen; d.stripe_bytes = stripe_bytes;
    if (!mb) { rc = NZ_ENOMEM; goto out; }
    {
        std::lock_guard<std::mutex> g(c->put_mu);
        e = cudaHostGetDevicePointer(&mb_dev, mb, 0);
        if (e == cudaSuccess) e = cudaMemcpyAsync(c->arena + off, stripes, (size_t)n_stripes * sizeof(NzDevStripe), cudaMemcpyHostToDevice, s);
        if (e == cudaSuccess) e = cudaMemcpyAsync(c->descs + desc_idx, &d, sizeof d, cudaMemcpyHostToDevice, s);
        if (e == cudaSuccess) e = nz_launch_insert(c->table, c->n_lines - 1, k0, k1, desc_idx, gen, &((Mailbox*)mb_dev)->status, s);
        c->launches += 1;
        if (e == cudaSuccess) e = cudaStreamSynchronize(s);
    }
    if (e != cudaSuccess) { rc = cuda_fail(e, "nz_put_striped"); goto out; }
    rc = mb->status;
out:
    if (mb) g_mail.put(mb);
    std::lock_guard<std::mutex> g(c->mu);
    if (rc == NZ_OK) {
        c->shadow[{k0, k1}].desc = desc_idx;
        c->n_striped++;
        c->puts++;
    } else {
        c->shadow.erase({k0, k1});
        arena_free(c, off, map_len);
        c->free_desc.push_back(desc_idx);
    }
    return rc;
}

int32_t nz_debug_poke(nz_ctx* c, const uint8_t key[16], uint64_t off, uint8_t xor_mask) {
    if (!c || !key) return NZ_EINVAL;
    CK(cudaSetDevice(c->device));
    NzHostObj ho;
    {
        std::lock_guard<std::mutex> g(c->mu);
        auto it = c->shadow.find({rd64(key), rd64(key + 8)});
        if (it == c->shadow.end() || it->second.desc == 0xFFFFFFFFu) return NZ_ENOTFOUND;
        ho = it->second;
    }
    if (off >= ho.stored_len) return NZ_EINVAL;
    uint8_t b = 0;
    CK(cudaMemcpy(&b, c->arena + ho.arena_off + off, 1, cudaMemcpyDeviceToHost));
    b ^= xor_mask;
    CK(cudaMemcpy(c->arena + ho.arena_off + off, &b, 1, cudaMemcpyHostToDevice));
    return NZ_OK;
}

int32_t nz_crc32c_device(nz_ctx* c, const void* dev, uint64_t n, uint32_t* out_crc, uint32_t* out_chunks_host) {
    if (!c || (!dev && n) || !out_crc) return NZ_EINVAL;
    CK(cudaSetDevice(c->device));
    const uint64_t nchunks = (n + NZ_CRC_CHUNK - 1) / NZ_CRC_CHUNK;
    uint32_t* tab = nullptr;
    CK(cudaMalloc(&tab, (nchunks + 1) * 4 + 256));
    cudaStream_t s = c->streams[c->rr.fetch_add(1) % c->streams.size()];
    uint32_t* whole = tab + nchunks;
    cudaError_t e = nz_launch_crc_chunks((const uint8_t*)dev, n, tab, whole, c->crc_tables, c->sm_count, s);
    c->launches += 2;
    uint32_t h = 0;
    if (e == cudaSuccess) e = cudaMemcpyAsync(&h, whole, 4, cudaMemcpyDeviceToHost, s);
    if (e == cudaSuccess && out_chunks_host && nchunks) e = cudaMemcpyAsync(out_chunks_host, tab, nchunks * 4, cudaMemcpyDeviceToHost, s);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s);
    cudaFree(tab);
    if (e != cudaSuccess) return cuda_fail(e, "crc32c_device");
    *out_crc = h;
    return NZ_OK;
}

}  // extern "C"
