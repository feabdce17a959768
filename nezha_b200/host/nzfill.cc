// nezha_b200/host/nzfill.cc — fill-side scale at the C ABI: how fast can a zone take a LARGE number of small
// objects (BASELINE configs[3] is a 1 TB store: ~10^6 keys per GPU), and what does a lookup cost afterwards?
//   nzfill [--device N] [--objects K] [--bytes B] [--batch M]
// Every batch: nz_put_begin + nz_put_write (from one pinned window, asynchronous) per object, then ONE
// nz_put_commit_many (batched CRC-32C kernels, one publish launch).  Then: nz_lookup of every key, and a wave of GETs
// over a sample whose bytes are compared with what was put.  Prints one JSON object.
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/nezha_gpu.h"

static inline uint64_t mix(uint64_t x) { x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33; return x; }
static void key_of(uint64_t i, uint8_t k[16]) { const uint64_t a = mix(i * 2 + 1), b = mix(i * 2 + 2); memcpy(k, &a, 8); memcpy(k + 8, &b, 8); }
static double now_s() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

int main(int argc, char** argv) {
    int device = 0;
    uint64_t objects = 1000000, bytes = 16384, batch = 4096;
    int packed = 0;
    for (int i = 1; i + 1 < argc; i += 2) {
        std::string a = argv[i];
        if (a == "--device") device = atoi(argv[i + 1]);
        else if (a == "--objects") objects = strtoull(argv[i + 1], nullptr, 10);
        else if (a == "--bytes") bytes = strtoull(argv[i + 1], nullptr, 10);
        else if (a == "--batch") batch = strtoull(argv[i + 1], nullptr, 10);
        else if (a == "--packed") packed = atoi(argv[i + 1]);
    }
    const uint64_t slot = (bytes + 4095) / 4096 * 4096 + 4096;        // extent + chunk table, 4 KiB granular
    nz_config cfg;
    memset(&cfg, 0, sizeof cfg);
    cfg.struct_size = sizeof cfg; cfg.device = device; cfg.arena_bytes = objects * slot + (256ull << 20); cfg.max_objects = (uint32_t)(objects + 1024);
    nz_ctx* ctx = nullptr;
    int32_t rc = nz_init(&cfg, &ctx);
    if (rc != NZ_OK) { fprintf(stderr, "nzfill: nz_init: %s (%s)\n", nz_strerror(rc), nz_last_error()); return 1; }
    nz_buf* win[2] = {nullptr, nullptr};
    for (auto& w : win) if (nz_buf_alloc(ctx, batch * bytes, NZ_BUF_PINNED_HOST, &w) != NZ_OK) { fprintf(stderr, "nzfill: window\n"); return 1; }
    for (auto& w : win) {
        uint8_t* p = (uint8_t*)nz_buf_ptr(w);
        for (uint64_t i = 0; i < batch * bytes; i += 8) { const uint64_t v = mix(i + (uint64_t)(w == win[1])); memcpy(p + i, &v, (batch * bytes - i) < 8 ? (batch * bytes - i) : 8); }
    }
    std::vector<nz_put*> puts(batch);
    std::vector<int32_t> st(batch);
    std::vector<uint32_t> crcs(batch);
    const double t0 = now_s();
    uint64_t done = 0, nb = 0;
    double t_begin = 0, t_commit = 0;
    while (done < objects) {
        const uint64_t n = objects - done < batch ? objects - done : batch;
        nz_buf* w = win[nb & 1];                         // the other window's copies were waited for by the previous commit
        const uint8_t* src = (const uint8_t*)nz_buf_ptr(w);
        const double a = now_s();
        double b = a;
        if (packed) {                                     // ONE copy + ONE scatter launch + batched checksums for the whole batch
            static std::vector<uint8_t> kk;
            static std::vector<uint64_t> so, ln, um;
            kk.resize(n * 16); so.resize(n); ln.resize(n); um.resize(n);
            for (uint64_t i = 0; i < n; ++i) { key_of(done + i, &kk[i * 16]); so[i] = i * bytes; ln[i] = bytes; um[i] = done + i; }
            b = now_s();
            if ((rc = nz_put_packed(ctx, (uint32_t)n, (const uint8_t(*)[16])kk.data(), so.data(), ln.data(), w, um.data(), crcs.data(), st.data())) != NZ_OK) {
                fprintf(stderr, "nzfill: nz_put_packed: %s (%s)\n", nz_strerror(rc), nz_last_error()); return 1; }
        } else {
            for (uint64_t i = 0; i < n; ++i) {
                uint8_t k[16];
                key_of(done + i, k);
                if ((rc = nz_put_begin(ctx, k, bytes, bytes, NZ_ENC_IDENTITY, done + i, &puts[i])) != NZ_OK ||
                    (rc = nz_put_write(puts[i], src + i * bytes, 0, bytes)) != NZ_OK) { fprintf(stderr, "nzfill: put %llu: %s\n", (unsigned long long)(done + i), nz_strerror(rc)); return 1; }
            }
            b = now_s();
            if ((rc = nz_put_commit_many(puts.data(), (uint32_t)n, nullptr, nullptr, crcs.data(), st.data())) != NZ_OK) { fprintf(stderr, "nzfill: commit_many: %s\n", nz_strerror(rc)); return 1; }
        }
        for (uint64_t i = 0; i < n; ++i) if (st[i] != NZ_OK) { fprintf(stderr, "nzfill: object %llu: %s\n", (unsigned long long)(done + i), nz_strerror(st[i])); return 1; }
        t_begin += b - a; t_commit += now_s() - b;
        done += n; ++nb;
    }
    const double fill_s = now_s() - t0;
    // lookups: every key, 64 Ki per call
    std::vector<uint8_t> keys(65536 * 16);
    std::vector<nz_extent> ext(65536);
    uint64_t found = 0;
    const double l0 = now_s();
    for (uint64_t base = 0; base < objects; base += 65536) {
        const uint64_t n = objects - base < 65536 ? objects - base : 65536;
        for (uint64_t i = 0; i < n; ++i) key_of(base + i, &keys[i * 16]);
        if (nz_lookup(ctx, (const uint8_t(*)[16])keys.data(), (uint32_t)n, ext.data()) != NZ_OK) { fprintf(stderr, "nzfill: lookup\n"); return 1; }
        for (uint64_t i = 0; i < n; ++i) found += ext[i].status == NZ_OK && ext[i].stored_length == bytes && ext[i].user_meta == base + i;
    }
    const double lookup_s = now_s() - l0;
    // GET a sample and compare with the window contents they were written from
    const uint32_t S = 4096;
    nz_buf* out = nullptr;
    nz_buf_alloc(ctx, (uint64_t)S * bytes, NZ_BUF_PINNED_HOST, &out);
    std::vector<nz_get_req> rq(S);
    std::vector<nz_get_res> rs(S);
    std::vector<uint64_t> which(S);
    for (uint32_t j = 0; j < S; ++j) {
        which[j] = mix(j) % objects;
        memset(&rq[j], 0, sizeof rq[j]);
        key_of(which[j], rq[j].key);
        rq[j].dst = out; rq[j].dst_base = (uint64_t)j * bytes;
    }
    nz_ticket* t = nullptr;
    uint64_t bad = 0;
    if (nz_get_submit(ctx, rq.data(), S, nullptr, &t) != NZ_OK || nz_get_wait(t, -1, rs.data()) != NZ_OK) { fprintf(stderr, "nzfill: get\n"); return 1; }
    nz_ticket_free(t);
    for (uint32_t j = 0; j < S; ++j) {
        const uint64_t i = which[j], b = i / batch, within = i % batch;
        const uint8_t* src = (const uint8_t*)nz_buf_ptr(win[b & 1]) + within * bytes;
        bad += rs[j].status != NZ_OK || memcmp((const uint8_t*)nz_buf_ptr(out) + (uint64_t)j * bytes, src, bytes) != 0;
    }
    nz_stats stt;
    nz_get_stats(ctx, &stt);
    printf("{\"packed\": %d, \"objects\": %llu, \"object_bytes\": %llu, \"batch\": %llu, \"fill_s\": %.3f, \"puts_per_s\": %.0f, \"fill_gbs\": %.2f, \"host_us_per_put_begin_write\": %.2f, "
           "\"host_us_per_put_commit\": %.2f, \"lookup_s\": %.3f, \"lookups_per_s\": %.0f, \"found\": %llu, \"sample_gets\": %u, \"sample_mismatches\": %llu, \"table_live\": %llu}\n",
           packed, (unsigned long long)objects, (unsigned long long)bytes, (unsigned long long)batch, fill_s, objects / fill_s, objects * bytes / fill_s / 1e9, t_begin / objects * 1e6,
           t_commit / objects * 1e6, lookup_s, objects / lookup_s, (unsigned long long)found, S, (unsigned long long)bad, (unsigned long long)stt.table_live);
    nz_shutdown(ctx);
    return (found == objects && bad == 0) ? 0 : 2;
}
