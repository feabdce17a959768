// nezha_b200/host/nzlat.cc — time-to-first-byte of one cached GET at the C ABI, measured from
// C++ (no interpreter in the loop): nz_get_submit + nz_get_wait of the first window of a GET
// into pinned host memory, one GET in flight.  Prints one JSON object.
//   nzlat [--device N] [--first-kib K] [--iters I] [--urgent 0|1] [--bg-threads B --bg-mib M]
// --bg-threads B: B background threads keep bulk waves of M MiB (4 MiB windows to pinned host memory) in
// flight, so the measured GET competes for the PCIe link the way a first window does in nzproxy.
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <atomic>
#include <string>
#include <thread>
#include <vector>

#include "../../include/nezha_gpu.h"

int main(int argc, char** argv) {
    int device = 0, iters = 2000, urgent = 0, bg_threads = 0;
    uint64_t first = 256 << 10, bg_mib = 16;
    for (int i = 1; i + 1 < argc; i += 2) {
        std::string a = argv[i];
        if (a == "--device") device = atoi(argv[i + 1]);
        else if (a == "--first-kib") first = (uint64_t)atoll(argv[i + 1]) << 10;
        else if (a == "--iters") iters = atoi(argv[i + 1]);
        else if (a == "--urgent") urgent = atoi(argv[i + 1]);
        else if (a == "--bg-threads") bg_threads = atoi(argv[i + 1]);
        else if (a == "--bg-mib") bg_mib = (uint64_t)atoll(argv[i + 1]);
    }
    nz_config cfg;
    memset(&cfg, 0, sizeof cfg);
    cfg.struct_size = sizeof cfg; cfg.device = device; cfg.arena_bytes = 256ull << 20; cfg.max_objects = 64;
    nz_ctx* ctx = nullptr;
    int32_t rc = nz_init(&cfg, &ctx);
    if (rc != NZ_OK) { fprintf(stderr, "nzlat: nz_init: %s (%s)\n", nz_strerror(rc), nz_last_error()); return 1; }
    const uint64_t L = 64ull << 20;
    std::vector<uint8_t> data(L);
    for (uint64_t i = 0; i < L; ++i) data[i] = (uint8_t)(i * 2654435761u >> 13);
    uint8_t key[16] = {1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16};
    nz_put* p = nullptr;
    uint32_t crc = 0;
    if (nz_put_begin(ctx, key, L, L, NZ_ENC_IDENTITY, 0, &p) != NZ_OK || nz_put_write(p, data.data(), 0, L) != NZ_OK ||
        nz_put_commit(p, 0, 0, &crc) != NZ_OK) { fprintf(stderr, "nzlat: put failed\n"); return 1; }
    nz_buf* win = nullptr;
    nz_buf_alloc(ctx, first, NZ_BUF_PINNED_HOST, &win);
    std::atomic<bool> stop{false};
    std::atomic<uint64_t> bg_bytes{0};
    std::vector<std::thread> bg;
    for (int b = 0; b < bg_threads; ++b)
        bg.emplace_back([&, b] {
            const uint32_t nw = (uint32_t)std::max<uint64_t>(1, bg_mib / 4);
            std::vector<nz_buf*> w(nw);
            for (auto& x : w) nz_buf_alloc(ctx, 4 << 20, NZ_BUF_PINNED_HOST, &x);
            std::vector<nz_get_req> rq(nw);
            std::vector<nz_seg> sg(nw);
            std::vector<nz_get_res> rs(nw);
            uint64_t k = b * 7919;
            while (!stop.load()) {
                for (uint32_t i = 0; i < nw; ++i, ++k) {
                    memset(&rq[i], 0, sizeof rq[i]);
                    memcpy(rq[i].key, key, 16);
                    sg[i] = nz_seg{(k * (4ull << 20)) % (L - (4 << 20)), 4ull << 20, 0};
                    rq[i].nseg = 1; rq[i].segs = &sg[i]; rq[i].dst = w[i];
                }
                nz_ticket* t = nullptr;
                if (nz_get_submit(ctx, rq.data(), nw, nullptr, &t) == NZ_OK) { nz_get_wait(t, -1, rs.data()); bg_bytes += (uint64_t)nw * (4ull << 20); }
                if (t) nz_ticket_free(t);
            }
            for (auto& x : w) nz_buf_free(x);
        });
    if (bg_threads) std::this_thread::sleep_for(std::chrono::milliseconds(200));
    const auto bg_t0 = std::chrono::steady_clock::now();
    const uint64_t bg_b0 = bg_bytes.load();
    if (bg_threads) nz_set_profiling(ctx, 1);      // device-side time of the measured wave (adds two event records)
    double gpu_us_sum = 0;
    std::vector<double> us;
    us.reserve(iters);
    uint64_t bad = 0;
    for (int it = 0; it < iters + 50; ++it) {
        nz_get_req rq;
        memset(&rq, 0, sizeof rq);
        memcpy(rq.key, key, 16);
        nz_seg sg = {(uint64_t)(it * 4099 % (L - first)), first, 0};
        rq.nseg = 1; rq.segs = &sg; rq.dst = win;
        rq.flags = urgent ? NZ_GET_URGENT : 0;
        if (bg_threads) std::this_thread::sleep_for(std::chrono::microseconds(300));     // arrivals spread over the bulk waves
        const auto t0 = std::chrono::steady_clock::now();
        nz_ticket* t = nullptr;
        nz_get_res rs;
        if (nz_get_submit(ctx, &rq, 1, nullptr, &t) != NZ_OK || nz_get_wait(t, -1, &rs) != NZ_OK || rs.status != NZ_OK) ++bad;
        const double dt = std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t0).count();
        float gms = 0.f;
        if (t && it >= 50 && nz_ticket_elapsed_ms(t, &gms) == NZ_OK) gpu_us_sum += gms * 1e3;
        if (t) nz_ticket_free(t);
        if (it >= 50) us.push_back(dt);
        if (memcmp(nz_buf_ptr(win), data.data() + sg.src_off, 64) != 0) ++bad;
    }
    const double bg_s = std::chrono::duration<double>(std::chrono::steady_clock::now() - bg_t0).count();
    const double bg_gbs = bg_threads ? (bg_bytes.load() - bg_b0) / bg_s / 1e9 : 0.0;
    stop = true;
    for (auto& th : bg) th.join();
    std::sort(us.begin(), us.end());
    printf("{\"first_bytes\": %llu, \"iters\": %d, \"errors\": %llu, \"p50_us\": %.2f, \"p90_us\": %.2f, \"p99_us\": %.2f, \"min_us\": %.2f, "
           "\"gpu_mean_us\": %.2f, \"urgent\": %d, \"bg_threads\": %d, \"bg_mib\": %llu, \"bg_gbs\": %.2f}\n",
           (unsigned long long)first, iters, (unsigned long long)bad, us[us.size() / 2], us[us.size() * 9 / 10], us[us.size() * 99 / 100], us[0],
           gpu_us_sum / std::max<size_t>(1, us.size()), urgent, bg_threads, (unsigned long long)bg_mib, bg_gbs);
    nz_buf_free(win);
    nz_shutdown(ctx);
    return bad ? 1 : 0;
}
