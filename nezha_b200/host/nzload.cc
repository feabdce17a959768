// nezha_b200/host/nzload.cc — HTTP load generator for the cached-GET metric.
//
// Measures what BASELINE.json's metric names at the wire: payload GB/s and time to first
// byte (ttfb: first byte of the response; ttfbody: first byte of the BODY) of cached GETs, the same way against the GPU proxy (nzproxy) and the nginx-restated
// CPU server.  One thread per connection (= one in-flight GET per connection, like the
// reference's wget clients, /root/reference/examples/demo/kubeflow-test.yaml:20-22);
// keep-alive; bodies are read and discarded.
//
//   nzload --target ip:port --conns N --log FILE [--repeat R] [--warmup W]
//   log line:  <host> <path> <range-spec or ->
// Prints one JSON object.
#include <arpa/inet.h>
#include <netinet/in.h>
#include <netinet/tcp.h>
#include <sys/socket.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <mutex>
#include <sstream>
#include <string>
#include <thread>
#include <vector>

struct Req { std::string host, path, range; };
using Clock = std::chrono::steady_clock;

static std::atomic<int> g_ready{0};
static std::atomic<bool> g_go{false};
static std::atomic<uint64_t> g_bytes{0}, g_reqs{0}, g_errors{0}, g_s2xx{0}, g_sother{0};
static std::mutex g_mu;
static std::vector<double> g_ttfb_us, g_ttfbody_us;

static bool g_discard = false;
static size_t g_buf_bytes = 256u << 10;    // --buf-kib: receive buffer per connection.  64 connections x 4 MiB (round 1) did not stay in any cache and the
                                           // CLIENT's copies cost the shared box 20 % of the figure (profiles/r02_sweep_load_client.txt); 256 KiB is what
                                           // curl/wget-sized clients use

static int dial(const std::string& ip, int port) {
    int fd = socket(AF_INET, SOCK_STREAM, 0);
    if (fd < 0) return -1;
    int one = 1, rcv = 8 << 20;
    setsockopt(fd, IPPROTO_TCP, TCP_NODELAY, &one, sizeof one);
    setsockopt(fd, SOL_SOCKET, SO_RCVBUF, &rcv, sizeof rcv);
    sockaddr_in sa;
    memset(&sa, 0, sizeof sa);
    sa.sin_family = AF_INET;
    sa.sin_port = htons((uint16_t)port);
    inet_pton(AF_INET, ip.c_str(), &sa.sin_addr);
    if (connect(fd, (sockaddr*)&sa, sizeof sa) != 0) { close(fd); return -1; }
    return fd;
}

// one GET on an open connection; returns body bytes or -1; *ttfb_us gets the first-byte latency
static int64_t one_get(int fd, const Req& r, std::vector<char>& buf, double* ttfb_us, int* status, double* ttfbody_us) {
    std::string q = "GET " + r.path + " HTTP/1.1\r\nHost: " + r.host + "\r\n";
    if (r.range != "-") q += "Range: " + r.range + "\r\n";
    q += "Connection: keep-alive\r\n\r\n";
    size_t off = 0;
    while (off < q.size()) {
        ssize_t w = send(fd, q.data() + off, q.size() - off, MSG_NOSIGNAL);
        if (w <= 0) return -1;
        off += (size_t)w;
    }
    const auto t0 = Clock::now();
    std::string head;
    size_t he = std::string::npos;
    bool first = true;
    while (he == std::string::npos) {
        ssize_t n = recv(fd, buf.data(), buf.size(), 0);
        if (n <= 0) return -1;
        if (first) { *ttfb_us = std::chrono::duration<double, std::micro>(Clock::now() - t0).count(); first = false; }
        head.append(buf.data(), (size_t)n);
        he = head.find("\r\n\r\n");
        if (head.size() > (1u << 20) && he == std::string::npos) return -1;
    }
    *status = atoi(head.c_str() + 9);
    int64_t clen = 0;
    {
        std::string lower = head.substr(0, he);
        for (char& c : lower) c = (char)tolower((unsigned char)c);
        size_t p = lower.find("\r\ncontent-length:");
        if (p != std::string::npos) clen = atoll(lower.c_str() + p + 17);
    }
    int64_t got = (int64_t)(head.size() - (he + 4));
    *ttfbody_us = -1.0;
    if (got > 0 || clen == 0) *ttfbody_us = std::chrono::duration<double, std::micro>(Clock::now() - t0).count();
    while (got < clen) {
        // --discard: the body is dropped inside the kernel (MSG_TRUNC on TCP: no copy to user space) - a sink that does not
        // spend the box's cores on a second copy of every byte; the head is always read for real
        ssize_t n = g_discard ? recv(fd, nullptr, (size_t)std::min<int64_t>((int64_t)1 << 30, clen - got), MSG_TRUNC)
                              : recv(fd, buf.data(), (size_t)std::min<int64_t>((int64_t)buf.size(), clen - got), 0);
        if (n <= 0) return -1;
        if (*ttfbody_us < 0) *ttfbody_us = std::chrono::duration<double, std::micro>(Clock::now() - t0).count();
        got += n;
    }
    return clen;
}

static void client(std::string ip, int port, std::vector<Req> mine, int repeat, int warmup) {
    std::vector<char> buf(g_buf_bytes);
    int fd = dial(ip, port);
    std::vector<double> ttfb, ttfbody;
    uint64_t bytes = 0, reqs = 0, errs = 0, ok2 = 0, oth = 0;
    auto run = [&](int rounds, bool count) {
        for (int k = 0; k < rounds; ++k)
            for (const Req& r : mine) {
                if (fd < 0) fd = dial(ip, port);
                if (fd < 0) { ++errs; continue; }
                double t = 0, tb = 0;
                int st = 0;
                const int64_t n = one_get(fd, r, buf, &t, &st, &tb);
                if (n < 0) { close(fd); fd = -1; if (count) ++errs; continue; }
                if (count) {
                    bytes += (uint64_t)n; ++reqs; ttfb.push_back(t); ttfbody.push_back(tb);
                    if (st >= 200 && st < 300) ++ok2; else ++oth;
                }
            }
    };
    run(warmup, false);
    g_ready++;
    while (!g_go.load()) std::this_thread::yield();
    run(repeat, true);
    if (fd >= 0) close(fd);
    g_bytes += bytes; g_reqs += reqs; g_errors += errs; g_s2xx += ok2; g_sother += oth;
    std::lock_guard<std::mutex> g(g_mu);
    g_ttfb_us.insert(g_ttfb_us.end(), ttfb.begin(), ttfb.end());
    g_ttfbody_us.insert(g_ttfbody_us.end(), ttfbody.begin(), ttfbody.end());
}

int main(int argc, char** argv) {
    std::string target = "127.0.0.1:8080", logf;
    int conns = 1, repeat = 1, warmup = 0;
    for (int i = 1; i < argc; ++i) {
        std::string a = argv[i];
        auto next = [&]() -> const char* { if (i + 1 >= argc) { fprintf(stderr, "missing value for %s\n", a.c_str()); exit(2); } return argv[++i]; };
        if (a == "--target") target = next();
        else if (a == "--conns") conns = atoi(next());
        else if (a == "--log") logf = next();
        else if (a == "--repeat") repeat = atoi(next());
        else if (a == "--warmup") warmup = atoi(next());
        else if (a == "--buf-kib") g_buf_bytes = (size_t)std::max(4, atoi(next())) << 10;
        else if (a == "--discard") g_discard = true;
        else { fprintf(stderr, "nzload --target ip:port --conns N --log FILE [--repeat R] [--warmup W] [--buf-kib K] [--discard]\n"); return 2; }
    }
    if (const char* e = getenv("NZ_LOAD_BUF_KIB")) g_buf_bytes = (size_t)std::max(4, atoi(e)) << 10;
    if (const char* e = getenv("NZ_LOAD_DISCARD")) g_discard = atoi(e) != 0;
    const size_t c = target.rfind(':');
    const std::string ip = target.substr(0, c);
    const int port = atoi(target.c_str() + c + 1);
    std::vector<Req> all;
    std::ifstream in(logf);
    std::string line;
    while (std::getline(in, line)) {
        std::istringstream ss(line);
        Req r;
        if (ss >> r.host >> r.path >> r.range) all.push_back(r);
    }
    if (all.empty() || conns < 1) { fprintf(stderr, "nzload: empty log\n"); return 2; }
    std::vector<std::vector<Req>> per(conns);
    for (size_t i = 0; i < all.size(); ++i) per[i % conns].push_back(all[i]);
    std::vector<std::thread> th;
    for (int t = 0; t < conns; ++t) th.emplace_back(client, ip, port, per[t], repeat, warmup);
    while (g_ready.load() < conns) std::this_thread::yield();
    const auto t0 = Clock::now();
    g_go = true;
    for (auto& t : th) t.join();
    const double secs = std::chrono::duration<double>(Clock::now() - t0).count();
    std::sort(g_ttfb_us.begin(), g_ttfb_us.end());
    std::sort(g_ttfbody_us.begin(), g_ttfbody_us.end());
    auto pct = [&](double p) { return g_ttfb_us.empty() ? 0.0 : g_ttfb_us[std::min(g_ttfb_us.size() - 1, (size_t)(p * g_ttfb_us.size()))]; };
    auto pctb = [&](double p) { return g_ttfbody_us.empty() ? 0.0 : g_ttfbody_us[std::min(g_ttfbody_us.size() - 1, (size_t)(p * g_ttfbody_us.size()))]; };
    printf("{\"bytes\": %llu, \"seconds\": %.6f, \"gbs\": %.4f, \"requests\": %llu, \"errors\": %llu, \"status_2xx\": %llu, \"status_other\": %llu, "
           "\"conns\": %d, \"ttfb_us_p50\": %.1f, \"ttfb_us_p90\": %.1f, \"ttfb_us_p99\": %.1f, \"ttfb_us_min\": %.1f, "
           "\"ttfbody_us_p50\": %.1f, \"ttfbody_us_p90\": %.1f, \"ttfbody_us_p99\": %.1f}\n",
           (unsigned long long)g_bytes.load(), secs, g_bytes.load() / secs / 1e9, (unsigned long long)g_reqs.load(),
           (unsigned long long)g_errors.load(), (unsigned long long)g_s2xx.load(), (unsigned long long)g_sother.load(), conns, pct(0.5),
           pct(0.9), pct(0.99), g_ttfb_us.empty() ? 0.0 : g_ttfb_us.front(), pctb(0.5), pctb(0.9), pctb(0.99));
    return 0;
}
