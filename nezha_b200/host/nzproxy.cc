// nezha_b200/host/nzproxy.cc — the HTTP/1.1 front end of the GPU cache (port-80 wire contract).
//
// Compiled stand-in for the Go `app/proxy` that BASELINE.json:north_star asks for (no `go` in
// this image).  It keeps the client contract of /root/reference/examples/demo/nginx.conf +
// nginx.yaml — plain HTTP virtual-hosted by Host:, GET/HEAD only, `Server: nginx` — so training
// pods are unchanged, and serves cache hits from HBM through the C ABI (include/nezha_gpu.h):
//   request -> nz_http (key, method gate, conditionals, Range) -> plan of literals + slices
//   slices  -> nz_get_submit into two pinned windows per connection (double buffered) -> write()
// Misses are filled from the origin the vhost's proxy_pass names (HTTP/1.0, full object, the
// three proxy_set_header lines of nginx.conf:52-54), admitted only on status 200
// (proxy_cache_valid 200 1d, nginx.conf:57), then served like a hit.
//
// Threading: one blocking thread per connection (cap --max-conns) whose windows go through per-GPU combining lanes, or
// - with --io-threads N - N event-loop workers (epoll), each owning its connections and submitting / polling / sending
// their windows itself, a miss being lent to a helper thread with blocking sockets.  The GPU library is thread safe.  TLS upstream for the GCS vhost (nginx.conf:113) with nginx's defaults (no verify, no SNI).  --access-log writes nginx's `main` format (nginx.conf:16-19).
#include <arpa/inet.h>
#include <errno.h>
#include <fcntl.h>
#include <netdb.h>
#include <netinet/in.h>
#include <netinet/tcp.h>
#include <sched.h>
#include <signal.h>
#include <openssl/evp.h>
#include <openssl/hmac.h>
#include <openssl/ssl.h>
#include <sys/epoll.h>
#include <sys/eventfd.h>
#include <sys/prctl.h>
#include <sys/socket.h>
#include <sys/syscall.h>
#include <sys/time.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <deque>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <map>
#include <memory>
#include <optional>
#include <mutex>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

#include "../../include/nezha_gpu.h"
#include "nz_http.h"

namespace {

struct Options {
    std::string listen_ip = "0.0.0.0";
    int port = 80;
    int device = 0;
    std::vector<int> devices;                      // --devices 0,1,...: one zone per GPU, GETs routed by key hash
    uint64_t arena_bytes = 8ull << 30;
    uint32_t max_objects = 1u << 16;
    uint64_t window_bytes = 4ull << 20;
    double host_rate_gbs = 38.0;            // all bulk windows together stay just under PCIe saturation (first windows then see a short queue)
    std::string access_log;                 // --access-log PATH: one line per request, nginx's `main` format (nginx.conf:16-19)
    bool proxy_ssl_server_name = false;     // nginx default: no SNI towards an https upstream
    bool early_header = true;               // --early-header 0: header goes out with the first body bytes (one segment)
    int host_ctas = 2;                      // CTAs per wave of windows: two saturate the link already
    uint64_t first_window = 16ull << 10;    // small first window keeps time-to-first-byte low (then x8 per window)
    int max_conns = 512;
    int io_threads = 0;                     // > 0: that many event-loop workers (epoll) own the connections; 0 (default): one blocking thread per
                                            // connection, which measured better wherever it was tried (profiles/r02_sweep_http_front_end.txt)
    bool gunzip = false;
    bool verify = false;
    // the conf's private-S3 server block (nginx.conf:138-166, commented out there): bucket, access key, secret
    std::string s3_bucket, s3_access, s3_secret;
    std::map<std::string, std::string> upstream;   // proxy_host -> ip:port override (tests; nginx resolves via DNS)
    std::string pv_dir;                            // durable tier: write-through on fill, bulk reload at start (README.md:123)
    int inactive_s = 3600;                         // proxy_cache_path ... inactive=60m (nginx.conf:38)
    int valid_s = 86400;                           // proxy_cache_valid 200 1d (nginx.conf:57)
};

struct Entry {
    nzh::ObjectMeta meta;
    uint8_t key128[16];
    time_t stored_at;
    int64_t valid_until = 0;               // SPEC H2: absolute seconds (upstream X-Accel-Expires / Cache-Control / Expires, else stored_at + 1 d)
    uint32_t crc32c = 0;
    std::atomic<int64_t> last_access{0};   // monotonic ms; inactive=60m bookkeeping (nginx.conf:38)
    std::atomic<int> inflight{0};          // responses currently streaming this object
};

struct Key16 {
    uint8_t b[16];
    bool operator==(const Key16& o) const { return memcmp(b, o.b, 16) == 0; }
};
struct Key16Hash {
    size_t operator()(const Key16& k) const { size_t v; memcpy(&v, k.b + 8, sizeof v); return v; }
};

constexpr int NZ_MAX_POOLS = 4;
struct Server {
    Options opt;
    nzh::Config cfg = nzh::Config::reference_default();
    nz_ctx* ctx = nullptr;                         // shard 0 (also owns the pinned windows)
    std::vector<nz_ctx*> ctxs;                      // one per GPU; shard = u64le(key128[0:8]) mod nGPU
    nz_ctx* zone_of(const uint8_t key128[16]) const { return ctxs[nz_shard_of(key128, (uint32_t)ctxs.size())]; }
    std::mutex mu;
    std::unordered_map<Key16, std::shared_ptr<Entry>, Key16Hash> index;   // headers blob per object (host side)
    // one fill per KEY at a time (concurrent misses of one object coalesce into one origin fetch - nginx would need
    // proxy_cache_lock for that); misses of different keys fill concurrently
    std::mutex fill_mu;
    std::condition_variable fill_cv;
    std::unordered_map<Key16, int, Key16Hash> filling;
    std::mutex win_mu;
    std::mutex pv_mu;
    // pinned windows, pooled per HOST MEMORY NODE: a window is written by one GPU's copy engine, and pages on the far
    // socket move at about a third of the near one's rate.  pool_of_shard maps a GPU to its node's pool.
    std::vector<nz_buf*> free_windows[NZ_MAX_POOLS];
    nz_ctx* pool_ctx[NZ_MAX_POOLS] = {};
    int n_pools = 1;
    std::vector<int> pool_of_shard;
    std::atomic<uint64_t> boundary{1};
    std::atomic<int> conns{0};
    std::atomic<uint64_t> hits{0}, misses{0}, bytes_out{0};
    int log_fd = -1;
    std::mutex log_mu;
};

bool write_all(int fd, const void* p, size_t n) {
    const char* c = static_cast<const char*>(p);
    while (n) {
        ssize_t w = ::send(fd, c, n, MSG_NOSIGNAL);
        if (w < 0) {
            if (errno == EINTR) continue;
            return false;
        }
        c += w;
        n -= (size_t)w;
    }
    return true;
}

std::string http_date(time_t t) {
    char buf[64];
    struct tm tmv;
    gmtime_r(&t, &tmv);
    strftime(buf, sizeof buf, "%a, %d %b %Y %H:%M:%S GMT", &tmv);
    return buf;
}

const char* status_text(int code) {
    switch (code) {
    case 200: return "OK";
    case 206: return "Partial Content";
    case 304: return "Not Modified";
    case 400: return "Bad Request";
    case 404: return "Not Found";
    case 405: return "Not Allowed";
    case 412: return "Precondition Failed";
    case 416: return "Requested Range Not Satisfiable";
    case 502: return "Bad Gateway";
    case 504: return "Gateway Time-out";
    case 505: return "HTTP Version Not Supported";
    default: return "Error";
    }
}

// ---- request parsing --------------------------------------------------------
// the request head is parsed by the host layer (nzh::parse_head, SPEC P): request-line and header-line rules of nginx,
// first-wins / unique / ignored header lines, 400 and 505 before anything else runs
using Parsed = nzh::Head;
static Parsed parse_head(const std::string& head) { return nzh::parse_head(head); }

std::string lower(std::string s) {
    for (char& c : s) c = (char)tolower((unsigned char)c);
    return s;
}
std::string trim(const std::string& s) {
    size_t a = 0, b = s.size();
    while (a < b && (s[a] == ' ' || s[a] == '\t')) ++a;
    while (b > a && (s[b - 1] == ' ' || s[b - 1] == '\t' || s[b - 1] == '\r')) --b;
    return s.substr(a, b - a);
}

// ---- origin fetch (miss path) ------------------------------------------------
int connect_to(const std::string& hostport, int default_port) {
    std::string host = hostport;
    int port = default_port;
    const size_t c = hostport.rfind(':');
    if (c != std::string::npos) { host = hostport.substr(0, c); port = atoi(hostport.c_str() + c + 1); }
    struct addrinfo hints, *res = nullptr;
    memset(&hints, 0, sizeof hints);
    hints.ai_family = AF_INET;
    hints.ai_socktype = SOCK_STREAM;
    char ps[16];
    snprintf(ps, sizeof ps, "%d", port);
    if (getaddrinfo(host.c_str(), ps, &hints, &res) != 0 || !res) return -1;
    int fd = socket(res->ai_family, res->ai_socktype, res->ai_protocol);
    if (fd >= 0) {
        struct timeval tv = {60, 0};     // proxy_connect/read timeouts default to 60 s
        setsockopt(fd, SOL_SOCKET, SO_RCVTIMEO, &tv, sizeof tv);
        setsockopt(fd, SOL_SOCKET, SO_SNDTIMEO, &tv, sizeof tv);
        if (connect(fd, res->ai_addr, res->ai_addrlen) != 0) { close(fd); fd = -1; }
    }
    freeaddrinfo(res);
    return fd;
}

int64_t now_ms() { return std::chrono::duration_cast<std::chrono::milliseconds>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

bool write_fd_all(int fd, const void* p, size_t n) {
    const char* c = static_cast<const char*>(p);
    while (n) {
        ssize_t w = ::write(fd, c, n);
        if (w < 0) { if (errno == EINTR) continue; return false; }
        c += w; n -= (size_t)w;
    }
    return true;
}

// ---- durable tier (PV): <dir>/<md5hex> extent files + <dir>/index.tsv ------------------------
// index line: md5hex \t stored_len \t raw_len \t crc32c(hex) \t stored_at \t content-type \t etag \t gzip(0/1) \t last-modified [\t valid_until [\t ranges(0/1)]]
std::string tsv(const std::optional<std::string>& s) { return s ? *s : std::string("\x01"); }   // \x01 = absent

void pv_index_append(Server& S, const Entry& e) {
    char hx[33];
    for (int i = 0; i < 16; ++i) snprintf(hx + 2 * i, 3, "%02x", e.key128[i]);
    char num[128];
    snprintf(num, sizeof num, "\t%llu\t%lld\t%08x\t%lld\t", (unsigned long long)e.meta.stored_length, (long long)e.meta.raw_length, e.crc32c,
             (long long)e.stored_at);
    const std::string line = std::string(hx) + num + e.meta.content_type + "\t" + tsv(e.meta.etag) + "\t" + (e.meta.stored_gzip ? "1" : "0") + "\t" +
                             tsv(e.meta.last_modified) + "\t" + std::to_string((long long)e.valid_until) + "\t" + (e.meta.allow_ranges ? "1" : "0") + "\n";
    std::lock_guard<std::mutex> g(S.pv_mu);
    int fd = open((S.opt.pv_dir + "/index.tsv").c_str(), O_WRONLY | O_CREAT | O_APPEND, 0644);
    if (fd >= 0) { write_fd_all(fd, line.data(), line.size()); close(fd); }
}

// The commit inflated a gzip member once (include/nezha_gpu.h): its real inflated length is in the descriptor, and
// NZ_ENC_INFLATABLE says the member checked out.  Anything else is served verbatim (SPEC G4: fail open).
void learn_raw_length(Server& S, Entry& e) {
    e.meta.raw_length = -1;
    if (!e.meta.stored_gzip) return;
    nz_extent ext;
    memset(&ext, 0, sizeof ext);
    uint8_t k[1][16];
    memcpy(k[0], e.key128, 16);
    if (nz_lookup(S.zone_of(e.key128), k, 1, &ext) == NZ_OK && ext.status == NZ_OK && (ext.flags & NZ_ENC_INFLATABLE))
        e.meta.raw_length = (int64_t)ext.raw_length;
}

// cold start: bulk-load every still-valid object of the PV into HBM (cuFile/GDS when available).  The index is
// read first; the files then load on several threads at once (nz_put_file hands out its bounce lanes per call).
void pv_reload(Server& S) {
    FILE* f = fopen((S.opt.pv_dir + "/index.tsv").c_str(), "r");
    if (!f) return;
    char line[4096];
    const time_t now = time(nullptr);
    const auto t0 = std::chrono::steady_clock::now();
    std::vector<std::pair<std::string, std::shared_ptr<Entry>>> todo;     // md5hex, entry (first line of a key wins)
    std::unordered_map<Key16, bool, Key16Hash> seen;
    while (fgets(line, sizeof line, f)) {
        std::vector<std::string> col;
        std::string l(line);
        if (!l.empty() && l.back() == '\n') l.pop_back();
        size_t pos = 0;
        for (;;) {
            const size_t t = l.find('\t', pos);
            col.push_back(l.substr(pos, t == std::string::npos ? std::string::npos : t - pos));
            if (t == std::string::npos) break;
            pos = t + 1;
        }
        if (col.size() < 9 || col.size() > 11 || col[0].size() != 32) continue;
        auto e = std::make_shared<Entry>();
        for (int i = 0; i < 16; ++i) e->key128[i] = (uint8_t)strtoul(col[0].substr(2 * i, 2).c_str(), nullptr, 16);
        e->meta.stored_length = strtoull(col[1].c_str(), nullptr, 10);
        e->meta.raw_length = atoll(col[2].c_str());
        e->crc32c = (uint32_t)strtoul(col[3].c_str(), nullptr, 16);
        e->stored_at = (time_t)atoll(col[4].c_str());
        e->meta.content_type = col[5];
        if (col[6] != "\x01") e->meta.etag = col[6];
        e->meta.stored_gzip = col[7] == "1";
        if (col[8] != "\x01") e->meta.last_modified = col[8];
        e->valid_until = col.size() >= 10 ? atoll(col[9].c_str()) : (int64_t)e->stored_at + S.opt.valid_s;
        e->meta.allow_ranges = col.size() < 11 || col[10] != "0";
        if ((int64_t)now > e->valid_until) continue;                            // expired: refetch on demand
        Key16 k;
        memcpy(k.b, e->key128, 16);
        if (seen.count(k)) continue;
        seen[k] = true;
        todo.emplace_back(col[0], e);
    }
    fclose(f);
    std::atomic<size_t> next{0};
    std::atomic<uint64_t> n_ok{0}, bytes{0}, n_bad{0};
    auto worker = [&] {
        for (;;) {
            const size_t i = next.fetch_add(1);
            if (i >= todo.size()) return;
            const std::shared_ptr<Entry>& e = todo[i].second;
            const int fd = open((S.opt.pv_dir + "/" + todo[i].first).c_str(), O_RDONLY);
            if (fd < 0) continue;
            nz_put* put = nullptr;
            bool ok = nz_put_begin(S.zone_of(e->key128), e->key128, e->meta.stored_length, 0, e->meta.stored_gzip ? NZ_ENC_GZIP : NZ_ENC_IDENTITY, 0, &put) == NZ_OK;
            if (ok) {
                ok = nz_put_file(put, fd, 0, 0, e->meta.stored_length) == NZ_OK;
                uint32_t crc = 0;
                if (ok) ok = nz_put_commit(put, e->crc32c, 1, &crc) == NZ_OK;          // the PV copy must still match its CRC-32C
                else nz_put_abort(put);
            }
            close(fd);
            if (!ok) { ++n_bad; continue; }
            learn_raw_length(S, *e);
            e->last_access = now_ms();
            Key16 k;
            memcpy(k.b, e->key128, 16);
            { std::lock_guard<std::mutex> g(S.mu); S.index[k] = e; }
            ++n_ok; bytes += e->meta.stored_length;
        }
    };
    std::vector<std::thread> th;
    const size_t nth = std::min<size_t>(8, std::max<size_t>(1, todo.size()));
    for (size_t i = 1; i < nth; ++i) th.emplace_back(worker);
    worker();
    for (auto& t : th) t.join();
    const double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    nz_stats st;
    nz_get_stats(S.ctx, &st);
    fprintf(stderr, "nzproxy: reloaded %llu objects (%.2f GiB) from %s in %.2f s (%.2f GB/s; cuFile %.2f GiB, bounce %.2f GiB), %llu rejected\n",
            (unsigned long long)n_ok.load(), bytes.load() / 1073741824.0, S.opt.pv_dir.c_str(), dt, dt > 0 ? bytes.load() / dt / 1e9 : 0.0,
            st.file_cufile_bytes / 1073741824.0, st.file_bounce_bytes / 1073741824.0, (unsigned long long)n_bad.load());
}

// drop one object everywhere (device extent, host index, PV file).  Responses take their reference on an entry under
// S.mu together with the index lookup, so an entry seen here with no response in flight cannot gain one afterwards.
bool drop_entry(Server& S, const std::shared_ptr<Entry>& e) {
    Key16 k;
    memcpy(k.b, e->key128, 16);
    {
        std::lock_guard<std::mutex> g(S.mu);
        if (e->inflight.load() > 0) return false;                   // still streaming: the next pass takes it
        auto it = S.index.find(k);
        if (it == S.index.end() || it->second != e) return false;   // already dropped (or replaced) by someone else
        S.index.erase(it);
    }
    nz_evict(S.zone_of(e->key128), e->key128);
    if (!S.opt.pv_dir.empty()) {
        char hx[33];
        for (int i = 0; i < 16; ++i) snprintf(hx + 2 * i, 3, "%02x", e->key128[i]);
        unlink((S.opt.pv_dir + "/" + hx).c_str());
    }
    return true;
}

// cache manager: `inactive` purge every minute; also called with need_room when the arena is full (LRU)
// need_bytes > 0: the arena of shard `shard` is full - evict least recently used objects OF THAT SHARD until at
// least need_bytes of extents were handed back.  Returns the bytes freed.
uint64_t cache_manager_pass(Server& S, uint64_t need_bytes = 0, uint32_t shard = 0) {
    std::vector<std::shared_ptr<Entry>> victims;
    const int64_t now = now_ms();
    {
        std::lock_guard<std::mutex> g(S.mu);
        std::vector<std::shared_ptr<Entry>> lru;
        for (auto& kv : S.index) {
            const std::shared_ptr<Entry>& e = kv.second;
            if (e->inflight.load() > 0) continue;
            if (now - e->last_access.load() > (int64_t)S.opt.inactive_s * 1000) victims.push_back(e);
            else if (need_bytes && nz_shard_of(e->key128, (uint32_t)S.ctxs.size()) == shard) lru.push_back(e);
        }
        if (need_bytes) {
            std::sort(lru.begin(), lru.end(), [](const auto& a, const auto& b) { return a->last_access.load() < b->last_access.load(); });
            uint64_t got = 0;
            for (auto& e : victims) if (nz_shard_of(e->key128, (uint32_t)S.ctxs.size()) == shard) got += e->meta.stored_length;
            for (auto& e : lru) { if (got >= need_bytes) break; victims.push_back(e); got += e->meta.stored_length; }
        }
    }
    uint64_t freed = 0;
    for (auto& e : victims)
        if (drop_entry(S, e) && nz_shard_of(e->key128, (uint32_t)S.ctxs.size()) == shard) freed += e->meta.stored_length;
    return freed;
}

// ---- upstream connection: plain TCP, or TLS for `proxy_pass https://...` (nginx.conf:113) ------------------
// nginx's defaults for an https upstream are mirrored: no certificate verification (proxy_ssl_verify off) and no
// SNI (proxy_ssl_server_name off) unless --proxy-ssl-server-name is given.
struct UpConn {
    int fd = -1;
    SSL* ssl = nullptr;
    bool write_all_(const void* p, size_t n) {
        if (!ssl) return write_all(fd, p, n);
        const char* c = static_cast<const char*>(p);
        while (n) {
            const int w = SSL_write(ssl, c, (int)std::min<size_t>(n, 1 << 30));
            if (w <= 0) return false;
            c += w; n -= (size_t)w;
        }
        return true;
    }
    ssize_t read_(void* p, size_t n) {
        if (!ssl) return recv(fd, p, n, 0);
        const int r = SSL_read(ssl, p, (int)std::min<size_t>(n, 1 << 30));
        return r > 0 ? r : (SSL_get_error(ssl, r) == SSL_ERROR_ZERO_RETURN ? 0 : -1);
    }
    void close_() {
        if (ssl) { SSL_shutdown(ssl); SSL_free(ssl); ssl = nullptr; }
        if (fd >= 0) { close(fd); fd = -1; }
    }
};
SSL_CTX* upstream_tls_ctx() {
    static SSL_CTX* ctx = [] {
        SSL_CTX* c = SSL_CTX_new(TLS_client_method());
        if (c) SSL_CTX_set_verify(c, SSL_VERIFY_NONE, nullptr);
        return c;
    }();
    return ctx;
}
bool b64_decode(const std::string& in, std::vector<uint8_t>& out) {
    static const std::string tb = "ABCDEFGHIJKLMNOPQRSTUVWXYZabcdefghijklmnopqrstuvwxyz0123456789+/";
    uint32_t acc = 0;
    int bits = 0;
    out.clear();
    for (char ch : in) {
        if (ch == '=') break;
        const size_t v = tb.find(ch);
        if (v == std::string::npos) return false;
        acc = (acc << 6) | (uint32_t)v;
        bits += 6;
        if (bits >= 8) { bits -= 8; out.push_back((uint8_t)(acc >> bits)); }
    }
    return true;
}

// returns 200 = admitted and indexed (fills `out`); -1 = the upstream response was not cacheable (SPEC H2) and has been
// relayed to the client as it came, uncached, the way nginx proxies such a response (the connection then closes);
// anything else = the status of the stock error page to send
int fill_from_origin(Server& S, int vhost, const std::string& host, const std::string& target, const std::string& client_ip,
                     const uint8_t key128[16], std::shared_ptr<Entry>& out, int client_fd, bool head_only) {
    const nzh::VHost& vh = S.cfg.vhosts[vhost];
    auto it = S.opt.upstream.find(vh.proxy_host);
    // --upstream host=ip:port keeps the vhost's scheme out of it (tests: plain origin); host=https://ip:port or no override
    // on an https vhost means TLS
    std::string addr = it != S.opt.upstream.end() ? it->second : vh.proxy_host;
    bool tls = it == S.opt.upstream.end() ? vh.upstream_tls : false;
    if (addr.rfind("https://", 0) == 0) { tls = true; addr = addr.substr(8); }
    else if (addr.rfind("http://", 0) == 0) { tls = false; addr = addr.substr(7); }
    UpConn up;
    up.fd = connect_to(addr, tls ? 443 : 80);
    if (up.fd < 0) return 502;
    if (tls) {
        SSL_CTX* cx = upstream_tls_ctx();
        up.ssl = cx ? SSL_new(cx) : nullptr;
        if (!up.ssl) { up.close_(); return 502; }
        SSL_set_fd(up.ssl, up.fd);
        if (S.opt.proxy_ssl_server_name) SSL_set_tlsext_host_name(up.ssl, vh.proxy_host.substr(0, vh.proxy_host.find(':')).c_str());
        if (SSL_connect(up.ssl) != 1) { up.close_(); return 502; }
    }
    const int fd = up.fd;
    (void)fd;
    // nginx talks HTTP/1.0 to the upstream and strips Range/If-* when caching; nginx.conf:52-54 headers
    std::string uri = target;
    if (uri.rfind("http://", 0) == 0 || uri.rfind("https://", 0) == 0) {
        const size_t s = uri.find('/', uri.find("://") + 3);
        uri = s == std::string::npos ? "/" : uri.substr(s);
    }
    std::string req;
    if (vh.server_name == "s3.amazonaws.com" && !S.opt.s3_bucket.empty()) {
        // private S3 (nginx.conf:138-166): AWS signature version 2.  The conf signs
        //   "$request_method\n\n\n\nx-amz-date:${now}\n/$bucket/$url_full"
        // with HMAC-SHA1 under the secret key, base64, and sends `Authorization: AWS access:signature`, `x-amz-date`
        // and `Host: $bucket.s3.amazonaws.com` over HTTP/1.1.  Its $url_full is "$1" of a regex that the block never
        // has (so the conf as written signs "/bucket/" for every object); here it is the object's path without the
        // leading slash, which is what S3 verifies.
        const std::string now = http_date(time(nullptr));
        const std::string path = uri.substr(0, uri.find('?'));
        const std::string to_sign = "GET\n\n\n\nx-amz-date:" + now + "\n/" + S.opt.s3_bucket + path;
        unsigned char mac[EVP_MAX_MD_SIZE];
        unsigned int mac_len = 0;
        HMAC(EVP_sha1(), S.opt.s3_secret.data(), (int)S.opt.s3_secret.size(), (const unsigned char*)to_sign.data(), to_sign.size(), mac, &mac_len);
        char b64[64];
        EVP_EncodeBlock((unsigned char*)b64, mac, (int)mac_len);
        req = "GET " + uri + " HTTP/1.1\r\nHost: " + S.opt.s3_bucket + ".s3.amazonaws.com\r\nx-amz-date: " + now + "\r\nAuthorization: AWS " + S.opt.s3_access + ":" +
              b64 + "\r\nConnection: close\r\n\r\n";
    } else {
        req = "GET " + uri + " HTTP/1.0\r\nX-Real-IP: " + client_ip + "\r\nHost: " + host + "\r\nX-Forwarded-For: " + client_ip +
              "\r\nConnection: close\r\n\r\n";
    }
    if (!up.write_all_(req.data(), req.size())) { up.close_(); return 502; }
    std::string head;
    std::vector<char> buf(1 << 16);
    size_t body_at = std::string::npos;
    while (body_at == std::string::npos) {
        ssize_t r = up.read_(buf.data(), buf.size());
        if (r <= 0) { up.close_(); return r == 0 ? 502 : 504; }
        head.append(buf.data(), (size_t)r);
        body_at = head.find("\r\n\r\n");
        if (head.size() > (1 << 20)) { up.close_(); return 502; }
    }
    std::string rest = head.substr(body_at + 4);
    head.resize(body_at + 2);
    int status = 0;
    if (sscanf(head.c_str(), "HTTP/%*d.%*d %d", &status) != 1) { up.close_(); return 502; }
    nzh::ObjectMeta meta;
    int64_t clen = -1;
    bool storable = true;                          // representable in the store (identity or gzip, known length)
    bool saw_accept_ranges = false;
    std::vector<std::pair<std::string, std::string>> hv;    // upstream headers in arrival order (admission, relay)
    bool have_crc = false;
    uint32_t want_crc = 0;
    std::vector<uint8_t> want_md5;                 // Content-MD5 (RFC 1864; Azure) or x-goog-hash: md5= (GCS): base64 of the body's MD5
    size_t pos = head.find("\r\n") + 2;
    while (pos < head.size()) {
        const size_t e = head.find("\r\n", pos);
        if (e == std::string::npos) break;
        const std::string h = head.substr(pos, e - pos);
        pos = e + 2;
        const size_t c = h.find(':');
        if (c == std::string::npos) continue;
        const std::string name = lower(h.substr(0, c)), val = trim(h.substr(c + 1));
        hv.emplace_back(h.substr(0, c), val);
        if (name == "content-length") clen = atoll(val.c_str());
        else if (name == "content-type") meta.content_type = val;
        else if (name == "etag") meta.etag = val;
        else if (name == "last-modified") meta.last_modified = val;
        else if (name == "accept-ranges") saw_accept_ranges = true;      // R0: nginx allows ranges on the cached copy only then
        else if (name == "content-encoding") { if (lower(val) == "gzip") meta.stored_gzip = true; else storable = false; }
        else if (name == "content-md5") { std::vector<uint8_t> m; if (b64_decode(val, m) && m.size() == 16) want_md5 = m; }
        else if (name == "x-goog-hash") {                                         // GCS: crc32c=<base64 of the big-endian u32>[, md5=<base64>]
            const size_t m5 = val.find("md5=");
            if (m5 != std::string::npos) {
                std::vector<uint8_t> m;
                const size_t e5 = val.find(',', m5);
                if (b64_decode(val.substr(m5 + 4, e5 == std::string::npos ? std::string::npos : e5 - m5 - 4), m) && m.size() == 16) want_md5 = m;
            }
            const size_t c32 = val.find("crc32c=");
            if (c32 != std::string::npos && val.size() >= c32 + 7 + 8) {
                static const std::string tb = "ABCDEFGHIJKLMNOPQRSTUVWXYZabcdefghijklmnopqrstuvwxyz0123456789+/";
                uint64_t acc = 0;
                int nb = 0;
                for (size_t i = c32 + 7; i < val.size() && nb < 6; ++i) {
                    const size_t v = tb.find(val[i]);
                    if (v == std::string::npos) break;
                    acc = (acc << 6) | (uint64_t)v;
                    ++nb;
                }
                if (nb == 6) { want_crc = (uint32_t)(acc >> 4); have_crc = true; }             // 36 bits read, 32 are data
            }
        }
    }
    // SPEC H2: status 200 only; Set-Cookie / Vary: * never; X-Accel-Expires > Cache-Control > Expires > proxy_cache_valid 200 1d
    const nzh::Admission adm = nzh::admission(status, hv, (int64_t)time(nullptr), S.opt.valid_s);
    // An object that can never fit the shard's arena is not worth flushing the cache for (and an origin may claim any length)
    const bool fits = clen >= 0 && (uint64_t)clen + (1u << 20) <= S.opt.arena_bytes;
    if (!adm.cacheable || !storable || !fits) {
        // relay uncached: upstream status line and headers as they came (hop-by-hop ones replaced), body until its end
        const size_t sl_end = head.find("\r\n");
        std::string sline = head.substr(0, sl_end);
        const size_t sp = sline.find(' ');
        std::string rh = "HTTP/1.1" + (sp == std::string::npos ? std::string(" 502 Bad Gateway") : sline.substr(sp)) + "\r\nServer: nginx\r\nDate: " + http_date(time(nullptr)) + "\r\n";
        for (const auto& kv : hv) {
            const std::string n = lower(kv.first);
            if (n == "server" || n == "date" || n == "connection" || n == "keep-alive" || n == "proxy-connection" || n == "transfer-encoding") continue;
            rh += kv.first + ": " + kv.second + "\r\n";
        }
        rh += "Connection: close\r\n\r\n";
        bool ok = write_all(client_fd, rh.data(), rh.size());
        uint64_t sent = 0;
        if (ok && !head_only) {
            if (!rest.empty()) { ok = write_all(client_fd, rest.data(), rest.size()); sent += rest.size(); }
            while (ok && (clen < 0 || sent < (uint64_t)clen)) {
                const ssize_t r = up.read_(buf.data(), buf.size());
                if (r <= 0) break;
                ok = write_all(client_fd, buf.data(), (size_t)r);
                sent += (size_t)r;
            }
        }
        up.close_();
        S.bytes_out += sent;
        return -1;
    }
    // stream the body into the arena through a pinned window
    nz_put* put = nullptr;
    int32_t prc = NZ_ENOMEM;
    const uint32_t shard = nz_shard_of(key128, (uint32_t)S.ctxs.size());
    for (int attempt = 0; attempt < 8 && prc == NZ_ENOMEM; ++attempt) {
        // raw_len 0: nz_put_commit inflates a gzip member once and records its real inflated length
        prc = nz_put_begin(S.ctxs[shard], key128, (uint64_t)clen, 0, meta.stored_gzip ? NZ_ENC_GZIP : NZ_ENC_IDENTITY, 0, &put);
        // arena of THIS shard full: hand back at least the bytes needed, least recently used first; give up when nothing was freed
        if (prc == NZ_ENOMEM && cache_manager_pass(S, (uint64_t)clen + (64u << 10), shard) == 0) break;
    }
    if (prc != NZ_OK) {
        up.close_();
        return 502;
    }
    nz_buf* win = nullptr;
    const int pool = S.pool_of_shard[shard];
    {
        std::lock_guard<std::mutex> g(S.win_mu);
        if (!S.free_windows[pool].empty()) { win = S.free_windows[pool].back(); S.free_windows[pool].pop_back(); }
    }
    if (!win && nz_buf_alloc(S.pool_ctx[pool], S.opt.window_bytes, NZ_BUF_PINNED_HOST, &win) != NZ_OK) { nz_put_abort(put); up.close_(); return 502; }
    char* wp = static_cast<char*>(nz_buf_ptr(win));
    uint64_t off = 0;
    bool ok = true;
    int pv_fd = -1;
    std::string pv_tmp, pv_final;
    if (!S.opt.pv_dir.empty()) {                       // write-through to the durable tier (temp file -> rename, like nginx)
        char hx[33];
        for (int i = 0; i < 16; ++i) snprintf(hx + 2 * i, 3, "%02x", key128[i]);
        pv_final = S.opt.pv_dir + "/" + hx;
        pv_tmp = pv_final + ".part";
        pv_fd = open(pv_tmp.c_str(), O_WRONLY | O_CREAT | O_TRUNC, 0644);
    }
    size_t have = std::min<size_t>(rest.size(), (size_t)clen);
    if (have) memcpy(wp, rest.data(), have);
    EVP_MD_CTX* md = nullptr;
    if (!want_md5.empty()) { md = EVP_MD_CTX_new(); if (md && EVP_DigestInit_ex(md, EVP_md5(), nullptr) != 1) { EVP_MD_CTX_free(md); md = nullptr; } }
    while (ok && off < (uint64_t)clen) {
        while (have < S.opt.window_bytes && off + have < (uint64_t)clen) {
            ssize_t r = up.read_(wp + have, std::min<uint64_t>(S.opt.window_bytes - have, (uint64_t)clen - off - have));
            if (r <= 0) { ok = false; break; }
            have += (size_t)r;
        }
        if (!ok) break;
        if (md) EVP_DigestUpdate(md, wp, have);
        if (pv_fd >= 0 && !write_fd_all(pv_fd, wp, have)) { close(pv_fd); unlink(pv_tmp.c_str()); pv_fd = -1; }
        if (nz_put_write(put, wp, off, have) != NZ_OK) { ok = false; break; }
        if (nz_put_sync(put) != NZ_OK) { ok = false; break; }   // the window is reused: its H2D copy must have left it
        off += have;
        have = 0;
    }
    up.close_();
    {
        std::lock_guard<std::mutex> g(S.win_mu);
        S.free_windows[pool].push_back(win);
    }
    if (md) {                                                   // the origin's MD5 of the body must match what arrived
        uint8_t dg[EVP_MAX_MD_SIZE];
        unsigned dl = 0;
        EVP_DigestFinal_ex(md, dg, &dl);
        EVP_MD_CTX_free(md);
        if (ok && (dl != 16 || memcmp(dg, want_md5.data(), 16) != 0)) ok = false;
    }
    if (!ok) { nz_put_abort(put); if (pv_fd >= 0) { close(pv_fd); unlink(pv_tmp.c_str()); } return 502; }
    uint32_t crc = 0;
    if (nz_put_commit(put, want_crc, have_crc ? 1 : 0, &crc) != NZ_OK) {          // NZ_ECORRUPT: origin's CRC-32C does not match
        if (pv_fd >= 0) { close(pv_fd); unlink(pv_tmp.c_str()); }
        return 502;
    }
    meta.stored_length = (uint64_t)clen;
    meta.raw_length = -1;
    meta.allow_ranges = saw_accept_ranges;
    auto e = std::make_shared<Entry>();
    e->meta = meta;
    memcpy(e->key128, key128, 16);
    e->stored_at = time(nullptr);
    e->valid_until = adm.valid_until;
    e->last_access = now_ms();
    e->crc32c = crc;
    learn_raw_length(S, *e);               // gzip: the inflated length the commit found (stays -1 if the member does not inflate: served verbatim, SPEC G4)
    if (pv_fd >= 0) {
        close(pv_fd);
        if (rename(pv_tmp.c_str(), pv_final.c_str()) == 0) pv_index_append(S, *e);
    }
    {
        std::lock_guard<std::mutex> g(S.mu);
        Key16 k;
        memcpy(k.b, key128, 16);
        S.index[k] = e;
    }
    out = e;
    return 200;
}

// ---- wave combiner ---------------------------------------------------------------
// Every connection streams its response through windows; submitting each window as its own
// wave would queue hundreds of tiny kernel pairs on the device.  Instead connection threads
// enqueue their window requests and a dispatcher thread turns whatever is queued into ONE wave
// (one plan + one gather kernel for all connections).  Two lanes: `urgent` carries the small
// first window of every response on a high-priority stream (time to first byte), `bulk` the rest.
struct WaveItem {
    nz_get_req rq;
    nz_seg sg;
    nz_get_res rs;
    int32_t rc = NZ_ECUDA;
    std::atomic<int> done{0};
    int64_t t_push = 0, t_done = 0;        // ns (steady clock), NZ_PROXY_TRACE only
};

// NZ_PROXY_TRACE=1: where a window's latency goes (queue -> submit -> device -> wake-up), printed once a second
static bool g_trace = false;
inline int64_t now_ns() { return std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
struct LaneTrace {
    std::atomic<uint64_t> items{0}, waves{0}, queue_ns{0}, submit_ns{0}, device_ns{0}, wake_ns{0}, bytes{0}, gpu_ns{0};
};
static LaneTrace g_tr[2];                  // 0 = urgent, 1 = bulk
static std::atomic<uint64_t> g_first_ns{0}, g_first_n{0};   // request parsed -> first body window written to the socket

static std::atomic<int> g_spinners{0};
static int g_spin_cap = 4;                   // set in main from the CPUs this process may actually use

struct Combiner {
    nz_ctx* ctx = nullptr;                         // shard 0 (also owns the pinned windows)
    std::vector<nz_ctx*> ctxs;                      // one per GPU; shard = u64le(key128[0:8]) mod nGPU
    nz_ctx* zone_of(const uint8_t key128[16]) const { return ctxs[nz_shard_of(key128, (uint32_t)ctxs.size())]; }
    std::mutex mu;
    std::condition_variable cv;
    std::deque<WaveItem*> q;
    uint64_t cap_bytes = 64ull << 20;     // payload per wave: bounds how long a wave holds the PCIe link
    int linger_us = 0;
    bool stop = false;

    std::atomic<int> pending{0};
    int spin_us = 0;
    int lane = 1;                           // trace slot
    void push(WaveItem* it) {
        if (g_trace) it->t_push = now_ns();
        { std::lock_guard<std::mutex> g(mu); q.push_back(it); pending.store(1, std::memory_order_release); }
        cv.notify_one();
    }
    static void wait(WaveItem* it) {
        // a window is ready in tens of microseconds: poll briefly before paying a futex sleep + wake-up - but only as many
        // threads at a time as the box can spare (g_spin_cap: an eighth of the CPUs the cgroup grants); the rest block
        if (g_spinners.fetch_add(1, std::memory_order_relaxed) < g_spin_cap) {
            const auto t0 = std::chrono::steady_clock::now();
            while (!it->done.load(std::memory_order_acquire)) {
                __builtin_ia32_pause();
                if (std::chrono::steady_clock::now() - t0 > std::chrono::microseconds(150)) break;
            }
        }
        g_spinners.fetch_sub(1, std::memory_order_relaxed);
        while (!it->done.load(std::memory_order_acquire)) it->done.wait(0, std::memory_order_acquire);
    }
    void run() {
        std::vector<WaveItem*> batch;
        std::vector<nz_get_req> reqs;
        std::vector<nz_get_res> res;
        for (;;) {
            batch.clear();
            {
                std::unique_lock<std::mutex> lk(mu);
                if (q.empty() && spin_us > 0) {          // hot dispatcher: poll the queue for a while before sleeping
                    lk.unlock();
                    const auto t0 = std::chrono::steady_clock::now();
                    for (;;) {
                        __builtin_ia32_pause();
                        if (std::chrono::steady_clock::now() - t0 > std::chrono::microseconds(spin_us)) break;
                        if (pending.load(std::memory_order_acquire)) break;
                    }
                    lk.lock();
                }
                cv.wait(lk, [&] { return stop || !q.empty(); });
                if (stop) return;
                if (linger_us > 0 && q.size() < 8) cv.wait_for(lk, std::chrono::microseconds(linger_us));
                uint64_t bytes = 0;
                while (!q.empty() && batch.size() < 512 && (batch.empty() || bytes + q.front()->sg.len <= cap_bytes)) {
                    bytes += q.front()->sg.len;
                    batch.push_back(q.front());
                    q.pop_front();
                }
                if (q.empty()) pending.store(0, std::memory_order_release);
            }
            reqs.resize(batch.size());
            res.resize(batch.size());
            for (size_t i = 0; i < batch.size(); ++i) { reqs[i] = batch[i]->rq; reqs[i].segs = &batch[i]->sg; }
            nz_ticket* t = nullptr;
            const int64_t t_take = g_trace ? now_ns() : 0;
            int32_t rc = nz_get_submit(ctx, reqs.data(), (uint32_t)reqs.size(), nullptr, &t);
            const int64_t t_sub = g_trace ? now_ns() : 0;
            if (rc == NZ_OK) rc = nz_get_wait(t, -1, res.data());
            const int64_t t_dn = g_trace ? now_ns() : 0;
            float gpu_ms = 0.f;
            if (g_trace && t && nz_ticket_elapsed_ms(t, &gpu_ms) != NZ_OK) gpu_ms = 0.f;     // needs NZ_PROFILE=1
            if (t) nz_ticket_free(t);
            if (g_trace) {
                g_tr[lane].gpu_ns += (uint64_t)(gpu_ms * 1e6) * batch.size();
                LaneTrace& T = g_tr[lane];
                T.waves++; T.items += batch.size();
                T.submit_ns += (uint64_t)(t_sub - t_take) * batch.size(); T.device_ns += (uint64_t)(t_dn - t_sub) * batch.size();
                for (WaveItem* it : batch) { T.queue_ns += (uint64_t)(t_take - it->t_push); T.bytes += it->sg.len; it->t_done = t_dn; }
            }
            for (size_t i = 0; i < batch.size(); ++i) {
                batch[i]->rc = rc;
                if (rc == NZ_OK) batch[i]->rs = res[i];
                batch[i]->done.store(1, std::memory_order_release);
                batch[i]->done.notify_one();
            }
        }
    }
};

static std::vector<std::unique_ptr<Combiner>> g_urgent, g_bulk;   // one pair of lanes per GPU
static std::atomic<int> g_direct{0};          // first windows being submitted straight from connection threads
static int g_direct_max = 4;                  // env NZ_PROXY_DIRECT (0 = always through the urgent combiner)
static std::atomic<uint64_t> g_direct_ns{0}, g_direct_n{0};

// ---- serving -------------------------------------------------------------------
// a connection's pinned windows: one pair per host memory node it has been served from (taken on first use)
struct ConnWin { nz_buf* w[NZ_MAX_POOLS][2] = {}; };
nz_buf** conn_windows(Server& S, ConnWin& cw, uint32_t shard);
void give_windows(Server& S, ConnWin& cw);

bool send_plan(Server& S, int fd, const nzh::Plan& plan, const Entry* ent, bool keep_alive, ConnWin& cw, int64_t t_req = 0) {
    std::string h = std::string("HTTP/1.1 ") + std::to_string(plan.status) + " " + status_text(plan.status) + "\r\nServer: nginx\r\nDate: " +
                    http_date(time(nullptr)) + "\r\n";
    for (const auto& kv : plan.headers) h += kv.first + ": " + kv.second + "\r\n";
    h += keep_alive ? "Connection: keep-alive\r\n" : "Connection: close\r\n";
    h += "\r\n";
    if (plan.head || plan.body.empty()) return write_all(fd, h.data(), h.size());
    // literals are tiny; slices stream through the two pinned windows.  The header goes out with
    // the first bytes of the body.
    std::string pending = h;
    const uint32_t flags = (plan.inflated ? NZ_GET_INFLATE : 0) | (S.opt.verify ? NZ_GET_VERIFY : 0);
    nz_buf** win = conn_windows(S, cw, nz_shard_of(ent->key128, (uint32_t)S.ctxs.size()));   // windows next to the GPU that holds the object
    if (!win) return false;
    bool first = true;                                   // first window of the RESPONSE: small and urgent
    for (const nzh::Part& part : plan.body) {
        if (part.literal) { pending += part.bytes; continue; }
        uint64_t off = part.off, left = part.len;
        WaveItem item[2];
        bool live[2] = {false, false};
        int cur = 0;
        uint64_t ramp = S.opt.first_window;
        // status line + headers (+ a multipart preamble) are known from host-side metadata: they leave before the
        // first window is even requested from the GPU, like a proxy forwarding upstream headers as they arrive
        if (first && !pending.empty() && S.opt.early_header) {
            if (!write_all(fd, pending.data(), pending.size())) return false;
            pending.clear();
        }
        auto submit = [&](int slot) {
            live[slot] = false;
            if (!left) return;
            const uint64_t n = std::min<uint64_t>(left, std::min<uint64_t>(ramp, S.opt.window_bytes));
            ramp = std::min<uint64_t>(ramp * 8, S.opt.window_bytes);          // 1st window small, then grow quickly
            WaveItem& it = item[slot];
            memset(&it.rq, 0, sizeof it.rq);
            memcpy(it.rq.key, ent->key128, 16);
            it.rq.flags = flags | (first ? NZ_GET_URGENT : 0u);
            it.sg = nz_seg{off, n, 0};
            it.rq.nseg = 1; it.rq.segs = &it.sg; it.rq.dst = win[slot];
            it.rc = NZ_ECUDA;
            it.done.store(0, std::memory_order_relaxed);
            {
                const uint32_t sh = nz_shard_of(ent->key128, (uint32_t)S.ctxs.size());
                // first window, urgent lane quiet: submit from this thread (no queue hop, no wake-up); a burst of
                // simultaneous requests still goes through the combiner and becomes one wave
                if (first && g_direct_max > 0 && g_direct.fetch_add(1) < g_direct_max) {
                    const int64_t t0 = g_trace ? now_ns() : 0;
                    nz_ticket* t = nullptr;
                    it.rc = nz_get_submit(S.ctxs[sh], &it.rq, 1, nullptr, &t);
                    if (it.rc == NZ_OK) it.rc = nz_get_wait(t, -1, &it.rs);
                    if (t) nz_ticket_free(t);
                    g_direct--;
                    it.done.store(1, std::memory_order_release);
                    if (g_trace) { it.t_done = now_ns(); g_direct_ns += (uint64_t)(it.t_done - t0); g_direct_n++; }
                } else {
                    if (first && g_direct_max > 0) g_direct--;
                    (first ? g_urgent[sh] : g_bulk[sh])->push(&it);
                }
            }
            first = false;
            live[slot] = true;
            off += n;
            left -= n;
        };
        submit(0);
        while (live[cur]) {
            submit(cur ^ 1);                                                  // next window in flight while this one is sent
            Combiner::wait(&item[cur]);
            if (g_trace) g_tr[(item[cur].rq.flags & NZ_GET_URGENT) ? 0 : 1].wake_ns += (uint64_t)(now_ns() - item[cur].t_done);
            live[cur] = false;
            bool ok = item[cur].rc == NZ_OK && item[cur].rs.status == NZ_OK;
            if (!ok) fprintf(stderr, "nzproxy: window failed: submit/wait rc=%d (%s) status=%d [%s]\n", item[cur].rc, nz_strerror(item[cur].rc), item[cur].rs.status, nz_last_error());
            if (ok && !pending.empty()) { ok = write_all(fd, pending.data(), pending.size()); pending.clear(); }
            if (ok) ok = write_all(fd, nz_buf_ptr(win[cur]), item[cur].sg.len);
            if (g_trace && t_req && (item[cur].rq.flags & NZ_GET_URGENT)) { g_first_ns += (uint64_t)(now_ns() - t_req); g_first_n++; }
            if (!ok) {
                if (live[cur ^ 1]) Combiner::wait(&item[cur ^ 1]);            // never leave a queued item pointing at this stack frame
                return false;                                                 // mid-body failure: drop the connection
            }
            S.bytes_out += item[cur].sg.len;
            cur ^= 1;
        }
    }
    if (!pending.empty() && !write_all(fd, pending.data(), pending.size())) return false;
    return true;
}

// access_log ... main (nginx.conf:16-19):
//   $remote_addr - $remote_user [$time_local] "$request" $status $body_bytes_sent "$http_referer" "$http_user_agent" "$http_x_forwarded_for"
void access_log_line(Server& S, const std::string& client_ip, const Parsed& pr, int status, uint64_t body_bytes) {
    if (S.log_fd < 0) return;
    char ts[64];
    time_t now = time(nullptr);
    struct tm tmv;
    localtime_r(&now, &tmv);
    strftime(ts, sizeof ts, "%d/%b/%Y:%H:%M:%S %z", &tmv);
    auto dash = [](const std::string& v) { return v.empty() ? std::string("-") : v; };
    const std::string line = client_ip + " - - [" + ts + "] \"" + pr.request_line + "\" " + std::to_string(status) + " " + std::to_string(body_bytes) + " \"" +
                             dash(pr.referer) + "\" \"" + dash(pr.user_agent) + "\" \"" + dash(pr.x_forwarded_for) + "\"\n";
    std::lock_guard<std::mutex> g(S.log_mu);
    write_fd_all(S.log_fd, line.data(), line.size());
}

// One request whose head is already in hand, served with blocking sockets: the miss path (origin fetch, uncached
// relay) and the legacy thread-per-connection mode (--io-threads 0).  Returns whether the connection stays open.
bool serve_one(Server& S, int fd, const std::string& client_ip, const std::string& head, ConnWin& cw) {
    const int64_t t_req = g_trace ? now_ns() : 0;
    Parsed pr = parse_head(head);
    nzh::Plan plan;
    std::shared_ptr<Entry> ent;
    bool keep = pr.keep_alive;
    auto stock = [&](int code) { plan = nzh::error_page(code, !pr.error && pr.rq.method == "HEAD"); };
    if (pr.error) {
        stock(pr.error);
        keep = false;
    } else if (!nzh::method_allowed(pr.rq.method) || (pr.rq.method != "GET" && pr.rq.method != "HEAD")) {
        stock(405);
    } else {
        int vhost = 0;
        auto key = nzh::cache_key(S.cfg, pr.rq.host, pr.rq.target, &vhost);
        if (!key) {
            stock(400);
            keep = false;
        } else {
            Key16 k;
            nzh::md5(key->data(), key->size(), k.b);
            // lookup and reference in one critical section: the cache manager only drops entries it sees unreferenced
            // under the same lock, so a response never streams an object that is being evicted (or its successor)
            auto lookup_ref = [&]() {
                std::lock_guard<std::mutex> g(S.mu);
                auto it = S.index.find(k);
                if (it == S.index.end()) return;
                if ((int64_t)time(nullptr) > it->second->valid_until) return;      // SPEC H2: expired -> MISS (refetched below)
                ent = it->second;
                ent->inflight++;
            };
            lookup_ref();
            int st = 200;
            if (!ent) {
                S.misses++;
                {   // one fill per key: later misses of the same key wait for the first and then find the entry
                    std::unique_lock<std::mutex> g(S.fill_mu);
                    S.fill_cv.wait(g, [&] { return S.filling.count(k) == 0; });
                    S.filling[k] = 1;
                }
                lookup_ref();                                                       // filled by a concurrent request while we waited
                if (!ent) {
                    {   // an expired (or otherwise stale) entry of this key goes first; if a response still streams it, serve as uncached miss
                        std::shared_ptr<Entry> old;
                        { std::lock_guard<std::mutex> g(S.mu); auto it = S.index.find(k); if (it != S.index.end()) old = it->second; }
                        if (old && !drop_entry(S, old)) st = 504;
                    }
                    if (st == 200) {
                        // $host for the upstream Host header = request host (proxy_set_header Host $host)
                        std::string host = key->substr(0, key->find('/'));
                        std::shared_ptr<Entry> fresh;
                        st = fill_from_origin(S, vhost, host, pr.rq.target, client_ip, k.b, fresh, fd, pr.rq.method == "HEAD");
                        if (st == 200) lookup_ref();
                    }
                }
                {
                    std::lock_guard<std::mutex> g(S.fill_mu);
                    S.filling.erase(k);
                }
                S.fill_cv.notify_all();
            } else {
                S.hits++;
            }
            if (st == -1) {                       // relayed uncached: the response is on the wire already
                access_log_line(S, client_ip, pr, 200, 0);
                return false;
            }
            if (!ent) {
                stock(st == 200 ? 502 : st);
            } else {
                ent->last_access = now_ms();
                nzh::Config cfg = S.cfg;
                cfg.gunzip = S.opt.gunzip;
                plan = nzh::plan_response(cfg, pr.rq, ent->meta, S.boundary.fetch_add(1));
            }
        }
    }
    // a request that carries a body is answered and the connection closed: the body is never read (nginx would
    // discard it; parsing it as the next request must not happen)
    if (!pr.error && pr.has_body) keep = false;
    const bool sent = send_plan(S, fd, plan, ent.get(), keep, cw, t_req);
    access_log_line(S, client_ip, pr, plan.status, plan.head ? 0 : plan.body_length());
    if (ent) ent->inflight--;
    return sent && keep;
}

void tune_socket(int fd) {
    int one = 1;
    setsockopt(fd, IPPROTO_TCP, TCP_NODELAY, &one, sizeof one);
    int sndbuf = 4 << 20;
    setsockopt(fd, SOL_SOCKET, SO_SNDBUF, &sndbuf, sizeof sndbuf);
    // nginx defaults: client_header_timeout 60s / keepalive_timeout 65s (one receive timeout covers both here), send_timeout 60s
    struct timeval rtv = {65, 0}, stv = {60, 0};
    setsockopt(fd, SOL_SOCKET, SO_RCVTIMEO, &rtv, sizeof rtv);
    setsockopt(fd, SOL_SOCKET, SO_SNDTIMEO, &stv, sizeof stv);
}

nz_buf** conn_windows(Server& S, ConnWin& cw, uint32_t shard) {
    const int pool = S.pool_of_shard[shard];
    nz_buf** win = cw.w[pool];
    if (win[0] && win[1]) return win;
    {
        std::lock_guard<std::mutex> g(S.win_mu);
        std::vector<nz_buf*>& fr = S.free_windows[pool];
        for (int i = 0; i < 2 && !fr.empty(); ++i) if (!win[i]) { win[i] = fr.back(); fr.pop_back(); }
    }
    for (int i = 0; i < 2; ++i)
        if (!win[i] && nz_buf_alloc(S.pool_ctx[pool], S.opt.window_bytes, NZ_BUF_PINNED_HOST, &win[i]) != NZ_OK) win[i] = nullptr;
    return win[0] && win[1] ? win : nullptr;
}
void give_windows(Server& S, ConnWin& cw) {
    std::lock_guard<std::mutex> g(S.win_mu);
    for (int p = 0; p < NZ_MAX_POOLS; ++p)
        for (int i = 0; i < 2; ++i) if (cw.w[p][i]) { S.free_windows[p].push_back(cw.w[p][i]); cw.w[p][i] = nullptr; }
}

// legacy mode (--io-threads 0): one thread per connection, blocking sockets
void serve_connection(Server* Sp, int fd, std::string client_ip) {
    Server& S = *Sp;
    tune_socket(fd);
    ConnWin cw;
    bool alive = true;
    std::string inbuf;
    std::vector<char> tmp(16384);
    while (alive) {
        size_t end;
        while ((end = inbuf.find("\r\n\r\n")) == std::string::npos) {
            if (inbuf.size() > 32768) { alive = false; break; }
            ssize_t r = recv(fd, tmp.data(), tmp.size(), 0);
            if (r <= 0) { alive = false; break; }
            inbuf.append(tmp.data(), (size_t)r);
        }
        if (!alive) break;
        const std::string head = inbuf.substr(0, end + 2);
        inbuf.erase(0, end + 4);
        if (!serve_one(S, fd, client_ip, head, cw)) break;
    }
    give_windows(S, cw);
    shutdown(fd, SHUT_RDWR);
    close(fd);
    S.conns--;
}


// ---------------------------------------------------------------------------------------------------------
// Event-driven front end (--io-threads N): N workers, each an epoll loop that OWNS its connections.
// Replaces thread-per-connection + dispatcher lanes on the hit path (nginx's own model is an event loop per worker,
// nginx.conf:2,8-10).  The thread that parsed a request is the thread that submits its windows to the GPU (one
// wave per loop turn and shard: every window its connections asked for in that turn = the combining submitter), polls
// their completion and writes them to the socket: no queue hop and no wake-up between "window ready" and send().
// While windows are in flight the loop polls (epoll_wait timeout 0); idle it blocks.  A MISS is the rare, slow case:
// the connection is lent to a helper thread that serves that one request with blocking sockets (serve_one), then
// comes back.  MSG_ZEROCOPY is not used: on loopback the kernel copies at delivery anyway, and the windows are
// reused as soon as send() returns, which zerocopy's completion notifications would only delay.
// ---------------------------------------------------------------------------------------------------------
struct Worker;
struct EvConn {
    int fd = -1;
    Worker* w = nullptr;
    std::string ip, in;
    ConnWin cw;
    nz_buf** win = nullptr;               // the pair the current response streams through (next to its GPU)
    enum WinState { FREE, INFLIGHT, READY, FAILED };
    struct Win { WinState st = FREE; uint64_t len = 0, sent = 0; bool urgent = false; } ws[2];
    // the response being sent
    bool busy = false, keep = true, dead = false, want_out = false, lent = false, queued_more = false;
    Parsed pr;
    nzh::Plan plan;
    std::shared_ptr<Entry> ent;
    size_t part = 0;                      // next plan.body part to schedule
    uint64_t part_off = 0;                // bytes of that slice already scheduled
    uint64_t ramp = 0;
    bool first = true;
    uint32_t flags = 0;
    struct Item { bool window; int slot; std::string bytes; size_t off; };
    std::deque<Item> q;                   // what goes to the socket, in order
    int inflight = 0;                     // windows the GPU may still be writing
    int64_t t_req = 0, last_active = 0;
    uint64_t body_sent = 0;
};

struct WinReq { EvConn* c; int slot; nz_get_req rq; nz_seg sg; };
struct OutTicket { nz_ticket* t; std::vector<std::pair<EvConn*, int>> items; bool urgent; int64_t t_sub; };

struct Worker {
    Server* S = nullptr;
    int ep = -1, evfd = -1;
    std::mutex mu;
    std::vector<EvConn*> inbox;                     // new connections and connections coming back from a miss helper
    std::vector<std::vector<WinReq>> pend_urgent, pend_bulk;     // per shard, this loop turn
    std::vector<OutTicket> out;
    std::vector<EvConn*> conns;
    std::deque<EvConn*> more;                       // connections that stopped sending only to let others have a turn
    std::vector<EvConn*> dying;                     // killed this turn or earlier; deleted at the end of a turn once the GPU is done with their windows
    std::vector<nz_get_req> reqs_scratch;
    std::vector<nz_get_res> res_scratch;
    void post(EvConn* c) {
        { std::lock_guard<std::mutex> g(mu); inbox.push_back(c); }
        uint64_t one = 1;
        if (write(evfd, &one, 8) < 0) {}
    }
    void run();
    void on_readable(EvConn* c);
    void try_start(EvConn* c);
    void start_response(EvConn* c);
    void schedule(EvConn* c);
    void pump(EvConn* c);
    void finish(EvConn* c);
    void kill(EvConn* c);
    void reap(EvConn* c);
    void submit_pending();
    void poll_tickets();
    void lend_to_helper(EvConn* c, const std::string& head);
    void arm(EvConn* c, bool out);
};

void Worker::arm(EvConn* c, bool out) {
    if (c->want_out == out) return;
    c->want_out = out;
    struct epoll_event ev;
    memset(&ev, 0, sizeof ev);
    ev.events = EPOLLIN | EPOLLRDHUP;
    if (out) ev.events |= EPOLLOUT;
    ev.data.ptr = c;
    epoll_ctl(ep, EPOLL_CTL_MOD, c->fd, &ev);
}

void Worker::kill(EvConn* c) {            // stop serving; the object itself goes away at the end of a loop turn (never under a caller's feet),
    if (c->dead) return;                  // and only once the GPU is done with its windows
    c->dead = true;
    epoll_ctl(ep, EPOLL_CTL_DEL, c->fd, nullptr);
    shutdown(c->fd, SHUT_RDWR);
    dying.push_back(c);
}
void Worker::reap(EvConn* c) {
    Server& Sv = *S;
    if (c->ent) { c->ent->inflight--; c->ent.reset(); }
    give_windows(Sv, c->cw);
    close(c->fd);
    Sv.conns--;
    conns.erase(std::remove(conns.begin(), conns.end(), c), conns.end());
    delete c;
}

void Worker::lend_to_helper(EvConn* c, const std::string& head) {
    // miss path: blocking sockets in a helper thread for this one request (origin fetch, fill, uncached relay), then back
    epoll_ctl(ep, EPOLL_CTL_DEL, c->fd, nullptr);
    c->lent = true;
    conns.erase(std::remove(conns.begin(), conns.end(), c), conns.end());
    Server* Sp = S;
    Worker* self = this;
    std::thread([Sp, self, c, head] {
        const int fl = fcntl(c->fd, F_GETFL, 0);
        fcntl(c->fd, F_SETFL, fl & ~O_NONBLOCK);
        const bool open_ = serve_one(*Sp, c->fd, c->ip, head, c->cw);
        if (!open_) {
            give_windows(*Sp, c->cw);
            shutdown(c->fd, SHUT_RDWR);
            close(c->fd);
            Sp->conns--;
            delete c;
            return;
        }
        fcntl(c->fd, F_SETFL, fl | O_NONBLOCK);
        c->lent = false;
        self->post(c);
    }).detach();
}

void Worker::try_start(EvConn* c) {
    Server& Sv = *S;
    while (!c->busy && !c->dead) {
        const size_t end = c->in.find("\r\n\r\n");
        if (end == std::string::npos) {
            if (c->in.size() > 32768) kill(c);
            return;
        }
        const std::string head = c->in.substr(0, end + 2);
        c->in.erase(0, end + 4);
        c->t_req = g_trace ? now_ns() : 0;
        c->pr = parse_head(head);
        Parsed& pr = c->pr;
        c->keep = pr.keep_alive;
        c->ent.reset();
        auto stock = [&](int code) { c->plan = nzh::error_page(code, !pr.error && pr.rq.method == "HEAD"); };
        if (pr.error) {
            stock(pr.error);
            c->keep = false;
        } else if (!nzh::method_allowed(pr.rq.method) || (pr.rq.method != "GET" && pr.rq.method != "HEAD")) {
            stock(405);
        } else {
            int vhost = 0;
            auto key = nzh::cache_key(Sv.cfg, pr.rq.host, pr.rq.target, &vhost);
            if (!key) {
                stock(400);
                c->keep = false;
            } else {
                Key16 k;
                nzh::md5(key->data(), key->size(), k.b);
                {
                    std::lock_guard<std::mutex> g(Sv.mu);
                    auto it = Sv.index.find(k);
                    if (it != Sv.index.end() && (int64_t)time(nullptr) <= it->second->valid_until) {
                        c->ent = it->second;
                        c->ent->inflight++;
                    }
                }
                if (!c->ent) {                    // MISS (or expired): the blocking path owns this request
                    lend_to_helper(c, head);
                    return;
                }
                Sv.hits++;
                c->ent->last_access = now_ms();
                nzh::Config cfg = Sv.cfg;
                cfg.gunzip = Sv.opt.gunzip;
                c->plan = nzh::plan_response(cfg, pr.rq, c->ent->meta, Sv.boundary.fetch_add(1));
            }
        }
        if (!pr.error && pr.has_body) c->keep = false;
        start_response(c);
    }
}

void Worker::start_response(EvConn* c) {
    Server& Sv = *S;
    c->win = nullptr;
    if (c->ent && !c->plan.head) {
        bool slices = false;
        for (const nzh::Part& p : c->plan.body) slices |= !p.literal;
        if (slices) {                                       // windows next to the GPU that holds the object
            c->win = conn_windows(Sv, c->cw, nz_shard_of(c->ent->key128, (uint32_t)Sv.ctxs.size()));
            if (!c->win) { c->plan = nzh::error_page(500, false); c->keep = false; }
        }
    }
    const nzh::Plan& plan = c->plan;
    std::string h = std::string("HTTP/1.1 ") + std::to_string(plan.status) + " " + status_text(plan.status) + "\r\nServer: nginx\r\nDate: " +
                    http_date(time(nullptr)) + "\r\n";
    for (const auto& kv : plan.headers) h += kv.first + ": " + kv.second + "\r\n";
    h += c->keep ? "Connection: keep-alive\r\n" : "Connection: close\r\n";
    h += "\r\n";
    c->busy = true;
    c->part = 0; c->part_off = 0; c->ramp = Sv.opt.first_window; c->first = true; c->body_sent = 0;
    c->flags = (plan.inflated ? NZ_GET_INFLATE : 0) | (Sv.opt.verify ? NZ_GET_VERIFY : 0);
    c->q.clear();
    c->q.push_back(EvConn::Item{false, 0, std::move(h), 0});
    if (plan.head) c->part = plan.body.size();          // HEAD: headers only
    schedule(c);
    pump(c);
}

// walk the plan: literals join the send queue as they come, slices are cut into windows while a window is free
void Worker::schedule(EvConn* c) {
    Server& Sv = *S;
    const nzh::Plan& plan = c->plan;
    while (c->part < plan.body.size()) {
        const nzh::Part& p = plan.body[c->part];
        if (p.literal) {
            if (!c->q.empty() && !c->q.back().window) c->q.back().bytes += p.bytes;
            else c->q.push_back(EvConn::Item{false, 0, p.bytes, 0});
            c->part++;
            continue;
        }
        if (c->part_off >= p.len) { c->part++; c->part_off = 0; continue; }
        int slot = -1;
        for (int i = 0; i < 2; ++i) if (c->ws[i].st == EvConn::FREE) { slot = i; break; }
        if (slot < 0) return;
        const uint64_t left = p.len - c->part_off;
        const uint64_t n = std::min<uint64_t>(left, std::min<uint64_t>(c->ramp, Sv.opt.window_bytes));
        c->ramp = std::min<uint64_t>(c->ramp * 8, Sv.opt.window_bytes);
        WinReq r;
        r.c = c; r.slot = slot;
        memset(&r.rq, 0, sizeof r.rq);
        memcpy(r.rq.key, c->ent->key128, 16);
        r.rq.flags = c->flags | (c->first ? NZ_GET_URGENT : 0u);
        r.sg = nz_seg{p.off + c->part_off, n, 0};
        r.rq.nseg = 1; r.rq.dst = c->win[slot];
        const uint32_t sh = nz_shard_of(c->ent->key128, (uint32_t)Sv.ctxs.size());
        (c->first ? pend_urgent : pend_bulk)[sh].push_back(r);
        c->ws[slot].st = EvConn::INFLIGHT; c->ws[slot].len = n; c->ws[slot].sent = 0; c->ws[slot].urgent = c->first;
        c->inflight++;
        c->q.push_back(EvConn::Item{true, slot, std::string(), 0});
        c->first = false;
        c->part_off += n;
    }
}

void Worker::pump(EvConn* c) {
    Server& Sv = *S;
    size_t budget = 512u << 10;                                         // bytes this call may hand to the socket
    while (!c->dead && !c->q.empty()) {
        EvConn::Item& it = c->q.front();
        const char* p;
        size_t n;
        if (!it.window) {
            // the status line and headers leave at once (like a proxy forwarding upstream headers as they arrive) unless
            // --early-header 0 asked for them to share a segment with the first body bytes
            if (!Sv.opt.early_header && c->q.size() > 1 && c->q[1].window && c->ws[c->q[1].slot].st == EvConn::INFLIGHT && it.off == 0) return;
            p = it.bytes.data() + it.off; n = it.bytes.size() - it.off;
        } else {
            EvConn::Win& w = c->ws[it.slot];
            if (w.st == EvConn::INFLIGHT) return;                       // the GPU is still filling it
            if (w.st == EvConn::FAILED) { kill(c); return; }             // mid-body failure: drop the connection
            p = (const char*)nz_buf_ptr(c->win[it.slot]) + w.sent; n = (size_t)(w.len - w.sent);
        }
        if (budget == 0) {                                              // fairness: a 4 MiB window is ~1 ms of copying; others wait their turn no longer than a slice
            if (!c->queued_more) { c->queued_more = true; more.push_back(c); }
            return;
        }
        if (n > budget) n = budget;
        budget -= n;
        while (n) {
            const ssize_t k = ::send(c->fd, p, n, MSG_NOSIGNAL | MSG_DONTWAIT);
            if (k < 0) {
                if (errno == EINTR) continue;
                if (errno == EAGAIN || errno == EWOULDBLOCK) { arm(c, true); return; }
                kill(c);
                return;
            }
            p += k; n -= (size_t)k;
            if (it.window) { c->ws[it.slot].sent += (uint64_t)k; c->body_sent += (uint64_t)k; Sv.bytes_out += (uint64_t)k; }
            else it.off += (size_t)k;
        }
        if (it.window) {
            EvConn::Win& w = c->ws[it.slot];
            if (w.sent < w.len) continue;                               // slice budget: the rest of this window next iteration of the loop
            if (g_trace && c->t_req && w.urgent) { g_first_ns += (uint64_t)(now_ns() - c->t_req); g_first_n++; }
            w.st = EvConn::FREE;
            c->q.pop_front();
            schedule(c);                                                // the freed window goes straight back to work
        } else {
            if (it.off < it.bytes.size()) continue;
            c->q.pop_front();
        }
    }
    if (c->dead) return;
    arm(c, false);
    if (c->q.empty() && c->part >= c->plan.body.size()) finish(c);
}

void Worker::finish(EvConn* c) {
    Server& Sv = *S;
    access_log_line(Sv, c->ip, c->pr, c->plan.status, c->plan.head ? 0 : c->plan.body_length());
    if (c->ent) { c->ent->inflight--; c->ent.reset(); }
    c->busy = false;
    c->last_active = now_ms();
    if (!c->keep) { kill(c); return; }
    // a pipelined request may already be waiting: try_start's loop (our caller, or the next readable event) picks it up
}

void Worker::on_readable(EvConn* c) {
    char tmp[16384];
    for (int turns = 0; turns < 4; ++turns) {
        const ssize_t r = recv(c->fd, tmp, sizeof tmp, MSG_DONTWAIT);
        if (r > 0) { c->in.append(tmp, (size_t)r); c->last_active = now_ms(); if ((size_t)r < sizeof tmp) break; continue; }
        if (r == 0) { kill(c); return; }
        if (errno == EINTR) continue;
        if (errno == EAGAIN || errno == EWOULDBLOCK) break;
        kill(c);
        return;
    }
    if (!c->busy) try_start(c);
}

void Worker::submit_pending() {
    Server& Sv = *S;
    for (int urgent = 1; urgent >= 0; --urgent) {
        auto& lanes = urgent ? pend_urgent : pend_bulk;
        for (size_t sh = 0; sh < lanes.size(); ++sh) {
            std::vector<WinReq>& v = lanes[sh];
            if (v.empty()) continue;
            reqs_scratch.resize(v.size());
            for (size_t i = 0; i < v.size(); ++i) { reqs_scratch[i] = v[i].rq; reqs_scratch[i].segs = &v[i].sg; }
            OutTicket ot;
            ot.t = nullptr; ot.urgent = urgent != 0; ot.t_sub = now_ns();
            const int32_t rc = nz_get_submit(Sv.ctxs[sh], reqs_scratch.data(), (uint32_t)v.size(), nullptr, &ot.t);
            for (const WinReq& r : v) ot.items.emplace_back(r.c, r.slot);
            if (rc != NZ_OK) {
                fprintf(stderr, "nzproxy: nz_get_submit failed: %s [%s]\n", nz_strerror(rc), nz_last_error());
                for (auto& it : ot.items) { it.first->ws[it.second].st = EvConn::FAILED; it.first->inflight--; }
                for (auto& it : ot.items) if (!it.first->dead) pump(it.first);
            } else {
                if (g_trace) { g_tr[urgent ? 0 : 1].waves++; g_tr[urgent ? 0 : 1].items += v.size(); }
                out.push_back(std::move(ot));
            }
            v.clear();
        }
    }
}

void Worker::poll_tickets() {
    for (size_t i = 0; i < out.size();) {
        OutTicket& ot = out[i];
        const int32_t pr = nz_get_poll(ot.t);
        if (pr == NZ_ETIMEOUT) { ++i; continue; }
        res_scratch.resize(ot.items.size());
        int32_t rc = pr;
        if (rc == NZ_OK) rc = nz_get_wait(ot.t, 0, res_scratch.data());
        nz_ticket_free(ot.t);
        if (g_trace) g_tr[ot.urgent ? 0 : 1].device_ns += (uint64_t)(now_ns() - ot.t_sub) * ot.items.size();
        std::vector<std::pair<EvConn*, int>> items = std::move(ot.items);
        std::vector<nz_get_res> res = res_scratch;
        out.erase(out.begin() + (long)i);
        for (size_t k = 0; k < items.size(); ++k) {
            EvConn* c = items[k].first;
            const bool ok = rc == NZ_OK && res[k].status == NZ_OK;
            if (!ok) fprintf(stderr, "nzproxy: window failed: rc=%d (%s) status=%d [%s]\n", rc, nz_strerror(rc), rc == NZ_OK ? res[k].status : 0, nz_last_error());
            c->ws[items[k].second].st = ok ? EvConn::READY : EvConn::FAILED;
            c->inflight--;
        }
        for (size_t k = 0; k < items.size(); ++k) {
            EvConn* c = items[k].first;
            bool seen = false;
            for (size_t j = 0; j < k; ++j) seen |= items[j].first == c;
            if (seen) continue;
            if (c->dead) continue;
            pump(c);
            if (!c->dead && !c->busy) try_start(c);                      // (c may be lent to a miss helper after this: not touched again)
        }
    }
}

void Worker::run() {
    Server& Sv = *S;
    pend_urgent.resize(Sv.ctxs.size());
    pend_bulk.resize(Sv.ctxs.size());
    struct epoll_event evs[128];
    int64_t last_sweep = now_ms();
    for (;;) {
        // How to wait.  Something to send right now: do not wait.  A first window (urgent) in flight: poll, its completion is
        // tens of microseconds away and is the latency a client sees.  Only bulk windows in flight: sleep in short slices
        // (a spinning worker per loop starves the very sockets it feeds on a box with fewer cores than threads).
        // Nothing in flight: block.
        const int64_t tnow = out.empty() ? 0 : now_ns();
        bool urgent_out = false;
        for (const OutTicket& ot : out) urgent_out |= ot.urgent && tnow - ot.t_sub < 300000;
        int n;
        if (!more.empty() || urgent_out) n = epoll_wait(ep, evs, 128, 0);
        else if (!out.empty()) {
            struct timespec ts = {0, 30000};
            n = (int)syscall(SYS_epoll_pwait2, ep, evs, 128, &ts, nullptr, 0);
            if (n < 0 && errno == ENOSYS) n = epoll_wait(ep, evs, 128, 0);
        } else n = epoll_wait(ep, evs, 128, 1000);
        const bool hot = urgent_out;
        for (int i = 0; i < n; ++i) {
            if (evs[i].data.ptr == nullptr) {                            // the eventfd: connections to adopt
                uint64_t cnt;
                if (read(evfd, &cnt, 8) < 0) {}
                std::vector<EvConn*> take;
                { std::lock_guard<std::mutex> g(mu); take.swap(inbox); }
                for (EvConn* c : take) {
                    c->w = this;
                    c->want_out = false;
                    c->last_active = now_ms();
                    struct epoll_event ev;
                    memset(&ev, 0, sizeof ev);
                    ev.events = EPOLLIN | EPOLLRDHUP;
                    ev.data.ptr = c;
                    conns.push_back(c);
                    if (epoll_ctl(ep, EPOLL_CTL_ADD, c->fd, &ev) != 0) { c->dead = true; shutdown(c->fd, SHUT_RDWR); dying.push_back(c); continue; }
                    if (!c->in.empty()) try_start(c);                    // back from a miss helper with a pipelined request
                }
                continue;
            }
            EvConn* c = static_cast<EvConn*>(evs[i].data.ptr);
            if (std::find(conns.begin(), conns.end(), c) == conns.end() || c->dead) continue;     // lent to a miss helper earlier in this batch / killed
            if (evs[i].events & (EPOLLERR | EPOLLHUP)) { kill(c); continue; }
            if (evs[i].events & EPOLLOUT) {
                pump(c);
                if (!c->dead && !c->busy) try_start(c);
                continue;                                                // (input, if any, is reported again: level triggered)
            }
            if (evs[i].events & (EPOLLIN | EPOLLRDHUP)) on_readable(c);
        }
        submit_pending();
        poll_tickets();
        for (size_t k = more.size(); k > 0 && !more.empty(); --k) {     // one more slice for everyone who was cut short
            EvConn* c = more.front();
            more.pop_front();
            c->queued_more = false;
            if (c->dead) continue;
            pump(c);
            if (!c->dead && !c->busy) try_start(c);
        }
        submit_pending();                                                // windows freed by this turn's sends
        if (hot && n == 0) __builtin_ia32_pause();
        const int64_t now = now_ms();
        if (now - last_sweep > 1000) {                                   // client_header_timeout 60 s / keepalive_timeout 65 s (nginx defaults)
            last_sweep = now;
            std::vector<EvConn*> idle;
            for (EvConn* c : conns)
                if (!c->dead && !c->busy && now - c->last_active > (c->in.empty() ? 65000 : 60000)) idle.push_back(c);
            for (EvConn* c : idle) kill(c);
        }
        for (size_t i = 0; i < dying.size();) {
            if (dying[i]->inflight == 0 && !dying[i]->queued_more) { reap(dying[i]); dying.erase(dying.begin() + (long)i); }
            else ++i;
        }
    }
}

void usage() {
    fprintf(stderr,
            "nzproxy [--listen ip:port] [--device N | --devices 0,1,..] [--arena-gib G] [--max-objects N] [--window-mib M] [--first-window-kib K] [--host-rate-gbs R] [--host-ctas N]\n"
            "        [--max-conns N] [--io-threads N] [--gunzip] [--s3-private bucket:access:secret] [--verify] [--upstream proxy_host=[https://]ip:port]... [--proxy-ssl-server-name] [--access-log PATH] [--pv DIR] [--inactive-s S] [--valid-s S]\n");
}

}  // namespace

int main(int argc, char** argv) {
    signal(SIGPIPE, SIG_IGN);
    Server S;
    // both `--opt value` and `--opt=value` (the form deploy/nezha-b200.yaml uses)
    std::vector<std::string> argx;
    for (int i = 1; i < argc; ++i) {
        const std::string a = argv[i];
        const size_t eq = a.rfind("--", 0) == 0 ? a.find('=') : std::string::npos;
        if (eq != std::string::npos && a != "--upstream") { argx.push_back(a.substr(0, eq)); argx.push_back(a.substr(eq + 1)); }
        else argx.push_back(a);
    }
    std::vector<const char*> argp;
    argp.push_back(argv[0]);
    for (const std::string& a : argx) argp.push_back(a.c_str());
    argc = (int)argp.size();
    argv = const_cast<char**>(argp.data());
    for (int i = 1; i < argc; ++i) {
        const std::string a = argv[i];
        auto next = [&]() -> const char* { if (i + 1 >= argc) { usage(); exit(2); } return argv[++i]; };
        if (a == "--listen") {
            const std::string v = next();
            const size_t c = v.rfind(':');
            if (c == std::string::npos) S.opt.port = atoi(v.c_str());
            else { S.opt.listen_ip = v.substr(0, c); S.opt.port = atoi(v.c_str() + c + 1); }
        } else if (a == "--device") S.opt.device = atoi(next());
        else if (a == "--devices") {
            std::string v = next();
            size_t q = 0;
            while (q <= v.size()) {
                size_t c2 = v.find(',', q);
                if (c2 == std::string::npos) c2 = v.size();
                if (c2 > q) S.opt.devices.push_back(atoi(v.substr(q, c2 - q).c_str()));
                q = c2 + 1;
            }
        }
        else if (a == "--arena-gib") S.opt.arena_bytes = (uint64_t)(atof(next()) * (1ull << 30));
        else if (a == "--max-objects") S.opt.max_objects = (uint32_t)atoi(next());
        else if (a == "--window-mib") S.opt.window_bytes = (uint64_t)(atof(next()) * (1ull << 20));
        else if (a == "--first-window-kib") S.opt.first_window = (uint64_t)atoll(next()) << 10;
        else if (a == "--host-rate-gbs") S.opt.host_rate_gbs = atof(next());
        else if (a == "--host-ctas") S.opt.host_ctas = atoi(next());
        else if (a == "--early-header") S.opt.early_header = atoi(next()) != 0;
        else if (a == "--proxy-ssl-server-name") S.opt.proxy_ssl_server_name = true;
        else if (a == "--access-log") S.opt.access_log = next();
        else if (a == "--max-conns") S.opt.max_conns = atoi(next());
        else if (a == "--io-threads") S.opt.io_threads = atoi(next());
        else if (a == "--gunzip") S.opt.gunzip = true;
        else if (a == "--s3-private") {                // bucket:access:secret
            const std::string v = next();
            const size_t c1 = v.find(':'), c2 = v.find(':', c1 == std::string::npos ? 0 : c1 + 1);
            if (c1 == std::string::npos || c2 == std::string::npos) { usage(); return 2; }
            S.opt.s3_bucket = v.substr(0, c1); S.opt.s3_access = v.substr(c1 + 1, c2 - c1 - 1); S.opt.s3_secret = v.substr(c2 + 1);
            S.cfg.vhosts.push_back(nzh::VHost{"s3.amazonaws.com", "s3.amazonaws.com", false});      // nginx.conf:141,164
        }
        else if (a == "--pv") S.opt.pv_dir = next();
        else if (a == "--inactive-s") S.opt.inactive_s = atoi(next());
        else if (a == "--valid-s") S.opt.valid_s = atoi(next());
        else if (a == "--verify") S.opt.verify = true;
        else if (a == "--upstream") {
            const std::string v = next();
            const size_t e = v.find('=');
            if (e == std::string::npos) { usage(); return 2; }
            S.opt.upstream[v.substr(0, e)] = v.substr(e + 1);
        } else { usage(); return 2; }
    }
    if (S.opt.first_window > S.opt.window_bytes) S.opt.first_window = S.opt.window_bytes;
    if (S.opt.devices.empty()) S.opt.devices.push_back(S.opt.device);
    for (size_t i = 0; i < S.opt.devices.size(); ++i) {
        nz_config cfg;
        memset(&cfg, 0, sizeof cfg);
        cfg.struct_size = sizeof cfg;
        cfg.device = S.opt.devices[i];
        cfg.arena_bytes = S.opt.arena_bytes;
        cfg.max_objects = S.opt.max_objects;
        cfg.shard_index = (uint32_t)i;
        cfg.shard_count = (uint32_t)S.opt.devices.size();
        nz_ctx* cx = nullptr;
        int32_t rc = nz_init(&cfg, &cx);
        if (rc != NZ_OK) {
            fprintf(stderr, "nzproxy: nz_init(device %d) failed: %s (%s) - there is no CPU fallback\n", cfg.device, nz_strerror(rc), nz_last_error());
            return 1;
        }
        if (!getenv("NZ_HOST_RATE_GBS") && !getenv("NZ_HOST_GRID")) nz_set_host_rate(cx, S.opt.host_rate_gbs, S.opt.host_ctas);
        S.ctxs.push_back(cx);
    }
    S.ctx = S.ctxs[0];
    {   // one window pool per host memory node that has a GPU of ours
        int node_of_pool[NZ_MAX_POOLS];
        S.n_pools = 0;
        for (nz_ctx* cx : S.ctxs) {
            const int node = nz_numa_node(cx);
            int pool = -1;
            for (int k = 0; k < S.n_pools; ++k) if (node_of_pool[k] == node) pool = k;
            if (pool < 0 && S.n_pools < NZ_MAX_POOLS) { pool = S.n_pools++; node_of_pool[pool] = node; S.pool_ctx[pool] = cx; }
            S.pool_of_shard.push_back(pool < 0 ? 0 : pool);
        }
        const int per_pool = std::min(S.opt.max_conns, 160) * 2 / S.n_pools;
        for (int k = 0; k < S.n_pools; ++k)
            for (int i = 0; i < per_pool; ++i) {                     // pinned windows up front: cudaHostAlloc is slow
                nz_buf* b = nullptr;
                if (nz_buf_alloc(S.pool_ctx[k], S.opt.window_bytes, NZ_BUF_PINNED_HOST, &b) != NZ_OK) break;
                S.free_windows[k].push_back(b);
            }
    }
    if (!S.opt.pv_dir.empty()) pv_reload(S);
    std::thread([&S] { for (;;) { std::this_thread::sleep_for(std::chrono::seconds(S.opt.inactive_s < 120 ? 1 : 60)); cache_manager_pass(S); } }).detach();
    {   // CPUs this process may really use: affinity mask and the cgroup's CPU quota (a container that shows 128 CPUs may own 16)
        cpu_set_t set;
        int cpus = (int)std::thread::hardware_concurrency();
        if (sched_getaffinity(0, sizeof set, &set) == 0) cpus = std::min(cpus, CPU_COUNT(&set));
        if (FILE* f = fopen("/sys/fs/cgroup/cpu.max", "r")) {
            long long quota = 0, period = 0;
            char q[32];
            if (fscanf(f, "%31s %lld", q, &period) == 2 && strcmp(q, "max") != 0 && (quota = atoll(q)) > 0 && period > 0)
                cpus = std::min<int>(cpus, (int)std::max<long long>(1, quota / period));
            fclose(f);
        }
        g_spin_cap = std::max(1, cpus / 8);
        if (const char* e = getenv("NZ_PROXY_SPIN_CAP")) g_spin_cap = atoi(e);
        fprintf(stderr, "nzproxy: %d usable CPUs, at most %d threads poll for a window at a time\n", cpus, g_spin_cap);
    }
    for (nz_ctx* cx : S.ctxs) {
        auto u = std::make_unique<Combiner>();
        u->ctx = cx; u->cap_bytes = 16ull << 20; u->linger_us = 0; u->spin_us = g_spin_cap >= 4 ? 2000 : 100; u->lane = 0;
        auto k = std::make_unique<Combiner>();
        k->ctx = cx; k->cap_bytes = 16ull << 20; k->linger_us = 20;
        if (const char* e = getenv("NZ_PROXY_BULK_CAP_MIB")) k->cap_bytes = (uint64_t)std::max(1, atoi(e)) << 20;
        Combiner* up = u.get();
        Combiner* kp = k.get();
        g_urgent.push_back(std::move(u));
        g_bulk.push_back(std::move(k));
        int n_urgent = 2;
        if (const char* e = getenv("NZ_PROXY_URGENT_THREADS")) n_urgent = std::max(1, atoi(e));
        for (int i = 0; i < n_urgent; ++i) std::thread([up] { up->run(); }).detach();
        int n_bulk = 2;
        if (const char* e = getenv("NZ_PROXY_BULK_THREADS")) n_bulk = std::max(1, atoi(e));
        for (int i = 0; i < n_bulk; ++i) std::thread([kp] { kp->run(); }).detach();
    }
    if (!S.opt.access_log.empty()) {
        S.log_fd = open(S.opt.access_log.c_str(), O_WRONLY | O_CREAT | O_APPEND, 0644);
        if (S.log_fd < 0) fprintf(stderr, "nzproxy: cannot open access log %s: %s\n", S.opt.access_log.c_str(), strerror(errno));
    }
    if (const char* e = getenv("NZ_PROXY_TRACE")) g_trace = atoi(e) != 0;
    if (const char* e = getenv("NZ_PROXY_DIRECT")) g_direct_max = atoi(e);
    prctl(PR_SET_TIMERSLACK, 1000UL, 0, 0, 0);      // 1 us: the library's sleep-polling of long waves stays precise (inherited by threads)
    if (g_trace)
        std::thread([] {
            uint64_t last = 0;
            for (;;) {
                std::this_thread::sleep_for(std::chrono::seconds(1));
                const uint64_t tot = g_tr[0].items + g_tr[1].items;
                if (tot == last) continue;
                last = tot;
                for (int l = 0; l < 2; ++l) {
                    LaneTrace& T = g_tr[l];
                    const double n = (double)std::max<uint64_t>(1, T.items.load());
                    fprintf(stderr, "nzproxy trace %s: %llu windows in %llu waves (%.1f MiB/wave)  mean us: queue %.1f  submit %.1f  submit->done %.1f (on the GPU %.1f)  wake %.1f\n",
                            l ? "bulk  " : "urgent", (unsigned long long)T.items.load(), (unsigned long long)T.waves.load(),
                            T.bytes.load() / 1048576.0 / (double)std::max<uint64_t>(1, T.waves.load()), T.queue_ns / n / 1e3, T.submit_ns / n / 1e3,
                            T.device_ns / n / 1e3, T.gpu_ns / n / 1e3, T.wake_ns / n / 1e3);
                }
                fprintf(stderr, "nzproxy trace: request parsed -> first window on the socket: mean %.1f us over %llu responses; %llu first windows submitted directly (mean %.1f us)\n",
                        g_first_ns / (double)std::max<uint64_t>(1, g_first_n.load()) / 1e3, (unsigned long long)g_first_n.load(),
                        (unsigned long long)g_direct_n.load(), g_direct_ns / (double)std::max<uint64_t>(1, g_direct_n.load()) / 1e3);
            }
        }).detach();
    std::vector<std::unique_ptr<Worker>> workers;
    size_t rr = 0;
    {
        int nio = S.opt.io_threads;
        if (const char* e = getenv("NZ_PROXY_IO_THREADS")) nio = atoi(e);
        if (nio < 0) nio = 0;
        for (int i = 0; i < nio; ++i) {
            auto w = std::make_unique<Worker>();
            w->S = &S;
            w->ep = epoll_create1(EPOLL_CLOEXEC);
            w->evfd = eventfd(0, EFD_NONBLOCK | EFD_CLOEXEC);
            struct epoll_event ev;
            memset(&ev, 0, sizeof ev);
            ev.events = EPOLLIN;
            ev.data.ptr = nullptr;
            if (w->ep < 0 || w->evfd < 0 || epoll_ctl(w->ep, EPOLL_CTL_ADD, w->evfd, &ev) != 0) { perror("nzproxy: epoll"); return 1; }
            Worker* wp = w.get();
            std::thread([wp] { wp->run(); }).detach();
            workers.push_back(std::move(w));
        }
    }
    int ls = socket(AF_INET, SOCK_STREAM, 0);
    int one = 1;
    setsockopt(ls, SOL_SOCKET, SO_REUSEADDR, &one, sizeof one);
    struct sockaddr_in sa;
    memset(&sa, 0, sizeof sa);
    sa.sin_family = AF_INET;
    sa.sin_port = htons((uint16_t)S.opt.port);
    inet_pton(AF_INET, S.opt.listen_ip.c_str(), &sa.sin_addr);
    if (bind(ls, (struct sockaddr*)&sa, sizeof sa) != 0 || listen(ls, 1024) != 0) {
        perror("nzproxy: bind/listen");
        return 1;
    }
    fprintf(stderr, "nzproxy: listening on %s:%d, %zu GPU shard(s), arena %.1f GiB each, %zu pinned windows of %.1f MiB in %d node pool(s), %zu event-loop worker(s)%s\n", S.opt.listen_ip.c_str(),
            S.opt.port, S.ctxs.size(), S.opt.arena_bytes / 1073741824.0, [&] { size_t n = 0; for (int k = 0; k < S.n_pools; ++k) n += S.free_windows[k].size(); return n; }(), S.opt.window_bytes / 1048576.0, S.n_pools, workers.size(),
            workers.empty() ? " (thread per connection)" : "");
    fflush(stderr);
    for (;;) {
        struct sockaddr_in ca;
        socklen_t cl = sizeof ca;
        int fd = accept(ls, (struct sockaddr*)&ca, &cl);
        if (fd < 0) {
            if (errno == EINTR) continue;
            break;
        }
        if (S.conns.load() >= S.opt.max_conns) { close(fd); continue; }   // worker_connections analogue
        S.conns++;
        char ip[64];
        inet_ntop(AF_INET, &ca.sin_addr, ip, sizeof ip);
        if (workers.empty()) { std::thread(serve_connection, &S, fd, std::string(ip)).detach(); continue; }
        EvConn* c = new EvConn();
        c->fd = fd;
        c->ip = ip;
        tune_socket(fd);
        fcntl(fd, F_SETFL, fcntl(fd, F_GETFL, 0) | O_NONBLOCK);
        workers[rr++ % workers.size()]->post(c);
    }
    for (nz_ctx* cx : S.ctxs) nz_shutdown(cx);
    return 0;
}
