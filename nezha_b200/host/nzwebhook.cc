// nezha_b200/host/nzwebhook.cc — HTTPS :443 server around webhook.cc.
// Flags mirror /root/reference/app/webhook/webhook.go:41-48 (single-dash Go style):
//   -config-file PATH -tls-cert-file PATH -tls-private-key-file PATH   [-listen ip:port] [-plain]
// The config file is re-read every 10 s (webhook.go:221-234); a bad file keeps the old config.
#include <arpa/inet.h>
#include <netinet/in.h>
#include <openssl/err.h>
#include <openssl/ssl.h>
#include <signal.h>
#include <sys/socket.h>
#include <unistd.h>

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <memory>
#include <mutex>
#include <sstream>
#include <string>
#include <thread>

#include "webhook.h"

static std::mutex g_mu;
static std::shared_ptr<std::vector<nzw::Config>> g_cfg;

static bool load(const std::string& path) {
    std::ifstream f(path);
    if (!f) return false;
    std::stringstream ss;
    ss << f.rdbuf();
    try {
        auto c = std::make_shared<std::vector<nzw::Config>>(nzw::parse_config(ss.str()));
        std::lock_guard<std::mutex> g(g_mu);
        g_cfg = c;
        return true;
    } catch (const std::exception& e) {
        fprintf(stderr, "nzwebhook: invalid config: %s\n", e.what());
        return false;
    }
}

struct Io {
    int fd;
    SSL* ssl;
    int rd(char* b, int n) { return ssl ? SSL_read(ssl, b, n) : (int)recv(fd, b, (size_t)n, 0); }
    int wr(const char* b, int n) { return ssl ? SSL_write(ssl, b, n) : (int)send(fd, b, (size_t)n, MSG_NOSIGNAL); }
};

static void handle(Io io) {
    std::string in;
    char buf[8192];
    size_t he;
    while ((he = in.find("\r\n\r\n")) == std::string::npos) {
        const int r = io.rd(buf, sizeof buf);
        if (r <= 0 || in.size() > (1 << 20)) return;
        in.append(buf, (size_t)r);
    }
    std::string head = in.substr(0, he + 2), lower = head;
    for (char& c : lower) c = (char)tolower((unsigned char)c);
    char method[16] = "", path[256] = "";
    sscanf(head.c_str(), "%15s %255s", method, path);
    size_t clen = 0;
    if (size_t p = lower.find("\r\ncontent-length:"); p != std::string::npos) clen = (size_t)atoll(lower.c_str() + p + 17);
    std::string ctype;
    bool has_ct = false;
    if (size_t p = lower.find("\r\ncontent-type:"); p != std::string::npos) {
        size_t v = p + 15, e = head.find("\r\n", v);
        while (v < e && head[v] == ' ') ++v;
        ctype = head.substr(v, e - v);
        has_ct = true;
    }
    std::string body = in.substr(he + 4);
    while (body.size() < clen) {
        const int r = io.rd(buf, sizeof buf);
        if (r <= 0) return;
        body.append(buf, (size_t)r);
    }
    body.resize(clen);
    std::string out;
    int code = 200;
    const std::string pth = path;
    if (pth == "/mutate-job" || pth == "/mutate-deployment") {                 // webhook.go:236-237
        std::shared_ptr<std::vector<nzw::Config>> cfg;
        { std::lock_guard<std::mutex> g(g_mu); cfg = g_cfg; }
        out = nzw::serve(*cfg, pth, has_ct ? ctype.c_str() : nullptr, body);
    } else {
        code = 404;
        out = "404 page not found\n";                                          // net/http default mux
    }
    char hdr[256];
    const int n = snprintf(hdr, sizeof hdr, "HTTP/1.1 %d %s\r\nContent-Length: %zu\r\n%sConnection: close\r\n\r\n", code, code == 200 ? "OK" : "Not Found",
                           out.size(), code == 200 ? "" : "Content-Type: text/plain; charset=utf-8\r\n");
    io.wr(hdr, n);
    if (!out.empty()) io.wr(out.data(), (int)out.size());
}

int main(int argc, char** argv) {
    signal(SIGPIPE, SIG_IGN);
    std::string cfgfile, cert, key, ip = "0.0.0.0";
    int port = 443;
    bool plain = false;
    for (int i = 1; i < argc; ++i) {
        std::string a = argv[i];
        while (a.size() > 1 && a[0] == '-' && a[1] == '-') a.erase(0, 1);
        auto val = [&](const char* name) -> const char* {
            const std::string n = std::string("-") + name;
            if (a == n && i + 1 < argc) return argv[++i];
            if (a.compare(0, n.size() + 1, n + "=") == 0) return argv[i] + (strlen(argv[i]) - (a.size() - n.size() - 1));
            return nullptr;
        };
        if (const char* v = val("config-file")) cfgfile = v;
        else if (const char* v = val("tls-cert-file")) cert = v;
        else if (const char* v = val("tls-private-key-file")) key = v;
        else if (const char* v = val("listen")) { std::string s = v; size_t c = s.rfind(':'); if (c == std::string::npos) port = atoi(v); else { ip = s.substr(0, c); port = atoi(s.c_str() + c + 1); } }
        else if (a == "-plain") plain = true;
        else if (a == "-v" || a.compare(0, 3, "-v=") == 0 || a == "-logtostderr" || a.compare(0, 13, "-logtostderr=") == 0) { if (a == "-v" && i + 1 < argc) ++i; }
        else { fprintf(stderr, "nzwebhook: unknown flag %s\n", argv[i]); return 2; }
    }
    if (cfgfile.empty()) { fprintf(stderr, "hostAliases config file is empty\n"); return 1; }          // webhook.go:213-215
    if (!load(cfgfile)) { fprintf(stderr, "failed to parse config file: %s\n", cfgfile.c_str()); return 1; }
    std::thread([cfgfile] { for (;;) { std::this_thread::sleep_for(std::chrono::seconds(10)); load(cfgfile); } }).detach();
    SSL_CTX* ctx = nullptr;
    if (!plain) {
        ctx = SSL_CTX_new(TLS_server_method());
        if (!ctx || SSL_CTX_use_certificate_chain_file(ctx, cert.c_str()) != 1 || SSL_CTX_use_PrivateKey_file(ctx, key.c_str(), SSL_FILETYPE_PEM) != 1) {
            ERR_print_errors_fp(stderr);
            return 1;
        }
    }
    int ls = socket(AF_INET, SOCK_STREAM, 0), one = 1;
    setsockopt(ls, SOL_SOCKET, SO_REUSEADDR, &one, sizeof one);
    sockaddr_in sa;
    memset(&sa, 0, sizeof sa);
    sa.sin_family = AF_INET; sa.sin_port = htons((uint16_t)port);
    inet_pton(AF_INET, ip.c_str(), &sa.sin_addr);
    if (bind(ls, (sockaddr*)&sa, sizeof sa) || listen(ls, 128)) { perror("nzwebhook: bind/listen"); return 1; }
    fprintf(stderr, "starting server\n");
    for (;;) {
        int fd = accept(ls, nullptr, nullptr);
        if (fd < 0) continue;
        std::thread([fd, ctx] {
            Io io{fd, nullptr};
            if (ctx) {
                io.ssl = SSL_new(ctx);
                SSL_set_fd(io.ssl, fd);
                if (SSL_accept(io.ssl) != 1) { SSL_free(io.ssl); close(fd); return; }
            }
            handle(io);
            if (io.ssl) { SSL_shutdown(io.ssl); SSL_free(io.ssl); }
            close(fd);
        }).detach();
    }
}
