/*
 * oracle/nz_cpu_server.c — the "nginx-restated" CPU server: the reference's cache-hit
 * path as a process, for HTTP-level parity replay and as the CPU arm of the HTTP bench.
 *
 * TEST INFRASTRUCTURE ONLY (like everything under oracle/).  PARITY UNPINNED: it restates
 * what /root/reference/examples/demo/nginx.conf makes stock nginx do, it is not nginx.
 *
 *   worker model   worker_processes 1 (nginx.conf:2) -> one single-threaded epoll loop;
 *                  --workers N forks N such loops on SO_REUSEPORT sockets (the "all cores" line)
 *   cache          objects are files in a directory named by the 32-hex md5 of the key
 *                  (proxy_cache_path /var/cache/nginx, no levels=, nginx.conf:38) plus a manifest
 *                  holding the stored upstream headers; files sit in the page cache
 *   request        key = "$host/$proxy_host$uri" (nginx.conf:56), method gate (:47-49),
 *                  not-modified + range filters = nzo_plan (oracle/SPEC.md C, R)
 *   body           sendfile off, output_buffers 2 32k (nginx defaults, nothing set in the conf):
 *                  pread 32 KiB into an output buffer, send it, repeat (SPEC B1)
 * A miss answers 504 (the fill path is not part of the hit-path baseline).
 * --gunzip: the operator switch of SPEC G (new behaviour): stored `Content-Encoding: gzip` objects are inflated with
 * zlib when they are loaded; requests without `Accept-Encoding: gzip` are served the inflated representation (a member
 * zlib declines stays as stored, SPEC G4).
 *
 * manifest line:  <md5hex> <stored_len> <raw_len|-1> <content-type|-> <etag|-> <encoding|->[,noranges] <last-modified...|->
 *                 (",noranges": the upstream response had no Accept-Ranges header, SPEC R0)
 */
#define _GNU_SOURCE
#include <arpa/inet.h>
#include <ctype.h>
#include <errno.h>
#include <fcntl.h>
#include <netinet/in.h>
#include <netinet/tcp.h>
#include <signal.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <strings.h>
#include <sys/epoll.h>
#include <sys/prctl.h>
#include <sys/socket.h>
#include <sys/wait.h>
#include <time.h>
#include <unistd.h>

#include "nz_oracle.h"

#define OUTBUF 32768
#define MAXSEG 140

typedef struct obj {
    uint8_t key[16];
    int fd;
    int64_t stored_len, raw_len;
    char ctype[128], etag[128], lm[64], enc[32];
    int has_etag, has_lm, has_enc, no_ranges;
    uint8_t* raw;                    /* --gunzip: inflated representation (NULL = none) */
} obj;

static obj* g_objs;
static int g_nobj;
static int g_gunzip;
static uint64_t g_boundary = 1;

typedef struct seg { int lit; char* bytes; int64_t off, len; } seg;

typedef struct conn {
    int fd;
    char in[16384];
    int in_len;
    /* response being sent */
    char* head; int head_len, head_off;
    seg segs[MAXSEG]; int nseg, cur;
    int64_t seg_done;
    const obj* o;
    uint8_t out[2][OUTBUF];          /* output_buffers 2 32k */
    int out_len, out_off, out_which;
    int keep, active;
    int inflated;                    /* slices address the inflated representation */
} conn;

static int hexv(int c) { return c <= '9' ? c - '0' : (c | 32) - 'a' + 10; }

static const obj* find_obj(const uint8_t k[16]) {
    for (int i = 0; i < g_nobj; i++)
        if (memcmp(g_objs[i].key, k, 16) == 0) return &g_objs[i];
    return NULL;
}

static int load_manifest(const char* dir) {
    char path[1024];
    snprintf(path, sizeof path, "%s/manifest", dir);
    FILE* f = fopen(path, "r");
    if (!f) return -1;
    char line[1024];
    int cap = 0;
    while (fgets(line, sizeof line, f)) {
        char hex[64], ct[128], et[128], en[32], lm[128];
        long long sl, rl;
        lm[0] = 0;
        int n = sscanf(line, "%63s %lld %lld %127s %127s %31s %127[^\n]", hex, &sl, &rl, ct, et, en, lm);
        if (n < 6 || strlen(hex) != 32) continue;
        if (g_nobj == cap) { cap = cap ? cap * 2 : 64; g_objs = realloc(g_objs, (size_t)cap * sizeof(obj)); }
        obj* o = &g_objs[g_nobj];
        memset(o, 0, sizeof *o);
        for (int i = 0; i < 16; i++) o->key[i] = (uint8_t)(hexv(hex[2 * i]) * 16 + hexv(hex[2 * i + 1]));
        o->stored_len = sl; o->raw_len = rl;
        if (strcmp(ct, "-")) snprintf(o->ctype, sizeof o->ctype, "%s", ct);
        if (strcmp(et, "-")) { snprintf(o->etag, sizeof o->etag, "%s", et); o->has_etag = 1; }
        { char* nr = strstr(en, ",noranges"); if (nr) { *nr = 0; o->no_ranges = 1; } }
        if (strcmp(en, "-")) { snprintf(o->enc, sizeof o->enc, "%s", en); o->has_enc = 1; }
        if (n == 7 && strcmp(lm, "-")) { snprintf(o->lm, sizeof o->lm, "%s", lm); o->has_lm = 1; }
        snprintf(path, sizeof path, "%s/%s", dir, hex);
        o->fd = open(path, O_RDONLY);
        if (o->fd < 0) continue;
        o->raw = NULL;
        if (o->has_enc && !strcasecmp(o->enc, "gzip")) {
            o->raw_len = -1;                                   /* known only if the member inflates (SPEC G3/G4) */
            if (g_gunzip && sl >= 18) {
                uint8_t* st = malloc((size_t)sl);
                if (st && pread(o->fd, st, (size_t)sl, 0) == (ssize_t)sl) {
                    int64_t isz = nzo_gzip_isize(st, (size_t)sl);
                    uint8_t* rw = isz >= 0 ? malloc((size_t)isz + 1) : NULL;
                    if (rw) {
                        int64_t got = nzo_gunzip(st, (size_t)sl, rw, (size_t)isz + 1);
                        if (got == isz) { o->raw = rw; o->raw_len = got; } else free(rw);
                    }
                }
                free(st);
            }
        }
        g_nobj++;
    }
    fclose(f);
    return g_nobj;
}

static const char* status_text(int c) {
    switch (c) {
    case 200: return "OK"; case 206: return "Partial Content"; case 304: return "Not Modified";
    case 400: return "Bad Request"; case 405: return "Not Allowed"; case 412: return "Precondition Failed";
    case 416: return "Requested Range Not Satisfiable"; case 504: return "Gateway Time-out"; case 505: return "HTTP Version Not Supported";
    default: return "Error";
    }
}

static void free_resp(conn* c) {
    free(c->head); c->head = NULL;
    for (int i = 0; i < c->nseg; i++) free(c->segs[i].bytes);
    c->nseg = 0; c->cur = 0; c->seg_done = 0; c->out_len = c->out_off = 0; c->active = 0;
}

static void simple_page(conn* c, int code, int keep) {
    char page[512];
    int n = snprintf(page, sizeof page, "<html>\r\n<head><title>%d %s</title></head>\r\n<body>\r\n<center><h1>%d %s</h1></center>\r\n"
                     "<hr><center>nginx</center>\r\n</body>\r\n</html>\r\n", code, status_text(code), code, status_text(code));
    c->head = malloc(1024);
    c->head_len = snprintf(c->head, 1024, "HTTP/1.1 %d %s\r\nServer: nginx\r\nContent-Type: text/html\r\nContent-Length: %d\r\nConnection: %s\r\n\r\n%s",
                           code, status_text(code), n, keep ? "keep-alive" : "close", page);
    c->head_off = 0; c->nseg = 0; c->keep = keep; c->active = 1;
}

/* build the response for one parsed request head (NUL-terminated, ends with CRLF) */
static void build_response(conn* c, char* head) {
    /* SPEC P: the head goes through the oracle's own parser (nzo_parse_head); its text form is read back here */
    static __thread char ptxt[40960];
    char method[64] = "", target[8192] = "";
    int minor = 1, keep = 1, body = 0, perr = 400;
    char *host = NULL, *a = NULL, *b = NULL, *d = NULL, *e = NULL, *f = NULL, *g = NULL, *h = NULL;
    if (nzo_parse_head(head, strlen(head), ptxt, sizeof ptxt) >= 0) {
        char* save0 = NULL;
        for (char* ln = strtok_r(ptxt, "\n", &save0); ln; ln = strtok_r(NULL, "\n", &save0)) {
            if (!strncmp(ln, "error ", 6)) perr = atoi(ln + 6);
            else if (!strncmp(ln, "method ", 7)) snprintf(method, sizeof method, "%s", ln + 7);
            else if (!strncmp(ln, "target ", 7)) snprintf(target, sizeof target, "%s", ln + 7);
            else if (!strncmp(ln, "minor ", 6)) minor = atoi(ln + 6);
            else if (!strncmp(ln, "keep ", 5)) keep = atoi(ln + 5);
            else if (!strncmp(ln, "body ", 5)) body = atoi(ln + 5);
            else if (!strncmp(ln, "f ", 2)) {
                char* colon = strstr(ln + 2, ": ");
                if (!colon) continue;
                *colon = 0;
                char* v = strdup(colon + 2);
                const char* n = ln + 2;
                if (!strcmp(n, "host")) host = v;
                else if (!strcmp(n, "range")) a = v;
                else if (!strcmp(n, "if-range")) b = v;
                else if (!strcmp(n, "if-none-match")) d = v;
                else if (!strcmp(n, "if-match")) e = v;
                else if (!strcmp(n, "if-modified-since")) f = v;
                else if (!strcmp(n, "if-unmodified-since")) g = v;
                else if (!strcmp(n, "accept-encoding")) h = v;
                else free(v);
            }
        }
    }
    char* plan = NULL;
    if (perr) { simple_page(c, perr, 0); goto done; }
    if (body) keep = 0;                       /* P10: a body is never read: answer, then close */
    nzo_request rq;
    memset(&rq, 0, sizeof rq);
    rq.method = method; rq.http_minor = minor;
    rq.range = a; rq.if_range = b; rq.if_none_match = d; rq.if_match = e; rq.if_modified_since = f; rq.if_unmodified_since = g; rq.accept_encoding = h;
    const obj* o = NULL;
    if (!nzo_method_gate(method) || (strcmp(method, "GET") && strcmp(method, "HEAD"))) {
        simple_page(c, 405, keep);
        goto done;
    }
    {
        char hl[512] = "";
        if (host) { size_t i = 0; for (; host[i] && host[i] != ':' && i < sizeof hl - 1; i++) hl[i] = (char)tolower((unsigned char)host[i]); hl[i] = 0; size_t n = strlen(hl); if (n && hl[n - 1] == '.') hl[n - 1] = 0; }
        if (!strncasecmp(target, "http://", 7) || !strncasecmp(target, "https://", 8)) {
            const char* hs = strstr(target, "://") + 3; size_t i = 0;
            for (; hs[i] && hs[i] != '/' && hs[i] != ':' && hs[i] != '?' && i < sizeof hl - 1; i++) hl[i] = (char)tolower((unsigned char)hs[i]);
            hl[i] = 0;
        }
        int vi = nzo_vhost_select(hl);
        char key[16384];
        int kl = nzo_cache_key(host, target, nzo_vhost_name(vi), nzo_vhost_proxy_host(vi), key, sizeof key);
        if (kl < 0) { simple_page(c, 400, 0); goto done; }
        uint8_t k128[16];
        nzo_md5(key, (size_t)kl, k128);
        o = find_obj(k128);
        if (!o) { simple_page(c, 504, keep); goto done; }
    }
    {
        nzo_object ob;
        memset(&ob, 0, sizeof ob);
        ob.stored_length = o->stored_len; ob.raw_length = o->raw_len; ob.content_type = o->ctype;
        ob.etag = o->has_etag ? o->etag : NULL; ob.last_modified = o->has_lm ? o->lm : NULL;
        ob.content_encoding = o->has_enc ? o->enc : NULL; ob.gunzip_enabled = g_gunzip; ob.no_ranges = o->no_ranges;
        plan = malloc(1 << 16);
        if (nzo_plan(&rq, &ob, g_boundary++, plan, 1 << 16) < 0) { simple_page(c, 400, 0); goto done; }
        int status = 0, headonly = 0;
        c->head = malloc(8192);
        int hl = 0;
        c->nseg = 0;
        char* save = NULL;
        char hdrs[4096]; int hdl = 0;
        for (char* ln = strtok_r(plan, "\n", &save); ln; ln = strtok_r(NULL, "\n", &save)) {
            if (!strncmp(ln, "status ", 7)) status = atoi(ln + 7);
            else if (!strncmp(ln, "repr ", 5)) c->inflated = !strcmp(ln + 5, "inflated");
            else if (!strncmp(ln, "head ", 5)) headonly = atoi(ln + 5);
            else if (!strncmp(ln, "hdr ", 4)) hdl += snprintf(hdrs + hdl, sizeof hdrs - (size_t)hdl, "%s\r\n", ln + 4);
            else if (!strncmp(ln, "lit ", 4) && c->nseg < MAXSEG) {
                size_t n = strlen(ln + 4) / 2;
                seg* s = &c->segs[c->nseg++];
                s->lit = 1; s->bytes = malloc(n + 1); s->len = (int64_t)n; s->off = 0;
                for (size_t i = 0; i < n; i++) s->bytes[i] = (char)(hexv(ln[4 + 2 * i]) * 16 + hexv(ln[5 + 2 * i]));
            } else if (!strncmp(ln, "slice ", 6) && c->nseg < MAXSEG) {
                long long so, sl;
                sscanf(ln + 6, "%lld %lld", &so, &sl);
                seg* s = &c->segs[c->nseg++];
                s->lit = 0; s->bytes = NULL; s->off = so; s->len = sl;
            }
        }
        char date[64];
        time_t now = time(NULL);
        struct tm tmv;
        gmtime_r(&now, &tmv);
        strftime(date, sizeof date, "%a, %d %b %Y %H:%M:%S GMT", &tmv);
        hl = snprintf(c->head, 8192, "HTTP/1.1 %d %s\r\nServer: nginx\r\nDate: %s\r\n%sConnection: %s\r\n\r\n", status, status_text(status), date,
                      hdrs, keep ? "keep-alive" : "close");
        c->head_len = hl; c->head_off = 0;
        if (headonly) { for (int i = 0; i < c->nseg; i++) free(c->segs[i].bytes); c->nseg = 0; }
        c->o = o; c->keep = keep; c->active = 1; c->cur = 0; c->seg_done = 0;
    }
done:
    free(plan); free(host); free(a); free(b); free(d); free(e); free(f); free(g); free(h);
}

/* push as much of the response as the socket takes; 1 = finished, 0 = would block, -1 = error */
static int pump(conn* c) {
    int budget = 16;   /* output buffers per event: round-robin fairness between connections of one worker */
    for (;;) {
        if (c->head_off < c->head_len) {
            ssize_t w = send(c->fd, c->head + c->head_off, (size_t)(c->head_len - c->head_off), MSG_NOSIGNAL);
            if (w < 0) return (errno == EAGAIN || errno == EWOULDBLOCK) ? 0 : -1;
            c->head_off += (int)w;
            continue;
        }
        if (c->out_off < c->out_len) {
            ssize_t w = send(c->fd, c->out[c->out_which] + c->out_off, (size_t)(c->out_len - c->out_off), MSG_NOSIGNAL);
            if (w < 0) return (errno == EAGAIN || errno == EWOULDBLOCK) ? 0 : -1;
            c->out_off += (int)w;
            continue;
        }
        if (c->cur >= c->nseg) return 1;
        seg* s = &c->segs[c->cur];
        if (c->seg_done >= s->len) { c->cur++; c->seg_done = 0; continue; }
        if (budget-- == 0) return 0;   /* yield; level-triggered EPOLLOUT brings us straight back */
        /* fill the other output buffer: the read loop of SPEC B1 */
        c->out_which ^= 1;
        int64_t want = s->len - c->seg_done;
        if (want > OUTBUF) want = OUTBUF;
        if (s->lit) {
            memcpy(c->out[c->out_which], s->bytes + c->seg_done, (size_t)want);
        } else {
            if (c->inflated) {
                if (!c->o->raw || s->off + c->seg_done + want > c->o->raw_len) return -1;
                memcpy(c->out[c->out_which], c->o->raw + s->off + c->seg_done, (size_t)want);
            } else {
                ssize_t r = pread(c->o->fd, c->out[c->out_which], (size_t)want, s->off + c->seg_done);
                if (r <= 0) return -1;
                want = r;
            }
        }
        c->out_len = (int)want; c->out_off = 0;
        c->seg_done += want;
    }
}

static void worker(int ls) {
    int ep = epoll_create1(0);
    struct epoll_event ev;
    ev.events = EPOLLIN; ev.data.ptr = NULL;
    epoll_ctl(ep, EPOLL_CTL_ADD, ls, &ev);
    struct epoll_event evs[256];
    for (;;) {
        int n = epoll_wait(ep, evs, 256, -1);
        for (int i = 0; i < n; i++) {
            conn* c = evs[i].data.ptr;
            if (!c) {
                for (;;) {
                    int fd = accept4(ls, NULL, NULL, SOCK_NONBLOCK);
                    if (fd < 0) break;
                    int one = 1;
                    setsockopt(fd, IPPROTO_TCP, TCP_NODELAY, &one, sizeof one);
                    conn* nc = calloc(1, sizeof(conn));
                    nc->fd = fd;
                    struct epoll_event e2;
                    e2.events = EPOLLIN; e2.data.ptr = nc;
                    epoll_ctl(ep, EPOLL_CTL_ADD, fd, &e2);
                }
                continue;
            }
            int dead = 0;
            if (!c->active && (evs[i].events & EPOLLIN)) {
                ssize_t r = recv(c->fd, c->in + c->in_len, sizeof c->in - 1 - (size_t)c->in_len, 0);
                if (r <= 0) { if (r == 0 || (errno != EAGAIN && errno != EWOULDBLOCK)) dead = 1; }
                else c->in_len += (int)r;
            }
            while (!dead) {
                if (!c->active) {
                    c->in[c->in_len] = 0;
                    char* end = strstr(c->in, "\r\n\r\n");
                    if (!end) { if (c->in_len >= (int)sizeof c->in - 1) dead = 1; break; }
                    end[2] = 0;
                    build_response(c, c->in);
                    int used = (int)(end + 4 - c->in);
                    memmove(c->in, c->in + used, (size_t)(c->in_len - used));
                    c->in_len -= used;
                }
                int st = pump(c);
                if (st < 0) { dead = 1; break; }
                struct epoll_event e2;
                e2.data.ptr = c;
                if (st == 0) { e2.events = EPOLLOUT; epoll_ctl(ep, EPOLL_CTL_MOD, c->fd, &e2); break; }
                int keep = c->keep;
                free_resp(c);
                e2.events = EPOLLIN; epoll_ctl(ep, EPOLL_CTL_MOD, c->fd, &e2);
                if (!keep) { dead = 1; break; }
                if (c->in_len == 0) break;
            }
            if (dead) { epoll_ctl(ep, EPOLL_CTL_DEL, c->fd, NULL); close(c->fd); free_resp(c); free(c); }
        }
    }
}

int main(int argc, char** argv) {
    const char* dir = NULL;
    const char* ip = "127.0.0.1";
    int port = 8080, workers = 1;
    for (int i = 1; i < argc; i++) {
        if (!strcmp(argv[i], "--root") && i + 1 < argc) dir = argv[++i];
        else if (!strcmp(argv[i], "--listen") && i + 1 < argc) { char* v = argv[++i]; char* c = strrchr(v, ':'); if (c) { *c = 0; ip = v; port = atoi(c + 1); } else port = atoi(v); }
        else if (!strcmp(argv[i], "--workers") && i + 1 < argc) workers = atoi(argv[++i]);
        else if (!strcmp(argv[i], "--gunzip")) g_gunzip = 1;
        else { fprintf(stderr, "nz_cpu_server --root DIR [--listen ip:port] [--workers N]\n"); return 2; }
    }
    signal(SIGPIPE, SIG_IGN);
    if (!dir || load_manifest(dir) < 0) { fprintf(stderr, "nz_cpu_server: cannot load %s/manifest\n", dir ? dir : "(null)"); return 1; }
    fprintf(stderr, "nz_cpu_server: %d objects, %d worker(s), %s:%d\n", g_nobj, workers, ip, port);
    for (int w = 0; w < workers; w++) {
        int ls = socket(AF_INET, SOCK_STREAM | SOCK_NONBLOCK, 0);
        int one = 1;
        setsockopt(ls, SOL_SOCKET, SO_REUSEADDR, &one, sizeof one);
        setsockopt(ls, SOL_SOCKET, SO_REUSEPORT, &one, sizeof one);
        struct sockaddr_in sa;
        memset(&sa, 0, sizeof sa);
        sa.sin_family = AF_INET; sa.sin_port = htons((uint16_t)port);
        inet_pton(AF_INET, ip, &sa.sin_addr);
        if (bind(ls, (struct sockaddr*)&sa, sizeof sa) || listen(ls, 1024)) { perror("bind/listen"); return 1; }
        if (workers == 1) { worker(ls); return 0; }
        pid_t p = fork();
        if (p == 0) { prctl(PR_SET_PDEATHSIG, SIGTERM); worker(ls); _exit(0); }
        close(ls);
    }
    while (wait(NULL) > 0) {}
    return 0;
}
