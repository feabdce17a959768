/*
 * oracle/nz_oracle.c — CPU restatement of the reference's cache-hit path.
 *
 * TEST INFRASTRUCTURE ONLY (see nz_oracle.h).  PARITY UNPINNED (see SPEC.md).
 *
 * What it follows.  The reference's data plane is the stock nginx binary driven
 * by /root/reference/examples/demo/nginx.conf; its arithmetic is in nginx
 * (version unpinned: examples/demo/nginx.yaml:30 `image: nginx`), zlib and the
 * RFCs, none of which are vendored.  Each function cites the conf line that
 * makes nginx do the work and the SPEC.md rule that restates nginx's published
 * behaviour.
 */
#define _GNU_SOURCE
#include "nz_oracle.h"

#include <ctype.h>
#include <errno.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <strings.h>
#include <time.h>
#include <unistd.h>
#include <zlib.h>

/* ======================================================================= */
/* MD5 — RFC 1321.  nginx.conf:38 (proxy_cache_path): cache identity = md5  */
/* of the proxy_cache_key string (SPEC K5).                                 */
/* ======================================================================= */
static uint32_t rol32(uint32_t x, int s) { return (x << s) | (x >> (32 - s)); }

static const uint32_t md5_k[64] = {
    0xd76aa478, 0xe8c7b756, 0x242070db, 0xc1bdceee, 0xf57c0faf, 0x4787c62a, 0xa8304613, 0xfd469501,
    0x698098d8, 0x8b44f7af, 0xffff5bb1, 0x895cd7be, 0x6b901122, 0xfd987193, 0xa679438e, 0x49b40821,
    0xf61e2562, 0xc040b340, 0x265e5a51, 0xe9b6c7aa, 0xd62f105d, 0x02441453, 0xd8a1e681, 0xe7d3fbc8,
    0x21e1cde6, 0xc33707d6, 0xf4d50d87, 0x455a14ed, 0xa9e3e905, 0xfcefa3f8, 0x676f02d9, 0x8d2a4c8a,
    0xfffa3942, 0x8771f681, 0x6d9d6122, 0xfde5380c, 0xa4beea44, 0x4bdecfa9, 0xf6bb4b60, 0xbebfbc70,
    0x289b7ec6, 0xeaa127fa, 0xd4ef3085, 0x04881d05, 0xd9d4d039, 0xe6db99e5, 0x1fa27cf8, 0xc4ac5665,
    0xf4292244, 0x432aff97, 0xab9423a7, 0xfc93a039, 0x655b59c3, 0x8f0ccc92, 0xffeff47d, 0x85845dd1,
    0x6fa87e4f, 0xfe2ce6e0, 0xa3014314, 0x4e0811a1, 0xf7537e82, 0xbd3af235, 0x2ad7d2bb, 0xeb86d391};
static const uint8_t md5_s[64] = {7, 12, 17, 22, 7, 12, 17, 22, 7, 12, 17, 22, 7, 12, 17, 22,
                                  5, 9,  14, 20, 5, 9,  14, 20, 5, 9,  14, 20, 5, 9,  14, 20,
                                  4, 11, 16, 23, 4, 11, 16, 23, 4, 11, 16, 23, 4, 11, 16, 23,
                                  6, 10, 15, 21, 6, 10, 15, 21, 6, 10, 15, 21, 6, 10, 15, 21};

static void md5_block(uint32_t st[4], const uint8_t* p) {
    uint32_t m[16];
    for (int i = 0; i < 16; i++)
        m[i] = (uint32_t)p[4 * i] | ((uint32_t)p[4 * i + 1] << 8) | ((uint32_t)p[4 * i + 2] << 16) |
               ((uint32_t)p[4 * i + 3] << 24);
    uint32_t a = st[0], b = st[1], c = st[2], d = st[3];
    for (int i = 0; i < 64; i++) {
        uint32_t f;
        int g;
        if (i < 16) { f = (b & c) | (~b & d); g = i; }
        else if (i < 32) { f = (d & b) | (~d & c); g = (5 * i + 1) & 15; }
        else if (i < 48) { f = b ^ c ^ d; g = (3 * i + 5) & 15; }
        else { f = c ^ (b | ~d); g = (7 * i) & 15; }
        uint32_t t = d;
        d = c;
        c = b;
        b = b + rol32(a + f + md5_k[i] + m[g], md5_s[i]);
        a = t;
    }
    st[0] += a; st[1] += b; st[2] += c; st[3] += d;
}

void nzo_md5(const void* data, size_t n, uint8_t out[16]) {
    uint32_t st[4] = {0x67452301, 0xefcdab89, 0x98badcfe, 0x10325476};
    const uint8_t* p = (const uint8_t*)data;
    size_t full = n / 64;
    for (size_t i = 0; i < full; i++) md5_block(st, p + 64 * i);
    uint8_t tail[128];
    size_t rem = n - full * 64;
    memset(tail, 0, sizeof tail);
    if (rem) memcpy(tail, p + full * 64, rem);
    tail[rem] = 0x80;
    size_t tl = (rem < 56) ? 64 : 128;
    uint64_t bits = (uint64_t)n * 8;
    for (int i = 0; i < 8; i++) tail[tl - 8 + i] = (uint8_t)(bits >> (8 * i));
    md5_block(st, tail);
    if (tl == 128) md5_block(st, tail + 64);
    for (int i = 0; i < 4; i++)
        for (int j = 0; j < 4; j++) out[4 * i + j] = (uint8_t)(st[i] >> (8 * j));
}

/* ======================================================================= */
/* CRC-32/IEEE of the key (SPEC K6) and CRC-32C (SPEC X1)                   */
/* ======================================================================= */
static uint32_t tab_ieee[256], tab_c[8][256];
static pthread_once_t tab_once = PTHREAD_ONCE_INIT;

static void tab_init(void) {
    for (uint32_t i = 0; i < 256; i++) {
        uint32_t a = i, b = i;
        for (int k = 0; k < 8; k++) {
            a = (a & 1) ? (a >> 1) ^ 0xEDB88320u : a >> 1;
            b = (b & 1) ? (b >> 1) ^ 0x82F63B78u : b >> 1;
        }
        tab_ieee[i] = a;
        tab_c[0][i] = b;
    }
    for (int t = 1; t < 8; t++)
        for (uint32_t i = 0; i < 256; i++)
            tab_c[t][i] = (tab_c[t - 1][i] >> 8) ^ tab_c[0][tab_c[t - 1][i] & 0xff];
}

uint32_t nzo_crc32_ieee(uint32_t crc, const void* data, size_t n) {
    pthread_once(&tab_once, tab_init);
    const uint8_t* p = (const uint8_t*)data;
    crc = ~crc;
    while (n--) crc = (crc >> 8) ^ tab_ieee[(crc ^ *p++) & 0xff];
    return ~crc;
}

/* definition-level form: one bit at a time (pins the table forms) */
uint32_t nzo_crc32c_bitwise(uint32_t crc, const void* data, size_t n) {
    const uint8_t* p = (const uint8_t*)data;
    crc = ~crc;
    while (n--) {
        crc ^= *p++;
        for (int k = 0; k < 8; k++) crc = (crc & 1) ? (crc >> 1) ^ 0x82F63B78u : crc >> 1;
    }
    return ~crc;
}

/* slice-by-8 table form */
uint32_t nzo_crc32c(uint32_t crc, const void* data, size_t n) {
    pthread_once(&tab_once, tab_init);
    const uint8_t* p = (const uint8_t*)data;
    crc = ~crc;
    while (n && ((uintptr_t)p & 7)) { crc = (crc >> 8) ^ tab_c[0][(crc ^ *p++) & 0xff]; n--; }
    while (n >= 8) {
        uint64_t w;
        memcpy(&w, p, 8);
        uint32_t lo = (uint32_t)w ^ crc, hi = (uint32_t)(w >> 32);
        crc = tab_c[7][lo & 0xff] ^ tab_c[6][(lo >> 8) & 0xff] ^ tab_c[5][(lo >> 16) & 0xff] ^
              tab_c[4][lo >> 24] ^ tab_c[3][hi & 0xff] ^ tab_c[2][(hi >> 8) & 0xff] ^
              tab_c[1][(hi >> 16) & 0xff] ^ tab_c[0][hi >> 24];
        p += 8;
        n -= 8;
    }
    while (n--) crc = (crc >> 8) ^ tab_c[0][(crc ^ *p++) & 0xff];
    return ~crc;
}

/* GF(2)[x] arithmetic mod the reflected Castagnoli polynomial: zlib 1.3's
 * crc32_combine (multmodp / x2nmodp) restated for poly 0x82F63B78 (SPEC X2). */
#define POLY_C 0x82F63B78u
static uint32_t multmodp(uint32_t a, uint32_t b) {
    uint32_t m = 1u << 31, p = 0;
    for (;;) {
        if (a & m) {
            p ^= b;
            if ((a & (m - 1)) == 0) break;
        }
        m >>= 1;
        b = (b & 1) ? (b >> 1) ^ POLY_C : b >> 1;
    }
    return p;
}
/* x^(n * 2^k) mod p.  zlib indexes its table with k & 31 because the IEEE
 * polynomial is irreducible (x^(2^32) == x).  The Castagnoli polynomial is
 * (x+1) * (primitive of degree 31), where that identity does NOT hold, so the
 * table simply holds every power 2^0 .. 2^71 by repeated squaring. */
static uint32_t x2n_tab[72];
static pthread_once_t x2n_once = PTHREAD_ONCE_INIT;
static void x2n_init(void) {
    uint32_t p = 1u << 30; /* x^1 */
    x2n_tab[0] = p;
    for (int i = 1; i < 72; i++) x2n_tab[i] = p = multmodp(p, p);
}
static uint32_t x2nmodp(uint64_t n, unsigned k) {
    pthread_once(&x2n_once, x2n_init);
    uint32_t p = 1u << 31; /* x^0 */
    while (n) {
        if (n & 1) p = multmodp(x2n_tab[k], p);
        n >>= 1;
        k++;
    }
    return p;
}
uint32_t nzo_crc32c_combine(uint32_t crc1, uint32_t crc2, uint64_t len2) {
    return multmodp(x2nmodp(len2, 3), crc1) ^ crc2;
}

uint32_t nzo_crc32c_chunks(const void* data, size_t n, size_t chunk, uint32_t* out) {
    const uint8_t* p = (const uint8_t*)data;
    uint32_t whole = 0;
    size_t i = 0;
    for (size_t off = 0; off < n; off += chunk, i++) {
        size_t l = n - off < chunk ? n - off : chunk;
        uint32_t c = nzo_crc32c(0, p + off, l);
        if (out) out[i] = c;
        whole = off == 0 ? c : nzo_crc32c_combine(whole, c, l);
    }
    return whole;
}

/* ======================================================================= */
/* vhosts, method gate (SPEC V1, V2) — examples/demo/nginx.conf:43-136       */
/* ======================================================================= */
static const char* const vhosts[5] = {"download.tensorflow.org", "yann.lecun.com",
                                      "www.cs.toronto.edu", "storage.googleapis.com",
                                      "images.cocodataset.org"};
/* proxy_pass host per vhost (nginx.conf:55,74,93,113,132); https://storage.googleapis.com
 * uses the scheme-default port so $proxy_host carries no port (SPEC K3). */
static const char* const vhost_proxy[5] = {"download.tensorflow.org", "yann.lecun.com",
                                           "www.cs.toronto.edu", "storage.googleapis.com",
                                           "images.cocodataset.org"};
int nzo_vhost_select(const char* host_lower) {
    if (host_lower)
        for (int i = 0; i < 5; i++)
            if (strcmp(host_lower, vhosts[i]) == 0) return i;
    return 0; /* first server block is the implicit default (SPEC V1) */
}
const char* nzo_vhost_name(int idx) { return (idx >= 0 && idx < 5) ? vhosts[idx] : NULL; }
const char* nzo_vhost_proxy_host(int idx) { return (idx >= 0 && idx < 5) ? vhost_proxy[idx] : NULL; }

int nzo_method_gate(const char* method) { /* nginx.conf:47-49: unanchored, case-sensitive */
    return strstr(method, "GET") != NULL || strstr(method, "HEAD") != NULL;
}

/* ======================================================================= */
/* cache key (SPEC K1-K4) — nginx.conf:56 "$host/$proxy_host$uri"            */
/* ======================================================================= */
static int hexval(int c) {
    if (c >= '0' && c <= '9') return c - '0';
    if (c >= 'a' && c <= 'f') return c - 'a' + 10;
    if (c >= 'A' && c <= 'F') return c - 'A' + 10;
    return -1;
}

/* $host: lower-case, strip :port and a trailing dot (SPEC K2) */
static int norm_host(const char* h, size_t n, char* out, size_t cap) {
    /* K2: the name ends at the first ':' outside an IPv6 literal or after its ']'; a NUL, a '/' or two dots in a
     * row anywhere in the header value (port part included) -> 400; one trailing dot is dropped */
    size_t end = n, dot = n;
    int st = 0; /* 0 name, 1 inside [..], 2 after the name */
    for (size_t i = 0; i < n; i++) {
        unsigned char c = (unsigned char)h[i];
        if (c == '.') {
            if (i > 0 && dot == i - 1) return -400;
            dot = i;
        } else if (c == ':') {
            if (st == 0) { end = i; st = 2; }
        } else if (c == '[') {
            if (i == 0) st = 1;
        } else if (c == ']') {
            if (st == 1) { end = i + 1; st = 2; }
        } else if (c == 0 || c == '/') {
            return -400;
        }
    }
    if (end && dot == end - 1) end--;
    if (end == 0 || end + 1 > cap) return -400;
    for (size_t i = 0; i < end; i++) {
        unsigned char c = (unsigned char)h[i];
        out[i] = (c >= 'A' && c <= 'Z') ? (char)(c | 0x20) : (char)c;
    }
    out[end] = 0;
    return (int)end;
}

/* ---------------------------------------------------------------------------
 * SPEC P: the request head.  Lines are walked in place; field values are (pointer, length) pairs into raw.
 * ------------------------------------------------------------------------- */
typedef struct { const char* p; size_t n; int set; } hfield;
static int name_is(const char* a, size_t n, const char* lit) {
    return strlen(lit) == n && strncasecmp(a, lit, n) == 0;
}
static int contains_ci(const char* p, size_t n, const char* needle) {
    size_t m = strlen(needle);
    for (size_t i = 0; i + m <= n; i++) if (strncasecmp(p + i, needle, m) == 0) return 1;
    return 0;
}
int nzo_parse_head(const char* raw, size_t n, char* out, size_t cap) {
    static const char* const kept[8] = {"accept-encoding", "host", "if-match", "if-modified-since", "if-none-match", "if-range", "if-unmodified-since", "range"};
    static const char* const uniq[9] = {"if-modified-since", "if-unmodified-since", "if-match", "if-none-match", "if-range", "content-length", "content-range", "expect", "authorization"};
    hfield kf[8];
    int useen[9];
    memset(kf, 0, sizeof kf);
    memset(useen, 0, sizeof useen);
    int error = 0, minor = 1, keep = 1, body = 0, hosts = 0;
    const char *method = NULL, *target = NULL;
    size_t mlen = 0, tlen = 0;
    while (n && (raw[0] == '\r' || raw[0] == '\n')) { raw++; n--; }          /* P: CR/LF before the request line */
    while (n >= 2 && raw[n - 2] == '\r' && raw[n - 1] == '\n') n -= 2;
    for (size_t i = 0; i < n; i++) {                                         /* P2: NUL, bare CR, bare LF */
        if (raw[i] == 0) { error = 400; goto emit; }
        if (raw[i] == '\r' && (i + 1 >= n || raw[i + 1] != '\n')) { error = 400; goto emit; }
        if (raw[i] == '\n' && (i == 0 || raw[i - 1] != '\r')) { error = 400; goto emit; }
    }
    {
        size_t le = 0;
        while (le < n && raw[le] != '\r') le++;
        size_t j = 0;
        while (j < le && ((raw[j] >= 'A' && raw[j] <= 'Z') || raw[j] == '_')) j++;
        if (j == 0 || j >= le || raw[j] != ' ') { error = 400; goto emit; }   /* P1 */
        method = raw; mlen = j;
        while (j < le && raw[j] == ' ') j++;
        size_t k = j;
        while (k < le && raw[k] != ' ') k++;
        target = raw + j; tlen = k - j;
        if (tlen == 0) { error = 400; goto emit; }
        if (target[0] != '/') {                                               /* P2: absolute-form */
            size_t m = 0;
            while (m < tlen && isalpha((unsigned char)target[m]) && (unsigned char)target[m] < 128) m++;
            if (m == 0 || m + 3 > tlen || memcmp(target + m, "://", 3) != 0) { error = 400; goto emit; }
        }
        size_t a = k, b = le;
        while (a < b && raw[a] == ' ') a++;
        while (b > a && raw[b - 1] == ' ') b--;
        if (a == b) {                                                         /* P3: HTTP/0.9 */
            if (!(mlen == 3 && memcmp(method, "GET", 3) == 0)) { error = 400; goto emit; }
            minor = -1; keep = 0;
        } else {                                                              /* P4 */
            const char* v = raw + a;
            size_t vn = b - a;
            if (vn < 8 || memcmp(v, "HTTP/", 5) != 0) { error = 400; goto emit; }
            size_t q = 5, d1 = 0, d2 = 0;
            int major = 0, mn = 0;
            while (q < vn && isdigit((unsigned char)v[q]) && d1 < 3) { major = major * 10 + (v[q] - '0'); q++; d1++; }
            if (d1 < 1 || d1 > 2 || v[5] == '0' || q >= vn || v[q] != '.') { error = 400; goto emit; }
            q++;
            while (q < vn && isdigit((unsigned char)v[q]) && d2 < 3) { mn = mn * 10 + (v[q] - '0'); q++; d2++; }
            if (d2 < 1 || d2 > 2 || q != vn) { error = 400; goto emit; }
            if (major > 1) { error = 505; goto emit; }
            minor = mn == 0 ? 0 : 1;
            keep = minor == 1;
        }
        size_t pos = le;
        while (pos < n) {                                                     /* P5-P10 */
            pos += 2;                                                         /* the CRLF that ended the previous line */
            size_t e = pos;
            while (e < n && raw[e] != '\r') e++;
            if (e == pos) break;
            const char* ln = raw + pos;
            size_t ll = e - pos, c = 0;
            pos = e;
            while (c < ll && ln[c] != ':') c++;
            size_t nl = c;
            const char* val = c < ll ? ln + c + 1 : ln + ll;
            size_t vl = c < ll ? ll - c - 1 : 0;
            int valid = nl > 0;
            for (size_t i = 0; i < nl && valid; i++) {
                unsigned char ch = (unsigned char)ln[i];
                if (!(ch < 128 && (isalnum(ch) || ch == '-'))) valid = 0;
            }
            if (!valid) continue;
            while (vl && val[0] == ' ') { val++; vl--; }
            while (vl && val[vl - 1] == ' ') vl--;
            if (name_is(ln, nl, "host")) {
                char tmp[1024];
                hosts++;
                if (vl >= sizeof tmp || norm_host(val, vl, tmp, sizeof tmp) < 0) { error = 400; goto emit; }
                if (hosts == 1) { kf[1].p = val; kf[1].n = vl; kf[1].set = 1; }
                continue;
            }
            if (name_is(ln, nl, "range") || name_is(ln, nl, "accept-encoding")) {   /* P6 */
                hfield* f = name_is(ln, nl, "range") ? &kf[7] : &kf[0];
                if (!f->set) { f->p = val; f->n = vl; f->set = 1; }
                continue;
            }
            int u = -1;
            for (int i = 0; i < 9; i++) if (name_is(ln, nl, uniq[i])) u = i;
            if (u >= 0) {                                                     /* P7 */
                if (useen[u]) { error = 400; goto emit; }
                useen[u] = 1;
                for (int i = 0; i < 8; i++) if (name_is(ln, nl, kept[i])) { kf[i].p = val; kf[i].n = vl; kf[i].set = 1; }
                if (u == 5) {                                                 /* P10 */
                    if (vl == 0) { error = 400; goto emit; }
                    int nz = 0;
                    for (size_t i = 0; i < vl; i++) { if (!isdigit((unsigned char)val[i])) { error = 400; goto emit; } if (val[i] != '0') nz = 1; }
                    if (nz) body = 1;
                }
                continue;
            }
            if (name_is(ln, nl, "transfer-encoding")) { body = 1; continue; }
            if (name_is(ln, nl, "connection")) {                              /* P9 */
                if (contains_ci(val, vl, "close")) keep = 0;
                else if (contains_ci(val, vl, "keep-alive")) keep = 1;
            }
        }
        if (minor == 1 && hosts == 0) error = 400;                            /* P8 */
    }
emit:;
    size_t o = 0;
    int w = snprintf(out, cap, "error %d\n", error);
    if (w < 0 || (size_t)w >= cap) return -1;
    o = (size_t)w;
    if (error) return (int)o;
    w = snprintf(out + o, cap - o, "method %.*s\ntarget %.*s\nminor %d\nkeep %d\nbody %d\n", (int)mlen, method, (int)tlen, target, minor, keep, body);
    if (w < 0 || (size_t)w >= cap - o) return -1;
    o += (size_t)w;
    for (int i = 0; i < 8; i++) {
        if (!kf[i].set) continue;
        w = snprintf(out + o, cap - o, "f %s: %.*s\n", kept[i], (int)kf[i].n, kf[i].p);
        if (w < 0 || (size_t)w >= cap - o) return -1;
        o += (size_t)w;
    }
    return (int)o;
}

/* $uri: decode, resolve dot segments, merge slashes, cut at ? or # (SPEC K4) */
static int norm_uri(const char* p, size_t n, char* out, size_t cap) {
    if (n == 0 || p[0] != '/') return -400;
    size_t o = 0;
    size_t i = 0;
    while (i < n) {
        int c = (unsigned char)p[i];
        if (c == '?' || c == '#') break;
        if (c == '%') {
            if (i + 2 >= n) return -400;
            int h = hexval((unsigned char)p[i + 1]), l = hexval((unsigned char)p[i + 2]);
            if (h < 0 || l < 0) return -400;
            c = h * 16 + l;
            i += 3;
            if (c == 0) return -400;
        } else {
            i++;
        }
        if (o + 2 > cap) return -400;
        out[o++] = (char)c;
    }
    out[o] = 0;
    /* second pass over decoded bytes: segment logic (decoded %2F counts as a slash, as in nginx) */
    size_t w = 0, r = 0;
    while (r < o) {
        /* at a '/' */
        if (out[r] == '/') {
            size_t s = r + 1;
            size_t e = s;
            while (e < o && out[e] != '/') e++;
            size_t len = e - s;
            if (len == 0 && e < o) { /* "//" -> merge */
                r = e;
                continue;
            }
            if (len == 1 && out[s] == '.') { /* "/." */
                r = e;
                if (e == o) out[w++] = '/';
                continue;
            }
            if (len == 2 && out[s] == '.' && out[s + 1] == '.') { /* "/.." */
                if (w == 0) return -400;
                while (w > 0 && out[w - 1] != '/') w--;
                if (w == 0) return -400;
                w--; /* drop the slash that began the removed segment */
                r = e;
                if (e == o) out[w++] = '/';
                continue;
            }
            out[w++] = '/';
            memmove(out + w, out + s, len);
            w += len;
            r = e;
        } else {
            return -400;
        }
    }
    if (w == 0) out[w++] = '/';
    out[w] = 0;
    return (int)w;
}

int nzo_cache_key(const char* host_hdr, const char* request_target, const char* server_name,
                  const char* proxy_host, char* out, size_t cap) {
    char host[512], uri[8192];
    const char* path = request_target;
    size_t plen = strlen(request_target);
    int hl = -1;
    /* absolute-form target: host comes from the request line */
    if (plen > 7 && (strncasecmp(request_target, "http://", 7) == 0 ||
                     (plen > 8 && strncasecmp(request_target, "https://", 8) == 0))) {
        const char* h = strstr(request_target, "://") + 3;
        const char* e = h;
        while (*e && *e != '/' && *e != '?' && *e != '#') e++;
        hl = norm_host(h, (size_t)(e - h), host, sizeof host);
        if (hl < 0) return -400;
        if (*e == '/') { path = e; plen = strlen(e); }
        else { path = "/"; plen = 1; }
    } else if (host_hdr) { /* a Host header that is present but empty is invalid (400), not absent */
        hl = norm_host(host_hdr, strlen(host_hdr), host, sizeof host);
        if (hl < 0) return -400;
    } else {
        if (!server_name) return -400;
        snprintf(host, sizeof host, "%s", server_name);
    }
    int ul = norm_uri(path, plen, uri, sizeof uri);
    if (ul < 0) return -400;
    int n = snprintf(out, cap, "%s/%s%s", host, proxy_host, uri);
    if (n < 0 || (size_t)n >= cap) return -400;
    return n;
}

/* ======================================================================= */
/* Range parser (SPEC R3-R7) — nginx's built-in range filter; active for    */
/* every `location /` of nginx.conf because no `max_ranges 0` is set.        */
/* ======================================================================= */
int nzo_range_parse(const char* p, int64_t content_length, nzo_range* out, int max_out, int* n_out) {
    const int64_t maxv = INT64_MAX;
    const int64_t cutoff = maxv / 10;
    const int cutlim = (int)(maxv % 10);
    int64_t size = 0;
    int n = 0;
    for (;;) {
        int64_t start = 0, end = 0;
        int suffix = 0, open_end = 0;
        while (*p == ' ') p++;
        if (*p != '-') {
            if (*p < '0' || *p > '9') return NZO_RANGE_416;
            while (*p >= '0' && *p <= '9') {
                if (start >= cutoff && (start > cutoff || *p - '0' > cutlim)) return NZO_RANGE_416;
                start = start * 10 + (*p++ - '0');
            }
            while (*p == ' ') p++;
            if (*p++ != '-') return NZO_RANGE_416;
            while (*p == ' ') p++;
            if (*p == ',' || *p == '\0') {
                end = content_length;
                open_end = 1;
            }
        } else {
            suffix = 1;
            p++;
        }
        if (!open_end) {
            if (*p < '0' || *p > '9') return NZO_RANGE_416;
            while (*p >= '0' && *p <= '9') {
                if (end >= cutoff && (end > cutoff || *p - '0' > cutlim)) return NZO_RANGE_416;
                end = end * 10 + (*p++ - '0');
            }
            while (*p == ' ') p++;
            if (*p != ',' && *p != '\0') return NZO_RANGE_416;
            if (suffix) {
                start = (end < content_length) ? content_length - end : 0;
                end = content_length - 1;
            }
            if (end >= content_length) end = content_length;
            else end++;
        }
        if (start < end) {
            if (n >= max_out) return NZO_RANGE_DECLINED;
            out[n].start = start;
            out[n].end = end;
            n++;
            if (size > maxv - (end - start)) return NZO_RANGE_416;
            size += end - start;
        } else if (start == 0) {
            return NZO_RANGE_DECLINED;
        }
        if (*p++ != ',') break;
    }
    *n_out = n;
    if (n == 0) return NZO_RANGE_416;
    if (size > content_length) return NZO_RANGE_DECLINED;
    return NZO_RANGE_OK;
}

/* ======================================================================= */
/* HTTP dates (SPEC C7)                                                     */
/* ======================================================================= */
static int64_t days_from_civil(int64_t y, unsigned m, unsigned d) {
    y -= m <= 2;
    const int64_t era = (y >= 0 ? y : y - 399) / 400;
    const unsigned yoe = (unsigned)(y - era * 400);
    const unsigned doy = (153 * (m + (m > 2 ? -3 : 9)) + 2) / 5 + d - 1;
    const unsigned doe = yoe * 365 + yoe / 4 - yoe / 100 + doy;
    return era * 146097 + (int64_t)doe - 719468;
}
static int month_of(const char* p) {
    static const char* mon = "JanFebMarAprMayJunJulAugSepOctNovDec";
    for (int i = 0; i < 12; i++)
        if (memcmp(p, mon + 3 * i, 3) == 0) return i + 1;
    return 0;
}
static int two(const char* p) {
    if (p[0] < '0' || p[0] > '9' || p[1] < '0' || p[1] > '9') return -1;
    return (p[0] - '0') * 10 + (p[1] - '0');
}
int64_t nzo_parse_http_time(const char* s, size_t n) {
    const char* p = s;
    const char* e = s + n;
    enum { RFC822, RFC850, ISOC } fmt = RFC822;
    int day = 0, mon = 0, year = 0, hh, mm, ss;
    for (; p < e; p++) {
        if (*p == ',') break;
        if (*p == ' ') { fmt = ISOC; break; }
    }
    for (p++; p < e; p++)
        if (*p != ' ') break;
    if (e - p < 18) return -1;
    if (fmt != ISOC) {
        day = two(p);
        if (day < 0) return -1;
        p += 2;
        if (*p == ' ') {
            if (e - p < 18) return -1;
            fmt = RFC822;
        } else if (*p == '-') {
            fmt = RFC850;
        } else {
            return -1;
        }
        p++;
    }
    mon = month_of(p);
    if (!mon) return -1;
    p += 3;
    if ((fmt == RFC822 && *p != ' ') || (fmt == RFC850 && *p != '-')) return -1;
    p++;
    if (fmt == RFC822) {
        int a = two(p), b = two(p + 2);
        if (a < 0 || b < 0) return -1;
        year = a * 100 + b;
        p += 4;
    } else if (fmt == RFC850) {
        year = two(p);
        if (year < 0) return -1;
        year += (year < 70) ? 2000 : 1900;
        p += 2;
    }
    if (fmt == ISOC) {
        if (*p == ' ') p++;
        if (*p < '0' || *p > '9') return -1;
        day = *p++ - '0';
        if (*p != ' ') {
            if (*p < '0' || *p > '9') return -1;
            day = day * 10 + (*p++ - '0');
        }
        if (e - p < 14) return -1;
    }
    if (*p++ != ' ') return -1;
    hh = two(p);
    if (hh < 0 || p[2] != ':') return -1;
    mm = two(p + 3);
    if (mm < 0 || p[5] != ':') return -1;
    ss = two(p + 6);
    if (ss < 0) return -1;
    p += 8;
    if (fmt == ISOC) {
        if (*p++ != ' ') return -1;
        int a = two(p), b = two(p + 2);
        if (a < 0 || b < 0) return -1;
        year = a * 100 + b;
    }
    if (hh > 23 || mm > 59 || ss > 59) return -1;
    if (day == 29 && mon == 2) {
        if ((year & 3) || ((year % 100 == 0) && (year % 400) != 0)) return -1;
    } else if (day > 30 + ((mon ^ (mon >> 3)) & 1) || (mon == 2 && day > 28)) { /* C7: no lower bound on the day, like ngx_parse_http_time */
        return -1;
    }
    return days_from_civil(year, (unsigned)mon, (unsigned)day) * 86400 + hh * 3600 + mm * 60 + ss;
}

/* ======================================================================= */
/* entity-tag list match (SPEC C6)                                          */
/* ======================================================================= */
static int etag_list_match(const char* list, const char* etag, int weak) {
    size_t ln = strlen(list);
    if (ln == 1 && list[0] == '*') return 1;
    if (!etag) return 0;
    const char* ed = etag;
    size_t el = strlen(etag);
    if (weak && el > 2 && ed[0] == 'W' && ed[1] == '/') { ed += 2; el -= 2; }
    const char* s = list;
    const char* e = list + ln;
    while (s < e) {
        if (weak && e - s > 2 && s[0] == 'W' && s[1] == '/') s += 2;
        if (el > (size_t)(e - s)) return 0;
        if (strncmp(s, ed, el) == 0) {
            const char* q = s + el;
            while (q < e && (*q == ' ' || *q == '\t')) q++;
            if (q == e || *q == ',') return 1;
        }
        while (s < e && *s != ',') s++;
        while (s < e && (*s == ' ' || *s == '\t' || *s == ',')) s++;
    }
    return 0;
}

/* SPEC G1: the Accept-Encoding test of ngx_http_gzip_ok, which the gunzip filter consults */
static unsigned gzip_qvalue(const char* p, const char* last) {
    char c = *p++;
    if (c != '0' && c != '1') return 0;
    unsigned q = (unsigned)(c - '0') * 100, n = 0;
    if (p == last) return q;
    c = *p++;
    if (c == ',' || c == ' ') return q;
    if (c != '.') return 0;
    while (p < last) {
        c = *p++;
        if (c == ',' || c == ' ') break;
        if (c < '0' || c > '9') return 0;
        q += (unsigned)(c - '0');
        n++;
    }
    return (q > 100 || n > 3) ? 0 : q;
}
static int has_gzip_token(const char* ae) {
    if (!ae) return 0;
    size_t len = strlen(ae);
    if (len < 4) return 0;
    if (len >= 5 && memcmp(ae, "gzip,", 5) == 0) return 1;
    const char *start = ae, *last = ae + len, *p;
    for (;;) {
        p = strcasestr(start, "gzip");
        if (!p) return 0;
        if (p == start || p[-1] == ',' || p[-1] == ' ') break;
        start = p + 4;
    }
    p += 4;
    int stage = 0; /* 0 after the token, 1 after ';' */
    while (p < last) {
        char c = *p++;
        if (stage == 0) {
            if (c == ',') return 1;
            if (c == ';') { stage = 1; continue; }
            if (c == ' ') continue;
            return 0;
        }
        if (c == 'q' || c == 'Q') {
            if (p + 2 > last || *p++ != '=') return 0;
            return gzip_qvalue(p, last) != 0;
        }
        if (c == ' ') continue;
        return 0;
    }
    return 1;
}

/* ======================================================================= */
/* plan serialisation                                                       */
/* ======================================================================= */
typedef struct sb { char* p; size_t cap, len; int ovf; } sb;
static void sb_put(sb* b, const char* s, size_t n) {
    if (b->len + n + 1 > b->cap) { b->ovf = 1; return; }
    memcpy(b->p + b->len, s, n);
    b->len += n;
    b->p[b->len] = 0;
}
static void sb_fmt(sb* b, const char* fmt, ...) __attribute__((format(printf, 2, 3)));
#include <stdarg.h>
static void sb_fmt(sb* b, const char* fmt, ...) {
    char tmp[1024];
    va_list ap;
    va_start(ap, fmt);
    int n = vsnprintf(tmp, sizeof tmp, fmt, ap);
    va_end(ap);
    if (n < 0 || (size_t)n >= sizeof tmp) { b->ovf = 1; return; }
    sb_put(b, tmp, (size_t)n);
}
static void sb_lit(sb* b, const char* s, size_t n) {
    sb_put(b, "lit ", 4);
    for (size_t i = 0; i < n; i++) sb_fmt(b, "%02x", (unsigned char)s[i]);
    sb_put(b, "\n", 1);
}

static const char* reason(int code) {
    switch (code) {
    case 400: return "Bad Request";
    case 405: return "Not Allowed";
    case 412: return "Precondition Failed";
    case 416: return "Requested Range Not Satisfiable";
    default: return "";
    }
}
static void error_page(sb* b, int code, int head, const char* extra_hdr) { /* SPEC E1 */
    char page[512];
    int n = snprintf(page, sizeof page,
                     "<html>\r\n<head><title>%d %s</title></head>\r\n<body>\r\n"
                     "<center><h1>%d %s</h1></center>\r\n<hr><center>nginx</center>\r\n"
                     "</body>\r\n</html>\r\n",
                     code, reason(code), code, reason(code));
    sb_fmt(b, "status %d\nhead %d\nrepr identity\n", code, head);
    sb_fmt(b, "hdr Content-Type: text/html\n");
    sb_fmt(b, "hdr Content-Length: %d\n", n);
    if (extra_hdr) sb_fmt(b, "hdr %s\n", extra_hdr);
    sb_lit(b, page, (size_t)n);
    sb_put(b, "end\n", 4);
}

int nzo_plan(const nzo_request* rq, const nzo_object* ob, uint64_t boundary, char* out, size_t cap) {
    sb b = {out, cap, 0, 0};
    if (cap) out[0] = 0;
    /* V2: method gate */
    if (!nzo_method_gate(rq->method)) {
        error_page(&b, 405, 0, NULL);
        return b.ovf ? -1 : (int)b.len;
    }
    int is_get = strcmp(rq->method, "GET") == 0, is_head = strcmp(rq->method, "HEAD") == 0;
    if (!is_get && !is_head) {
        sb_fmt(&b, "status 0\nbypass\nend\n");
        return b.ovf ? -1 : (int)b.len;
    }
    int head = is_head;

    /* representation: SPEC G1/G3 vs B3 */
    int gz_stored = ob->content_encoding && strcasecmp(ob->content_encoding, "gzip") == 0;
    int inflated = gz_stored && ob->gunzip_enabled && ob->raw_length >= 0 && !has_gzip_token(rq->accept_encoding);
    int64_t L = inflated ? ob->raw_length : ob->stored_length;
    char etag_buf[300];
    const char* etag = ob->etag;
    if (inflated && etag && !(etag[0] == 'W' && etag[1] == '/')) { /* G3: weaken */
        snprintf(etag_buf, sizeof etag_buf, "W/%s", etag);
        etag = etag_buf;
    }
    int64_t lm = ob->last_modified ? nzo_parse_http_time(ob->last_modified, strlen(ob->last_modified)) : -1;

    /* C1, C2 -> 412 */
    if (rq->if_unmodified_since) {
        int64_t t = nzo_parse_http_time(rq->if_unmodified_since, strlen(rq->if_unmodified_since));
        if (lm == -1 || t < lm) { /* C1: no stored date, or an unparseable one (-1), also fails the precondition */
            error_page(&b, 412, head, NULL);
            return b.ovf ? -1 : (int)b.len;
        }
    }
    if (rq->if_match && !etag_list_match(rq->if_match, etag, 0)) {
        error_page(&b, 412, head, NULL);
        return b.ovf ? -1 : (int)b.len;
    }
    /* C3 -> 304 */
    if (rq->if_modified_since || rq->if_none_match) {
        int serve_full = 0;
        if (rq->if_modified_since) {
            int64_t t = nzo_parse_http_time(rq->if_modified_since, strlen(rq->if_modified_since));
            if (lm == -1 || t != lm) serve_full = 1; /* C5 */
        }
        if (!serve_full && rq->if_none_match && !etag_list_match(rq->if_none_match, etag, 1)) serve_full = 1;
        if (!serve_full) {
            sb_fmt(&b, "status 304\nhead %d\nrepr %s\n", head, inflated ? "inflated" : "identity");
            if (etag) sb_fmt(&b, "hdr ETag: %s\n", etag);
            if (ob->last_modified) sb_fmt(&b, "hdr Last-Modified: %s\n", ob->last_modified);
            sb_put(&b, "end\n", 4);
            return b.ovf ? -1 : (int)b.len;
        }
    }

    /* R1, R2 */
    nzo_range rg[64];
    int nr = 0, rc = NZO_RANGE_DECLINED;
    if (ob->no_ranges) { /* R0: the range filter is not in play at all - full 200 and no Accept-Ranges line */
        rc = NZO_RANGE_DECLINED;
    } else if (rq->http_minor >= 0 && rq->range && strlen(rq->range) >= 7 && strncasecmp(rq->range, "bytes=", 6) == 0) {
        int use = 1;
        if (rq->if_range) {
            size_t n = strlen(rq->if_range);
            if (n && rq->if_range[n - 1] == '"') {
                if (!etag || strlen(etag) != n || memcmp(etag, rq->if_range, n) != 0) use = 0;
            } else {
                int64_t t = nzo_parse_http_time(rq->if_range, n);
                if (lm == -1 || t != lm) use = 0;
            }
        }
        if (use) rc = nzo_range_parse(rq->range + 6, L, rg, 64, &nr);
    }
    const char* repr = inflated ? "inflated" : "identity";
    if (rc == NZO_RANGE_416) { /* R6 */
        char cr[96];
        snprintf(cr, sizeof cr, "Content-Range: bytes */%lld", (long long)L);
        error_page(&b, 416, head, cr);
        return b.ovf ? -1 : (int)b.len;
    }
    if (rc == NZO_RANGE_DECLINED) { /* full 200 */
        sb_fmt(&b, "status 200\nhead %d\nrepr %s\n", head, repr);
        if (ob->content_type && *ob->content_type) sb_fmt(&b, "hdr Content-Type: %s\n", ob->content_type);
        sb_fmt(&b, "hdr Content-Length: %lld\n", (long long)L);
        if (etag) sb_fmt(&b, "hdr ETag: %s\n", etag);
        if (ob->last_modified) sb_fmt(&b, "hdr Last-Modified: %s\n", ob->last_modified);
        if (!ob->no_ranges) sb_fmt(&b, "hdr Accept-Ranges: bytes\n");
        if (gz_stored && !inflated) sb_fmt(&b, "hdr Content-Encoding: gzip\n");
        if (L > 0) sb_fmt(&b, "slice 0 %lld\n", (long long)L);
        sb_put(&b, "end\n", 4);
        return b.ovf ? -1 : (int)b.len;
    }
    if (nr == 1) { /* R8 */
        sb_fmt(&b, "status 206\nhead %d\nrepr %s\n", head, repr);
        if (ob->content_type && *ob->content_type) sb_fmt(&b, "hdr Content-Type: %s\n", ob->content_type);
        sb_fmt(&b, "hdr Content-Length: %lld\n", (long long)(rg[0].end - rg[0].start));
        sb_fmt(&b, "hdr Content-Range: bytes %lld-%lld/%lld\n", (long long)rg[0].start,
               (long long)(rg[0].end - 1), (long long)L);
        if (etag) sb_fmt(&b, "hdr ETag: %s\n", etag);
        if (ob->last_modified) sb_fmt(&b, "hdr Last-Modified: %s\n", ob->last_modified);
        if (gz_stored && !inflated) sb_fmt(&b, "hdr Content-Encoding: gzip\n");
        sb_fmt(&b, "slice %lld %lld\n", (long long)rg[0].start, (long long)(rg[0].end - rg[0].start));
        sb_put(&b, "end\n", 4);
        return b.ovf ? -1 : (int)b.len;
    }
    /* R9 multipart */
    char bnd[32];
    snprintf(bnd, sizeof bnd, "%020llu", (unsigned long long)boundary);
    char part[64][400];
    int plen[64];
    int64_t total = 0;
    for (int i = 0; i < nr; i++) {
        if (ob->content_type && *ob->content_type)
            plen[i] = snprintf(part[i], sizeof part[i],
                               "\r\n--%s\r\nContent-Type: %s\r\nContent-Range: bytes %lld-%lld/%lld\r\n\r\n", bnd,
                               ob->content_type, (long long)rg[i].start, (long long)(rg[i].end - 1), (long long)L);
        else
            plen[i] = snprintf(part[i], sizeof part[i], "\r\n--%s\r\nContent-Range: bytes %lld-%lld/%lld\r\n\r\n",
                               bnd, (long long)rg[i].start, (long long)(rg[i].end - 1), (long long)L);
        total += plen[i] + (rg[i].end - rg[i].start);
    }
    char trailer[64];
    int tl = snprintf(trailer, sizeof trailer, "\r\n--%s--\r\n", bnd);
    total += tl;
    sb_fmt(&b, "status 206\nhead %d\nrepr %s\n", head, repr);
    sb_fmt(&b, "hdr Content-Type: multipart/byteranges; boundary=%s\n", bnd);
    sb_fmt(&b, "hdr Content-Length: %lld\n", (long long)total);
    if (etag) sb_fmt(&b, "hdr ETag: %s\n", etag);
    if (ob->last_modified) sb_fmt(&b, "hdr Last-Modified: %s\n", ob->last_modified);
    if (gz_stored && !inflated) sb_fmt(&b, "hdr Content-Encoding: gzip\n");
    for (int i = 0; i < nr; i++) {
        sb_lit(&b, part[i], (size_t)plen[i]);
        sb_fmt(&b, "slice %lld %lld\n", (long long)rg[i].start, (long long)(rg[i].end - rg[i].start));
    }
    sb_lit(&b, trailer, (size_t)tl);
    sb_put(&b, "end\n", 4);
    return b.ovf ? -1 : (int)b.len;
}

/* ======================================================================= */
/* gunzip (SPEC G2, G4): zlib inflate with wbits=31, single member          */
/* ======================================================================= */
int64_t nzo_gzip_isize(const void* src, size_t n) {
    if (n < 18) return -1;
    const uint8_t* p = (const uint8_t*)src + n - 4;
    return (int64_t)((uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24));
}
int64_t nzo_gunzip(const void* src, size_t n, void* dst, size_t cap) {
    z_stream z;
    memset(&z, 0, sizeof z);
    if (inflateInit2(&z, 31) != Z_OK) return -1;
    const uint8_t* in = (const uint8_t*)src;
    uint8_t* o = (uint8_t*)dst;
    size_t ip = 0, op = 0;
    int rc = Z_OK;
    uint8_t dummy;
    do { /* feed in <=1 GiB windows: z_stream counters are 32-bit */
        size_t ia = n - ip > (1u << 30) ? (1u << 30) : n - ip;
        size_t oa = cap - op > (1u << 30) ? (1u << 30) : cap - op;
        z.next_in = (Bytef*)(in + ip);
        z.avail_in = (uInt)ia;
        z.next_out = oa ? o + op : &dummy;
        z.avail_out = (uInt)oa;
        rc = inflate(&z, Z_NO_FLUSH);
        ip += ia - z.avail_in;
        op += oa - z.avail_out;
        if (rc == Z_BUF_ERROR && ia == z.avail_in && oa == z.avail_out) break;
    } while (rc == Z_OK || rc == Z_BUF_ERROR);
    inflateEnd(&z);
    if (rc != Z_STREAM_END) return -1;
    if (ip != n) return -1; /* trailing bytes / second member: declined (G4) */
    return (int64_t)op;
}

/* ======================================================================= */
/* the hot loop (SPEC B1): nginx with sendfile off reads the cache file in   */
/* output_buffers (2 x 32k default) and writes them to the connection.       */
/* ======================================================================= */
#define NZO_OUTBUF 32768
int64_t nzo_serve_slice(int fd, int64_t off, int64_t len, uint8_t* dst) {
    uint8_t buf[NZO_OUTBUF];
    int64_t done = 0;
    while (done < len) {
        size_t want = len - done > NZO_OUTBUF ? NZO_OUTBUF : (size_t)(len - done);
        ssize_t r = pread(fd, buf, want, off + done);
        if (r < 0) {
            if (errno == EINTR) continue;
            return -1;
        }
        if (r == 0) return -1;
        memcpy(dst + done, buf, (size_t)r); /* stands in for writev(socket) */
        done += r;
    }
    return done;
}
int64_t nzo_serve_slice_mem(const uint8_t* obj, int64_t off, int64_t len, uint8_t* dst) {
    uint8_t buf[NZO_OUTBUF];
    int64_t done = 0;
    while (done < len) {
        size_t want = len - done > NZO_OUTBUF ? NZO_OUTBUF : (size_t)(len - done);
        memcpy(buf, obj + off + done, want);
        memcpy(dst + done, buf, want);
        done += (int64_t)want;
    }
    return done;
}

/* ---- threaded CPU arm -------------------------------------------------- */
typedef struct work {
    const nzo_job* jobs;
    nzo_zjob* zjobs;
    int n;
    int next;
    int err;
    pthread_mutex_t mu;
} work;
static void* serve_worker(void* arg) {
    work* w = (work*)arg;
    for (;;) {
        pthread_mutex_lock(&w->mu);
        int i = w->next++;
        pthread_mutex_unlock(&w->mu);
        if (i >= w->n) break;
        if (w->jobs) {
            const nzo_job* j = &w->jobs[i];
            if (nzo_serve_slice(j->fd, j->off, j->len, j->dst) != j->len) w->err = 1;
        } else {
            nzo_zjob* j = &w->zjobs[i];
            j->out = nzo_gunzip(j->src, (size_t)j->n, j->dst, (size_t)j->cap);
            if (j->out < 0) w->err = 1;
        }
    }
    return NULL;
}
static double now_s(void) {
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}
static double run_workers(work* w, int threads) {
    if (threads < 1) threads = 1;
    if (threads > 256) threads = 256;
    pthread_t th[256];
    pthread_mutex_init(&w->mu, NULL);
    double t0 = now_s();
    if (threads == 1) {
        serve_worker(w);
    } else {
        for (int i = 0; i < threads; i++) pthread_create(&th[i], NULL, serve_worker, w);
        for (int i = 0; i < threads; i++) pthread_join(th[i], NULL);
    }
    double t1 = now_s();
    pthread_mutex_destroy(&w->mu);
    return w->err ? -1.0 : t1 - t0;
}
double nzo_serve_jobs(const nzo_job* jobs, int n, int threads) {
    work w = {jobs, NULL, n, 0, 0, PTHREAD_MUTEX_INITIALIZER};
    return run_workers(&w, threads);
}
double nzo_gunzip_jobs(nzo_zjob* jobs, int n, int threads) {
    work w = {NULL, jobs, n, 0, 0, PTHREAD_MUTEX_INITIALIZER};
    return run_workers(&w, threads);
}


/* ------------------------------------------------------------------------- */
/* SPEC H2: admission / validity of an upstream response (nginx.conf:57)      */
/* ------------------------------------------------------------------------- */
static int h2_name_is(const char* line, size_t nlen, const char* want) {
    return strlen(want) == nlen && strncasecmp(line, want, nlen) == 0;
}
static const char* h2_casestr(const char* s, size_t n, const char* needle) {
    size_t m = strlen(needle);
    for (size_t i = 0; m && i + m <= n; i++)
        if (strncasecmp(s + i, needle, m) == 0) return s + i;
    return NULL;
}
static int64_t h2_digits(const char* s, size_t n) {      /* all digits -> value, else -1 */
    if (n == 0) return -1;
    int64_t v = 0;
    for (size_t i = 0; i < n; i++) {
        if (!isdigit((unsigned char)s[i])) return -1;
        if (v > (INT64_MAX - 9) / 10) return -1;
        v = v * 10 + (s[i] - '0');
    }
    return v;
}
int nzo_admission(int status, const char* headers, int64_t now, int64_t default_valid_s, int64_t* valid_until) {
    int cacheable = 1, seen_xae = 0;
    int64_t valid = 0;
    if (valid_until) *valid_until = 0;
    if (status != 200) return 0;                                           /* H2a */
    const char* p = headers ? headers : "";
    while (*p) {
        const char* eol = strchr(p, '\n');
        size_t ll = eol ? (size_t)(eol - p) : strlen(p);
        const char* colon = memchr(p, ':', ll);
        if (colon) {
            size_t nlen = (size_t)(colon - p);
            const char* v = colon + 1;
            const char* ve = p + ll;
            while (v < ve && (*v == ' ' || *v == '\t')) v++;
            while (ve > v && (ve[-1] == ' ' || ve[-1] == '\t' || ve[-1] == '\r')) ve--;
            size_t vl = (size_t)(ve - v);
            if (h2_name_is(p, nlen, "Set-Cookie")) {                        /* H2b */
                cacheable = 0;
            } else if (h2_name_is(p, nlen, "Vary")) {
                if ((vl == 1 && v[0] == '*') || vl > 128) cacheable = 0;   /* NGX_HTTP_CACHE_VARY_LEN (128 since 1.11.10) */
            } else if (h2_name_is(p, nlen, "X-Accel-Expires")) {            /* H2c */
                seen_xae = 1;
                if (vl && v[0] == '@') {
                    int64_t n = h2_digits(v + 1, vl - 1);
                    if (n >= 0) valid = n;
                } else {
                    int64_t n = h2_digits(v, vl);
                    if (n == 0) cacheable = 0;
                    if (n > 0) valid = now + n;
                }
            } else if (h2_name_is(p, nlen, "Cache-Control")) {              /* H2d */
                if (!(valid != 0 && seen_xae)) {
                    if (h2_casestr(v, vl, "no-cache") || h2_casestr(v, vl, "no-store") || h2_casestr(v, vl, "private")) {
                        cacheable = 0;
                    } else {
                        const char* q = h2_casestr(v, vl, "s-maxage=");
                        size_t skip = 9;
                        if (!q) { q = h2_casestr(v, vl, "max-age="); skip = 8; }
                        if (q) {
                            int64_t n = 0;
                            int broke = 0;
                            for (q += skip; q < ve; q++) {
                                if (*q == ',' || *q == ';' || *q == ' ') break;
                                if (*q < '0' || *q > '9') { broke = 1; break; }
                                n = n * 10 + (*q - '0');
                                if (n > ((int64_t)1 << 40)) n = (int64_t)1 << 40;
                            }
                            if (broke || n == 0) cacheable = 0;
                            else valid = now + n;
                        }
                    }
                }
            } else if (h2_name_is(p, nlen, "Expires")) {                    /* H2e */
                if (valid == 0) {
                    int64_t t = nzo_parse_http_time(v, vl);
                    if (t < 0 || t < now) cacheable = 0;
                    else valid = t;
                }
            }
        }
        if (!eol) break;
        p = eol + 1;
    }
    if (cacheable && valid == 0) valid = now + default_valid_s;            /* H2f */
    if (valid_until) *valid_until = cacheable ? valid : 0;
    return cacheable;
}
