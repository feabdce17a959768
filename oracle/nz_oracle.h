/*
 * oracle/nz_oracle.h — CPU restatement of the reference's cache-hit path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under nezha_b200/ may include, link or
 * load this.  Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline
 * legs use it, as the checker / the timed CPU arm.
 *
 * PARITY UNPINNED: the reference holds no tests or golden vectors for this path
 * and its data plane (stock nginx) cannot run here; see oracle/SPEC.md.  Rule
 * numbers (K1, R4, ...) below refer to oracle/SPEC.md.
 */
#ifndef NZ_ORACLE_H
#define NZ_ORACLE_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* ---- identity (SPEC K5, K6) ------------------------------------------- */
void     nzo_md5(const void* data, size_t n, uint8_t out[16]);
uint32_t nzo_crc32_ieee(uint32_t crc, const void* data, size_t n);

/* ---- integrity (SPEC X1, X2) ------------------------------------------ */
uint32_t nzo_crc32c(uint32_t crc, const void* data, size_t n);
uint32_t nzo_crc32c_bitwise(uint32_t crc, const void* data, size_t n);
uint32_t nzo_crc32c_combine(uint32_t crc1, uint32_t crc2, uint64_t len2);
/* chunk table of ceil(n/chunk) entries; returns the whole-object CRC obtained
 * by combining the chunk CRCs (must equal nzo_crc32c(0,data,n)). */
uint32_t nzo_crc32c_chunks(const void* data, size_t n, size_t chunk, uint32_t* out);

/* ---- cache key (SPEC K1-K4) ------------------------------------------- */
/* host_hdr: value of Host: (may be NULL); request_target: the raw request-line
 * target ("/a/b?x" or absolute "http://h/p"); server_name, proxy_host: from the
 * selected vhost.  Writes a NUL-terminated key; returns its length, or -400. */
int nzo_cache_key(const char* host_hdr, const char* request_target,
                  const char* server_name, const char* proxy_host,
                  char* out, size_t cap);
/* vhost table of examples/demo/nginx.conf (SPEC V1): returns index 0..4 */
int nzo_vhost_select(const char* host_lower);
const char* nzo_vhost_name(int idx);
const char* nzo_vhost_proxy_host(int idx);
/* SPEC V2: 1 = passes the unanchored GET|HEAD regex gate, 0 = 405 */
int nzo_method_gate(const char* method);

/* ---- ranges (SPEC R3-R7) ---------------------------------------------- */
typedef struct nzo_range { int64_t start, end; /* [start,end) */ } nzo_range;
#define NZO_RANGE_OK        0
#define NZO_RANGE_DECLINED  1   /* ignore header, full 200 */
#define NZO_RANGE_416       416
/* spec: the Range value AFTER the 6 bytes "bytes=" */
int nzo_range_parse(const char* spec, int64_t content_length,
                    nzo_range* out, int max_out, int* n_out);

/* ---- HTTP dates (SPEC C7): seconds since epoch or -1 ------------------- */
int64_t nzo_parse_http_time(const char* s, size_t n);
/* SPEC H2: headers = "Name: value\n" lines in arrival order; returns 1 cacheable / 0; *valid_until absolute seconds */
int nzo_admission(int status, const char* headers, int64_t now, int64_t default_valid_s, int64_t* valid_until);

/* ---- request head (SPEC P) --------------------------------------------- */
/* raw: the bytes before the blank line.  Writes a text form: "error N" (0/400/505), and for error 0 "method M",
 * "target T", "minor N" (-1/0/1), "keep 0|1", "body 0|1" and one "f name: value" line per field K/C/R/G read
 * (host, range, if-range, if-none-match, if-match, if-modified-since, if-unmodified-since, accept-encoding), sorted by
 * name.  Returns the length written or -1 (cap too small). */
int nzo_parse_head(const char* raw, size_t n, char* out, size_t cap);

/* ---- whole request plan (SPEC V,C,R,G,E) ------------------------------- */
typedef struct nzo_request {
    const char* method;             /* "GET" */
    int         http_minor;         /* 0 or 1 (HTTP/1.x); -1 = HTTP/0.9 */
    const char* range;              /* header values or NULL */
    const char* if_range;
    const char* if_none_match;
    const char* if_match;
    const char* if_modified_since;
    const char* if_unmodified_since;
    const char* accept_encoding;
} nzo_request;

typedef struct nzo_object {         /* what the cache entry stores */
    int64_t     stored_length;      /* bytes at rest */
    int64_t     raw_length;         /* inflated length if content_encoding=gzip and known, else -1 */
    const char* content_type;       /* may be "" */
    const char* etag;               /* NULL if origin sent none */
    const char* last_modified;      /* NULL if none */
    const char* content_encoding;   /* NULL or "gzip" */
    int         gunzip_enabled;     /* operator switch for SPEC G */
    int         no_ranges;          /* SPEC R0: the upstream response carried no Accept-Ranges header */
} nzo_object;

/* Serialises the response plan as text into out (NUL-terminated):
 *   status N / head 0|1 / repr identity|inflated / hdr Name: value ... /
 *   lit <hex> | slice <start> <len> ... / end
 * boundary: the 20-digit multipart boundary counter to use.  Returns length
 * or -1 if cap is too small. */
int nzo_plan(const nzo_request* rq, const nzo_object* ob, uint64_t boundary,
             char* out, size_t cap);

/* ---- gunzip (SPEC G2, G4): zlib is nginx's own inflate ------------------ */
/* returns inflated length, or -1 declined (malformed / multi-member / trailing) */
int64_t nzo_gunzip(const void* src, size_t n, void* dst, size_t cap);
/* peeks ISIZE (last 4 bytes LE); -1 if n < 18 */
int64_t nzo_gzip_isize(const void* src, size_t n);

/* ---- the hot loop (SPEC B1): pread 32 KiB -> copy out ------------------- */
/* Copies [off, off+len) of fd into dst in 32 KiB units through a bounce buffer
 * (nginx: pread into an output buffer, then writev from it).  Returns bytes
 * copied or -1. */
int64_t nzo_serve_slice(int fd, int64_t off, int64_t len, uint8_t* dst);
/* Same from memory (page-cache stand-in when no file is wanted). */
int64_t nzo_serve_slice_mem(const uint8_t* obj, int64_t off, int64_t len, uint8_t* dst);

/* ---- CPU serving arm for bench.py (--impl reference / cpu_baseline) ----- */
typedef struct nzo_job { int fd; int64_t off, len; uint8_t* dst; } nzo_job;
/* Serves n jobs with `threads` worker threads (1 = nginx.conf:2); returns
 * seconds of wall time, negative on error. */
double nzo_serve_jobs(const nzo_job* jobs, int n, int threads);
/* Inflates n gzip members with `threads` workers; returns seconds. */
typedef struct nzo_zjob { const uint8_t* src; int64_t n; uint8_t* dst; int64_t cap; int64_t out; } nzo_zjob;
double nzo_gunzip_jobs(nzo_zjob* jobs, int n, int threads);

#ifdef __cplusplus
}
#endif
#endif
