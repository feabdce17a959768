// nezha_b200/host/nz_http.h — host-side HTTP semantics of the cache-hit path.
//
// This is PRODUCT code (the compiled stand-in for the Go handler north_star asks
// for; `go` is absent from the image): what nginx does around the body loop for
// the vhosts of /root/reference/examples/demo/nginx.conf — pick the server block,
// gate the method, build "$host/$proxy_host$uri", evaluate conditionals and Range,
// and lay the response out as literals (written by the host) and slices (gathered
// by the GPU through include/nezha_gpu.h).  Rule ids in comments (K1, R4, C6 ...)
// are the numbered rules of the behaviour being matched; tests diff this layer
// against the independent CPU restatement.
#pragma once
#include <cstdint>
#include <optional>
#include <string>
#include <string_view>
#include <utility>
#include <vector>

namespace nzh {

struct VHost {
    std::string server_name;   // nginx.conf:45 etc.
    std::string proxy_host;    // host[:port] of proxy_pass, port only when not the scheme default
    bool upstream_tls = false; // nginx.conf:113 (https://storage.googleapis.com)
};

struct Config {
    std::vector<VHost> vhosts;           // first entry is the implicit default server
    bool gunzip = false;                 // operator switch for inflate-on-hit (new behaviour)
    static Config reference_default();   // the five live server blocks of examples/demo/nginx.conf
    int select(std::string_view host_lower) const;
};

struct Request {
    std::string method;
    std::string target;                  // raw request-line target
    int http_minor = 1;                  // -1 = HTTP/0.9
    std::optional<std::string> host, range, if_range, if_none_match, if_match, if_modified_since,
        if_unmodified_since, accept_encoding;
};

struct ObjectMeta {
    uint64_t stored_length = 0;
    int64_t raw_length = -1;             // inflated length when stored gzip and known
    std::string content_type;
    std::optional<std::string> etag, last_modified;
    bool stored_gzip = false;
    bool allow_ranges = true;            // R0: the upstream response carried an Accept-Ranges header (no proxy_force_ranges in the conf)
};

struct Part {
    bool literal = false;
    std::string bytes;                   // literal
    uint64_t off = 0, len = 0;           // slice of the served representation
};

struct Plan {
    int status = 0;
    bool head = false, inflated = false, bypass = false;
    std::vector<std::pair<std::string, std::string>> headers;
    std::vector<Part> body;
    uint64_t body_length() const;
    std::string to_text() const;         // canonical text form used by the parity tests
};

// MD5 (RFC 1321): cache identity of the key string (nginx.conf:38)
void md5(const void* data, size_t n, uint8_t out[16]);

// "$host/$proxy_host$uri" (nginx.conf:56); nullopt = 400 Bad Request.  vhost_out gets the block index.
std::optional<std::string> cache_key(const Config& cfg, const std::optional<std::string>& host_header,
                                     std::string_view target, int* vhost_out);

bool method_allowed(std::string_view method);   // nginx.conf:47-49 (unanchored GET|HEAD regex)

struct ByteRange { int64_t start, end; };       // [start, end)
enum class RangeVerdict { Ok, Ignore, Unsatisfiable };
RangeVerdict parse_ranges(std::string_view spec_after_bytes_eq, int64_t length, std::vector<ByteRange>& out,
                          size_t max_ranges = 64);

int64_t parse_http_date(std::string_view s);    // seconds since the epoch, -1 if unparseable

// Cache admission and validity of an upstream response (SPEC H2a-H2e: nginx's proxy_cache_valid 200 1d, nginx.conf:57,
// overridden by X-Accel-Expires / Cache-Control / Expires exactly as nginx's upstream header handlers do, in
// arrival order).  headers: (name, value) as received.  valid_until: absolute seconds.
struct Admission { bool cacheable = false; int64_t valid_until = 0; };
Admission admission(int status, const std::vector<std::pair<std::string, std::string>>& headers, int64_t now,
                    int64_t default_valid_s = 86400);

// The request head (SPEC P: nginx's request-line and header-line parsers and the ngx_http_headers_in handlers): what reaches
// the vhost / key / planner above.  error != 0 (400, 505) is answered with the stock page before anything else runs.
struct Head {
    int error = 0;
    Request rq;
    bool keep_alive = true;
    bool has_body = false;               // Content-Length > 0 or a Transfer-Encoding line: answered, then the connection closes
    std::string request_line, referer, user_agent, x_forwarded_for;   // access log only (nginx.conf:16-18)
    std::string to_text() const;         // canonical text form used by the parity tests
};
Head parse_head(std::string_view raw);   // raw = the bytes before the blank line

Plan plan_response(const Config& cfg, const Request& rq, const ObjectMeta& ob, uint64_t boundary);
Plan error_page(int code, bool head);           // stock nginx page (server_tokens off, nginx.conf:20)

}  // namespace nzh

// C view for the ctypes parity tests and for the proxy binary
extern "C" {
int nzh_cache_key(const char* host_hdr, const char* target, char* out, size_t cap);   // length or -400
int nzh_method_allowed(const char* method);
int nzh_parse_ranges(const char* spec, int64_t length, int64_t* out_pairs, int max_pairs, int* n_out);   // 0 ok,1 ignore,416
int64_t nzh_parse_http_date(const char* s, size_t n);
void nzh_md5(const void* data, size_t n, uint8_t out[16]);
struct nzh_request {
    const char* method; int http_minor;
    const char *range, *if_range, *if_none_match, *if_match, *if_modified_since, *if_unmodified_since, *accept_encoding;
};
struct nzh_object {
    int64_t stored_length, raw_length;
    const char *content_type, *etag, *last_modified, *content_encoding;
    int gunzip_enabled;
    int no_ranges;
};
int nzh_plan_text(const nzh_request* rq, const nzh_object* ob, uint64_t boundary, char* out, size_t cap);
// headers: "Name: value\n" lines in arrival order.  Returns 1 cacheable / 0 not; *valid_until = absolute seconds.
int nzh_admission(int status, const char* headers, int64_t now, int64_t default_valid_s, int64_t* valid_until);
int nzh_parse_head_text(const char* raw, size_t n, char* out, size_t cap);   // Head::to_text(); length or -1
}
