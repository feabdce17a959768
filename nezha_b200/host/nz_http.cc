// nezha_b200/host/nz_http.cc — see nz_http.h.  Product host code: no CUDA, no oracle.
#include "nz_http.h"

#include <algorithm>
#include <array>
#include <cstdio>
#include <cstring>
#include <limits>

namespace nzh {

// ---------------------------------------------------------------------------
// configuration: the server blocks of /root/reference/examples/demo/nginx.conf:43-136
// ---------------------------------------------------------------------------
Config Config::reference_default() {
    Config c;
    c.vhosts = {
        {"download.tensorflow.org", "download.tensorflow.org", false},   // nginx.conf:43-59
        {"yann.lecun.com", "yann.lecun.com", false},                     // :62-78
        {"www.cs.toronto.edu", "www.cs.toronto.edu", false},             // :81-97
        {"storage.googleapis.com", "storage.googleapis.com", true},      // :101-117 (https upstream, default port)
        {"images.cocodataset.org", "images.cocodataset.org", false},     // :120-136
    };
    return c;
}

int Config::select(std::string_view host_lower) const {
    for (size_t i = 0; i < vhosts.size(); ++i)
        if (vhosts[i].server_name == host_lower) return (int)i;
    return 0;   // no default_server flag anywhere: the first block catches unknown hosts
}

// ---------------------------------------------------------------------------
// MD5 (RFC 1321)
// ---------------------------------------------------------------------------
namespace {
struct Md5 {
    uint32_t a0 = 0x67452301u, b0 = 0xefcdab89u, c0 = 0x98badcfeu, d0 = 0x10325476u;
    static constexpr std::array<uint32_t, 64> K = {
        0xd76aa478, 0xe8c7b756, 0x242070db, 0xc1bdceee, 0xf57c0faf, 0x4787c62a, 0xa8304613, 0xfd469501, 0x698098d8,
        0x8b44f7af, 0xffff5bb1, 0x895cd7be, 0x6b901122, 0xfd987193, 0xa679438e, 0x49b40821, 0xf61e2562, 0xc040b340,
        0x265e5a51, 0xe9b6c7aa, 0xd62f105d, 0x02441453, 0xd8a1e681, 0xe7d3fbc8, 0x21e1cde6, 0xc33707d6, 0xf4d50d87,
        0x455a14ed, 0xa9e3e905, 0xfcefa3f8, 0x676f02d9, 0x8d2a4c8a, 0xfffa3942, 0x8771f681, 0x6d9d6122, 0xfde5380c,
        0xa4beea44, 0x4bdecfa9, 0xf6bb4b60, 0xbebfbc70, 0x289b7ec6, 0xeaa127fa, 0xd4ef3085, 0x04881d05, 0xd9d4d039,
        0xe6db99e5, 0x1fa27cf8, 0xc4ac5665, 0xf4292244, 0x432aff97, 0xab9423a7, 0xfc93a039, 0x655b59c3, 0x8f0ccc92,
        0xffeff47d, 0x85845dd1, 0x6fa87e4f, 0xfe2ce6e0, 0xa3014314, 0x4e0811a1, 0xf7537e82, 0xbd3af235, 0x2ad7d2bb,
        0xeb86d391};
    static uint32_t rotl(uint32_t v, unsigned s) { return (v << s) | (v >> (32 - s)); }
    void block(const uint8_t* p) {
        static constexpr unsigned S[4][4] = {{7, 12, 17, 22}, {5, 9, 14, 20}, {4, 11, 16, 23}, {6, 10, 15, 21}};
        uint32_t w[16];
        for (int i = 0; i < 16; ++i) std::memcpy(&w[i], p + 4 * i, 4);   // little-endian host
        uint32_t A = a0, B = b0, C = c0, D = d0;
        for (unsigned i = 0; i < 64; ++i) {
            uint32_t F;
            unsigned g;
            switch (i >> 4) {
            case 0: F = D ^ (B & (C ^ D)); g = i; break;
            case 1: F = C ^ (D & (B ^ C)); g = (5 * i + 1) & 15; break;
            case 2: F = B ^ C ^ D; g = (3 * i + 5) & 15; break;
            default: F = C ^ (B | ~D); g = (7 * i) & 15; break;
            }
            const uint32_t t = A + F + K[i] + w[g];
            A = D; D = C; C = B;
            B += rotl(t, S[i >> 4][i & 3]);
        }
        a0 += A; b0 += B; c0 += C; d0 += D;
    }
};
}  // namespace

void md5(const void* data, size_t n, uint8_t out[16]) {
    Md5 st;
    const uint8_t* p = static_cast<const uint8_t*>(data);
    size_t i = 0;
    for (; i + 64 <= n; i += 64) st.block(p + i);
    uint8_t pad[128] = {0};
    const size_t rem = n - i;
    if (rem) std::memcpy(pad, p + i, rem);
    pad[rem] = 0x80;
    const size_t total = rem + 1 + 8 <= 64 ? 64 : 128;
    const uint64_t bits = (uint64_t)n << 3;
    std::memcpy(pad + total - 8, &bits, 8);
    st.block(pad);
    if (total == 128) st.block(pad + 64);
    std::memcpy(out, &st.a0, 4); std::memcpy(out + 4, &st.b0, 4);
    std::memcpy(out + 8, &st.c0, 4); std::memcpy(out + 12, &st.d0, 4);
}

// ---------------------------------------------------------------------------
// key
// ---------------------------------------------------------------------------
namespace {
bool iprefix(std::string_view s, std::string_view pre) {
    if (s.size() < pre.size()) return false;
    for (size_t i = 0; i < pre.size(); ++i)
        if (std::tolower((unsigned char)s[i]) != pre[i]) return false;
    return true;
}

// $host (K2): the name ends at the first ':' outside an IPv6 literal, or right after the literal's ']'; upper case
// is folded; one trailing dot goes; a NUL, a '/' or ".." anywhere in the value (port part included) is a 400
std::optional<std::string> normal_host(std::string_view h) {
    enum { Name, Literal, Rest } st = Name;
    size_t end = h.size(), last_dot = std::string_view::npos;
    for (size_t i = 0; i < h.size(); ++i) {
        switch (h[i]) {
        case '.':
            if (last_dot != std::string_view::npos && last_dot + 1 == i) return std::nullopt;
            last_dot = i;
            break;
        case ':': if (st == Name) { end = i; st = Rest; } break;
        case '[': if (i == 0) st = Literal; break;
        case ']': if (st == Literal) { end = i + 1; st = Rest; } break;
        case '\0': case '/': return std::nullopt;
        default: break;
        }
    }
    if (end && last_dot != std::string_view::npos && last_dot + 1 == end) --end;
    if (end == 0 || end > 500) return std::nullopt;
    std::string out(h.substr(0, end));
    for (char& ch : out)
        if (ch >= 'A' && ch <= 'Z') ch = (char)(ch | 0x20);
    return out;
}

int hexdigit(char c) {
    if (c >= '0' && c <= '9') return c - '0';
    c = (char)(c | 0x20);
    return (c >= 'a' && c <= 'f') ? c - 'a' + 10 : -1;
}

// $uri: decoded, dot segments resolved, slashes merged, query/fragment dropped (K4)
std::optional<std::string> normal_uri(std::string_view target) {
    if (target.empty() || target.front() != '/') return std::nullopt;
    std::string dec;
    dec.reserve(target.size());
    for (size_t i = 0; i < target.size();) {
        const char c = target[i];
        if (c == '?' || c == '#') break;
        if (c != '%') { dec.push_back(c); ++i; continue; }
        if (i + 2 >= target.size()) return std::nullopt;
        const int hi = hexdigit(target[i + 1]), lo = hexdigit(target[i + 2]);
        if (hi < 0 || lo < 0 || (hi == 0 && lo == 0)) return std::nullopt;
        dec.push_back((char)(hi * 16 + lo));
        i += 3;
    }
    if (dec.size() > 8000) return std::nullopt;
    std::vector<std::string_view> keep;
    bool trailing = false;
    size_t pos = 1;
    const std::string_view d(dec);
    for (;;) {
        const size_t nx = d.find('/', pos);
        const bool last = nx == std::string_view::npos;
        const std::string_view seg = d.substr(pos, last ? std::string_view::npos : nx - pos);
        if (seg.empty()) {
            if (last) trailing = true;
        } else if (seg == ".") {
            if (last) trailing = true;
        } else if (seg == "..") {
            if (keep.empty()) return std::nullopt;   // escapes the root
            keep.pop_back();
            if (last) trailing = true;
        } else {
            keep.push_back(seg);
            trailing = false;
        }
        if (last) break;
        pos = nx + 1;
    }
    std::string out;
    for (auto s : keep) { out.push_back('/'); out.append(s); }
    if (out.empty() || trailing) out.push_back('/');
    return out;
}
}  // namespace

std::optional<std::string> cache_key(const Config& cfg, const std::optional<std::string>& host_header,
                                     std::string_view target, int* vhost_out) {
    std::optional<std::string> host;
    std::string_view path = target;
    if (iprefix(target, "http://") || iprefix(target, "https://")) {   // absolute-form: request line wins
        const size_t hs = target.find("://") + 3;
        size_t he = hs;
        while (he < target.size() && target[he] != '/' && target[he] != '?' && target[he] != '#') ++he;
        host = normal_host(target.substr(hs, he - hs));
        if (!host) return std::nullopt;
        path = (he < target.size() && target[he] == '/') ? target.substr(he) : std::string_view("/");
    } else if (host_header) {                          // present but empty: invalid (400), not "absent"
        host = normal_host(*host_header);
        if (!host) return std::nullopt;
    }
    const int vi = cfg.select(host ? std::string_view(*host) : std::string_view());
    if (vhost_out) *vhost_out = vi;
    if (!host) host = cfg.vhosts[vi].server_name;
    auto uri = normal_uri(path);
    if (!uri) return std::nullopt;
    std::string key = *host;
    key.push_back('/');
    key += cfg.vhosts[vi].proxy_host;
    key += *uri;
    if (key.size() >= 16000) return std::nullopt;
    return key;
}

bool method_allowed(std::string_view m) {   // V2
    return m.find("GET") != std::string_view::npos || m.find("HEAD") != std::string_view::npos;
}

// ---------------------------------------------------------------------------
// Range (R3-R7)
// ---------------------------------------------------------------------------
namespace {
struct Cursor {
    std::string_view s;
    size_t i = 0;
    char peek() const { return i < s.size() ? s[i] : '\0'; }
    void blanks() { while (peek() == ' ') ++i; }
    // decimal int64; false on overflow.  Requires a digit at the cursor.
    bool number(int64_t& v) {
        constexpr int64_t lim = std::numeric_limits<int64_t>::max();
        v = 0;
        while (peek() >= '0' && peek() <= '9') {
            const int d = peek() - '0';
            if (v > (lim - d) / 10) return false;
            v = v * 10 + d;
            ++i;
        }
        return true;
    }
    bool digit() const { return peek() >= '0' && peek() <= '9'; }
};
}  // namespace

RangeVerdict parse_ranges(std::string_view spec, int64_t length, std::vector<ByteRange>& out, size_t max_ranges) {
    // an embedded NUL ends the header value, as a C string would
    if (const size_t z = spec.find('\0'); z != std::string_view::npos) spec = spec.substr(0, z);
    out.clear();
    Cursor c{spec};
    int64_t total = 0;
    for (;;) {
        int64_t first = 0, last = 0;
        c.blanks();
        if (c.peek() == '-') {                      // "-suffix"
            ++c.i;
            if (!c.digit()) return RangeVerdict::Unsatisfiable;
            if (!c.number(last)) return RangeVerdict::Unsatisfiable;
            c.blanks();
            if (c.peek() != ',' && c.peek() != '\0') return RangeVerdict::Unsatisfiable;
            first = last < length ? length - last : 0;
            last = length;                          // exclusive
        } else {
            if (!c.digit()) return RangeVerdict::Unsatisfiable;
            if (!c.number(first)) return RangeVerdict::Unsatisfiable;
            c.blanks();
            if (c.peek() != '-') return RangeVerdict::Unsatisfiable;
            ++c.i;
            c.blanks();
            if (c.peek() == ',' || c.peek() == '\0') {
                last = length;                      // "first-"
            } else {
                if (!c.digit()) return RangeVerdict::Unsatisfiable;
                if (!c.number(last)) return RangeVerdict::Unsatisfiable;
                c.blanks();
                if (c.peek() != ',' && c.peek() != '\0') return RangeVerdict::Unsatisfiable;
                last = last >= length ? length : last + 1;
            }
        }
        if (first < last) {
            if (out.size() >= max_ranges) return RangeVerdict::Ignore;
            out.push_back({first, last});
            if (total > std::numeric_limits<int64_t>::max() - (last - first)) return RangeVerdict::Unsatisfiable;
            total += last - first;
        } else if (first == 0) {
            return RangeVerdict::Ignore;            // R5
        }
        if (c.peek() != ',') break;
        ++c.i;
    }
    if (out.empty()) return RangeVerdict::Unsatisfiable;
    if (total > length) return RangeVerdict::Ignore;   // R7
    return RangeVerdict::Ok;
}

// ---------------------------------------------------------------------------
// HTTP dates (C7): RFC 822/1123, RFC 850, asctime
// ---------------------------------------------------------------------------
namespace {
int two_digits(std::string_view s, size_t at) {
    if (at + 2 > s.size()) return -1;
    const char a = s[at], b = s[at + 1];
    if (a < '0' || a > '9' || b < '0' || b > '9') return -1;
    return (a - '0') * 10 + (b - '0');
}
int month_number(std::string_view s, size_t at) {
    static const char* names[12] = {"Jan", "Feb", "Mar", "Apr", "May", "Jun", "Jul", "Aug", "Sep", "Oct", "Nov", "Dec"};
    if (at + 3 > s.size()) return 0;
    for (int m = 0; m < 12; ++m)
        if (s.compare(at, 3, names[m]) == 0) return m + 1;
    return 0;
}
int64_t civil_to_days(int64_t y, int m, int d) {   // proleptic Gregorian, days since 1970-01-01
    y -= m <= 2;
    const int64_t era = (y >= 0 ? y : y - 399) / 400;
    const int64_t yoe = y - era * 400;
    const int64_t doy = (153 * (m + (m > 2 ? -3 : 9)) + 2) / 5 + d - 1;
    const int64_t doe = yoe * 365 + yoe / 4 - yoe / 100 + doy;
    return era * 146097 + doe - 719468;
}
}  // namespace

int64_t parse_http_date(std::string_view s) {
    enum { Rfc822, Rfc850, Asctime } form = Rfc822;
    size_t i = 0;
    for (; i < s.size(); ++i) {
        if (s[i] == ',') break;
        if (s[i] == ' ') { form = Asctime; break; }
    }
    for (++i; i < s.size() && s[i] == ' '; ++i) {}
    if (i > s.size() || s.size() - i < 18) return -1;
    int day = 0, mon = 0, year = 0;
    if (form != Asctime) {
        day = two_digits(s, i);
        if (day < 0) return -1;
        i += 2;
        if (s[i] == ' ') {
            if (s.size() - i < 18) return -1;
            form = Rfc822;
        } else if (s[i] == '-') {
            form = Rfc850;
        } else {
            return -1;
        }
        ++i;
    }
    mon = month_number(s, i);
    if (!mon) return -1;
    i += 3;
    if (i >= s.size()) return -1;
    if ((form == Rfc822 && s[i] != ' ') || (form == Rfc850 && s[i] != '-')) return -1;
    ++i;
    if (form == Rfc822) {
        if (i + 4 > s.size()) return -1;
        const int hi = two_digits(s, i), lo = two_digits(s, i + 2);
        if (hi < 0 || lo < 0) return -1;
        year = hi * 100 + lo;
        i += 4;
    } else if (form == Rfc850) {
        if (i + 2 > s.size()) return -1;
        year = two_digits(s, i);
        if (year < 0) return -1;
        year += year < 70 ? 2000 : 1900;
        i += 2;
    } else {
        if (i < s.size() && s[i] == ' ') ++i;
        if (i >= s.size() || s[i] < '0' || s[i] > '9') return -1;
        day = s[i++] - '0';
        if (i < s.size() && s[i] != ' ') {
            if (s[i] < '0' || s[i] > '9') return -1;
            day = day * 10 + (s[i++] - '0');
        }
        if (s.size() - i < 14) return -1;
    }
    if (i >= s.size() || s[i] != ' ') return -1;
    ++i;
    if (i + 8 > s.size()) return -1;
    const int hh = two_digits(s, i), mm = two_digits(s, i + 3), ss = two_digits(s, i + 6);
    if (hh < 0 || mm < 0 || ss < 0 || s[i + 2] != ':' || s[i + 5] != ':') return -1;
    i += 8;
    if (form == Asctime) {
        if (i >= s.size() || s[i] != ' ') return -1;
        ++i;
        if (i + 4 > s.size()) return -1;
        const int hi = two_digits(s, i), lo = two_digits(s, i + 2);
        if (hi < 0 || lo < 0) return -1;
        year = hi * 100 + lo;
    }
    if (hh > 23 || mm > 59 || ss > 59) return -1;
    if (day == 29 && mon == 2) {
        if ((year & 3) || (year % 100 == 0 && year % 400 != 0)) return -1;
    } else if (day > 30 + ((mon ^ (mon >> 3)) & 1) || (mon == 2 && day > 28)) {      // C7: no lower bound on the day (nginx has none)
        return -1;
    }
    return civil_to_days(year, mon, day) * 86400 + hh * 3600 + mm * 60 + ss;
}

// ---------------------------------------------------------------------------
// entity tags (C6)
// ---------------------------------------------------------------------------
namespace {
bool etag_in_list(std::string_view list, const std::optional<std::string>& etag, bool weak) {
    if (list == "*") return true;
    if (!etag) return false;
    std::string_view want(*etag);
    if (weak && want.size() > 2 && want[0] == 'W' && want[1] == '/') want.remove_prefix(2);
    size_t i = 0;
    const size_t n = list.size();
    while (i < n) {
        if (weak && n - i > 2 && list[i] == 'W' && list[i + 1] == '/') i += 2;
        if (want.size() > n - i) return false;
        if (list.compare(i, want.size(), want) == 0) {
            size_t j = i + want.size();
            while (j < n && (list[j] == ' ' || list[j] == '\t')) ++j;
            if (j == n || list[j] == ',') return true;
        }
        while (i < n && list[i] != ',') ++i;
        while (i < n && (list[i] == ' ' || list[i] == '\t' || list[i] == ',')) ++i;
    }
    return false;
}

const char* reason_phrase(int code) {
    switch (code) {
    case 400: return "Bad Request";
    case 404: return "Not Found";
    case 405: return "Not Allowed";
    case 412: return "Precondition Failed";
    case 416: return "Requested Range Not Satisfiable";
    case 502: return "Bad Gateway";
    case 504: return "Gateway Time-out";
    case 505: return "HTTP Version Not Supported";
    default: return "";
    }
}

Plan stock_page(int code, bool head, std::optional<std::pair<std::string, std::string>> extra = std::nullopt) {   // E1
    const std::string title = std::to_string(code) + " " + reason_phrase(code);
    std::string page = "<html>\r\n<head><title>" + title + "</title></head>\r\n<body>\r\n<center><h1>" + title +
                       "</h1></center>\r\n<hr><center>nginx</center>\r\n</body>\r\n</html>\r\n";
    Plan p;
    p.status = code;
    p.head = head;
    p.headers.emplace_back("Content-Type", "text/html");
    p.headers.emplace_back("Content-Length", std::to_string(page.size()));
    if (extra) p.headers.push_back(*extra);
    Part lit;
    lit.literal = true;
    lit.bytes = std::move(page);
    p.body.push_back(std::move(lit));
    return p;
}

Part slice(uint64_t off, uint64_t len) {
    Part s;
    s.off = off;
    s.len = len;
    return s;
}
Part literal(std::string b) {
    Part s;
    s.literal = true;
    s.bytes = std::move(b);
    return s;
}
}  // namespace

Plan error_page(int code, bool head) { return stock_page(code, head); }

// ---- request head (SPEC P) ----------------------------------------------------------------------
namespace {
bool ieq(std::string_view a, std::string_view b) {
    if (a.size() != b.size()) return false;
    for (size_t i = 0; i < a.size(); ++i)
        if (std::tolower((unsigned char)a[i]) != std::tolower((unsigned char)b[i])) return false;
    return true;
}
bool has_ci(std::string_view hay, std::string_view needle) {
    if (needle.size() > hay.size()) return false;
    for (size_t i = 0; i + needle.size() <= hay.size(); ++i)
        if (ieq(hay.substr(i, needle.size()), needle)) return true;
    return false;
}
std::string_view trim_spaces(std::string_view v) {       // P5: spaces only
    while (!v.empty() && v.front() == ' ') v.remove_prefix(1);
    while (!v.empty() && v.back() == ' ') v.remove_suffix(1);
    return v;
}
}  // namespace

Head parse_head(std::string_view raw) {
    Head bad;
    bad.error = 400;
    bad.keep_alive = false;
    while (!raw.empty() && (raw.front() == '\r' || raw.front() == '\n')) raw.remove_prefix(1);
    while (raw.size() >= 2 && raw.substr(raw.size() - 2) == "\r\n") raw.remove_suffix(2);
    for (size_t i = 0; i < raw.size(); ++i) {                                  // P2: NUL, bare CR, bare LF
        const char c = raw[i];
        if (c == '\0') return bad;
        if (c == '\r' && (i + 1 >= raw.size() || raw[i + 1] != '\n')) return bad;
        if (c == '\n' && (i == 0 || raw[i - 1] != '\r')) return bad;
    }
    std::vector<std::string_view> lines;
    for (size_t pos = 0;;) {
        const size_t e = raw.find("\r\n", pos);
        lines.push_back(raw.substr(pos, e == std::string_view::npos ? std::string_view::npos : e - pos));
        if (e == std::string_view::npos) break;
        pos = e + 2;
    }
    Head h;
    const std::string_view rl = lines[0];
    h.request_line = std::string(rl);
    size_t j = 0;
    while (j < rl.size() && ((rl[j] >= 'A' && rl[j] <= 'Z') || rl[j] == '_')) ++j;
    if (j == 0 || j >= rl.size() || rl[j] != ' ') return bad;                  // P1
    h.rq.method = std::string(rl.substr(0, j));
    while (j < rl.size() && rl[j] == ' ') ++j;
    size_t k = rl.find(' ', j);
    if (k == std::string_view::npos) k = rl.size();
    const std::string_view target = rl.substr(j, k - j);
    if (target.empty()) return bad;
    if (target[0] != '/') {                                                    // P2: absolute-form
        size_t m = 0;
        while (m < target.size() && (unsigned char)target[m] < 128 && std::isalpha((unsigned char)target[m])) ++m;
        if (m == 0 || target.substr(m, 3) != "://") return bad;
    }
    h.rq.target = std::string(target);
    const std::string_view ver = trim_spaces(rl.substr(k));
    if (ver.empty()) {                                                         // P3
        if (h.rq.method != "GET") return bad;
        h.rq.http_minor = -1;
        h.keep_alive = false;
    } else {                                                                   // P4
        if (ver.substr(0, 5) != "HTTP/") return bad;
        const std::string_view num = ver.substr(5);
        const size_t dot = num.find('.');
        if (dot == std::string_view::npos) return bad;
        const std::string_view mj = num.substr(0, dot), mn = num.substr(dot + 1);
        auto digits = [](std::string_view d) {
            if (d.empty() || d.size() > 2) return false;
            for (char c : d) if (c < '0' || c > '9') return false;
            return true;
        };
        if (!digits(mj) || !digits(mn) || mj[0] == '0') return bad;
        if (std::stoi(std::string(mj)) > 1) { Head v; v.error = 505; v.keep_alive = false; return v; }
        h.rq.http_minor = std::stoi(std::string(mn)) == 0 ? 0 : 1;
        h.keep_alive = h.rq.http_minor == 1;
    }
    static const char* const unique[] = {"if-modified-since", "if-unmodified-since", "if-match", "if-none-match", "if-range", "content-length",
                                         "content-range", "expect", "authorization"};
    bool seen[sizeof unique / sizeof unique[0]] = {};
    int hosts = 0;
    for (size_t li = 1; li < lines.size(); ++li) {
        const std::string_view ln = lines[li];
        if (ln.empty()) break;
        const size_t c = ln.find(':');
        const std::string_view name = ln.substr(0, c);                         // no colon: the whole line is a name (P5)
        const std::string_view val = c == std::string_view::npos ? std::string_view() : trim_spaces(ln.substr(c + 1));
        bool valid = !name.empty();
        for (char ch : name)
            if (!((unsigned char)ch < 128 && (std::isalnum((unsigned char)ch) || ch == '-'))) { valid = false; break; }
        if (!valid) continue;                                                  // ignore_invalid_headers on
        if (ieq(name, "host")) {                                               // P8
            if (!normal_host(val)) return bad;
            if (++hosts == 1) h.rq.host = std::string(val);
            continue;
        }
        if (ieq(name, "range")) { if (!h.rq.range) h.rq.range = std::string(val); continue; }                          // P6
        if (ieq(name, "accept-encoding")) { if (!h.rq.accept_encoding) h.rq.accept_encoding = std::string(val); continue; }
        if (ieq(name, "referer")) { if (h.referer.empty()) h.referer = std::string(val); continue; }
        if (ieq(name, "user-agent")) { if (h.user_agent.empty()) h.user_agent = std::string(val); continue; }
        if (ieq(name, "x-forwarded-for")) { h.x_forwarded_for += (h.x_forwarded_for.empty() ? "" : ", ") + std::string(val); continue; }
        int u = -1;
        for (size_t i = 0; i < sizeof unique / sizeof unique[0]; ++i) if (ieq(name, unique[i])) u = (int)i;
        if (u >= 0) {                                                          // P7
            if (seen[u]) return bad;
            seen[u] = true;
            switch (u) {
            case 0: h.rq.if_modified_since = std::string(val); break;
            case 1: h.rq.if_unmodified_since = std::string(val); break;
            case 2: h.rq.if_match = std::string(val); break;
            case 3: h.rq.if_none_match = std::string(val); break;
            case 4: h.rq.if_range = std::string(val); break;
            case 5: {                                                          // P10
                if (val.empty()) return bad;
                bool nonzero = false;
                for (char ch : val) { if (ch < '0' || ch > '9') return bad; if (ch != '0') nonzero = true; }
                if (nonzero) h.has_body = true;
                break;
            }
            default: break;
            }
            continue;
        }
        if (ieq(name, "transfer-encoding")) { h.has_body = true; continue; }
        if (ieq(name, "connection")) {                                         // P9
            if (has_ci(val, "close")) h.keep_alive = false;
            else if (has_ci(val, "keep-alive")) h.keep_alive = true;
        }
    }
    if (h.rq.http_minor == 1 && hosts == 0) return bad;                        // P8
    return h;
}

std::string Head::to_text() const {
    std::string t = "error " + std::to_string(error) + "\n";
    if (error) return t;
    t += "method " + rq.method + "\ntarget " + rq.target + "\nminor " + std::to_string(rq.http_minor) + "\nkeep " + (keep_alive ? "1" : "0") +
         "\nbody " + (has_body ? "1" : "0") + "\n";
    auto f = [&](const char* n, const std::optional<std::string>& v) { if (v) t += std::string("f ") + n + ": " + *v + "\n"; };
    f("accept-encoding", rq.accept_encoding); f("host", rq.host); f("if-match", rq.if_match); f("if-modified-since", rq.if_modified_since);
    f("if-none-match", rq.if_none_match); f("if-range", rq.if_range); f("if-unmodified-since", rq.if_unmodified_since); f("range", rq.range);
    return t;
}

// ---- admission / validity (SPEC H2) -------------------------------------------------------------
namespace {
size_t ifind(std::string_view hay, std::string_view needle) {      // case-insensitive substring (ngx_strlcasestrn)
    if (needle.empty() || hay.size() < needle.size()) return std::string_view::npos;
    for (size_t i = 0; i + needle.size() <= hay.size(); ++i)
        if (ieq(hay.substr(i, needle.size()), needle)) return i;
    return std::string_view::npos;
}
int64_t atoi_strict(std::string_view s) {                          // ngx_atoi: digits only; -1 on anything else or overflow
    if (s.empty()) return -1;
    int64_t v = 0;
    for (char c : s) {
        if (c < '0' || c > '9') return -1;
        if (v > (INT64_MAX - (c - '0')) / 10) return -1;
        v = v * 10 + (c - '0');
    }
    return v;
}
}  // namespace

Admission admission(int status, const std::vector<std::pair<std::string, std::string>>& headers, int64_t now, int64_t default_valid_s) {
    Admission a;
    if (status != 200) return a;                                   // H2a: proxy_cache_valid names 200 only
    bool cacheable = true, xae_seen = false;
    int64_t valid = 0;
    for (const auto& h : headers) {
        const std::string_view name = h.first, val = h.second;
        if (ieq(name, "set-cookie")) cacheable = false;            // H2b
        else if (ieq(name, "vary")) { if (val == "*" || val.size() > 128) cacheable = false; }   // "*" or longer than NGX_HTTP_CACHE_VARY_LEN
        else if (ieq(name, "x-accel-expires")) {                   // H2c
            xae_seen = true;
            if (!val.empty() && val[0] == '@') {
                const int64_t n = atoi_strict(val.substr(1));
                if (n >= 0) valid = n;
            } else {
                const int64_t n = atoi_strict(val);
                if (n == 0) cacheable = false;
                else if (n > 0) valid = now + n;
            }
        } else if (ieq(name, "cache-control")) {                   // H2d
            if (valid != 0 && xae_seen) continue;
            if (ifind(val, "no-cache") != std::string_view::npos || ifind(val, "no-store") != std::string_view::npos ||
                ifind(val, "private") != std::string_view::npos) { cacheable = false; continue; }
            size_t p = ifind(val, "s-maxage="), skip = 9;
            if (p == std::string_view::npos) { p = ifind(val, "max-age="); skip = 8; }
            if (p == std::string_view::npos) continue;
            int64_t n = 0;
            bool bad = false;
            for (size_t i = p + skip; i < val.size(); ++i) {
                const char c = val[i];
                if (c == ',' || c == ';' || c == ' ') break;
                if (c >= '0' && c <= '9') { n = n * 10 + (c - '0'); if (n > (int64_t)1 << 40) n = (int64_t)1 << 40; continue; }
                bad = true;
                break;
            }
            if (bad || n == 0) { cacheable = false; continue; }
            valid = now + n;
        } else if (ieq(name, "expires")) {                         // H2e
            if (valid != 0) continue;
            const int64_t t = parse_http_date(val);
            if (t < 0 || t < now) { cacheable = false; continue; }
            valid = t;
        }
    }
    if (cacheable && valid == 0) valid = now + default_valid_s;    // H2f: proxy_cache_valid 200 1d
    a.cacheable = cacheable;
    a.valid_until = cacheable ? valid : 0;
    return a;
}

// G1: does the client take gzip?  The Accept-Encoding test nginx's gunzip filter relies on (ngx_http_gzip_ok):
// the literal prefix "gzip,", else a case-insensitive "gzip" that starts the value or follows ',' / ' ' and is
// followed by blanks then ',' / end, or by ";q=" with a value that is not zero.
namespace {
bool qvalue_nonzero(std::string_view v) {            // "0" | "1" [ "." up to 3 digits ], ends at ',' / ' ' / end
    if (v.empty() || (v[0] != '0' && v[0] != '1')) return false;
    unsigned q = (unsigned)(v[0] - '0') * 100, digits = 0;
    if (v.size() == 1 || v[1] == ',' || v[1] == ' ') return q != 0;
    if (v[1] != '.') return false;
    for (size_t i = 2; i < v.size() && v[i] != ',' && v[i] != ' '; ++i) {
        if (v[i] < '0' || v[i] > '9') return false;
        q += (unsigned)(v[i] - '0');
        ++digits;
    }
    return q != 0 && q <= 100 && digits <= 3;
}
bool accepts_gzip(std::string_view ae) {
    if (ae.size() < 4) return false;
    if (ae.substr(0, 5) == "gzip,") return true;
    auto is_gzip_at = [&](size_t i) {
        return i + 4 <= ae.size() && (ae[i] | 0x20) == 'g' && (ae[i + 1] | 0x20) == 'z' && (ae[i + 2] | 0x20) == 'i' && (ae[i + 3] | 0x20) == 'p';
    };
    size_t from = 0, at = std::string_view::npos;
    for (size_t i = 0; i + 4 <= ae.size(); ++i) {
        if (!is_gzip_at(i)) continue;
        if (i == from || ae[i - 1] == ',' || ae[i - 1] == ' ') { at = i; break; }
        from = i + 4;                                 // the search resumes right behind a rejected occurrence
        i += 3;
    }
    if (at == std::string_view::npos) return false;
    size_t i = at + 4;
    for (; i < ae.size() && ae[i] == ' '; ++i) {}
    if (i == ae.size() || ae[i] == ',') return true;
    if (ae[i] != ';') return false;
    for (++i; i < ae.size() && ae[i] == ' '; ++i) {}
    if (i == ae.size()) return true;
    if (ae[i] != 'q' && ae[i] != 'Q') return false;
    ++i;
    if (i + 2 > ae.size() || ae[i] != '=') return false;
    return qvalue_nonzero(ae.substr(i + 1));
}
}  // namespace

Plan plan_response(const Config& cfg, const Request& rq, const ObjectMeta& ob, uint64_t boundary) {
    if (!method_allowed(rq.method)) return stock_page(405, false);
    const bool head = rq.method == "HEAD";
    if (!head && rq.method != "GET") {
        Plan p;
        p.bypass = true;
        return p;
    }
    // representation (G1/G3 vs B3)
    const bool client_takes_gzip = rq.accept_encoding && accepts_gzip(*rq.accept_encoding);
    const bool inflated = ob.stored_gzip && cfg.gunzip && ob.raw_length >= 0 && !client_takes_gzip;
    const int64_t L = inflated ? ob.raw_length : (int64_t)ob.stored_length;
    std::optional<std::string> etag = ob.etag;
    if (inflated && etag && !(etag->size() >= 2 && (*etag)[0] == 'W' && (*etag)[1] == '/')) etag = "W/" + *etag;
    const int64_t lm = ob.last_modified ? parse_http_date(*ob.last_modified) : -1;

    // conditionals (C1-C3)
    if (rq.if_unmodified_since) {
        const int64_t t = parse_http_date(*rq.if_unmodified_since);
        if (lm == -1 || t < lm) return stock_page(412, head);      // C1: also when nothing is stored to compare with, or the date does not parse (-1)
    }
    if (rq.if_match && !etag_in_list(*rq.if_match, etag, false)) return stock_page(412, head);
    if (rq.if_modified_since || rq.if_none_match) {
        bool modified = false;
        if (rq.if_modified_since) modified = lm == -1 || parse_http_date(*rq.if_modified_since) != lm;
        if (!modified && rq.if_none_match && !etag_in_list(*rq.if_none_match, etag, true)) modified = true;
        if (!modified) {
            Plan p;
            p.status = 304;
            p.head = head;
            p.inflated = inflated;
            if (etag) p.headers.emplace_back("ETag", *etag);
            if (ob.last_modified) p.headers.emplace_back("Last-Modified", *ob.last_modified);
            return p;
        }
    }

    // Range / If-Range (R1, R2)
    std::vector<ByteRange> rg;
    RangeVerdict verdict = RangeVerdict::Ignore;
    // R0: without an Accept-Ranges header on the stored upstream response nginx's range filter stays out of it
    if (ob.allow_ranges && rq.http_minor >= 0 && rq.range && rq.range->size() >= 7 && iprefix(*rq.range, "bytes=")) {
        bool honour = true;
        if (rq.if_range) {
            const std::string& ir = *rq.if_range;
            if (!ir.empty() && ir.back() == '"') honour = etag && *etag == ir;
            else honour = lm != -1 && parse_http_date(ir) == lm;
        }
        if (honour) verdict = parse_ranges(std::string_view(*rq.range).substr(6), L, rg);
    }
    if (verdict == RangeVerdict::Unsatisfiable)
        return stock_page(416, head, std::make_pair(std::string("Content-Range"), "bytes */" + std::to_string(L)));

    Plan p;
    p.head = head;
    p.inflated = inflated;
    const bool say_gzip = ob.stored_gzip && !inflated;
    auto validators = [&] {
        if (etag) p.headers.emplace_back("ETag", *etag);
        if (ob.last_modified) p.headers.emplace_back("Last-Modified", *ob.last_modified);
    };
    if (verdict == RangeVerdict::Ignore) {
        p.status = 200;
        if (!ob.content_type.empty()) p.headers.emplace_back("Content-Type", ob.content_type);
        p.headers.emplace_back("Content-Length", std::to_string(L));
        validators();
        if (ob.allow_ranges) p.headers.emplace_back("Accept-Ranges", "bytes");
        if (say_gzip) p.headers.emplace_back("Content-Encoding", "gzip");
        if (L > 0) p.body.push_back(slice(0, (uint64_t)L));
        return p;
    }
    p.status = 206;
    if (rg.size() == 1) {   // R8
        if (!ob.content_type.empty()) p.headers.emplace_back("Content-Type", ob.content_type);
        p.headers.emplace_back("Content-Length", std::to_string(rg[0].end - rg[0].start));
        p.headers.emplace_back("Content-Range", "bytes " + std::to_string(rg[0].start) + "-" + std::to_string(rg[0].end - 1) +
                                                    "/" + std::to_string(L));
        validators();
        if (say_gzip) p.headers.emplace_back("Content-Encoding", "gzip");
        p.body.push_back(slice((uint64_t)rg[0].start, (uint64_t)(rg[0].end - rg[0].start)));
        return p;
    }
    // R9 multipart/byteranges
    char bnd[32];
    std::snprintf(bnd, sizeof bnd, "%020llu", (unsigned long long)boundary);
    uint64_t total = 0;
    for (const ByteRange& r : rg) {
        std::string h = std::string("\r\n--") + bnd + "\r\n";
        if (!ob.content_type.empty()) h += "Content-Type: " + ob.content_type + "\r\n";
        h += "Content-Range: bytes " + std::to_string(r.start) + "-" + std::to_string(r.end - 1) + "/" + std::to_string(L) +
             "\r\n\r\n";
        total += h.size() + (uint64_t)(r.end - r.start);
        p.body.push_back(literal(std::move(h)));
        p.body.push_back(slice((uint64_t)r.start, (uint64_t)(r.end - r.start)));
    }
    std::string tail = std::string("\r\n--") + bnd + "--\r\n";
    total += tail.size();
    p.body.push_back(literal(std::move(tail)));
    p.headers.emplace_back("Content-Type", std::string("multipart/byteranges; boundary=") + bnd);
    p.headers.emplace_back("Content-Length", std::to_string(total));
    validators();
    if (say_gzip) p.headers.emplace_back("Content-Encoding", "gzip");
    return p;
}

uint64_t Plan::body_length() const {
    uint64_t n = 0;
    for (const Part& b : body) n += b.literal ? b.bytes.size() : b.len;
    return n;
}

std::string Plan::to_text() const {
    if (bypass) return "status 0\nbypass\nend\n";
    std::string t = "status " + std::to_string(status) + "\nhead " + (head ? "1" : "0") + "\nrepr " +
                    (inflated ? "inflated" : "identity") + "\n";
    for (const auto& h : headers) t += "hdr " + h.first + ": " + h.second + "\n";
    static const char* hexd = "0123456789abcdef";
    for (const Part& b : body) {
        if (b.literal) {
            t += "lit ";
            for (unsigned char ch : b.bytes) { t.push_back(hexd[ch >> 4]); t.push_back(hexd[ch & 15]); }
            t.push_back('\n');
        } else {
            t += "slice " + std::to_string(b.off) + " " + std::to_string(b.len) + "\n";
        }
    }
    t += "end\n";
    return t;
}

}  // namespace nzh

// ---------------------------------------------------------------------------
// C view
// ---------------------------------------------------------------------------
namespace {
std::optional<std::string> opt(const char* s) { return s ? std::optional<std::string>(s) : std::nullopt; }
const nzh::Config& default_config() {
    static const nzh::Config c = nzh::Config::reference_default();
    return c;
}
}  // namespace

extern "C" {
int nzh_cache_key(const char* host_hdr, const char* target, char* out, size_t cap) {
    auto k = nzh::cache_key(default_config(), opt(host_hdr), target ? target : "", nullptr);
    if (!k || k->size() + 1 > cap) return -400;
    std::memcpy(out, k->c_str(), k->size() + 1);
    return (int)k->size();
}
int nzh_method_allowed(const char* method) { return nzh::method_allowed(method ? method : "") ? 1 : 0; }
int nzh_parse_ranges(const char* spec, int64_t length, int64_t* out_pairs, int max_pairs, int* n_out) {
    std::vector<nzh::ByteRange> v;
    const nzh::RangeVerdict r = nzh::parse_ranges(spec ? spec : "", length, v, (size_t)max_pairs);
    *n_out = 0;
    if (r == nzh::RangeVerdict::Unsatisfiable) return 416;
    if (r == nzh::RangeVerdict::Ignore) return 1;
    for (size_t i = 0; i < v.size(); ++i) { out_pairs[2 * i] = v[i].start; out_pairs[2 * i + 1] = v[i].end; }
    *n_out = (int)v.size();
    return 0;
}
int64_t nzh_parse_http_date(const char* s, size_t n) { return nzh::parse_http_date(std::string_view(s, n)); }
void nzh_md5(const void* data, size_t n, uint8_t out[16]) { nzh::md5(data, n, out); }
int nzh_plan_text(const nzh_request* rq, const nzh_object* ob, uint64_t boundary, char* out, size_t cap) {
    nzh::Request r;
    r.method = rq->method ? rq->method : "";
    r.http_minor = rq->http_minor;
    r.range = opt(rq->range); r.if_range = opt(rq->if_range); r.if_none_match = opt(rq->if_none_match);
    r.if_match = opt(rq->if_match); r.if_modified_since = opt(rq->if_modified_since);
    r.if_unmodified_since = opt(rq->if_unmodified_since); r.accept_encoding = opt(rq->accept_encoding);
    nzh::ObjectMeta o;
    o.stored_length = (uint64_t)ob->stored_length;
    o.raw_length = ob->raw_length;
    o.content_type = ob->content_type ? ob->content_type : "";
    o.etag = opt(ob->etag);
    o.last_modified = opt(ob->last_modified);
    o.stored_gzip = ob->content_encoding && strcasecmp(ob->content_encoding, "gzip") == 0;
    o.allow_ranges = !ob->no_ranges;
    nzh::Config cfg = default_config();
    cfg.gunzip = ob->gunzip_enabled != 0;
    const std::string t = nzh::plan_response(cfg, r, o, boundary).to_text();
    if (t.size() + 1 > cap) return -1;
    std::memcpy(out, t.c_str(), t.size() + 1);
    return (int)t.size();
}
int nzh_parse_head_text(const char* raw, size_t n, char* out, size_t cap) {
    const std::string t = nzh::parse_head(std::string_view(raw, n)).to_text();
    if (t.size() + 1 > cap) return -1;
    std::memcpy(out, t.c_str(), t.size() + 1);
    return (int)t.size();
}
int nzh_admission(int status, const char* headers, int64_t now, int64_t default_valid_s, int64_t* valid_until) {
    std::vector<std::pair<std::string, std::string>> hv;
    std::string_view all(headers ? headers : "");
    while (!all.empty()) {
        const size_t e = all.find('\n');
        const std::string_view line = all.substr(0, e);
        all = e == std::string_view::npos ? std::string_view() : all.substr(e + 1);
        const size_t c = line.find(':');
        if (c == std::string_view::npos) continue;
        size_t v = c + 1;
        while (v < line.size() && (line[v] == ' ' || line[v] == '\t')) ++v;
        size_t ve = line.size();
        while (ve > v && (line[ve - 1] == ' ' || line[ve - 1] == '\t' || line[ve - 1] == '\r')) --ve;
        hv.emplace_back(std::string(line.substr(0, c)), std::string(line.substr(v, ve - v)));
    }
    const nzh::Admission a = nzh::admission(status, hv, now, default_valid_s);
    if (valid_until) *valid_until = a.valid_until;
    return a.cacheable ? 1 : 0;
}
}
