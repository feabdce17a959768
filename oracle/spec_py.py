"""oracle/spec_py.py — a SECOND, independent restatement of the host-side rules (SPEC K, V, H2, C, R, E).

TEST INFRASTRUCTURE ONLY (tests/ may import it; the product never does).  PARITY UNPINNED, like everything
in oracle/: nginx cannot run here.

Why it exists: `nz_oracle.c`, `nz_cpu_server.c` and the product's `nz_http.cc` were written by one hand from one
reading of SPEC.md, so a misreading of nginx passes every differential between them.
This file is pure Python, uses no ctypes, shares no code or tables with those three, and was written from the
algorithms of the nginx modules the reference's nginx.conf configures, function by function:

  * ngx_http_parse_complex_uri()          -> normalize_uri()      (state machine over the request path)
  * ngx_http_validate_host()              -> normalize_host()
  * ngx_http_not_modified_header_filter() -> conditional()        (test_if_unmodified / test_if_match / test_if_modified)
  * ngx_http_range_header_filter(), ngx_http_range_parse(), *_singlepart_header(), *_multipart_header()
                                          -> ranges(), plan()
  * ngx_http_upstream_process_{accel_expires,cache_control,expires,set_cookie,vary}() -> admission()
  * ngx_parse_http_time()                 -> parse_http_time()
  * RFC 1321 / RFC 3720 for md5() / crc32c()

tests/test_spec_triple.py diffs it against the C oracle and the product under hypothesis.  Where this reading of
nginx and the first one disagreed, SPEC.md section Z records which one was taken and why.
"""
from __future__ import annotations

import struct
from dataclasses import dataclass, field
from typing import List, Optional, Sequence, Tuple

# ----------------------------------------------------------------------------------------------------------
# the reference's configuration (/root/reference/examples/demo/nginx.conf): server blocks in file order
# ----------------------------------------------------------------------------------------------------------
SERVERS = (                       # (server_name, proxy_pass)                                  nginx.conf lines
    ("download.tensorflow.org", "http://download.tensorflow.org"),       # 45, 55   <- first block = default server
    ("yann.lecun.com", "http://yann.lecun.com"),                         # 64, 74
    ("www.cs.toronto.edu", "http://www.cs.toronto.edu"),                 # 83, 93
    ("storage.googleapis.com", "https://storage.googleapis.com"),        # 103, 113
    ("images.cocodataset.org", "http://images.cocodataset.org"),         # 122, 132   (s3.amazonaws.com, 139-166, is commented out)
)
MAX_OFF_T = (1 << 63) - 1


def server_blocks() -> List[Tuple[str, str]]:
    """(server_name, proxy_pass) of every live server block, in file order, read from the conf text itself when the
    reference tree is present (this container), else the table above.  Only the first block matters beyond its name:
    it is the default server (no `default_server` flag anywhere)."""
    import os
    import re

    p = "/root/reference/examples/demo/nginx.conf"
    if not os.path.exists(p):
        return list(SERVERS)
    out, name = [], None
    for raw in open(p, encoding="utf-8", errors="replace"):
        line = raw.split("#", 1)[0].strip()
        m = re.match(r"server_name\s+([^;\s]+)\s*;", line)
        if m:
            name = m.group(1)
        m = re.match(r"proxy_pass\s+([^;\s]+)\s*;", line)
        if m and name is not None:
            out.append((name, m.group(1)))
            name = None
    return out or list(SERVERS)


# ----------------------------------------------------------------------------------------------------------
# K: host, uri, key
# ----------------------------------------------------------------------------------------------------------
def normalize_host(host: str) -> Optional[str]:
    """ngx_http_validate_host (1.15-era: no control-character / space check yet): lower-case; the name ends at the first
    ':' outside an IPv6 literal or right after the literal's ']'; ONE trailing dot is dropped; two dots in a row, a NUL
    or a '/' anywhere (the port part included) are invalid.  None = invalid or empty (400)."""
    USUAL, LITERAL, REST = 0, 1, 2
    state = USUAL
    n = len(host)
    host_len = n
    dot_pos = n
    for i, ch in enumerate(host):
        if ch == ".":
            if dot_pos == i - 1:
                return None
            dot_pos = i
        elif ch == ":":
            if state == USUAL:
                host_len = i
                state = REST
        elif ch == "[":
            if i == 0:
                state = LITERAL
        elif ch == "]":
            if state == LITERAL:
                host_len = i + 1
                state = REST
        elif ch == "\0" or ch == "/":
            return None
    if dot_pos == host_len - 1:
        host_len -= 1
    if host_len == 0:
        return None
    return "".join(c.lower() if "A" <= c <= "Z" else c for c in host[:host_len])


def split_target(target: str) -> Tuple[Optional[str], str]:
    """request-target -> (host from an absolute-form target or None, path-and-rest)."""
    tl = target[:8].lower()
    for scheme in ("http://", "https://"):
        if tl.startswith(scheme):
            rest = target[len(scheme):]
            cut = len(rest)
            for j, c in enumerate(rest):
                if c in "/?#":
                    cut = j
                    break
            path = rest[cut:]
            if not path.startswith("/"):
                path = "/" + path
            return rest[:cut], path
    return None, target


_HEX = "0123456789abcdefABCDEF"


def normalize_uri(path: str) -> Optional[str]:
    """ngx_http_parse_complex_uri as a state machine over CHARACTERS of the path (latin-1 code points).
    Returns the decoded, dot-resolved, slash-merged $uri, or None (400).  The query (first literal '?') and a
    literal '#' end the path.  %00 is invalid; an escape that decodes to '%', '#' or '?' is literal text."""
    if not path.startswith("/"):
        return None
    USUAL, SLASH, DOT, DOT_DOT = 0, 1, 2, 3
    out: List[str] = []
    state = USUAL
    i, n = 0, len(path)

    def feed(ch: str, literal_only: bool) -> Optional[bool]:
        """process one (possibly decoded) character in the current state; True = stop (args/fragment)."""
        nonlocal state
        if literal_only:                           # a decoded '%', '#', '?': plain text, back to USUAL
            out.append(ch)
            state = USUAL
            return False
        if state == USUAL:
            if ch == "/":
                state = SLASH
                out.append(ch)
            elif ch == "?" or ch == "#":
                return True
            else:
                out.append(ch)
            return False
        if state == SLASH:
            if ch == "/":
                return False                       # merge_slashes on (default): the duplicate is dropped
            if ch == ".":
                state = DOT
                out.append(ch)
            elif ch == "?" or ch == "#":
                return True
            else:
                state = USUAL
                out.append(ch)
            return False
        if state == DOT:
            if ch == "/":
                state = SLASH
                out.pop()                          # "/./" -> "/"
            elif ch == ".":
                state = DOT_DOT
                out.append(ch)
            elif ch == "?" or ch == "#":
                out.pop()                          # a trailing "/." resolves to "/"
                return True
            else:
                state = USUAL
                out.append(ch)
            return False
        # DOT_DOT: out ends with "/.."
        if ch == "/" or ch == "?" or ch == "#":
            # go up one segment: drop "/..", then the previous segment back to (and keeping) its leading slash
            del out[-3:]
            if not out:
                raise ValueError("above root")
            while out and out[-1] != "/":
                out.pop()
            if not out:
                raise ValueError("above root")
            # out now ends with the parent's trailing '/'
            state = SLASH
            if ch != "/":
                return True
            return False
        state = USUAL
        out.append(ch)
        return False

    try:
        while i < n:
            ch = path[i]
            i += 1
            if ch == "%":
                if i + 2 > n:
                    return None
                h1, h2 = path[i], path[i + 1]
                if h1 not in _HEX or h2 not in _HEX:
                    return None
                i += 2
                dec = chr(int(h1 + h2, 16))
                if dec == "\0":
                    return None
                if feed(dec, dec in "%#?"):
                    break
                continue
            if ch == "\0":
                return None
            if feed(ch, False):
                break
        else:
            # end of input: a path ending in "/." or "/.." resolves like one followed by a delimiter
            if state == DOT:
                out.pop()
            elif state == DOT_DOT:
                feed("?", False)
    except ValueError:
        return None
    return "".join(out)


def select_server(host: Optional[str], servers: Optional[Sequence[Tuple[str, str]]] = None) -> Tuple[str, str]:
    servers = list(servers) if servers is not None else list(SERVERS)
    if host is not None:
        for name, pp in servers:
            if name.lower() == host:
                return name, pp
    return servers[0]


def proxy_host_of(proxy_pass: str) -> str:
    """$proxy_host: host[:port] of the proxy_pass URL, the port only when it is not the scheme's default."""
    scheme, rest = proxy_pass.split("://", 1)
    hostport = rest.split("/", 1)[0]
    if ":" in hostport:
        h, p = hostport.rsplit(":", 1)
        if (scheme == "http" and p == "80") or (scheme == "https" and p == "443"):
            return h
    return hostport


def cache_key(host_header: Optional[str], target: str, servers: Optional[Sequence[Tuple[str, str]]] = None) -> Optional[bytes]:
    """proxy_cache_key "$host/$proxy_host$uri" (nginx.conf:56).  None = 400."""
    thost, path = split_target(target)
    raw = thost if thost is not None else host_header
    host = None
    if raw is not None:                               # a host that is present but empty / invalid is a 400, in either place
        host = normalize_host(raw)
        if host is None:
            return None
    name, pp = select_server(host, servers)
    uri = normalize_uri(path)
    if uri is None:
        return None
    h = host if host is not None else name            # $host falls back to the server name
    return (h + "/" + proxy_host_of(pp) + uri).encode("latin-1")


def method_gate(method: str) -> bool:
    """`if ($request_method !~ "GET|HEAD") { return 405; }`: an unanchored, case-sensitive search."""
    return ("GET" in method) or ("HEAD" in method)


# ----------------------------------------------------------------------------------------------------------
# RFC 1321 / RFC 3720
# ----------------------------------------------------------------------------------------------------------
def md5(data: bytes) -> bytes:
    import math

    s = [7, 12, 17, 22] * 4 + [5, 9, 14, 20] * 4 + [4, 11, 16, 23] * 4 + [6, 10, 15, 21] * 4
    K = [int(abs(math.sin(i + 1)) * 2 ** 32) & 0xFFFFFFFF for i in range(64)]
    a0, b0, c0, d0 = 0x67452301, 0xEFCDAB89, 0x98BADCFE, 0x10325476
    msg = bytearray(data)
    bitlen = (8 * len(data)) & 0xFFFFFFFFFFFFFFFF
    msg.append(0x80)
    while len(msg) % 64 != 56:
        msg.append(0)
    msg += struct.pack("<Q", bitlen)
    for off in range(0, len(msg), 64):
        M = struct.unpack("<16I", msg[off:off + 64])
        A, B, Cc, D = a0, b0, c0, d0
        for i in range(64):
            if i < 16:
                F, g = (B & Cc) | (~B & D), i
            elif i < 32:
                F, g = (D & B) | (~D & Cc), (5 * i + 1) % 16
            elif i < 48:
                F, g = B ^ Cc ^ D, (3 * i + 5) % 16
            else:
                F, g = Cc ^ (B | ~D), (7 * i) % 16
            F = (F + A + K[i] + M[g]) & 0xFFFFFFFF
            A, D, Cc = D, Cc, B
            B = (B + ((F << s[i]) | (F >> (32 - s[i])))) & 0xFFFFFFFF
        a0, b0, c0, d0 = (a0 + A) & 0xFFFFFFFF, (b0 + B) & 0xFFFFFFFF, (c0 + Cc) & 0xFFFFFFFF, (d0 + D) & 0xFFFFFFFF
    return struct.pack("<4I", a0, b0, c0, d0)


def crc32c(data: bytes, crc: int = 0) -> int:
    """bit at a time, reflected 0x1EDC6F41."""
    c = crc ^ 0xFFFFFFFF
    for b in data:
        c ^= b
        for _ in range(8):
            c = (c >> 1) ^ (0x82F63B78 if c & 1 else 0)
    return c ^ 0xFFFFFFFF


# ----------------------------------------------------------------------------------------------------------
# ngx_parse_http_time
# ----------------------------------------------------------------------------------------------------------
_MONTHS = {"Jan": 0, "Feb": 1, "Mar": 2, "Apr": 3, "May": 4, "Jun": 5, "Jul": 6, "Aug": 7, "Sep": 8, "Oct": 9, "Nov": 10, "Dec": 11}
_MDAY = [31, 28, 31, 30, 31, 30, 31, 31, 30, 31, 30, 31]


def parse_http_time(value: Optional[str]) -> int:
    """RFC 822 / RFC 850 / asctime -> seconds since the epoch, -1 on anything else.  Follows ngx_parse_http_time: the
    weekday is skipped without being checked, the time zone is never looked at."""
    if value is None:
        return -1
    s = value
    n = len(s)
    p = 0
    fmt = None
    # skip the weekday: up to ',' (rfc822 / rfc850) or ' ' (isoc = asctime)
    while p < n:
        if s[p] == ",":
            break
        if s[p] == " ":
            fmt = "isoc"
            break
        p += 1
    else:
        return -1
    p += 1
    while p < n and s[p] == " ":
        p += 1
    if p == n:
        return -1

    def two(q):
        if q + 2 > n or not (s[q].isdigit() and s[q + 1].isdigit()) or not (s[q].isascii() and s[q + 1].isascii()):
            return None
        return int(s[q:q + 2])

    day = month = year = 0
    if fmt != "isoc":
        if n - p < 18:
            return -1
        d = two(p)
        if d is None:
            return -1
        day = d
        p += 2
        if s[p] == " ":
            if n - p < 18:
                return -1
            fmt = "rfc822"
        elif s[p] == "-":
            fmt = "rfc850"
        else:
            return -1
        p += 1
    mon = s[p:p + 3]
    if mon not in _MONTHS:
        return -1
    month = _MONTHS[mon]
    p += 3
    if (fmt == "rfc822" and (p >= n or s[p] != " ")) or (fmt == "rfc850" and (p >= n or s[p] != "-")):
        return -1
    p += 1
    if fmt == "rfc822":
        if p + 4 > n or not all(c.isascii() and c.isdigit() for c in s[p:p + 4]):
            return -1
        year = int(s[p:p + 4])
        p += 4
    elif fmt == "rfc850":
        y = two(p)
        if y is None:
            return -1
        year = y + (2000 if y < 70 else 1900)
        p += 2
    if fmt == "isoc":
        if p < n and s[p] == " ":
            p += 1
        if p >= n or not (s[p].isascii() and s[p].isdigit()):
            return -1
        day = int(s[p])
        p += 1
        if p < n and s[p] != " ":
            if not (s[p].isascii() and s[p].isdigit()):
                return -1
            day = day * 10 + int(s[p])
            p += 1
        if n - p < 14:
            return -1
    if p >= n or s[p] != " ":
        return -1
    p += 1
    hh = two(p)
    if hh is None or p + 2 >= n or s[p + 2] != ":":
        return -1
    mm = two(p + 3)
    if mm is None or p + 5 >= n or s[p + 5] != ":":
        return -1
    ss = two(p + 6)
    if ss is None:
        return -1
    p += 8
    if fmt == "isoc":
        if p >= n or s[p] != " ":
            return -1
        p += 1
        if p + 4 > n or not all(c.isascii() and c.isdigit() for c in s[p:p + 4]):
            return -1
        year = int(s[p:p + 4])
    if hh > 23 or mm > 59 or ss > 59:
        return -1
    if day == 29 and month == 1:
        if (year & 3) or ((year % 100 == 0) and (year % 400) != 0):
            return -1
    elif day > _MDAY[month]:
        return -1
    # Gauss' formula as nginx uses it (March-based year)
    month -= 1
    if month <= 0:
        month += 12
        year -= 1
    days = 365 * year + year // 4 - year // 100 + year // 400 + 367 * month // 12 - 30 + day - 1 - 719527 + 31 + 28
    t = days * 86400 + hh * 3600 + mm * 60 + ss
    return t


# ----------------------------------------------------------------------------------------------------------
# H2: admission / validity of an upstream response (header handlers run in arrival order)
# ----------------------------------------------------------------------------------------------------------
def _find_ci(hay: str, needle: str) -> int:
    return hay.lower().find(needle.lower())


def admission(status: int, headers: Sequence[Tuple[str, str]], now: int, default_valid: int = 86400) -> Tuple[bool, int]:
    """(cacheable, valid_until) with proxy_cache_valid 200 1d and no proxy_ignore_headers (nginx.conf:57)."""
    cacheable = True
    valid = 0
    seen_accel = False
    for name, value in headers:
        ln = name.lower()
        if ln == "set-cookie":
            cacheable = False
        elif ln == "vary":
            if value == "*" or len(value) > 128:       # nginx: "*" or longer than NGX_HTTP_CACHE_VARY_LEN is never cached
                cacheable = False
        elif ln == "x-accel-expires":
            seen_accel = True
            if not value:
                continue
            if value[0] != "@":
                if value.isascii() and value.isdigit():
                    nsec = int(value)
                    if nsec == 0:
                        cacheable = False
                    else:
                        valid = now + nsec
            else:
                rest = value[1:]
                if rest and rest.isascii() and rest.isdigit():
                    valid = int(rest)
        elif ln == "cache-control":
            if valid != 0 and seen_accel:
                continue
            if _find_ci(value, "no-cache") >= 0 or _find_ci(value, "no-store") >= 0 or _find_ci(value, "private") >= 0:
                cacheable = False
                continue
            pos = _find_ci(value, "s-maxage=")
            off = 9
            if pos < 0:
                pos = _find_ci(value, "max-age=")
                off = 8
            if pos < 0:
                continue
            nsec = 0
            bad = False
            for ch in value[pos + off:]:
                if ch in ",; ":
                    break
                if "0" <= ch <= "9":
                    nsec = nsec * 10 + ord(ch) - 48
                else:
                    bad = True
                    break
            if bad or nsec == 0:
                cacheable = False
            else:
                valid = now + nsec
        elif ln == "expires":
            if valid != 0:
                continue
            t = parse_http_time(value)
            if t == -1 or t < now:
                cacheable = False
            else:
                valid = t
    if status != 200:
        cacheable = False
    if cacheable and valid == 0:
        valid = now + default_valid
    return cacheable, (valid if cacheable else 0)


# ----------------------------------------------------------------------------------------------------------
# C / R: the not-modified and range header filters
# ----------------------------------------------------------------------------------------------------------
@dataclass
class Stored:
    length: int                               # stored (at-rest) length
    content_type: str = "application/octet-stream"
    etag: Optional[str] = None
    last_modified: Optional[str] = None
    content_encoding: Optional[str] = None
    raw_length: int = -1                      # inflated length of a gzip object, -1 unknown
    gunzip: bool = False                      # operator enabled inflate-on-hit (SPEC G)
    accept_ranges: bool = True                # the stored upstream response carried an Accept-Ranges header (any value)


@dataclass
class Request:
    method: str = "GET"
    http_minor: int = 1
    range: Optional[str] = None
    if_range: Optional[str] = None
    if_none_match: Optional[str] = None
    if_match: Optional[str] = None
    if_modified_since: Optional[str] = None
    if_unmodified_since: Optional[str] = None
    accept_encoding: Optional[str] = None


@dataclass
class Response:
    status: int
    head: bool = False
    repr: str = "identity"
    headers: List[Tuple[str, str]] = field(default_factory=list)
    body: List[Tuple[str, object]] = field(default_factory=list)       # ("lit", bytes) | ("slice", (start, len))
    bypass: bool = False


_REASON = {400: "Bad Request", 405: "Not Allowed", 412: "Precondition Failed", 416: "Requested Range Not Satisfiable"}


def error_page(code: int) -> bytes:
    t = f"{code} {_REASON[code]}"
    return (f"<html>\r\n<head><title>{t}</title></head>\r\n<body>\r\n<center><h1>{t}</h1></center>\r\n"
            "<hr><center>nginx</center>\r\n</body>\r\n</html>\r\n").encode("latin-1")


def _special(code: int, head: bool, extra: Sequence[Tuple[str, str]] = ()) -> Response:
    page = error_page(code)
    r = Response(code, head)
    r.headers.append(("Content-Type", "text/html"))
    r.headers.append(("Content-Length", str(len(page))))
    r.headers.extend(extra)
    r.body.append(("lit", page))
    return r


def etag_match(header: str, etag: Optional[str], weak: bool) -> bool:
    """ngx_http_test_if_match."""
    if header == "*":
        return True
    if etag is None:
        return False
    if weak and len(etag) > 2 and etag.startswith("W/"):
        etag = etag[2:]
    p, end = 0, len(header)
    while p < end:
        if weak and end - p > 2 and header[p] == "W" and header[p + 1] == "/":
            p += 2
        if len(etag) > end - p:
            return False
        if header[p:p + len(etag)] == etag:
            q = p + len(etag)
            while q < end and header[q] in " \t":
                q += 1
            if q == end or header[q] == ",":
                return True
            p = q
        while p < end and header[p] != ",":
            p += 1
        while p < end and header[p] in " \t,":
            p += 1
    return False


def _gzip_quantum(s: str, p: int) -> int:
    """ngx_http_gzip_quantum: 0 / 1 [ "." up to three digits ] -> hundredths, 0 when malformed."""
    n = len(s)
    c = s[p]
    p += 1
    if c != "0" and c != "1":
        return 0
    q = (ord(c) - 48) * 100
    if p == n:
        return q
    c = s[p]
    p += 1
    if c == "," or c == " ":
        return q
    if c != ".":
        return 0
    digits = 0
    while p < n:
        c = s[p]
        p += 1
        if c == "," or c == " ":
            break
        if "0" <= c <= "9":
            q += ord(c) - 48
            digits += 1
            continue
        return 0
    if q > 100 or digits > 3:
        return 0
    return q


def client_accepts_gzip(ae: Optional[str]) -> bool:
    """The Accept-Encoding part of ngx_http_gzip_ok (which the gunzip filter consults): the literal prefix "gzip," or
    ngx_http_gzip_accept_encoding - a case-insensitive "gzip" at the start or after ',' / ' ', followed by blanks and
    ',' or the end, or by ";q=" with a non-zero value."""
    if ae is None or len(ae) < 4:
        return False
    if ae.startswith("gzip,"):
        return True
    low = ae.lower()
    start = 0
    while True:
        p = low.find("gzip", start)
        if p < 0:
            return False
        if p == start or ae[p - 1] in ", ":
            break
        start = p + 4
    p += 4
    n = len(ae)
    while p < n:
        c = ae[p]
        p += 1
        if c == ",":
            return True
        if c == ";":
            break
        if c == " ":
            continue
        return False
    else:
        return True
    while p < n:                                  # after ';'
        c = ae[p]
        p += 1
        if c in "qQ":
            break
        if c == " ":
            continue
        return False
    else:
        return True
    if p + 2 > n or ae[p] != "=":
        return False
    return _gzip_quantum(ae, p + 1) != 0


NOT_SATISFIABLE, DECLINED, OK = 416, -5, 0


def range_parse(spec: str, length: int) -> Tuple[int, List[Tuple[int, int]]]:
    """ngx_http_range_parse over the text after "bytes=".  Returns (OK|DECLINED|NOT_SATISFIABLE, [(start, end_exclusive)])."""
    s = spec + "\0"
    p = 0
    size = 0
    out: List[Tuple[int, int]] = []
    cutoff, cutlim = MAX_OFF_T // 10, MAX_OFF_T % 10
    while True:
        start = end = 0
        suffix = False
        have = False
        while s[p] == " ":
            p += 1
        if s[p] != "-":
            if not ("0" <= s[p] <= "9"):
                return NOT_SATISFIABLE, []
            while "0" <= s[p] <= "9":
                d = ord(s[p]) - 48
                if start >= cutoff and (start > cutoff or d > cutlim):
                    return NOT_SATISFIABLE, []
                start = start * 10 + d
                p += 1
            while s[p] == " ":
                p += 1
            if s[p] != "-":
                return NOT_SATISFIABLE, []
            p += 1
            while s[p] == " ":
                p += 1
            if s[p] == "," or s[p] == "\0":
                end = length
                have = True
        else:
            suffix = True
            p += 1
        if not have:
            if not ("0" <= s[p] <= "9"):
                return NOT_SATISFIABLE, []
            while "0" <= s[p] <= "9":
                d = ord(s[p]) - 48
                if end >= cutoff and (end > cutoff or d > cutlim):
                    return NOT_SATISFIABLE, []
                end = end * 10 + d
                p += 1
            while s[p] == " ":
                p += 1
            if s[p] != "," and s[p] != "\0":
                return NOT_SATISFIABLE, []
            if suffix:
                start = length - end if end < length else 0
                end = length - 1
            if end >= length:
                end = length
            else:
                end += 1
        if start < end:
            out.append((start, end))
            if size > MAX_OFF_T - (end - start):
                return NOT_SATISFIABLE, []
            size += end - start
        elif start == 0:
            return DECLINED, []
        if s[p] != ",":
            break
        p += 1
    if not out:
        return NOT_SATISFIABLE, []
    if size > length:
        return DECLINED, []
    return OK, out


def plan(rq: Request, ob: Stored, boundary: int = 1) -> Response:
    """What the cached GET/HEAD answers: method gate -> representation -> not-modified filter -> range filter."""
    if not method_gate(rq.method):
        return _special(405, False)
    if rq.method not in ("GET", "HEAD"):
        return Response(0, bypass=True)
    head = rq.method == "HEAD"
    gz = ob.content_encoding is not None and ob.content_encoding.lower() == "gzip"
    inflated = gz and ob.gunzip and ob.raw_length >= 0 and not client_accepts_gzip(rq.accept_encoding)
    L = ob.raw_length if inflated else ob.length
    etag = ob.etag
    if inflated and etag is not None and not etag.startswith("W/"):
        etag = "W/" + etag
    lm = parse_http_time(ob.last_modified) if ob.last_modified is not None else -1
    rep = "inflated" if inflated else "identity"

    # ---- not-modified filter (status 200 only) ----
    if rq.if_unmodified_since is not None:
        ok = lm != -1 and parse_http_time(rq.if_unmodified_since) >= lm          # test_if_unmodified
        if not ok:
            return _special(412, head)
    if rq.if_match is not None and not etag_match(rq.if_match, etag, False):
        return _special(412, head)
    if rq.if_modified_since is not None or rq.if_none_match is not None:
        modified = False
        if rq.if_modified_since is not None:
            modified = lm == -1 or parse_http_time(rq.if_modified_since) != lm    # if_modified_since exact
        if not modified and rq.if_none_match is not None and not etag_match(rq.if_none_match, etag, True):
            modified = True
        if not modified:
            r = Response(304, head, rep)
            if etag is not None:
                r.headers.append(("ETag", etag))
            if ob.last_modified is not None:
                r.headers.append(("Last-Modified", ob.last_modified))
            return r

    # ---- range filter ----
    rc, rs = DECLINED, []
    # ngx_http_upstream_copy_allow_ranges: on a cached response r->allow_ranges is set only by the upstream's own
    # Accept-Ranges header (the conf has no proxy_force_ranges); without it the range filter returns at once
    if ob.accept_ranges and rq.http_minor >= 0 and rq.range is not None and len(rq.range) >= 7 and rq.range[:6].lower() == "bytes=":
        use = True
        if rq.if_range is not None:
            ir = rq.if_range
            if len(ir) >= 2 and ir[-1] == '"':
                use = etag is not None and ir == etag
            else:
                use = lm != -1 and parse_http_time(ir) == lm
        if use:
            rc, rs = range_parse(rq.range[6:], L)
    common: List[Tuple[str, str]] = []
    if etag is not None:
        common.append(("ETag", etag))
    if ob.last_modified is not None:
        common.append(("Last-Modified", ob.last_modified))
    enc = [("Content-Encoding", "gzip")] if gz and not inflated else []
    if rc == NOT_SATISFIABLE:
        return _special(416, head, [("Content-Range", f"bytes */{L}")])
    if rc == DECLINED:
        r = Response(200, head, rep)
        if ob.content_type:
            r.headers.append(("Content-Type", ob.content_type))
        r.headers.append(("Content-Length", str(L)))
        r.headers += common + ([("Accept-Ranges", "bytes")] if ob.accept_ranges else []) + enc
        if L > 0:
            r.body.append(("slice", (0, L)))
        return r
    if len(rs) == 1:
        s0, e0 = rs[0]
        r = Response(206, head, rep)
        if ob.content_type:
            r.headers.append(("Content-Type", ob.content_type))
        r.headers.append(("Content-Length", str(e0 - s0)))
        r.headers.append(("Content-Range", f"bytes {s0}-{e0 - 1}/{L}"))
        r.headers += common + enc
        r.body.append(("slice", (s0, e0 - s0)))
        return r
    b = "%020d" % boundary
    r = Response(206, head, rep)
    total = 0
    pieces: List[Tuple[str, object]] = []
    for s0, e0 in rs:
        hdr = f"\r\n--{b}\r\n"
        if ob.content_type:
            hdr += f"Content-Type: {ob.content_type}\r\n"
        hdr += f"Content-Range: bytes {s0}-{e0 - 1}/{L}\r\n\r\n"
        pieces.append(("lit", hdr.encode("latin-1")))
        pieces.append(("slice", (s0, e0 - s0)))
        total += len(hdr) + (e0 - s0)
    tail = f"\r\n--{b}--\r\n".encode("latin-1")
    pieces.append(("lit", tail))
    total += len(tail)
    r.headers.append(("Content-Type", f"multipart/byteranges; boundary={b}"))
    r.headers.append(("Content-Length", str(total)))
    r.headers += common + enc
    r.body = pieces
    return r


# ----------------------------------------------------------------------------------------------------------
# P: the request head (ngx_http_parse_request_line, ngx_http_parse_header_line, the ngx_http_headers_in handlers)
# ----------------------------------------------------------------------------------------------------------
@dataclass
class Head:
    error: int = 0                      # 0, 400 or 505: answered before any vhost / cache logic runs
    method: str = ""
    target: str = ""
    minor: int = 1                      # -1 = HTTP/0.9, 0 = HTTP/1.0, 1 = HTTP/1.1 or later
    keep_alive: bool = True
    has_body: bool = False
    fields: dict = field(default_factory=dict)   # lower-case name -> value, only the names K/C/R/G read


_FIRST_WINS = ("range", "accept-encoding")
_UNIQUE = ("if-modified-since", "if-unmodified-since", "if-match", "if-none-match", "if-range", "content-length", "content-range", "expect", "authorization")


def parse_head(raw: bytes) -> Head:
    """SPEC P.  `raw` = everything before the blank line (with or without the final CRLF), as bytes."""
    text = raw.decode("latin-1")
    i = 0
    while i < len(text) and text[i] in "\r\n":            # sw_start skips CR and LF before the request line
        i += 1
    text = text[i:]
    while text.endswith("\r\n"):
        text = text[:-2]
    lines = text.split("\r\n")
    bad = Head(error=400, keep_alive=False)
    if "\0" in text or any("\r" in ln or "\n" in ln for ln in lines):
        return bad
    # ---- request line (P1-P4)
    rl = lines[0]
    j = 0
    while j < len(rl) and ("A" <= rl[j] <= "Z" or rl[j] == "_"):
        j += 1
    if j == 0 or j >= len(rl) or rl[j] != " ":
        return bad
    h = Head(method=rl[:j])
    while j < len(rl) and rl[j] == " ":
        j += 1
    k = rl.find(" ", j)
    h.target = rl[j:] if k < 0 else rl[j:k]
    if not h.target:
        return bad
    if not h.target.startswith("/"):
        m = 0
        while m < len(h.target) and h.target[m].isascii() and h.target[m].isalpha():
            m += 1
        if m == 0 or not h.target.startswith("://", m):
            return bad
    rest = "" if k < 0 else rl[k:].strip(" ")
    if rest == "":
        if h.method != "GET":
            return bad
        h.minor, h.keep_alive = -1, False
    else:
        if not rest.startswith("HTTP/"):
            return bad
        major, dot, minor = rest[5:].partition(".")
        ok = (dot == "." and 1 <= len(major) <= 2 and 1 <= len(minor) <= 2 and major.isascii() and minor.isascii() and major.isdigit() and minor.isdigit()
              and major[0] != "0")
        if not ok:
            return bad
        if int(major) > 1:
            return Head(error=505, keep_alive=False)
        h.minor = 0 if int(minor) == 0 else 1
        h.keep_alive = h.minor == 1
    # ---- header lines (P5-P10)
    hosts = 0
    for ln in lines[1:]:
        if ln == "":
            break
        name, colon, value = ln.partition(":")
        if not colon:
            name, value = ln, ""
        if name == "" or any(not (c.isascii() and (c.isalnum() or c == "-")) for c in name):
            continue                                          # ignore_invalid_headers on
        name = name.lower()
        value = value.strip(" ")
        if name == "host":
            hosts += 1
            if normalize_host(value) is None:
                return bad
            if hosts == 1:
                h.fields["host"] = value
        elif name in _FIRST_WINS:
            h.fields.setdefault(name, value)
        elif name in _UNIQUE:
            if name in h.fields:
                return bad
            h.fields[name] = value
            if name == "content-length":
                if value == "" or not (value.isascii() and value.isdigit()):
                    return bad
                if int(value) > 0:
                    h.has_body = True
        elif name == "transfer-encoding":
            h.has_body = True
        elif name == "connection":
            low = value.lower()
            if "close" in low:
                h.keep_alive = False
            elif "keep-alive" in low:
                h.keep_alive = True
    if h.minor == 1 and hosts == 0:
        return bad
    return h
