"""-m "not gpu": THREE restatements of the host-side rules diffed against each other under hypothesis:

  * oracle/spec_py.py        pure Python, written from the nginx modules' algorithms (no ctypes, no shared code)
  * oracle/nz_oracle.c       the C oracle (through oracle/oracle.py)
  * nezha_b200/host/nz_http.cc  the product's host layer (through nezha_b200/hostlib.py)

The first two come from different readings of nginx (the C oracle and the product shared one).
What the second reading changed is listed in oracle/SPEC.md, section Z; those cases are pinned below by name."""
import os
import string

from hypothesis import HealthCheck, given, settings, strategies as st

from nezha_b200 import hostlib as h
from oracle import oracle as o, spec_py as sp

SET = settings(max_examples=int(os.environ.get("NZ_HYP_EXAMPLES", "400")), deadline=None, suppress_health_check=[HealthCheck.too_slow])

HOSTS = ["www.cs.toronto.edu", "WWW.CS.TORONTO.EDU:80", "yann.lecun.com.", "storage.googleapis.com", "example.org", "images.cocodataset.org:8080",
         "download.tensorflow.org", "[::1]:80", "[::1]", "[::1", "bad host", "", None, "a/b", "x:y:z", "x:y/z", "a..b", "a.", ".", "a.:80", "Ä.example"]
path_tok = st.sampled_from(list("ab/.%?#2F0e~-_ +") + ["%2e", "%2F", "%00", "%zz", "..", "//", "/./", "%41", "%3f", "%3F", "%23", "%25", "%2", "/../", "/.."])


def test_conf_server_blocks_match_the_table():
    # in this container the table is re-read from the reference's own nginx.conf; on the GPU box the committed table is used
    assert sp.server_blocks() == list(sp.SERVERS)
    assert sp.SERVERS[0][0] == "download.tensorflow.org"


@SET
@given(st.sampled_from(HOSTS), st.lists(path_tok, max_size=14).map("".join), st.sampled_from(["", "http://yann.lecun.com", "HTTPS://Storage.GoogleApis.com:443", "http://"]))
def test_cache_key_three_way(host, tail, prefix):
    target = prefix + "/" + tail
    a, b, c = sp.cache_key(host, target), o.cache_key(host, target), h.cache_key(host, target)
    assert a == b == c, (host, target, a, b, c)
    if a is not None:
        assert sp.md5(a) == o.md5(a) == h.md5(a)


@SET
@given(st.text(alphabet="GETHADPOSTgeth X", max_size=8))
def test_method_gate_three_way(m):
    assert sp.method_gate(m) == o.method_gate(m) == h.method_allowed(m)


range_tok = st.sampled_from(["0", "1", "5", "99", "100", "999", "1000", "1001", "65535", "9223372036854775807", "9223372036854775808",
                             "99999999999999999999", "-", ",", " ", "  ", "a", "x", "-5", "0-", "-0", "\t"])


@SET
@given(st.lists(range_tok, max_size=9).map("".join), st.sampled_from([0, 1, 2, 100, 1000, 1 << 40, (1 << 63) - 1]))
def test_range_parse_three_way(spec, length):
    rc_p, rg_p = sp.range_parse(spec, length)
    rc_o, rg_o = o.range_parse(spec, length)
    rc_h, rg_h = h.parse_ranges(spec, length)
    assert (rc_p == sp.OK) == (rc_o == 0) == (rc_h == 0), (spec, length, rc_p, rc_o, rc_h)
    assert (rc_p == sp.NOT_SATISFIABLE) == (rc_o == 416) == (rc_h == 416), (spec, length, rc_p, rc_o, rc_h)
    if rc_p == sp.OK:
        assert rg_p == rg_o == rg_h, (spec, length)


DATES = ["Sun, 06 Nov 1994 08:49:37 GMT", "Sunday, 06-Nov-94 08:49:37 GMT", "Sun Nov  6 08:49:37 1994", "Wed, 07 Nov 2018 15:20:01 GMT",
         "Wed, 07 Nov 2018 15:20:02 GMT", "Tue, 29 Feb 2000 12:00:00 GMT", "Mon, 29 Feb 2001 12:00:00 GMT", "Thu, 31 Apr 2020 00:00:00 GMT",
         "Thu, 01 Jan 1970 00:00:00 GMT", "Fri, 31 Dec 1999 23:59:60 GMT", "Sat, 01 Jan 2000 24:00:00 GMT", "Sun Nov 16 08:49:37 1994",
         "Sunday, 06-Nov-69 08:49:37 GMT", "Sunday, 06-Nov-70 08:49:37 GMT", "garbage", "", "Sun, 06 Nov 1994", "Sun, 6 Nov 1994 08:49:37 GMT",
         "Sun, 06 Nox 1994 08:49:37 GMT", "Sun, 06-Nov 1994 08:49:37 GMT", "Wed, 00 Nov 2018 15:20:01 GMT", "Sun Nov  0 08:49:37 1994",
         "Mon, 29 Feb 1900 00:00:00 GMT", "Sat, 01 Jan 2000 00:00:00 +0100", "xxx, 01 Jan 2000 00:00:00 GMT"]


@SET
@given(st.sampled_from(DATES), st.lists(st.tuples(st.integers(0, 33), st.sampled_from(string.ascii_letters + string.digits + " ,:-")), max_size=3))
def test_http_date_three_way(base, edits):
    l = list(base)
    for pos, ch in edits:
        if pos < len(l):
            l[pos] = ch
    d = "".join(l)
    assert sp.parse_http_time(d) == o.parse_http_time(d) == h.parse_http_date(d), d


ETAGS = [None, '"abc"', 'W/"abc"', '"5be2f2c5-a12c15d"', "", "W/"]
ETAG_LISTS = [None, "*", '"abc"', 'W/"abc"', '"x", "abc"', '"x","abc"', '"abc" , "y"', '"abcd"', 'W/"x", W/"abc"', "", " ", '"ab', 'abc',
              '"5be2f2c5-a12c15d"', "W/", ",", '"abc"x', ' "abc"', '"abc"\t', "* ", '"abc"x, "abc"']
RANGES = [None, "bytes=0-0", "bytes=-1", "bytes=5-", "bytes=0-9,20-29", "bytes=10-5", "bytes=2000-", "BYTES=1-2", "bytes=", "bytes=0-", "items=0-1",
          "bytes=0-999,0-999", "bytes= 1 - 2 ,-3", "bytes=0-0,0-0,0-0", "bytes=a", "bytes=-0", "bytes=99999999999999999999-", "bytes=0-0,", "bytes=,0-0"]
IF_RANGES = [None, '"abc"', 'W/"abc"', "Wed, 07 Nov 2018 15:20:01 GMT", "Wed, 07 Nov 2018 15:20:02 GMT", "junk", '"', '""']
LMS = [None, "Wed, 07 Nov 2018 15:20:01 GMT", "not a date"]
AES = [None, "gzip", "gzip, deflate", "gzip,deflate", "identity", "GZIP", "gzip;q=0", "gzip;q=0.0", "gzip; q=0.001", "gzip;q=1.0001", "deflate, gzip;q=0", "xgzip",
       "gzipx", "xgzipgzip", "x-gzip", "gzip;", "gzip;x", "gzip ;q=0", "deflate,GZip", "gz", "br;q=1, gzip"]


def _norm_py(r):
    return (r.status, r.head if r.status else False, r.repr if r.status in (200, 206, 304) else "", sorted(r.headers),
            [(k, (tuple(v) if k == "slice" else bytes(v))) for k, v in r.body], r.bypass)


def _norm_text(t):
    p = o.parse_plan(t)
    return (p.status, p.head if p.status else False, p.repr if p.status in (200, 206, 304) else "", sorted(p.headers),
            [(k, (tuple(v) if k == "slice" else bytes(v))) for k, v in p.body], p.bypass)


@SET
@given(st.sampled_from(["GET", "HEAD", "POST", "XGET", "get"]), st.sampled_from([-1, 0, 1]), st.sampled_from(RANGES), st.sampled_from(IF_RANGES),
       st.sampled_from(ETAG_LISTS), st.sampled_from(ETAG_LISTS), st.sampled_from([None, "junk"] + DATES[:6]), st.sampled_from([None, "junk"] + DATES[:6]),
       st.sampled_from(AES), st.sampled_from([0, 1, 10, 1000]), st.sampled_from(ETAGS), st.sampled_from(LMS),
       st.sampled_from(["", "application/x-gzip", "text/plain"]), st.sampled_from([None, "gzip", "GZIP", "br"]), st.sampled_from([-1, 0, 5000]), st.booleans(),
       st.integers(0, 2**63), st.sampled_from([True, True, True, False]))
def test_plan_three_way(method, minor, rng, if_range, inm, im, ims, ius, ae, slen, etag, lm, ctype, cenc, rawlen, gunzip, bnd, ar):
    a = _norm_py(sp.plan(sp.Request(method, minor, rng, if_range, inm, im, ims, ius, ae), sp.Stored(slen, ctype, etag, lm, cenc, rawlen, gunzip, ar), bnd))
    b = _norm_text(o.plan(o.Req(method, minor, rng, if_range, inm, im, ims, ius, ae), o.Obj(slen, ctype, etag, lm, cenc, rawlen, gunzip, ar), bnd))
    c = _norm_text(h.plan_text(method, minor, rng, if_range, inm, im, ims, ius, ae, stored_length=slen, raw_length=rawlen, content_type=ctype, etag=etag,
                               last_modified=lm, content_encoding=cenc, gunzip_enabled=gunzip, boundary=bnd, accept_ranges=ar))
    assert a == b == c, (method, minor, rng, if_range, inm, im, ims, ius, ae, slen, etag, lm, ctype, cenc, rawlen, gunzip, ar)


HN = ["Cache-Control", "X-Accel-Expires", "Expires", "Set-Cookie", "Vary", "ETag", "cache-control", "EXPIRES"]
HV = ["max-age=60", "s-maxage=5, max-age=9", "no-cache", "private, max-age=5", "max-age=0", "max-age=x", "0", "30", "@1700000000", "@", "abc",
      "Wed, 07 Nov 2018 15:20:01 GMT", "Wed, 07 Nov 2040 15:20:01 GMT", "junk", "*", "Accept-Encoding", "a=b", "MAX-AGE=7", "public,max-age=10;x", "", "@12x", "007",
      "Origin, " * 16, "x" * 128, "x" * 129]


@SET
@given(st.sampled_from([200, 200, 200, 206, 301, 404]), st.lists(st.tuples(st.sampled_from(HN), st.sampled_from(HV)), max_size=5))
def test_admission_three_way(status, headers):
    now = 1_600_000_000
    a, b, c = sp.admission(status, headers, now), o.admission(status, headers, now), h.admission(status, headers, now)
    assert a[0] == b[0] == c[0], (status, headers, a, b, c)
    if a[0]:
        assert a[1] == b[1] == c[1], (status, headers, a, b, c)


# ---- what the second reading changed (oracle/SPEC.md section Z), pinned by name in all three ------------------------------
def _status3(**kw):
    rq = dict(method="GET", http_minor=1, range=None, if_range=None, if_none_match=None, if_match=None, if_modified_since=None, if_unmodified_since=None,
              accept_encoding=None)
    ob = dict(stored_length=1000, raw_length=-1, content_type="application/x-gzip", etag='"abc"', last_modified="Wed, 07 Nov 2018 15:20:01 GMT",
              content_encoding=None, gunzip_enabled=False)
    rq.update({k: v for k, v in kw.items() if k in rq})
    ob.update({k: v for k, v in kw.items() if k in ob})
    a = sp.plan(sp.Request(rq["method"], rq["http_minor"], rq["range"], rq["if_range"], rq["if_none_match"], rq["if_match"], rq["if_modified_since"],
                           rq["if_unmodified_since"], rq["accept_encoding"]),
                sp.Stored(ob["stored_length"], ob["content_type"], ob["etag"], ob["last_modified"], ob["content_encoding"], ob["raw_length"], ob["gunzip_enabled"]), 7)
    b = o.parse_plan(o.plan(o.Req(**rq), o.Obj(ob["stored_length"], ob["content_type"], ob["etag"], ob["last_modified"], ob["content_encoding"], ob["raw_length"],
                                                ob["gunzip_enabled"]), 7))
    c = o.parse_plan(h.plan_text(rq["method"], rq["http_minor"], rq["range"], rq["if_range"], rq["if_none_match"], rq["if_match"], rq["if_modified_since"],
                                 rq["if_unmodified_since"], rq["accept_encoding"], boundary=7, **ob))
    assert a.status == b.status == c.status and a.repr == b.repr == c.repr, (kw, a.status, b.status, c.status)
    return a.status, a.repr


def test_Z1_if_unmodified_since_without_a_comparable_date_is_412():
    assert _status3(if_unmodified_since="Wed, 07 Nov 2018 15:20:01 GMT")[0] == 200
    assert _status3(if_unmodified_since="Wed, 07 Nov 2018 15:20:00 GMT")[0] == 412
    assert _status3(if_unmodified_since="not a date")[0] == 412                      # unparseable = -1 < any stored date
    assert _status3(if_unmodified_since="Wed, 07 Nov 2018 15:20:01 GMT", last_modified=None)[0] == 412      # nothing stored to compare with
    assert _status3(if_unmodified_since="Wed, 07 Nov 2018 15:20:01 GMT", last_modified="junk")[0] == 412


def test_Z2_accept_encoding_gzip_is_nginx_gzip_ok():
    gz = dict(content_encoding="gzip", gunzip_enabled=True, raw_length=5000)
    for ae, inflated in [(None, True), ("identity", True), ("gzip", False), ("GZIP", False), ("deflate, gzip", False), ("gzip;q=0", True), ("gzip; q=0.000", True),
                         ("gzip;q=0.001", False), ("xgzip", True), ("gzipx", True), ("gz", True), ("gzip,", False), ("x gzip ; Q=1", False)]:
        assert _status3(accept_encoding=ae, **gz)[1] == ("inflated" if inflated else "identity"), ae


def test_Z3_host_with_a_space_is_a_name_and_a_slash_after_the_port_is_not():
    for fn in (sp.cache_key, o.cache_key, h.cache_key):
        assert fn("bad host", "/a") == b"bad host/download.tensorflow.org/a"
        assert fn("x:80/y", "/a") is None and fn("a..b", "/a") is None
        assert fn("[::1]:80", "/a") == b"[::1]/download.tensorflow.org/a"


def test_Z5_no_ranges_unless_the_upstream_response_said_accept_ranges():
    for ar, want in [(True, 206), (False, 200)]:
        a = sp.plan(sp.Request(range="bytes=0-9"), sp.Stored(1000, accept_ranges=ar))
        b = o.parse_plan(o.plan(o.Req(range="bytes=0-9"), o.Obj(1000, accept_ranges=ar)))
        c = o.parse_plan(h.plan_text(range="bytes=0-9", stored_length=1000, content_type="application/octet-stream", accept_ranges=ar))
        assert a.status == b.status == c.status == want
        for hdrs in (a.headers, b.headers, c.headers):
            assert (("Accept-Ranges", "bytes") in hdrs) == (ar and want == 200)
    full = sp.plan(sp.Request(), sp.Stored(1000, accept_ranges=False))
    assert full.status == 200 and all(k != "Accept-Ranges" for k, _ in full.headers)


def test_Z4_day_zero_parses_like_nginx():
    for fn in (sp.parse_http_time, o.parse_http_time, h.parse_http_date):
        assert fn("Wed, 00 Nov 2018 15:20:01 GMT") == fn("Wed, 31 Oct 2018 15:20:01 GMT") != -1


# ---- P: the request head -------------------------------------------------------------------------------------------
KEPT = ("accept-encoding", "host", "if-match", "if-modified-since", "if-none-match", "if-range", "if-unmodified-since", "range")


def _head_py(raw: bytes):
    r = sp.parse_head(raw)
    if r.error:
        return (r.error,)
    return (0, r.method, r.target, r.minor, r.keep_alive, r.has_body, tuple((k, r.fields[k]) for k in KEPT if k in r.fields))


def _head_text(t: str):
    lines = t.split("\n")
    err = int(lines[0].split(" ", 1)[1])
    if err:
        return (err,)
    d = {}
    fields = []
    for ln in lines[1:]:
        if not ln:
            continue
        k, _, v = ln.partition(" ")
        if k == "f":
            n, _, val = v.partition(": ")
            fields.append((n, val))
        else:
            d[k] = v
    return (0, d["method"], d["target"], int(d["minor"]), d["keep"] == "1", d["body"] == "1", tuple(fields))


REQ_LINES = ["GET / HTTP/1.1", "GET /a/b?x=1 HTTP/1.0", "HEAD /x HTTP/1.1", "GET /", "POST /", "get / HTTP/1.1", "GET  /two-spaces   HTTP/1.1  ", "GET / HTTP/1.2", "GET / HTTP/2.0",
             "GET / HTTP/1.10", "GET / HTTP/01.1", "GET / HTTP/1.", "GET / HTTP/1.100", "GET / http/1.1", "GET / HTTP/1.1 x", "GET http://yann.lecun.com/p HTTP/1.1",
             "GET yann.lecun.com/p HTTP/1.1", "GET ://x HTTP/1.1", "G_T / HTTP/1.1", "G-T / HTTP/1.1", "GET", "GET ", " GET / HTTP/1.1", "GET /a b HTTP/1.1", "PURGE / HTTP/1.1",
             "GET / HTTP/9.9", "GET / HTTP/10.0", "GET / HTTP/0.9", "GET /\t HTTP/1.1", "GET * HTTP/1.1"]
HDR_NAMES = ["Host", "host", "HOST", "Range", "range", "If-Range", "If-Match", "if-match", "If-None-Match", "If-Modified-Since", "If-Unmodified-Since", "Accept-Encoding",
             "Connection", "Content-Length", "Transfer-Encoding", "Expect", "Authorization", "X-Real-IP", "X_Underscore", "Host ", " Host", "", "Ra nge", "Räng", "User-Agent",
             "Content-Range", "Referer"]
HDR_VALUES = ["a", "A.B.", "www.cs.toronto.edu", "a..b", "", " ", "  bytes=0-1  ", "bytes=5-", "\tx", "x\t", "close", "Keep-Alive", "keep-alive, close", "Upgrade", "0", "00", "10", "1x", "-1",
              '"etag"', "Wed, 07 Nov 2018 15:20:01 GMT", "gzip", "chunked", "a/b", "a:80", "[::1]:8"]


@SET
@given(st.sampled_from(REQ_LINES), st.lists(st.tuples(st.sampled_from(HDR_NAMES), st.sampled_from([":", ": ", ":  ", " : ", ""]), st.sampled_from(HDR_VALUES)), max_size=7),
       st.sampled_from(["", "\r\n", "\r\n\r\n", "\n"]), st.sampled_from(["\r\n", "\r\n", "\r\n", "\n", "\r", "\r\n\0"]))
def test_request_head_three_way(rl, hdrs, lead, eol):
    raw = (lead + rl + eol + "".join(n + sep + v + "\r\n" for n, sep, v in hdrs)).encode("latin-1")
    a, b, c = _head_py(raw), _head_text(o.parse_head_text(raw)), _head_text(h.parse_head_text(raw))
    assert a == b == c, (raw, a, b, c)


def test_P_rules_by_name():
    def every(raw):
        a, b, c = _head_py(raw), _head_text(o.parse_head_text(raw)), _head_text(h.parse_head_text(raw))
        assert a == b == c, (raw, a, b, c)
        return a

    ok = every(b"GET /x HTTP/1.1\r\nHost: A\r\n")
    assert ok[:6] == (0, "GET", "/x", 1, True, False)
    assert every(b"get /x HTTP/1.1\r\nHost: a\r\n") == (400,)                                  # P1: lower-case method never reaches the 405 gate
    assert every(b"POST /x\r\n") == (400,) and every(b"GET /x\r\n")[3] == -1                     # P3
    assert every(b"GET / HTTP/2.0\r\nHost: a\r\n") == (505,)                                     # P4
    assert every(b"GET / HTTP/1.7\r\nHost: a\r\n")[3] == 1                                       # P4: any minor but 0 behaves as 1.1
    assert every(b"GET / HTTP/1.1\r\n") == (400,) and every(b"GET / HTTP/1.0\r\n")[0] == 0       # P8: Host required from 1.1 on
    r = every(b"GET / HTTP/1.1\r\nHost: first\r\nHost: second\r\nRange: bytes=0-0\r\nRange: bytes=5-5\r\n")
    assert dict(r[6]) == {"host": "first", "range": "bytes=0-0"}                                 # P6/P8: first line counts
    assert every(b"GET / HTTP/1.1\r\nHost: ok\r\nHost: a..b\r\n") == (400,)                       # P8: every Host line is validated
    for name in (b"If-Match", b"If-None-Match", b"If-Range", b"If-Modified-Since", b"If-Unmodified-Since", b"Content-Length"):
        assert every(b"GET / HTTP/1.1\r\nHost: a\r\n" + name + b": 1\r\n" + name.lower() + b": 1\r\n") == (400,)     # P7
    r = every(b"GET / HTTP/1.1\r\nHost: a\r\nX_Y: 1\r\nRange : bytes=0-0\r\n: v\r\n")
    assert dict(r[6]) == {"host": "a"}                                                            # P5: invalid names are ignored
    assert every(b"GET / HTTP/1.1\r\nHost: a\r\nConnection: close\r\nConnection: keep-alive\r\n")[4] is True          # P9: in order
    assert every(b"GET / HTTP/1.0\r\nConnection: Keep-Alive\r\n")[4] is True
    assert every(b"GET / HTTP/1.1\r\nHost: a\r\nContent-Length: 1x\r\n") == (400,)                # P10
    assert every(b"GET / HTTP/1.1\r\nHost: a\r\nContent-Length: 5\r\n")[5] is True
    assert every(b"\r\n\r\nGET / HTTP/1.1\r\nHost: a\r\n")[0] == 0                                # CR/LF before the request line
    assert every(b"GET / HTTP/1.1\r\nHost: a\0b\r\n") == (400,)
