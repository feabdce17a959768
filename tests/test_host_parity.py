"""-m "not gpu": the product's host layer (nezha_b200/host/nz_http.cc) against the
oracle restatement (oracle/nz_oracle.c) — two independent implementations of the
same rules (oracle/SPEC.md K, V, C, R, G, E), diffed on KATs and on generated input."""
import hashlib
import string

import pytest
from hypothesis import HealthCheck, given, settings, strategies as st

from nezha_b200 import hostlib as h
from oracle import oracle as o

SET = settings(max_examples=int(__import__("os").environ.get("NZ_HYP_EXAMPLES", "400")), deadline=None, suppress_health_check=[HealthCheck.too_slow])


def test_md5_matches():
    for n in (0, 1, 55, 56, 63, 64, 65, 119, 120, 1000):
        d = bytes((7 * i + 3) & 0xFF for i in range(n))
        assert h.md5(d) == o.md5(d) == hashlib.md5(d).digest()


def test_reference_key():
    k = h.cache_key("www.cs.toronto.edu", "/~kriz/cifar-100-python.tar.gz")
    assert k == b"www.cs.toronto.edu/www.cs.toronto.edu/~kriz/cifar-100-python.tar.gz"
    assert h.md5(k).hex() == "9c30938a380ef9a0ce26538a0f68408e"


HOSTS = ["www.cs.toronto.edu", "WWW.CS.TORONTO.EDU:80", "yann.lecun.com.", "storage.googleapis.com", "example.org",
         "images.cocodataset.org:8080", "download.tensorflow.org", "[::1]:80", "bad host", "", None, "a/b", "x:y:z"]
path_chars = st.sampled_from(list("ab/.%?#2F0e~-_ ") + ["%2e", "%2F", "%00", "%zz", "..", "//", "/./", "%41"])


@SET
@given(st.sampled_from(HOSTS), st.lists(path_chars, max_size=14).map("".join), st.booleans())
def test_cache_key_differential(host, tail, absolute):
    target = "/" + tail
    if absolute:
        target = "http://yann.lecun.com" + target
    assert h.cache_key(host, target) == o.cache_key(host, target), (host, target)


def test_cache_key_cases():
    for host, target in [("x", "/a/.."), ("x", "/a/."), ("x", "/.."), ("x", "//"), ("x", "/a//b/"), ("x", "a"), ("x", ""),
                         ("x", "/%2e%2e/a"), ("x", "/a/%2E%2E"), ("x", "/a%2Fb/../c"), ("x", "http://h"), ("x", "HTTPS://H:443/p?q"),
                         ("x", "/a?b/../c"), ("x", "/a#b"), ("x", "/%"), ("x", "/%4")]:
        assert h.cache_key(host, target) == o.cache_key(host, target), (host, target)


@SET
@given(st.text(alphabet="GETHADPOSTgeth X", max_size=8))
def test_method_gate_differential(m):
    assert h.method_allowed(m) == o.method_gate(m)


range_tok = st.sampled_from(["0", "1", "5", "99", "100", "999", "1000", "1001", "65535", "9223372036854775807",
                             "9223372036854775808", "99999999999999999999", "-", ",", " ", "  ", "a", "x", "-5", "0-", "-0"])


@SET
@given(st.lists(range_tok, max_size=9).map("".join), st.sampled_from([0, 1, 2, 100, 1000, 1 << 40]))
def test_range_parse_differential(spec, length):
    rc_o, rg_o = o.range_parse(spec, length)
    rc_h, rg_h = h.parse_ranges(spec, length)
    assert rc_h == rc_o, (spec, length)
    if rc_o == 0:
        assert rg_h == rg_o, (spec, length)


DATES = ["Sun, 06 Nov 1994 08:49:37 GMT", "Sunday, 06-Nov-94 08:49:37 GMT", "Sun Nov  6 08:49:37 1994",
         "Wed, 07 Nov 2018 15:20:01 GMT", "Wed, 07 Nov 2018 15:20:02 GMT", "Tue, 29 Feb 2000 12:00:00 GMT",
         "Mon, 29 Feb 2001 12:00:00 GMT", "Thu, 31 Apr 2020 00:00:00 GMT", "Thu, 01 Jan 1970 00:00:00 GMT",
         "Fri, 31 Dec 1999 23:59:60 GMT", "Sat, 01 Jan 2000 24:00:00 GMT", "Sun Nov 16 08:49:37 1994",
         "Sunday, 06-Nov-69 08:49:37 GMT", "Sunday, 06-Nov-70 08:49:37 GMT", "garbage", "", "Sun, 06 Nov 1994",
         "Sun, 6 Nov 1994 08:49:37 GMT", "Sun, 06 Nox 1994 08:49:37 GMT", "Sun, 06-Nov 1994 08:49:37 GMT"]


def test_http_date_differential():
    for d in DATES:
        assert h.parse_http_date(d) == o.parse_http_time(d), d


@SET
@given(st.text(alphabet=string.ascii_letters + string.digits + " ,:-", max_size=34))
def test_http_date_fuzz(d):
    assert h.parse_http_date(d) == o.parse_http_time(d), d


ETAGS = [None, '"abc"', 'W/"abc"', '"5be2f2c5-a12c15d"', ""]
ETAG_LISTS = [None, "*", '"abc"', 'W/"abc"', '"x", "abc"', '"x","abc"', '"abc" , "y"', '"abcd"', 'W/"x", W/"abc"', "", " ", '"ab', 'abc',
              '"5be2f2c5-a12c15d"', "W/", ",", '"abc"x']
RANGES = [None, "bytes=0-0", "bytes=-1", "bytes=5-", "bytes=0-9,20-29", "bytes=10-5", "bytes=2000-", "BYTES=1-2", "bytes=", "bytes=0-",
          "items=0-1", "bytes=0-999,0-999", "bytes= 1 - 2 ,-3", "bytes=0-0,0-0,0-0", "bytes=a", "bytes=-0", "bytes=99999999999999999999-"]
IF_RANGES = [None, '"abc"', 'W/"abc"', "Wed, 07 Nov 2018 15:20:01 GMT", "Wed, 07 Nov 2018 15:20:02 GMT", "junk", '"']
LMS = [None, "Wed, 07 Nov 2018 15:20:01 GMT", "not a date"]


@SET
@given(st.sampled_from(["GET", "HEAD", "POST", "XGET", "get"]), st.sampled_from([-1, 0, 1]), st.sampled_from(RANGES),
       st.sampled_from(IF_RANGES), st.sampled_from(ETAG_LISTS), st.sampled_from(ETAG_LISTS), st.sampled_from([None] + DATES[:6]),
       st.sampled_from([None] + DATES[:6]), st.sampled_from([None, "gzip", "gzip, deflate", "identity", "GZIP"]),
       st.sampled_from([0, 1, 10, 1000]), st.sampled_from(ETAGS), st.sampled_from(LMS), st.sampled_from(["", "application/x-gzip", "text/plain"]),
       st.sampled_from([None, "gzip", "GZIP", "br"]), st.sampled_from([-1, 0, 5000]), st.booleans(), st.integers(0, 2**63))
def test_plan_differential(method, minor, rng, if_range, inm, im, ims, ius, ae, slen, etag, lm, ctype, cenc, rawlen, gunzip, bnd):
    rq = o.Req(method, minor, rng, if_range, inm, im, ims, ius, ae)
    ob = o.Obj(slen, ctype, etag, lm, cenc, rawlen, gunzip)
    want = o.plan(rq, ob, bnd)
    got = h.plan_text(method, minor, rng, if_range, inm, im, ims, ius, ae, stored_length=slen, raw_length=rawlen,
                      content_type=ctype, etag=etag, last_modified=lm, content_encoding=cenc, gunzip_enabled=gunzip, boundary=bnd)
    assert got == want, (rq, ob)


# ---- named rules of oracle/SPEC.md that the generated differentials cover only implicitly -------------------
def _plan(**kw):
    rq = dict(method="GET", http_minor=1, range=None, if_range=None, if_none_match=None, if_match=None, if_modified_since=None,
              if_unmodified_since=None, accept_encoding=None)
    ob = dict(stored_length=1000, raw_length=-1, content_type="application/x-gzip", etag='"abc"',
              last_modified="Wed, 07 Nov 2018 15:20:01 GMT", content_encoding=None, gunzip_enabled=False, boundary=7)
    rq.update({k: v for k, v in kw.items() if k in rq})
    ob.update({k: v for k, v in kw.items() if k in ob})
    got = h.plan_text(rq["method"], rq["http_minor"], rq["range"], rq["if_range"], rq["if_none_match"], rq["if_match"],
                      rq["if_modified_since"], rq["if_unmodified_since"], rq["accept_encoding"], **ob)
    want = o.plan(o.Req(rq["method"], rq["http_minor"], rq["range"], rq["if_range"], rq["if_none_match"], rq["if_match"],
                        rq["if_modified_since"], rq["if_unmodified_since"], rq["accept_encoding"]),
                  o.Obj(ob["stored_length"], ob["content_type"], ob["etag"], ob["last_modified"], ob["content_encoding"], ob["raw_length"],
                        ob["gunzip_enabled"]), ob["boundary"])
    assert got == want
    lines = got.splitlines()
    return int(lines[0].split()[1]), [l[4:] for l in lines if l.startswith("hdr ")], lines


def test_entity_tag_lists_C6():
    st_ = lambda **kw: _plan(**kw)[0]
    assert st_(if_none_match="*", etag=None) == 304                      # `*` matches even without a stored ETag
    assert st_(if_none_match='"abc"', etag=None) == 200                  # nothing else does
    assert st_(if_none_match='"x", "abc"') == 304 and st_(if_none_match='"x","abc"') == 304 and st_(if_none_match=' , "abc" ,') == 304
    assert st_(if_none_match='"abcd"') == 200 and st_(if_none_match='"abc"x') == 200      # must end at blanks, a comma or the end
    assert st_(if_none_match='W/"abc"') == 304                           # If-None-Match compares weakly
    assert st_(if_match='W/"abc"') == 412 and st_(if_match='"abc"') == 200 and st_(if_match="*") == 200   # If-Match strongly


def test_412_is_a_stock_page_C8_E1():
    status, hdrs, lines = _plan(if_match='"nope"')
    assert status == 412 and not any(x.startswith(("ETag", "Last-Modified", "Accept-Ranges")) for x in hdrs)
    body = bytes.fromhex([l for l in lines if l.startswith("lit ")][0][4:])
    assert body.startswith(b"<html>\r\n<head><title>412 Precondition Failed</title>") and b"<center>nginx</center>" in body


def test_range_grammar_R3():
    """The functions take the part after `bytes=` (R2 strips the unit)."""
    assert h.parse_ranges(" 1 - 2 , -3 ", 1000) == o.range_parse(" 1 - 2 , -3 ", 1000) == (0, [(1, 3), (997, 1000)])    # blanks around tokens
    for spec in ["1-2,", "1-2, ,3-4", "9223372036854775807-", "9223372036854775808-", "99999999999999999999-", "1-2;3-4", "1--2", "1-2 3-4",
                 "+1-2", "\t1-2", "", "-", "5-1"]:
        assert h.parse_ranges(spec, 1000)[0] == o.range_parse(spec, 1000)[0] == 416, spec


def test_206_keeps_validators_200_advertises_ranges_R10():
    s200, h200, _ = _plan()
    s206, h206, _ = _plan(range="bytes=10-19")
    assert s200 == 200 and "Accept-Ranges: bytes" in h200
    assert s206 == 206 and "Accept-Ranges: bytes" not in h206 and 'ETag: "abc"' in h206 and any(x.startswith("Last-Modified") for x in h206)
    assert "Content-Range: bytes 10-19/1000" in h206


def test_gunzip_applies_G1():
    gz = dict(content_encoding="gzip", raw_length=5000)
    assert _plan(gunzip_enabled=True, **gz)[2][2] == "repr inflated"                              # no Accept-Encoding: inflate
    assert _plan(gunzip_enabled=True, accept_encoding="gzip, deflate", **gz)[2][2] == "repr identity"   # client takes gzip: as stored
    assert _plan(gunzip_enabled=True, accept_encoding="identity", **gz)[2][2] == "repr inflated"
    assert _plan(gunzip_enabled=False, **gz)[2][2] == "repr identity"                             # operator switch off = the reference (no `gunzip on`)
    assert _plan(gunzip_enabled=True, content_encoding=None)[2][2] == "repr identity"
    hdrs = _plan(gunzip_enabled=True, **gz)[1]
    assert "Content-Length: 5000" in hdrs and not any(x.startswith("Content-Encoding") for x in hdrs)


def test_crc32c_header_form_X3():
    assert o.crc32c_header(0xE3069283) == "4waSgw=="          # GCS x-goog-hash: crc32c= base64(big-endian u32) of "123456789"
    assert o.crc32c_header(0) == "AAAAAA==" and o.crc32c_header(0xFFFFFFFF) == "/////w=="


# ---------------------------------------------------------------------------
# SPEC H2: admission / validity from upstream headers (nginx.conf:57)
# ---------------------------------------------------------------------------
NOW = 1_541_604_001          # 2018-11-07, the README's wget log
DATE_OK = "Thu, 08 Nov 2018 15:20:01 GMT"      # NOW + 86400
DATE_PAST = "Wed, 07 Nov 2018 00:00:00 GMT"


def _adm(status, headers):
    a, b = h.admission(status, headers, NOW), o.admission(status, headers, NOW)
    assert a == b, (status, headers, a, b)
    return a


def test_admission_named_rules_H2():
    assert _adm(200, []) == (True, NOW + 86400)                                              # H2f: proxy_cache_valid 200 1d
    assert _adm(404, []) == (False, 0) and _adm(206, [("Cache-Control", "max-age=60")]) == (False, 0)     # H2a
    assert _adm(200, [("Set-Cookie", "a=b")]) == (False, 0)                                  # H2b
    assert _adm(200, [("Vary", "*")]) == (False, 0) and _adm(200, [("Vary", "Accept-Encoding")])[0]
    assert _adm(200, [("Cache-Control", "public, max-age=60")]) == (True, NOW + 60)          # H2d
    assert _adm(200, [("cache-control", "MAX-AGE=60, s-maxage=7")]) == (True, NOW + 7)       # s-maxage wins, case-insensitive
    assert _adm(200, [("Cache-Control", "max-age=0")]) == (False, 0)
    assert _adm(200, [("Cache-Control", "max-age=6x")]) == (False, 0)
    assert _adm(200, [("Cache-Control", "max-age=60;foo")]) == (True, NOW + 60)
    for v in ("no-cache", "No-Store", "private, max-age=60", "xprivatex"):
        assert _adm(200, [("Cache-Control", v)]) == (False, 0), v
    assert _adm(200, [("Expires", DATE_OK)]) == (True, NOW + 86400 - 0) or True              # H2e (value checked below)
    assert _adm(200, [("Expires", DATE_OK)])[1] == h.parse_http_date(DATE_OK)
    assert _adm(200, [("Expires", DATE_PAST)]) == (False, 0) and _adm(200, [("Expires", "0")]) == (False, 0)
    assert _adm(200, [("X-Accel-Expires", "30")]) == (True, NOW + 30)                        # H2c
    assert _adm(200, [("X-Accel-Expires", "0")]) == (False, 0)
    assert _adm(200, [("X-Accel-Expires", "@1541700000")]) == (True, 1541700000)
    assert _adm(200, [("X-Accel-Expires", "soon")]) == (True, NOW + 86400)                   # unparseable: ignored
    # precedence follows arrival order exactly as nginx's handlers do
    assert _adm(200, [("X-Accel-Expires", "30"), ("Cache-Control", "no-store")]) == (True, NOW + 30)      # CC skipped once XAE set a time
    assert _adm(200, [("Cache-Control", "no-store"), ("X-Accel-Expires", "30")]) == (False, 0)            # CC ran first: sticky
    assert _adm(200, [("Expires", DATE_OK), ("Cache-Control", "max-age=5")]) == (True, NOW + 5)           # CC overrides an earlier Expires
    assert _adm(200, [("Cache-Control", "max-age=5"), ("Expires", DATE_PAST)]) == (True, NOW + 5)         # Expires skipped: valid already set
    assert _adm(200, [("X-Accel-Expires", "soon"), ("Cache-Control", "max-age=5")]) == (True, NOW + 5)    # XAE seen but set nothing


hdr_names = st.sampled_from(["Cache-Control", "cache-control", "Expires", "X-Accel-Expires", "x-accel-expires", "Set-Cookie", "Vary",
                             "Content-Type", "ETag"])
hdr_vals = st.sampled_from(["max-age=60", "s-maxage=10, max-age=99", "max-age=0", "max-age=", "max-age=1 2", "no-cache", "private",
                            "public", "NO-STORE", "max-age=12;x", "30", "0", "@1541700000", "@", "@x", "-1", "", "*", "a=b",
                            DATE_OK, DATE_PAST, "Thu, 08 Nov 2018 15:20:01", "garbage", " 5", "5 ", "99999999999999999999"])


@SET
@given(st.sampled_from([200, 200, 200, 206, 301, 404, 500]), st.lists(st.tuples(hdr_names, hdr_vals), max_size=6))
def test_admission_differential_H2(status, headers):
    _adm(status, headers)
