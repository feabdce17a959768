"""-m "not gpu": two THIRD-PARTY implementations of RFC 7232/7233 as witnesses for the range filter and the
not-modified filter.

The three restatements in this repo (oracle/spec_py.py, oracle/nz_oracle.c, nezha_b200/host/nz_http.cc) were all
written here, from one person's readings of nginx; nginx itself cannot run in this image (parity unpinned,
oracle/SPEC.md).  Starlette's StaticFiles/FileResponse and aiohttp's FileResponse ship in the image and were written
by other people from the RFCs, so on every request where nginx follows the RFC they are independent witnesses of
what the answer must be: status, Content-Range, Content-Length and the body bytes (multipart bodies compared part by
part: boundaries and line endings differ).  Where nginx is KNOWN to differ from the RFC reading (or a library has a
quirk of its own) the case is listed under DIVERGENCES with the nginx rule it follows, and the test pins ours to
that rule while showing that the witnesses split - nothing is excluded silently.
"""
import asyncio
import os
import re

import pytest

from nezha_b200 import hostlib as h
from oracle import oracle as o, spec_py as sp

L = 10_000
DATA = bytes((i * 7 + 3) & 0xFF for i in range(L))
MTIME = 1541604001                                   # Wed, 07 Nov 2018 15:20:01 GMT
LM = "Wed, 07 Nov 2018 15:20:01 GMT"
LM_MINUS = "Wed, 07 Nov 2018 15:20:00 GMT"
LM_PLUS = "Wed, 07 Nov 2018 15:20:02 GMT"
CT = "application/octet-stream"


@pytest.fixture(scope="module")
def objdir(tmp_path_factory):
    d = tmp_path_factory.mktemp("witness")
    p = d / "obj.bin"
    p.write_bytes(DATA)
    os.utime(p, (MTIME, MTIME))
    return str(d)


def ours(method, hd, etag):
    """status, Content-Range, Content-Length, body (bytes, or [(content-range, bytes)] for multipart) - from all three
    restatements, which must agree before anybody else is asked."""
    args = dict(range=hd.get("Range"), if_range=hd.get("If-Range"), if_none_match=hd.get("If-None-Match"), if_match=hd.get("If-Match"),
                if_modified_since=hd.get("If-Modified-Since"), if_unmodified_since=hd.get("If-Unmodified-Since"))
    a = sp.plan(sp.Request(method, 1, args["range"], args["if_range"], args["if_none_match"], args["if_match"], args["if_modified_since"],
                           args["if_unmodified_since"], None), sp.Stored(L, CT, etag, LM, None, -1, False, True), 7)
    b = o.parse_plan(o.plan(o.Req(method, 1, args["range"], args["if_range"], args["if_none_match"], args["if_match"], args["if_modified_since"],
                                  args["if_unmodified_since"], None), o.Obj(L, CT, etag, LM, None, -1, False, True), 7))
    c = o.parse_plan(h.plan_text(method, 1, args["range"], args["if_range"], args["if_none_match"], args["if_match"], args["if_modified_since"],
                                 args["if_unmodified_since"], None, stored_length=L, raw_length=-1, content_type=CT, etag=etag, last_modified=LM,
                                 content_encoding=None, gunzip_enabled=False, boundary=7, accept_ranges=True))

    def norm(p):
        hdrs = {k.lower(): v for k, v in p.headers}
        body = b"" if p.head else b"".join(DATA[v[0]:v[0] + v[1]] if k == "slice" else bytes(v) for k, v in p.body)
        return p.status, hdrs.get("content-range"), hdrs.get("content-length"), hdrs.get("content-type", ""), body

    na, nb, nc = norm(a), norm(b), norm(c)
    assert na == nb == nc, (method, hd, na[:4], nb[:4], nc[:4])
    return na


def parts_of(ctype, body):
    """multipart/byteranges -> [(content-range, payload)], whatever the boundary token and the line endings"""
    m = re.search(r'boundary="?([^";]+)"?', ctype)
    assert m, ctype
    bnd = m.group(1).encode()
    out = []
    for chunk in body.split(b"--" + bnd)[1:]:
        if chunk.startswith(b"--"):
            break
        chunk = chunk.lstrip(b"\r\n")
        sep = b"\r\n\r\n" if b"\r\n\r\n" in chunk else b"\n\n"
        head, _, payload = chunk.partition(sep)
        cr = re.search(rb"(?i)content-range:\s*([^\r\n]+)", head).group(1).decode().strip()
        a, b_ = map(int, re.match(r"bytes (\d+)-(\d+)/", cr).groups())
        out.append((cr, payload[:b_ - a + 1]))
    return out


class Starlette:
    def __init__(self, objdir):
        from starlette.applications import Starlette as App
        from starlette.staticfiles import StaticFiles
        from starlette.testclient import TestClient

        app = App()
        app.mount("/", StaticFiles(directory=objdir))
        self.c = TestClient(app)
        self.etag = self.c.get("/obj.bin").headers["etag"]

    def ask(self, method, hd):
        r = self.c.request(method, "/obj.bin", headers=hd)
        return r.status_code, r.headers.get("content-range"), r.headers.get("content-length"), r.headers.get("content-type", ""), r.content


class Aiohttp:
    def __init__(self, objdir):
        self.path = os.path.join(objdir, "obj.bin")
        self.etag = self.ask("GET", {})[5]

    def ask(self, method, hd):
        from aiohttp import web
        from aiohttp.test_utils import TestClient, TestServer

        async def run():
            app = web.Application()

            async def hnd(_req):
                return web.FileResponse(self.path)

            app.router.add_route("*", "/obj.bin", hnd)
            async with TestClient(TestServer(app)) as c:
                r = await c.request(method, "/obj.bin", headers=hd)
                body = await r.read()
                return r.status, r.headers.get("Content-Range"), r.headers.get("Content-Length"), r.headers.get("Content-Type", ""), body, r.headers.get("ETag")

        return asyncio.run(run())


@pytest.fixture(scope="module")
def star(objdir):
    pytest.importorskip("starlette")
    pytest.importorskip("httpx")
    return Starlette(objdir)


@pytest.fixture(scope="module")
def aio(objdir):
    pytest.importorskip("aiohttp")
    return Aiohttp(objdir)


def test_validators_are_what_the_witnesses_say(star, aio):
    s = star.ask("GET", {})
    a = aio.ask("GET", {})
    assert s[0] == a[0] == 200 and s[4] == a[4] == DATA
    assert star.c.get("/obj.bin").headers["last-modified"] == LM          # same Last-Modified string as our stored meta


# ---- requests where nginx follows the RFC: every witness must say what we say ----------------------------------
AGREED = [
    ("GET", {}),
    ("HEAD", {}),
    ("GET", {"Range": "bytes=0-0"}),
    ("GET", {"Range": "bytes=0-9"}),
    ("GET", {"Range": "bytes=5-"}),
    ("GET", {"Range": "bytes=-1"}),
    ("GET", {"Range": "bytes=-9999"}),
    ("GET", {"Range": "bytes=9999-9999"}),
    ("GET", {"Range": "bytes=9999-20000"}),              # last-byte-pos beyond the end is clamped (RFC 7233 2.1)
    ("GET", {"Range": "bytes=0-9999"}),
    ("GET", {"Range": "bytes=0-"}),
    ("GET", {"Range": "bytes=1234-5678"}),
    ("GET", {"Range": "bytes=10000-"}),                  # first-byte-pos == length: 416, Content-Range: bytes */L
    ("GET", {"Range": "bytes=20000-30000"}),
    ("HEAD", {"Range": "bytes=0-9"}),
    ("GET", {"Range": "bytes=0-9", "If-Range": "ETAG"}),
    ("GET", {"Range": "bytes=0-9", "If-Range": LM}),
    ("GET", {"If-None-Match": "ETAG"}),
    ("GET", {"If-None-Match": '"something-else"'}),
    ("GET", {"If-None-Match": '"a", ETAG, "b"'}),
    ("GET", {"If-Modified-Since": LM}),
    ("GET", {"If-Modified-Since": LM_MINUS}),
    ("GET", {"If-None-Match": '"something-else"', "If-Modified-Since": LM}),     # If-None-Match wins over If-Modified-Since (RFC 7232 3.3 / 6)
    ("HEAD", {"If-None-Match": "ETAG"}),
]
# aiohttp alone also implements these (Starlette's StaticFiles has no If-Match / If-Unmodified-Since / "*")
AGREED_AIO = [
    ("GET", {"If-None-Match": "*"}),
    ("GET", {"If-None-Match": "W/ETAG"}),                # weak comparison for If-None-Match
    ("GET", {"If-Match": "ETAG"}),
    ("GET", {"If-Match": '"something-else"'}),
    ("GET", {"If-Match": "*"}),
    ("GET", {"If-Unmodified-Since": LM}),
    ("GET", {"If-Unmodified-Since": LM_MINUS}),
    ("GET", {"If-Unmodified-Since": LM_PLUS}),
    ("GET", {"Range": "bytes=-20000"}),                  # suffix longer than the object: the whole object as 206 (RFC 7233 2.1)
    ("GET", {"Range": "bytes=10-5"}),                    # last < first: not a satisfiable spec -> 416
]
# Starlette alone answers several ranges (aiohttp: 416 for any list)
MULTI = [
    {"Range": "bytes=0-9,20-29"},
    {"Range": "bytes=0-0,9999-9999"},
    {"Range": "bytes=100-199,300-399,-50"},
    {"Range": "bytes=5000-5999,7000-"},
]


def with_etag(hd, etag):
    return {k: v.replace("ETAG", etag) for k, v in hd.items()}


def same_answer(mine, theirs, who, case):
    st, cr, cl, _ct, body = mine
    assert theirs[0] == st, (who, case, theirs[:3], mine[:3])
    if st in (200, 206):
        assert theirs[4] == body, (who, case)
        assert (theirs[1] or None) == cr, (who, case, theirs[1], cr)
        assert theirs[2] == cl, (who, case, theirs[2], cl)
    elif st == 416:
        assert (theirs[1] or "").replace("bytes ", "") == (cr or "").replace("bytes ", "") == f"*/{L}", (who, case, theirs[1], cr)
    else:
        assert theirs[4] == b"", (who, case)


@pytest.mark.parametrize("method,hd", AGREED, ids=lambda v: str(v))
def test_both_witnesses_agree_with_us(star, aio, method, hd):
    same_answer(ours(method, with_etag(hd, star.etag), star.etag), star.ask(method, with_etag(hd, star.etag)), "starlette", (method, hd))
    same_answer(ours(method, with_etag(hd, aio.etag), aio.etag), aio.ask(method, with_etag(hd, aio.etag)), "aiohttp", (method, hd))


@pytest.mark.parametrize("method,hd", AGREED_AIO, ids=lambda v: str(v))
def test_aiohttp_agrees_with_us(aio, method, hd):
    same_answer(ours(method, with_etag(hd, aio.etag), aio.etag), aio.ask(method, with_etag(hd, aio.etag)), "aiohttp", (method, hd))


@pytest.mark.parametrize("hd", MULTI, ids=lambda v: str(v))
def test_starlette_multipart_parts_equal_ours(star, hd):
    st, cr, _cl, ct, body = ours("GET", hd, star.etag)
    s = star.ask("GET", hd)
    assert st == s[0] == 206 and cr is None
    # Starlette 0.52 announces the boundary in Content-Range (and leaves Content-Type alone): take it where it is
    s_ct = s[3] if "boundary=" in s[3] else (s[1] or "")
    assert ct.startswith("multipart/byteranges; boundary=") and s_ct.startswith("multipart/byteranges; boundary=")
    mine, theirs = parts_of(ct, body), parts_of(s_ct, s[4])
    assert mine == theirs, (hd, [c for c, _ in mine], [c for c, _ in theirs])
    for cr_, payload in mine:
        a, b = map(int, re.match(r"bytes (\d+)-(\d+)/10000$", cr_).groups())
        assert payload == DATA[a:b + 1]


# ---- where nginx (hence ours) and a witness part ways: stated, not hidden --------------------------------------
def test_divergence_if_modified_since_is_exact_in_nginx(star, aio):
    """nginx's default `if_modified_since exact` (the reference's conf does not set it): 304 only when the date EQUALS
    Last-Modified (ngx_http_not_modified_filter_module: ngx_http_test_if_modified).  The RFC reading - and both
    witnesses - answer 304 for any date at or after it.  SPEC C5."""
    hd = {"If-Modified-Since": LM_PLUS}
    assert ours("GET", hd, star.etag)[0] == 200
    assert star.ask("GET", hd)[0] == 304 and aio.ask("GET", hd)[0] == 304


def test_divergence_if_range_with_another_etag(star, aio):
    """If-Range carrying a different validator: the whole object, 200 (RFC 7233 3.2; nginx range header filter).
    Starlette agrees; aiohttp only understands dates in If-Range and still answers 206."""
    hd = {"Range": "bytes=0-9", "If-Range": '"another"'}
    mine = ours("GET", hd, star.etag)
    assert mine[0] == 200 and mine[4] == DATA
    same_answer(mine, star.ask("GET", hd), "starlette", hd)
    assert aio.ask("GET", hd)[0] == 206


def test_divergence_if_range_date_must_match_exactly(star, aio):
    """If-Range with a date is an exact match in nginx (`if_range_time != last_modified_time` -> full response) and in
    RFC 7233 3.2; Starlette says the same; aiohttp treats it like If-Unmodified-Since (later date still 206)."""
    hd = {"Range": "bytes=0-9", "If-Range": LM_PLUS}
    mine = ours("GET", hd, star.etag)
    assert mine[0] == 200 and mine[4] == DATA
    same_answer(mine, star.ask("GET", hd), "starlette", hd)
    assert aio.ask("GET", hd)[0] == 206


def test_divergence_suffix_longer_than_the_object(star, aio):
    """`bytes=-20000` of 10 000: nginx starts at 0 (ngx_http_range_parse: suffix > length -> start = 0) = RFC 7233 2.1;
    aiohttp agrees (AGREED_AIO); Starlette 0.52 answers 416."""
    hd = {"Range": "bytes=-20000"}
    mine = ours("GET", hd, star.etag)
    assert mine[:2] == (206, f"bytes 0-{L - 1}/{L}") and mine[4] == DATA
    assert star.ask("GET", hd)[0] == 416


def test_divergence_empty_suffix_and_reversed_spec(star, aio):
    """`bytes=-0` selects nothing and `bytes=10-5` is not a valid spec: with no satisfiable range left nginx answers 416
    (ngx_http_range_parse returns NGX_HTTP_RANGE_NOT_SATISFIABLE; SPEC R5/R6).  Starlette: 416 for the first, 400 for the second;
    aiohttp: the whole object as 206 for the first, 416 for the second - each witness backs one of the two."""
    assert ours("GET", {"Range": "bytes=-0"}, star.etag)[0] == 416
    assert ours("GET", {"Range": "bytes=10-5"}, star.etag)[0] == 416
    assert star.ask("GET", {"Range": "bytes=-0"})[0] == 416 and aio.ask("GET", {"Range": "bytes=10-5"})[0] == 416
    assert star.ask("GET", {"Range": "bytes=10-5"})[0] == 400 and aio.ask("GET", {"Range": "bytes=-0"})[0] == 206


def test_divergence_overlapping_ranges_are_not_merged_by_nginx(star):
    """nginx answers the parts as asked, in order, and only falls back to the whole object when their sum exceeds it
    (ngx_http_range_parse: `size > content_length` -> 200); Starlette merges overlapping ranges into one.  SPEC R7/R9."""
    hd = {"Range": "bytes=0-99,50-149"}
    st, _cr, _cl, ct, body = ours("GET", hd, star.etag)
    assert st == 206 and [c for c, _ in parts_of(ct, body)] == [f"bytes 0-99/{L}", f"bytes 50-149/{L}"]
    s = star.ask("GET", hd)
    assert s[0] == 206 and s[1] == f"bytes 0-149/{L}"


def test_divergence_both_validators_must_say_not_modified_in_nginx(star, aio):
    """If-None-Match matches but If-Modified-Since is an earlier date: RFC 7232 (6) lets If-None-Match decide -> 304, and
    so do both witnesses; nginx's not-modified filter needs EVERY validator present to say "not modified"
    (`if_modified_since && ngx_http_test_if_modified(r)` -> next filter, before If-None-Match is looked at) -> 200.  SPEC C3(a)."""
    for w in (star, aio):
        hd = {"If-None-Match": w.etag, "If-Modified-Since": LM_MINUS}
        mine = ours("GET", hd, w.etag)
        assert mine[0] == 200 and mine[4] == DATA
        assert w.ask("GET", hd)[0] == 304
