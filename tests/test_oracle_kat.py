"""Pins the oracle (oracle/nz_oracle.c) before anything trusts it.

The reference ships no tests for this path ("parity
unpinned"), so the pins are (a) the only vectors the reference itself fixes —
the key template, the demo URL, the object length — and (b) published KATs of
the algorithms nginx's directives invoke (RFC 1321, RFC 3720 B.4, zlib), each
cross-checked against an independent implementation in the Python stdlib.
Rule ids refer to oracle/SPEC.md.
"""
import hashlib
import os
import random
import zlib

import numpy as np
import pytest

from oracle import oracle as o


# ---- K: key + identity ---------------------------------------------------
def test_reference_demo_key_K7():
    # /root/reference/examples/demo/kubeflow-test.yaml:22 + nginx.conf:94
    key = o.cache_key("www.cs.toronto.edu", "/~kriz/cifar-100-python.tar.gz")
    assert key == b"www.cs.toronto.edu/www.cs.toronto.edu/~kriz/cifar-100-python.tar.gz"
    assert len(key) == 67
    assert o.md5(key).hex() == "9c30938a380ef9a0ce26538a0f68408e"
    assert o.crc32_ieee(key) == 0xA0FD5DC9


@pytest.mark.parametrize("n", [0, 1, 55, 56, 57, 63, 64, 65, 119, 120, 121, 1000, 4097])
def test_md5_vs_hashlib(n):
    d = bytes((i * 131 + 7) & 0xFF for i in range(n))
    assert o.md5(d) == hashlib.md5(d).digest()


def test_md5_rfc1321_suite():
    kat = {
        b"": "d41d8cd98f00b204e9800998ecf8427e",
        b"a": "0cc175b9c0f1b6a831c399e269772661",
        b"abc": "900150983cd24fb0d6963f7d28e17f72",
        b"message digest": "f96b697d7cb7938d525a2f31aaf161d0",
        b"abcdefghijklmnopqrstuvwxyz": "c3fcd3d76192e4007dfb496cca67e13b",
    }
    for m, h in kat.items():
        assert o.md5(m).hex() == h


def test_crc32_ieee_vs_zlib():
    assert o.crc32_ieee(b"123456789") == 0xCBF43926
    for n in (0, 1, 7, 8, 9, 1000):
        d = os.urandom(n)
        assert o.crc32_ieee(d) == zlib.crc32(d)


def test_key_rules_K2_K4():
    ck = o.cache_key
    assert ck("WWW.CS.Toronto.EDU:80", "/a") == b"www.cs.toronto.edu/www.cs.toronto.edu/a"
    assert ck("yann.lecun.com.", "/exdb/mnist/") == b"yann.lecun.com/yann.lecun.com/exdb/mnist/"
    # query string is NOT part of the key (nginx.conf:56 uses $uri)
    assert ck("storage.googleapis.com", "/b/o?alt=media&sig=1") == ck("storage.googleapis.com", "/b/o")
    # percent-decoding, dot segments, slash merging
    assert ck("images.cocodataset.org", "/zips//./x/../val%32017.zip") == b"images.cocodataset.org/images.cocodataset.org/zips/val2017.zip"
    assert ck("images.cocodataset.org", "/a/..") .endswith(b"org/")
    assert ck("images.cocodataset.org", "/a/.").endswith(b"org/a/")
    assert ck("images.cocodataset.org", "/../etc") is None
    assert ck("images.cocodataset.org", "/a%00b") is None
    assert ck("images.cocodataset.org", "/a%zz") is None
    # absolute-form target: host from the request line wins
    assert ck("ignored", "http://yann.lecun.com/x?q") == b"yann.lecun.com/yann.lecun.com/x"
    # unknown Host -> first server block's proxy_host (SPEC V1)
    assert ck("example.org", "/f") == b"example.org/download.tensorflow.org/f"


def test_method_gate_V2():
    assert o.method_gate("GET") and o.method_gate("HEAD")
    assert o.method_gate("XGETX")          # unanchored regex, nginx.conf:47
    assert not o.method_gate("get") and not o.method_gate("POST") and not o.method_gate("PUT")


# ---- X: CRC-32C -----------------------------------------------------------
def test_crc32c_rfc3720():
    assert o.crc32c(b"123456789") == 0xE3069283
    assert o.crc32c(bytes(32)) == 0x8A9136AA
    assert o.crc32c(b"\xff" * 32) == 0x62A8AB43
    assert o.crc32c(bytes(range(32))) == 0x46DD794E
    assert o.crc32c(bytes(range(31, -1, -1))) == 0x113FDB5C
    assert o.crc32c_header(0xE3069283) == "4waSgw=="
    assert o.crc32c(b"") == 0


def test_crc32c_table_vs_bitwise():
    rnd = random.Random(5)
    for n in [0, 1, 3, 7, 8, 9, 15, 16, 17, 255, 256, 4095, 70001]:
        d = bytes(rnd.getrandbits(8) for _ in range(n))
        assert o.crc32c(d) == o.crc32c_bitwise(d)
        # running form
        k = n // 3
        assert o.crc32c(d[k:], o.crc32c(d[:k])) == o.crc32c(d)


def test_crc32c_combine_and_chunks():
    rnd = np.random.default_rng(7)
    for n in [1, 65535, 65536, 65537, 200000, 3 * 65536]:
        d = rnd.integers(0, 256, n, dtype=np.uint8)
        whole = o.crc32c(d)
        for cut in (0, 1, n // 2, n - 1, n):
            a, b = d[:cut], d[cut:]
            assert o.crc32c_combine(o.crc32c(a), o.crc32c(b), len(b)) == whole
        chunks, w = o.crc32c_chunks(d, 65536)
        assert w == whole and len(chunks) == (n + 65535) // 65536
        assert chunks[0] == o.crc32c(d[:65536])
    # long shifts: the Castagnoli polynomial is not irreducible, x^(2^32) != x
    a, b = b"hello", bytes(5)
    big = (1 << 33) + 12345
    # crc(a || 0^big) by doubling: combine with a block of zeros of length big built from halves
    z1 = o.crc32c(bytes(1 << 16))
    acc, ln = z1, 1 << 16
    while ln < (1 << 33):
        acc = o.crc32c_combine(acc, acc, ln)
        ln *= 2
    rest = o.crc32c(bytes(12345))
    zeros_big = o.crc32c_combine(acc, rest, 12345)
    assert o.crc32c_combine(o.crc32c(a), zeros_big, big) == o.crc32c_combine(
        o.crc32c_combine(o.crc32c(a), acc, 1 << 33), rest, 12345)


# ---- R: ranges --------------------------------------------------------------
def test_range_parse_rules():
    rp = o.range_parse
    assert rp("0-99", 1000) == (0, [(0, 100)])
    assert rp("-5", 1000) == (0, [(995, 1000)])
    assert rp("-5000", 1000) == (0, [(0, 1000)])          # R4 suffix longer than object
    assert rp("990-", 1000) == (0, [(990, 1000)])
    assert rp("0-5000", 1000) == (0, [(0, 1000)])         # end clamped
    assert rp(" 1 - 2 , 4-4", 1000) == (0, [(1, 3), (4, 5)])
    assert rp("5-2", 1000)[0] == 416                       # dropped, none kept (R6)
    assert rp("1000-", 1000)[0] == 416
    assert rp("0-", 0)[0] == 1                             # R5: start 0 dropped spec -> ignore header
    assert rp("-0", 1000)[0] == 416                        # suffix 0: start=1000,end=1000 -> dropped, start != 0
    assert rp("a-b", 1000)[0] == 416 and rp("1", 1000)[0] == 416 and rp("1-2x", 1000)[0] == 416
    assert rp("0-999,0-999", 1000)[0] == 1                 # R7: total > length -> ignore
    assert rp("0-499,500-999", 1000) == (0, [(0, 500), (500, 1000)])
    assert rp("99999999999999999999-", 1000)[0] == 416     # overflow
    assert rp("9223372036854775807-", 1000)[0] == 416
    assert rp("500-600,5-2", 1000) == (0, [(500, 601)])    # a dropped spec with start != 0 is skipped


def test_http_time_C7():
    t = 784111777
    assert o.parse_http_time("Sun, 06 Nov 1994 08:49:37 GMT") == t
    assert o.parse_http_time("Sunday, 06-Nov-94 08:49:37 GMT") == t
    assert o.parse_http_time("Sun Nov  6 08:49:37 1994") == t
    assert o.parse_http_time("Thu, 01 Jan 1970 00:00:00 GMT") == 0
    assert o.parse_http_time("Tue, 29 Feb 2000 12:00:00 GMT") == 951825600
    assert o.parse_http_time("Mon, 29 Feb 2001 12:00:00 GMT") == -1
    assert o.parse_http_time("garbage") == -1
    import email.utils

    for s in ["Wed, 07 Nov 2018 15:20:01 GMT", "Fri, 31 Dec 1999 23:59:59 GMT", "Sat, 01 Mar 2036 00:00:01 GMT"]:
        assert o.parse_http_time(s) == int(email.utils.parsedate_to_datetime(s).timestamp())


# ---- plan: C, R, G, E --------------------------------------------------------
OBJ = o.Obj(1000, "application/x-gzip", '"5be2f2c5-a12c15d"', "Wed, 07 Nov 2018 15:20:01 GMT")


def P(rq, ob=OBJ, boundary=1):
    return o.parse_plan(o.plan(rq, ob, boundary))


def test_plan_full_and_head():
    p = P(o.Req())
    assert p.status == 200 and not p.head
    assert dict(p.headers)["Content-Length"] == "1000" and dict(p.headers)["Accept-Ranges"] == "bytes"
    assert p.body == [("slice", (0, 1000))]
    h = P(o.Req(method="HEAD"))
    assert h.status == 200 and h.head and h.headers == p.headers        # V3
    assert P(o.Req(method="POST")).status == 405
    assert P(o.Req(method="XGET")).bypass


def test_plan_single_range_R8():
    p = P(o.Req(range="bytes=10-19"))
    hd = dict(p.headers)
    assert p.status == 206 and hd["Content-Range"] == "bytes 10-19/1000" and hd["Content-Length"] == "10"
    assert "Accept-Ranges" not in hd and p.body == [("slice", (10, 10))]
    assert P(o.Req(range="BYTES=0-0")).status == 206                     # R1 case-insensitive
    assert P(o.Req(range="bytes=")).status == 200                        # R1 shorter than 7
    assert P(o.Req(range="items=0-1")).status == 200
    assert P(o.Req(range="bytes=0-0", http_minor=-1)).status == 200


def test_plan_multipart_R9():
    p = P(o.Req(range="bytes=0-9,20-29"), boundary=42)
    hd = dict(p.headers)
    assert p.status == 206
    assert hd["Content-Type"] == "multipart/byteranges; boundary=00000000000000000042"
    obj = bytes(range(256)) * 4
    body = p.body_bytes(obj[:1000])
    assert int(hd["Content-Length"]) == len(body)
    exp = (b"\r\n--00000000000000000042\r\nContent-Type: application/x-gzip\r\nContent-Range: bytes 0-9/1000\r\n\r\n" + obj[0:10] +
           b"\r\n--00000000000000000042\r\nContent-Type: application/x-gzip\r\nContent-Range: bytes 20-29/1000\r\n\r\n" + obj[20:30] +
           b"\r\n--00000000000000000042--\r\n")
    assert body == exp


def test_plan_416_R6_E1():
    p = P(o.Req(range="bytes=2000-"))
    hd = dict(p.headers)
    assert p.status == 416 and hd["Content-Range"] == "bytes */1000" and hd["Content-Type"] == "text/html"
    page = p.body_bytes(b"")
    assert page.startswith(b"<html>\r\n<head><title>416 Requested Range Not Satisfiable</title></head>")
    assert page.endswith(b"<hr><center>nginx</center>\r\n</body>\r\n</html>\r\n") and int(hd["Content-Length"]) == len(page)


def test_plan_conditionals_C():
    et, lm = OBJ.etag, OBJ.last_modified
    assert P(o.Req(if_none_match=et)).status == 304
    assert P(o.Req(if_none_match="W/" + et)).status == 304               # weak compare
    assert P(o.Req(if_none_match='"zzz", ' + et)).status == 304
    assert P(o.Req(if_none_match='"zzz"')).status == 200
    assert P(o.Req(if_none_match="*")).status == 304
    assert P(o.Req(if_modified_since=lm)).status == 304
    assert P(o.Req(if_modified_since="Wed, 07 Nov 2018 15:20:02 GMT")).status == 200    # exact (C5)
    assert P(o.Req(if_modified_since=lm, if_none_match='"zzz"')).status == 200
    assert P(o.Req(if_match='"zzz"')).status == 412
    assert P(o.Req(if_match=et)).status == 200
    assert P(o.Req(if_match="W/" + et)).status == 412                    # strong compare (C2)
    assert P(o.Req(if_unmodified_since="Wed, 07 Nov 2018 15:20:00 GMT")).status == 412
    assert P(o.Req(if_unmodified_since=lm)).status == 200
    p304 = P(o.Req(if_none_match=et))
    assert [n for n, _ in p304.headers] == ["ETag", "Last-Modified"] and p304.body == []   # C4
    # a 304 wins over Range; Range needs status 200
    assert P(o.Req(if_none_match=et, range="bytes=0-1")).status == 304


def test_plan_if_range_R2():
    et, lm = OBJ.etag, OBJ.last_modified
    assert P(o.Req(range="bytes=0-1", if_range=et)).status == 206
    assert P(o.Req(range="bytes=0-1", if_range='"other"')).status == 200
    assert P(o.Req(range="bytes=0-1", if_range=lm)).status == 206
    assert P(o.Req(range="bytes=0-1", if_range="Wed, 07 Nov 2018 15:20:02 GMT")).status == 200
    assert P(o.Req(range="bytes=0-1", if_range="W/" + et)).status == 200   # exact bytes only


def test_plan_gunzip_G():
    ob = o.Obj(300, "text/plain", '"e"', None, "gzip", raw_length=1000, gunzip_enabled=True)
    p = P(o.Req(), ob)
    hd = dict(p.headers)
    assert p.repr == "inflated" and hd["Content-Length"] == "1000" and "Content-Encoding" not in hd and hd["ETag"] == 'W/"e"'
    q = P(o.Req(accept_encoding="gzip, deflate"), ob)
    hq = dict(q.headers)
    assert q.repr == "identity" and hq["Content-Length"] == "300" and hq["Content-Encoding"] == "gzip" and hq["ETag"] == '"e"'
    r = P(o.Req(range="bytes=900-"), ob)
    assert r.status == 206 and dict(r.headers)["Content-Range"] == "bytes 900-999/1000" and r.repr == "inflated"
    ob2 = o.Obj(300, "text/plain", '"e"', None, "gzip", raw_length=1000, gunzip_enabled=False)   # B3: reference behaviour
    assert P(o.Req(), ob2).repr == "identity"


def test_gunzip_G2_G4():
    raw = (b"the quick brown fox jumps over the lazy dog\n" * 3000)[:100000]
    for level in (1, 6, 9):
        c = zlib.compressobj(level, zlib.DEFLATED, 31)
        gz = c.compress(raw) + c.flush()
        assert gz[:3] == b"\x1f\x8b\x08"
        assert o.gunzip(gz) == raw == zlib.decompress(gz, 31)
        assert o.gunzip(gz + b"\x00") is None          # trailing garbage -> declined
        assert o.gunzip(gz + gz) is None               # multi-member -> declined
        assert o.gunzip(gz[:-1]) is None
        bad = bytearray(gz); bad[len(gz) // 2] ^= 0x10
        assert o.gunzip(bytes(bad)) is None or o.gunzip(bytes(bad)) != raw
    c = zlib.compressobj(0, zlib.DEFLATED, 31)         # stored blocks
    gz = c.compress(raw) + c.flush()
    assert o.gunzip(gz) == raw
    assert o.gunzip(b"") is None


# ---- B: the hot loop ---------------------------------------------------------
def test_serve_slice_B1(tmp_path):
    rnd = np.random.default_rng(1)
    obj = rnd.integers(0, 256, 200001, dtype=np.uint8)
    f = tmp_path / "obj"
    f.write_bytes(obj.tobytes())
    fd = os.open(f, os.O_RDONLY)
    try:
        for off, ln in [(0, 200001), (0, 0), (1, 32768), (32767, 2), (100, 100000), (200000, 1)]:
            assert o.serve_slice_fd(fd, off, ln) == obj[off:off + ln].tobytes() == o.serve_slice_mem(obj, off, ln)
    finally:
        os.close(fd)
