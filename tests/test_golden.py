"""Frozen fixtures (tests/golden/c1_small.json, made by tests/golden/make_golden.py).
CPU part: the oracle still reproduces them.  GPU part: the CUDA path serves every
body of the request log byte-for-byte (SHA-256) — SURVEY.md §8c parity definition."""
import hashlib
import json
import os

import numpy as np
import pytest

from nezha_b200 import workloads as wl
from oracle import oracle as o

G = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "c1_small.json")))


def _obj():
    d = wl.synth_bytes(G["object"]["length"], G["object"]["seed"])
    assert hashlib.sha256(d.tobytes()).hexdigest() == G["object"]["sha256"]
    return d


def test_reference_pinned_vectors():
    rp = G["reference_pinned"]
    key = o.cache_key("www.cs.toronto.edu", "/~kriz/cifar-100-python.tar.gz")
    assert key.decode() == rp["key"] and o.md5(key).hex() == rp["md5"] == "9c30938a380ef9a0ce26538a0f68408e"
    assert f"{o.crc32_ieee(key):08x}" == rp["crc32_ieee"] and rp["object_length"] == 169001437 == wl.C1_LENGTH


def test_oracle_reproduces_golden_log():
    d = _obj()
    assert f"{o.crc32c(d):08x}" == G["object"]["crc32c"]
    assert [f"{int(c):08x}" for c in o.crc32c_chunks(d, 65536)[0]] == G["object"]["chunk_crc32c"]
    ob = o.Obj(len(d), "application/x-gzip", '"5be2f2c5-a12c15d"', "Wed, 07 Nov 2018 15:20:01 GMT")
    for e in G["requests"]:
        p = o.parse_plan(o.plan(o.Req(**e["request"]), ob, e["boundary"]))
        assert p.status == e["status"] and p.head == e["head"]
        assert [list(h) for h in p.headers] == e["headers"]
        body = b"" if p.bypass else p.body_bytes(d.tobytes())
        assert len(body) == e["body_len"] and hashlib.sha256(body).hexdigest() == e["body_sha256"]
    raw = wl.text_like(G["gzip"]["raw_len"], G["gzip"]["raw_seed"]).tobytes()
    assert hashlib.sha256(raw).hexdigest() == G["gzip"]["raw_sha256"]
    assert o.gunzip(wl.gzip_member(raw)) == raw


@pytest.mark.gpu
def test_cuda_path_serves_golden_log():
    import nezha_b200 as nz
    from nezha_b200 import capi

    d = _obj()
    with nz.Zone(device=0, arena_bytes=64 << 20, max_objects=64) as z:
        key = wl.key128(wl.C1_KEY)
        crc = z.put(key, d)
        assert f"{crc:08x}" == G["object"]["crc32c"]
        for e in G["requests"]:
            if e["status"] not in (200, 206) or e["head"]:
                continue
            buf = z.alloc_host(max(e["body_len"], 1))
            try:
                view = buf.numpy()
                segs, off = [], 0
                for kind, v in e["body"]:
                    if kind == "lit":
                        b = bytes.fromhex(v)
                        view[off:off + len(b)] = np.frombuffer(b, dtype=np.uint8)
                        off += len(b)
                    else:
                        segs.append((v[0], v[1], off))
                        off += v[1]
                t = z.submit([nz.Get(key, buf, segs, capi.NZ_GET_VERIFY)])
                r = t.wait()[0]
                t.free()
                assert r.status == capi.NZ_OK and off == e["body_len"]
                assert hashlib.sha256(view[:off].tobytes()).hexdigest() == e["body_sha256"], e["request"]
            finally:
                buf.free()
        raw = wl.text_like(G["gzip"]["raw_len"], G["gzip"]["raw_seed"]).tobytes()
        gk = wl.key128(b"golden/gz")
        z.put(gk, np.frombuffer(wl.gzip_member(raw), dtype=np.uint8), enc=capi.NZ_ENC_GZIP, raw_len=len(raw))
        res, body = z.get(gk, flags=capi.NZ_GET_INFLATE)
        assert res.status == capi.NZ_OK and hashlib.sha256(body).hexdigest() == G["gzip"]["raw_sha256"]
