#!/usr/bin/env python
"""Regenerates tests/golden/*.json.

The reference cannot run here (stock nginx, no tests — SURVEY.md §8c), so these
fixtures are produced by the oracle restatement, cross-checked against the Python
stdlib (hashlib / zlib) where an independent implementation exists, and frozen so
that any later change to the oracle OR the CUDA path shows up as a diff.  The
only values that come from the reference itself are marked "reference-pinned".
Run:  python tests/golden/make_golden.py
"""
import hashlib
import json
import os
import sys
import zlib

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from nezha_b200 import workloads as wl   # noqa: E402  (inputs only)
from oracle import oracle as o           # noqa: E402

OBJ_LEN = 300_007
OBJ_SEED = 1234


def request_log():
    et, lm = '"5be2f2c5-a12c15d"', "Wed, 07 Nov 2018 15:20:01 GMT"
    R = o.Req
    return [
        R(), R(method="HEAD"), R(method="POST"), R(range="bytes=0-0"), R(range="bytes=-1"), R(range="bytes=1000-"),
        R(range="bytes=65535-65537"), R(range="bytes=299999-400000"), R(range="bytes=300007-"), R(range="bytes=5-2"),
        R(range="bytes=0-99,200-299"), R(range="bytes=-100,0-9, 150000-150099"), R(range="bytes=0-300006,0-300006"),
        R(range="bytes=7-131077", if_range=et), R(range="bytes=7-131077", if_range='"nope"'),
        R(if_none_match=et), R(if_none_match='W/"x", W/' + et), R(if_modified_since=lm), R(if_match='"x"'),
        R(if_unmodified_since="Wed, 07 Nov 2018 15:20:00 GMT"), R(range="bytes=1-2", if_none_match="*"),
        R(range="bytes=33333-66666", http_minor=0), R(range="bytes= 10 - 20 "), R(range="bytes=a-b"),
    ]


def main():
    obj = wl.synth_bytes(OBJ_LEN, OBJ_SEED)
    ob = o.Obj(OBJ_LEN, "application/x-gzip", '"5be2f2c5-a12c15d"', "Wed, 07 Nov 2018 15:20:01 GMT")
    assert hashlib.md5(wl.C1_KEY).digest() == o.md5(wl.C1_KEY)
    assert zlib.crc32(wl.C1_KEY) == o.crc32_ieee(wl.C1_KEY)
    out = {
        "reference_pinned": {
            "key": wl.C1_KEY.decode(), "key_source": "examples/demo/nginx.conf:94 + examples/demo/kubeflow-test.yaml:22",
            "md5": o.md5(wl.C1_KEY).hex(), "crc32_ieee": f"{o.crc32_ieee(wl.C1_KEY):08x}",
            "object_length": wl.C1_LENGTH, "length_source": "README.md:94,116",
        },
        "object": {"length": OBJ_LEN, "seed": OBJ_SEED, "sha256": hashlib.sha256(obj.tobytes()).hexdigest(),
                   "crc32c": f"{o.crc32c(obj):08x}",
                   "chunk_crc32c": [f"{int(c):08x}" for c in o.crc32c_chunks(obj, 65536)[0]]},
        "requests": [],
    }
    for i, rq in enumerate(request_log()):
        text = o.plan(rq, ob, boundary=i + 1)
        p = o.parse_plan(text)
        body = b"" if p.bypass else p.body_bytes(obj.tobytes())
        out["requests"].append({
            "request": {k: v for k, v in rq.__dict__.items() if v is not None},
            "boundary": i + 1, "status": p.status, "head": p.head, "headers": p.headers,
            "body": [[k, (v.hex() if k == "lit" else list(v))] for k, v in p.body],
            "body_len": len(body), "body_sha256": hashlib.sha256(body).hexdigest(),
        })
    raw = wl.text_like(200_000, seed=77).tobytes()
    gz = wl.gzip_member(raw)
    assert zlib.decompress(gz, 31) == raw == o.gunzip(gz)
    out["gzip"] = {"raw_seed": 77, "raw_len": len(raw), "raw_sha256": hashlib.sha256(raw).hexdigest(),
                   "raw_crc32": f"{zlib.crc32(raw):08x}",
                   "note": "member bytes are zlib-version dependent and are NOT frozen; the inflated bytes are"}
    with open(os.path.join(HERE, "c1_small.json"), "w") as f:
        json.dump(out, f, indent=1)
    print("wrote", os.path.join(HERE, "c1_small.json"), len(out["requests"]), "requests")


if __name__ == "__main__":
    main()
