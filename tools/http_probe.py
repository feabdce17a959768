#!/usr/bin/env python
"""HTTP-level section of bench.py alone, with nzproxy's latency trace (NZ_PROXY_TRACE=1) captured.
Usage (GPU box): python tools/http_probe.py [extra nzproxy env as K=V ...]  -> gpurun_out/http_probe.json + nzproxy_trace.log"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
for kv in sys.argv[1:]:
    k, v = kv.split("=", 1)
    os.environ[k] = v
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
os.environ.setdefault("NZ_PROXY_TRACE", "1")
os.environ.setdefault("NZ_PROXY_LOG", os.path.join(ROOT, "gpurun_out", "nzproxy_trace.log"))
import bench
from nezha_b200 import workloads as wl

L = 1 << 30
out = bench.http_bench(L, wl.c2_requests(L, 64, seed=22), 0)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "http_probe.json"), "w"), indent=1)
for k in ("gpu", "cpu_1worker", "cpu_allcores"):
    if k in out:
        print(k, {x: out[k][x] for x in ("gbs", "ttfb_us_p50", "ttfb_us_p90", "ttfb_us_p99", "ttfbody_us_p50", "ttfbody_us_p90", "ttfbody_us_p99")})
print("\n".join(open(os.environ["NZ_PROXY_LOG"]).read().strip().splitlines()[-3:]))
