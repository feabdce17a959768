#!/usr/bin/env python
"""The load generator's share of the loopback HTTP figure: nzload's receive buffer (NZ_LOAD_BUF_KIB; 64 connections x 4 MiB
does not stay in any cache) and a sink that drops the body inside the kernel (NZ_LOAD_DISCARD=1, recv MSG_TRUNC).  Both
servers (nzproxy and the nginx-restated CPU server) face the same client in each line."""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CASES = [("4096", "0"), ("256", "0"), ("64", "0"), ("256", "1")]
for buf, disc in CASES:
    env = dict(os.environ, NZ_LOAD_BUF_KIB=buf, NZ_LOAD_DISCARD=disc)
    if "--gpu-only" in sys.argv:
        env["NZ_BENCH_HTTP_GPU_ONLY"] = "1"
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "3", "--warmup", "3", "--no-e2e", "--no-cpu", "--no-sections"],
                       env=env, capture_output=True, text=True)
    line = [l for l in r.stdout.splitlines() if l.startswith("{")]
    if not line:
        print(buf, disc, "FAILED", r.stderr[-300:]); continue
    h = json.loads(line[-1])["http"]
    def f(k):
        g = h.get(k) or {}
        return f"{g.get('gbs')} GB/s, first byte p50 {g.get('ttfb_us_p50')} us, first body byte p50 {g.get('ttfbody_us_p50')} p90 {g.get('ttfbody_us_p90')} us"
    print(f"client buffer {buf:>4} KiB discard={disc}: nzproxy {f('gpu')} | cpu all cores {f('cpu_allcores')} | cpu 1 worker {(h.get('cpu_1worker') or {}).get('gbs')} GB/s", flush=True)
