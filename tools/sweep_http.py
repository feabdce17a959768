#!/usr/bin/env python
"""nzproxy front end: event-loop workers (--io-threads via NZ_PROXY_IO_THREADS) vs the round-1 thread-per-connection model,
same C2 log, loopback.  One line per configuration."""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for io in (sys.argv[1:] or ["0", "4", "8", "16", "32"]):
    env = dict(os.environ, NZ_PROXY_IO_THREADS=io, NZ_BENCH_HTTP_GPU_ONLY="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "3", "--warmup", "3", "--no-e2e", "--no-cpu", "--no-sections"],
                       env=env, capture_output=True, text=True)
    line = [l for l in r.stdout.splitlines() if l.startswith("{")]
    if not line:
        print(io, "FAILED", r.stderr[-300:]); continue
    h = json.loads(line[-1])["http"]
    g = h.get("gpu", {})
    c1 = h.get("c1", {}).get("gpu", {})
    print(f"io-threads {io:>2}: {g.get('gbs')} GB/s  ttfb p50 {g.get('ttfb_us_p50')} us  first body byte p50 {g.get('ttfbody_us_p50')} p90 {g.get('ttfbody_us_p90')} p99 {g.get('ttfbody_us_p99')} us  "
          f"errors {g.get('errors')}  | C1 one connection {c1.get('gbs')} GB/s  {h.get('error', '')}", flush=True)
