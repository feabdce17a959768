#!/usr/bin/env python
"""Sweep k_gather_tma: stages x tile size, with and without request coalescing (C2 wave)."""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for coal in ("1", "0"):
    for st, tk, ah in [("2", "32", "0"), ("2", "16", "0"), ("2", "64", "0"), ("3", "16", "0"), ("3", "32", "0"), ("4", "16", "0"), ("1", "32", "0")]:
        if True:
            env = dict(os.environ, NZ_GATHER_VARIANT="14", NZ_TMA_STAGES=st, NZ_TMA_TILE_KB=tk, NZ_COALESCE=coal, NZ_TMA_AHEAD=ah)
            r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "10", "--warmup", "3", "--no-e2e", "--no-cpu", "--no-http"],
                               env=env, capture_output=True, text=True)
            line = [l for l in r.stdout.splitlines() if l.startswith("{")]
            if not line:
                print(coal, st, tk, "FAILED", r.stderr[-300:]); continue
            d = json.loads(line[-1])
            print(f"coalesce {coal} stages {st} tile {tk} KiB: value {d['value']:.1f} GB/s  frac {d['roofline']['frac']:.4f}  k_gather {d['roofline']['kernel_ms']:.3f} ms", flush=True)
