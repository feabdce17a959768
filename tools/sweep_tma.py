#!/usr/bin/env python
"""Sweep k_gather_tma on the C2 wave: stages x tile size x L2 policy x CTAs/SM, with and without request coalescing.
Prints one line per configuration (kernel-level only: --no-e2e --no-cpu --no-http --no-sections)."""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GRID = [  # (stages, tile KiB, l2hint, ctas/sm or "")
    ("2", "32", "0", ""), ("2", "32", "1", ""), ("2", "32", "2", ""), ("3", "32", "0", ""), ("3", "32", "1", ""),
    ("2", "32", "0", "2"), ("2", "32", "1", "2"), ("2", "16", "1", ""), ("3", "16", "1", ""), ("2", "64", "1", ""),
]
for coal in ("1", "0"):
    for st, tk, l2, cps in (GRID if coal == "1" else GRID[:5]):
        env = dict(os.environ, NZ_TMA_STAGES=st, NZ_TMA_TILE_KB=tk, NZ_COALESCE=coal, NZ_TMA_L2HINT=l2)
        if cps:
            env["NZ_GATHER_CTAS_PER_SM"] = cps
        r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "10", "--warmup", "3", "--no-e2e", "--no-cpu", "--no-http", "--no-sections"],
                           env=env, capture_output=True, text=True)
        line = [l for l in r.stdout.splitlines() if l.startswith("{")]
        if not line:
            print(coal, st, tk, l2, cps, "FAILED", r.stderr[-300:]); continue
        d = json.loads(line[-1])
        print(f"coalesce {coal} stages {st} tile {tk} KiB l2hint {l2} ctas/sm {cps or 'occ'}: value {d['value']:.1f} GB/s  frac {d['roofline']['frac']:.4f}  k_gather {d['roofline']['kernel_ms']:.3f} ms", flush=True)
