#!/usr/bin/env python
"""Sweep the k_gather variants (threads x unroll x CTAs/SM x cache hints) on the C2 wave.
Usage (GPU box): python tools/sweep_gather.py [steps]   -> one line per variant."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
steps = sys.argv[1] if len(sys.argv) > 1 else "20"
rows = []
for v in (int(x) for x in os.environ.get('NZ_SWEEP', ','.join(map(str, range(15)))).split(',')):
    env = dict(os.environ, NZ_GATHER_VARIANT=str(v))
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", steps, "--warmup", "3", "--no-e2e", "--no-cpu", "--no-http"],
                       env=env, capture_output=True, text=True)
    line = [l for l in r.stdout.splitlines() if l.startswith("{")]
    if not line:
        print(v, "FAILED", r.stderr[-300:])
        continue
    d = json.loads(line[-1])
    rows.append((v, d["value"], d["roofline"]["achieved"], d["roofline"]["frac"], d["roofline"]["kernel_ms"]))
    print(f"variant {v}: value {d['value']:.1f} GB/s  traffic {d['roofline']['achieved']:.1f} GB/s  frac {d['roofline']['frac']:.4f}  k_gather {d['roofline']['kernel_ms']:.3f} ms", flush=True)
best = max(rows, key=lambda x: x[3])
print("best", best)
