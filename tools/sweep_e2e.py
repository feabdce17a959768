#!/usr/bin/env python
"""PCIe leg of the e2e path: k_gather storing straight into mapped pinned host windows, swept over the
number of CTAs (NZ_HOST_GRID), next to a plain cudaMemcpyAsync D2H of the same bytes (copy engine)."""
import json, os, subprocess, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
x = torch.empty(1 << 30, dtype=torch.uint8, device="cuda")
h = torch.empty(1 << 30, dtype=torch.uint8).pin_memory()
for _ in range(2):
    h.copy_(x, non_blocking=True); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    h.copy_(x, non_blocking=True)
e1.record(); e1.synchronize()
print("cudaMemcpyAsync D2H (copy engine): %.1f GB/s" % (5 * (1 << 30) / (e0.elapsed_time(e1) * 1e-3) / 1e9), flush=True)
del x, h
for g in (4, 8, 16, 32, 64, 128, 296):
    env = dict(os.environ, NZ_HOST_GRID=str(g))
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "3", "--warmup", "3", "--no-cpu", "--no-http", "--no-sections"], env=env, capture_output=True, text=True)
    line = [l for l in r.stdout.splitlines() if l.startswith("{")]
    if not line:
        print(g, "FAILED", r.stderr[-200:]); continue
    d = json.loads(line[-1])
    print(f"host_grid {g}: e2e {d['e2e']['value']:.1f} GB/s  ttfb native p50 {d['ttfb']['native'].get('p50_us')}", flush=True)
