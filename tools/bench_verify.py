#!/usr/bin/env python
"""NZ_GET_VERIFY on the C2 wave: fused (gather checksums tiles in shared memory) vs separate
(k_verify re-reads every chunk), with and without request coalescing.  Device destinations,
CUDA events around whole waves (plan + gather + verify), payload GB/s.
Usage (GPU box): python tools/bench_verify.py [steps]"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

CHILD = r'''
import json, os, sys, time
import numpy as np, torch
sys.path.insert(0, %r)
import nezha_b200 as nz
from nezha_b200 import capi, workloads as wl
steps = int(sys.argv[1]); flags = int(sys.argv[2])
L = 1 << 30
z = nz.Zone(device=0, arena_bytes=L + (64 << 20), max_objects=64)
key = wl.key128(wl.object_key_string("c2", 0))
z.put(key, wl.synth_bytes(L, 2))
reqs = wl.c2_requests(L, 64, seed=22)
gets, bufs, payload = [], [], 0
for r in reqs:
    n = L if r is None else r[1]
    b = z.alloc_device(n)
    bufs.append(b)
    gets.append(nz.Get(key, b, [] if r is None else [(r[0], n, 0)], flags))
    payload += n
st = torch.cuda.Stream()
pw = z.prepare(gets) if hasattr(z, "prepare") else None
def wave():
    t = z.submit(gets, stream=st.cuda_stream)
    res = t.wait(); t.free()
    return res
for _ in range(3): res = wave()
assert all(r.status == 0 for r in res), [r.status for r in res]
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize()
ms = []
for _ in range(steps):
    e0.record(st); t = z.submit(gets, stream=st.cuda_stream); e1.record(st); t.wait(); t.free()
    torch.cuda.synchronize(); ms.append(e0.elapsed_time(e1))
s = z.stats()
print(json.dumps({"ms": float(np.median(ms)), "gbs": payload / (float(np.median(ms)) * 1e-3) / 1e9, "payload": payload,
                  "verify_chunks": s.verify_chunks, "fused": s.verify_fused_chunks}))
''' % ROOT

steps = sys.argv[1] if len(sys.argv) > 1 else "10"
rows = {}
for name, flags, env in (("no verify, coalesced", 0, {}),
                         ("verify fused, coalesced", 2, {}),
                         ("verify separate, coalesced", 2, {"NZ_FUSE_VERIFY": "0"}),
                         ("no verify, uncoalesced", 0, {"NZ_COALESCE": "0"}),
                         ("verify fused, uncoalesced", 2, {"NZ_COALESCE": "0"}),
                         ("verify separate, uncoalesced", 2, {"NZ_COALESCE": "0", "NZ_FUSE_VERIFY": "0"})):
    r = subprocess.run([sys.executable, "-c", CHILD, steps, str(flags)], env=dict(os.environ, **env), capture_output=True, text=True)
    line = [l for l in r.stdout.splitlines() if l.startswith("{")]
    if not line:
        print(name, "FAILED", r.stderr[-600:])
        continue
    rows[name] = json.loads(line[-1])
    print(f"{name:32s} {rows[name]['ms']:8.3f} ms  {rows[name]['gbs']:8.1f} GB/s payload  chunks {rows[name]['verify_chunks']} fused {rows[name]['fused']}", flush=True)
out = os.path.join(ROOT, "gpurun_out")
os.makedirs(out, exist_ok=True)
json.dump(rows, open(os.path.join(out, "bench_verify.json"), "w"), indent=1)
