#!/usr/bin/env python
"""K1 probe rate: k_lookup over a table holding N live keys, 1 Mi keys per launch, half of them hits.
Run under:  ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum -k regex:k_lookup --csv python tools/bench_probe.py
(the kernel's device time is what matters: nz_lookup itself also copies keys/results through pinned memory)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import nezha_b200 as nz
from nezha_b200 import workloads as wl

N = int(sys.argv[1]) if len(sys.argv) > 1 else 200_000
z = nz.Zone(device=0, arena_bytes=N * 4096 * 2 + (64 << 20), max_objects=N + 1024)
t0 = time.perf_counter()
keys = [wl.key128(b"probe/%d" % i) for i in range(N)]
blob = np.arange(64, dtype=np.uint8)
for k in keys:
    z.put(k, blob)
print(f"filled {N} objects in {time.perf_counter() - t0:.1f} s; table slots {z.stats().table_slots}", flush=True)
rng = np.random.default_rng(1)
q = []
for i in range(1 << 20):
    q.append(keys[int(rng.integers(0, N))] if i & 1 else rng.bytes(16))
for rep in range(3):
    t0 = time.perf_counter()
    ext = z.lookup(q)
    dt = time.perf_counter() - t0
    hits = sum(1 for e in ext if e.status == 0)
    print(f"nz_lookup of {len(q)} keys: {dt * 1e3:.1f} ms host wall ({len(q) / dt / 1e6:.1f} Mkeys/s incl. host copies), hits {hits}", flush=True)
z.close()
