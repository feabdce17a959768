#!/usr/bin/env python
"""One small NZ_GET_VERIFY wave for an ncu capture of the fused-verify gather:
   NZ_COALESCE=0 ncu --set full --clock-control none --import-source on -k regex:k_gather_tma -s 1 -c 1 -o gpurun_out/verify python tools/ncu_verify.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import nezha_b200 as nz
from nezha_b200 import capi, workloads as wl

L = 1 << 30
z = nz.Zone(device=0, arena_bytes=L + (64 << 20), max_objects=16)
key = wl.key128(b"ncu/verify")
z.put(key, wl.synth_bytes(L, 2))
bufs = [z.alloc_device(L) for _ in range(4)]
gets = [nz.Get(key, b, [], capi.NZ_GET_VERIFY) for b in bufs]
for _ in range(3):
    t = z.submit(gets)
    res = t.wait()
    t.free()
assert all(r.status == 0 for r in res)
print("ok", z.stats().verify_fused_chunks)
