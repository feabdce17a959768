#!/usr/bin/env python
"""K3 alone: nz_crc32c_device over a 1 GiB device buffer (chunk table + fold), CUDA-synchronised wall time.
Usage (GPU box): python tools/bench_crc.py"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import nezha_b200 as nz

n = 1 << 30
z = nz.Zone(device=0, arena_bytes=64 << 20, max_objects=16)
t = torch.randint(0, 256, (n,), dtype=torch.uint8, device="cuda")
for _ in range(3):
    crc = z.crc32c_device(t.data_ptr(), n)
torch.cuda.synchronize()
best = 1e9
for _ in range(10):
    t0 = time.perf_counter()
    z.crc32c_device(t.data_ptr(), n)
    best = min(best, time.perf_counter() - t0)
print(f"k_crc_chunks + k_crc_fold, 1 GiB: {best * 1e3:.3f} ms = {n / best / 1e9:.0f} GB/s  (crc {crc:08x})")
