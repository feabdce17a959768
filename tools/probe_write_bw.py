#!/usr/bin/env python
"""HBM write-only / read-only / copy rates with plain torch kernels (context for the write-dominated coalesced gather)."""
import torch
def rate(fn, nbytes, reps=10):
    for _ in range(3): fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(reps): fn()
    e1.record(); e1.synchronize()
    return reps * nbytes / (e0.elapsed_time(e1) * 1e-3) / 1e9
n = 1 << 31
a = torch.empty(n, dtype=torch.uint8, device="cuda"); b = torch.empty(n, dtype=torch.uint8, device="cuda")
print("fill  (write only) u8 : %.0f GB/s" % rate(lambda: a.zero_(), n))
a32 = a.view(torch.int32); b32 = b.view(torch.int32)
print("fill  (write only) i32: %.0f GB/s" % rate(lambda: a32.zero_(), n))
print("copy  (read+write)    : %.0f GB/s (2 x bytes)" % rate(lambda: b32.copy_(a32), 2 * n))
print("sum   (read only)     : %.0f GB/s" % rate(lambda: a32.sum(), n))
# 32 destination buffers written round-robin in 32 KiB pieces is what the coalesced gather does; emulate with expand+copy
src = torch.empty(1 << 26, dtype=torch.int32, device="cuda")            # 256 MiB source
dst = torch.empty((16, 1 << 26), dtype=torch.int32, device="cuda")      # 16 x 256 MiB destinations
print("bcast (1 read, 16 writes): %.0f GB/s written" % rate(lambda: dst.copy_(src.unsqueeze(0).expand(16, -1)), 16 * (1 << 28)))
