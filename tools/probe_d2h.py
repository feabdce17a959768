#!/usr/bin/env python
"""What bounds the e2e leg at N GPUs: the copy engines, or the host DRAM they write into?

Launch under torchrun (N = 1, 2, 4, 8).  Every rank moves bytes device -> pinned host with plain cudaMemcpyAsync (no
kernels of ours), all ranks at the same time, and rank 0 prints one line per case:

  placement   local   = pinned pages on the NUMA node the GPU hangs off (what bench.py's e2e does)
              remote  = pinned pages on the other socket
              spread  = half local, half remote
  flags       plain | wc (cudaHostAllocWriteCombined)
  footprint   bytes of host memory the copies cycle through per rank (2 GiB = what the e2e windows add up to; 32 MiB ~ LLC)

Usage: python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/probe_d2h.py
"""
import ctypes as C
import glob
import json
import os
import sys
import time

import torch
import torch.distributed as dist


def cudart():
    d = os.path.join(os.path.dirname(torch.__file__), "lib")
    for pat in ("libcudart.so*", "../../nvidia/cuda_runtime/lib/libcudart.so*"):
        c = sorted(glob.glob(os.path.join(d, pat)))
        if c:
            return C.CDLL(c[0])
    return C.CDLL("libcudart.so")


def node_cpus(node):
    cpus = set()
    for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
        a, _, b = part.partition("-")
        cpus.update(range(int(a), int(b or a) + 1))
    return cpus


def main():
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rt = cudart()
    rt.cudaHostAlloc.argtypes = [C.POINTER(C.c_void_p), C.c_size_t, C.c_uint]
    rt.cudaFreeHost.argtypes = [C.c_void_p]
    rt.cudaMemcpyAsync.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p]
    p = torch.cuda.get_device_properties(local)
    bus = f"{p.pci_domain_id:04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"
    try:
        node = int(open(f"/sys/bus/pci/devices/{bus}/numa_node").read())
    except Exception:
        node = -1
    nodes = sorted(int(x[4:]) for x in os.listdir("/sys/devices/system/node") if x.startswith("node") and x[4:].isdigit())
    all_cpus = os.sched_getaffinity(0)
    other = [n for n in nodes if n != node]
    src = torch.empty(256 << 20, dtype=torch.uint8, device="cuda").random_(0, 256)
    stream = torch.cuda.Stream()
    CHUNK = 8 << 20
    out = []

    def alloc(nbytes, where, flags):
        if node >= 0 and where == "local":
            os.sched_setaffinity(0, node_cpus(node))
        elif where == "remote" and other:
            os.sched_setaffinity(0, node_cpus(other[0]))
        ptr = C.c_void_p()
        rc = rt.cudaHostAlloc(C.byref(ptr), nbytes, flags)
        os.sched_setaffinity(0, all_cpus)
        if rc != 0:
            raise RuntimeError(f"cudaHostAlloc rc={rc}")
        C.memset(ptr, 0, nbytes)          # (first touch happened at pin time; this only makes sure the pages are resident)
        return ptr

    cases = [("local", 0, 2 << 30), ("local", 0, 256 << 20), ("local", 0, 32 << 20), ("local", 4, 2 << 30), ("remote", 0, 2 << 30), ("spread", 0, 2 << 30)]
    for where, flags, foot in cases:
        if where == "spread":
            bufs = [alloc(foot // 2, "local", flags), alloc(foot // 2, "remote", flags)]
            sizes = [foot // 2, foot // 2]
        else:
            bufs = [alloc(foot, where, flags)]
            sizes = [foot]
        slots = []
        for b, sz in zip(bufs, sizes):
            slots += [b.value + o for o in range(0, sz, CHUNK)]
        if where == "spread":                       # alternate local / remote chunk by chunk
            h = len(slots) // 2
            slots = [x for pair in zip(slots[:h], slots[h:]) for x in pair]
        total = 16 << 30
        n = total // CHUNK
        with torch.cuda.stream(stream):
            for i in range(64):                     # warm-up
                rt.cudaMemcpyAsync(slots[i % len(slots)], src.data_ptr() + (i % 32) * CHUNK, CHUNK, 2, stream.cuda_stream)
            stream.synchronize()
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for i in range(n):
                rt.cudaMemcpyAsync(slots[i % len(slots)], src.data_ptr() + (i % 32) * CHUNK, CHUNK, 2, stream.cuda_stream)
            e1.record(stream)
            e1.synchronize()
        ms = e0.elapsed_time(e1)
        t = torch.tensor([ms], device="cuda", dtype=torch.float64)
        if world > 1:
            allt = [torch.zeros_like(t) for _ in range(world)]
            dist.all_gather(allt, t)
            times = [float(x.item()) for x in allt]
        else:
            times = [ms]
        for b in bufs:
            rt.cudaFreeHost(b)
        if rank == 0:
            rates = [total / (x * 1e-3) / 1e9 for x in times]
            out.append({"placement": where, "flags": "wc" if flags else "plain", "footprint_mib": foot >> 20, "aggregate_gbs": world * total / (max(times) * 1e-3) / 1e9,
                        "per_rank_gbs": [round(r, 1) for r in rates]})
            print(json.dumps(out[-1]), flush=True)
    if rank == 0:
        print(json.dumps({"n_gpus": world, "numa_node_rank0": node, "nodes": nodes, "cpus": len(all_cpus)}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
