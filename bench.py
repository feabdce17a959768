#!/usr/bin/env python
"""bench.py — cached-GET GB/s of the B200 cache-hit path (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
  python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path (restated)

A "step" is one wave of BASELINE config C2: one 1 GiB object per GPU, 64
concurrent GETs (50 % full, 50 % single range), cache-hit only.
  value     payload GB/s with response buffers resident in HBM (kernel level)
  e2e       the same wave served through the C ABI into pinned HOST windows, the
            device->host copy inside the timed region (what the Go proxy would see)
  roofline  achieved HBM traffic of the gather kernel (2 x payload: read + write)
            / MEASURED_PEAKS.json hbm_gbs
  cpu_baseline  the nginx-restated CPU loop (oracle/) on this box's cores

Multi-GPU: key-hash sharding — every rank owns one shard and serves
its own object; there is no data-path collective, so scaling is weak.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WINDOW = int(os.environ.get("NZ_BENCH_WINDOW_KIB", "8192")) << 10          # e2e: per-connection pinned window (double buffered)
FALLBACK_HBM_GBS = 6650.0


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:
        return FALLBACK_HBM_GBS, "fallback"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks + throttle reasons DURING the timed region."""

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index, self.samples, self.stop_ev = index, [], threading.Event()

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        while not self.stop_ev.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}", "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            self.stop_ev.wait(0.2)

    def summary(self):
        sm = sorted(int(s[0]) for s in self.samples if s and s[0].isdigit())
        mx = [int(s[1]) for s in self.samples if len(s) > 1 and s[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for s in self.samples if len(s) >= 6 for i in range(4) if s[2 + i].lower().startswith("active")})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm)}


def bind_to_gpu_numa_node(local: int):
    """Run this rank on the CPU cores of the NUMA node its GPU hangs off, so the pinned
    response windows are allocated next to the GPU's PCIe root (first touch)."""
    try:
        import torch

        p = torch.cuda.get_device_properties(local)
        bus = f"{p.pci_domain_id:04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"
        node = int(open(f"/sys/bus/pci/devices/{bus}/numa_node").read())
        if node < 0:
            return None
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        os.sched_setaffinity(0, cpus)
        return node
    except Exception:
        return None


def c2_config(L_obj: int, nreq: int, world: int):
    """The workload both arms run (one dict, so the two JSON lines carry the same `config`)."""
    from nezha_b200 import workloads as wl

    spans = [((0, L_obj) if r is None else r) for r in wl.c2_requests(L_obj, nreq, seed=22)]
    payload = sum(l for _, l in spans)
    return {"workload": "C2: one 1 GiB object per GPU, 64 concurrent GETs (32 full + 32 single-range, log-uniform 4 KiB-256 MiB), cache-hit only",
            "object_bytes": L_obj, "requests": nreq, "request_seed": 22, "payload_bytes_per_gpu_step": payload,
            "l2": "inputs larger than L2 (1 GiB object, 33 GiB served per step); no flush needed",
            "sharding": f"key-hash x{world}, no data-path collective"}


def dist_env():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


# --------------------------------------------------------------------------- #
# the CPU arm: the reference's cache-hit loop restated (oracle/), all host cores
# --------------------------------------------------------------------------- #
def cpu_serve(requests, length, seed, threads, budget_s, reps=1, shm_dir="/dev/shm"):
    """pread(32 KiB)+copy of the C2 wave from a page-cache file; returns dict with GB/s.
    The wave is cut into WINDOW-sized jobs; a bounded sample (prefix of the shuffled job
    list) is timed so one call stays within ~budget_s."""
    import ctypes as C

    import numpy as np

    from nezha_b200 import workloads as wl
    from oracle import oracle as o

    L = o.lib()
    d = shm_dir if os.path.isdir(shm_dir) and os.access(shm_dir, os.W_OK) else "/tmp"
    path = os.path.join(d, f"nezha_bench_obj_{os.getpid()}.bin")
    data = wl.synth_bytes(length, seed)
    with open(path, "wb") as f:
        f.write(memoryview(data))
    del data
    fd = os.open(path, os.O_RDONLY)
    try:
        spans = []
        for ci, rq in enumerate(requests):
            s, l = (0, length) if rq is None else rq
            for off in range(0, l, WINDOW):
                spans.append((ci, s + off, min(WINDOW, l - off)))
        rng = np.random.default_rng(1)
        rng.shuffle(spans)
        nconn = len(requests)
        win = np.empty((nconn, WINDOW), dtype=np.uint8)      # one window per connection (response buffer)
        win[:] = 0

        def run(sample):
            jobs = (o.Job * len(sample))()
            for i, (ci, off, ln) in enumerate(sample):
                jobs[i].fd, jobs[i].off, jobs[i].len = fd, off, ln
                jobs[i].dst = win[ci].ctypes.data
            t = L.nzo_serve_jobs(jobs, len(sample), threads)
            if t < 0:
                raise RuntimeError("cpu serve failed")
            return t, sum(x[2] for x in sample)

        # calibrate on ~1 GiB, then size the sample
        probe = []
        acc = 0
        for sp in spans:
            probe.append(sp)
            acc += sp[2]
            if acc >= (1 << 30):
                break
        t, b = run(probe)           # also warms the page cache mapping
        t, b = run(probe)
        rate = b / t
        total = sum(x[2] for x in spans)
        want = min(total, int(rate * budget_s))
        sample, acc = [], 0
        for sp in spans:
            sample.append(sp)
            acc += sp[2]
            if acc >= want:
                break
        times = []
        for _ in range(reps):
            t, b = run(sample)
            times.append(t)
        return {"gbs": b / min(times) / 1e9, "times": times, "bytes": b, "total_bytes": total,
                "sample": f"{len(sample)} of {len(spans)} 8 MiB windows of the C2 wave ({b / 2**30:.2f} of {total / 2**30:.2f} GiB), "
                          f"pread 32 KiB + copy from a page-cache file, {threads} threads"}
    finally:
        os.close(fd)
        try:
            os.unlink(path)
        except OSError:
            pass


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path.  nginx is not installable here (no
    network, not in the image), so this is the oracle port of its cache-hit loop (kind "port").  `value` is the
    reference's own configuration - ONE worker (`worker_processes 1`, /root/reference/examples/demo/nginx.conf:2);
    the same loop on every host thread is reported beside it (`all_cores`)."""
    rank, world, _ = dist_env()
    if rank != 0:
        return 0
    from nezha_b200 import workloads as wl

    threads = os.cpu_count() or 1
    reqs = wl.c2_requests(args.object_bytes, args.requests, seed=22)
    nsteps = args.steps + args.warmup
    budget = max(1.0, min(10.0, 90.0 / max(nsteps, 1)))
    r1 = cpu_serve(reqs, args.object_bytes, 2, 1, budget, reps=nsteps)
    times = r1["times"][args.warmup:]
    t_step = sum(times) / len(times)
    gbs = r1["bytes"] / t_step / 1e9     # the box has one set of host cores whatever N is: the CPU rate does not scale with GPUs
    ra = cpu_serve(reqs, args.object_bytes, 2, threads, 6.0, reps=3)
    line = {
        "impl": "reference", "metric": "cached-GET GB/s", "value": gbs, "unit": "GB/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": t_step * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": c2_config(args.object_bytes, args.requests, world),
        "cpu_baseline": {"value": gbs, "unit": "GB/s", "cores": 1, "kind": "port", "sample": r1["sample"]},
        "all_cores": {"value": ra["gbs"], "unit": "GB/s", "cores": threads, "sample": ra["sample"],
                      "note": "the same loop on every host thread: NOT a configuration the reference ships (nginx.conf:2 is worker_processes 1)"},
        "e2e": {"value": gbs, "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "nginx-restated CPU path (oracle port of nginx.conf's cache-hit loop: pread 32 KiB + copy), not nginx itself; one worker as "
                "nginx.conf:2 configures it; README.md:116 quotes 323 MB/s for one client of that configuration over a pod network",
    }
    print(json.dumps(line), flush=True)
    return 0


# --------------------------------------------------------------------------- #
# the CUDA arm
# --------------------------------------------------------------------------- #
def run_ours(args):
    import numpy as np
    import torch

    import nezha_b200 as nz
    from nezha_b200 import capi, workloads as wl

    rank, world, local = dist_env()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (this path has no CPU fallback)")
    torch.cuda.set_device(local)
    all_cpus = os.sched_getaffinity(0)
    numa = bind_to_gpu_numa_node(local)
    if world > 1:
        import torch.distributed as dist

        import datetime

        dist.init_process_group("nccl", device_id=torch.device("cuda", local), timeout=datetime.timedelta(seconds=240))
    L_obj = args.object_bytes
    reqs = wl.c2_requests(L_obj, args.requests, seed=22)      # the same request set on every rank: per-rank work is identical
    spans = [((0, L_obj) if r is None else r) for r in reqs]
    payload = sum(l for _, l in spans)

    zone = nz.Zone(device=local, arena_bytes=L_obj + (256 << 20), max_objects=1024, shard_index=rank, shard_count=world)
    zone.set_profiling(True)     # per-kernel CUDA events (roofline timing)
    # the shard's object: first bench key this shard owns (key-hash placement)
    i = 0
    while not zone.owns(wl.key128(wl.object_key_string("c2", i))):
        i += 1
    key = wl.key128(wl.object_key_string("c2", i))
    data = wl.synth_bytes(L_obj, seed=2 + rank)
    t0 = time.perf_counter()
    crc = zone.put(key, data)
    fill_s = time.perf_counter() - t0

    # ---- kernel level: response buffers in HBM --------------------------------
    offs, total = [], 0
    for _, l in spans:
        offs.append(total)
        total += (l + 255) // 256 * 256
    dbuf = zone.alloc_device(total)
    gets = [nz.Get(key, dbuf, [] if r is None else [(r[0], r[1], 0)], 0, offs[j]) for j, r in enumerate(reqs)]
    wave = zone.prepare(gets)
    # a real (non-default) stream: handle 0 would mean "use the zone's stream pool" to the C ABI
    stream = torch.cuda.Stream(device=local)
    torch.cuda.set_stream(stream)
    sptr = stream.cuda_stream
    assert sptr != 0

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local)
    sampler.start()          # samples from warm-up through the timed loop (both run the same wave back to back)
    for _ in range(max(args.warmup, 3)):
        t = wave.submit(sptr)
        res = t.wait()
        t.free()
    bad = [r.status for r in res if r.status != capi.NZ_OK]
    if bad or sum(r.bytes for r in res) != payload:
        raise SystemExit(f"bench.py: wave failed: {bad[:4]} bytes={sum(r.bytes for r in res)} want={payload}")
    # spot parity of the wave inside bench (cheap): three slices against the source bytes
    for j in (0, len(reqs) // 2, len(reqs) - 1):
        s, l = spans[j]
        n = min(l, 1 << 16)
        got = dbuf.bytes(offs[j] + l - n, n)
        if got != data[s + l - n:s + l].tobytes():
            raise SystemExit("bench.py: served bytes differ from the source object")

    launches0 = zone.stats().kernel_launches
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record(stream)
    tickets = [wave.submit(sptr) for _ in range(args.steps)]
    ev1.record(stream)
    ev1.synchronize()
    barrier()
    sampler.stop_ev.set()
    ms = ev0.elapsed_time(ev1)
    gather_ms = [t.gather_ms() for t in tickets]
    for t in tickets:
        t.wait()
        t.free()
    launches = zone.stats().kernel_launches - launches0
    if world > 1:
        tt = torch.tensor([ms], device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms = float(tt.item())
        pt = torch.tensor([float(payload)], device="cuda", dtype=torch.float64)
        dist.all_reduce(pt)
        payload_all = float(pt.item())
    else:
        payload_all = float(payload)
    ms_step = ms / args.steps
    value = payload_all / (ms_step * 1e-3) / 1e9
    hbm, which = peaks()
    # context for write-dominated waves: what a pure fill reaches on this GPU right now
    wbuf = torch.empty(1 << 29, dtype=torch.int32, device="cuda")      # 2 GiB, filled as 32-bit words
    for _ in range(2):
        wbuf.zero_()
    w0, w1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    w0.record(stream)
    for _ in range(5):
        wbuf.zero_()
    w1.record(stream)
    w1.synchronize()
    write_peak = 5 * (1 << 31) / (w0.elapsed_time(w1) * 1e-3) / 1e9
    del wbuf
    g_ms = sum(gather_ms) / len(gather_ms)
    coalesce = os.environ.get("NZ_COALESCE", "1") != "0"
    # algorithmic bytes of one launch: every payload byte is written once; the source is read once per
    # DISTINCT slice when request coalescing is on (identical concurrent GETs are chained behind one
    # leader and read once), once per request when it is off (2 x P)
    distinct = {}
    for j, (s_, l_) in enumerate(spans):
        distinct[(s_, l_, offs[j] & 127)] = l_
    read_bytes = sum(distinct.values()) if coalesce else payload
    algo = read_bytes + payload
    achieved = algo / (g_ms * 1e-3) / 1e9
    roofline = {"bound": "hbm", "achieved": achieved, "peak": hbm, "unit": "GB/s", "frac": achieved / hbm,
                "traffic": None, "peak_source": which, "kernel": "k_gather_tma",
                "algorithmic_bytes_per_launch": algo, "kernel_ms": g_ms,
                "accounting": ("source read once per distinct slice (request coalescing) + every payload byte written" if coalesce
                               else "2 x payload (read + write per request)"),
                "uncoalesced_equivalent_gbs": 2.0 * payload / (g_ms * 1e-3) / 1e9,
                "write_only_fill_gbs": write_peak, "read_bytes": read_bytes, "write_bytes": payload,
                "kernel_share_of_step": g_ms / (ms / args.steps) if world == 1 else None}

    try:
        with open(os.path.join(ROOT, "profiles", "k_gather_traffic.json")) as f:
            tr = json.load(f)
        if tr.get("request_coalescing") == coalesce and tr.get("payload_bytes") == payload:
            roofline["traffic"] = tr["dram_bytes_read"] + tr["dram_bytes_write"]
            roofline["traffic_source"] = tr.get("source")
    except Exception:
        pass

    # ---- e2e: through the C ABI into pinned host windows ----------------------
    e2e = None
    ttfb = None
    if not args.no_e2e:
        nreq = len(reqs)
        hbuf = zone.alloc_host(nreq * 2 * WINDOW)
        rounds = max((l + WINDOW - 1) // WINDOW for _, l in spans)
        waves, h2d = [], 0
        for r_i in range(rounds):
            g = []
            for j, (s, l) in enumerate(spans):
                off = r_i * WINDOW
                if off < l:
                    g.append(nz.Get(key, hbuf, [(s + off, min(WINDOW, l - off), 0)], 0, (2 * j + (r_i & 1)) * WINDOW))
            waves.append(zone.prepare(g))
            h2d += len(g) * (64 + 32)
        d2h = payload + nreq * 32 * 1

        def e2e_step():
            # like a server: a connection's window is reused only after the wave that filled it has completed, so
            # two waves are in flight (double-buffered windows), on the library's stream pool: the staging gather
            # of wave k+1 overlaps the copy-engine transfers of wave k
            inflight = []
            last = None
            for w in waves:
                if len(inflight) == 2:
                    t = inflight.pop(0)
                    last = t.wait()
                    t.free()
                inflight.append(w.submit(None))
            for t in inflight:
                last = t.wait()
                t.free()
            return last

        for _ in range(1):
            e2e_step()
        # parity of the last window of connection 0 as landed in host memory
        s0, l0 = spans[0]
        last_r = (l0 - 1) // WINDOW
        n0 = l0 - last_r * WINDOW
        if hbuf.bytes((0 + (last_r & 1)) * WINDOW, min(n0, 4096)) != data[s0 + last_r * WINDOW:s0 + last_r * WINDOW + min(n0, 4096)].tobytes():
            raise SystemExit("bench.py: e2e host window differs from the source object")
        e_steps = max(1, min(args.steps, 3))
        barrier()
        t0 = time.perf_counter()
        for _ in range(e_steps):
            e2e_step()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        if world > 1:
            tt = torch.tensor([dt], device="cuda", dtype=torch.float64)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            dt = float(tt.item())
        e2e = {"value": payload_all * e_steps / dt / 1e9, "unit": "GB/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
               "steps": e_steps, "window_bytes": WINDOW, "note": "two waves in flight (double-buffered windows per connection, reused only after their wave completed); each wave "
               "is gathered into an HBM staging slot and moved to the pinned windows by the copy engine (PCIe D2H in the timed region; "
               "NZ_HOST_DMA=0: the gather kernel stores across PCIe itself); request descriptors are read by the plan kernel from pinned host memory"}
        # ---- TTFB: first 256 KiB of each GET, unloaded, host clock around submit+wait ---
        lat = []
        for j, (s, l) in enumerate(spans):
            w = zone.prepare([nz.Get(key, hbuf, [(s, min(l, 256 << 10), 0)], 0, 2 * j * WINDOW)])
            for rep in range(3):
                t0 = time.perf_counter()
                t = w.submit(None)
                t.wait()
                dtt = time.perf_counter() - t0
                t.free()
            lat.append(dtt * 1e6)
        lat.sort()
        native = None
        try:        # the same measurement without the interpreter in the loop
            r = subprocess.run([os.path.join(ROOT, "nezha_b200", "_build", "nzlat"), "--device", str(local), "--iters", "2000"],
                               capture_output=True, text=True, timeout=120)
            native = json.loads(r.stdout.strip().splitlines()[-1])
        except Exception as e:
            native = {"error": repr(e)}
        ttfb = {"native": native, "p50_us": lat[len(lat) // 2], "p99_us": lat[min(len(lat) - 1, int(len(lat) * 0.99))], "first_bytes": 256 << 10,
                "note": "host clock around nz_get_submit+nz_get_wait through Python ctypes, one GET in flight, first 256 KiB to pinned host"}
        hbuf.free()

    # ---- C1: the reference's own published case (README.md:116): one client, one full GET of the 169 001 437 B object ----
    c1 = None
    if rank == 0 and world == 1 and not args.no_e2e:
        try:
            c1 = c1_section(zone, nz, wl, capi, data_seed=1)
        except Exception as e:
            c1 = {"error": repr(e)}

    # ---- CPU baseline beside it (rank 0, N=1 only) -----------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        os.sched_setaffinity(0, all_cpus)      # the CPU arm gets every host core
        threads = os.cpu_count() or 1
        r1 = cpu_serve(wl.c2_requests(L_obj, args.requests, seed=22), L_obj, 2, 1, min(args.cpu_budget, 6.0))
        r = cpu_serve(wl.c2_requests(L_obj, args.requests, seed=22), L_obj, 2, threads, args.cpu_budget)
        cpu = {"value": r1["gbs"], "unit": "GB/s", "cores": 1, "kind": "port", "sample": r1["sample"],
               "note": "worker_processes 1 (nginx.conf:2): the reference's actual configuration",
               "all_cores_gbs": r["gbs"], "all_cores": threads, "all_cores_sample": r["sample"]}

    dbuf.free()
    zone.close()
    zone = None
    # ---- C4 / K5: the one path with an exchange step, whenever there is more than one GPU -------------
    span_res = None
    if world > 1 and not args.no_span:
        try:
            span_res = span_section(args, dist, rank, world, local)
        except Exception as e:          # never take the C2 numbers down with it
            span_res = {"error": repr(e)}
    c5_res = None
    if world > 1 and not args.no_sections:
        try:
            c5_res = c5_fill_section(args, dist, rank, world, local)
        except Exception as e:
            c5_res = {"error": repr(e)}
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()       # rank 0 goes on alone: the HTTP front end owns every GPU from here
        if rank != 0:
            return 0
    os.sched_setaffinity(0, all_cpus)
    http = None
    if not args.no_http:
        http = http_bench(L_obj, reqs, list(range(world)) if world > 1 else [local], with_c1=(world == 1))
    sections = {}
    if world == 1 and not args.no_sections:
        # the other BASELINE configs, each by its own process (a failure or a leak there cannot touch this line)
        sections["c3"] = run_section(["--workload", "c3", "--steps", "3", "--warmup", "3", "--no-cpu"], 600)
        sections["c5"] = run_section(["--workload", "c5", "--steps", "5"], 600)
    line = {
        "metric": "cached-GET GB/s", "value": value, "unit": "GB/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": c2_config(L_obj, len(reqs), world),
        "arm": {"response_buffers": "HBM", "request_coalescing": coalesce, "gather_kernel": "k_gather_tma"},
        "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "ttfb": ttfb, "http": http, "gpu_launches": int(launches),
        "clocks": sampler.summary(), "numa_node": numa, "fill_gbs": L_obj / fill_s / 1e9, "object_crc32c": f"{crc:08x}",
        "c1": c1, "span": span_res, "c5_fill": c5_res, **sections,
    }
    print(json.dumps(line), flush=True)
    return 0


def run_section(extra, timeout_s):
    """One of the other BASELINE configs in its own process; returns its JSON line (or the error)."""
    env = dict(os.environ)
    for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK", "MASTER_ADDR", "MASTER_PORT"):
        env.pop(k, None)
    try:
        r = subprocess.run([sys.executable, os.path.abspath(__file__)] + extra, capture_output=True, text=True, timeout=timeout_s, env=env)
        for ln in reversed(r.stdout.strip().splitlines()):
            if ln.startswith("{"):
                return json.loads(ln)
        return {"error": f"rc={r.returncode}: {r.stderr.strip()[-400:]}"}
    except Exception as e:
        return {"error": repr(e)}


def c1_section(zone, nz, wl, capi, data_seed=1):
    """BASELINE configs[0]: ONE connection, one full GET of cifar-100-python.tar.gz's length (169 001 437 B), the case
    /root/reference/README.md:116 publishes as 323 MB/s (wget through the one-worker nginx proxy cache over a pod network).
    Here: the C ABI serving that GET window by window into pinned host memory, one request in flight (what a single
    HTTP connection of the proxy does), and the same bytes into an HBM response buffer."""
    import time as _t

    L = wl.C1_LENGTH
    key = wl.key128(wl.C1_KEY)
    data = wl.synth_bytes(L, seed=data_seed)
    zone.put(key, data)
    hbuf = zone.alloc_host(2 * WINDOW)
    waves = []
    for i, off in enumerate(range(0, L, WINDOW)):
        waves.append(zone.prepare([nz.Get(key, hbuf, [(off, min(WINDOW, L - off), 0)], 0, (i & 1) * WINDOW)]))

    def once():
        inflight = []
        for w in waves:                      # double-buffered windows, like one proxy connection
            if len(inflight) == 2:
                t = inflight.pop(0); t.wait(); t.free()
            inflight.append(w.submit(None))
        for t in inflight:
            t.wait(); t.free()

    once()
    last = (L - 1) // WINDOW
    n0 = L - last * WINDOW
    if hbuf.bytes((last & 1) * WINDOW + n0 - 4096, 4096) != data[L - 4096:].tobytes():
        raise SystemExit("bench.py c1: host window differs from the source object")
    ts = []
    for _ in range(5):
        t0 = _t.perf_counter(); once(); ts.append(_t.perf_counter() - t0)
    hbuf.free()
    dbuf = zone.alloc_device(L + 256)
    w = zone.prepare([nz.Get(key, dbuf, [], 0, 0)])
    dev = []
    for _ in range(6):
        t = w.submit(None); t.wait(); dev.append(t.elapsed_ms()); t.free()
    dbuf.free()
    zone.evict(key)
    best = min(ts)
    return {"workload": "C1: single cached GET of a 169 001 437 B object, one connection (examples/demo/kubeflow-test.yaml:22)",
            "object_bytes": L, "c_abi_host_mbs": L / best / 1e6, "c_abi_host_ms": best * 1e3,
            "device_ms": sorted(dev[1:])[len(dev[1:]) // 2], "device_gbs": L / (sorted(dev[1:])[len(dev[1:]) // 2] * 1e-3) / 1e9,
            "reference_published_mbs": 323, "reference_source": "README.md:116 (wget via the nginx proxy cache, pod network)",
            "note": "c_abi_host: 8 MiB windows, two in flight, bytes land in pinned host memory (PCIe D2H inside the timed region); "
                    "the HTTP figures for the same GET are in http.c1"}


# --------------------------------------------------------------------------- #
# HTTP level: the metric at the wire (GB/s and p50 TTFB), GPU proxy vs restated CPU server
# --------------------------------------------------------------------------- #
def http_bench(L_obj, requests, devices, seed=2, with_c1=False):
    """nzload (64 keep-alive connections per GPU, the C2 log) against nzproxy (bodies from HBM; `--devices` shards the
    objects by key hash over the GPUs behind one front end) and against oracle/nz_cpu_server (pread 32 KiB + send) with
    1 worker (nginx.conf:2) and with one worker per core.  Loopback TCP on the same box, same run.  with_c1: also the
    reference's published case, ONE connection fetching the 169 001 437 B object (README.md:116: 323 MB/s)."""
    import hashlib
    import http.server
    import shutil
    import signal
    import socket
    import tempfile
    import threading

    from nezha_b200 import shard, workloads as wl

    def free_port():
        s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p

    def wait_port(port, timeout=180):
        t0 = time.time()
        while time.time() - t0 < timeout:
            try:
                socket.create_connection(("127.0.0.1", port), 0.2).close(); return True
            except OSError:
                time.sleep(0.1)
        return False

    base = "/dev/shm" if os.access("/dev/shm", os.W_OK) else "/tmp"
    d = tempfile.mkdtemp(prefix="nezha_http_", dir=base)
    procs = []
    out = {}
    ngpu = len(devices)
    try:
        host = "storage.googleapis.com"
        # one C2 object per GPU: the first bench keys whose hash lands on each shard (key-hash placement)
        objects = {}                # path -> (file, length, etag)
        paths = []
        i = 0
        while len(paths) < ngpu:
            ks = wl.object_key_string("c2", i)
            if shard.shard_of(wl.key128(ks), ngpu) == len(paths):
                path = "/nezha-bench/c2/obj-%08d.bin" % i
                hx = hashlib.md5(ks).hexdigest()
                with open(os.path.join(d, hx), "wb") as f:
                    f.write(memoryview(wl.synth_bytes(L_obj, seed + len(paths))))
                objects[path] = (hx, L_obj, '"c2-obj-%d"' % i)
                paths.append(path)
                i = -1              # restart the scan for the next shard
            i += 1
        c1_path, c1_host = "/~kriz/cifar-100-python.tar.gz", "www.cs.toronto.edu"
        if with_c1:
            hx = hashlib.md5(wl.C1_KEY).hexdigest()
            with open(os.path.join(d, hx), "wb") as f:
                f.write(memoryview(wl.synth_bytes(wl.C1_LENGTH, 1)))
            objects[c1_path] = (hx, wl.C1_LENGTH, '"5a2d5e4a-a12c59d"')
        with open(os.path.join(d, "manifest"), "w") as f:
            for path, (hx, ln, etag) in objects.items():
                f.write(f'{hx} {ln} -1 application/octet-stream {etag} - Wed, 07 Nov 2018 15:20:01 GMT\n')
        with open(os.path.join(d, "log"), "w") as f:
            for r in requests:
                for path in paths:
                    f.write(f"{host} {path} {'-' if r is None else f'bytes={r[0]}-{r[0] + r[1] - 1}'}\n")
        with open(os.path.join(d, "log_c1"), "w") as f:
            f.write(f"{c1_host} {c1_path} -\n")
        nconn = len(requests) * ngpu
        nzload = os.path.join(ROOT, "nezha_b200", "_build", "nzload")

        class Origin(http.server.BaseHTTPRequestHandler):
            protocol_version = "HTTP/1.0"

            def do_GET(self):
                ent = objects.get(self.path)
                if ent is None:
                    self.send_error(404); return
                self.send_response(200)
                self.send_header("Content-Type", "application/octet-stream")
                self.send_header("Content-Length", str(ent[1]))
                self.send_header("ETag", ent[2])
                self.send_header("Last-Modified", "Wed, 07 Nov 2018 15:20:01 GMT")
                self.send_header("Accept-Ranges", "bytes")        # GCS / S3 / Apache say so; nginx serves ranges from its cache only then (SPEC R0)
                self.end_headers()
                with open(os.path.join(d, ent[0]), "rb") as fh:
                    shutil.copyfileobj(fh, self.wfile, 4 << 20)

            def log_message(self, *a):
                pass

        oport, gport = free_port(), free_port()
        httpd = http.server.ThreadingHTTPServer(("127.0.0.1", oport), Origin)
        threading.Thread(target=httpd.serve_forever, daemon=True).start()

        def load(port, repeat, warmup, log="log", conns=None):
            r = subprocess.run([nzload, "--target", f"127.0.0.1:{port}", "--conns", str(conns or nconn), "--log", os.path.join(d, log),
                                "--repeat", str(repeat), "--warmup", str(warmup)], capture_output=True, text=True, timeout=900)
            return json.loads(r.stdout.strip().splitlines()[-1])

        # ---- GPU proxy ----
        dev_args = ["--devices", ",".join(str(x) for x in devices)] if ngpu > 1 else ["--device", str(devices[0])]
        p = subprocess.Popen([os.path.join(ROOT, "nezha_b200", "_build", "nzproxy"), "--listen", f"127.0.0.1:{gport}", *dev_args,
                              "--arena-gib", str(L_obj / 2**30 + 0.75), "--max-conns", str(2 * nconn + 8), "--window-mib", "4", "--first-window-kib", os.environ.get("NZ_BENCH_FIRST_KIB", "16"),
                              "--upstream", f"{host}=127.0.0.1:{oport}", "--upstream", f"{c1_host}=127.0.0.1:{oport}"], start_new_session=True,
                             stderr=open(os.environ["NZ_PROXY_LOG"], "w") if os.environ.get("NZ_PROXY_LOG") else subprocess.DEVNULL)
        procs.append(p)
        if wait_port(gport):
            import http.client
            t0 = time.perf_counter()
            for hh, path in [(host, q) for q in paths] + ([(c1_host, c1_path)] if with_c1 else []):
                c = http.client.HTTPConnection("127.0.0.1", gport, timeout=600)
                c.request("GET", path, headers={"Host": hh, "Range": "bytes=0-0"})      # miss -> fill from the origin
                c.getresponse().read()
                c.close()
            fill_s = time.perf_counter() - t0
            out["gpu"] = load(gport, 2, 1)
            out["gpu"]["fill_from_origin_s"] = fill_s
            out["gpu"]["devices"] = ngpu
            if with_c1:
                out.setdefault("c1", {})["gpu"] = load(gport, 5, 1, "log_c1", 1)
        os.killpg(p.pid, signal.SIGTERM)
        httpd.shutdown()
        # ---- CPU servers ----
        for name, workers, repeat in (("cpu_1worker", 1, 1), ("cpu_allcores", os.cpu_count() or 1, 2)):
            if os.environ.get("NZ_BENCH_HTTP_GPU_ONLY"):        # tools: front-end sweeps
                break
            if ngpu > 2 and workers == 1:
                repeat = 1
            cport = free_port()
            p = subprocess.Popen([os.path.join(ROOT, "oracle", "_build", "nz_cpu_server"), "--root", d, "--listen", f"127.0.0.1:{cport}",
                                  "--workers", str(workers)], start_new_session=True, stderr=subprocess.DEVNULL)
            procs.append(p)
            if wait_port(cport, 30):
                if workers == 1 and ngpu > 1:        # one worker serves ~4 GB/s whatever the GPU count: a bounded sample (one GPU's share of the log)
                    with open(os.path.join(d, "log_1w"), "w") as f:
                        for r in requests:
                            f.write(f"{host} {paths[0]} {'-' if r is None else f'bytes={r[0]}-{r[0] + r[1] - 1}'}\n")
                    out[name] = load(cport, 1, 0, "log_1w", len(requests))
                    out[name]["sample"] = "one GPU's share of the log (64 connections, one object)"
                else:
                    out[name] = load(cport, repeat, 1 if workers > 1 else 0)
                out[name]["workers"] = workers
                if with_c1 and workers == 1:
                    out.setdefault("c1", {})["cpu_1worker"] = load(cport, 3, 1, "log_c1", 1)
            os.killpg(p.pid, signal.SIGTERM)
        if with_c1 and "c1" in out:
            out["c1"]["reference_published_mbs"] = 323
            out["c1"]["note"] = ("ONE connection, full GET of the 169 001 437 B object (README.md:116 publishes 323 MB/s for this through the one-worker "
                                 "nginx proxy cache over a pod network); here loopback TCP, same box, same run: gbs x 1000 = MB/s")
        out["note"] = ("loopback TCP, nzload with %d keep-alive connections (256 KiB receive buffer each) replaying the C2 log against %d object(s), one per GPU; TTFB = request sent -> first response byte (ttfbody: first BODY byte; nzproxy sends the header before asking the GPU for the first window); "
                       "cpu_* is the nginx-restated server (oracle/nz_cpu_server.c), not nginx" % (nconn, ngpu))
    except Exception as e:      # the HTTP section must never take the kernel-level numbers down with it
        out["error"] = repr(e)
    finally:
        import signal as _s
        for p in procs:
            try:
                os.killpg(p.pid, _s.SIGKILL)
            except Exception:
                pass
        shutil.rmtree(d, ignore_errors=True)
    return out


# --------------------------------------------------------------------------- #
# C3: 10k x 1 MiB gzip objects, Zipf(0.99) access, inflate on every hit
# --------------------------------------------------------------------------- #
def _c3_make(i):
    import hashlib

    from nezha_b200 import workloads as wl

    raw = wl.text_like(wl.C3_RAW, seed=30_000 + i)
    return wl.gzip_member(raw), hashlib.sha256(raw.tobytes()).hexdigest()


def run_c3(args):
    import hashlib
    import multiprocessing as mp

    import numpy as np
    import torch

    import nezha_b200 as nz
    from nezha_b200 import capi, workloads as wl

    rank, world, local = dist_env()
    if world > 1:
        raise SystemExit("bench.py --workload c3 is a single-GPU configuration (BASELINE configs[2])")
    torch.cuda.set_device(local)
    n_obj, wave_n = args.c3_objects, args.c3_wave
    t0 = time.perf_counter()
    cache = os.environ.get("NZ_C3_CACHE")          # tools: keep the generated members between runs of one session
    cpath = os.path.join(cache, f"c3_{n_obj}.pkl") if cache else None
    if cpath and os.path.exists(cpath):
        import pickle
        with open(cpath, "rb") as f:
            made = pickle.load(f)
    else:
        with mp.get_context("fork").Pool(min(64, os.cpu_count() or 1)) as pool:
            made = pool.map(_c3_make, range(n_obj), chunksize=8)
        if cpath:
            import pickle
            os.makedirs(cache, exist_ok=True)
            with open(cpath, "wb") as f:
                pickle.dump(made, f, protocol=4)
    gen_s = time.perf_counter() - t0
    comp_total = sum(len(g) for g, _ in made)
    # arena: the members + their checkpoint indexes (~7 % of the raw size each, 4 KiB granular)
    zone = nz.Zone(device=local, arena_bytes=comp_total + n_obj * (8192 + (128 << 10)) + (64 << 20), max_objects=max(2 * n_obj, 1024))
    zone.set_profiling(True)
    keys = [wl.key128(wl.object_key_string("c3", i)) for i in range(n_obj)]
    t0 = time.perf_counter()
    used0 = zone.stats().arena_used
    for b0 in range(0, n_obj, 1024):              # bulk load: nz_put_commit_many (CRC, index build and publish batched)
        zone.put_many([(keys[i], np.frombuffer(made[i][0], dtype=np.uint8), wl.C3_RAW) for i in range(b0, min(n_obj, b0 + 1024))],
                      enc=capi.NZ_ENC_GZIP)
    fill_s = time.perf_counter() - t0
    arena_used = zone.stats().arena_used - used0
    ranks = wl.zipf_ranks(1_000_000, n_obj, 0.99, seed=3)
    nwaves = args.steps + max(args.warmup, 3)
    dbuf = zone.alloc_device(wave_n * wl.C3_RAW)
    waves = []
    for w in range(nwaves):
        rr = ranks[(w * wave_n) % (len(ranks) - wave_n):][:wave_n]
        waves.append((rr, zone.prepare([nz.Get(keys[r], dbuf, [], capi.NZ_GET_INFLATE, j * wl.C3_RAW) for j, r in enumerate(rr)])))
    stream = torch.cuda.Stream(device=local)
    torch.cuda.set_stream(stream)
    sptr = stream.cuda_stream
    sampler = ClockSampler(local)
    sampler.start()
    for rr, w in waves[:max(args.warmup, 3)]:
        t = w.submit(sptr)
        res = t.wait()
        t.free()
    if any(r.status != capi.NZ_OK for r in res):
        raise SystemExit("bench.py c3: inflate wave failed")
    # steady state keeps as many tickets pooled as there are waves in flight: create them (pinned descriptor block, device
    # block, inflate scratch) before the timed region, like every other one-time allocation
    nwarm = max(args.warmup, 3)
    pool_warm = [waves[i % nwarm][1].submit(sptr) for i in range(args.steps)]
    for t in pool_warm:
        t.wait()
        t.free()
    rr = waves[max(args.warmup, 3) - 1][0]
    for j in (0, wave_n // 2, wave_n - 1):       # spot parity vs the SHA-256 of the generator's raw bytes
        if hashlib.sha256(dbuf.bytes(j * wl.C3_RAW, wl.C3_RAW)).hexdigest() != made[rr[j]][1]:
            raise SystemExit("bench.py c3: inflated bytes differ from the source")
    # one member alone (latency of a single GET that inflates): one 1 MiB object, nothing else on the GPU
    single = []
    for j in range(8):
        g1 = zone.prepare([nz.Get(keys[int(rr[j])], dbuf, [], capi.NZ_GET_INFLATE, 0)])
        for _ in range(3):
            t = g1.submit(sptr)
            t.wait()
            single.append((t.inflate_ms(), t.elapsed_ms()))
            t.free()
    single_infl_ms = sorted(x[0] for x in single)[len(single) // 2]
    single_wave_ms = sorted(x[1] for x in single)[len(single) // 2]
    launches0 = zone.stats().kernel_launches
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    ev0.record(stream)
    th0 = time.perf_counter()
    tickets, submit_each_ms = [], []
    for _, w in waves[max(args.warmup, 3):]:
        ts = time.perf_counter()
        tickets.append(w.submit(sptr))
        submit_each_ms.append((time.perf_counter() - ts) * 1e3)
    host_submit_ms = (time.perf_counter() - th0) * 1e3 / max(1, len(tickets))
    ev1.record(stream)
    ev1.synchronize()
    sampler.stop_ev.set()
    ms = ev0.elapsed_time(ev1)
    infl_ms = [t.inflate_ms() for t in tickets]
    gath_ms = [t.gather_ms() for t in tickets]
    for t in tickets:
        t.wait()
        t.free()
    launches = zone.stats().kernel_launches - launches0
    raw_bytes = wave_n * wl.C3_RAW
    # the inflate kernel decodes every DISTINCT member of a wave once (duplicates are copies made by the gather)
    timed = waves[max(args.warmup, 3):]
    distinct = [len(set(int(r) for r in rr)) for rr, _ in timed]
    comp_wave = [sum(len(made[r][0]) for r in set(int(x) for x in rr)) for rr, _ in timed]
    ms_step = ms / args.steps
    hbm, which = peaks()
    i_ms = sum(infl_ms) / len(infl_ms)
    g_ms = sum(gath_ms) / len(gath_ms)
    dedupe = os.environ.get("NZ_INFL_DEDUPE", "1") != "0"
    decoded = (sum(distinct) / len(distinct) if dedupe else wave_n) * wl.C3_RAW
    achieved = ((sum(comp_wave) / len(comp_wave)) * (1 if dedupe else wave_n / (sum(distinct) / len(distinct))) + decoded) / (i_ms * 1e-3) / 1e9
    cpu = None
    if not args.no_cpu:
        import ctypes as C

        from oracle import oracle as o

        threads = os.cpu_count() or 1
        L = o.lib()
        m = min(wave_n, 4 * threads)
        outs = np.empty((m, wl.C3_RAW + 8), dtype=np.uint8)
        srcs = [np.frombuffer(made[r][0], dtype=np.uint8) for r in waves[0][0][:m]]
        jobs = (o.ZJob * m)()
        for j in range(m):
            jobs[j].src, jobs[j].n, jobs[j].dst, jobs[j].cap = srcs[j].ctypes.data, srcs[j].size, outs[j].ctypes.data, wl.C3_RAW + 8
        best = min(L.nzo_gunzip_jobs(jobs, m, threads) for _ in range(3))
        cpu = {"value": m * wl.C3_RAW / best / 1e9, "unit": "GB/s", "cores": threads, "kind": "port",
               "sample": f"{m} of the wave's {wave_n} members, zlib inflate (wbits 31), {threads} threads"}
    line = {
        "metric": "cached-GET GB/s (inflated bytes out)", "value": raw_bytes / (ms_step * 1e-3) / 1e9, "unit": "GB/s", "n_gpus": 1,
        "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": f"C3: {n_obj} x 1 MiB text-like objects stored gzip -6 ({wl.C3_RAW * n_obj / comp_total:.2f}:1), Zipf(0.99) access, "
                               f"{wave_n} GETs per wave, inflate on every hit", "objects": n_obj, "wave": wave_n,
                   "response_buffers": "HBM", "l2": f"{wave_n} MiB written per wave (> L2)"},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": hbm, "unit": "GB/s", "frac": achieved / hbm, "traffic": None,
                     "peak_source": which, "kernel": "k_inflate_seg", "kernel_ms": i_ms,
                     "note": "bit-serial DEFLATE: latency/ALU bound, not HBM bound; algorithmic bytes = compressed in + raw out of the members "
                             "the kernel decodes (each distinct member of the wave once)",
                     "kernel_share_of_step": i_ms / ms_step, "decoded_gbs": decoded / (i_ms * 1e-3) / 1e9},
        "c3": {"single_stream_ms": single_infl_ms, "single_get_wave_ms": single_wave_ms, "distinct_members_per_wave": sum(distinct) / len(distinct),
               "inflate_kernel_ms": i_ms, "fanout_gather_ms": g_ms, "host_submit_ms": host_submit_ms, "host_submit_each_ms": submit_each_ms, "dedupe": dedupe,
               "index_bytes_per_object": (arena_used - comp_total) / n_obj, "index_segment_kib": int(os.environ.get("NZ_INFL_SEG_KB", "128")),
               "indexed": os.environ.get("NZ_INFL_INDEX", "1") != "0"},
        "cpu_baseline": cpu, "e2e": None, "gpu_launches": int(launches), "clocks": sampler.summary(),
        "gen_s": gen_s, "fill_s": fill_s, "compressed_bytes_total": comp_total,
    }
    print(json.dumps(line), flush=True)
    dbuf.free()
    zone.close()
    return 0


# --------------------------------------------------------------------------- #
# C4: key-hash-sharded store, objects > 64 MiB striped over the shards, GETs that span shards
# --------------------------------------------------------------------------- #
def span_section(args, dist, rank, world, local):
    """BASELINE configs[3]: key-hash-sharded store of `--c4-gib-per-gpu` per GPU (the named config: 125 GiB/GPU = 1 TB at 8
    GPUs; clamped to what fits next to the response buffers), objects > 64 MiB striped round-robin over the shards, 256
    in-flight full GETs routed to their owners.  A GET of a striped object pulls the remote stripes with direct NVLink peer
    loads inside k_gather_tma (K5) - the exchange IS the gather kernel.  Beside it: an NCCL all-gather of one object striped
    64 MiB per rank (the collective north_star names).  Collective: every rank calls it; every rank gets the dict."""
    import numpy as np
    import torch

    import nezha_b200 as nz
    from nezha_b200 import capi, shard, span, workloads as wl

    want = int(args.c4_gib_per_gpu * (1 << 30))
    inflight = args.c4_inflight
    # what fits: arena + this rank's response buffers (in-flight GETs x mean object size, x1.6 for the draw) + 6 GiB for
    # NCCL, the CUDA context, staging
    free_b, _ = torch.cuda.mem_get_info()
    mean_obj = ((1 << 30) - (1 << 20)) / math.log(1024.0)
    resp_est = int(1.6 * inflight / world * mean_obj) + (2 << 30)
    fit = torch.tensor([max(free_b - resp_est - (6 << 30), 1 << 30)], device="cuda", dtype=torch.int64)
    dist.all_reduce(fit, op=dist.ReduceOp.MIN)
    per_gpu = min(want, int(fit.item()))
    total_target = per_gpu * world
    # population: log-uniform 1 MiB .. 1 GiB until 85 % of the arenas is reached (same list on every rank)
    g = np.random.Generator(np.random.Philox(key=4))
    sizes = []
    acc = 0
    while acc < total_target * 0.85:
        sz = int(math.exp(g.uniform(math.log(1 << 20), math.log(1 << 30))))
        sz = min(sz, int(total_target * 0.85) - acc)
        sizes.append(max(sz, 1))
        acc += sizes[-1]
    keys = [wl.key128(wl.object_key_string("c4", i)) for i in range(len(sizes))]
    zone = nz.Zone(device=local, arena_bytes=per_gpu + (256 << 20), max_objects=1 << 16, shard_index=rank, shard_count=world)
    zone.set_profiling(True)
    dbuf = None
    try:
        store = span.ShardedStore(zone, dist)
        store.attach_peers()
        pool = torch.empty(span.STRIPE_BYTES, dtype=torch.uint8, device="cuda")     # one stripe of scratch for the generator

        def gen_for(i):
            def gen(off, ln):
                tg = torch.Generator(device="cuda")
                tg.manual_seed(1_000_003 * i + off // (1 << 20))
                if ln <= pool.numel():
                    return pool[:ln].random_(0, 256, generator=tg)
                return torch.randint(0, 256, (ln,), dtype=torch.uint8, device="cuda", generator=tg)
            return gen

        t0 = time.perf_counter()
        crcs = {}
        for i, (k, sz) in enumerate(zip(keys, sizes)):
            c = store.put(k, sz, gen_for(i))
            if c is not None:
                crcs[i] = c
        torch.cuda.synchronize()
        dist.barrier()
        fill_s = time.perf_counter() - t0
        stored = zone.stats().arena_used
        # request log: `inflight` GETs over the whole population (seed 44), routed to the owners
        gq = np.random.Generator(np.random.Philox(key=44))
        picks = gq.integers(0, len(sizes), inflight)
        mine = [int(i) for i in picks if shard.shard_of(keys[int(i)], world) == rank]
        offs, total = [], 0
        for i in mine:
            offs.append(total)
            total += (sizes[i] + 255) // 256 * 256
        dbuf = zone.alloc_device(max(total, 256))
        wave = zone.prepare([nz.Get(keys[i], dbuf, [], 0, offs[j]) for j, i in enumerate(mine)]) if mine else None
        payload = float(sum(sizes[i] for i in mine))
        remote = 0.0          # NVLink bytes per wave: identical concurrent GETs are coalesced, so count distinct objects
        local_src = 0.0
        for i in sorted(set(mine)):
            for sh, _, ln in shard.stripe_plan(keys[i], sizes[i], world):
                if sh != rank:
                    remote += ln
                else:
                    local_src += ln
        stream = torch.cuda.Stream(device=local)
        torch.cuda.set_stream(stream)
        sptr = stream.cuda_stream

        def one():
            return wave.submit(sptr) if wave is not None else None

        for _ in range(max(args.warmup, 3)):
            t = one()
            if t:
                res = t.wait()
                t.free()
                if any(r.status != capi.NZ_OK for r in res):
                    raise RuntimeError(f"rank {rank} span wave failed {[r.status for r in res if r.status][:4]}")
        # spot parity: the CRC-32C of one served striped object equals the CRC recorded at fill
        big = [j for j, i in enumerate(mine) if sizes[i] > span.STRIPE_BYTES and i in crcs]
        if big:
            j = big[0]
            i = mine[j]
            if zone.crc32c_device(dbuf.dev + offs[j], sizes[i]) != crcs[i]:
                raise RuntimeError("gathered striped object differs from what was stored (CRC-32C)")
        steps = max(3, min(args.steps, 10))
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        dist.barrier()
        torch.cuda.synchronize()
        ev0.record(stream)
        tickets = [one() for _ in range(steps)]
        ev1.record(stream)
        ev1.synchronize()
        dist.barrier()
        ms = ev0.elapsed_time(ev1)
        g_ms = [t.gather_ms() for t in tickets if t]
        for t in tickets:
            if t:
                t.wait()
                t.free()
        k_ms = sum(g_ms) / len(g_ms) if g_ms else 0.0
        ms_all, payload_all = shard.reduce_timing(dist, ms, payload, device="cuda")
        _, remote_all = shard.reduce_timing(dist, ms, remote, device="cuda")
        ms_step = ms_all / steps
        value = payload_all / (ms_step * 1e-3) / 1e9
        # NVLink ingress of every rank: bytes pulled from peers / its own gather kernel time
        ingress = remote / (k_ms * 1e-3) / 1e9 if k_ms > 0 else 0.0
        ing = torch.tensor([ingress, k_ms, remote, payload], device="cuda", dtype=torch.float64)
        allr = [torch.zeros_like(ing) for _ in range(world)]
        dist.all_gather(allr, ing)
        per_rank = [{"rank": r, "ingress_gbs": float(x[0]), "gather_ms": float(x[1]), "nvlink_bytes": float(x[2]), "payload_bytes": float(x[3])}
                    for r, x in enumerate(allr)]
        best = max(p["ingress_gbs"] for p in per_rank)
        # the rank with the most NVLink bytes bounds the wave: its ingress is the figure of merit
        busiest = max(per_rank, key=lambda p: p["nvlink_bytes"])
        # ---- NCCL all-gather baseline on one evenly striped object ----------------
        nccl = None
        try:
            part = 64 << 20
            src = torch.randint(0, 256, (part,), dtype=torch.uint8, device="cuda")
            outb = torch.empty(part * world, dtype=torch.uint8, device="cuda")
            for _ in range(3):
                dist.all_gather_into_tensor(outb, src)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            dist.barrier()
            e0.record()
            reps = 10
            for _ in range(reps):
                dist.all_gather_into_tensor(outb, src)
            e1.record()
            e1.synchronize()
            tms = e0.elapsed_time(e1) / reps
            tt = torch.tensor([tms], device="cuda", dtype=torch.float64)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            nccl = {"object_bytes": part * world, "ms": float(tt.item()), "gbs_per_gpu_assembled": part * world / (float(tt.item()) * 1e-3) / 1e9,
                    "ingress_gbs_per_gpu": part * (world - 1) / (float(tt.item()) * 1e-3) / 1e9,
                    "note": "ncclAllGather of one object striped 64 MiB per rank: every rank ends with the whole object (then still needs a copy into the response)"}
            del src, outb
        except Exception as e:
            nccl = {"error": repr(e)}
        peak = 770.0
        return {
            "value": value, "unit": "GB/s", "ms_per_step": ms_step, "steps": steps,
            "config": {"workload": f"C4: key-hash-sharded store, {len(sizes)} objects 1 MiB-1 GiB (log-uniform), {per_gpu / 2**30:.1f} GiB arena/GPU filled to 85 % "
                                   f"(named config: 125 GiB/GPU; asked {args.c4_gib_per_gpu}), objects > 64 MiB striped round-robin, {inflight} in-flight full GETs routed to owners",
                       "objects": len(sizes), "bytes_total": int(acc), "arena_gib_per_gpu": per_gpu / 2**30, "stored_bytes_rank0": int(stored), "inflight": inflight,
                       "response_buffers": "HBM",
                       "exchange": "direct NVLink peer loads (cp.async.bulk from the peer's arena) inside k_gather_tma over CUDA-IPC-mapped arenas; no separate collective",
                       "tile_order": "spread over the wave (NZ_SPAN_SPREAD rows): concurrent CTAs pull from different peers"},
            "nvlink_bytes_over_payload": remote_all / payload_all if payload_all else 0.0, "nvlink_bytes_per_step": remote_all,
            "busiest_rank_ingress_gbs": busiest["ingress_gbs"], "max_rank_ingress_gbs": best, "nvlink_peak_gbs_per_dir": peak,
            "ingress_frac_of_peak": busiest["ingress_gbs"] / peak, "per_rank": per_rank,
            "nccl_allgather": nccl, "fill_s": fill_s, "fill_gbs": acc / fill_s / 1e9, "gpu_launches": 2 * steps,
        }
    finally:
        if dbuf is not None:
            dbuf.free()
        zone.close()
        torch.cuda.empty_cache()


def c5_fill_section(args, dist, rank, world, local):
    """BASELINE configs[4] at N > 1: every rank cold-fills its shard from files on the PV at the same time (cuFile/GDS first,
    else pread -> pinned bounce -> H2D).  Bounded: 2 GiB per rank.  Returns the aggregate rate (max over ranks' time)."""
    import shutil
    import tempfile

    import torch

    import nezha_b200 as nz
    from nezha_b200 import capi, workloads as wl

    plan = [(64, 16), (256, 4)]
    pv = tempfile.mkdtemp(prefix=f"nezha_pv_r{rank}_", dir=args.pv_dir)
    zone = None
    try:
        files = []
        for mib, cnt in plan:
            blob = memoryview(wl.synth_bytes(mib << 20, seed=5_000 + mib + rank))
            for j in range(cnt):
                k = wl.key128(wl.object_key_string(f"c5-r{rank}-{mib}", j))
                path = os.path.join(pv, k.hex())
                with open(path, "wb") as f:
                    f.write(blob)
                files.append((k, path, mib << 20))
        total = sum(f[2] for f in files)
        zone = nz.Zone(device=local, arena_bytes=total + (256 << 20), max_objects=1024, shard_index=0, shard_count=1)
        wpath = os.path.join(pv, "warm")
        with open(wpath, "wb") as f:
            f.write(b"\0" * 65536)
        fd = os.open(wpath, os.O_RDONLY)
        zone.put_file(wl.key128(b"c5/warm"), fd, 65536)      # pays cuFileDriverOpen once (a per-process start-up cost)
        os.close(fd)
        dist.barrier()
        t0 = time.perf_counter()

        def load(item):
            k, path, n = item
            fd = os.open(path, os.O_RDONLY)
            try:
                zone.put_file(k, fd, n)
            finally:
                os.close(fd)

        from concurrent.futures import ThreadPoolExecutor
        with ThreadPoolExecutor(4) as ex:
            list(ex.map(load, sorted(files, key=lambda f: -f[2])))
        dt = time.perf_counter() - t0
        st = zone.stats()
        tt = torch.tensor([dt], device="cuda", dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        cb = torch.tensor([float(st.file_cufile_bytes), float(st.file_bounce_bytes)], device="cuda", dtype=torch.float64)
        dist.all_reduce(cb)
        return {"workload": f"C5 cold fill at {world} GPUs: every rank loads {total / 2**30:.1f} GiB of files from {args.pv_dir} into its shard at the same time",
                "gbs": total * world / float(tt.item()) / 1e9, "seconds": float(tt.item()), "cufile_bytes": int(cb[0].item()), "bounce_bytes": int(cb[1].item()),
                "cufile_status": capi.load().nz_cufile_status().decode()}
    finally:
        if zone is not None:
            zone.close()
        shutil.rmtree(pv, ignore_errors=True)


def run_c4(args):
    """`--workload c4`: the span section on its own (needs torchrun, N >= 2)."""
    import torch
    import torch.distributed as dist

    rank, world, local = dist_env()
    torch.cuda.set_device(local)
    bind_to_gpu_numa_node(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    r = span_section(args, dist, rank, world, local)
    if rank == 0:
        line = {"metric": "cached-GET GB/s", "value": r["value"], "unit": "GB/s", "n_gpus": world, "steps": r["steps"], "warmup": max(args.warmup, 3),
                "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
                "config": r["config"], "span": {k: v for k, v in r.items() if k not in ("config", "value", "unit", "ms_per_step", "steps")},
                "gpu_launches": r["gpu_launches"]}
        print(json.dumps(line), flush=True)
    dist.destroy_process_group()
    return 0


# --------------------------------------------------------------------------- #
# C5: cold fill from files on the PV (cuFile/GDS, else pinned bounce), then a hit sweep 1 MiB .. 1 GiB
# --------------------------------------------------------------------------- #
def run_c5(args):
    import shutil
    import tempfile

    import numpy as np
    import torch

    import nezha_b200 as nz
    from nezha_b200 import capi, workloads as wl

    rank, world, local = dist_env()
    torch.cuda.set_device(local)
    bind_to_gpu_numa_node(local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    plan = [(1, 64), (4, 64), (16, 32), (64, 16), (256, 4), (1024, 2)]          # (MiB, objects) per size class
    pv = tempfile.mkdtemp(prefix=f"nezha_pv_r{rank}_", dir=args.pv_dir)
    zone = None
    try:
        files = []
        t0 = time.perf_counter()
        for mib, cnt in plan:
            for j in range(cnt):
                n = mib << 20
                k = wl.key128(wl.object_key_string(f"c5-r{rank}-{mib}", j))
                path = os.path.join(pv, k.hex())
                with open(path, "wb") as f:
                    f.write(memoryview(wl.synth_bytes(n, seed=5_000 + 100 * mib + j)))
                files.append((mib, k, path, n))
        write_s = time.perf_counter() - t0
        total = sum(f[3] for f in files)
        zone = nz.Zone(device=local, arena_bytes=total + (512 << 20), max_objects=4096, shard_index=0, shard_count=1)
        zone.set_profiling(True)
        # ---- cold fill: PV -> HBM ----
        # one throw-away file first: the first nz_put_file pays cuFileDriverOpen (seconds), which is a
        # per-process start-up cost, not a per-byte one; it is reported on its own
        wpath = os.path.join(pv, "warm")
        with open(wpath, "wb") as f:
            f.write(b"\0" * 65536)
        t0 = time.perf_counter()
        fd = os.open(wpath, os.O_RDONLY)
        zone.put_file(wl.key128(b"c5/warm"), fd, 65536)
        os.close(fd)
        driver_init_s = time.perf_counter() - t0
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()

        def load(item):                                # like nzproxy's cold start: several files at once, largest first;
            mib, k, path, n = item                     # nz_put_file hands every call its share of the bounce lanes
            fd = os.open(path, os.O_RDONLY)
            try:
                zone.put_file(k, fd, n)
            finally:
                os.close(fd)

        from concurrent.futures import ThreadPoolExecutor
        with ThreadPoolExecutor(8) as ex:
            list(ex.map(load, sorted(files, key=lambda f: -f[3])))
        fill_s = time.perf_counter() - t0
        st = zone.stats()
        # ---- hit sweep ----
        stream = torch.cuda.Stream(device=local)
        torch.cuda.set_stream(stream)
        sptr = stream.cuda_stream
        rng = np.random.default_rng(55 + rank)
        sweep = []
        for mib, cnt in plan:
            n = mib << 20
            gets_n = 100 if mib <= 256 else 32
            keys = [f[1] for f in files if f[0] == mib]
            picks = rng.integers(0, len(keys), gets_n)
            dbuf = zone.alloc_device(gets_n * n)
            wave = zone.prepare([nz.Get(keys[int(p)], dbuf, [], 0, j * n) for j, p in enumerate(picks)])
            for _ in range(3):
                t = wave.submit(sptr); res = t.wait(); t.free()
            if any(r.status != capi.NZ_OK for r in res):
                raise SystemExit("bench.py c5: wave failed")
            reps = max(3, min(args.steps, 20))
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record(stream)
            tk = [wave.submit(sptr) for _ in range(reps)]
            e1.record(stream)
            e1.synchronize()
            ms = e0.elapsed_time(e1) / reps
            for t in tk:
                t.wait(); t.free()
            distinct = len(set(int(p) for p in picks))
            sweep.append({"object_mib": mib, "gets": gets_n, "distinct_objects": distinct, "ms_per_wave": ms,
                          "payload_gbs": gets_n * n / (ms * 1e-3) / 1e9,
                          "hbm_traffic_gbs": (distinct + gets_n) * n / (ms * 1e-3) / 1e9})
            dbuf.free()
        fill_gbs = total / fill_s / 1e9
        if world > 1:
            tt = torch.tensor([fill_s], device="cuda", dtype=torch.float64)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            fill_gbs = total * world / float(tt.item()) / 1e9
            for row in sweep:
                v = torch.tensor([row["payload_gbs"]], device="cuda", dtype=torch.float64)
                dist.all_reduce(v)
                row["payload_gbs_all_gpus"] = float(v.item())
        if rank == 0:
            big = sweep[-1]
            line = {
                "metric": "cached-GET GB/s", "value": big.get("payload_gbs_all_gpus", big["payload_gbs"]), "unit": "GB/s", "n_gpus": world,
                "steps": args.steps, "warmup": 3, "ms_per_step": big["ms_per_wave"], "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "u8", "data": "synthetic",
                "config": {"workload": "C5: cold fill from files on the PV, then hit sweep 1 MiB..1 GiB (100 GETs per size, 32 for 1 GiB); value = the 1 GiB row",
                           "pv_dir": args.pv_dir, "bytes_per_gpu": total, "response_buffers": "HBM"},
                "fill": {"gbs": fill_gbs, "seconds": fill_s, "cufile_bytes": int(st.file_cufile_bytes), "bounce_bytes": int(st.file_bounce_bytes),
                         "note": "cuFile/GDS is tried first; on a file system without GDS support the bytes go pread -> pinned bounce -> cudaMemcpyAsync",
                         "file_write_s": write_s, "first_put_file_s": driver_init_s, "cufile_status": capi.load().nz_cufile_status().decode()},
                "sweep": sweep, "gpu_launches": 2 * sum(max(3, min(args.steps, 20)) for _ in plan),
            }
            print(json.dumps(line), flush=True)
    finally:
        if zone is not None:
            zone.close()
        shutil.rmtree(pv, ignore_errors=True)
        if world > 1:
            dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--object-bytes", type=int, default=1 << 30)
    ap.add_argument("--requests", type=int, default=64)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-http", action="store_true")
    ap.add_argument("--cpu-budget", type=float, default=12.0)
    ap.add_argument("--workload", default="c2", choices=["c2", "c3", "c4", "c5"])
    ap.add_argument("--pv-dir", default="/tmp")
    ap.add_argument("--c4-gib-per-gpu", type=float, default=125.0)
    ap.add_argument("--no-span", action="store_true")
    ap.add_argument("--no-sections", action="store_true")
    ap.add_argument("--c4-inflight", type=int, default=256)
    ap.add_argument("--c3-objects", type=int, default=10_000)
    ap.add_argument("--c3-wave", type=int, default=8192)
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    if args.workload == "c3":
        return run_c3(args)
    if args.workload == "c4":
        return run_c4(args)
    if args.workload == "c5":
        return run_c5(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
