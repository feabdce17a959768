#!/usr/bin/env python
"""Why does (or does not) cuFile/GDS carry bytes on this box?  Tries nz_put_file on a few file systems and prints the
library's own account of the last attempt (nz_cufile_status), plus the mount types."""
import os, subprocess, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import nezha_b200 as nz
from nezha_b200 import capi, workloads as wl

print(subprocess.run("df -hT /tmp /dev/shm . 2>&1; lsmod 2>/dev/null | grep -i nvidia_fs; ls /dev/nvidia-fs* 2>&1 | head -3; cat /proc/driver/nvidia-fs/stats 2>&1 | head -5",
                     shell=True, capture_output=True, text=True).stdout)
d = wl.synth_bytes(256 << 20, seed=9)
for base in ("/tmp", "/dev/shm", os.getcwd()):
    p = os.path.join(base, "nz_cufile_probe.bin")
    try:
        with open(p, "wb") as f:
            f.write(memoryview(d))
        with nz.Zone(device=0, arena_bytes=1 << 30, max_objects=64) as z:
            fd = os.open(p, os.O_RDONLY)
            s0 = z.stats()
            t0 = time.perf_counter()
            z.put_file(wl.key128(b"probe/" + base.encode()), fd, len(d))
            dt = time.perf_counter() - t0
            os.close(fd)
            s1 = z.stats()
            print(base, f"{len(d) / dt / 1e9:.2f} GB/s", "cufile", s1.file_cufile_bytes - s0.file_cufile_bytes, "bounce", s1.file_bounce_bytes - s0.file_bounce_bytes,
                  "|", capi.load().nz_cufile_status().decode(), flush=True)
    finally:
        try:
            os.unlink(p)
        except OSError:
            pass
