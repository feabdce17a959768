"""-m gpu: pinned windows are placed on the host memory node next to the GPU (a window on the far socket
moves at about a third of the near one's rate), whatever thread allocates them — the C ABI's callers are goroutines /
connection threads that run anywhere."""
import ctypes as C
import os
import threading

import pytest

import nezha_b200 as nz

pytestmark = pytest.mark.gpu

SYS_move_pages = 279            # x86_64


def page_nodes(addr: int, nbytes: int, samples: int = 16):
    """node of `samples` pages of [addr, addr+nbytes) (move_pages with nodes=NULL only queries); None if refused"""
    libc = C.CDLL(None, use_errno=True)
    step = max(4096, (nbytes // samples) & ~4095)
    pages = [(addr & ~4095) + i * step for i in range(samples) if i * step < nbytes]
    arr = (C.c_void_p * len(pages))(*pages)
    status = (C.c_int * len(pages))()
    rc = libc.syscall(SYS_move_pages, 0, C.c_ulong(len(pages)), arr, None, status, 0)
    if rc != 0:
        return None
    return list(status)


def nodes_online():
    return sorted(int(d[4:]) for d in os.listdir("/sys/devices/system/node") if d.startswith("node") and d[4:].isdigit())


def test_pinned_windows_live_next_to_the_gpu():
    with nz.Zone(device=0, arena_bytes=64 << 20, max_objects=64) as z:
        node = z.numa_node()
        assert node >= -1
        if node < 0 or len(nodes_online()) < 2:
            pytest.skip(f"single host memory node (gpu node {node}, online {nodes_online()}): nothing to place")
        others = [n for n in nodes_online() if n != node]
        far_cpus = set()
        for n in others:
            txt = open(f"/sys/devices/system/node/node{n}/cpulist").read().strip()
            for part in txt.split(","):
                a, _, b = part.partition("-")
                far_cpus.update(range(int(a), int(b or a) + 1))
        far_cpus &= os.sched_getaffinity(0)
        got = {}

        def alloc_from_far_socket():
            if far_cpus:
                os.sched_setaffinity(0, far_cpus)            # this thread only: a caller running on the other socket
            b = z.alloc_host(8 << 20)
            got["nodes"] = page_nodes(b.host, 8 << 20)
            got["after"] = os.sched_getaffinity(0)
            b.free()

        t = threading.Thread(target=alloc_from_far_socket)
        t.start()
        t.join()
        if got["nodes"] is None:
            pytest.skip("move_pages refused in this container: cannot read page placement")
        if not all(n == node for n in got["nodes"]):
            # best effort by design: without set_mempolicy (seccomp) AND without a CPU of the GPU's node in this cgroup
            # there is no way to place the pages - that is the container's doing, not a failure of the library
            libc = C.CDLL(None, use_errno=True)
            mask = C.c_ulong(1 << node)
            policy_ok = libc.syscall(238, 1, C.byref(mask), 65) == 0          # SYS_set_mempolicy, MPOL_PREFERRED
            if policy_ok:
                libc.syscall(238, 0, None, 0)
            near = set()
            for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
                a, _, b = part.partition("-")
                near.update(range(int(a), int(b or a) + 1))
            if not policy_ok and not (near & os.sched_getaffinity(0)):
                pytest.skip("neither set_mempolicy nor a CPU of the GPU's node is available here: placement impossible")
        assert all(n == node for n in got["nodes"]), (node, got["nodes"])
        if far_cpus:
            assert got["after"] == far_cpus                   # the caller's affinity is handed back untouched


def test_placement_can_be_turned_off(monkeypatch):
    monkeypatch.setenv("NZ_NUMA", "0")
    with nz.Zone(device=0, arena_bytes=64 << 20, max_objects=64) as z:
        assert z.numa_node() == -1
        b = z.alloc_host(1 << 20)
        b.numpy()[:] = 7
        b.free()
