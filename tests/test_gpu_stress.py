"""-m gpu: size and concurrency edges of the hot path.

* an object larger than 4 GiB (64-bit offsets through plan, tiles, TMA and register paths, CRC fold);
* many host threads submitting waves while others fill and evict (the C ABI must be callable
  concurrently from any OS thread — cgo moves goroutines between threads).
"""
import ctypes as C
import threading
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import pytest
import torch

import nezha_b200 as nz
from nezha_b200 import capi, workloads as wl
from oracle import oracle as o

pytestmark = pytest.mark.gpu


def test_object_larger_than_4gib():
    n = (5 << 30) + 12_345
    with nz.Zone(device=0, arena_bytes=n + (256 << 20), max_objects=64) as z:
        g = torch.Generator(device="cuda")
        g.manual_seed(99)
        data = torch.randint(0, 256, (n,), dtype=torch.uint8, device="cuda", generator=g)
        key = wl.key128(b"stress/5gib")
        crc = z.put(key, data)
        e = z.lookup([key])[0]
        assert e.status == capi.NZ_OK and e.stored_length == n
        host = data.cpu().numpy()
        assert crc == o.crc32c(host)                                   # chunk table + fold across > 65536 chunks
        four = 1 << 32
        cases = [(four - 100_000, 200_001), (four - 7, 15), (four, 1), (n - 3, 3), (four + 12_345, 70_001), ((3 << 30) + 5, (1 << 30) + 77)]
        for s, l in cases:
            res, body = z.get(key, segs=[(s, l)], device_dst=(l > (64 << 20)))
            assert res.status == capi.NZ_OK and res.bytes == l
            assert body == host[s:s + l].tobytes(), (s, l)
        # whole object into a device buffer, checked by CRC-32C on the device (no 5 GiB download)
        buf = z.alloc_device(n)
        t = z.submit([nz.Get(key, buf)])
        r = t.wait()[0]
        t.free()
        assert r.status == capi.NZ_OK and r.bytes == n
        assert z.crc32c_device(buf.dev, n) == crc
        # verify-on-GET over a range that straddles 4 GiB
        res, body = z.get(key, segs=[(four - 65_536 - 1, 131_074)], flags=capi.NZ_GET_VERIFY)
        assert res.status == capi.NZ_OK and body == host[four - 65_537:four + 65_537].tobytes()
        buf.free()
        del data


def test_concurrent_submitters_with_fill_and_evict():
    with nz.Zone(device=0, arena_bytes=1 << 30, max_objects=4096) as z:
        stable = {}
        for i in range(24):
            d = wl.synth_bytes(50_000 + 7_777 * i, seed=300 + i)
            k = wl.key128(b"stress/stable/%d" % i)
            z.put(k, d)
            stable[k] = d
        keys = list(stable)
        stop = threading.Event()
        errors = []

        def churn():                                     # fill + evict other keys while readers run
            j = 0
            try:
                while not stop.is_set():
                    k = wl.key128(b"stress/churn/%d" % (j % 40))
                    d = wl.synth_bytes(10_000 + (j % 13) * 1_000, seed=j)
                    try:
                        z.put(k, d)
                    except nz.NezhaError as e:
                        if e.code != capi.NZ_EEXIST:
                            raise
                    if j % 3 == 2:
                        z.evict(k)
                    j += 1
            except Exception as e:                       # pragma: no cover
                errors.append(repr(e))

        def reader(seed):
            rng = np.random.default_rng(seed)
            try:
                for _ in range(60):
                    picks = [keys[int(x)] for x in rng.integers(0, len(keys), 6)]
                    bufs, gets, want = [], [], []
                    for k in picks:
                        d = stable[k]
                        s = int(rng.integers(0, len(d) - 10)); l = int(rng.integers(1, len(d) - s))
                        b = z.alloc_host(l) if rng.random() < 0.5 else z.alloc_device(l)
                        bufs.append(b); gets.append(nz.Get(k, b, [(s, l, 0)], capi.NZ_GET_VERIFY if rng.random() < 0.3 else 0))
                        want.append(d[s:s + l].tobytes())
                    t = z.submit(gets)
                    res = t.wait()
                    t.free()
                    for r, b, w in zip(res, bufs, want):
                        if r.status != capi.NZ_OK or b.bytes(0, len(w)) != w:
                            errors.append(f"mismatch status={r.status}")
                        b.free()
            except Exception as e:                       # pragma: no cover
                errors.append(repr(e))

        ch = threading.Thread(target=churn)
        ch.start()
        with ThreadPoolExecutor(12) as ex:
            list(ex.map(reader, range(12)))
        stop.set()
        ch.join()
        assert not errors, errors[:5]
        st = z.stats()
        assert st.gets >= 12 * 60 * 6 and st.corrupt == 0
