"""-m gpu: the CUDA path, called through the C ABI, against the oracle on the same
seeded inputs — bit-exact (byte/integer work).  SPEC rule ids from oracle/SPEC.md.
"""
import os

import numpy as np
import pytest

import nezha_b200 as nz
from nezha_b200 import capi, workloads as wl
from oracle import oracle as o

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def zone():
    z = nz.Zone(device=0, arena_bytes=3 << 30, max_objects=1 << 12)
    yield z
    z.close()


def K(i, cfg="t"):
    return wl.key128(wl.object_key_string(cfg, i))


SIZES = [0, 1, 15, 16, 17, 127, 128, 129, 4095, 4096, 32767, 32768, 32769, 65535, 65536, 65537, 100_000, 1_000_003, 5 * 65536 + 7]


@pytest.fixture(scope="module")
def objects(zone):
    objs = {}
    for i, n in enumerate(SIZES):
        d = wl.synth_bytes(n, seed=100 + i)
        crc = zone.put(K(i), d)
        objs[i] = (d, crc)
    return objs


def test_put_crc_matches_oracle_X1_X2(zone, objects):
    for i, (d, crc) in objects.items():
        assert crc == o.crc32c(d), (i, len(d))
    ext = zone.lookup([K(i) for i in objects])
    for i, e in zip(objects, ext):
        assert e.status == capi.NZ_OK and e.stored_length == len(objects[i][0]) == e.raw_length
        assert e.crc32c == objects[i][1] and e.flags == capi.NZ_ENC_IDENTITY


def test_lookup_miss_and_hit_K1(zone, objects):
    keys = [K(0), K(10_000), K(3), os.urandom(16)]
    ext = zone.lookup(keys)
    assert [e.status for e in ext] == [capi.NZ_OK, capi.NZ_ENOTFOUND, capi.NZ_OK, capi.NZ_ENOTFOUND]
    # many keys in one launch (several probe groups per warp)
    many = [K(i % len(SIZES)) if i % 3 else os.urandom(16) for i in range(1000)]
    ext = zone.lookup(many)
    for i, e in enumerate(ext):
        assert (e.status == capi.NZ_OK) == bool(i % 3)


@pytest.mark.parametrize("device_dst", [False, True])
def test_full_get_bit_exact_B1(zone, objects, device_dst):
    for i, (d, crc) in objects.items():
        res, body = zone.get(K(i), device_dst=device_dst)
        assert res.status == capi.NZ_OK and res.bytes == len(d) == res.length and res.crc32c == crc
        assert body == o.serve_slice_mem(d, 0, len(d))


def test_single_ranges_all_alignments_R8(zone, objects):
    """Range starts at every byte phase (source misalignment 0..15 vs the 16 B vector path)."""
    d, _ = objects[SIZES.index(1_000_003)]
    key = K(SIZES.index(1_000_003))
    rng = np.random.default_rng(9)
    cases = [(s, l) for s in range(0, 18) for l in (1, 15, 16, 17, 100, 4097)]
    cases += [(int(rng.integers(0, len(d) - 70000)), int(rng.integers(1, 70000))) for _ in range(60)]
    cases += [(len(d) - 1, 1), (0, len(d)), (1, len(d) - 1), (32768 - 5, 10), (65536 - 1, 2)]
    for s, l in cases:
        res, body = zone.get(key, segs=[(s, l)])
        assert res.status == capi.NZ_OK and res.bytes == l
        assert body == o.serve_slice_mem(d, s, l), (s, l)


def test_ranges_to_misaligned_destination(zone, objects):
    """dst_off at arbitrary byte phases (multipart bodies put slices after part headers)."""
    d, _ = objects[SIZES.index(1_000_003)]
    key = K(SIZES.index(1_000_003))
    buf = zone.alloc_host(400_000)
    try:
        for dphase in range(0, 17):
            for sphase in (0, 1, 7, 16, 21):
                view = buf.numpy()
                view[:] = 0xEE
                segs = [(1000 + sphase, 70_001, dphase), (500_000, 33_000, 100_000 + dphase), (5, 3, 200_000 + dphase)]
                t = zone.submit([nz.Get(key, buf, segs)])
                r = t.wait()[0]
                t.free()
                assert r.status == capi.NZ_OK and r.bytes == sum(s[1] for s in segs)
                for so, ln, do in segs:
                    assert view[do:do + ln].tobytes() == d[so:so + ln].tobytes(), (dphase, sphase, so)
                    assert view[do - 1] == 0xEE if do else True
                    assert view[do + ln] == 0xEE        # no write past a slice
    finally:
        buf.free()


def test_multipart_plan_served_R9(zone, objects):
    """Whole oracle plan -> product: host writes the literals, the device gathers the slices."""
    idx = SIZES.index(5 * 65536 + 7)
    d, _ = objects[idx]
    ob = o.Obj(len(d), "application/octet-stream", '"abc"', "Wed, 07 Nov 2018 15:20:01 GMT")
    for spec in ["bytes=0-0,-1", "bytes=5-70000,65536-131071,300000-", "bytes=1-2,4-5,7-8,10-11,13-14"]:
        plan = o.parse_plan(o.plan(o.Req(range=spec), ob, boundary=7))
        assert plan.status == 206
        want = plan.body_bytes(d.tobytes())
        buf = zone.alloc_host(len(want))
        try:
            segs, off = [], 0
            view = buf.numpy()
            for kind, v in plan.body:
                if kind == "lit":
                    view[off:off + len(v)] = np.frombuffer(v, dtype=np.uint8)
                    off += len(v)
                else:
                    segs.append((v[0], v[1], off))
                    off += v[1]
            t = zone.submit([nz.Get(K(idx), buf, segs)])
            r = t.wait()[0]
            t.free()
            assert r.status == capi.NZ_OK
            assert view[:off].tobytes() == want
            assert int(dict(plan.headers)["Content-Length"]) == off
        finally:
            buf.free()


def test_wave_of_mixed_requests(zone, objects):
    """64 concurrent GETs in one wave: hits, misses, bad ranges, small buffers."""
    rng = np.random.default_rng(3)
    gets, expect, bufs = [], [], []
    for j in range(64):
        i = int(rng.integers(0, len(SIZES)))
        d = objects[i][0]
        kind = j % 8
        if kind == 0:                       # miss
            b = zone.alloc_device(64)
            gets.append(nz.Get(os.urandom(16), b)); expect.append((capi.NZ_ENOTFOUND, b""))
        elif kind == 1 and len(d) > 1:      # range outside the object
            b = zone.alloc_device(64)
            gets.append(nz.Get(K(i), b, [(len(d) - 1, 2, 0)])); expect.append((capi.NZ_EINVAL, b""))
        elif kind == 2 and len(d) > 8:      # destination too small
            b = zone.alloc_device(8)
            gets.append(nz.Get(K(i), b)); expect.append((capi.NZ_ERANGE, b""))
        elif kind in (3, 4) and len(d) > 2:
            s = int(rng.integers(0, len(d) - 1)); l = int(rng.integers(1, len(d) - s))
            b = zone.alloc_device(l)
            gets.append(nz.Get(K(i), b, [(s, l, 0)])); expect.append((capi.NZ_OK, d[s:s + l].tobytes()))
        else:
            b = zone.alloc_device(max(len(d), 1))
            gets.append(nz.Get(K(i), b)); expect.append((capi.NZ_OK, d.tobytes()))
        bufs.append(b)
    zone.set_profiling(True)
    t = zone.submit(gets)
    res = t.wait()
    assert t.elapsed_ms() >= 0 and t.gather_ms() >= 0
    t.free()
    zone.set_profiling(False)
    for r, (st, body), b in zip(res, expect, bufs):
        assert r.status == st
        if st == capi.NZ_OK:
            assert r.bytes == len(body) and b.bytes(0, len(body)) == body
        else:
            assert r.bytes == 0
    for b in bufs:
        b.free()


def test_verify_flag_and_corruption_X4(zone):
    d = wl.synth_bytes(300_000, seed=77)
    key = K(900)
    crc = zone.put(key, d, expected_crc32c=o.crc32c(d))
    assert crc == o.crc32c(d)
    res, body = zone.get(key, flags=capi.NZ_GET_VERIFY)
    assert res.status == capi.NZ_OK and body == d.tobytes()
    res, body = zone.get(key, segs=[(70_000, 10)], flags=capi.NZ_GET_VERIFY)
    assert res.status == capi.NZ_OK and body == d[70_000:70_010].tobytes()
    assert zone.debug_poke(key, 200_000, 0x40) == capi.NZ_OK      # flip a bit in chunk 3
    res, _ = zone.get(key, segs=[(0, 65536)], flags=capi.NZ_GET_VERIFY)
    assert res.status == capi.NZ_OK                                # chunk 0 untouched
    res, _ = zone.get(key, segs=[(199_000, 2000)], flags=capi.NZ_GET_VERIFY)
    assert res.status == capi.NZ_ECORRUPT
    res, _ = zone.get(key, flags=capi.NZ_GET_VERIFY)
    assert res.status == capi.NZ_ECORRUPT
    res, body = zone.get(key)                                      # without verify: bytes as stored (reference behaviour)
    assert res.status == capi.NZ_OK and body != d.tobytes()
    assert zone.evict(key) == capi.NZ_OK


def test_commit_rejects_wrong_crc(zone):
    d = wl.synth_bytes(10_000, seed=5)
    with pytest.raises(nz.NezhaError) as ei:
        zone.put(K(901), d, expected_crc32c=o.crc32c(d) ^ 1)
    assert ei.value.code == capi.NZ_ECORRUPT
    assert zone.lookup([K(901)])[0].status == capi.NZ_ENOTFOUND    # nothing published
    zone.put(K(901), d, expected_crc32c=o.crc32c(d))               # and the key is free again
    with pytest.raises(nz.NezhaError) as ei:
        zone.put(K(901), d)
    assert ei.value.code == capi.NZ_EEXIST
    assert zone.evict(K(901)) == capi.NZ_OK
    assert zone.evict(K(901)) == capi.NZ_ENOTFOUND
    assert zone.lookup([K(901)])[0].status == capi.NZ_ENOTFOUND


def test_crc32c_device_kernel_K3(zone):
    import torch
    rng = np.random.default_rng(11)
    for n in [1, 255, 256, 257, 65535, 65536, 65537, 1_000_000, 16 * 65536]:
        h = rng.integers(0, 256, n, dtype=np.uint8)
        t = torch.from_numpy(h).cuda()
        crc, chunks = zone.crc32c_device(t.data_ptr(), n, want_chunks=True)
        ochunks, owhole = o.crc32c_chunks(h, 65536)
        assert crc == owhole == o.crc32c(h)
        assert np.array_equal(chunks, ochunks)
        if n > 17:   # unaligned device pointer takes the byte path
            crc2 = zone.crc32c_device(t.data_ptr() + 1, n - 1)
            assert crc2 == o.crc32c(h[1:])


def test_K3_and_commit_crc_against_the_cpu_crc32_instruction(zone):
    """X1 with a witness nobody here wrote: the x86 `crc32` instruction (tests/hwcrc.py), at sizes the oracle's own loop
    would take long on - whole CRC, the 64 KiB chunk table, an odd base address, and the CRC a put commits."""
    import torch

    import hwcrc
    if hwcrc.lib() is None:
        pytest.skip("no SSE4.2 crc32 instruction / gcc here")
    torch.manual_seed(7)
    for n in (1, 65536, 65537, 5 * 65536 + 7, (64 << 20) + 13, 256 << 20):
        t = torch.randint(0, 256, (n,), dtype=torch.uint8, device="cuda")
        host = t.cpu().numpy()
        crc, chunks = zone.crc32c_device(t.data_ptr(), n, want_chunks=True)
        assert crc == hwcrc.crc32c(host), n
        for c in {0, len(chunks) // 2, len(chunks) - 1}:
            assert int(chunks[c]) == hwcrc.crc32c(host[c * 65536:(c + 1) * 65536]), (n, c)
        if n > 1:
            assert zone.crc32c_device(t.data_ptr() + 1, n - 1) == hwcrc.crc32c(host[1:]), n
    d = wl.synth_bytes(3_000_017, seed=77)
    assert zone.put(K(9001), d) == hwcrc.crc32c(d)


def test_table_churn_insert_evict(zone):
    """Tombstones + re-insert keep every live key reachable (K1/K6)."""
    live = {}
    rng = np.random.default_rng(21)
    for rnd in range(3):
        for i in range(200):
            k = K(5000 + rnd * 1000 + i, "churn")
            d = wl.synth_bytes(int(rng.integers(1, 3000)), seed=i + rnd)
            zone.put(k, d)
            live[k] = d
        victims = list(live)[::3]
        for k in victims:
            assert zone.evict(k) == capi.NZ_OK
            del live[k]
        ext = zone.lookup(list(live) + victims)
        assert all(e.status == capi.NZ_OK for e in ext[:len(live)])
        assert all(e.status == capi.NZ_ENOTFOUND for e in ext[len(live):])
    for k, d in list(live.items())[:50]:
        res, body = zone.get(k)
        assert res.status == capi.NZ_OK and body == d.tobytes()
    for k in live:
        zone.evict(k)


def test_c1_reference_object_roundtrip(zone):
    """BASELINE config C1: the 169,001,437-byte object under the reference's own key;
    size-independent properties at full size: CRC-of-chunks == whole CRC, slices == source."""
    d = wl.synth_bytes(wl.C1_LENGTH, seed=1)
    key = wl.key128(wl.C1_KEY)
    assert key.hex() == "9c30938a380ef9a0ce26538a0f68408e"
    crc = zone.put(key, d)
    assert crc == o.crc32c(d)
    res, body = zone.get(key, device_dst=True)
    assert res.status == capi.NZ_OK and res.bytes == wl.C1_LENGTH
    assert o.crc32c(body) == crc and body[-100:] == d[-100:].tobytes()
    for s, l in [(0, 1), (wl.C1_LENGTH - 1, 1), (123_456_789, 1 << 20), (7, 33_333_333)]:
        res, body = zone.get(key, segs=[(s, l)])
        assert body == d[s:s + l].tobytes()
    assert zone.evict(key) == capi.NZ_OK


def test_put_file_fill_path(zone, tmp_path):
    d = wl.synth_bytes(40 << 20, seed=8)      # spans the 16 MiB bounce buffers
    f = tmp_path / "obj.bin"
    f.write_bytes(d.tobytes())
    fd = os.open(f, os.O_RDONLY)
    try:
        crc = zone.put_file(K(902), fd, len(d))
    finally:
        os.close(fd)
    assert crc == o.crc32c(d)
    res, body = zone.get(K(902), segs=[(16 << 20, 1 << 20)])
    assert body == d[16 << 20:17 << 20].tobytes()
    zone.evict(K(902))
    st = zone.stats()
    assert st.file_cufile_bytes + st.file_bounce_bytes >= len(d)      # one of the two paths carried the file
    # unaligned file offset / length / destination offset through whichever path is active
    fd = os.open(f, os.O_RDONLY)
    try:
        k = K(903)
        import ctypes as C
        kk = (C.c_uint8 * 16).from_buffer_copy(k)
        p = C.c_void_p()
        n = 1_000_003
        capi.check(zone.L.nz_put_begin(zone.ctx, kk, n, n, 0, 0, C.byref(p)), "put_begin")
        capi.check(zone.L.nz_put_file(p, fd, 12_345, 0, 500_001), "put_file a")
        capi.check(zone.L.nz_put_file(p, fd, 12_345 + 500_001, 500_001, n - 500_001), "put_file b")
        crc = C.c_uint32()
        capi.check(zone.L.nz_put_commit(p, 0, 0, C.byref(crc)), "commit")
    finally:
        os.close(fd)
    assert crc.value == o.crc32c(d[12_345:12_345 + n])
    res, body = zone.get(K(903))
    assert body == d[12_345:12_345 + n].tobytes()
    zone.evict(K(903))


def test_sharded_zone_rejects_foreign_keys():
    z = nz.Zone(device=0, arena_bytes=64 << 20, max_objects=256, shard_index=1, shard_count=4)
    try:
        mine = [K(i, "sh") for i in range(64) if z.owns(K(i, "sh"))]
        other = [K(i, "sh") for i in range(64) if not z.owns(K(i, "sh"))]
        assert mine and other
        z.put(mine[0], b"abc")
        with pytest.raises(nz.NezhaError) as ei:
            z.put(other[0], b"abc")
        assert ei.value.code == capi.NZ_EINVAL
        st = [e.status for e in z.lookup([mine[0], other[0]])]
        assert st == [capi.NZ_OK, capi.NZ_ENOTFOUND]
    finally:
        z.close()


def test_stats_and_launch_count(zone, objects):
    s0 = zone.stats()
    zone.get(K(1))
    s1 = zone.stats()
    assert s1.gets == s0.gets + 1 and s1.hits == s0.hits + 1
    assert s1.kernel_launches > s0.kernel_launches


def test_coalesced_identical_gets(zone, objects):
    """Many concurrent GETs of the same bytes (the usual dataset-cache pattern) are chained behind
    one leader slice: read once, written to every destination.  Every destination must still get
    exactly the oracle's bytes, whatever its alignment class or memory kind."""
    idx = SIZES.index(1_000_003)
    d, _ = objects[idx]
    key = K(idx)
    n = len(d)
    gets, checks, bufs = [], [], []
    for j in range(40):                                   # identical full GETs: device and host destinations
        b = zone.alloc_device(n) if j % 3 else zone.alloc_host(n)
        bufs.append(b)
        gets.append(nz.Get(key, b))
        checks.append((b, 0, 0, n))
    for j in range(24):                                   # identical misaligned range, two destination phases
        b = zone.alloc_host(400_000)
        bufs.append(b)
        phase = 0 if j % 2 else 37
        gets.append(nz.Get(key, b, [(12_345, 333_333, phase)]))
        checks.append((b, phase, 12_345, 333_333))
    for j in range(8):                                    # a multi-slice request repeated (multipart layout)
        b = zone.alloc_device(300_000)
        bufs.append(b)
        gets.append(nz.Get(key, b, [(5, 70_000, 11), (500_000, 33_001, 100_000), (999_999, 4, 200_000)]))
        checks += [(b, 11, 5, 70_000), (b, 100_000, 500_000, 33_001), (b, 200_000, 999_999, 4)]
    b = zone.alloc_device(n)                              # and a different object in the same wave
    bufs.append(b)
    other = SIZES.index(100_000)
    gets.append(nz.Get(K(other), b))
    t = zone.submit(gets)
    res = t.wait()
    t.free()
    assert all(r.status == capi.NZ_OK for r in res)
    assert res[0].bytes == n and res[40].bytes == 333_333 and res[64].bytes == 70_000 + 33_001 + 4
    for buf, doff, soff, ln in checks:
        assert buf.bytes(doff, ln) == o.serve_slice_mem(d, soff, ln), (doff, soff, ln)
    assert b.bytes(0, 100_000) == objects[other][0].tobytes()
    for x in bufs:
        x.free()


def test_probe_chains_collisions_and_full_table():
    """Every key lands on the same table line (crafted key[8:16]): linear probing walks several
    128 B lines, tombstones are skipped on lookup and reused on insert, and the load limit answers NZ_EFULL."""
    z = nz.Zone(device=0, arena_bytes=32 << 20, max_objects=32, table_slots=64)     # 16 lines of 4 slots
    try:
        def key(i):
            return bytes([i, 0xAB, 0, 0, 0, 0, 0, 0]) + (0x5555_0000_0000_0003).to_bytes(8, "little")   # same line, distinct key128
        data = {i: wl.synth_bytes(1000 + i, seed=700 + i) for i in range(32)}
        for i in range(32):
            z.put(key(i), data[i])
        with pytest.raises(nz.NezhaError) as ei:                                     # 32 live in 64 slots: load limit 0.5
            z.put(key(40), data[0])
        assert ei.value.code in (capi.NZ_EFULL, capi.NZ_ENOMEM)
        assert all(e.status == capi.NZ_OK and e.stored_length == 1000 + i for i, e in enumerate(z.lookup([key(i) for i in range(32)])))
        for i in range(0, 32, 3):                                                    # tombstones inside the chain
            assert z.evict(key(i)) == capi.NZ_OK
        ext = z.lookup([key(i) for i in range(32)] + [key(99)])
        for i in range(32):
            assert (ext[i].status == capi.NZ_OK) == (i % 3 != 0), i
        assert ext[32].status == capi.NZ_ENOTFOUND
        for i in range(0, 32, 3):                                                    # re-insert: tombstones are reused
            z.put(key(i), data[i])
        for i in range(32):
            res, body = z.get(key(i))
            assert res.status == capi.NZ_OK and body == data[i].tobytes(), i
        # a wave that mixes all of them with misses
        bufs = [z.alloc_device(2000) for _ in range(40)]
        gets = [nz.Get(key(i) if i < 32 else key(100 + i), bufs[i]) for i in range(40)]
        t = z.submit(gets)
        res = t.wait()
        t.free()
        assert [r.status for r in res] == [capi.NZ_OK] * 32 + [capi.NZ_ENOTFOUND] * 8
        for b in bufs:
            b.free()
    finally:
        z.close()


def test_paced_host_waves_same_bytes_bounded_rate():
    """nz_set_host_rate: waves that only write pinned host memory are paced under the PCIe link rate (so a
    first window behind them sees a short queue); the bytes are the same, urgent waves are never paced,
    device destinations are never paced, and 0/0 restores the defaults."""
    import time
    z = nz.Zone(device=0, arena_bytes=512 << 20, max_objects=16)
    n = 96 << 20
    d = wl.synth_bytes(n, seed=4242)
    key = K(0, "pace")
    z.put(key, d)
    wins = [z.alloc_host(8 << 20) for _ in range(12)]
    gets = [nz.Get(key, w, [(i * (8 << 20), 8 << 20, 0)]) for i, w in enumerate(wins)]

    def wave(gs):
        t0 = time.perf_counter()
        t = z.submit(gs)
        res = t.wait()
        dt = time.perf_counter() - t0
        t.free()
        assert all(r.status == capi.NZ_OK for r in res)
        return dt

    wave(gets)
    fast = min(wave(gets) for _ in range(3))
    z.set_host_rate(8.0, 2)                                  # 96 MiB at 8 GB/s: >= 12 ms
    slow = min(wave(gets) for _ in range(3))
    for i, w in enumerate(wins):
        assert w.bytes(0, 8 << 20) == d[i * (8 << 20):(i + 1) * (8 << 20)].tobytes()
    assert slow >= 0.8 * (n / 8e9), (fast, slow)
    assert slow > 2 * fast, (fast, slow)
    urgent = min(wave([nz.Get(key, w, [(i * (8 << 20), 8 << 20, 0)], capi.NZ_GET_URGENT) for i, w in enumerate(wins)]) for _ in range(3))
    assert urgent < 0.6 * slow, (urgent, slow)
    z.set_host_rate(0.0, 0)
    again = min(wave(gets) for _ in range(3))
    assert again < 0.6 * slow, (again, slow)
    with pytest.raises(capi.NezhaError):
        z.set_host_rate(-1.0, 0)
    for w in wins:
        w.free()
    z.close()


def test_staged_dma_host_waves_bit_exact():
    """Waves of 8 MiB .. 1 GiB to pinned host memory are gathered into an HBM staging slot and moved out by
    the copy engine, slice by slice: bytes identical to the oracle's slices, gaps between slices untouched,
    coalesced duplicates fine, and the ring of staging slots survives many waves in flight."""
    z = nz.Zone(device=0, arena_bytes=1 << 30, max_objects=64)
    objs = []
    for i, n in enumerate([40 << 20, (9 << 20) + 12345, 3_000_001]):
        d = wl.synth_bytes(n, seed=9_000 + i)
        k = K(i, "dma")
        z.put(k, d)
        objs.append((k, d))
    s0 = z.stats()
    # (1) one request, three slices with gaps (a multipart body: the host wrote part headers into the gaps)
    k, d = objs[0]
    b = z.alloc_host(40 << 20)
    view = b.numpy()
    view[:] = 0xEE
    segs = [(5, 6 << 20, 100), (20 << 20, (7 << 20) + 3, (6 << 20) + 300), ((39 << 20) + 1, (1 << 20) - 1, (14 << 20) + 1000)]
    t = z.submit([nz.Get(k, b, segs)])
    (r,) = t.wait()
    t.free()
    assert r.status == capi.NZ_OK and r.bytes == sum(l for _, l, _ in segs)
    for s, l, off in segs:
        assert b.bytes(off, l) == d[s:s + l].tobytes()
    assert bytes(view[:100]) == b"\xEE" * 100 and bytes(view[(6 << 20) + 100:(6 << 20) + 300]) == b"\xEE" * 200     # gaps untouched
    # (2) many waves in flight over the slot ring, mixed objects, duplicates (coalesced), different sizes
    wins = [z.alloc_host(4 << 20) for _ in range(24)]
    tickets, expect = [], []
    rng = np.random.default_rng(5)
    for w in range(10):
        gets, exp = [], []
        for j in range(len(wins) // 2):
            k, d = objs[int(rng.integers(0, 3))]
            l = min(int(rng.integers(1 << 20, 4 << 20)), len(d))
            s = int(rng.integers(0, len(d) - l + 1)) if j % 3 else 0
            win = wins[(w % 2) * 12 + j]
            gets.append(nz.Get(k, win, [(s, l, 0)]))
            exp.append((win, d, s, l))
        tickets.append(z.submit(gets))
        expect.append(exp)
        if w >= 1:                                  # two waves in flight; windows of wave w-1 are checked before wave w+1 reuses them
            res = tickets[w - 1].wait()
            assert all(x.status == capi.NZ_OK for x in res)
            for win, d, s, l in expect[w - 1]:
                assert win.bytes(0, l) == d[s:s + l].tobytes()
            tickets[w - 1].free()
    res = tickets[-1].wait()
    assert all(x.status == capi.NZ_OK for x in res)
    for win, d, s, l in expect[-1]:
        assert win.bytes(0, l) == d[s:s + l].tobytes()
    tickets[-1].free()
    s1 = z.stats()
    assert s1.host_dma_waves - s0.host_dma_waves == 11
    # (3) verify + a miss in the same staged wave: the miss reports NZ_ENOTFOUND, the others are served and verified
    gets = [nz.Get(objs[1][0], wins[0], [(0, 4 << 20, 0)], capi.NZ_GET_VERIFY), nz.Get(K(99, "dma"), wins[1], [(0, 4 << 20, 0)]),
            nz.Get(objs[0][0], wins[2], [(1 << 20, 4 << 20, 0)], capi.NZ_GET_VERIFY)]
    t = z.submit(gets)
    res = t.wait()
    t.free()
    assert [x.status for x in res] == [capi.NZ_OK, capi.NZ_ENOTFOUND, capi.NZ_OK]
    assert wins[0].bytes(0, 4 << 20) == objs[1][1][:4 << 20].tobytes() and wins[2].bytes(0, 4 << 20) == objs[0][1][1 << 20:5 << 20].tobytes()
    for w in wins + [b]:
        w.free()
    z.close()


def test_concurrent_put_file_shares_bounce_lanes(tmp_path):
    """A cold start loads many files at once: concurrent nz_put_file calls take bounce lanes from one pool
    (large files several, small files one) and every object must arrive intact (CRC-32C vs the oracle)."""
    from concurrent.futures import ThreadPoolExecutor
    z = nz.Zone(device=0, arena_bytes=1 << 30, max_objects=64)
    sizes = [70 << 20, 1, 4096, (33 << 20) + 5, 1 << 20, 123_457, 48 << 20, (16 << 20) - 1, 9 << 20, 65_536, 200 << 20, 3]
    items = []
    for i, n in enumerate(sizes):
        d = wl.synth_bytes(n, seed=12_000 + i)
        path = tmp_path / f"f{i}"
        path.write_bytes(d.tobytes())
        items.append((K(i, "cpf"), str(path), n, d))

    def load(it):
        k, path, n, d = it
        fd = os.open(path, os.O_RDONLY)
        try:
            return z.put_file(k, fd, n)
        finally:
            os.close(fd)

    with ThreadPoolExecutor(6) as ex:
        crcs = list(ex.map(load, items))
    for (k, path, n, d), crc in zip(items, crcs):
        assert crc == o.crc32c(d.tobytes())
    res, body = z.get(items[3][0], segs=[(1 << 20, 1 << 20)])
    assert res.status == capi.NZ_OK and body == items[3][3][1 << 20:2 << 20].tobytes()
    st = z.stats()
    assert st.file_cufile_bytes + st.file_bounce_bytes == sum(sizes)
    z.close()


def test_put_packed_bulk_fill_of_small_objects():
    """nz_put_packed: hundreds of objects of every small size and byte phase through one pinned window - one copy, one scatter
    launch, batched CRC-32C.  Every object is then found, checksummed like the oracle does, and served bit-exact."""
    rng = np.random.default_rng(77)
    sizes = [0, 1, 15, 16, 17, 4095, 4096, 4097, 65_535, 65_536, 65_537, 200_003, 300_000] + [int(x) for x in rng.integers(1, 150_000, 300)]
    objs = []
    for i, n in enumerate(sizes):
        objs.append((wl.key128(b"packed/%d" % i), wl.synth_bytes(n, seed=8_000 + i).tobytes()))
    with nz.Zone(device=0, arena_bytes=256 << 20, max_objects=2048) as z:
        crcs = z.put_packed(objs)
        assert crcs == [o.crc32c(d) for _, d in objs]
        ext = z.lookup([k for k, _ in objs])
        assert all(e.status == capi.NZ_OK and e.stored_length == len(d) and e.crc32c == c for e, (_, d), c in zip(ext, objs, crcs))
        for j in (0, 1, 3, 7, 9, 10, 11, 12, 50, 311):
            res, body = z.get(objs[j][0], flags=capi.NZ_GET_VERIFY)
            assert res.status == capi.NZ_OK and body == objs[j][1], j
        res, body = z.get(objs[12][0], segs=[(123_457, 100_001)])
        assert res.status == capi.NZ_OK and body == o.serve_slice_mem(objs[12][1], 123_457, 100_001)
        # the second load of the same keys is refused as a whole (nothing half-stored), the zone still serves
        with pytest.raises(nz.NezhaError):
            z.put_packed(objs[:3])
        res, body = z.get(objs[2][0])
        assert res.status == capi.NZ_OK and body == objs[2][1]
