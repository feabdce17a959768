"""-m gpu: K2+K3 fusion.  With NZ_GET_VERIFY the gather checksums every tile from the copy in shared
memory and accumulates per 64 KiB chunk; only chunks a slice covers partly, and slices the bulk-copy
engine cannot take, are re-read by k_verify.  Both routes must agree with the chunk table (built at
fill time by K3, itself pinned to the oracle in test_gpu_parity) and catch a flipped bit exactly
where the oracle's CRC-32C would."""
import numpy as np
import pytest

import nezha_b200 as nz
from nezha_b200 import capi, workloads as wl
from oracle import oracle as o

pytestmark = pytest.mark.gpu

SIZES = [1, 15, 16, 17, 2047, 2048, 2049, 32_768, 65_535, 65_536, 65_537, 98_304, 131_072 + 5, 1_000_003, 5 * 1_048_576 + 11]


@pytest.fixture(scope="module")
def zone():
    z = nz.Zone(device=0, arena_bytes=1 << 30, max_objects=256)
    yield z
    z.close()


@pytest.fixture(scope="module")
def objs(zone):
    out = []
    for i, n in enumerate(SIZES):
        d = wl.synth_bytes(n, seed=7_000 + i).tobytes()
        k = wl.key128(b"vf/%d" % i)
        zone.put(k, d)
        out.append((k, d))
    return out


def test_whole_objects_verify_clean_every_size(zone, objs):
    """Aligned device destinations: every tile is an engine tile or a short tile of an engine slice."""
    gets, bufs = [], []
    for k, d in objs:
        b = zone.alloc_device(len(d) + 256)
        bufs.append(b)
        gets.append(nz.Get(k, b, [], capi.NZ_GET_VERIFY))
    s0 = zone.stats()
    t = zone.submit(gets)
    res = t.wait()
    t.free()
    s1 = zone.stats()
    chunks = sum((len(d) + 65_535) // 65_536 for _, d in objs)
    assert s1.verify_chunks - s0.verify_chunks == chunks
    assert s1.verify_fused_chunks - s0.verify_fused_chunks == chunks          # no second pass over the source at all
    for r, (k, d), b in zip(res, objs, bufs):
        assert r.status == capi.NZ_OK and r.bytes == len(d)
        assert b.bytes(0, len(d)) == d
        assert r.crc32c == o.crc32c(d)
        b.free()


@pytest.mark.parametrize("phase", [0, 16, 112, 3, 77])
def test_destination_phase_and_ranges(zone, objs, phase):
    """Slices starting and ending inside chunks (partly covered chunks are re-read, inner ones fused);
    odd destination phases push the whole slice onto the register path (no fusion at all)."""
    k, d = objs[-1]
    L = len(d)
    segs = [(0, 70_000, 0), (65_536 + 32, 300_000, 70_016), (16 * 9, 1_000_000, 70_016 + 300_032), (L - 200_011, 200_011, 70_016 + 300_032 + 1_000_016)]
    need = segs[-1][2] + segs[-1][1] + phase
    b = zone.alloc_device(need + 256)
    s0 = zone.stats()
    t = zone.submit([nz.Get(k, b, segs, capi.NZ_GET_VERIFY, phase)])
    (r,) = t.wait()
    t.free()
    s1 = zone.stats()
    assert r.status == capi.NZ_OK
    fused = s1.verify_fused_chunks - s0.verify_fused_chunks
    if phase % 16:
        assert fused == 0                                   # register path: every touched chunk is re-read
    else:
        # chunks wholly inside the 16 B aligned slices: [0,70000) -> 1, [65568,365568) -> 3, [144,1000144) -> 14,
        # [L-200011, L) -> 3 full ones + the object's 11-byte last chunk
        assert fused == 1 + 3 + 14 + 4
    for s, l, off in segs:
        assert b.bytes(phase + off, l) == d[s:s + l]
    b.free()


def test_coalesced_group_mixed_flags_and_corruption(zone):
    """One flipped bit: every member of the coalesced group that asked for verification reports
    NZ_ECORRUPT, members that did not ask are served as before, other chunks stay clean."""
    n = 3 * 1_048_576 + 123
    d = wl.synth_bytes(n, seed=7_777).tobytes()
    k = wl.key128(b"vf/corrupt")
    zone.put(k, d)
    bufs = [zone.alloc_device(n + 256) for _ in range(6)]
    flags = [capi.NZ_GET_VERIFY, 0, capi.NZ_GET_VERIFY, 0, capi.NZ_GET_VERIFY, 0]

    def wave(extra=()):
        gets = [nz.Get(k, b, [], f) for b, f in zip(bufs, flags)] + list(extra)
        t = zone.submit(gets)
        res = t.wait()
        t.free()
        return res

    assert [r.status for r in wave()] == [capi.NZ_OK] * 6
    pos = 2 * 1_048_576 + 70_001                                     # inside chunk 33
    assert zone.debug_poke(k, pos, 0x10) == capi.NZ_OK
    bad = bytearray(d)
    bad[pos] ^= 0x10
    assert o.crc32c(bytes(bad)) != o.crc32c(d)
    # slices on either side of the damaged chunk stay clean; one that contains it (fused) and one that touches it (re-read) do not
    c = pos // 65_536
    side = zone.alloc_device(4 << 20)
    extra = [nz.Get(k, side, [(0, c * 65_536, 0)], capi.NZ_GET_VERIFY),
             nz.Get(k, side, [((c + 1) * 65_536, n - (c + 1) * 65_536, 0)], capi.NZ_GET_VERIFY, 0),
             nz.Get(k, side, [((c - 1) * 65_536, 3 * 65_536, 0)], capi.NZ_GET_VERIFY),
             nz.Get(k, side, [(pos - 5, 10, 0)], capi.NZ_GET_VERIFY)]
    res = wave(extra)
    assert [r.status for r in res[:6]] == [capi.NZ_ECORRUPT, capi.NZ_OK, capi.NZ_ECORRUPT, capi.NZ_OK, capi.NZ_ECORRUPT, capi.NZ_OK]
    assert [r.status for r in res[6:]] == [capi.NZ_OK, capi.NZ_OK, capi.NZ_ECORRUPT, capi.NZ_ECORRUPT]
    assert bufs[1].bytes(0, n) == bytes(bad)                          # unverified members get the stored bytes as they are
    assert zone.debug_poke(k, pos, 0x10) == capi.NZ_OK                # flip back
    assert [r.status for r in wave()] == [capi.NZ_OK] * 6
    for b in bufs + [side]:
        b.free()


def test_host_destination_and_last_chunk_tail(zone, objs):
    """Pinned host windows (the HTTP front end's case) and objects whose last chunk is short and odd."""
    for k, d in objs:
        b = zone.alloc_host(len(d) + 64)
        t = zone.submit([nz.Get(k, b, [], capi.NZ_GET_VERIFY)])
        (r,) = t.wait()
        t.free()
        assert r.status == capi.NZ_OK and b.bytes(0, len(d)) == d
        b.free()


def test_large_wave_throughput_sanity(zone):
    """64 verified full GETs of one 256 MiB object: coalesced, fused; just has to finish clean."""
    n = 256 << 20
    d = wl.synth_bytes(n, seed=7_900)
    k = wl.key128(b"vf/big")
    zone.put(k, d)
    bufs = [zone.alloc_device(n) for _ in range(2)]
    t = zone.submit([nz.Get(k, b, [], capi.NZ_GET_VERIFY) for b in bufs])
    res = t.wait()
    t.free()
    assert all(r.status == capi.NZ_OK for r in res)
    assert np.array_equal(np.frombuffer(bufs[1].bytes(n - 4096, 4096), dtype=np.uint8), d[-4096:])
    for b in bufs:
        b.free()
