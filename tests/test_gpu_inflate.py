"""-m gpu: K4 gzip inflate through the C ABI vs the oracle (zlib inflate, wbits=31),
bit-exact (SPEC G2), declined exactly when the oracle declines (SPEC G4)."""
import gzip
import io
import struct
import zlib

import numpy as np
import pytest

import nezha_b200 as nz
from nezha_b200 import capi, workloads as wl
from oracle import oracle as o

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def zone():
    z = nz.Zone(device=0, arena_bytes=2 << 30, max_objects=1 << 12)
    yield z
    z.close()


_n = [0]


def put_gz(zone, member: bytes, raw_len: int):
    _n[0] += 1
    key = wl.key128(b"inflate-test/%d" % _n[0])
    zone.put(key, np.frombuffer(member, dtype=np.uint8) if member else b"", enc=capi.NZ_ENC_GZIP, raw_len=raw_len)
    return key


def member(raw: bytes, level=6, strategy=zlib.Z_DEFAULT_STRATEGY) -> bytes:
    c = zlib.compressobj(level, zlib.DEFLATED, 31, 8, strategy)
    return c.compress(raw) + c.flush()


def inputs():
    rng = np.random.default_rng(4)
    txt = wl.text_like(1 << 20, seed=31).tobytes()
    return {
        "empty": b"",
        "one": b"x",
        "short_fixed": b"hello hello hello hello",
        "text_1MiB": txt,
        "text_odd": txt[: 333_333],
        "zeros_1MiB": bytes(1 << 20),                       # dist=1 overlapping copies, max-length matches
        "period3": b"abc" * 100_000,
        "period_long": bytes(rng.integers(0, 256, 300, dtype=np.uint8)) * 2000,
        "random_64k": rng.integers(0, 256, 65_536, dtype=np.uint8).tobytes(),   # stored / near-incompressible
        "random_300k": rng.integers(0, 256, 300_001, dtype=np.uint8).tobytes(),
        "lowent": rng.integers(0, 4, 500_000, dtype=np.uint8).tobytes(),
        "far_matches": (rng.integers(0, 256, 32_768, dtype=np.uint8).tobytes()) * 8,   # distance 32768
    }


@pytest.fixture(scope="module")
def corpus():
    return inputs()


@pytest.mark.parametrize("level", [0, 1, 6, 9])
def test_inflate_levels_bit_exact(zone, corpus, level):
    for name, raw in corpus.items():
        gz = member(raw, level)
        assert o.gunzip(gz) == raw
        key = put_gz(zone, gz, len(raw))
        res, body = zone.get(key, flags=capi.NZ_GET_INFLATE, device_dst=True)
        assert res.status == capi.NZ_OK, (name, level, res.status)
        assert res.length == len(raw) and res.bytes == len(raw)
        assert body == raw, (name, level)
        zone.evict(key)


@pytest.mark.parametrize("strategy", [zlib.Z_FIXED, zlib.Z_HUFFMAN_ONLY, zlib.Z_RLE, zlib.Z_FILTERED])
def test_inflate_strategies(zone, corpus, strategy):
    for name in ("short_fixed", "text_odd", "lowent", "zeros_1MiB"):
        raw = corpus[name]
        gz = member(raw, 6, strategy)
        key = put_gz(zone, gz, len(raw))
        res, body = zone.get(key, flags=capi.NZ_GET_INFLATE, device_dst=True)
        assert res.status == capi.NZ_OK and body == raw == o.gunzip(gz), (name, strategy)
        zone.evict(key)


def test_inflate_to_host_and_ranges_G3(zone, corpus):
    raw = corpus["text_1MiB"]
    key = put_gz(zone, member(raw), len(raw))
    res, body = zone.get(key, flags=capi.NZ_GET_INFLATE)                  # pinned host destination
    assert res.status == capi.NZ_OK and body == raw
    for s, l in [(0, 1), (len(raw) - 1, 1), (12345, 67890), (1 << 19, 1 << 19), (7, 1_000_000)]:
        res, body = zone.get(key, segs=[(s, l)], flags=capi.NZ_GET_INFLATE)
        assert res.status == capi.NZ_OK and res.length == len(raw) and body == raw[s:s + l], (s, l)
    res, _ = zone.get(key, segs=[(len(raw), 1)], flags=capi.NZ_GET_INFLATE)
    assert res.status == capi.NZ_EINVAL
    # without the flag the stored member is served verbatim (reference behaviour, SPEC B3)
    res, body = zone.get(key)
    assert res.status == capi.NZ_OK and body == member(raw) and res.flags == capi.NZ_ENC_GZIP | capi.NZ_ENC_INFLATABLE
    # verify + inflate: every stored chunk is re-checksummed
    res, body = zone.get(key, flags=capi.NZ_GET_INFLATE | capi.NZ_GET_VERIFY)
    assert res.status == capi.NZ_OK and body == raw
    zone.evict(key)


def test_inflate_flag_on_identity_object_is_ignored(zone):
    d = wl.synth_bytes(70_000, seed=3)
    key = wl.key128(b"inflate-test/identity")
    zone.put(key, d)
    res, body = zone.get(key, flags=capi.NZ_GET_INFLATE)
    assert res.status == capi.NZ_OK and body == d.tobytes()
    zone.evict(key)


def test_gzip_header_fields(zone, corpus):
    raw = corpus["text_odd"]
    # FNAME via the gzip module
    bio = io.BytesIO()
    with gzip.GzipFile(filename="dataset.bin", mode="wb", fileobj=bio, mtime=1541604001) as f:
        f.write(raw)
    cases = {"fname": bio.getvalue()}
    # FEXTRA + FNAME + FCOMMENT + FHCRC built by hand around a raw deflate body
    body = zlib.compress(raw, 6)[2:-4]
    hdr = bytearray(b"\x1f\x8b\x08" + bytes([4 | 8 | 16 | 2]) + struct.pack("<I", 1541604001) + b"\x00\x03")
    hdr += struct.pack("<H", 6) + b"AB\x02\x00xy" + b"name.bin\x00" + b"a comment\x00"
    hdr += struct.pack("<H", zlib.crc32(bytes(hdr)) & 0xFFFF)
    trailer = struct.pack("<II", zlib.crc32(raw), len(raw) & 0xFFFFFFFF)
    cases["all_fields"] = bytes(hdr) + body + trailer
    bad = bytearray(cases["all_fields"]); bad[len(hdr) - 1] ^= 0xFF
    cases["bad_hcrc"] = bytes(bad)
    bad = bytearray(cases["fname"]); bad[3] |= 0x20
    cases["reserved_flag"] = bytes(bad)
    for name, gz in cases.items():
        want = o.gunzip(gz)
        key = put_gz(zone, gz, len(raw))
        res, got = zone.get(key, flags=capi.NZ_GET_INFLATE, device_dst=True)
        if want is None:
            assert res.status == capi.NZ_ECORRUPT, name
        else:
            assert res.status == capi.NZ_OK and got == want == raw, name
        zone.evict(key)
    assert o.gunzip(cases["bad_hcrc"]) is None and o.gunzip(cases["reserved_flag"]) is None


def test_declined_streams_G4(zone, corpus):
    raw = corpus["text_odd"]
    gz = member(raw)
    variants = {
        "trailing_byte": gz + b"\x00",
        "two_members": gz + gz,
        "truncated_1": gz[:-1],
        "truncated_half": gz[: len(gz) // 2],
        "bad_crc": gz[:-8] + struct.pack("<I", zlib.crc32(raw) ^ 1) + gz[-4:],
        "bad_isize": gz[:-4] + struct.pack("<I", len(raw) + 1),
        "bad_magic": b"\x1f\x8c" + gz[2:],
        "bad_method": gz[:2] + b"\x07" + gz[3:],
        "tiny": gz[:10],
    }
    for name, v in variants.items():
        assert o.gunzip(v) is None, name
        key = put_gz(zone, v, len(raw))
        res, _ = zone.get(key, flags=capi.NZ_GET_INFLATE, device_dst=True)
        assert res.status == capi.NZ_ECORRUPT and res.bytes == 0, (name, res.status)
        zone.evict(key)


def test_bit_flips_agree_with_oracle(zone, corpus):
    """Random single-bit corruption: the device accepts iff zlib accepts, and then the bytes agree."""
    raw = corpus["text_odd"][:60_000]
    gz = member(raw)
    rng = np.random.default_rng(12)
    agree_ok = agree_bad = 0
    for _ in range(80):
        pos = int(rng.integers(0, len(gz)))
        v = bytearray(gz)
        v[pos] ^= 1 << int(rng.integers(0, 8))
        v = bytes(v)
        want = o.gunzip(v)
        key = put_gz(zone, v, len(raw) if want is None else len(want))
        res, got = zone.get(key, flags=capi.NZ_GET_INFLATE, device_dst=True)
        if want is None or len(want) != len(raw):
            assert res.status == capi.NZ_ECORRUPT or (want is not None and res.status == capi.NZ_OK and got == want), pos
            agree_bad += 1
        else:
            assert res.status == capi.NZ_OK and got == want, pos
            agree_ok += 1
        zone.evict(key)
    assert agree_bad > 40          # most flips break the stream (CRC at the latest)


def test_wave_of_many_streams(zone):
    """C3-shaped: hundreds of gzip objects inflated in one wave, Zipf access."""
    n_obj = 300
    raws, keys = [], []
    for i in range(n_obj):
        raw = wl.text_like(20_000 + 997 * (i % 40), seed=500 + i).tobytes()
        raws.append(raw)
        keys.append(put_gz(zone, member(raw), len(raw)))
    ranks = wl.zipf_ranks(1000, n_obj, seed=3)
    sizes = [len(raws[r]) for r in ranks]
    offs = np.concatenate([[0], np.cumsum([(s + 255) // 256 * 256 for s in sizes])])
    buf = zone.alloc_device(int(offs[-1]))
    try:
        gets = [nz.Get(keys[r], buf, [], capi.NZ_GET_INFLATE, int(offs[j])) for j, r in enumerate(ranks)]
        t = zone.submit(gets)
        res = t.wait()
        t.free()
        assert all(r.status == capi.NZ_OK for r in res)
        for j in list(range(0, 1000, 37)) + [999]:
            assert buf.bytes(int(offs[j]), sizes[j]) == raws[ranks[j]], j
    finally:
        buf.free()
        for k in keys:
            zone.evict(k)


# ---------------------------------------------------------------------------
# checkpoint index (built at commit) + inflate dedupe
# ---------------------------------------------------------------------------
def test_commit_learns_the_inflated_length(zone, corpus):
    """nz_put_commit inflates the member once: raw_length is the member's, whatever nz_put_begin was told."""
    raw = corpus["text_odd"]
    for told in (0, len(raw), 12345):
        key = put_gz(zone, member(raw), told)
        ext = zone.lookup([key])[0]
        assert ext.status == capi.NZ_OK and ext.raw_length == len(raw) and ext.flags == capi.NZ_ENC_GZIP | capi.NZ_ENC_INFLATABLE, told
        res, body = zone.get(key, flags=capi.NZ_GET_INFLATE, device_dst=True)
        assert res.status == capi.NZ_OK and body == raw
        zone.evict(key)


@pytest.mark.parametrize("level", [1, 6, 9])
def test_multi_segment_members(zone, level):
    """Members of several index segments: every segment decodes on its own warp; matches that reach back across a
    segment boundary take their bytes from the index."""
    rng = np.random.default_rng(77)
    block = rng.integers(0, 256, 30_000, dtype=np.uint8).tobytes()
    cases = {
        "text_5MiB": wl.text_like(5 << 20, seed=91).tobytes(),
        "far_repeat": block * 40,                                          # distance ~30000 matches across every boundary
        "mixed": wl.text_like(700_000, seed=92).tobytes() + bytes(300_000) + rng.integers(0, 256, 200_000, dtype=np.uint8).tobytes()
                 + wl.text_like(900_000, seed=93).tobytes(),                # huffman blocks, long runs, stored blocks
    }
    for name, raw in cases.items():
        gz = member(raw, level)
        key = put_gz(zone, gz, 0)
        st0 = zone.stats()
        res, body = zone.get(key, flags=capi.NZ_GET_INFLATE, device_dst=True)
        assert res.status == capi.NZ_OK and res.length == len(raw), (name, level, res.status)
        assert body == raw == o.gunzip(gz), (name, level)
        st1 = zone.stats()
        # the member went through the index (a segment is >= 128 KiB of output and ends with a group), none decoded whole
        assert st1.inflate_segments - st0.inflate_segments >= 2 and st1.inflate_whole_members == st0.inflate_whole_members
        assert st1.index_bytes > 0
        for s, l in [(131_071, 3), (len(raw) - 5, 5), (262_144 - 100, 300_000)]:
            res, body = zone.get(key, segs=[(s, l)], flags=capi.NZ_GET_INFLATE)
            assert res.status == capi.NZ_OK and body == raw[s:s + l], (name, s, l)
        zone.evict(key)


def test_fragmented_member_without_index(zone, corpus):
    """A member cut into thousands of tiny deflate blocks outgrows the index caps: it is served by the whole-member
    decode instead, bit-exact all the same."""
    raw = corpus["text_odd"][:120_000]
    c = zlib.compressobj(6, zlib.DEFLATED, 31)
    parts = []
    for i in range(0, len(raw), 40):
        parts.append(c.compress(raw[i:i + 40]))
        parts.append(c.flush(zlib.Z_FULL_FLUSH))
    gz = b"".join(parts) + c.flush()
    assert o.gunzip(gz) == raw
    key = put_gz(zone, gz, 0)
    assert zone.lookup([key])[0].raw_length == len(raw)
    st0 = zone.stats()
    res, body = zone.get(key, flags=capi.NZ_GET_INFLATE, device_dst=True)
    assert res.status == capi.NZ_OK and body == raw
    st1 = zone.stats()
    assert st1.inflate_whole_members == st0.inflate_whole_members + 1 and st1.inflate_segments == st0.inflate_segments
    zone.evict(key)


def test_dedupe_one_decode_many_requesters(zone, corpus):
    """A wave asking for the same member many times inflates it once; every requester gets its own bytes, and a
    requester with a bad slice fails alone."""
    raw = corpus["text_1MiB"]
    other = corpus["text_odd"]
    k1 = put_gz(zone, member(raw), len(raw))
    k2 = put_gz(zone, member(other), len(other))
    dev = zone.alloc_device(8 << 20)
    host = zone.alloc_host(4 << 20)
    try:
        gets = [
            nz.Get(k1, dev, [(len(raw), 1, 0)], capi.NZ_GET_INFLATE, 0),                      # 0: leader, slice outside -> EINVAL
            nz.Get(k1, dev, [], capi.NZ_GET_INFLATE, 1 << 20),                                 # 1: whole -> HBM
            nz.Get(k1, host, [], capi.NZ_GET_INFLATE, 0),                                      # 2: whole -> pinned host
            nz.Get(k1, dev, [(5, 70_000, 0), (900_000, 100_000, 70_000)], capi.NZ_GET_INFLATE, 3 << 20),   # 3: two slices
            nz.Get(k2, dev, [], capi.NZ_GET_INFLATE, 4 << 20),                                 # 4: sole requester of k2: direct
            nz.Get(k1, dev, [], capi.NZ_GET_INFLATE | capi.NZ_GET_VERIFY, 5 << 20),            # 5: whole, verified
            nz.Get(k1, dev, [], 0, 7 << 20),                                                   # 6: the stored member, no inflate
        ]
        st0 = zone.stats()
        t = zone.submit(gets)
        res = t.wait()
        t.free()
        st1 = zone.stats()
        # k1 (1 MiB = 8 segments) decoded once for its five inflate requesters, k2 (333 333 B = 3 segments) once
        assert 8 + 2 <= st1.inflate_segments - st0.inflate_segments <= 9 + 3
        assert res[0].status == capi.NZ_EINVAL
        assert [r.status for r in res[1:]] == [capi.NZ_OK] * 6
        assert dev.bytes(1 << 20, len(raw)) == raw
        assert host.bytes(0, len(raw)) == raw
        assert dev.bytes(3 << 20, 170_000) == raw[5:70_005] + raw[900_000:1_000_000]
        assert dev.bytes(4 << 20, len(other)) == other
        assert dev.bytes(5 << 20, len(raw)) == raw
        assert dev.bytes(7 << 20, res[6].bytes) == member(raw)
    finally:
        dev.free()
        host.free()
        zone.evict(k1)
        zone.evict(k2)


def test_put_many_bulk_commit(zone):
    """nz_put_commit_many: one batched commit of gzip and identity objects."""
    raws = [wl.text_like(50_000 + 4001 * i, seed=700 + i).tobytes() for i in range(40)]
    keys = [wl.key128(b"inflate-test/many/%d" % i) for i in range(40)]
    crcs = zone.put_many([(k, member(r), 0) for k, r in zip(keys, raws)], enc=capi.NZ_ENC_GZIP)
    assert crcs == [o.crc32c(member(r)) for r in raws]
    ext = zone.lookup(keys)
    assert [e.raw_length for e in ext] == [len(r) for r in raws]
    buf = zone.alloc_device(40 * (256 << 10))
    try:
        t = zone.submit([nz.Get(k, buf, [], capi.NZ_GET_INFLATE, i * (256 << 10)) for i, k in enumerate(keys)])
        res = t.wait()
        t.free()
        assert all(r.status == capi.NZ_OK for r in res)
        for i in range(40):
            assert buf.bytes(i * (256 << 10), len(raws[i])) == raws[i], i
    finally:
        buf.free()
        for k in keys:
            zone.evict(k)


def test_stored_byte_corrupted_after_commit(zone, corpus):
    """A member whose bytes rot in HBM after its index was built: verify catches it; without verify the decode either
    notices (recorded lengths no longer reproduce) or returns bytes, but never crashes the wave."""
    raw = corpus["text_1MiB"]
    gz = member(raw)
    key = put_gz(zone, gz, len(raw))
    zone.debug_poke(key, len(gz) // 2, 0x10)
    res, _ = zone.get(key, flags=capi.NZ_GET_INFLATE | capi.NZ_GET_VERIFY, device_dst=True)
    assert res.status == capi.NZ_ECORRUPT
    res, _ = zone.get(key, flags=capi.NZ_GET_INFLATE, device_dst=True)
    assert res.status in (capi.NZ_OK, capi.NZ_ECORRUPT)
    zone.evict(key)


def test_inflate_flag_without_a_gzip_object_in_a_recycled_ticket(zone, corpus):
    """A wave that carries NZ_GET_INFLATE but names only identity objects launches the segment decoder with nothing to do.
    Its work total must not be whatever an earlier inflate wave left in the (recycled) ticket: found by the property
    test as cudaErrorMisalignedAddress, round 2."""
    raw = np.frombuffer(corpus["text_odd"], dtype=np.uint8)
    kz = wl.key128(b"gz/recycle/z")
    ki = wl.key128(b"gz/recycle/i")
    zone.put(kz, np.frombuffer(wl.gzip_member(raw), dtype=np.uint8), enc=capi.NZ_ENC_GZIP, raw_len=len(raw))
    zone.put(ki, raw[:70_001])
    for _ in range(3):
        # same shape (6 requests) so the second wave takes the first one's ticket back from the pool
        bufs = [zone.alloc_device(len(raw) + 64) for _ in range(6)]
        t = zone.submit([nz.Get(kz, b, [], capi.NZ_GET_INFLATE, 0) for b in bufs])
        res = t.wait(); t.free()
        assert all(r.status == capi.NZ_OK for r in res) and bufs[0].bytes(0, len(raw)) == raw.tobytes()
        t = zone.submit([nz.Get(ki, b, [], capi.NZ_GET_INFLATE | (capi.NZ_GET_VERIFY if j & 1 else 0), j) for j, b in enumerate(bufs)])
        res = t.wait(); t.free()
        assert all(r.status == capi.NZ_OK for r in res)
        assert bufs[3].bytes(3, 70_001) == raw[:70_001].tobytes()
        for b in bufs:
            b.free()
