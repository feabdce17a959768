"""-m gpu: randomised waves against the oracle (hypothesis).  Every example builds a wave of GETs with
random slices, destination phases, memory kinds and flags over a fixed population (identity + gzip
objects) and requires byte-for-byte agreement with the oracle's slices / gunzip, plus exact statuses."""
import numpy as np
import pytest
from hypothesis import HealthCheck, given, settings, strategies as st

import nezha_b200 as nz
from nezha_b200 import capi, workloads as wl
from oracle import oracle as o

pytestmark = pytest.mark.gpu

_STATE = {}


def population():
    if "z" not in _STATE:
        z = nz.Zone(device=0, arena_bytes=512 << 20, max_objects=256)
        objs = []
        for i, n in enumerate([1, 17, 4096, 65_535, 65_536, 200_003, 1_048_576, 3_000_001]):
            d = wl.synth_bytes(n, seed=4_000 + i).tobytes()
            k = wl.key128(b"prop/id/%d" % i)
            z.put(k, d)
            objs.append((k, d, d, 0))                        # key, stored, served-when-inflated, enc
        for i, n in enumerate([10, 50_000, 400_000]):
            raw = wl.text_like(n, seed=4_100 + i).tobytes()
            gz = wl.gzip_member(raw)
            k = wl.key128(b"prop/gz/%d" % i)
            z.put(k, np.frombuffer(gz, dtype=np.uint8), enc=capi.NZ_ENC_GZIP, raw_len=len(raw))
            objs.append((k, gz, raw, 1))
        _STATE["z"], _STATE["objs"] = z, objs
    return _STATE["z"], _STATE["objs"]


req = st.tuples(st.integers(0, 10), st.integers(0, 3), st.lists(st.tuples(st.floats(0, 1), st.floats(0, 1), st.integers(0, 200)), min_size=0, max_size=3),
                st.booleans(), st.integers(0, 130))


@settings(max_examples=60, deadline=None, suppress_health_check=[HealthCheck.too_slow, HealthCheck.data_too_large])
@given(st.lists(req, min_size=1, max_size=24))
def test_random_waves_match_oracle(reqs):
    z, objs = population()
    gets, expect, bufs = [], [], []
    for oi, flagsel, slices, host, base in reqs:
        key, stored, raw, enc = objs[oi]
        flags = [0, capi.NZ_GET_VERIFY, capi.NZ_GET_INFLATE, capi.NZ_GET_INFLATE | capi.NZ_GET_VERIFY][flagsel]
        served = raw if (enc and flags & capi.NZ_GET_INFLATE) else stored
        L = len(served)
        segs, off = [], 0
        for a, b, gap in slices:
            s = int(a * L)
            l = int(b * (L - s))
            if l == 0:
                continue
            off += gap
            segs.append((s, l, off))
            off += l
        need = (off if segs else L) + base
        buf = (z.alloc_host if host else z.alloc_device)(max(need, 1))
        bufs.append(buf)
        gets.append(nz.Get(key, buf, segs, flags, base))
        expect.append((served, segs, base))
    t = z.submit(gets)
    res = t.wait()
    t.free()
    for r, (served, segs, base), buf in zip(res, expect, bufs):
        assert r.status == capi.NZ_OK, r.status
        if segs:
            assert r.bytes == sum(l for _, l, _ in segs)
            for s, l, d in segs:
                assert buf.bytes(base + d, l) == served[s:s + l]
        else:
            assert r.bytes == len(served) and buf.bytes(base, len(served)) == served
        buf.free()


def test_population_oracle_agreement():
    z, objs = population()
    for key, stored, raw, enc in objs:
        e = z.lookup([key])[0]
        assert e.crc32c == o.crc32c(stored) and e.stored_length == len(stored)
        if enc:
            assert o.gunzip(stored) == raw and e.raw_length == len(raw)
