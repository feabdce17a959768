"""-m gpu: objects that span stripes / shards (K5).

* one GPU: stripes all local (shard_count 1) — exercises the host split at stripe
  boundaries, the stripe map and the plan kernel's per-stripe source resolution;
* two GPUs (skipped on a 1-GPU box; run with `gpurun --gpus 2`): one process per GPU,
  arenas mapped through CUDA IPC, a GET on the owner pulls the remote stripes with peer
  loads inside k_gather.  Bodies are compared with the oracle's slices, CRCs with the
  oracle's CRC-32C.
"""
import os
import socket

import numpy as np
import pytest
import torch

import nezha_b200 as nz
from nezha_b200 import capi, shard, span, workloads as wl
from oracle import oracle as o

pytestmark = pytest.mark.gpu


def test_striped_object_single_gpu():
    stripe = 96 * 1024          # not a multiple of the 32 KiB tile or of anything convenient
    with nz.Zone(device=0, arena_bytes=256 << 20, max_objects=512) as z:
        st = span.ShardedStore(z, None, stripe_bytes=stripe)
        length = 5 * stripe + 12_345
        d = wl.synth_bytes(length, seed=61)
        key = wl.key128(b"span/one-gpu")
        # world 1: place the stripes by hand on the only shard
        stripes, whole = [], 0
        for i, off in enumerate(range(0, length, stripe)):
            ln = min(stripe, length - off)
            sk = span.stripe_key(key, i)
            z.put(sk, d[off:off + ln], enc=capi.NZ_ENC_STRIPE)
            a_off, a_len, tab, crc = z.extent_location(sk)
            stripes.append((0, a_off, ln, tab))
            whole = crc if i == 0 else z.crc32c_combine(whole, crc, ln)
        z.put_striped(key, length, stripe, stripes, whole)
        assert whole == o.crc32c(d)
        e = z.lookup([key])[0]
        assert e.status == capi.NZ_OK and e.stored_length == length and e.flags & capi.NZ_ENC_STRIPED and e.crc32c == whole
        res, body = z.get(key)
        assert res.status == capi.NZ_OK and res.bytes == length and body == d.tobytes()
        rng = np.random.default_rng(2)
        for _ in range(40):
            s = int(rng.integers(0, length - 1)); l = int(rng.integers(1, min(length - s, 4 * stripe)))
            res, body = z.get(key, segs=[(s, l)], device_dst=bool(_ % 2))
            assert res.status == capi.NZ_OK and body == o.serve_slice_mem(d, s, l), (s, l)
        res, _ = z.get(key, segs=[(length - 1, 2)])
        assert res.status == capi.NZ_EINVAL
        # mixed wave: striped + plain objects together
        z.put(wl.key128(b"span/plain"), d[:70_000])
        b1, b2 = z.alloc_host(length), z.alloc_host(70_000)
        t = z.submit([nz.Get(key, b1), nz.Get(wl.key128(b"span/plain"), b2)])
        r = t.wait(); t.free()
        assert [x.status for x in r] == [0, 0] and b1.bytes() == d.tobytes() and b2.bytes() == d[:70_000].tobytes()
        assert z.evict(key) == capi.NZ_OK
        assert z.lookup([key])[0].status == capi.NZ_ENOTFOUND


def test_striped_verify_single_gpu():
    """NZ_GET_VERIFY on a striped object checks every piece against its stripe's own chunk table (fused into the
    gather for 16 B-aligned slices, k_verify otherwise); a flipped stored byte is NZ_ECORRUPT, never a silent NZ_OK."""
    stripe = 192 * 1024          # 3 chunks per stripe
    with nz.Zone(device=0, arena_bytes=256 << 20, max_objects=512) as z:
        length = 7 * stripe + 4_321
        d = wl.synth_bytes(length, seed=62)
        key = wl.key128(b"span/verify")
        stripes, whole = [], 0
        for i, off in enumerate(range(0, length, stripe)):
            ln = min(stripe, length - off)
            sk = span.stripe_key(key, i)
            z.put(sk, d[off:off + ln], enc=capi.NZ_ENC_STRIPE)
            a_off, a_len, tab, crc = z.extent_location(sk)
            stripes.append((0, a_off, ln, tab))
            whole = crc if i == 0 else z.crc32c_combine(whole, crc, ln)
        z.put_striped(key, length, stripe, stripes, whole)
        st0 = z.stats()
        cases = [[], [(0, stripe)], [(stripe - 16, 4 * stripe)], [(12_345, 3 * stripe + 17)], [(length - 100, 100)]]
        for segs in cases:
            for dev in (True, False):
                res, body = z.get(key, segs=segs, flags=capi.NZ_GET_VERIFY, device_dst=dev)
                want = b"".join(o.serve_slice_mem(d, s, l) for s, l in segs) if segs else d.tobytes()
                assert res.status == capi.NZ_OK and body == want, (segs, dev, res.status)
        st1 = z.stats()
        assert st1.verify_chunks - st0.verify_chunks >= 2 * (length // 65536)      # the two whole-object GETs alone
        assert st1.verify_fused_chunks > st0.verify_fused_chunks
        # flip one stored byte of stripe 4: every verified GET that touches it fails, the others still pass
        assert z.debug_poke(span.stripe_key(key, 4), 70_000, 0x10) == capi.NZ_OK
        res, _ = z.get(key, flags=capi.NZ_GET_VERIFY, device_dst=True)
        assert res.status == capi.NZ_ECORRUPT
        res, _ = z.get(key, segs=[(4 * stripe + 65_536, 10_000)], flags=capi.NZ_GET_VERIFY)
        assert res.status == capi.NZ_ECORRUPT
        res, body = z.get(key, segs=[(0, 4 * stripe)], flags=capi.NZ_GET_VERIFY, device_dst=True)
        assert res.status == capi.NZ_OK and body == d[:4 * stripe].tobytes()
        res, _ = z.get(key, segs=[(4 * stripe, 100)])          # unverified: served as stored
        assert res.status == capi.NZ_OK


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs (gpurun --gpus 2)")
def test_span_gather_two_gpus_ipc():
    """One process per GPU under torchrun (the launch path bench.py uses), NCCL for the handshakes."""
    import subprocess
    import sys
    port = _free_port()
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", str(port), os.path.join(os.path.dirname(os.path.abspath(__file__)), "span_worker.py")],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and r.stdout.count("span worker ok") == 2, (r.stdout[-2000:], r.stderr[-4000:])


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p
