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


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _span_worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ok = False
    try:
        stripe = 4 << 20
        z = nz.Zone(device=rank, arena_bytes=512 << 20, max_objects=1024, shard_index=rank, shard_count=world)
        st = span.ShardedStore(z, dist, stripe_bytes=stripe)
        st.attach_peers()
        objs = []
        for i in range(6):
            key = wl.key128(b"span/two-gpu/%d" % i)
            length = [3 * stripe + 777, stripe, stripe + 1, 10 * stripe, 1000, 7 * stripe + 3][i]
            data = wl.synth_bytes(length, seed=900 + i)
            crc = st.put(key, length, lambda off, ln, d=data: d[off:off + ln])
            objs.append((key, length, data, crc))
        dist.barrier()
        for key, length, data, crc in objs:
            if shard.shard_of(key, world) != rank:
                assert z.lookup([key])[0].status == capi.NZ_ENOTFOUND       # GETs are routed to the owner
                continue
            assert crc == o.crc32c(data)
            res, body = z.get(key, device_dst=True)
            assert res.status == capi.NZ_OK and body == data.tobytes(), length
            for s, l in [(0, 1), (length - 1, 1), (length // 3, length // 2), (max(0, stripe - 5), min(10, length - max(0, stripe - 5)))]:
                if l <= 0:
                    continue
                res, body = z.get(key, segs=[(s, l)])
                assert res.status == capi.NZ_OK and body == o.serve_slice_mem(data, s, l), (length, s, l)
        dist.barrier()
        z.close()
        ok = True
    finally:
        q.put((rank, ok))
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs (gpurun --gpus 2)")
def test_span_gather_two_gpus_ipc():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_span_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(300)
    res = dict(q.get(timeout=10) for _ in range(2))
    assert res == {0: True, 1: True} and all(p.exitcode == 0 for p in procs)
