"""-m "not gpu": the N>1 host logic under torch.distributed/gloo, world_size 2 (and 3):
routing by key hash is a partition, agrees with the C ABI's nz_shard_of, striping covers
every byte once, and the whole-job metric is sum(bytes) / max(time)."""
import ctypes as C
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from nezha_b200 import capi, shard, workloads as wl


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        keys = [wl.key128(wl.object_key_string("gloo", i)) for i in range(500)]
        mine = shard.partition(keys, world)[rank]
        # agree with the C ABI
        L = capi.load()
        for i in mine[:50]:
            k = (C.c_uint8 * 16).from_buffer_copy(keys[i])
            assert L.nz_shard_of(k, world) == rank
        # every request has exactly one owner across ranks
        owned = torch.zeros(len(keys), dtype=torch.int32)
        owned[mine] = 1
        dist.all_reduce(owned)
        assert bool((owned == 1).all())
        # whole-job metric: bytes summed, time = slowest rank
        t, b = shard.reduce_timing(dist, 10.0 * (rank + 1), 1000.0 * (rank + 1))
        assert t == 10.0 * world and b == 1000.0 * world * (world + 1) / 2
        # stripe placement is identical on every rank (pure function of key and length)
        plan = shard.stripe_plan(keys[7], (300 << 20) + 5, world)
        flat = torch.tensor([x for st in plan for x in st], dtype=torch.int64)
        ref = flat.clone()
        dist.broadcast(ref, 0)
        assert torch.equal(flat, ref)
        q.put((rank, len(mine)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_partition_under_gloo(world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    got = dict(q.get(timeout=5) for _ in range(world))
    assert sum(got.values()) == 500 and all(v > 100 // world for v in got.values())


def test_stripe_plan_covers_object():
    key = wl.key128(b"stripe/obj")
    for world in (1, 2, 8):
        for length in (0, 1, 64 << 20, (64 << 20) + 1, (1 << 30), (1 << 30) + 12345):
            plan = shard.stripe_plan(key, length, world)
            assert sum(l for _, _, l in plan) == length
            assert [o for _, o, _ in plan] == [sum(l for _, _, l in plan[:i]) for i in range(len(plan))]
            assert plan[0][0] == shard.shard_of(key, world)
            if world > 1 and length > (64 << 20):
                assert len({s for s, _, _ in plan}) == min(world, len(plan))
