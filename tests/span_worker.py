"""tests/span_worker.py - one rank of the 2-GPU span test (launched by test_gpu_span.py under torchrun).

Every rank maps the other's arena through CUDA IPC; a GET on the owner pulls the remote stripes with peer loads
inside the gather kernel.  Bodies are compared with the oracle's slices, CRCs with the oracle's CRC-32C, and
NZ_GET_VERIFY is exercised against chunk tables that live on the other GPU.
"""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import nezha_b200 as nz                                     # noqa: E402
from nezha_b200 import capi, shard, span, workloads as wl   # noqa: E402
from oracle import oracle as o                              # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    stripe = 4 << 20
    z = nz.Zone(device=local, arena_bytes=512 << 20, max_objects=1024, shard_index=rank, shard_count=world)
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
        for flags in (0, capi.NZ_GET_VERIFY):
            res, body = z.get(key, device_dst=True, flags=flags)
            assert res.status == capi.NZ_OK and body == data.tobytes(), (length, flags, res.status)
            for s, l in [(0, 1), (length - 1, 1), (length // 3, length // 2), (max(0, stripe - 5), min(10, length - max(0, stripe - 5)))]:
                if l <= 0:
                    continue
                res, body = z.get(key, segs=[(s, l)], flags=flags)
                assert res.status == capi.NZ_OK and body == o.serve_slice_mem(data, s, l), (length, s, l, flags, res.status)
    dist.barrier()
    # corruption on the OTHER GPU: object 3 (10 stripes); its stripe 1 lives on the shard after the owner
    key, length, data, _ = objs[3]
    owner = shard.shard_of(key, world)
    holder = (owner + 1) % world
    if rank == holder:
        assert z.debug_poke(span.stripe_key(key, 1), 123_456, 0x04) == capi.NZ_OK
    torch.cuda.synchronize()
    dist.barrier()
    if rank == owner:
        res, _ = z.get(key, device_dst=True, flags=capi.NZ_GET_VERIFY)
        assert res.status == capi.NZ_ECORRUPT, res.status
        res, _ = z.get(key, segs=[(stripe + 100_000, 50_001)], flags=capi.NZ_GET_VERIFY)
        assert res.status == capi.NZ_ECORRUPT, res.status
        res, body = z.get(key, segs=[(2 * stripe, 3 * stripe)], flags=capi.NZ_GET_VERIFY, device_dst=True)
        assert res.status == capi.NZ_OK and body == data[2 * stripe:5 * stripe].tobytes()
    dist.barrier()
    z.close()
    print("span worker ok", rank, flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
