"""nezha_b200/span.py — objects that span shards (BASELINE config C4).

One process per GPU.  Objects up to one stripe live whole on the shard their key
hashes to; larger ones are striped round-robin (shard.stripe_plan).  Each rank maps
every other rank's arena through CUDA IPC once (`attach_peers`), after which a GET on
the owner gathers all stripes with direct NVLink peer loads inside the ordinary K2
gather kernel — the transfer IS the kernel, there is no separate collective.
torch.distributed only carries the handshakes (IPC handles, stripe locations).
"""
from __future__ import annotations

import hashlib
from typing import Callable, List, Optional

from . import capi, shard
from .zone import Zone

STRIPE_BYTES = 64 << 20


def stripe_key(key: bytes, index: int) -> bytes:
    return hashlib.md5(key + index.to_bytes(4, "little")).digest()


class ShardedStore:
    def __init__(self, zone: Zone, dist=None, stripe_bytes: int = STRIPE_BYTES):
        self.zone, self.dist, self.stripe_bytes = zone, dist, stripe_bytes
        self.rank = dist.get_rank() if dist is not None and dist.is_initialized() else 0
        self.world = dist.get_world_size() if dist is not None and dist.is_initialized() else 1
        assert zone.shard_index == self.rank and zone.shard_count == self.world
        self.attached = False

    def attach_peers(self):
        """Exchange IPC handles of the arenas and map every peer (collective)."""
        if self.world == 1 or self.attached:
            self.attached = True
            return
        handles: List[Optional[bytes]] = [None] * self.world
        self.dist.all_gather_object(handles, self.zone.ipc_export())
        for r, h in enumerate(handles):
            if r != self.rank:
                self.zone.peer_attach(r, h)
        self.attached = True
        self.dist.barrier()

    def put(self, key: bytes, length: int, gen: Callable[[int, int], object]) -> Optional[int]:
        """Collective put of one object; `gen(off, len)` yields that slice's bytes (numpy/bytes/
        CUDA tensor) so that each rank materialises only what it stores.  Returns the whole-object
        CRC-32C on the owner, None elsewhere."""
        plan = shard.stripe_plan(key, length, self.world, self.stripe_bytes)
        owner = plan[0][0]
        if len(plan) == 1:
            return self.zone.put(key, gen(0, length)) if owner == self.rank else None
        if not self.attached:
            self.attach_peers()
        local = []
        for idx, (sh, off, ln) in enumerate(plan):
            if sh == self.rank:
                sk = stripe_key(key, idx)
                self.zone.put(sk, gen(off, ln), enc=capi.NZ_ENC_STRIPE)
                a_off, a_len, tab, crc = self.zone.extent_location(sk)
                assert a_len == ln
                local.append((idx, a_off, ln, tab, crc))
        if self.world > 1:
            everyone: List[Optional[list]] = [None] * self.world
            self.dist.all_gather_object(everyone, local)
        else:
            everyone = [local]
        if owner != self.rank:
            return None
        table = {}
        for r, lst in enumerate(everyone):
            for idx, a_off, ln, tab, crc in lst:
                table[idx] = (r, a_off, ln, tab, crc)
        stripes, whole = [], 0
        for idx in range(len(plan)):
            r, a_off, ln, tab, crc = table[idx]
            assert r == plan[idx][0] and ln == plan[idx][2]
            stripes.append((r, a_off, ln, tab))
            whole = crc if idx == 0 else self.zone.crc32c_combine(whole, crc, ln)
        self.zone.put_striped(key, length, self.stripe_bytes, stripes, whole)
        return whole
