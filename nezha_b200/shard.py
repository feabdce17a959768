"""nezha_b200/shard.py — key-hash sharding across the box's GPUs.

Objects are independent, so the store partitions by `u64le(key128[0:8]) mod nGPU`
(the same function the C ABI exports as nz_shard_of): each rank owns one shard,
the host routes every GET to its owner and there is NO collective on the common
path.  torch.distributed (NCCL on GPUs, gloo in the CPU tests) is only plumbing:
barriers, max-over-ranks timing, summing served bytes, exchanging IPC handles for
the one exchange step that exists (a GET that spans shards, nz_span.cu).
"""
from __future__ import annotations

from typing import Dict, List, Sequence, Tuple


def shard_of(key: bytes, world: int) -> int:
    """Owner shard of a key128; must agree with nz_shard_of in libnezha_gpu.so."""
    if world <= 1:
        return 0
    return int.from_bytes(key[:8], "little") % world


def partition(keys: Sequence[bytes], world: int) -> Dict[int, List[int]]:
    """request index lists per owner shard (every index appears exactly once)."""
    out: Dict[int, List[int]] = {r: [] for r in range(max(world, 1))}
    for i, k in enumerate(keys):
        out[shard_of(k, world)].append(i)
    return out


def stripe_plan(key: bytes, length: int, world: int, stripe_bytes: int = 64 << 20) -> List[Tuple[int, int, int]]:
    """C4 placement of one object: [(shard, offset, length)] per stripe.  Objects up to one
    stripe live whole on their owner; larger ones are striped round-robin starting at the owner
    so that a large GET spans shards."""
    owner = shard_of(key, world)
    if length <= stripe_bytes or world <= 1:
        return [(owner, 0, length)]
    out = []
    off, i = 0, 0
    while off < length:
        ln = min(stripe_bytes, length - off)
        out.append(((owner + i) % world, off, ln))
        off += ln
        i += 1
    return out


def reduce_timing(dist, elapsed_ms: float, payload_bytes: float, device=None) -> Tuple[float, float]:
    """(max over ranks of elapsed, sum over ranks of payload): whole-job throughput =
    all bytes served / slowest rank's time."""
    import torch

    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return elapsed_ms, payload_bytes
    t = torch.tensor([elapsed_ms], dtype=torch.float64, device=device)
    b = torch.tensor([payload_bytes], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dist.all_reduce(b, op=dist.ReduceOp.SUM)
    return float(t.item()), float(b.item())
