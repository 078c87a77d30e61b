"""Read sharding across the GPUs of one box: reads are independent, the graph is replicated per device,
there is no collective on the data path — only the records / GAF text are gathered on rank 0 in input order.

`partition` gives every rank one CONTIGUOUS range of reads, balanced by an estimated cost per read
(cells ~ graph rows x read length for every mode of the reference; pass your own cost for banded runs).
"""
from typing import List, Sequence, Tuple


def partition(costs: Sequence[float], world: int) -> List[Tuple[int, int]]:
    """Contiguous, cost-balanced ranges [lo, hi) per rank; every read belongs to exactly one rank."""
    n = len(costs)
    total = float(sum(costs))
    out = []
    lo = 0
    acc = 0.0
    for r in range(world):
        if r == world - 1:
            hi = n
        else:
            target = total * (r + 1) / world
            hi = lo
            while hi < n and (acc + costs[hi] <= target or hi == lo and n - hi > world - 1 - r):
                acc += costs[hi]
                hi += 1
            hi = min(hi, n - (world - 1 - r)) if n >= world else hi
            hi = max(hi, lo)
            # keep acc consistent with hi
            acc = float(sum(costs[:hi]))
        out.append((lo, hi))
        lo = hi
    return out


def shard_for_rank(read_lengths: Sequence[int], world: int, rank: int) -> Tuple[int, int]:
    return partition([float(x) for x in read_lengths], world)[rank]


def gather_in_order(local_items: list, world: int, rank: int, group=None):
    """Gather per-rank lists on rank 0 and concatenate them in rank (= input) order. Uses torch.distributed when
    world > 1 (gloo on CPU tests, NCCL-backed object gather on the GPU box)."""
    if world == 1:
        return list(local_items)
    import torch.distributed as dist
    bucket = [None] * world if rank == 0 else None
    dist.gather_object(local_items, bucket, dst=0, group=group)
    if rank != 0:
        return None
    out = []
    for part in bucket:
        out.extend(part)
    return out
