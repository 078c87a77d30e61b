"""Host logic of the multi-GPU path (read sharding + ordered gather), world_size 2 over gloo on CPU."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from recgraph_b200 import shard


def test_partition_is_contiguous_exact_and_balanced():
    import random
    rnd = random.Random(5)
    for n, world in [(1, 1), (7, 2), (100, 8), (1000, 4), (3, 8), (8, 8)]:
        costs = [rnd.randint(1, 10000) for _ in range(n)]
        parts = shard.partition(costs, world)
        assert len(parts) == world
        assert parts[0][0] == 0 and parts[-1][1] == n
        for (a, b), (c, d) in zip(parts, parts[1:]):
            assert b == c and a <= b and c <= d
        if n >= 50 * world:
            loads = [sum(costs[a:b]) for a, b in parts]
            assert max(loads) <= 1.5 * sum(costs) / world


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, lengths, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = shard.shard_for_rank(lengths, world, rank)
    local = [f"read{i}:{lengths[i]}" for i in range(lo, hi)]  # stands for the GAF lines of this rank's shard
    t = torch.tensor([float(hi - lo)])
    dist.all_reduce(t)  # every read is owned exactly once
    out = shard.gather_in_order(local, world, rank)
    if rank == 0:
        q.put((out, int(t.item())))
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks_gather_in_input_order():
    lengths = [1000 + (i * 37) % 500 for i in range(101)]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, lengths, q)) for r in range(2)]
    for p in procs:
        p.start()
    out, total = q.get(timeout=120)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert total == len(lengths)
    assert out == [f"read{i}:{lengths[i]}" for i in range(len(lengths))]
