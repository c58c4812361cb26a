"""world_size-2 gloo test (CPU) of the multi-GPU layer: shard bounds + the final all-gather, with equal
and unequal shards.  The generator itself is GPU-only; here the shards are rank-coded tensors."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, nreal, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from pta_replicator_b200.distributed import allgather_rows, shard_bounds
    ld = 8
    start, count = shard_bounds(nreal, world, rank)
    local = (torch.arange(start, start + count, dtype=torch.float64)[:, None] * 10 + torch.arange(ld, dtype=torch.float64)[None, :])
    counts = [shard_bounds(nreal, world, r)[1] for r in range(world)]
    full = allgather_rows(local, counts)
    expect = (torch.arange(nreal, dtype=torch.float64)[:, None] * 10 + torch.arange(ld, dtype=torch.float64)[None, :])
    ok = torch.equal(full, expect)
    dist.barrier()
    dist.destroy_process_group()
    q.put((rank, ok, start, count))


@pytest.mark.parametrize("nreal", [16, 22])
def test_gloo_allgather_of_shards(nreal):
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, nreal, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, ok, _, _ in res), res
    assert sum(c for _, _, _, c in res) == nreal
