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


class _FakeBatch:
    """Stands in for PulsarBatch on the CPU: row r of a realization with global id g is g * 10 + column."""
    ld = 8
    device = torch.device("cpu")

    def generate(self, n, seed=0, real0=0, out=None, **kw):
        out[:n] = (torch.arange(real0, real0 + n, dtype=torch.float64)[:, None] * 10
                   + torch.arange(self.ld, dtype=torch.float64)[None, :]) + seed
        return out


def _worker_gathered(rank, world, port, nreal, chunk, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from pta_replicator_b200.distributed import gather_plan, generate_gathered
    C, n_chunks, padded = gather_plan(nreal, world, chunk)
    full = generate_gathered(_FakeBatch(), nreal, seed=3, chunk=chunk)
    expect = (torch.arange(padded, dtype=torch.float64)[:, None] * 10 + torch.arange(8, dtype=torch.float64)[None, :]) + 3
    ok = full.shape[0] == padded and padded >= nreal and C % 4 == 0 and torch.equal(full, expect)
    dist.barrier()
    dist.destroy_process_group()
    q.put((rank, ok, C, n_chunks))


@pytest.mark.parametrize("nreal,chunk", [(64, 8), (50, 12), (7, 512)])
def test_gloo_chunked_gather_places_rows_in_global_id_order(nreal, chunk):
    """generate_gathered: chunk-interleaved ownership; the all-gather of chunk c is a plain concat into rows
    [c G C, (c+1) G C) of the result, which therefore equals a single-process run over ids 0 .. padded-1."""
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_gathered, args=(r, world, port, nreal, chunk, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, ok, _, _ in res), res


def test_gather_plan_and_usable_cores():
    from pta_replicator_b200.distributed import chunk_ids, gather_plan, usable_cores
    C, n, padded = gather_plan(100000, 8, 512)
    assert C == 512 and n == 25 and padded == 25 * 8 * 512
    ids = sorted(chunk_ids(c, r, 8, C) for c in range(n) for r in range(8))
    assert ids == list(range(0, padded, C))                      # every block of C ids owned exactly once
    assert gather_plan(7, 2, 512) == (4, 1, 8)
    assert 1 <= usable_cores() <= (os.cpu_count() or 1)


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
