"""Multi-GPU layer: realizations shard embarrassingly across ranks (one process per GPU,
``torch.distributed``), with an optional final all-gather of the residuals over NCCL.

There is no data-path collective: a realization depends only on (seed, global realization id)
(Philox counters, ``csrc/ptar_rng.cuh``), so rank r simply generates the ids of its block and the
result is bit-identical to a single-GPU run.  The all-gather is a delivery step: it moves 8 bytes
per TOA per realization across NVLink (~0.77 TB/s per direction per GPU measured) while the
generator produces them at several TB/s, so it is reported separately from the generation rate
(multi-GPU section).
"""
from __future__ import annotations

ALIGN = 4  # Philox counters carry 4 consecutive realizations


def shard_bounds(nreal: int, world: int, rank: int, align: int = ALIGN):
    """(start, count) of rank's block of realization ids; blocks are contiguous, ordered by rank,
    start at multiples of ``align`` and cover [0, nreal) exactly."""
    if world < 1 or not (0 <= rank < world) or nreal < 0:
        raise ValueError("bad shard request")
    nblk = (nreal + align - 1) // align
    base, extra = divmod(nblk, world)
    b0 = rank * base + min(rank, extra)
    b1 = b0 + base + (1 if rank < extra else 0)
    start = min(b0 * align, nreal)
    stop = min(b1 * align, nreal)
    return start, stop - start


def allgather_rows(local, counts, group=None):
    """Concatenate per-rank row blocks ``local`` [counts[rank], ld] into [sum(counts), ld] on every rank
    (``all_gather_into_tensor`` when the blocks are equal, padded ``all_gather`` otherwise)."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    ld = local.shape[1]
    if len(set(counts)) == 1:
        full = torch.empty((world * counts[0], ld), dtype=local.dtype, device=local.device)
        dist.all_gather_into_tensor(full, local.contiguous(), group=group)
        return full
    m = max(counts)
    pad = torch.zeros((m, ld), dtype=local.dtype, device=local.device)
    pad[:local.shape[0]] = local
    parts = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad, group=group)
    return torch.cat([p[:c] for p, c in zip(parts, counts)], dim=0)


def generate_sharded(batch, nreal: int, seed: int = 0, gather: bool = False, group=None, **kw):
    """Generate this rank's block of ``nreal`` realizations; with ``gather`` return all of them."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        world, rank = dist.get_world_size(group), dist.get_rank(group)
    else:
        world, rank = 1, 0
    start, count = shard_bounds(nreal, world, rank)
    out = batch.generate(count, seed=seed, real0=start, **kw) if count else None
    if not gather or world == 1:
        return out, (start, count)
    import torch
    if out is None:
        out = torch.empty((0, batch.ld), dtype=torch.float64, device=batch.device)
    counts = [shard_bounds(nreal, world, r)[1] for r in range(world)]
    return allgather_rows(out, counts, group), (0, nreal)
