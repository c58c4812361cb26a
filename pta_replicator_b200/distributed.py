"""Multi-GPU layer: realizations shard embarrassingly across ranks (one process per GPU,
``torch.distributed``), with the final all-gather of the residuals over NCCL (north star,
SURVEY.md 8e) chunked and overlapped with generation.

There is no data-path collective inside generation: a realization depends only on (seed, global
realization id) (Philox counters, ``csrc/ptar_rng.cuh``), so a rank simply generates the ids it owns
and the result is bit-identical to a single-GPU run whatever the split.

Two ownership patterns:

* ``shard_bounds`` / ``generate_sharded``: contiguous blocks, rank r owns ids
  ``[r R/G, (r+1) R/G)`` -- each rank keeps (or reduces, or writes) its own shard; one blocking
  all-gather at the end if asked for.
* ``gather_plan`` / ``generate_gathered``: chunk-interleaved blocks, chunk c of rank r holds ids
  ``c G C + r C .. + C``.  The all-gather of chunk c is then a plain concat into rows
  ``[c G C, (c+1) G C)`` of the full array (global-id order, no reshuffle), and it runs on a second
  stream while the generator already produces chunk c+1 into the other staging buffer.  The gather
  moves 8 (G-1)/G bytes per TOA per realization into every GPU over NVLink (~0.7-0.77 TB/s received
  per GPU measured) while HBM absorbs the generator's 8 bytes at several TB/s, so for large rows the
  gather is the ceiling (DESIGN.md section 8 gives the numbers).
"""
from __future__ import annotations

import os

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
    """Generate this rank's block of ``nreal`` realizations; with ``gather`` return all of them
    (one blocking all-gather after generation; ``generate_gathered`` overlaps it instead)."""
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


def gather_plan(nreal: int, world: int, chunk: int, align: int = ALIGN):
    """Chunk-interleaved ownership for ``generate_gathered``: returns ``(C, n_chunks, padded)`` with C the
    realizations per rank per chunk (multiple of ``align``), ``padded = n_chunks * world * C >= nreal``.
    Chunk c of rank r holds the global ids ``c*world*C + r*C + [0, C)``."""
    if world < 1 or nreal <= 0 or chunk <= 0:
        raise ValueError("bad gather plan")
    C = max(align, (min(chunk, -(-nreal // world)) + align - 1) // align * align)
    per_round = world * C
    n_chunks = -(-nreal // per_round)
    return C, n_chunks, n_chunks * per_round


def chunk_ids(c: int, rank: int, world: int, C: int):
    """First global realization id of chunk ``c`` of ``rank`` (it holds C consecutive ids)."""
    return c * world * C + rank * C


def generate_gathered(batch, nreal: int, seed: int = 0, chunk: int = 512, group=None, out=None, gather: bool = True,
                      comm_stream=None, **kw):
    """All ``nreal`` realizations on every rank, rows in global-id order: generate chunk c (stream A) while the
    NCCL all-gather of chunk c-1 runs (stream B); two staging buffers.  ``gather=False`` generates the same ids
    without the collective (every rank then holds only its own chunks' rows of ``out``; used to time the
    generator alone on the identical schedule).  Returns ``out`` [padded, ld]; rows >= nreal are valid extra
    realizations (ids nreal .. padded-1)."""
    import torch
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        world, rank = dist.get_world_size(group), dist.get_rank(group)
    else:
        world, rank = 1, 0
    C, n_chunks, padded = gather_plan(nreal, world, chunk)
    dev, ld = batch.device, batch.ld
    if out is None:
        out = torch.empty((padded, ld), dtype=torch.float64, device=dev)
    if out.shape[0] < padded or out.shape[1] != ld:
        raise ValueError(f"out must be at least [{padded}, {ld}]")
    if world == 1 or not gather:
        for c in range(n_chunks):
            r0 = chunk_ids(c, rank, world, C)
            batch.generate(C, seed=seed, real0=r0, out=out[r0:r0 + C], **kw)
        return out
    if torch.device(dev).type != "cuda":   # host-logic path (gloo tests): same ownership and placement, no streams
        stage = torch.zeros((C, ld), dtype=torch.float64)
        for c in range(n_chunks):
            batch.generate(C, seed=seed, real0=chunk_ids(c, rank, world, C), out=stage, **kw)
            dist.all_gather_into_tensor(out[c * world * C:(c + 1) * world * C], stage, group=group)
        return out
    main = torch.cuda.current_stream(dev)
    comm = comm_stream if comm_stream is not None else torch.cuda.Stream(dev)
    stage = getattr(batch, "_gather_stage", None)
    if stage is None or stage[0].shape != (C, ld):
        stage = (torch.zeros((C, ld), dtype=torch.float64, device=dev), torch.zeros((C, ld), dtype=torch.float64, device=dev))
        batch._gather_stage = stage
    gen_done = [torch.cuda.Event(), torch.cuda.Event()]
    gat_done = [torch.cuda.Event(), torch.cuda.Event()]
    for c in range(n_chunks):
        b = c & 1
        if c >= 2:
            main.wait_event(gat_done[b])          # staging buffer b has been sent
        batch.generate(C, seed=seed, real0=chunk_ids(c, rank, world, C), out=stage[b], **kw)
        gen_done[b].record(main)
        comm.wait_event(gen_done[b])
        with torch.cuda.stream(comm):
            dist.all_gather_into_tensor(out[c * world * C:(c + 1) * world * C], stage[b], group=group)
            gat_done[b].record(comm)
    main.wait_stream(comm)
    return out


class PeerDelivery:
    """All-gather without a collective library: every rank maps the result buffer of every other rank into its own
    address space (CUDA IPC through ``ptar_peer_export`` / ``ptar_peer_open``) and PUSHES the rows it generates
    straight into them with the copy engines over NVLink (``ptar_peer_copy`` = an asynchronous UVA device-to-device
    copy) on side streams, while the generator already produces the next chunk.  No SM is taken from the generator
    (NCCL's all-gather kernels run on SMs), no staging buffers: a chunk is generated into its rows of the local
    result and copied from there.  Rank r sends to r+1, r+2, ... so that at any moment every GPU receives from one
    peer.  Rows are complete on every rank after ``finish()`` (stream sync + barrier)."""

    def __init__(self, full, group=None, n_streams: int = 2):
        import ctypes as C
        import torch
        import torch.distributed as dist
        from . import _cabi
        self.torch, self.dist, self.group, self.lib = torch, dist, group, _cabi.lib()
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        self.full = full
        handle = (C.c_ubyte * 64)()
        off = C.c_int64(0)
        _cabi.check(self.lib.ptar_peer_export(full.data_ptr(), handle, C.byref(off)), "ptar_peer_export")
        mine = (bytes(handle), int(off.value), tuple(full.shape))
        everyone = [None] * self.world
        dist.all_gather_object(everyone, mine, group=group)
        self.bases, self.ptrs = {}, {}
        for r, (h, o, shape) in enumerate(everyone):
            if r == self.rank:
                continue
            if shape != tuple(full.shape):
                raise ValueError("PeerDelivery: result buffers differ in shape between ranks")
            base = C.c_void_p()
            buf = (C.c_ubyte * 64).from_buffer_copy(h)
            _cabi.check(self.lib.ptar_peer_open(buf, C.byref(base)), "ptar_peer_open")
            self.bases[r], self.ptrs[r] = base.value, base.value + o
        self.streams = [torch.cuda.Stream(full.device) for _ in range(max(1, n_streams))]
        self.row_bytes = full.shape[1] * full.element_size()

    def push_rows(self, row0: int, nrows: int, after_event):
        """Copy rows [row0, row0 + nrows) of the local result into every peer's result, once ``after_event`` (recorded
        on the generating stream) has completed."""
        src = self.full.data_ptr() + row0 * self.row_bytes
        from . import _cabi
        for k in range(1, self.world):
            r = (self.rank + k) % self.world
            st = self.streams[k % len(self.streams)]
            st.wait_event(after_event)
            _cabi.check(self.lib.ptar_peer_copy(self.ptrs[r] + row0 * self.row_bytes, src, nrows * self.row_bytes, st.cuda_stream),
                        "ptar_peer_copy")

    def finish(self):
        """All pushes of all ranks have landed when this returns."""
        for st in self.streams:
            st.synchronize()
        self.torch.cuda.synchronize(self.full.device)
        self.dist.barrier(group=self.group)

    def close(self):
        for r, base in list(self.bases.items()):
            self.lib.ptar_peer_close(base)
        self.bases.clear()
        self.ptrs.clear()


def generate_gathered_p2p(batch, nreal: int, seed: int = 0, chunk: int = 512, group=None, out=None, delivery=None, **kw):
    """``generate_gathered`` with the all-gather replaced by peer pushes (``PeerDelivery``): same ownership, same
    result (rows in global-id order on every rank).  Pass ``delivery`` (built once on ``out``) to reuse the IPC
    mappings across calls."""
    import torch
    import torch.distributed as dist
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    C, n_chunks, padded = gather_plan(nreal, world, chunk)
    if out is None:
        out = torch.empty((padded, batch.ld), dtype=torch.float64, device=batch.device)
    own = delivery is None
    if own:
        delivery = PeerDelivery(out, group)
    elif delivery.full.data_ptr() != out.data_ptr():
        raise ValueError("generate_gathered_p2p: `delivery` was built on a different buffer")
    main = torch.cuda.current_stream(batch.device)
    for c in range(n_chunks):
        r0 = chunk_ids(c, rank, world, C)
        batch.generate(C, seed=seed, real0=r0, out=out[r0:r0 + C], **kw)
        ev = torch.cuda.Event()
        ev.record(main)
        delivery.push_rows(r0, C, ev)
    delivery.finish()
    if own:
        delivery.close()
    return out


def bind_to_gpu_numa(device_index: int):
    """Pin this process to the CPU cores local to GPU ``device_index`` (``/sys/bus/pci/devices/<bdf>/local_cpulist``)
    so that pinned host buffers allocated afterwards live on the GPU's NUMA node -- with one process per GPU
    started without CPU binding, every rank's staging memory otherwise lands on the node torchrun runs on and
    half the GPUs copy across the socket interconnect.  Returns a description, or None if nothing was changed."""
    try:
        import torch
        pr = torch.cuda.get_device_properties(device_index)
        bdf = "%04x:%02x:%02x.0" % (int(pr.pci_domain_id), int(pr.pci_bus_id), int(pr.pci_device_id))
    except Exception:
        bdf = None
    if not bdf:
        try:
            import subprocess
            q = subprocess.run(["nvidia-smi", f"--id={device_index}", "--query-gpu=pci.bus_id", "--format=csv,noheader"],
                               capture_output=True, text=True, timeout=10)
            bdf = q.stdout.strip().splitlines()[0].strip()
        except Exception:
            return None
    bdf = bdf.lower()
    if bdf.count(":") == 2 and len(bdf.split(":")[0]) == 8:
        bdf = bdf[4:]                                  # nvidia-smi prints an 8-digit domain, sysfs uses 4
    path = f"/sys/bus/pci/devices/{bdf}"
    try:
        with open(os.path.join(path, "local_cpulist")) as fh:
            cpulist = fh.read().strip()
        with open(os.path.join(path, "numa_node")) as fh:
            node = int(fh.read().strip())
    except OSError:
        return None
    cpus = set()
    for part in cpulist.split(","):
        if not part:
            continue
        lo, _, hi = part.partition("-")
        cpus.update(range(int(lo), int(hi or lo) + 1))
    allowed = os.sched_getaffinity(0)
    cpus &= allowed
    if not cpus or cpus == allowed:
        return {"numa_node": node, "cpus": len(allowed), "changed": False}
    os.sched_setaffinity(0, cpus)
    return {"numa_node": node, "cpus": len(cpus), "changed": True}


def usable_cores() -> int:
    """Cores this process may actually use: the scheduler affinity mask capped by the cgroup CPU quota."""
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    for path in ("/sys/fs/cgroup/cpu.max", "/sys/fs/cgroup/cpu/cpu.cfs_quota_us"):
        try:
            with open(path) as fh:
                txt = fh.read().split()
            if path.endswith("cpu.max"):
                if txt[0] != "max":
                    n = min(n, max(1, int(int(txt[0]) / int(txt[1]))))
            else:
                q = int(txt[0])
                if q > 0:
                    with open("/sys/fs/cgroup/cpu/cpu.cfs_period_us") as fh:
                        n = min(n, max(1, q // int(fh.read().strip())))
            break
        except (OSError, ValueError, IndexError):
            continue
    return max(1, n)
