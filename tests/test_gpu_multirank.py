"""Rank-level bitwise test of the multi-GPU layer on real devices (needs >= 2 GPUs in the box; skips otherwise, the
gloo tests in test_distributed_cpu.py cover the same ownership / placement logic on the CPU): two NCCL ranks run
``generate_gathered`` (chunk-interleaved ownership, all-gather overlapped with generation), ``generate_gathered_p2p``
(the same delivery by CUDA-IPC peer pushes with the copy engines) and ``generate_sharded``;
every rank's result must equal a single-GPU ``generate`` of the same global realization ids bit for bit."""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _batch():
    import pta_replicator_b200 as P
    from pta_replicator_b200.engine import PulsarBatch
    from tests.fixtures import load_flags_case
    _, spec = load_flags_case()
    psrs = []
    for s in spec:
        p = P.pulsar_from_arrays(s["name"], s["loc"], s["mjd"].astype(np.longdouble), s["err_us"],
                                 flags=[{"f": f, "pta": "SYN"} for f in s["flag"]])
        P.make_ideal(p)
        psrs.append(p)
    b = PulsarBatch(psrs)
    for i, s in enumerate(spec):
        be = np.array(s["backends"])
        b.set_white(i, efac=s["efac"], log10_equad=s["l10_equad"], flagid="f", flags=be)
        b.set_ecorr(i, s["l10_ecorr"], flagid="f", flags=be, coarsegrain=1.0 / 86400.0)
        b.set_red(i, s["rn_l10A"], s["rn_gamma"], components=30)
    b.set_gwb(-14.2, 13.0 / 3.0)
    return b


def _worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from pta_replicator_b200 import distributed as D
    b = _batch()
    nreal, seed = 200, 77
    full = D.generate_gathered(b, nreal, seed=seed, chunk=24)
    torch.cuda.synchronize()
    C, n_chunks, padded = D.gather_plan(nreal, world, 24)
    ref = b.generate(padded, seed=seed, real0=0)                      # single-GPU run over the same ids
    ok_gather = bool(torch.equal(full[:padded], ref))
    part, (start, count) = D.generate_sharded(b, nreal, seed=seed)
    ok_shard = bool(torch.equal(part, ref[start:start + count]))
    allr, _ = D.generate_sharded(b, nreal, seed=seed, gather=True)
    ok_all = bool(torch.equal(allr, ref[:nreal]))
    # the same delivery by peer pushes over NVLink (CUDA IPC + copy engines) instead of the NCCL collective
    pushed = torch.zeros((padded, b.ld), dtype=torch.float64, device=b.device)
    D.generate_gathered_p2p(b, nreal, seed=seed, chunk=24, out=pushed)
    ok_p2p = bool(torch.equal(pushed, ref))
    dist.barrier()
    dist.destroy_process_group()
    q.put((rank, ok_gather, ok_shard, ok_all and ok_p2p))


def test_two_ranks_equal_a_single_gpu_run_bit_for_bit():
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip(f"needs 2 GPUs in one box, found {torch.cuda.device_count()} (bench.py checks the same property under "
                    "torchrun: config5.shard_bitwise_ok)")
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
    assert all(a and b and c for _, a, b, c in res), res
