#!/usr/bin/env python
"""Benchmark of the hot path: realizations/s of 67-pulsar ng15-shaped EFAC/EQUAD + ECORR + red noise
+ HD-correlated GWB residuals (BASELINE.json metric), on N GPUs of one node.

    python bench.py --gpus N --steps K --warmup W            # our arm (torchrun for N > 1)
    python bench.py --impl reference --gpus N --steps K ...   # CPU arm: the oracle port on host cores

A step = one batch of R realizations (default 1000) of the whole 67-pulsar array through
``PulsarBatch.generate`` (Philox mode; outputs stay in HBM).  ``value`` = realizations of all ranks /
max-over-ranks device time.  ``e2e`` = the same metric through ``ptar_run_job_to_host``: per-step noise
parameters are copied host->device from pinned memory and every residual is copied back to pinned host
memory inside the timed region.  ``roofline`` is for the dominant kernel (the fused generator):
algorithmic bytes = 8 B x sum(N_toa) x realizations per launch, duration from CUDA events
around every launch in the timed region, peak = MEASURED_PEAKS.json ``hbm_gbs``.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "realizations/sec (67-psr ng15 GWB+RN+ECORR)"
FALLBACK_HBM_GBS = 6650.0


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        with open(p) as fh:
            d = json.load(fh)
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        super().__init__(daemon=True)
        self.index, self.samples, self.stop_flag, self.proc = index, [], False, None

    def run(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                if self.stop_flag:
                    break
                self.samples.append([x.strip() for x in line.split(",")])
        except Exception:
            pass

    def finish(self):
        self.stop_flag = True
        if self.proc is not None:
            self.proc.terminate()
        sm = sorted(float(s[0]) for s in self.samples if s and s[0].replace(".", "").isdigit())
        mx = [float(s[1]) for s in self.samples if len(s) > 1 and s[1].replace(".", "").isdigit()]
        reasons = set()
        for s in self.samples:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), s[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(self.samples)}


def cpu_arm(ds, n_real, cores, budget_s=None):
    """realizations/s of the oracle port on `cores` processes; returns (rate, wall, completed)."""
    from oracle import recipe
    done, wall = recipe.timed_realizations(ds, n_real, cores, budget_s)
    return done / wall, wall, done


def run_reference(args):
    """The reference arm: the oracle port (the reference is Python + PINT and cannot travel to the box;
    oracle/recipe.py) on all host cores; each step = one realization per core."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import recipe
    from pta_replicator_b200 import synthetic
    psrs, noise = synthetic.make_ng15_like(args.kind)
    ds = recipe.dataset_from_pulsars(psrs, noise)
    cores = min(os.cpu_count() or 1, 64)
    per_step = cores
    done_total, wall = 0, 0.0
    for _ in range(args.steps):   # a step = one realization per core; pool start-up and one warm-up pass are untimed
        _, w, done = cpu_arm(ds, per_step, cores, budget_s=max(20.0, 180.0 / max(args.steps, 1)))
        done_total += done
        wall += w
    value = done_total / wall
    ntoa = sum(p.toas.ntoas for p in psrs)
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "realizations/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * wall / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"ng15-{args.kind} 67 psr, sum N_toa={ntoa}, EFAC/EQUAD+ECORR(1s)+RN(30)+HD GWB", "realizations_per_step": per_step},
            "cpu_baseline": {"value": value, "unit": "realizations/s", "cores": cores, "kind": "port",
                             "sample": f"{per_step} realizations/step x {args.steps} steps, one process per core, numpy oracle port "
                                       "(no PINT, no dense U: faster than the unmodified reference)"},
            "e2e": {"value": value, "unit": "realizations/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--nreal", type=int, default=1000, help="realizations per step per GPU")
    ap.add_argument("--kind", default="full", choices=["full", "epoch"])
    ap.add_argument("--rc", type=int, default=0)
    ap.add_argument("--chunk", type=int, default=0)
    ap.add_argument("--merged-white", action="store_true", help="one N(0,w1^2+w2^2) draw per TOA instead of two")
    ap.add_argument("--e2e-nreal", type=int, default=256)
    ap.add_argument("--cpu-seconds", type=float, default=20.0)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-variants", action="store_true")
    ap.add_argument("--split", action="store_true", help="epoch kernel + TOA kernel (two launches) instead of the fused generator")
    ap.add_argument("--taylor-tol", type=float, default=1e-14,
                    help="PulsarBatch(rn_taylor_tol=...): remainder bound of the in-epoch Taylor step, relative to the red-noise rms")
    ap.add_argument("--gather", action="store_true", help="also time an NCCL all-gather of one chunk of residuals")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import numpy as np
    import torch
    import __graft_entry__ as ge
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    if rank == 0:
        ge.build()               # no-op when the in-tree .so is up to date
    if dist is not None:
        dist.barrier()           # the other ranks load the library only after rank 0 has (re)built it
    from pta_replicator_b200 import synthetic
    from pta_replicator_b200.engine import PulsarBatch

    psrs, noise = synthetic.make_ng15_like(args.kind)
    b = PulsarBatch(psrs, rn_taylor_tol=args.taylor_tol)
    b.white_merged = bool(args.merged_white)
    synthetic.ng15_recipe(b, noise)
    if args.chunk:
        b.default_chunk = args.chunk
    b.split_epoch = bool(args.split)
    st = b.compile()
    R = args.nreal
    out = torch.empty((R, b.ld), dtype=torch.float64, device=b.device)
    out.zero_()
    seed = 20250922

    def step(k, timers=None):
        # global realization ids: rank-major blocks so any shard is reproducible on any GPU
        real0 = ((k * world + rank) * R + 3) // 4 * 4
        b.generate(R, seed=seed, real0=real0, out=out, rc=args.rc, timers=timers)

    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()          # samples span warm-up, the timed region and a short soak of the same steps
        t_wait = time.time()
        while not sampler.samples and time.time() - t_wait < 8.0:
            time.sleep(0.05)     # nvidia-smi can take a second to deliver its first sample
    for k in range(args.warmup):
        step(k)
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for k in range(args.steps):
        step(args.warmup + k)
    e1.record()
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    ms = e0.elapsed_time(e1)
    # per-kernel durations: the same K steps again, serialized on one stream with CUDA events around every launch
    timers = {}
    for k in range(args.steps):
        step(args.warmup + args.steps + k, timers)
    torch.cuda.synchronize()
    if sampler:
        # the timed region lasts ~10-20 ms, shorter than nvidia-smi's 100 ms period: keep the identical load
        # running for ~0.7 s so that the clock / throttle record is taken under this load
        n0, t_soak, k = len(sampler.samples), time.time(), 0
        while time.time() - t_soak < 0.7:
            step(args.warmup + 2 * args.steps + k)
            torch.cuda.synchronize()
            k += 1
        clocks = sampler.finish()
        clocks["window"] = "warm-up + timed region + %d soak steps of the same load (%d samples before the soak)" % (k, n0)
    else:
        clocks = None
    tms = torch.tensor([ms], dtype=torch.float64, device=b.device)
    if dist is not None:
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
    ms_max = float(tms.item())
    value = world * R * args.steps / (ms_max * 1e-3)

    # per-kernel durations inside the timed region (same stream)
    per = {}
    for name, a, z in timers["events"]:
        per.setdefault(name, []).append(a.elapsed_time(z))
    kern = {k: {"launches": len(v), "avg_ms": float(np.mean(v)), "total_ms": float(np.sum(v))} for k, v in per.items()}
    chunk_real = min(R, b.default_chunk)
    hbm, how = peaks()
    gen = kern["generate"]
    n_gen = gen["launches"]
    alg_bytes_per_launch = 8.0 * b.n_toa_total * (R * args.steps / n_gen)
    achieved = alg_bytes_per_launch / (gen["avg_ms"] * 1e-3) / 1e9
    roof = {"bound": "hbm", "kernel": "gen_kernel (fused white+ECORR+red+GWB-interp generator)", "achieved": achieved,
            "peak": hbm, "unit": "GB/s", "frac": achieved / hbm, "traffic": None, "peak_source": how,
            "algorithmic_bytes_per_launch": alg_bytes_per_launch, "avg_launch_ms": gen["avg_ms"],
            "share_of_step": gen["total_ms"] / sum(k["total_ms"] for k in kern.values())}
    prof = os.path.join(ROOT, "profiles", "r01_traffic.json")
    if os.path.isfile(prof):
        with open(prof) as fh:
            tj = json.load(fh)
        cap_real = float(tj.get("gen_kernel_realizations_per_launch", 512))
        cap_bytes = float(tj.get("gen_kernel_dram_bytes_per_launch", 0.0))
        roof["traffic"] = cap_bytes / cap_real * (R * args.steps / n_gen)
        roof["traffic_note"] = ("dram__bytes_read.sum + dram__bytes_write.sum of one ncu --set full capture of a %d-realization "
                                "launch (%.4g B; algorithmic %.4g B)%s" % (cap_real, cap_bytes, 8.0 * b.n_toa_total * cap_real,
                                "" if cap_real == R * args.steps / n_gen else ", scaled to this run's realizations per launch"))

    line = {"metric": METRIC, "value": value, "unit": "realizations/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"ng15-{args.kind}: 67 psr, sum N_toa={b.n_toa_total}, EFAC/EQUAD+ECORR(1s)+RN(30 comp)+HD GWB(npts=600,howml=10)",
                       "realizations_per_step_per_gpu": R, "rng": "in-kernel Philox4x32-10, fp32 Box-Muller",
                       "white_draws_per_toa": 1 if args.merged_white else 2, "gwb_chunk": chunk_real,
                       "rn_taylor_tol": args.taylor_tol,
                       "l2": f"output per step {R * b.ld * 8 / 1e9:.2f} GB > L2 (126 MB); no flush needed",
                       "parallelism": f"realization-sharded x{world}, no data-path collective"},
            "kernels": kern, "kernels_timing": "second pass of the same K steps with CUDA events around every launch (same stream, same schedule)",
            "roofline": roof, "clocks": clocks,
            "gpu_launches": int(sum(k["launches"] for k in kern.values()))}

    # ---- variant: one merged white draw per TOA (identical distribution; PTAR_F_WHITE1), rank 0, N == 1 only
    if world == 1 and not args.merged_white and not args.no_variants:
        b1 = PulsarBatch(psrs, rn_taylor_tol=args.taylor_tol)
        b1.white_merged = True
        b1.split_epoch = bool(args.split)
        synthetic.ng15_recipe(b1, noise)
        if args.chunk:
            b1.default_chunk = args.chunk
        for k in range(2):
            b1.generate(R, seed=seed, real0=0, out=out, rc=args.rc)
        tv = {}
        v0, v1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        v0.record()
        for k in range(3):
            b1.generate(R, seed=seed, real0=4 * R * (k + 1), out=out, rc=args.rc, timers=tv)
        v1.record()
        torch.cuda.synchronize()
        gv = [a.elapsed_time(z) for n, a, z in tv["events"] if n == "generate"]
        line["variants"] = {"merged_white_draw": {
            "value": 3 * R / (v0.elapsed_time(v1) * 1e-3), "unit": "realizations/s",
            "roofline_frac": 8.0 * b.n_toa_total * (3 * R / len(gv)) / (float(np.mean(gv)) * 1e-3) / 1e9 / hbm,
            "note": "w1 z1 + w2 z2 replaced by sqrt(w1^2+w2^2) z: same Gaussian law, one Philox draw per TOA; not the headline"}}
        del b1

    # ---- e2e through the C ABI with host buffers (rank-local; aggregated like `value`)
    Re = min(args.e2e_nreal, R)
    pinned_out = torch.empty((Re, b.ld), dtype=torch.float64, pin_memory=True)
    host_in = {k: st[k].cpu().pin_memory() for k in ("w1", "w2", "ep_ecorr", "rn_scale")}
    h2d = int(sum(v.numel() * v.element_size() for v in host_in.values()))

    def e2e_step(k):
        for name, h in host_in.items():
            st[name].copy_(h, non_blocking=True)        # this step's noise parameters, pinned host -> device
        b.generate_to_host(Re, seed=seed + 1, real0=4 * k * Re, out_host=pinned_out, chunk=32, rc=args.rc)

    for k in range(2):
        e2e_step(k)
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    ne2e = max(2, min(args.steps, 5))
    t0 = time.perf_counter()
    for k in range(ne2e):
        e2e_step(2 + k)
    torch.cuda.synchronize()
    te = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=b.device)
    if dist is not None:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    line["e2e"] = {"value": world * Re * ne2e / float(te.item()), "unit": "realizations/s", "h2d_bytes_per_step": h2d,
                   "d2h_bytes_per_step": int(Re * b.ld * 8), "realizations_per_step_per_gpu": Re, "steps": ne2e,
                   "path": "ptar_run_job_to_host: pinned H2D of noise parameters, generate in chunks of 32, D2H of every residual overlapped on a second stream"}

    if dist is not None:   # final NCCL all-gather of residuals (north star): timed separately on a bounded block
        n = min(R, 64)
        full = torch.empty((world * n, b.ld), dtype=torch.float64, device=b.device)
        dist.all_gather_into_tensor(full, out[:n])
        torch.cuda.synchronize()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record()
        dist.all_gather_into_tensor(full, out[:n])
        g1.record()
        torch.cuda.synchronize()
        gms = g0.elapsed_time(g1)
        gather_rate = world * n / (gms * 1e-3)           # realizations/s the gather alone can deliver to every rank
        line["allgather"] = {"realizations_per_rank": n, "ms": gms, "recv_GBps_per_gpu": (world - 1) * n * b.ld * 8 / gms / 1e6,
                             "value_if_every_realization_were_gathered": 1.0 / (1.0 / value + 1.0 / gather_rate),
                             "note": "final NCCL all_gather_into_tensor of residuals, not in `value`: NVLink moves 8 B/TOA/realization "
                                     "~6x slower than one GPU generates them; the last key is the serial (un-overlapped) estimate"}

    # ---- CPU baseline (rank 0, N == 1 only): oracle port on the host cores, bounded sample
    if rank == 0 and world == 1 and not args.no_cpu:
        from oracle import recipe
        ds = recipe.dataset_from_pulsars(psrs, noise)
        cores = min(os.cpu_count() or 1, 64)
        _, one = recipe.timed_realizations(ds, 1, 1)
        v, wall, done = cpu_arm(ds, cores * 4, cores, budget_s=args.cpu_seconds)
        line["cpu_baseline"] = {"value": v, "unit": "realizations/s", "cores": cores, "kind": "port",
                                "sample": f"{done} realizations of the same workload in {wall:.1f} s, one single-threaded process per core, "
                                          f"numpy oracle port (one core alone: {1 / one:.2f}/s); no PINT, no dense U => faster than the unmodified reference"}
    if rank == 0:
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
