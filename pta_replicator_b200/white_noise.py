"""EFAC/EQUAD white noise and ECORR (jitter) injection -- drop-in for
``/root/reference/pta_replicator/white_noise.py`` (``quantize_fast`` :7-44,
``add_measurement_noise`` :47-125, ``add_jitter`` :128-198).

Same names, arguments, ledger keys, error behaviour and -- because the standard normals are
drawn on the host from the global legacy ``np.random`` stream in the reference's order
 -- the same numbers for the same seeds.  The arithmetic runs in the fused
sm_100a generator (``ptar_generate``); there is no CPU path.
"""
from __future__ import annotations

import numpy as np

from .engine import PulsarBatch, greedy_buckets
from .simulate import SimulatedPulsar, TimeArray


def quantize_fast(times, flags=None, dt=1.0):
    """Bucket TOAs greedily in bins of ``dt`` days.  Returns ``(avetoas, U)`` or
    ``(avetoas, aveflags, U)`` like the reference; ``U`` is built only for API compatibility --
    the kernels consume the per-TOA bucket index (``quantize_index``)."""
    times = np.asarray(times, dtype=float)
    bucket, firsts = quantize_index(times, dt)
    nb = len(firsts)
    avetoas = np.bincount(bucket, weights=times, minlength=nb) / np.bincount(bucket, minlength=nb)
    U = np.zeros((len(times), nb), "d")
    U[np.arange(len(times)), bucket] = 1
    if flags is not None:
        return avetoas, np.asarray(flags)[firsts], U
    return avetoas, U


def quantize_index(times, dt):
    """``(bucket_of_toa, first_toa_of_bucket)`` in the caller's TOA order."""
    times = np.asarray(times, dtype=float)
    order = np.argsort(times, kind="stable")
    b_sorted = greedy_buckets(times[order], dt)
    bucket = np.empty(len(times), dtype=np.int64)
    bucket[order] = b_sorted
    nb = int(b_sorted[-1]) + 1 if len(times) else 0
    firsts = order[np.searchsorted(b_sorted, np.arange(nb), side="left")]
    return bucket, firsts


def add_measurement_noise(psr: SimulatedPulsar, efac: float = 1.0, log10_equad: float = None, flagid: str = "f",
                          flags: list = None, seed: int = None, tnequad: bool = False):
    """EFAC * (TOA error (+) EQUAD) white noise [default, t2equad] or EFAC*error (+) EQUAD."""
    equad_str = "tnequad" if tnequad else "t2equad"
    if seed is not None:
        np.random.seed(seed)
    batch = PulsarBatch([psr], exact_epochs=True)
    batch.set_white(0, efac=efac, log10_equad=log10_equad, flagid=flagid, flags=flags, tnequad=tnequad)
    n = psr.toas.ntoas
    z1 = np.random.randn(n)          # white_noise.py:105
    z2 = np.random.randn(n)          # :107/:109 -- drawn even when equad == 0
    torch = batch.torch
    inject = dict(z1=torch.from_numpy(batch.pack_table_order([z1]))[None], z2=torch.from_numpy(batch.pack_table_order([z2]))[None])
    row = batch.generate(1, inject=inject)[0]
    dt = TimeArray(batch.unpack(row, 0), "s")
    if flags is None:
        psr.update_added_signals("{}_measurement_noise".format(psr.name),
                                 {"efac": efac, "log10_" + equad_str: log10_equad}, dt)
    else:
        psr.update_added_signals("{}_measurement_noise".format(psr.name), {}, dt)
        for i, flag in enumerate(flags):
            psr.update_added_signals("{}_{}_measurement_noise".format(psr.name, flag),
                                     {"efac": efac[i], "log10_" + equad_str: log10_equad[i]})
    psr.toas.adjust_TOAs(dt.to("day"))
    psr.update_residuals()


def add_jitter(psr: SimulatedPulsar, log10_ecorr: float, flagid: str = "f", flags: list = None,
               coarsegrain: float = 0.1, seed: int = None):
    """Epoch-correlated (ECORR) noise of rms ``10**log10_ecorr`` s in buckets of ``coarsegrain`` days."""
    if seed is not None:
        np.random.seed(seed)
    batch = PulsarBatch([psr], exact_epochs=True)
    batch.set_ecorr(0, log10_ecorr, flagid=flagid, flags=flags, coarsegrain=coarsegrain)
    nb = len(batch._ecorr[0][1])
    zb = np.random.randn(nb)          # white_noise.py:182, bucket order = time order
    torch = batch.torch
    row = batch.generate(1, inject=dict(zb=torch.from_numpy(zb)[None]))[0]
    dt = TimeArray(batch.unpack(row, 0), "s")
    if flags is None:
        psr.update_added_signals("{}_jitter".format(psr.name), {"log10_ecorr": log10_ecorr}, dt)
    else:
        psr.update_added_signals("{}_jitter".format(psr.name), {}, dt)
        for i, flag in enumerate(flags):
            psr.update_added_signals("{}_{}_jitter".format(psr.name, flag), {"log10_ecorr": log10_ecorr[i]})
    psr.toas.adjust_TOAs(dt.to("day"))
    psr.update_residuals()


# north_star / libstempo spellings
def add_efac(psr, efac=1.0, flagid="f", flags=None, seed=None):
    """libstempo-style alias: EFAC only."""
    if flags is None:
        return add_measurement_noise(psr, efac=efac, log10_equad=None, flagid=flagid, flags=None, seed=seed)
    return add_measurement_noise(psr, efac=efac, log10_equad=np.full(len(flags), -300.0), flagid=flagid, flags=flags, seed=seed)


add_ecorr = add_jitter
