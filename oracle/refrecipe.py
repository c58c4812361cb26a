"""TEST INFRASTRUCTURE ONLY -- the 15-yr recipe (examples/add_noise.ipynb cells 9, 11) for ONE realization through
the UNMODIFIED reference functions ``add_measurement_noise`` (white_noise.py:47), ``add_jitter`` (:128),
``add_red_noise`` (red_noise.py:106) and ``add_gwb`` (:138), run under the ``sys.modules`` stubs of
``oracle/refstubs.py`` on duck-typed pulsars (one fresh set per realization, like the reference: one
``SimulatedPulsar`` = one realization).  This is the CPU arm of ``bench.py`` (``--impl reference`` and
``cpu_baseline``, kind "reference"); PINT's ``adjust_TOAs`` / ``Residuals`` are stubs, which flatters the
reference.  ``oracle/recipe.py`` is the numpy port of the same recipe (kind "port", faster).
"""
from __future__ import annotations

import numpy as np

from oracle import refstubs


def dataset_from_pulsars(psrs, noise, gw_log10_A=-14.6733, gw_gamma=13.0 / 3.0, components=30,
                         coarsegrain=1.0 / 86400.0):
    ds = dict(gw_log10_A=gw_log10_A, gw_gamma=gw_gamma, components=components, coarsegrain=coarsegrain, psrs=[])
    for p in psrs:
        pp = noise[p.name]
        ds["psrs"].append(dict(
            name=p.name, loc=dict(p.loc), mjd=np.asarray(p.toas.table["tdbld"], dtype=np.longdouble),
            err_us=np.asarray(p.toas.get_errors().to("us").value, float),
            flags=[dict(f) for f in p.toas.table["flags"]],
            backends=list(pp["backends"]), efac=np.asarray(pp["efac"], float),
            l10_equad=np.asarray(pp["log10_equad"], float), l10_ecorr=np.asarray(pp["log10_ecorr"], float),
            rn_l10A=float(pp["rn_log10_A"]), rn_gamma=float(pp["rn_gamma"])))
    return ds


def realization(ds, seed):
    """One realization; returns the per-pulsar sum of the injected delays [s] (table order)."""
    ref = refstubs.reference_modules()
    psrs = [refstubs.StubPulsar(p["name"], p["loc"], p["mjd"], p["err_us"], p["flags"], freeze_toas=True) for p in ds["psrs"]]
    for i, (sp, p) in enumerate(zip(psrs, ds["psrs"])):
        ref.white_noise.add_measurement_noise(sp, efac=p["efac"], log10_equad=p["l10_equad"], flagid="f",
                                              flags=p["backends"], seed=10660 + i + 1000 * seed)
        ref.white_noise.add_jitter(sp, log10_ecorr=p["l10_ecorr"], flagid="f", flags=p["backends"],
                                   coarsegrain=ds["coarsegrain"], seed=17763 + i + 1000 * seed)
        ref.red_noise.add_red_noise(sp, p["rn_l10A"], p["rn_gamma"], components=ds["components"],
                                    seed=19870 + i + 1000 * seed)
    ref.red_noise.add_gwb(psrs, ds["gw_log10_A"], ds["gw_gamma"], seed=16672 + 1000 * seed)
    return [np.asarray(sp.toas.delta, dtype=float) for sp in psrs]


_DS = None


def _init_worker():
    try:
        from threadpoolctl import threadpool_limits
        global _LIMIT
        _LIMIT = threadpool_limits(1)
    except Exception:
        pass
    refstubs.reference_modules()


def _worker(seed):
    return float(sum(np.sum(x) for x in realization(_DS, seed)))


def timed_realizations(ds, n_real, n_proc, budget_s=None):
    """Same contract as ``oracle.recipe.timed_realizations``: (realizations completed, wall seconds); one
    single-threaded process per core, pool start-up and one warm-up realization per worker untimed."""
    import multiprocessing as mp
    import time
    global _DS
    _DS = ds
    if n_proc <= 1:
        _init_worker()
        t0 = time.perf_counter()
        done = 0
        for s in range(n_real):
            _worker(s)
            done += 1
            if budget_s is not None and time.perf_counter() - t0 > budget_s:
                break
        return done, time.perf_counter() - t0
    ctx = mp.get_context("fork")
    with ctx.Pool(n_proc, initializer=_init_worker) as pool:
        pool.map(_worker, [10**6 + k for k in range(n_proc)], chunksize=1)
        t0 = time.perf_counter()
        done = 0
        for _ in pool.imap_unordered(_worker, range(n_real), chunksize=1):
            done += 1
            if budget_s is not None and time.perf_counter() - t0 > budget_s:
                pool.terminate()
                break
        return done, time.perf_counter() - t0
