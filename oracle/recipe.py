"""TEST INFRASTRUCTURE ONLY -- the 15-yr recipe (examples/add_noise.ipynb cells 9, 11) for ONE
realization through the numpy oracle, on plain (picklable) arrays.  Used as the checker for the
batched engine and as the CPU arm of ``bench.py`` (``cpu_baseline`` / ``--impl reference``).

Like the reference, nothing is cached between realizations: every call re-buckets the TOAs,
rebuilds the Fourier basis, the ORF and its Cholesky factor.  Unlike the reference
it skips PINT's ``adjust_TOAs`` + ``Residuals`` after each injection and the dense ``U`` matrix of
``quantize_fast`` -- so it is FASTER than the real reference (0.167 realizations/s/core measured
for the unmodified functions in the authoring container).
"""
from __future__ import annotations

import numpy as np

from oracle import refnumpy as O


def dataset_from_pulsars(psrs, noise, gw_log10_A=-14.6733, gw_gamma=13.0 / 3.0, components=30,
                         coarsegrain=1.0 / 86400.0):
    ds = dict(gw_log10_A=gw_log10_A, gw_gamma=gw_gamma, components=components, coarsegrain=coarsegrain, psrs=[])
    for p in psrs:
        pp = noise[p.name]
        ds["psrs"].append(dict(
            name=p.name, loc=dict(p.loc), mjd=np.asarray(p.toas.get_mjds().value, float),
            tdb=np.asarray(p.toas.table["tdbld"], float), err_s=np.asarray(p.toas.get_errors().to("s").value, float),
            flag=np.array([f.get("f") for f in p.toas.table["flags"]]),
            backends=list(pp["backends"]), efac=np.asarray(pp["efac"]), l10_equad=np.asarray(pp["log10_equad"]),
            l10_ecorr=np.asarray(pp["log10_ecorr"]), rn_l10A=pp["rn_log10_A"], rn_gamma=pp["rn_gamma"]))
    return ds


def realization(ds, seed, white=True, ecorr=True, red=True, gwb=True):
    """One realization of the recipe with the legacy global-stream draws, seeds as in the notebook
    (cell 8: 10660 / 17763 / 19870 + pulsar index, 16672 for the GWB) offset by ``seed``."""
    out = []
    P = ds["psrs"]
    for i, p in enumerate(P):
        n = len(p["mjd"])
        d = np.zeros(n)
        if white:
            ef = O.per_toa_params(p["efac"], p["backends"], p["flag"], n)
            eq = O.per_toa_params(10 ** p["l10_equad"], p["backends"], p["flag"], n)
            z1, z2 = O.legacy_randn(10660 + i + 1000 * seed, n, n)
            d += O.white_noise(p["err_s"], ef, eq, z1, z2)
        if ecorr:
            b, firsts = O.epoch_buckets(p["mjd"], ds["coarsegrain"])
            ec = O.ecorr_per_bucket(10 ** p["l10_ecorr"], p["backends"], p["flag"], firsts)
            (zb,) = O.legacy_randn(17763 + i + 1000 * seed, len(firsts))
            d += O.jitter(b, ec, zb)
        if red:
            (zr,) = O.legacy_randn(19870 + i + 1000 * seed, 2 * ds["components"])
            d += O.red_noise(p["tdb"], p["rn_l10A"], p["rn_gamma"], zr, components=ds["components"])
        out.append(d)
    if gwb:
        setup = O.gwb_grid_setup([p["mjd"].min() for p in P], [p["mjd"].max() for p in P])
        nf = len(setup["f"])
        draws = O.legacy_randn(16672 + 1000 * seed, *([nf] * (2 * len(P))))
        w = np.array([draws[2 * i] + 1j * draws[2 * i + 1] for i in range(len(P))])
        Cf = O.gwb_spectrum(setup["f"], setup["dur"], 10, ds["gw_log10_A"], ds["gw_gamma"])
        M = np.linalg.cholesky(O.orf_matrix([p["loc"] for p in P]))
        g, _ = O.gwb_from_draws(setup, Cf, M, w, [p["mjd"] for p in P])
        for i in range(len(P)):
            out[i] += g[i]
    return out


_DS = None  # inherited by forked workers; tasks carry only a seed


def _init_worker():
    """One BLAS/OpenMP thread per worker process (the pool supplies the parallelism)."""
    try:
        from threadpoolctl import threadpool_limits
        global _LIMIT
        _LIMIT = threadpool_limits(1)
    except Exception:
        pass


def _worker(seed):
    r = realization(_DS, seed)
    return float(sum(np.sum(x) for x in r))


def timed_realizations(ds, n_real, n_proc, budget_s=None):
    """Run up to ``n_real`` realizations on ``n_proc`` processes (1 numpy thread each); stop handing out
    work once ``budget_s`` seconds have passed.  Returns (realizations completed, wall seconds)."""
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
        pool.map(_worker, [10**6 + k for k in range(n_proc)], chunksize=1)  # warm every worker (untimed)
        t0 = time.perf_counter()
        done = 0
        for _ in pool.imap_unordered(_worker, range(n_real), chunksize=1):
            done += 1
            if budget_s is not None and time.perf_counter() - t0 > budget_s:
                pool.terminate()
                break
        return done, time.perf_counter() - t0
