"""Shared loaders for the committed golden fixtures (tests/golden; see oracle/make_golden.py)."""
import os

import numpy as np

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_flags_case():
    """Return (npz, list of per-pulsar dicts) for tests/golden/ref_flags.npz."""
    z = np.load(os.path.join(GOLD, "ref_flags.npz"), allow_pickle=False)
    n = int(z["npsr"])
    psrs = []
    for i in range(n):
        raj, decj = z[f"raj_decj_{i}"]
        psrs.append(dict(
            name=str(z[f"name_{i}"]), loc={"RAJ": float(raj), "DECJ": float(decj)},
            mjd=z[f"mjd_{i}"], err_us=z[f"err_{i}"], flag=[str(s) for s in z[f"flag_{i}"]],
            backends=[str(s) for s in z[f"backends_{i}"]],
            efac=z[f"efac_{i}"], l10_equad=z[f"l10_equad_{i}"], l10_ecorr=z[f"l10_ecorr_{i}"],
            rn_l10A=float(z[f"rn_l10A_{i}"]), rn_gamma=float(z[f"rn_gamma_{i}"])))
    return z, psrs


def load_small_case():
    from pta_replicator_b200 import partim
    import glob
    pars = sorted(glob.glob(os.path.join(GOLD, "partim_small", "par", "*.par")))
    tims = sorted(glob.glob(os.path.join(GOLD, "partim_small", "tim", "*.tim")))
    psrs = []
    for p, t in zip(pars, tims):
        par = partim.read_par(p)
        c = partim.read_tim(t)
        psrs.append(dict(name=par["_name"], loc=par["_loc"], mjd=c["mjd"], err_us=c["err_us"], flags=c["flags"]))
    return np.load(os.path.join(GOLD, "ref_small.npz")), psrs


def rel_rms(a, b):
    """max|a-b| / rms(b)."""
    return float(np.max(np.abs(np.asarray(a) - np.asarray(b))) / np.sqrt(np.mean(np.square(b))))


# Waveforms of the burst / transient cases (shared by oracle/make_golden.py and the tests; t = TOA - tref in seconds)
F4_TREF = 55000.0 * 86400.0
F4_T0_MJD = 55800.0


def burst_plus(t):
    return 2.0e-7 * np.exp(-0.5 * ((t - 3.0e7) / 6.0e6) ** 2) * np.sin(2 * np.pi * t / 2.5e7)


def burst_cross(t):
    return 1.3e-7 * np.exp(-0.5 * ((t - 3.4e7) / 9.0e6) ** 2) * np.cos(2 * np.pi * t / 3.1e7)


def transient_waveform(t):
    return 5.0e-7 * np.where(t > 1.0e7, np.exp(-(t - 1.0e7) / 4.0e7), 0.0)


def outlier_population(seed=77, n=4000):
    """Synthetic SMBHB population in holodeck's layout: vals = [Mtot [g], q, z, f_obs [Hz]], weights, bin edges."""
    msol = 1.988409870698051e33
    rng = np.random.default_rng(seed)
    fobs = np.arange(1, 8) / (16.03 * 365.25 * 86400.0)              # 6 bins at multiples of 1/T
    vals = np.stack([10 ** rng.uniform(8.5, 10.3, n) * msol, rng.uniform(0.1, 1.0, n), 10 ** rng.uniform(-1.5, 0.3, n),
                     rng.uniform(fobs[0], fobs[-1], n)])
    weights = 10 ** rng.uniform(-1.0, 2.5, n)
    return vals, weights, fobs, 16.03 * 365.25 * 86400.0
