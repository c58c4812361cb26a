"""Pin the numpy oracle (oracle/refnumpy.py) to the reference.

Two anchors: the reference's own libstempo golden vector, and outputs of
the UNMODIFIED reference functions run under the stub harness in the authoring
container (tests/golden/ref_*.npz, generator oracle/make_golden.py).  CPU only.
"""
import os

import numpy as np
import pytest

from oracle import refnumpy as O
from tests.fixtures import GOLD, load_flags_case, load_small_case, rel_rms

TIGHT = 1e-12   # relative to the signal's rms: same arithmetic, different summation order


def test_constants_match_reference_values():
    from pta_replicator_b200 import constants as C
    assert (C.SOLAR2S, C.KPC2S, C.MPC2S, C.YEAR_IN_SEC) == (O.SOLAR2S, O.KPC2S, O.MPC2S, O.YEAR)


def _small_gwb(z, psrs, nf):
    first = [float(np.min(p["mjd"])) for p in psrs]
    last = [float(np.max(p["mjd"])) for p in psrs]
    setup = O.gwb_grid_setup(first, last, nf=nf)
    return setup


def test_small_recipe_per_signal():
    z, psrs = load_small_case()
    nf = int(z["nf"])
    # GWB: draws in reference order
    first = [float(np.min(p["mjd"])) for p in psrs]
    last = [float(np.max(p["mjd"])) + float(z["last_mjd_nudge_days"]) for p in psrs]
    setup = O.gwb_grid_setup(first, [max(last)] * 3)
    assert len(setup["f"]) == nf
    draws = O.legacy_randn(123456, *([nf] * 6))
    w = np.array([draws[2 * i] + 1j * draws[2 * i + 1] for i in range(3)])
    C = O.gwb_spectrum(setup["f"], setup["dur"], 10, -14, 4.33)
    M = np.linalg.cholesky(O.orf_matrix([p["loc"] for p in psrs]))
    res, _ = O.gwb_from_draws(setup, C, M, w, [np.asarray(p["mjd"], float) for p in psrs])
    for i in range(3):
        assert rel_rms(res[i], z["gwb"][i]) < TIGHT
    for i, p in enumerate(psrs):
        n = len(p["mjd"])
        z1, z2 = O.legacy_randn(54321 + i, n, n)
        wn = O.white_noise(p["err_us"] * 1e-6, np.ones(n), np.zeros(n), z1, z2)
        assert rel_rms(wn, z["measurement_noise"][i]) < 1e-15
        bucket, firsts = O.epoch_buckets(np.asarray(p["mjd"], float), 0.1)
        (zb,) = O.legacy_randn(54321 + i, len(firsts))
        jit = O.jitter(bucket, np.ones(len(firsts)) * 10 ** np.log10(3e-7), zb)
        assert rel_rms(jit, z["jitter"][i]) < 1e-15
        (zr,) = O.legacy_randn(12345 + i, 60)
        rn = O.red_noise(p["mjd"], -15, 4.2, zr, components=30, libstempo_convention=True)
        assert rel_rms(rn, z["red_noise"][i]) < TIGHT
        cw = O.cgw(np.asarray(p["mjd"], float), p["loc"], np.pi / 2, 2.5, 1e9, 5.0, 1e-8, 0.5, 1.5, np.pi / 4,
                   pdist=1.0, psrTerm=True, evolve=True, tref=53000 * 86400)
        assert rel_rms(cw, z["cw"][i]) < TIGHT


def test_small_recipe_reproduces_libstempo_golden():
    """Sum of the oracle's signals minus the mean vs the reference's golden npz, with the
    reference test's own criterion (tests/test_against_libstempo.py:66) made two-sided."""
    z, psrs = load_small_case()
    gold = np.load(os.path.join(GOLD, "libstempo_golden.npz"))["residuals"]
    total = z["gwb"] + z["measurement_noise"] + z["jitter"] + z["red_noise"] + z["cw"]
    res = total - total.mean(axis=1, keepdims=True)
    assert np.all(np.abs(res - gold) / np.sqrt(np.mean(res**2)) < 1e-3)


def test_flags_case_white_jitter_rn():
    z, psrs = load_flags_case()
    for i, p in enumerate(psrs):
        n = len(p["mjd"])
        err = p["err_us"] * 1e-6
        ef = O.per_toa_params(p["efac"], p["backends"], p["flag"], n)
        eq = O.per_toa_params(10 ** p["l10_equad"], p["backends"], p["flag"], n)
        z1, z2 = O.legacy_randn(10660 + i, n, n)
        assert rel_rms(O.white_noise(err, ef, eq, z1, z2), z[f"wn_flags_{i}"]) < 1e-15
        z1, z2 = O.legacy_randn(333 + i, n, n)
        assert rel_rms(O.white_noise(err, ef, eq, z1, z2, tnequad=True), z[f"wn_tn_{i}"]) < 1e-15
        z1, z2 = O.legacy_randn(777 + i, n, n)
        assert rel_rms(O.white_noise(err, np.ones(n) * 1.1, np.ones(n) * 10 ** -6.5, z1, z2), z[f"wn_scalar_{i}"]) < 1e-15
        b, firsts = O.epoch_buckets(p["mjd"], 1.0 / 86400.0)
        ec = O.ecorr_per_bucket(10 ** p["l10_ecorr"], p["backends"], p["flag"], firsts)
        (zb,) = O.legacy_randn(17763 + i, len(firsts))
        assert rel_rms(O.jitter(b, ec, zb), z[f"jit_flags_{i}"]) < 1e-15
        b, firsts = O.epoch_buckets(p["mjd"], 0.1)
        (zb,) = O.legacy_randn(444 + i, len(firsts))
        assert rel_rms(O.jitter(b, O.ecorr_per_bucket(10 ** -6.7, None, None, firsts), zb), z[f"jit_scalar_{i}"]) < 1e-15
        (zr,) = O.legacy_randn(19870 + i, 60)
        assert rel_rms(O.red_noise(p["mjd"], p["rn_l10A"], p["rn_gamma"], zr), z[f"rn_default_{i}"]) < TIGHT
        (zr,) = O.legacy_randn(555 + i, 24)
        assert rel_rms(O.red_noise(p["mjd"], -13.7, 3.1, zr, components=12, libstempo_convention=True), z[f"rn_ls12_{i}"]) < TIGHT
        (zr,) = O.legacy_randn(888 + i, 10)
        assert rel_rms(O.red_noise(p["mjd"], -14.0, 2.5, zr, modes=z["rn_modes"]), z[f"rn_modes_{i}"]) < TIGHT


GWB_VARIANTS = {
    "hd": dict(l10A=-14.2, gamma=13.0 / 3.0, seed=16672),
    "turnover": dict(l10A=-14.0, gamma=4.0, seed=16673, turnover=True, f0=3e-9, beta=1.2, power=1.5),
    "nocorr": dict(l10A=-14.3, gamma=3.0, seed=16674, no_correlations=True),
    "userspec": dict(l10A=-14.0, gamma=4.0, seed=16675, userspec=True),
    "npts300": dict(l10A=-14.1, gamma=13.0 / 3.0, seed=16677, npts=300, howml=4),
}


@pytest.mark.parametrize("tag", sorted(GWB_VARIANTS))
def test_flags_case_gwb(tag):
    z, psrs = load_flags_case()
    kw = GWB_VARIANTS[tag]
    npts, howml = kw.get("npts", 600), kw.get("howml", 10)
    setup = O.gwb_grid_setup([p["mjd"].min() for p in psrs], [p["mjd"].max() for p in psrs], npts=npts, howml=howml)
    nf = len(setup["f"])
    assert nf == int(z[f"gwb_{tag}_nf"])
    n = len(psrs)
    draws = O.legacy_randn(kw["seed"], *([nf] * (2 * n)))
    w = np.array([draws[2 * i] + 1j * draws[2 * i + 1] for i in range(n)])
    C = O.gwb_spectrum(setup["f"], setup["dur"], howml, kw["l10A"], kw["gamma"], turnover=kw.get("turnover", False),
                       f0=kw.get("f0", 1e-9), beta=kw.get("beta", 1), power=kw.get("power", 1),
                       userSpec=z["gwb_userspec"] if kw.get("userspec") else None)
    M = np.linalg.cholesky(O.orf_matrix([p["loc"] for p in psrs], no_correlations=kw.get("no_correlations", False)))
    res, _ = O.gwb_from_draws(setup, C, M, w, [p["mjd"] for p in psrs])
    for i in range(n):
        assert rel_rms(res[i], z[f"gwb_{tag}_{i}"]) < TIGHT


CGW_BASE = dict(gwtheta=1.1, gwphi=4.0, mc=3e9, dist=40.0, fgw=2.2e-8, phase0=1.3, psi=0.4, inc=1.0, tref=53000 * 86400)
CGW_VARIANTS = {
    "evolve": dict(CGW_BASE, pdist=1.3, psrTerm=True, evolve=True),
    "earth": dict(CGW_BASE, psrTerm=False, evolve=True),
    "approx": dict(CGW_BASE, pdist=0.9, psrTerm=True, evolve=False, phase_approx=True),
    "mono": dict(CGW_BASE, pdist=0.9, psrTerm=True, evolve=False, phase_approx=False),
    "pphase": dict(CGW_BASE, pphase=2.0, psrTerm=True, evolve=True),
}


@pytest.mark.parametrize("tag", sorted(CGW_VARIANTS))
def test_flags_case_cgw(tag):
    z, psrs = load_flags_case()
    for i, p in enumerate(psrs):
        got = O.cgw(p["mjd"], p["loc"], **CGW_VARIANTS[tag])
        assert rel_rms(got, z[f"cgw_{tag}_{i}"]) < TIGHT


def test_hd_basis_against_reference():
    z = np.load(os.path.join(GOLD, "ref_orf.npz"))
    got = O.hd_basis_l0(z["locs_hd"])
    assert np.max(np.abs(got - z["basis_hd"][0])) < 1e-15
    got = O.hd_basis_l0(z["locs"])
    assert np.max(np.abs(got - z["basis_l6"][0])) < 1e-15


CAT_VARIANTS = {"evolve": dict(pdist=1.2, psrTerm=True, evolve=True), "earth": dict(psrTerm=False, evolve=True),
                "approx": dict(pdist=0.8, psrTerm=True, evolve=False, phase_approx=True),
                "mono": dict(pdist=0.8, psrTerm=True, evolve=False, phase_approx=False),
                "pphase": dict(pphase=1.5, psrTerm=True, evolve=True)}


def load_catalog():
    c = np.load(os.path.join(GOLD, "ref_catalog.npz"))
    return c, {k: c[f"cat_{k}"] for k in ("gwtheta", "gwphi", "mc", "dist", "fgw", "phase0", "psi", "inc")}


@pytest.mark.parametrize("tag", sorted(CAT_VARIANTS))
def test_cw_catalog_oracle_against_reference_numba_loops(tag):
    """``add_catalog_of_cws`` (numba ``loop_over_CWs``), 300 sources incl. one that merges inside the data span.
    Evolving branches: numpy vs LLVM ``pow`` differ by an ulp, amplified like in ``add_cgw`` (see test_gpu_parity)."""
    _, spec = load_flags_case()
    c, cat = load_catalog()
    tol = 1e-9 if tag in ("evolve", "earth", "pphase") else TIGHT
    for i, s in enumerate(spec):
        got = O.cw_catalog(s["mjd"], s["loc"], cat, tref=53000 * 86400, **CAT_VARIANTS[tag])
        assert rel_rms(got, c[f"cat_{tag}_{i}"]) < tol


def test_burst_transient_memory_against_reference_outputs():
    """Rows f4 of SURVEY.md section 8: the restatements of add_burst / add_noise_transient / add_gw_memory against
    the unmodified reference run on the 4 synthetic pulsars (tests/golden/ref_f4.npz, oracle/make_golden.py:make_f4)."""
    from tests import fixtures as fx
    z = np.load(os.path.join(fx.GOLD, "ref_f4.npz"))
    _, spec = fx.load_flags_case()
    for i, s in enumerate(spec):
        for tag, quad in (("burst", False), ("burstq", True)):
            got = O.burst(s["mjd"], s["loc"], 1.1, 4.0, fx.burst_plus, fx.burst_cross, psi=0.7, tref=fx.F4_TREF, remove_quad=quad)
            assert fx.rel_rms(got, z[f"{tag}_{i}"]) < 1e-12, (tag, i)
        assert np.array_equal(O.noise_transient(s["mjd"], fx.transient_waveform, tref=fx.F4_TREF), z[f"transient_{i}"])
        got = O.gw_memory(s["mjd"], s["loc"], 3.0e-14, 0.9, 2.2, 0.4, fx.F4_T0_MJD)
        assert fx.rel_rms(got, z[f"memory_{i}"]) < 1e-13, i
        assert (z[f"memory_{i}"] == 0).any() and (z[f"memory_{i}"] != 0).any()
