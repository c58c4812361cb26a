"""GPU tests on the reference's real NANOGrav 15-yr files (tests/golden/ref_real3.npz: B1855+09, B1937+21, J1909-3744;
unsorted TOAs, 7,758 / 23,023 / 35,037 per pulsar, 360 / 629 / 831 one-second ECORR buckets, ELONG/ELAT positions, the
B-name epoch-1950 rule) and of the Fourier-matrix options the injection path never uses (pshift, logf, fmin/fmax)."""
import os

import numpy as np
import pytest

from oracle import refnumpy as O
from tests.fixtures import GOLD
from tests.test_bridge_and_fixtures import real3_pulsars

pytestmark = pytest.mark.gpu


def _api_pulsars(spec):
    import pta_replicator_b200 as P
    out = []
    for s in spec:
        p = P.pulsar_from_arrays(s["name"], s["loc"], s["mjd"], s["err_us"], flags=[{"f": f} for f in s["flag"]])
        P.make_ideal(p)
        out.append(p)
    return out


def test_fourier_matrix_options_against_the_unmodified_reference():
    """create_fourier_design_matrix_red(pshift / logf / fmin / fmax / Tspan) on the device against outputs of the
    unmodified function (red_noise.py:36-103); pshift draws its phases from the global legacy stream like the reference."""
    import pta_replicator_b200 as P
    z = np.load(os.path.join(GOLD, "ref_fourier.npz"))
    t = z["t"]
    np.random.seed(4242)
    F, f = P.create_fourier_design_matrix_red(t, nmodes=20, pshift=True)
    assert np.array_equal(f, z["f_pshift"]) and np.max(np.abs(F - z["F_pshift"])) < 5e-16
    np.random.seed(4243)
    F, _ = P.create_fourier_design_matrix_red(t, nmodes=20, pshift=True, libstempo_convention=True)
    assert np.max(np.abs(F - z["F_pshift_ls"])) < 5e-16
    for key, kw in (("logf", dict(logf=True, fmin=2e-9, fmax=3e-7)), ("lin", dict(fmin=2e-9, fmax=3e-7)), ("tspan", dict(Tspan=6.0e8))):
        F, f = P.create_fourier_design_matrix_red(t, nmodes=20, **kw)
        assert np.array_equal(f, z[f"f_{key}"]) and np.max(np.abs(F - z[f"F_{key}"])) < 5e-16, key


def test_drop_in_injections_on_the_real_ng15_files():
    """add_measurement_noise / add_jitter / add_red_noise with the 15-yr noise dictionary and the legacy seeds on the
    real, unsorted TOAs: every 40th TOA against the unmodified reference, all TOAs against the oracle."""
    import pta_replicator_b200 as P
    spec, z = real3_pulsars()
    psrs = _api_pulsars(spec)
    for i, (p, s) in enumerate(zip(psrs, spec)):
        pp = s["pp"]
        be = np.array(pp["backends"])
        n = p.toas.ntoas
        P.add_measurement_noise(p, efac=np.asarray(pp["efac"]), log10_equad=np.asarray(pp["log10_equad"]), flagid="f", flags=be,
                                seed=10660 + i)
        wn = p.added_signals_time[f"{p.name}_measurement_noise"].to("s").value
        assert np.max(np.abs(wn[::40] - z[f"measurement_noise_{i}"])) < 1e-14 * np.sqrt(np.mean(wn ** 2))
        p.toas.table["tdbld"] = p.toas.table["tdbld"] - np.asarray(p.toas.delay_s, np.longdouble) / np.longdouble(86400)
        p.toas.delay_s = np.zeros(n)
        P.add_jitter(p, log10_ecorr=np.asarray(pp["log10_ecorr"]), flagid="f", flags=be, coarsegrain=1.0 / 86400.0, seed=17763 + i)
        jit = p.added_signals_time[f"{p.name}_jitter"].to("s").value
        assert np.max(np.abs(jit[::40] - z[f"jitter_{i}"])) < 1e-14 * np.sqrt(np.mean(jit ** 2))
        p.toas.table["tdbld"] = p.toas.table["tdbld"] - np.asarray(p.toas.delay_s, np.longdouble) / np.longdouble(86400)
        p.toas.delay_s = np.zeros(n)
        P.add_red_noise(p, pp["rn_log10_A"], pp["rn_gamma"], components=30, seed=19870 + i)
        rn = p.added_signals_time[f"{p.name}_red_noise"].to("s").value
        assert np.max(np.abs(rn[::40] - z[f"red_noise_{i}"])) < 1e-11 * np.sqrt(np.mean(rn ** 2))
        mjd = np.asarray(s["mjd"], dtype=float)
        (zr,) = O.legacy_randn(19870 + i, 60)
        full = O.red_noise(mjd, pp["rn_log10_A"], pp["rn_gamma"], zr)
        assert np.max(np.abs(rn - full)) < 1e-11 * np.sqrt(np.mean(full ** 2))


def test_batched_engine_on_the_real_ng15_files():
    """PulsarBatch on the real TOAs (Taylor epochs on real sub-band structure, 831 buckets, ELONG/ELAT -> ORF with the
    B1950 rule): all random terms with injected draws against the oracle, every TOA; then a Philox run."""
    import torch
    from pta_replicator_b200 import orf
    from pta_replicator_b200.engine import PulsarBatch
    spec, _ = real3_pulsars()
    psrs = _api_pulsars(spec)
    b = PulsarBatch(psrs)
    for i, s in enumerate(spec):
        pp, be = s["pp"], np.array(s["pp"]["backends"])
        b.set_white(i, efac=pp["efac"], log10_equad=pp["log10_equad"], flagid="f", flags=be)
        b.set_ecorr(i, pp["log10_ecorr"], flagid="f", flags=be, coarsegrain=1.0 / 86400.0)
        b.set_red(i, pp["rn_log10_A"], pp["rn_gamma"], components=30)
    b.set_gwb(-14.6733, 13.0 / 3.0)
    st = b.compile()
    assert st["n_epochs"] < 0.2 * b.n_toa_total            # real sub-banded epochs compress like the synthetic ones
    R, P = 3, 3
    rng = np.random.default_rng(8)
    Jg = st["gwb_T_Jreal"]
    z1 = rng.standard_normal((R, b.ld)); z2 = rng.standard_normal((R, b.ld))
    zb = rng.standard_normal((R, st["n_bucket_total"])); zrn = rng.standard_normal((R, P, 60)); zg = rng.standard_normal((R, P, Jg))
    out = b.generate(R, inject=dict(z1=torch.from_numpy(z1), z2=torch.from_numpy(z2), zb=torch.from_numpy(zb),
                                    zrn=torch.from_numpy(zrn), gwb_z=torch.from_numpy(zg))).cpu().numpy()
    g = b._gwb
    Nf = g["Nf"]
    locs = orf.psrlocs_from_pulsars(psrs)
    M = np.linalg.cholesky(g["ORF"])
    assert np.allclose(g["ORF"], O.orf_matrix([{"RAJ": ra * 12 / np.pi, "DECJ": np.degrees(dec)} for ra, dec in locs]), atol=1e-12)
    boff = st["psr_bucket_off"].cpu().numpy()
    mjds = [np.asarray(s["mjd"], dtype=float) for s in spec]
    worst = 0.0
    for r in range(R):
        w = np.zeros((P, Nf), complex)
        w[:, 1:Nf - 1] = zg[r, :, 0::2] + 1j * zg[r, :, 1::2]
        gw, _ = O.gwb_from_draws(dict(npts=g["npts"], dt=g["dt"], ut=g["ut"]), g["C"], M, w, mjds)
        for i, s in enumerate(spec):
            pp = s["pp"]
            n, off, o = b.ntoa[i], b.toa_off[i], b.order[i]
            flag = np.array(s["flag"])
            zz1 = np.empty(n); zz1[o] = z1[r, off:off + n]
            zz2 = np.empty(n); zz2[o] = z2[r, off:off + n]
            ef = O.per_toa_params(pp["efac"], pp["backends"], flag, n)
            eq = O.per_toa_params(10 ** np.asarray(pp["log10_equad"]), pp["backends"], flag, n)
            tot = O.white_noise(s["err_us"] * 1e-6, ef, eq, zz1, zz2)
            bk, firsts = O.epoch_buckets(mjds[i], 1.0 / 86400.0)
            ec = O.ecorr_per_bucket(10 ** np.asarray(pp["log10_ecorr"]), pp["backends"], flag, firsts)
            tot = tot + O.jitter(bk, ec, zb[r, boff[i]:boff[i] + len(firsts)])
            tot = tot + O.red_noise(mjds[i], pp["rn_log10_A"], pp["rn_gamma"], zrn[r, i]) + gw[i]
            worst = max(worst, float(np.max(np.abs(b.unpack(out[r], i) - tot)) / np.sqrt(np.mean(tot ** 2))))
    assert worst < 1e-11, worst
    x = b.generate(64, seed=3)
    assert torch.isfinite(x).all() and float(x.std()) > 1e-8
