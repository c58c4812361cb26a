"""GPU tests at the benchmark's full size (synthetic ng15-full: 67 pulsars, sum N_toa = 639,453; BASELINE.json
configs 2/3/5): parity with injected draws against the numpy oracle for every pulsar, and size-independent
properties of the throughput mode (linearity over signals, shard invariance, finite output)."""
import numpy as np
import pytest

from oracle import refnumpy as O

pytestmark = pytest.mark.gpu

CGW = dict(gwtheta=np.pi / 2, gwphi=2.5, mc=1e9, dist=5.0, fgw=1e-8, phase0=0.5, psi=1.5, inc=np.pi / 4, pdist=1.0,
           psrTerm=True, evolve=True, tref=53000 * 86400)


@pytest.fixture(scope="module")
def ng15():
    from pta_replicator_b200 import synthetic
    psrs, noise = synthetic.make_ng15_like("full")
    return psrs, noise


def test_full_size_parity_all_signals_with_injected_draws(ng15):
    """Config 3 (EFAC/EQUAD + ECORR + RN + HD GWB + CGW, the libstempo-test CGW parameters) on all 67
    pulsars, 2 realizations, draws injected; epoch (Taylor) mode as used by the benchmark."""
    import torch
    from pta_replicator_b200 import synthetic
    from pta_replicator_b200.engine import PulsarBatch
    psrs, noise = ng15
    b = PulsarBatch(psrs)
    synthetic.ng15_recipe(b, noise)
    b.add_cgw(**CGW)
    st = b.compile()
    R, P = 2, len(psrs)
    rng = np.random.default_rng(99)
    Jg = st["gwb_T_Jreal"]
    z1 = rng.standard_normal((R, b.ld)); z2 = rng.standard_normal((R, b.ld))
    zb = rng.standard_normal((R, st["n_bucket_total"])); zrn = rng.standard_normal((R, P, 60))
    zg = rng.standard_normal((R, P, Jg))
    out = b.generate(R, inject=dict(z1=torch.from_numpy(z1), z2=torch.from_numpy(z2), zb=torch.from_numpy(zb),
                                    zrn=torch.from_numpy(zrn), gwb_z=torch.from_numpy(zg))).cpu().numpy()
    g = b._gwb
    Nf = g["Nf"]
    M = np.linalg.cholesky(g["ORF"])
    boff = st["psr_bucket_off"].cpu().numpy()
    mjds = [np.asarray(p.toas.get_mjds().value) for p in psrs]
    worst = 0.0
    for r in range(R):
        w = np.zeros((P, Nf), complex)
        w[:, 1:Nf - 1] = zg[r, :, 0::2] + 1j * zg[r, :, 1::2]
        gw, _ = O.gwb_from_draws(dict(npts=g["npts"], dt=g["dt"], ut=g["ut"]), g["C"], M, w, mjds)
        for i, p in enumerate(psrs):
            pp = noise[p.name]
            n, off, o = b.ntoa[i], b.toa_off[i], b.order[i]
            flag = np.array([f["f"] for f in p.toas.table["flags"]])
            zz1 = np.empty(n); zz1[o] = z1[r, off:off + n]
            zz2 = np.empty(n); zz2[o] = z2[r, off:off + n]
            ef = O.per_toa_params(pp["efac"], pp["backends"], flag, n)
            eq = O.per_toa_params(10 ** pp["log10_equad"], pp["backends"], flag, n)
            tot = O.white_noise(p.toas.get_errors().to("s").value, ef, eq, zz1, zz2)
            bk, firsts = O.epoch_buckets(mjds[i], 1.0 / 86400.0)
            ec = O.ecorr_per_bucket(10 ** pp["log10_ecorr"], pp["backends"], flag, firsts)
            tot = tot + O.jitter(bk, ec, zb[r, boff[i]:boff[i] + len(firsts)])
            tot = tot + O.red_noise(mjds[i], pp["rn_log10_A"], pp["rn_gamma"], zrn[r, i])
            tot = tot + gw[i] + O.cgw(mjds[i], p.loc, **{k: v for k, v in CGW.items()})
            got = b.unpack(out[r], i)
            worst = max(worst, float(np.max(np.abs(got - tot)) / np.sqrt(np.mean(tot**2))))
    assert worst < 1e-10, worst


def test_full_size_linearity_and_shards(ng15):
    """Philox streams are keyed by signal kind, so the all-signal output equals the sum of single-signal
    runs with the same seed (to fp64 addition order); shards reproduce bitwise; output is finite."""
    import torch
    from pta_replicator_b200 import synthetic
    from pta_replicator_b200.engine import PulsarBatch
    psrs, noise = ng15
    R, seed = 8, 31337
    parts = []
    for kw in (dict(white=True, ecorr=False, red=False, gwb=False), dict(white=False, ecorr=True, red=False, gwb=False),
               dict(white=False, ecorr=False, red=True, gwb=False), dict(white=False, ecorr=False, red=False, gwb=True)):
        b = PulsarBatch(psrs)
        synthetic.ng15_recipe(b, noise, **kw)
        parts.append(b.generate(R, seed=seed, real0=64))
        del b
    b = PulsarBatch(psrs)
    synthetic.ng15_recipe(b, noise)
    full = b.generate(R, seed=seed, real0=64)
    assert torch.isfinite(full).all()
    total = parts[0] + parts[1] + parts[2] + parts[3]
    scale = full.std().item()
    assert (full - total).abs().max().item() < 1e-12 * scale
    a = b.generate(4, seed=seed, real0=64)
    c = b.generate(4, seed=seed, real0=68)
    assert torch.equal(full[:4], a) and torch.equal(full[4:], c)
    # rms per signal is in the expected range (white ~ us; a 15-yr GWB at A = 10^-14.67 ~ tens of us)
    assert 1e-7 < parts[0].std().item() < 1e-5 and 1e-7 < parts[3].std().item() < 1e-3
