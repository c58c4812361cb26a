"""TEST INFRASTRUCTURE ONLY -- numpy restatement of the reference hot path.

Array-in / array-out versions of the algorithms in ``/root/reference/pta_replicator``
(``white_noise.py``, ``red_noise.py``, ``deterministic.py``, ``spharmORFbasis.py``).
Every function names the reference lines it restates.  Pinned by
``tests/test_oracle_pinning.py`` against the reference's libstempo golden vector and
against outputs of the unmodified reference run under ``oracle/refstubs.py``
(``tests/golden/ref_*.npz``; generator: ``oracle/make_golden.py``).

Never imported by the product package.
"""
from __future__ import annotations

import numpy as np

DAY = 86400.0
YEAR = 365.25 * DAY            # constants.py:3-4
F1YR_GWB = 1 / 3.16e7          # red_noise.py:248
SOLAR2S = 4.925838061995516e-06    # constants.py:6 (value checked in make_golden.py)
KPC2S = 102927125054.33899         # constants.py:7
MPC2S = 102927125054338.98         # constants.py:8


# --------------------------------------------------------------------------- RNG order
def legacy_randn(seed, *sizes):
    """Draws in the order the reference consumes them: one
    ``np.random.seed(seed)`` then consecutive ``np.random.randn(n)`` calls."""
    rs = np.random.RandomState(seed)
    return [rs.randn(n) for n in sizes]


# --------------------------------------------------------------------------- white + ECORR
def epoch_buckets(mjd, width_days):
    """Greedy time buckets of ``quantize_fast`` (white_noise.py:21-31).

    Returns ``(bucket_of_toa[int64 N], first_toa_of_bucket[int64 Nb])`` with buckets
    numbered in time order; a TOA opens a new bucket when it lies ``>= width`` after the
    *first* TOA of the current bucket.
    """
    mjd = np.asarray(mjd, dtype=float)
    order = np.argsort(mjd)
    bucket = np.empty(len(mjd), dtype=np.int64)
    firsts = [order[0]]
    ref = mjd[order[0]]
    bucket[order[0]] = 0
    for i in order[1:]:
        if not (mjd[i] - ref < width_days):
            firsts.append(i)
            ref = mjd[i]
        bucket[i] = len(firsts) - 1
    return bucket, np.asarray(firsts, dtype=np.int64)


def per_toa_params(values, flags, toa_flags, n):
    """Expand per-backend values onto TOAs (white_noise.py:95-101): TOAs whose flag is
    not listed keep 0; scalar ``values`` with ``flags=None`` broadcast (:92-93)."""
    out = np.zeros(n)
    if flags is None:
        return np.ones(n) * values
    tf = np.asarray(toa_flags)
    for v, f in zip(values, flags):
        out[tf == f] = v
    return out


def white_noise(err_s, efac_toa, equad_toa, z1, z2, tnequad=False):
    """white_noise.py:105-109."""
    dt = efac_toa * err_s * z1
    if tnequad:
        dt = dt + equad_toa * z2
    else:
        dt = dt + efac_toa * equad_toa * z2
    return dt


def jitter(bucket_of_toa, ecorr_bucket, zb):
    """white_noise.py:182: ``dot(U*ecorrvec, z)``; U has a single 1 per row."""
    return ecorr_bucket[bucket_of_toa] * zb[bucket_of_toa]


def ecorr_per_bucket(values, flags, toa_flags, firsts):
    """white_noise.py:165-178: scalar broadcast, or by the flag of each bucket's FIRST TOA."""
    if flags is None:
        return np.ones(len(firsts)) * values
    bf = np.asarray(toa_flags)[firsts]
    out = np.zeros(len(firsts))
    for v, f in zip(values, flags):
        out[bf == f] = v
    return out


# --------------------------------------------------------------------------- red noise
def fourier_basis(t_sec, nmodes=30, Tspan=None, libstempo_convention=False, modes=None, logf=False, fmin=None, fmax=None,
                  ranphase=None):
    """red_noise.py:61-101.  ``ranphase`` = the ``pshift`` phases (the reference draws them with
    ``np.random.uniform(0, 2 pi, nmodes)`` at :83; pass that array).  Returns ``(F[N,2K], Ffreqs[2K])``."""
    t = np.asarray(t_sec, dtype=float)
    T = Tspan if Tspan is not None else t.max() - t.min()
    if modes is not None:
        f = np.asarray(modes, dtype=float)
    elif fmin is None and fmax is None and not logf:
        f = 1.0 * np.arange(1, nmodes + 1) / T
    else:
        fmin = 1 / T if fmin is None else fmin
        fmax = nmodes / T if fmax is None else fmax
        f = np.logspace(np.log10(fmin), np.log10(fmax), nmodes) if logf else np.linspace(fmin, fmax, nmodes)
    ph = np.zeros(len(f)) if ranphase is None else np.asarray(ranphase, dtype=float)
    F = np.zeros((len(t), 2 * len(f)))
    if libstempo_convention:
        arg = 2 * np.pi * (t[:, None] - t[0, None]) * f[None, :] + ph[None, :]
        F[:, 0::2] = np.cos(arg)
        F[:, 1::2] = np.sin(arg)
    else:
        arg = 2 * np.pi * t[:, None] * f[None, :] + ph[None, :]
        F[:, 0::2] = np.sin(arg)
        F[:, 1::2] = np.cos(arg)
    return F, np.repeat(f, 2)


def red_noise_prior(freqs2, log10_A, gamma, Tspan):
    """red_noise.py:126."""
    A = 10 ** log10_A
    return A**2 * (freqs2 / (1 / YEAR)) ** (-gamma) / (12 * np.pi**2 * Tspan) * YEAR**3


def red_noise(tdb_mjd, log10_A, gamma, z, components=30, libstempo_convention=False, modes=None):
    """red_noise.py:123-128 with the 2K draws ``z`` supplied."""
    t = np.asarray(tdb_mjd, dtype=float) * DAY
    T = t.max() - t.min()
    F, ff = fourier_basis(t, nmodes=components, Tspan=T, libstempo_convention=libstempo_convention, modes=modes)
    y = np.sqrt(red_noise_prior(ff, log10_A, gamma, T)) * z
    return F @ y


# --------------------------------------------------------------------------- GWB
def gwb_grid_setup(first_mjds, last_mjds, npts=600, howml=10, nf=None):
    """red_noise.py:182-197, :230-232.  ``nf`` overrides the fragile ``len(arange)``."""
    start = float(np.min(np.asarray(first_mjds, float) * 86400) - 86400)
    stop = float(np.max(np.asarray(last_mjds, float) * 86400) + 86400)
    dur = stop - start
    ut = np.linspace(start, stop, npts)
    dt = dur / npts
    f = np.arange(0, 1 / (2 * dt), 1 / (dur * howml))
    if nf is not None and nf != len(f):
        f = np.arange(nf) * (1 / (dur * howml))
    f[0] = f[1]
    return dict(start=start, stop=stop, dur=dur, ut=ut, dt=dt, f=f, npts=npts, howml=howml)


def gwb_spectrum(f, dur, howml, log10_A, gamma, turnover=False, f0=1e-9, beta=1, power=1, userSpec=None):
    """red_noise.py:243-265 -> ``C(f)``."""
    if userSpec is None:
        alpha = -0.5 * (gamma - 3)
        hcf = 10**log10_A * (f / F1YR_GWB) ** alpha
        if turnover:
            hcf = hcf / (1 + (f / f0) ** (power * (alpha - beta))) ** (1 / power)
    else:
        lx, ly = np.log10(userSpec[:, 0]), np.log10(userSpec[:, 1])
        lf = np.log10(f)
        hcf = 10.0 ** np.where(lf < lx[0], ly[0], np.where(lf > lx[-1], ly[-1], np.interp(lf, lx, ly)))
    return 1 / 96 / np.pi**2 * hcf**2 / f**3 * dur * howml


def gwb_from_draws(setup, C, M, w, mjds_per_psr):
    """red_noise.py:268-287: colour, Hermitian pack, IFFT, crop, interpolate."""
    npsr, Nf = w.shape
    npts, dt = setup["npts"], setup["dt"]
    Res_f = (M @ w) * np.sqrt(C)[None, :]
    Res_f[:, 0] = 0
    Res_f[:, -1] = 0
    full = np.zeros((npsr, 2 * Nf - 2), complex)
    full[:, :Nf] = Res_f
    full[:, Nf:] = np.conj(Res_f[:, Nf - 2:0:-1])
    Res_t = np.real(np.fft.ifft(full) / dt)
    grid = Res_t[:, 10:npts + 10]
    out = [np.interp(np.asarray(m, float) * 86400, setup["ut"], grid[p]) for p, m in enumerate(mjds_per_psr)]
    return out, grid


def orf_matrix(locs, names=None, lmax=0, clm=(np.sqrt(4.0 * np.pi),), no_correlations=False):
    """red_noise.py:200-226 for RAJ/DECJ positions (hours, deg)."""
    n = len(locs)
    if no_correlations:
        return np.diag(np.ones(n) * 2)
    pl = np.zeros((n, 2))
    for i, lc in enumerate(locs):
        pl[i] = lc["RAJ"] * np.pi / 12.0, lc["DECJ"] * np.pi / 180.0
    pl[:, 1] = np.pi / 2.0 - pl[:, 1]
    if lmax == 0:
        basis = [hd_basis_l0(pl)]
    else:
        raise NotImplementedError("l>0: compare against tests/golden/ref_orf.npz (generated from the reference)")
    return 2.0 * sum(c * b for c, b in zip(clm, basis))


def hd_basis_l0(psrlocs):
    """spharmORFbasis.py:385-434 at lmax=0: arbCompFrame_ORF(0,0,zeta) (:309-344 ->
    arbORF :164-189) -- the closed form; the l=0 rotation is the identity."""
    phi, th = psrlocs[:, 0], psrlocs[:, 1]
    n = len(phi)
    out = np.zeros((n, n))
    norm = 3.0 / (8 * np.pi)
    for a in range(n):
        for b in range(a, n):
            if phi[a] == phi[b] and th[a] == th[b]:
                val = 2.0 * norm * 0.25 * np.sqrt(np.pi * 4) * (1 + 1.0 / 3.0)
            else:
                arg = np.sin(th[a]) * np.sin(th[b]) * np.cos(phi[a] - phi[b]) + np.cos(th[a]) * np.cos(th[b])
                zeta = np.pi if arg < -1 else (0.0 if arg > 1 else np.arccos(arg))
                c = np.cos(zeta)
                if zeta == 0.0:
                    val = 2.0 * norm * 0.25 * np.sqrt(np.pi * 4) * (1 + c / 3.0)
                else:
                    # Fminus00(0,0,0,z) (:43-67) = 2 - (1+c)
                    # Fplus01(1,0,0,z) (:97-134) = -(2 - (1-c)) + 2 log(2/(1-c))
                    fm = 2.0 - (1.0 + c)
                    fp = -(2.0 - (1.0 - c)) + 2.0 * np.log(2.0 / (1.0 - c))
                    val = norm * 0.5 * np.sqrt(np.pi) * (1.0 + c / 3.0 - (1.0 + c) * fm - (1.0 - c) * fp)
            out[a, b] = out[b, a] = val
    return out


# --------------------------------------------------------------------------- CGW
def cgw(mjd, loc, gwtheta, gwphi, mc, dist, fgw, phase0, psi, inc, pdist=1.0, pphase=None,
        psrTerm=True, evolve=True, phase_approx=False, tref=0):
    """deterministic.py:50-163 for RAJ/DECJ positions."""
    mc = mc * SOLAR2S
    dist = dist * MPC2S
    w0 = np.pi * fgw
    phase0 = phase0 / 2
    w053 = w0 ** (-5 / 3)
    ct, cp, st, sp = np.cos(gwtheta), np.cos(gwphi), np.sin(gwtheta), np.sin(gwphi)
    s2p, c2p = np.sin(2 * psi), np.cos(2 * psi)
    inc1, inc2 = 0.5 * (3 + np.cos(2 * inc)), 2 * np.cos(inc)
    m = np.array([sp, -cp, 0.0])
    n = np.array([-ct * cp, -ct * sp, st])
    om = np.array([-st * cp, -st * sp, -ct])
    fac1 = 256 / 5 * mc ** (5 / 3) * w0 ** (8 / 3)
    fac2 = 1 / 32 / mc ** (5 / 3)
    fac3 = mc ** (5 / 3) / dist
    ptheta = np.pi / 2 - loc["DECJ"] * np.pi / 180.0
    pphi = loc["RAJ"] * np.pi / 12.0
    phat = np.array([np.sin(ptheta) * np.cos(pphi), np.sin(ptheta) * np.sin(pphi), np.cos(ptheta)])
    fplus = 0.5 * (np.dot(m, phat) ** 2 - np.dot(n, phat) ** 2) / (1 + np.dot(om, phat))
    fcross = (np.dot(m, phat) * np.dot(n, phat)) / (1 + np.dot(om, phat))
    cosMu = -np.dot(om, phat)
    toas = np.asarray(mjd, float) * 86400 - tref
    pd = pphase / (2 * np.pi * fgw * (1 - cosMu)) / KPC2S if pphase is not None else pdist
    pd = pd * KPC2S
    tp = toas - pd * (1 - cosMu)
    if evolve:
        omega = w0 * (1 - fac1 * toas) ** (-3 / 8)
        omega_p = w0 * (1 - fac1 * tp) ** (-3 / 8)
        phase = phase0 + fac2 * (w053 - omega ** (-5 / 3))
        phase_p = phase0 + fac2 * (w053 - omega_p ** (-5 / 3))
    elif phase_approx:
        omega = w0
        omega_p = w0 * (1 + fac1 * pd * (1 - cosMu)) ** (-3 / 8)
        phase = phase0 + omega * toas
        phase_p = phase0 + fac2 * (w053 - omega_p ** (-5 / 3)) + omega_p * toas
    else:
        omega = w0
        omega_p = omega
        phase = phase0 + omega * toas
        phase_p = phase0 + omega * tp
    At, Bt = np.sin(2 * phase) * inc1, np.cos(2 * phase) * inc2
    Atp, Btp = np.sin(2 * phase_p) * inc1, np.cos(2 * phase_p) * inc2
    alpha, alpha_p = fac3 / omega ** (1 / 3), fac3 / omega_p ** (1 / 3)
    rplus = alpha * (At * c2p + Bt * s2p)
    rcross = alpha * (-At * s2p + Bt * c2p)
    rplus_p = alpha_p * (Atp * c2p + Btp * s2p)
    rcross_p = alpha_p * (-Atp * s2p + Btp * c2p)
    if psrTerm:
        return fplus * (rplus_p - rplus) + fcross * (rcross_p - rcross)
    return -fplus * rplus - fcross * rcross


def cw_catalog(mjd, loc, cat, **kw):
    """deterministic.py:443-561 (``loop_over_CWs``): per-source ``cgw`` with NaNs dropped (:553-559), summed."""
    out = np.zeros(len(mjd))
    with np.errstate(invalid="ignore"):
        for k in range(len(cat["mc"])):
            r = cgw(mjd, loc, cat["gwtheta"][k], cat["gwphi"][k], cat["mc"][k], cat["dist"][k], cat["fgw"][k],
                    cat["phase0"][k], cat["psi"][k], cat["inc"][k], **kw)
            out += np.where(np.isnan(r), 0.0, r)
    return out


# --------------------------------------------------------------------------- bursts, transients, memory
def antenna(loc, gwtheta, gwphi):
    """(fplus, fcross) as in deterministic.py:733-759 / :837-862 for RAJ/DECJ positions."""
    ct, cp, st, sp = np.cos(gwtheta), np.cos(gwphi), np.sin(gwtheta), np.sin(gwphi)
    m = np.array([sp, -cp, 0.0])
    n = np.array([-ct * cp, -ct * sp, st])
    om = np.array([-st * cp, -st * sp, -ct])
    ptheta = np.pi / 2 - loc["DECJ"] * np.pi / 180.0
    pphi = loc["RAJ"] * np.pi / 12.0
    phat = np.array([np.sin(ptheta) * np.cos(pphi), np.sin(ptheta) * np.sin(pphi), np.cos(ptheta)])
    fplus = 0.5 * (np.dot(m, phat) ** 2 - np.dot(n, phat) ** 2) / (1 + np.dot(om, phat))
    fcross = (np.dot(m, phat) * np.dot(n, phat)) / (1 + np.dot(om, phat))
    return fplus, fcross


def burst(mjd, loc, gwtheta, gwphi, waveform_plus, waveform_cross, psi=0.0, tref=0, remove_quad=False):
    """deterministic.py:761-780: elliptically polarised burst, optional quadratic removal (np.polyfit)."""
    fplus, fcross = antenna(loc, gwtheta, gwphi)
    toas = np.asarray(mjd, float) * 86400 - tref
    hplus, hcross = waveform_plus(toas), waveform_cross(toas)
    rplus = hplus * np.cos(2 * psi) - hcross * np.sin(2 * psi)
    rcross = hplus * np.sin(2 * psi) + hcross * np.cos(2 * psi)
    res = -fplus * rplus - fcross * rcross
    if remove_quad:
        pp = np.polyfit(toas, res, 2)
        res = res - pp[0] * toas ** 2 - pp[1] * toas - pp[2]
    return res


def noise_transient(mjd, waveform, tref=0):
    """deterministic.py:805-810."""
    return waveform(np.asarray(mjd, float) * 86400 - tref)


def gw_memory(mjd, loc, strain, gwtheta, gwphi, bwm_pol, t0_mjd):
    """deterministic.py:864-873: ramp pol * strain * (t - t0) after the burst epoch."""
    fplus, fcross = antenna(loc, gwtheta, gwphi)
    pol = np.cos(2 * bwm_pol) * fplus + np.sin(2 * bwm_pol) * fcross
    toas = np.asarray(mjd, float) * 86400
    t0 = t0_mjd * 86400
    return np.where(toas < t0, 0.0, pol * strain * (toas - t0))


# --------------------------------------------------------------------------- whole recipes
def weighted_mean_residual(delay, err):
    """residual formation without PINT."""
    w = 1.0 / np.square(err)
    return delay - np.sum(w * delay) / np.sum(w)
