"""Overlap-reduction-function basis in real spherical harmonics (setup; realization-independent).

Restates ``/root/reference/pta_replicator/spharmORFbasis.py`` (``correlated_basis`` :385-434 and
its helpers :14-382; Gair et al. 2014, Mingarelli et al. 2013) vectorised over pulsar pairs: the
reference evaluates every pair and every (l, m) in scalar Python (about 160 s for 67 pulsars at
lmax = 6).  ``lmax = 0`` is the Hellings-Downs curve.  The matrices feed
``ptar_cholesky_lower``.
"""
from __future__ import annotations

import math

import numpy as np

NORM = 3.0 / (8 * np.pi)


def ecliptic_to_equatorial(elong_deg, elat_deg, epoch="2000"):
    """(RA, DEC) in radians from ecliptic (ELONG, ELAT) in degrees.

    The reference calls ``ephem.Equatorial(ephem.Ecliptic(lon, lat), epoch=...)`` with epoch 1950
    for B-names (red_noise.py:210-221).  PyEphem is not available here; this is the standard
    rotation by the mean obliquity at J2000 followed, for epoch 1950, by IAU-1976 precession.
    PARITY UNPINNED (no reference fixture uses ELONG/ELAT).
    """
    lam, bet = np.radians(elong_deg), np.radians(elat_deg)
    eps = np.radians(23.4392911)
    x = np.cos(bet) * np.cos(lam)
    y = np.cos(bet) * np.sin(lam) * np.cos(eps) - np.sin(bet) * np.sin(eps)
    z = np.cos(bet) * np.sin(lam) * np.sin(eps) + np.sin(bet) * np.cos(eps)
    v = np.array([x, y, z])
    if str(epoch) == "1950":
        T = (2433282.4235 - 2451545.0) / 36525.0  # B1950.0 relative to J2000, Julian centuries
        asec = np.pi / 180 / 3600
        zeta = (2306.2181 * T + 0.30188 * T**2 + 0.017998 * T**3) * asec
        zz = (2306.2181 * T + 1.09468 * T**2 + 0.018203 * T**3) * asec
        th = (2004.3109 * T - 0.42665 * T**2 - 0.041833 * T**3) * asec

        def rz(a):
            return np.array([[np.cos(a), np.sin(a), 0], [-np.sin(a), np.cos(a), 0], [0, 0, 1]])

        def ry(a):
            return np.array([[np.cos(a), 0, -np.sin(a)], [0, 1, 0], [np.sin(a), 0, np.cos(a)]])

        v = rz(-zz) @ ry(th) @ rz(-zeta) @ v
    ra = math.atan2(v[1], v[0]) % (2 * np.pi)
    dec = math.asin(max(-1.0, min(1.0, v[2])))
    return ra, dec


def psrlocs_from_pulsars(psrs) -> np.ndarray:
    """[n_psr, 2] = (RA [rad], DEC [rad]) as ``add_gwb`` / ``add_cgw`` derive them
    (red_noise.py:205-221, deterministic.py:76-88): RAJ in hours, DECJ in degrees."""
    out = np.zeros((len(psrs), 2))
    for i, p in enumerate(psrs):
        if "DECJ" in p.loc:
            out[i] = float(p.loc["RAJ"] * np.pi / 12.0), float(p.loc["DECJ"] * np.pi / 180.0)
        elif "ELAT" in p.loc:
            out[i] = ecliptic_to_equatorial(p.loc["ELONG"], p.loc["ELAT"], "1950" if "B" in p.name else "2000")
        else:
            raise AttributeError("No pulsar location information (RAJ/DECJ or ELONG/ELAT).")
    return out


def _zeta(phi, theta):
    """Pairwise angular separation with the reference's clamping (spharmORFbasis.py:14-35)."""
    same = (phi[:, None] == phi[None, :]) & (theta[:, None] == theta[None, :])
    arg = (np.sin(theta)[:, None] * np.sin(theta)[None, :] * np.cos(phi[:, None] - phi[None, :])
           + np.cos(theta)[:, None] * np.cos(theta)[None, :])
    z = np.arccos(np.clip(arg, -1.0, 1.0))
    z = np.where(arg < -1, np.pi, np.where(arg > 1, 0.0, z))
    return np.where(same, 0.0, z)


def hellings_downs_l0(psrlocs) -> np.ndarray:
    """The (l, m) = (0, 0) basis matrix: ``arbCompFrame_ORF(0, 0, zeta)`` (:309-344), which for
    zeta != 0 is ``arbORF`` (:164-189) with ``Fminus00(0,0,0)`` = 2 - (1+c) and ``Fplus01(1,0,0)`` =
    -(2 - (1-c)) + 2 log(2/(1-c)); the l = 0 rotation to the cosmic frame is the identity."""
    phi, theta = np.asarray(psrlocs[:, 0], float), np.asarray(psrlocs[:, 1], float)
    zeta = _zeta(phi, theta)
    c = np.cos(zeta)
    with np.errstate(divide="ignore", invalid="ignore"):
        fm = 2.0 - (1.0 + c)
        fp = -(2.0 - (1.0 - c)) + 2.0 * np.log(2.0 / (1.0 - c))
        off = NORM * 0.5 * np.sqrt(np.pi) * (1.0 + c / 3.0 - (1.0 + c) * fm - (1.0 - c) * fp)
    on = 2.0 * NORM * 0.25 * np.sqrt(np.pi * 4) * (1 + c / 3.0)
    return np.where(zeta == 0.0, on, off)


def correlated_basis(psrlocs, lmax: int):
    """List of (lmax+1)^2 real-harmonic ORF basis matrices, ordered (l, m = -l..l) like the
    reference.  ``psrlocs[:, 0]`` = azimuth (RA), ``psrlocs[:, 1]`` = polar angle (pi/2 - DEC)."""
    psrlocs = np.asarray(psrlocs, dtype=float)
    if lmax == 0:
        return [hellings_downs_l0(psrlocs)]
    from . import orf_aniso
    return orf_aniso.correlated_basis(psrlocs, lmax)
