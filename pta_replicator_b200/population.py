"""Holodeck-free helpers for ``add_gwb_plus_outlier_cws`` (``/root/reference/pta_replicator/deterministic.py:565-715``).

The reference takes these four functions and two constants from ``holodeck`` (``utils.m1m2_from_mtmr``,
``utils.chirp_mass``, ``utils.gw_strain_source``, ``cosmo.z_to_dcom``) and ``astropy.constants``; neither package is
available here, so they are restated from their published definitions -- **parity unpinned** against holodeck itself
(the partition / injection logic built on top of them IS pinned to the unmodified reference, tests/golden/ref_outliers.npz).
All quantities in cgs, like holodeck.
"""
from __future__ import annotations

import numpy as np

NWTG = 6.6743e-8                 # G [cm^3 g^-1 s^-2]  (CODATA 2018)
SPLC = 2.99792458e10             # c [cm/s]
MSOL = 1.988409870698051e33      # solar mass [g]   (astropy.constants.M_sun.cgs)
PC = 3.0856775814913674e18       # parsec [cm]      (astropy.constants.pc.cgs)
# flat LambdaCDM of holodeck's default cosmology (WMAP9)
H0_KMS_MPC = 69.32
OMEGA_M = 0.2865


def m1m2_from_mtmr(mt, mr):
    """Component masses from total mass and mass ratio q = m2/m1 <= 1."""
    mt, mr = np.asarray(mt, dtype=float), np.asarray(mr, dtype=float)
    m1 = mt / (1.0 + mr)
    return m1, mt - m1


def chirp_mass(m1, m2):
    return np.power(m1 * m2, 3.0 / 5.0) / np.power(m1 + m2, 1.0 / 5.0)


def gw_strain_source(mchirp, dcom, freq_rest_orb):
    """Sky- and polarisation-averaged source strain of a circular binary,
    h_s = 8/sqrt(10) (G Mc)^(5/3) (2 pi f_orb)^(2/3) / (c^4 d_com)  (the formula quoted at deterministic.py:636-637)."""
    return (8.0 / np.sqrt(10.0)) * np.power(NWTG * mchirp, 5.0 / 3.0) * np.power(2.0 * np.pi * freq_rest_orb, 2.0 / 3.0) / (
        SPLC ** 4 * dcom)


_GL_X, _GL_W = np.polynomial.legendre.leggauss(64)


def z_to_dcom(z):
    """Comoving distance [cm] in flat LambdaCDM: (c/H0) int_0^z dz' / sqrt(Om (1+z')^3 + 1 - Om), 64-point
    Gauss-Legendre per redshift (relative error < 1e-12 for z < 10)."""
    z = np.atleast_1d(np.asarray(z, dtype=float))
    h0 = H0_KMS_MPC * 1.0e5 / (1.0e6 * PC)          # 1/s
    zz = 0.5 * z[:, None] * (_GL_X[None, :] + 1.0)
    integrand = 1.0 / np.sqrt(OMEGA_M * (1.0 + zz) ** 3 + (1.0 - OMEGA_M))
    return (SPLC / h0) * 0.5 * z * np.sum(_GL_W[None, :] * integrand, axis=1)


def partition_population(vals, weights, fobs, T_obs, outlier_per_bin=100):
    """The realization-independent part of ``add_gwb_plus_outlier_cws`` (deterministic.py:616-669, :685-689): per
    frequency bin keep the ``outlier_per_bin`` binaries with the largest weighted h_c^2 as individual sources and sum
    the rest into a free spectrum.  Returns (f_centers, free_spec, outlier_fo, outlier_hs, outlier_mc [Msun],
    outlier_dl [Mpc]) with empty slots dropped like the reference does."""
    vals = np.asarray(vals, dtype=float)
    weights = np.asarray(weights, dtype=float)
    fobs = np.asarray(fobs, dtype=float)
    f_centers = np.array([(fobs[i + 1] + fobs[i]) / 2 for i in range(fobs.size - 1)])
    mc = chirp_mass(*m1m2_from_mtmr(vals[0], vals[1]))          # rest frame
    rz = vals[2, :]
    frst = vals[3] * (1.0 + rz)
    dc = z_to_dcom(rz)
    dl = np.copy(dc) * (1.0 + rz)
    hs = gw_strain_source(mc, dc, frst / 2)
    fo = vals[-1]
    mc = mc * (1.0 + rz)                                        # observer frame for the injections
    bin_of = np.digitize(fo, fobs) - 1
    nb = fobs.shape[0] - 1
    free_spec = np.ones(nb) * 1e-100
    slots = nb * outlier_per_bin
    o_hs, o_fo, o_mc, o_dl = np.zeros(slots), np.zeros(slots), np.zeros(slots), np.zeros(slots)
    whs = weights * hs ** 2 * fo * T_obs
    for k in range(nb):
        members = np.flatnonzero(bin_of == k)
        order = members[np.argsort(whs[members])[::-1]]         # loudest first
        top = order[:outlier_per_bin]
        s = slice(outlier_per_bin * k, outlier_per_bin * k + len(top))
        o_hs[s], o_fo[s] = whs[top], fo[top]
        o_mc[s], o_dl[s] = mc[top] / MSOL, dl[top] / PC / 1e6
        free_spec[k] += np.sum(whs[order[outlier_per_bin:]])
    return f_centers, free_spec, o_fo[o_fo > 0], o_hs[o_hs > 0], o_mc[o_mc > 0], o_dl[o_dl > 0]
