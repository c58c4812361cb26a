"""Anisotropic ORF basis for lmax > 0, vectorised over pulsar pairs.

Restates ``/root/reference/pta_replicator/spharmORFbasis.py``: the computational-frame ORFs of Gair
et al. (2014) (``Fminus00`` :43-67, ``Fminus01`` :70-94, ``Fplus01`` :97-134, ``Fplus00`` :137-161,
``arbORF`` :164-248, ``arbCompFrame_ORF`` :309-344), the rotation to the cosmic frame with Wigner
D-matrices (``dlmk`` :251-268, ``Dlmk`` :271-279, ``gamma`` :282-306, ``rotated_Gamma_ml`` :347-359)
and the real-harmonic combinations of Mingarelli et al. (2013) (``real_rotated_Gammas`` :362-382),
assembled like ``correlated_basis`` :385-434.  The reference loops over pairs and (l, m) in scalar
Python (~160 s for 67 pulsars at lmax = 6); here every formula is evaluated once per (l, m) on the
array of all pairs, with the double sums kept in the reference's order so the rounding matches.
Setup only (realization independent); the result feeds ``ptar_cholesky_lower``.
"""
from __future__ import annotations

from math import factorial as _fact

import numpy as np

NORM = 3.0 / (8 * np.pi)


def _f(n):
    return float(_fact(n))


def _Fminus00(qq, mm, ll, c):
    out = np.zeros_like(c)
    for ii in range(0, qq + 1):
        for jj in range(mm, ll + 1):
            p = qq - ii + jj - mm + 1
            out = out + ((2.0 ** (ii - jj) * (-1.0) ** (qq - ii + jj + mm))
                         * (_f(qq) * _f(ll + jj) * (2.0 ** p - (1.0 + c) ** p))
                         / (_f(ii) * _f(qq - ii) * _f(jj) * _f(ll - jj) * _f(jj - mm) * p))
    return out


def _Fminus01(qq, mm, ll, c):
    out = np.zeros_like(c)
    for ii in range(0, qq + 1):
        for jj in range(mm, ll + 1):
            p = qq - ii + jj - mm + 2
            out = out + ((2.0 ** (ii - jj) * (-1.0) ** (qq - ii + jj + mm))
                         * (_f(qq) * _f(ll + jj) * (2.0 ** p - (1.0 + c) ** p))
                         / (_f(ii) * _f(qq - ii) * _f(jj) * _f(ll - jj) * _f(jj - mm) * p))
    return out


def _Fplus01(qq, mm, ll, c):
    out = np.zeros_like(c)
    for ii in range(0, qq):
        for jj in range(mm, ll + 1):
            p = qq - ii + jj - mm
            out = out + ((2.0 ** (ii - jj) * (-1.0) ** (ll + qq - ii + jj))
                         * (_f(qq) * _f(ll + jj) * (2.0 ** p - (1.0 - c) ** p))
                         / (_f(ii) * _f(qq - ii) * _f(jj) * _f(ll - jj) * _f(jj - mm) * p))
    if mm != ll:
        for jj in range(mm + 1, ll + 1):
            out = out + ((2.0 ** (qq - jj) * (-1.0) ** (ll + jj))
                         * (_f(ll + jj) * (2.0 ** (jj - mm) - (1.0 - c) ** (jj - mm)))
                         / (_f(jj) * _f(ll - jj) * _f(jj - mm) * (jj - mm)))
    out = out + ((-1.0) ** (ll + mm) * 2.0 ** (qq - mm) * _f(ll + mm) * np.log(2.0 / (1.0 - c))) / (
        1.0 * _f(mm) * _f(ll - mm))
    return out


def _Fplus00(qq, mm, ll, c):
    out = np.zeros_like(c)
    for ii in range(0, qq + 1):
        for jj in range(mm, ll + 1):
            p = qq - ii + jj - mm + 1
            out = out + ((2.0 ** (ii - jj) * (-1.0) ** (ll + qq - ii + jj))
                         * (_f(qq) * _f(ll + jj) * (2.0 ** p - (1.0 - c) ** p))
                         / (_f(ii) * _f(qq - ii) * _f(jj) * _f(ll - jj) * _f(jj - mm) * p))
    return out


def _arbORF(mm, ll, zeta):
    """``arbORF`` for zeta != 0 (the zeta == 0 branch is handled by ``_comp_frame``)."""
    c = np.cos(zeta)
    pre = np.sqrt((2.0 * ll + 1.0) * np.pi)
    if mm == 0:
        body = -(1.0 + c) * _Fminus00(0, 0, ll, c) - (1.0 - c) * _Fplus01(1, 0, ll, c)
        if 0 <= ll <= 2:
            delta = [1.0 + c / 3.0, -(1.0 + c) / 3.0, 2.0 * c / 15.0][ll]
            # reference order: delta - (1+c) Fm - (1-c) Fp
            body = delta - (1.0 + c) * _Fminus00(0, 0, ll, c) - (1.0 - c) * _Fplus01(1, 0, ll, c)
        return NORM * 0.5 * pre * body
    if mm == 1:
        t1 = ((1.0 + c) ** (3.0 / 2.0) / (1.0 - c) ** (1.0 / 2.0)) * _Fminus00(1, 1, ll, c)
        t2 = ((1.0 - c) ** (3.0 / 2.0) / (1.0 + c) ** (1.0 / 2.0)) * _Fplus01(2, 1, ll, c)
        rt = np.sqrt((1.0 * _f(ll - 1)) / (1.0 * _f(ll + 1)))
        if ll in (1, 2):
            delta = [2.0 * np.sin(zeta) / 3.0, -2.0 * np.sin(zeta) / 5.0][ll - 1]
            return NORM * 0.25 * pre * rt * (delta - t1 - t2)
        return NORM * 0.25 * pre * rt * (-t1 - t2)
    h = mm / 2.0
    rt = np.sqrt((1.0 * _f(ll - mm)) / (1.0 * _f(ll + mm)))
    return (-NORM * 0.25 * pre * rt
            * (((1.0 + c) ** (h + 1) / (1.0 - c) ** h) * _Fminus00(mm, mm, ll, c)
               - ((1.0 + c) ** h / (1.0 - c) ** (h - 1.0)) * _Fminus01(mm - 1, mm, ll, c)
               + ((1.0 - c) ** (h + 1) / (1.0 + c) ** h) * _Fplus01(mm + 1, mm, ll, c)
               - ((1.0 - c) ** h / (1.0 + c) ** (h - 1.0)) * _Fplus00(mm, mm, ll, c)))


def _comp_frame(mm, ll, zeta):
    """``arbCompFrame_ORF`` on an array of separations."""
    z0 = zeta == 0.0
    zpi = zeta == np.pi
    safe = np.where(z0, 1.0, zeta)  # any non-special value; masked below
    with np.errstate(all="ignore"):
        gen = _arbORF(mm, ll, safe)
    out = gen
    # zeta == pi
    if ll > 2 or (ll in (1, 2) and mm != 0):
        out = np.where(zpi, 0.0, out)
    # zeta == 0 (pulsar-term doubling on the diagonal and for coincident pulsars)
    c0 = 1.0  # cos(0)
    if ll > 2:
        v0 = 0.0
    elif ll == 2:
        v0 = 2 * 0.25 * NORM * (4.0 / 3) * (np.sqrt(np.pi / 5)) * c0 if mm == 0 else 0.0
    elif ll == 1:
        v0 = -2 * 0.5 * NORM * (np.sqrt(np.pi / 3.0)) * (1.0 + c0) if mm == 0 else 0.0
    else:
        v0 = 2.0 * NORM * 0.25 * np.sqrt(np.pi * 4) * (1 + (c0 / 3.0))
    return np.where(z0, v0, out)


def _dlmk(l, m, k, theta1):
    """Wigner small-d as in Allen & Ottewill 97 (``dlmk`` :251-268)."""
    from scipy import special as sp
    if m >= k:
        factor = np.sqrt(_f(l - k) * _f(l + m) / _f(l + k) / _f(l - m))
        part2 = (np.cos(theta1 / 2)) ** (2 * l + k - m) * (-np.sin(theta1 / 2)) ** (m - k) / _f(m - k)
        part3 = sp.hyp2f1(m - l, -k - l, m - k + 1, -((np.tan(theta1 / 2)) ** 2))
        return factor * part2 * part3
    return (-1) ** (m - k) * _dlmk(l, k, m, theta1)


def _third_angle(phi1, phi2, theta1, theta2):
    """``gamma`` :282-306."""
    same = (phi1 == phi2) & (theta1 == theta2)
    with np.errstate(all="ignore"):
        g = np.arctan(np.sin(theta2) * np.sin(phi2 - phi1)
                      / (np.cos(theta1) * np.sin(theta2) * np.cos(phi1 - phi2) - np.sin(theta1) * np.cos(theta2)))
    g = np.where(same, 0.0, g)
    dummy = (np.cos(g) * np.cos(theta1) * np.sin(theta2) * np.cos(phi1 - phi2)
             + np.sin(g) * np.sin(theta2) * np.sin(phi2 - phi1) - np.cos(g) * np.sin(theta1) * np.cos(theta2))
    return np.where(dummy >= 0, g, np.pi + g)


def correlated_basis(psrlocs, lmax):
    """List of (lmax+1)^2 matrices ordered (l, m = -l..l); ``psrlocs[:, 0]`` azimuth, ``[:, 1]`` polar angle."""
    psrlocs = np.asarray(psrlocs, dtype=float)
    n = len(psrlocs)
    ia, ib = np.triu_indices(n)
    phi1, phi2 = psrlocs[ia, 0], psrlocs[ib, 0]
    th1, th2 = psrlocs[ia, 1], psrlocs[ib, 1]
    same = (phi1 == phi2) & (th1 == th2)
    arg = np.sin(th1) * np.sin(th2) * np.cos(phi1 - phi2) + np.cos(th1) * np.cos(th2)
    zeta = np.where(arg < -1, np.pi, np.where(arg > 1, 0.0, np.arccos(np.clip(arg, -1.0, 1.0))))
    zeta = np.where(same, 0.0, zeta)
    gam = _third_angle(phi1, phi2, th1, th2)
    out = []
    for ll in range(lmax + 1):
        plus = [_comp_frame(mm, ll, zeta) for mm in range(ll + 1)]
        gamma_ml = [(-1) ** mm * plus[mm] for mm in range(ll, 0, -1)] + plus  # index k + ll, k = -ll..ll

        def rotated(m):
            acc = np.zeros(len(zeta), dtype=complex)
            for ii in range(2 * ll + 1):
                k = ii - ll
                D = np.exp(-1j * m * phi1) * _dlmk(ll, m, k, th1) * np.exp(-1j * k * gam)
                acc = acc + np.conj(D) * gamma_ml[ii]
            return acc

        cache = {m: rotated(m) for m in range(-ll, ll + 1)}
        for m in range(-ll, ll + 1):
            if m > 0:
                val = ((1.0 / np.sqrt(2)) * (cache[m] + (-1) ** m * cache[-m])).real
            elif m == 0:
                val = cache[0].real
            else:
                val = ((1.0 / np.sqrt(2) / complex(0.0, 1)) * (cache[-m] - (-1) ** m * cache[m])).real
            mat = np.zeros((n, n))
            mat[ia, ib] = val
            mat[ib, ia] = val
            out.append(mat)
    return out
