"""TEST INFRASTRUCTURE ONLY -- run the UNMODIFIED reference without PINT/astropy.

The reference modules import astropy / pint / enterprise / ephem / holodeck at
module top (``/root/reference/pta_replicator/simulate.py:10-20``,
``white_noise.py:3-4``, ``red_noise.py:4-6``, ``deterministic.py:3-8``).  None of
those is installed here, and none of their arithmetic is on the hot path: they
carry units and apply ``dt`` to the TOAs.  This module installs ~100 lines of
``sys.modules`` stubs and offers a duck-typed pulsar so the five hot functions
run byte-for-byte as shipped.  The package is imported from ``/root/reference``
where that exists (the authoring container) and otherwise from ``oracle/_ref``,
the byte-compiled copy staged by ``oracle/build_ref.py`` (git-ignored; it travels
to the GPU box like a built ``.so``), so the CPU arm of ``bench.py`` times the
reference's own functions there.
"""
from __future__ import annotations

import importlib
import sys
import types

import numpy as np

import os

REFERENCE_ROOT = os.environ.get("PTAR_REFERENCE_ROOT", "/root/reference")   # the override lets a test exercise the staged copy
STAGED_ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")


def available() -> str:
    """Where the unmodified reference can be imported from ('' if nowhere)."""
    if os.path.isdir(os.path.join(REFERENCE_ROOT, "pta_replicator")):
        return REFERENCE_ROOT
    if os.path.isfile(os.path.join(STAGED_ROOT, "pta_replicator", "white_noise.pyc")):
        return STAGED_ROOT
    return ""

_UNIT_SECONDS = {"s": 1.0, "day": 86400.0, "us": 1e-6, "MHz": 1.0}


class _Unit:
    __array_ufunc__ = None

    def __init__(self, name):
        self.name = name

    def __rmul__(self, other):
        return _Quantity(np.asarray(other, dtype=float), self.name)

    def __mul__(self, other):
        return _Quantity(np.asarray(other, dtype=float), self.name)

    def __repr__(self):
        return f"Unit({self.name})"


def _unit_name(u):
    return u.name if isinstance(u, _Unit) else str(u)


class _Quantity:
    """Minimal astropy.units.Quantity stand-in (time units only)."""

    __array_ufunc__ = None

    def __init__(self, value, unit):
        self.value = np.asarray(value, dtype=float)
        self.unit = _unit_name(unit)

    def to(self, unit):
        unit = _unit_name(unit)
        fac = _UNIT_SECONDS[self.unit] / _UNIT_SECONDS[unit]
        return _Quantity(self.value * fac, unit)

    def to_value(self, unit):
        return self.to(unit).value

    def _other(self, other):
        if isinstance(other, _Quantity):
            return other.to(self.unit).value
        return np.asarray(other, dtype=float)

    def __mul__(self, other):
        if isinstance(other, _Unit):
            raise TypeError("unit*unit not needed")
        return _Quantity(self.value * np.asarray(other, dtype=float), self.unit)

    __rmul__ = __mul__

    def __truediv__(self, other):
        return _Quantity(self.value / np.asarray(other, dtype=float), self.unit)

    def __add__(self, other):
        return _Quantity(self.value + self._other(other), self.unit)

    __radd__ = __add__

    def __iadd__(self, other):
        self.value = self.value + self._other(other)
        return self

    def __neg__(self):
        return _Quantity(-self.value, self.unit)

    def __len__(self):
        return len(self.value)

    def __getitem__(self, k):
        return _Quantity(self.value[k], self.unit)


class _TimeDelta:
    def __init__(self, q, format=None):
        self.sec = q.to("s").value if isinstance(q, _Quantity) else np.asarray(q, float) * 86400.0


def _module(name, **attrs):
    m = types.ModuleType(name)
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m
    return m


# ---- stand-ins for the four holodeck functions / two astropy constants used by add_gwb_plus_outlier_cws
# (deterministic.py:614-631).  holodeck is not installed: these restate the published definitions (the strain
# formula is the one quoted at deterministic.py:636-637; flat LambdaCDM with holodeck's default WMAP9 parameters)
# so that the UNMODIFIED reference function can run and pin the partition / draw-order / injection logic.
_H_G, _H_C = 6.6743e-8, 2.99792458e10
_H_MSOL, _H_PC = 1.988409870698051e33, 3.0856775814913674e18
_H_H0, _H_OM = 69.32, 0.2865


def _h_m1m2_from_mtmr(mt, mr):
    mt, mr = np.asarray(mt, dtype=float), np.asarray(mr, dtype=float)
    m1 = mt / (1.0 + mr)
    return m1, mt - m1


def _h_chirp_mass(m1, m2):
    return np.power(m1 * m2, 3.0 / 5.0) / np.power(m1 + m2, 1.0 / 5.0)


def _h_gw_strain_source(mchirp, dcom, freq_rest_orb):
    return (8.0 / np.sqrt(10.0)) * np.power(_H_G * mchirp, 5.0 / 3.0) * np.power(2.0 * np.pi * freq_rest_orb, 2.0 / 3.0) / (
        _H_C ** 4 * dcom)


def _h_z_to_dcom(z):
    x, w = np.polynomial.legendre.leggauss(64)
    z = np.atleast_1d(np.asarray(z, dtype=float))
    h0 = _H_H0 * 1.0e5 / (1.0e6 * _H_PC)
    zz = 0.5 * z[:, None] * (x[None, :] + 1.0)
    integrand = 1.0 / np.sqrt(_H_OM * (1.0 + zz) ** 3 + (1.0 - _H_OM))
    return (_H_C / h0) * 0.5 * z * np.sum(w[None, :] * integrand, axis=1)


def install():
    """Install the stubs and put the reference on ``sys.path``.  Idempotent."""
    if "pta_replicator" in sys.modules and getattr(sys.modules["pta_replicator"], "_ptar_stubbed", False):
        return
    units = _module("astropy.units", s=_Unit("s"), day=_Unit("day"), us=_Unit("us"), MHz=_Unit("MHz"),
                    Quantity=_Quantity)
    time = _module("astropy.time", TimeDelta=_TimeDelta)
    _module("astropy", units=units, time=time)

    class _Dummy:  # annotation targets for simulate.py:29-31
        pass

    residuals = _module("pint.residuals", Residuals=_Dummy)
    toa = _module("pint.toa", TOAs=_Dummy, get_TOAs=None)
    models = _module("pint.models", TimingModel=_Dummy, get_model=None)
    simulation = _module("pint.simulation", make_fake_toas_fromMJDs=None)
    fitter = _module("pint.fitter")
    _module("pint", residuals=residuals, toa=toa, models=models, simulation=simulation, fitter=fitter)
    epulsar = _module("enterprise.pulsar", Pulsar=_Dummy)
    _module("enterprise", pulsar=epulsar)
    _module("ephem")
    hutils = _module("holodeck.utils", m1m2_from_mtmr=_h_m1m2_from_mtmr, chirp_mass=_h_chirp_mass,
                     gw_strain_source=_h_gw_strain_source)
    hcosmo = _module("holodeck.cosmo", z_to_dcom=_h_z_to_dcom)
    _module("holodeck", utils=hutils, cosmo=hcosmo)
    cgs = lambda v: types.SimpleNamespace(cgs=types.SimpleNamespace(value=v))  # noqa: E731
    sys.modules["astropy"].constants = _module("astropy.constants", pc=cgs(_H_PC), M_sun=cgs(_H_MSOL))
    root = available()
    if not root:
        raise ImportError("the reference is neither at /root/reference nor staged under oracle/_ref "
                          "(python -m oracle.build_ref in the authoring container)")
    if root not in sys.path:
        sys.path.insert(0, root)
    pkg = importlib.import_module("pta_replicator")
    pkg._ptar_stubbed = True


class _Col:
    def __init__(self, data):
        self.data = data


class StubTOAs:
    """Duck-typed pint.toa.TOAs (fields the hot functions touch)."""

    def __init__(self, mjd_ld, err_us, flags):
        self.mjd0 = np.asarray(mjd_ld, dtype=np.longdouble)
        self.ntoas = len(self.mjd0)
        self.err_us = np.asarray(err_us, dtype=float)
        self.flags = list(flags)
        self.delta = np.zeros(self.ntoas)  # accumulated seconds
        self.table = {"tdbld": self.mjd0.copy(), "flags": _Col(self.flags)}

    def get_mjds(self):
        return _Quantity(np.asarray(self.table["tdbld"], dtype=float), "day")

    def get_errors(self):
        return _Quantity(self.err_us, "us")

    first_override = None  # float MJD; lets a test pin add_gwb's Nf
    last_override = None

    @property
    def first_MJD(self):
        v = self.first_override if self.first_override is not None else float(np.min(self.table["tdbld"]))
        return types.SimpleNamespace(value=v)

    @property
    def last_MJD(self):
        v = self.last_override if self.last_override is not None else float(np.max(self.table["tdbld"]))
        return types.SimpleNamespace(value=v)

    def adjust_TOAs(self, td):
        self.delta = self.delta + td.sec
        self.table["tdbld"] = self.table["tdbld"] + np.asarray(td.sec, dtype=np.longdouble) / np.longdouble(86400)


class StubPulsar:
    """Duck-typed SimulatedPulsar with the ledger semantics of simulate.py:79-89."""

    def __init__(self, name, loc, mjd_ld, err_us, flags, freeze_toas=True):
        self.name = name
        self.loc = dict(loc)
        self.toas = StubTOAs(mjd_ld, err_us, flags)
        self.added_signals = {}
        self.added_signals_time = {}
        self._freeze = freeze_toas

    def update_added_signals(self, signal_name, param_dict, dt=None):
        if self.added_signals is None:
            raise ValueError("make_ideal() must be called on SimulatedPulsar before adding new signals.")
        if signal_name in self.added_signals:
            raise ValueError(f"{signal_name} already exists in the model.")
        self.added_signals[signal_name] = param_dict
        if dt is not None:
            self.added_signals_time[signal_name] = dt

    def update_residuals(self):
        if self._freeze:  # keep the TOAs ideal so each signal is evaluated at the same epochs
            self.toas.table["tdbld"] = self.toas.mjd0.copy()

    def signal_seconds(self, name):
        return np.asarray(self.added_signals_time[name].to("s").value, dtype=float)

    @property
    def resids_value(self):
        d = self.toas.delta
        return d - d.mean()


def reference_modules():
    install()
    wn = importlib.import_module("pta_replicator.white_noise")
    rn = importlib.import_module("pta_replicator.red_noise")
    det = importlib.import_module("pta_replicator.deterministic")
    orf = importlib.import_module("pta_replicator.spharmORFbasis")
    const = importlib.import_module("pta_replicator.constants")
    return types.SimpleNamespace(white_noise=wn, red_noise=rn, deterministic=det, orf=orf, constants=const)
