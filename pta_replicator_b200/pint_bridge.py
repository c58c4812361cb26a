"""PINT / enterprise bridge: push device-generated delays back into real PINT ``TOAs``.

The reference hands every injected delay to PINT (``psr.toas.adjust_TOAs(TimeDelta(dt))`` followed by
``psr.update_residuals()``, white_noise.py:124-125, red_noise.py:132-133, deterministic.py:169-170) and writes
par / tim files through PINT (simulate.py:71-77).  PINT, astropy and enterprise are not part of this image, so
everything here is import-guarded: with the packages installed the functions below work on real ``pint.toa.TOAs``
/ ``pint.models.TimingModel`` objects; without them they raise ``PintUnavailable`` with the reason.

``PulsarBatch`` itself only reads the duck-typed surface of ``TOAs`` that the reference's hot functions use
(``ntoas``, ``table['tdbld']``, ``table['flags']``, ``get_mjds()``, ``get_errors()``, ``first_MJD`` / ``last_MJD``), so
a list of PINT-backed ``SimulatedPulsar`` objects from ``load_pulsar_pint`` can be batched directly.
"""
from __future__ import annotations

import os

import numpy as np


class PintUnavailable(ImportError):
    pass


def _need(*modules):
    import importlib
    out = []
    for m in modules:
        try:
            out.append(importlib.import_module(m))
        except ImportError as e:
            raise PintUnavailable(f"{m} is not installed ({e}); the PINT bridge needs pint-pulsar + astropy "
                                  "(and enterprise for to_enterprise)") from e
    return out


def have_pint() -> bool:
    """True only for the real packages (test harnesses may have put stand-in modules into ``sys.modules``)."""
    try:
        toa, models, _, _, _ = _need("pint.toa", "pint.models", "pint.residuals", "astropy.time", "astropy.units")
    except PintUnavailable:
        return False
    return callable(getattr(toa, "get_TOAs", None)) and callable(getattr(models, "get_model", None)) and \
        getattr(toa, "__file__", None) is not None


def _need_real_pint():
    if not have_pint():
        raise PintUnavailable("pint-pulsar / astropy are not installed; the PINT bridge needs them")


def load_pulsar_pint(parfile: str, timfile: str, ephem: str = "DE440"):
    """The reference's ``load_pulsar`` (simulate.py:138-167) on real PINT objects."""
    from .simulate import SimulatedPulsar
    if not os.path.isfile(parfile):
        raise FileNotFoundError("par file does not exist.")
    if not os.path.isfile(timfile):
        raise FileNotFoundError("tim file does not exist.")
    _need_real_pint()
    toa, models, residuals = _need("pint.toa", "pint.models", "pint.residuals")
    model = models.get_model(parfile)
    toas = toa.get_TOAs(timfile, ephem=ephem, planets=True)
    res = residuals.Residuals(toas, model)
    if hasattr(model, "RAJ") and hasattr(model, "DECJ"):
        loc = {"RAJ": model.RAJ.value, "DECJ": model.DECJ.value}
    elif hasattr(model, "ELONG") and hasattr(model, "ELAT"):
        loc = {"ELONG": model.ELONG.value, "ELAT": model.ELAT.value}
    else:
        raise AttributeError("No pulsar location information (RAJ/DECJ or ELONG/ELAT) in parfile.")
    return SimulatedPulsar(ephem=ephem, model=model, toas=toas, residuals=res, name=model.PSR.value, loc=loc)


def make_ideal_pint(psr, iterations: int = 2):
    """simulate.py:193-202 on PINT objects."""
    (residuals, atime) = _need("pint.residuals", "astropy.time")
    for _ in range(iterations):
        res = residuals.Residuals(psr.toas, psr.model)
        psr.toas.adjust_TOAs(atime.TimeDelta(-1.0 * res.time_resids))
    psr.added_signals = {}
    psr.added_signals_time = {}
    psr.residuals = residuals.Residuals(psr.toas, psr.model)


def apply_delay(psr, delay_s, signal_name: str = None, params: dict = None):
    """Shift the TOAs of a PINT-backed (or PINT-like) pulsar by ``delay_s`` [s, table order] exactly as every
    ``add_*`` of the reference does: ``toas.adjust_TOAs(TimeDelta(dt))``, ledger entry, ``update_residuals()``."""
    atime, units = _need("astropy.time", "astropy.units")
    dt = np.asarray(delay_s, dtype=float) * units.s
    if signal_name is not None:
        psr.update_added_signals(signal_name, params or {}, dt)
    psr.toas.adjust_TOAs(atime.TimeDelta(dt))
    psr.update_residuals()


def apply_realization(batch, psrs, row, signal_name: str = "b200_batch", params: dict = None):
    """One realization row of ``PulsarBatch.generate`` (device tensor or host array, ENGINE order) -> every pulsar's
    TOAs.  ``psrs`` are the PINT-backed pulsars the batch was built from (same order)."""
    if hasattr(row, "cpu"):
        row = row.cpu().numpy()
    for i, p in enumerate(psrs):
        apply_delay(p, batch.unpack(row, i), f"{p.name}_{signal_name}", params)


def fit(psr, fitter: str = "auto", **fitter_kwargs):
    """simulate.py:44-69."""
    _need_real_pint()
    (pfit,) = _need("pint.fitter")
    if fitter == "wls":
        f = pfit.WLSFitter(psr.toas, psr.model)
    elif fitter == "gls":
        f = pfit.GLSFitter(psr.toas, psr.model)
    elif fitter == "downhill":
        f = pfit.DownhillGLSFitter(psr.toas, psr.model)
    elif fitter == "auto":
        f = pfit.Fitter.auto(psr.toas, psr.model)
    else:
        raise ValueError(f"{fitter=} must be one of 'wls', 'gls', 'downhill' or 'auto'")
    f.fit_toas(**fitter_kwargs)
    psr.f = f
    psr.model = f.model
    psr.update_residuals()


def to_enterprise(psr, ephem: str = "DE440"):
    """simulate.py:91-95."""
    _need_real_pint()
    (epulsar,) = _need("enterprise.pulsar")
    if not callable(getattr(epulsar, "Pulsar", None)) or getattr(epulsar, "__file__", None) is None:
        raise PintUnavailable("enterprise is not installed")
    return epulsar.Pulsar(psr.toas, psr.model, ephem=ephem, timing_package="pint")


def write_partim(psr, outpar: str, outtim: str, tempo2: bool = False):
    """simulate.py:71-77 for PINT-backed pulsars."""
    psr.model.write_parfile(outpar)
    if tempo2:
        psr.toas.write_TOA_file(outtim, format="Tempo2")
    else:
        psr.toas.write_TOA_file(outtim)
