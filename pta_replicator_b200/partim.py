"""PINT-free readers for tempo2-style ``.par`` / ``.tim`` files.

The reference delegates ingest to PINT (``simulate.py:154-156``: ``get_model``,
``get_TOAs``).  PINT is out of scope for the hot path;
what the injection kernels need from the files is small: the pulsar name and sky
position from the par file; per TOA the MJD (kept in long double like PINT's
``tdbld``), the uncertainty in microseconds, the observing frequency, the site
and the ``-flag value`` pairs from the tim file.
"""
from __future__ import annotations

import os

import numpy as np

_SKIP_PREFIXES = ("C ", "#", "FORMAT", "MODE", "TIME", "EFAC", "EQUAD", "JUMP", "SKIP", "NOSKIP", "END")


def _sexagesimal(text: str) -> float:
    """'-75:42:35.3' -> -75.70980...; plain decimal strings pass through."""
    if ":" not in text:
        return float(text)
    sign = -1.0 if text.strip().startswith("-") else 1.0
    parts = [abs(float(p)) for p in text.split(":")]
    while len(parts) < 3:
        parts.append(0.0)
    return sign * (parts[0] + parts[1] / 60.0 + parts[2] / 3600.0)


def read_par(path: str) -> dict:
    """Return ``{KEY: [tokens...]}`` for every par line plus parsed position fields.

    Adds ``'_name'`` (PSR/PSRJ/PSRB) and ``'_loc'``: ``{'RAJ': hours, 'DECJ': deg}``
    or ``{'ELONG': deg, 'ELAT': deg}`` -- the units PINT's ``model.RAJ.value`` etc.
    carry at ``simulate.py:159-162``.  Raises ``AttributeError`` if neither pair is
    present (``simulate.py:164``).
    """
    entries: dict = {}
    with open(path) as fh:
        for raw in fh:
            line = raw.strip()
            if not line or line.startswith("#") or line.startswith("C "):
                continue
            toks = line.split()
            entries[toks[0].upper()] = toks[1:]
    name = None
    for key in ("PSR", "PSRJ", "PSRB"):
        if key in entries and entries[key]:
            name = entries[key][0]
            break
    if "RAJ" in entries and "DECJ" in entries:
        loc = {"RAJ": _sexagesimal(entries["RAJ"][0]), "DECJ": _sexagesimal(entries["DECJ"][0])}
    elif "ELONG" in entries and "ELAT" in entries:
        loc = {"ELONG": float(entries["ELONG"][0]), "ELAT": float(entries["ELAT"][0])}
    elif "LAMBDA" in entries and "BETA" in entries:
        loc = {"ELONG": float(entries["LAMBDA"][0]), "ELAT": float(entries["BETA"][0])}
    else:
        raise AttributeError("No pulsar location information (RAJ/DECJ or ELONG/ELAT) in parfile.")
    entries["_name"] = name
    entries["_loc"] = loc
    return entries


def read_tim(path: str, _depth: int = 0) -> dict:
    """Parse a tempo2 ``FORMAT 1`` tim file (``INCLUDE`` followed one level deep).

    Returns a dict of columns: ``name`` (list), ``freq`` [MHz] f64, ``mjd`` long
    double, ``err_us`` f64, ``site`` (list), ``flags`` (list of dicts, values str).
    """
    names, freqs, mjds, errs, sites, flags = [], [], [], [], [], []
    base = os.path.dirname(path)
    with open(path) as fh:
        for raw in fh:
            line = raw.strip()
            if not line:
                continue
            if line.upper().startswith("INCLUDE"):
                if _depth > 4:
                    raise ValueError("tim INCLUDE nesting too deep")
                sub = read_tim(os.path.join(base, line.split()[1]), _depth + 1)
                names += sub["name"]; freqs += list(sub["freq"]); mjds += list(sub["mjd"])
                errs += list(sub["err_us"]); sites += sub["site"]; flags += sub["flags"]
                continue
            if line.startswith(_SKIP_PREFIXES) or line == "C":
                continue
            toks = line.split()
            if len(toks) < 5:
                continue
            try:
                fq = float(toks[1]); mj = np.longdouble(toks[2]); er = float(toks[3])
            except ValueError:
                continue
            fl = {}
            k = 5
            while k < len(toks):
                if toks[k].startswith("-") and not _is_number(toks[k]):
                    key = toks[k][1:]
                    if k + 1 < len(toks) and not (toks[k + 1].startswith("-") and not _is_number(toks[k + 1])):
                        fl[key] = toks[k + 1]
                        k += 2
                    else:
                        fl[key] = ""
                        k += 1
                else:
                    k += 1
            names.append(toks[0]); freqs.append(fq); mjds.append(mj); errs.append(er)
            sites.append(toks[4]); flags.append(fl)
    return {
        "name": names,
        "freq": np.asarray(freqs, dtype=float),
        "mjd": np.asarray(mjds, dtype=np.longdouble),
        "err_us": np.asarray(errs, dtype=float),
        "site": sites,
        "flags": flags,
    }


def _is_number(tok: str) -> bool:
    try:
        float(tok)
        return True
    except ValueError:
        return False


def write_tim(path: str, cols: dict, mjd=None) -> None:
    """Write a tempo2 ``FORMAT 1`` tim file from the column dict of :func:`read_tim`."""
    mjd = cols["mjd"] if mjd is None else mjd
    with open(path, "w") as fh:
        fh.write("FORMAT 1\nMODE 1\n")
        for i in range(len(mjd)):
            fl = " ".join(f"-{k} {v}".rstrip() for k, v in cols["flags"][i].items())
            fh.write(f" {cols['name'][i]} {cols['freq'][i]:.8f} {np.format_float_positional(mjd[i], precision=20, unique=False)} "
                     f"{cols['err_us'][i]:.5f} {cols['site'][i]} {fl}\n")
