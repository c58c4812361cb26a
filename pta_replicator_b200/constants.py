"""Unit constants of the hot path.

Values must equal ``/root/reference/pta_replicator/constants.py:3-8`` (which
derives them from ``scipy.constants``); they are restated here from CODATA so
the product does not need scipy at import time.  ``GWB_F1YR`` is the *different*
year ``add_gwb`` uses (``red_noise.py:248``).
"""
DAY_IN_SEC = 86400
YEAR_IN_SEC = 365.25 * DAY_IN_SEC
DMk = 4.15e3  # MHz^2 cm^3 pc s

_G = 6.67430e-11          # CODATA 2018, as in scipy.constants
_C = 299792458.0
_PARSEC = 3.085677581491367e16  # au / tan(1 arcsec), scipy.constants.parsec

SOLAR2S = _G / _C**3 * 1.98855e30
KPC2S = _PARSEC / _C * 1e3
MPC2S = _PARSEC / _C * 1e6

GWB_F1YR = 1 / 3.16e7
