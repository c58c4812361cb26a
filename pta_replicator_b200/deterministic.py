"""Deterministic signals -- drop-in for ``add_cgw`` of
``/root/reference/pta_replicator/deterministic.py:13-185`` (the only deterministic signal on the
hot path; catalogs, bursts and memory are "next" rows f1/f4).
"""
from __future__ import annotations

import numpy as np

from .constants import MPC2S, SOLAR2S
from .engine import PulsarBatch
from .simulate import TimeArray


def add_cgw(psr, gwtheta, gwphi, mc, dist, fgw, phase0, psi, inc, pdist=1.0, pphase=None, psrTerm=True, evolve=True,
            phase_approx=False, tref=0, signal_name="cw"):
    """Continuous-wave residuals of one circular SMBHB (Earth + optional pulsar term), evaluated by
    ``ptar_cgw_delay`` on the GPU.  Units as in the reference: mc [Msun], dist [Mpc], fgw [Hz],
    pdist [kpc], angles [rad], tref [s]."""
    batch = PulsarBatch([psr], exact_epochs=True)
    d = batch.cgw_delays(gwtheta, gwphi, mc, dist, fgw, phase0, psi, inc, pdist=pdist, pphase=pphase,
                         psrTerm=psrTerm, evolve=evolve, phase_approx=phase_approx, tref=tref)[0]
    res = np.empty(len(d))
    res[batch.order[0]] = d
    dt = TimeArray(res, "s")
    # the reference stores the unit-converted mc / dist and the halved phase0 (deterministic.py:51-56, :167-182)
    psr.update_added_signals("{}_".format(psr.name) + signal_name,
                             {"gwtheta": gwtheta, "gwphi": gwphi, "mc": mc * SOLAR2S, "dist": dist * MPC2S, "fgw": fgw,
                              "phase0": phase0 / 2, "psi": psi, "inc": inc, "pdist": pdist, "pphase": pphase,
                              "psrTerm": psrTerm, "evolve": evolve, "phase_approx": phase_approx, "tref": tref}, dt)
    psr.toas.adjust_TOAs(dt.to("day"))
    psr.update_residuals()
