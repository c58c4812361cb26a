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


def add_catalog_of_cws(psr, gwtheta_list, gwphi_list, mc_list, dist_list, fgw_list, phase0_list, psi_list, inc_list,
                       pdist=1.0, pphase=None, psrTerm=True, evolve=True, phase_approx=False, tref=0,
                       chunk_size=10_000_000, signal_name="cw_catalog"):
    """Sum of many circular SMBHBs -- drop-in for ``add_catalog_of_cws`` (deterministic.py:188-318) with the numba
    loops (:321-561) replaced by ``ptar_cw_catalog`` (thread per TOA, catalog slices staged through shared memory,
    deterministic slice reduction, NaN contributions of already-merged binaries dropped).  ``chunk_size`` is accepted
    for compatibility; the whole catalog is processed in one call and recorded once in the ledger."""
    import ctypes as C

    import torch

    from . import _cabi
    from . import orf as orf_mod
    dev = _cabi.require_cuda()
    radec = orf_mod.psrlocs_from_pulsars([psr])[0]
    ptheta, pphi = np.pi / 2 - radec[1], radec[0]
    phat = (C.c_double * 3)(np.sin(ptheta) * np.cos(pphi), np.sin(ptheta) * np.sin(pphi), np.cos(ptheta))
    cat = np.ascontiguousarray(np.stack([np.asarray(a, dtype=np.float64) for a in
                                         (gwtheta_list, gwphi_list, mc_list, dist_list, fgw_list, phase0_list, psi_list,
                                          inc_list)]))
    n_src = cat.shape[1]
    toas = np.ascontiguousarray(np.asarray(psr.toas.get_mjds().value, dtype=np.float64) * 86400 - tref)
    n = len(toas)
    n_slices = int(max(1, min(64, (n_src + 255) // 256)))
    t_d, cat_d = torch.from_numpy(toas).to(dev), torch.from_numpy(cat).to(dev)
    out = torch.empty(n, dtype=torch.float64, device=dev)
    pre = torch.empty(n_src * 16, dtype=torch.float64, device=dev)
    partial = torch.empty(n_slices * n, dtype=torch.float64, device=dev)
    mode = 0 if evolve else (1 if phase_approx else 2)
    _cabi.check(_cabi.lib().ptar_cw_catalog(out.data_ptr(), t_d.data_ptr(), n, phat, cat_d.data_ptr(), n_src, float(pdist),
                                            float(pphase) if pphase is not None else 0.0, int(pphase is not None), mode,
                                            int(bool(psrTerm)), 0, pre.data_ptr(), partial.data_ptr(), n_slices,
                                            _cabi.current_stream()), "ptar_cw_catalog")
    dt = TimeArray(out.cpu().numpy(), "s")
    psr.update_added_signals("{}_".format(psr.name) + signal_name,
                             {"gwtheta_list": gwtheta_list, "gwphi_list": gwphi_list, "mc_list": mc_list, "dist_list": dist_list,
                              "fgw_list": fgw_list, "phase0_list": phase0_list, "psi_list": psi_list, "inc_list": inc_list,
                              "pdist": pdist, "pphase": pphase, "psrTerm": psrTerm, "evolve": evolve,
                              "phase_approx": phase_approx, "tref": tref}, dt)
    psr.toas.adjust_TOAs(dt.to("day"))
    psr.update_residuals()
