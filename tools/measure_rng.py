"""Measure the throughput-mode normal stream (fp32 MUFU Box-Muller) against a float64 Box-Muller of the same
uniforms: quantiles of |d|, where the large differences sit, KS / AD statistics.  Prints one JSON object.
    python tools/measure_rng.py        (GPU box)"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from oracle import philox as PH  # noqa: E402
from pta_replicator_b200 import _cabi  # noqa: E402

dev = _cabi.require_cuda()
L = _cabi.lib()
n = 1 << 22
out = torch.empty(n, dtype=torch.float32, device=dev)
res = {}
allg, allr = [], []
for kind, psr, real, seed in ((1, 0, 0, 1), (2, 66, 123457, 0xDEADBEEFCAFE1234), (5, 12, 99998, 77)):
    _cabi.check(L.ptar_philox_normals(out.data_ptr(), kind, psr, real, 0, n, seed, None))
    torch.cuda.synchronize()
    allg.append(out.cpu().numpy().astype(np.float64))
    allr.append(PH.normals(kind, psr, real, n, seed))
g, r = np.concatenate(allg), np.concatenate(allr)
d = np.abs(g - r)
res["n"] = int(len(d))
res["abs_err_quantiles"] = {str(q): float(np.quantile(d, q)) for q in (0.5, 0.9, 0.99, 0.999, 0.9999, 0.99999)}
res["abs_err_max"] = float(d.max())
for thr in (1e-5, 2e-5, 1e-4, 1e-3):
    big = d > thr
    res[f"frac_err_gt_{thr:g}"] = float(big.mean())
    res[f"max_abs_z_where_err_gt_{thr:g}"] = float(np.abs(r[big]).max()) if big.any() else None
rel = d / np.maximum(np.abs(r), 1e-3)
res["rel_err_quantiles_for_absz_gt_1e-3"] = {str(q): float(np.quantile(rel, q)) for q in (0.5, 0.99, 0.9999)}
res["rel_err_max_absz_gt_0.1"] = float((d[np.abs(r) > 0.1] / np.abs(r[np.abs(r) > 0.1])).max())
print(json.dumps(res, indent=1))
