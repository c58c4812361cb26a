"""TEST INFRASTRUCTURE ONLY -- numpy restatement of the throughput-mode normal stream
(``pta_replicator_b200/csrc/ptar_rng.cuh``): Philox4x32-10 (Salmon, Moraes, Dror, Shaw 2011;
constants as published / as in Random123 and cuRAND) followed by Box-Muller.

The integer part is bit-exact.  The kernel evaluates Box-Muller with fp32 MUFU intrinsics;
here the same fp32 uniforms are pushed through float64 log/sin/cos, so the two differ only by
the intrinsics' errors (tests/test_gpu_statistics.py measures and bounds them).
"""
from __future__ import annotations

import numpy as np

M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
W0, W1 = 0x9E3779B9, 0xBB67AE85
MASK = np.uint64(0xFFFFFFFF)

K_WHITE1, K_WHITE2, K_ECORR, K_RED, K_GWB = 1, 2, 3, 4, 5


def philox4x32_10(c0, c1, c2, c3, seed):
    """Vectorised Philox4x32-10.  ``c*`` broadcastable uint32-valued arrays; returns 4 uint32 arrays."""
    c0, c1, c2, c3 = [np.asarray(x, dtype=np.uint64) & MASK for x in np.broadcast_arrays(c0, c1, c2, c3)]
    k0 = int(seed) & 0xFFFFFFFF
    k1 = (int(seed) >> 32) & 0xFFFFFFFF
    for _ in range(10):
        p0 = M0 * c0
        p1 = M1 * c2
        hi0, lo0 = p0 >> np.uint64(32), p0 & MASK
        hi1, lo1 = p1 >> np.uint64(32), p1 & MASK
        c0, c1, c2, c3 = (hi1 ^ c1 ^ np.uint64(k0)), lo1, (hi0 ^ c3 ^ np.uint64(k1)), lo0
        k0 = (k0 + W0) & 0xFFFFFFFF
        k1 = (k1 + W1) & 0xFFFFFFFF
    return c0.astype(np.uint32), c1.astype(np.uint32), c2.astype(np.uint32), c3.astype(np.uint32)


def _box_muller(a, b):
    """ptar_rng.cuh::box_muller (mode 0) with the SAME fp32 uniforms (u1 and the angle are rounded to float32 exactly
    as the kernel's two FFMAs do) and the transcendental functions in float64."""
    u1 = (a.astype(np.float32).astype(np.float64) * 2.0 ** -32 + 2.0 ** -33).astype(np.float32).astype(np.float64)  # = fmaf
    r = np.sqrt(np.maximum(-1.3862943611198906 * np.log2(u1), 0.0))
    th = (b.astype(np.float32).astype(np.float64) * np.float64(np.float32(1.4629180792671596e-9))
          + np.float64(np.float32(-3.1415926535897931))).astype(np.float32).astype(np.float64)                    # = fmaf
    return r * np.cos(th), r * np.sin(th)


def normals4(block, kind, psr, rfield, seed):
    """The 4 normals of one counter (arrays broadcast) -> array [..., 4] float64."""
    block = np.asarray(block, dtype=np.uint64)
    rfield = np.asarray(rfield, dtype=np.uint64)
    c1 = np.uint64(int(kind) | (int(psr) << 8))
    w = philox4x32_10(block, c1, rfield & MASK, rfield >> np.uint64(32), seed)
    n0, n1 = _box_muller(w[0], w[1])
    n2, n3 = _box_muller(w[2], w[3])
    return np.stack([n0, n1, n2, n3], axis=-1)


def normals(kind, psr, realization, n, seed, idx0=0):
    """Elements idx0 .. idx0+n-1 of stream (kind, psr) for one global realization id."""
    z = normals4(np.arange(idx0, idx0 + n), kind, psr, int(realization) >> 2, seed)
    return z[:, int(realization) & 3]
