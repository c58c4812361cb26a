"""ctypes binding of ``libptar_b200.so`` (C ABI declared in ``include/ptar.h``).

The shared library is built in-tree by ``__graft_entry__.build()`` (plain ``nvcc``; no torch
in the ABI).  There is no CPU fallback: if the library is missing, or there is no CUDA
device, the product raises instead of computing anywhere else.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# PTAR_B200_LIB points at an alternative build of the same library (A/B measurements of kernel variants)
LIB_PATH = os.environ.get("PTAR_B200_LIB") or os.path.join(_HERE, "csrc", "libptar_b200.so")

TILE_TOAS = 1024
TILE_EPOCHS = 64
TOA_ALIGN = 4
I8_SLICES, I8_BM, I8_BN, I8_BK = 6, 128, 32, 32   # tcgen05 GWB synthesis tile (csrc/ptar_gwb_i8.cuh)

F_WHITE, F_ECORR, F_RED, F_GWB, F_DET, F_WHITE1 = 1, 2, 4, 8, 16, 32
K_WHITE1, K_WHITE2, K_ECORR, K_RED, K_GWB = 1, 2, 3, 4, 5

EXPORTS = ("ptar_version", "ptar_last_error", "ptar_cholesky_lower", "ptar_fourier_basis", "ptar_cgw_delay", "ptar_cw_catalog", "ptar_burst_delay", "ptar_memory_delay",
           "ptar_gwb_mix", "ptar_gwb_synth", "ptar_gwb_mix_i8", "ptar_gwb_slice_i8", "ptar_gwb_synth_i8", "ptar_debug_i8_timestamps", "ptar_generate", "ptar_generate_stage", "ptar_philox_normals", "ptar_peer_export", "ptar_peer_open", "ptar_peer_close", "ptar_peer_copy", "ptar_run_job",
           "ptar_run_job_to_host")


class Tile(C.Structure):
    _fields_ = [("toa_start", C.c_int32), ("n_toa", C.c_int32), ("toa_local0", C.c_int32), ("ep_start", C.c_int32),
                ("n_ep", C.c_int32), ("psr", C.c_int32), ("nd", C.c_int32), ("reserved", C.c_int32)]


class GenParams(C.Structure):
    _fields_ = [
        ("n_psr", C.c_int32), ("n_tiles", C.c_int32), ("J", C.c_int32), ("npts", C.c_int32),
        ("flags", C.c_uint32), ("rn_convention", C.c_int32),
        ("tiles", C.c_void_p),
        ("w1", C.c_void_p), ("w2", C.c_void_p), ("dtau", C.c_void_p), ("eloc", C.c_void_p), ("det", C.c_void_p),
        ("ep_ecorr", C.c_void_p), ("ep_bucket", C.c_void_p), ("ep_gidx", C.c_void_p), ("ep_gw", C.c_void_p),
        ("ep_ginv", C.c_void_p), ("psr_bucket_off", C.c_void_p), ("Ftile", C.c_void_p),
        ("rn_scale", C.c_void_p), ("rn_omega", C.c_void_p),
        ("G", C.c_void_p), ("g_ld", C.c_int64), ("g_ldr", C.c_int64),
        ("z1", C.c_void_p), ("z2", C.c_void_p), ("zb", C.c_void_p), ("zrn", C.c_void_p),
        ("n_bucket_total", C.c_int64),
        ("seed", C.c_uint64), ("real0", C.c_int64),
        ("out", C.c_void_p), ("ld_out", C.c_int64), ("nreal", C.c_int32), ("rc", C.c_int32),
        ("Cbuf", C.c_void_p), ("cbuf_len", C.c_int64), ("c_rows", C.c_int64),
    ]


class Job(C.Structure):
    _fields_ = [("gen", GenParams), ("M", C.c_void_p), ("A", C.c_void_p), ("lda", C.c_int64), ("Jg", C.c_int32),
                ("lower_tri", C.c_int32), ("tile_list", C.c_void_p), ("knots", C.c_void_p), ("n_syn_tiles", C.c_int32),
                ("reserved", C.c_int32), ("Zm", C.c_void_p), ("Gbuf", C.c_void_p), ("gwb_zin", C.c_void_p),
                ("AS", C.c_void_p), ("colscale", C.c_void_p), ("ZS", C.c_void_p), ("zscale", C.c_void_p), ("zinv", C.c_void_p),
                ("tile_list_i8", C.c_void_p), ("rcap", C.c_int64), ("Jpad", C.c_int32), ("n_syn_tiles_i8", C.c_int32)]


_lib = None


class PtarError(RuntimeError):
    pass


def lib():
    """Load the shared library (once).  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise PtarError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                        "(nvcc -gencode arch=compute_100a,code=sm_100a).  There is no CPU fallback.")
    L = C.CDLL(LIB_PATH)
    vp, i32, i64, u64 = C.c_void_p, C.c_int, C.c_int64, C.c_uint64
    L.ptar_version.restype = C.c_int
    L.ptar_last_error.restype = C.c_char_p
    L.ptar_cholesky_lower.argtypes = [vp, vp, i32, i32, vp, vp]
    L.ptar_fourier_basis.argtypes = [vp, vp, i64, vp, vp, vp, vp, i32, i32, i64, vp]
    L.ptar_cgw_delay.argtypes = [vp, vp, vp, vp, vp, i32, i32, i32, i64, vp]
    L.ptar_cw_catalog.argtypes = [vp, vp, i64, C.POINTER(C.c_double), vp, i64, C.c_double, C.c_double, i32, i32, i32, i32, vp, vp, i32, vp]
    L.ptar_burst_delay.argtypes = [vp, vp, vp, C.c_double, C.c_double, C.c_double, C.c_double, i32, i64, vp]
    L.ptar_memory_delay.argtypes = [vp, vp, C.c_double, C.c_double, i32, i64, vp]
    L.ptar_gwb_mix.argtypes = [vp, vp, vp, i32, i32, i64, u64, i64, vp]
    L.ptar_gwb_synth.argtypes = [vp, i64, i64, vp, i64, vp, i32, i64, vp, i32, vp, i32, vp]
    L.ptar_gwb_mix_i8.argtypes = [vp, vp, vp, i32, i32, i32, i64, i64, u64, i64, vp]
    L.ptar_gwb_slice_i8.argtypes = [vp, vp, vp, i32, i32, i32, i64, i64, vp]
    L.ptar_gwb_synth_i8.argtypes = [vp, i64, i64, vp, vp, vp, vp, i32, i32, i32, i64, i64, vp, i32, vp]
    L.ptar_debug_i8_timestamps.argtypes = [vp]
    L.ptar_generate.argtypes = [C.POINTER(GenParams), vp]
    L.ptar_generate_stage.argtypes = [C.POINTER(GenParams), i32, vp]
    L.ptar_philox_normals.argtypes = [vp, i32, i32, i64, i64, i64, u64, vp]
    L.ptar_peer_export.argtypes = [vp, vp, C.POINTER(C.c_int64)]
    L.ptar_peer_open.argtypes = [vp, C.POINTER(C.c_void_p)]
    L.ptar_peer_close.argtypes = [vp]
    L.ptar_peer_copy.argtypes = [vp, vp, i64, vp]
    L.ptar_run_job.argtypes = [C.POINTER(Job), i64, C.c_int32, vp, vp]
    L.ptar_run_job_to_host.argtypes = [C.POINTER(Job), i64, i64, C.c_int32, vp, vp, vp, vp, vp]
    for name in EXPORTS:
        if name not in ("ptar_version", "ptar_last_error"):
            getattr(L, name).restype = C.c_int
    _lib = L
    return L


def check(rc: int, what: str = "ptar"):
    if rc != 0:
        raise PtarError(f"{what} failed ({rc}): {lib().ptar_last_error().decode()}")


def require_cuda():
    """The product path runs on the GPU only."""
    import torch
    if not torch.cuda.is_available():
        raise PtarError("pta_replicator_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback.")
    lib()
    return torch.device("cuda", torch.cuda.current_device())


def ptr(t):
    """Device pointer of a torch tensor (None -> NULL)."""
    return None if t is None else t.data_ptr()


def current_stream():
    import torch
    return torch.cuda.current_stream().cuda_stream
