"""ctypes binding of the C ABI declared in include/pdm_b200.h.

Loads the in-tree ``lib/libpdm_b200.so`` (built by ``build.py`` with nvcc for sm_100a).  There is no
fallback: if the library is missing or a call fails, a ``PdmError`` is raised.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

HERE = os.path.dirname(os.path.abspath(__file__))
PKG_ROOT = os.path.dirname(HERE)
LIB_PATH = os.environ.get("PDM_B200_LIB") or os.path.join(PKG_ROOT, "lib", "libpdm_b200.so")   # env: dev variants

PDM_OK = 0
PREC_EXACT_F32, PREC_F16X3, PREC_F16X1, PREC_F16X2, PREC_F8X1 = 0, 1, 2, 3, 4
PART_STRIDE = 8
TOPK_SLOTS = 8
OUT_E_MIN, OUT_LOG_L, OUT_MEAN_E, OUT_MEAN_E2, OUT_VAR_E, OUT_AUX_MEAN, OUT_ENTROPY, OUT_L = range(8)
OUT_ROWS = 8


class PdmError(RuntimeError):
    pass


class StatsArgs(C.Structure):
    """Mirror of ``struct pdm_stats_args`` (include/pdm_b200.h) -- field order and types must match."""
    _fields_ = [
        ("precision", C.c_int32), ("n_splits", C.c_int32), ("m_group", C.c_int32), ("cta_group", C.c_int32),
        ("records_per_row", C.c_int32), ("reserved0", C.c_int32),
        ("M", C.c_int64), ("N", C.c_int64), ("d", C.c_int64), ("index_offset", C.c_int64),
        ("q", C.c_void_p), ("ldq", C.c_int64), ("y", C.c_void_p), ("ldy", C.c_int64),
        ("q_hi", C.c_void_p), ("q_lo", C.c_void_p), ("ldqh", C.c_int64), ("q_inv_scale", C.c_void_p),
        ("y_hi", C.c_void_p), ("y_lo", C.c_void_p), ("ldyh", C.c_int64), ("y_inv_scale", C.c_float),
        ("q_norm", C.c_void_p), ("y_norm", C.c_void_p), ("inv_temp", C.c_void_p), ("y_aux", C.c_void_p),
        ("partials", C.c_void_p), ("energy_out", C.c_void_p), ("lde", C.c_int64), ("energy_mult", C.c_float),
        ("row_tiles", C.c_void_p), ("n_row_tiles", C.c_int64), ("n_row_tiles_dev", C.c_void_p),
        ("topk_val", C.c_void_p), ("topk_idx", C.c_void_p),
    ]


_P, _I64, _I32, _F = C.c_void_p, C.c_int64, C.c_int32, C.c_float

# name -> (restype, argtypes): every symbol include/pdm_b200.h declares
SIGNATURES = {
    "pdm_last_error": (C.c_char_p, []),
    "pdm_abi_version": (C.c_int, []),
    "pdm_device_info": (C.c_int, [C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "pdm_row_norms_f32": (C.c_int, [_P, _I64, _I64, _I64, _P, _P]),
    "pdm_prepare_rows": (C.c_int, [_P, _I64, _I64, _P, _I64, _P, _P, _I64, _I64, _F, _P, _I64, _P, _P, _P, _I64, _P, _P]),
    "pdm_noised_rows_philox": (C.c_int, [C.c_uint64, C.c_uint64, C.c_uint64, _I64, _P, _I64, _I64, _I64, _P, _I64, _P, _P, _I64,
                                         _P, _P, _I64, _P, _P]),
    "pdm_split_row_norms": (C.c_int, [_P, _P, _I64, _P, _I64, _I64, _P, _P]),
    "pdm_row_absmax_f32": (C.c_int, [_P, _I64, _I64, _I64, _P, _P]),
    "pdm_absmax_f32": (C.c_int, [_P, _I64, _I64, _I64, _P, _P]),
    "pdm_lattice_residual_f32": (C.c_int, [_P, _I64, _I64, _I64, _F, _P, _P]),
    "pdm_transpose_split_f16": (C.c_int, [_P, _I64, _I64, _I64, _F, _P, _P, _I64, _P]),
    "pdm_column_moments_f32": (C.c_int, [_P, _I64, _I64, _I64, _P, _P, _P, _P]),
    "pdm_posterior_stats_plan": (C.c_int, [C.POINTER(StatsArgs), C.c_int, C.POINTER(C.c_int64)]),
    "pdm_posterior_stats": (C.c_int, [C.POINTER(StatsArgs), _P]),
    "pdm_merge_partials": (C.c_int, [_P, _I64, _I64, _I64, _I64, _I64, _I64, _P, _I64, _P, _P, _P]),
    "pdm_reduce_partials": (C.c_int, [_P, _I64, _I64, _I64, _I64, _I64, _I64, _P, _P, _P]),
    "pdm_screen_temperatures": (C.c_int, [_P, _P, _I64, _P, _F, _F, _F, _P, _P]),
    "pdm_screen_certify": (C.c_int, [_P, _I64, _F, _I32, _P, _P, _P, _P]),
    "pdm_split_to_e4m3": (C.c_int, [_P, _P, _I64, _I64, _I64, _P, _I64, _P, _P]),
    "pdm_screen_temperatures_f8": (C.c_int, [_P, _P, _P, _P, _I64, _P, _P, _F, _F, _F, _P, _P]),
    "pdm_screen_tile_list": (C.c_int, [_P, _I64, _I32, _P, _P, _P]),
    "pdm_screen_finalize": (C.c_int, [_P, _P, _I64, _I64, _P, _P, _I64, _P, _P, _P, _P, _I64, _F, _P, _P, _I64, _I64, _I64,
                                      _P, _P, _P]),
    "pdm_weights_from_energy": (C.c_int, [_P, _I64, _I64, _I64, _P, _P, _P, _P, _I64, _P, _P, _I64, _P]),
    "pdm_weights_from_energy_tiles": (C.c_int, [_P, _I64, _I64, _I64, _P, _P, _P, _P, _I64, _P, _P, _I64, _P, _I32, _I64, _P, _P]),
    "pdm_split_gemm_f16x3_tiles": (C.c_int, [_P, _P, _I64, _I64, _P, _P, _I64, _I64, _I64, _F, _P, _I64, _I32, _I32, _P, _I64, _P, _P]),
    "pdm_delta_tile_list": (C.c_int, [_P, _I64, _I32, _P, _P, _P, _P]),
    "pdm_screen_merge_stage": (C.c_int, [_P, _P, _I64, _I32, _I64, _P, _P, _P, _P, _P]),
    "pdm_topk_merge": (C.c_int, [_P, _P, _I64, _I64, _I64, _I32, _P, _P, _P]),
    "pdm_refine_neighbours_f32": (C.c_int, [_P, _I64, _I64, _I64, _P, _I64, _I64, _I64, _I32, _P, _P, _P]),
    "pdm_gather_rows_f32": (C.c_int, [_P, _I64, _I64, _I64, _P, _I64, _P, _I64, _P, _I64, _P]),
    "pdm_split_gemm_f16x3": (C.c_int, [_P, _P, _I64, _I64, _P, _P, _I64, _I64, _I64, _F, _P, _I64, _I32, _I32, _P]),
    "pdm_sampler_step_f32": (C.c_int, [_P, _P, _P, _F, _F, _F, _P, _I64, _P]),
    "pdm_sampler_step_dev_f32": (C.c_int, [_P, _P, _P, _P, _P, _I64, _P]),
    "pdm_topk_smallest_f32": (C.c_int, [_P, _I64, _I64, _I64, _I32, _P, _P, _P]),
    "pdm_denoiser_backward_weights": (C.c_int, [_P, _I64, _P, _I64, _I64, _I64, _P, _P, _P, _P, _P, _P, _I64, _P, _P]),
    "pdm_weighted_mean_exact_f32": (C.c_int, [_P, _I64, _I64, _I64, _P, _I64, _I64, _P, _I64, _I32, _P]),
}

_lib = None
_lock = threading.Lock()


def load() -> C.CDLL:
    """Load the shared library (once).  Raises PdmError when it has not been built."""
    global _lib
    with _lock:
        if _lib is None:
            if not os.path.exists(LIB_PATH):
                raise PdmError(
                    f"{LIB_PATH} is missing: build it with `python {os.path.join(PKG_ROOT, 'build.py')}` "
                    "(needs nvcc).  pdm_b200 has no CPU or PyTorch fallback.")
            lib = C.CDLL(LIB_PATH)
            for name, (res, args) in SIGNATURES.items():
                fn = getattr(lib, name)
                fn.restype, fn.argtypes = res, args
            if lib.pdm_abi_version() != 4:
                raise PdmError("libpdm_b200.so ABI version mismatch; rebuild the library")
            _lib = lib
    return _lib


def check(rc: int, what: str) -> None:
    if rc != PDM_OK:
        msg = load().pdm_last_error()
        raise PdmError(f"{what} failed (code {rc}): {msg.decode() if msg else '?'}")
