"""ctypes binding of ``libpvb200.so`` -- the C ABI declared in ``include/pvb200.h``.

This is the only place the product touches native code.  There is NO fallback: if the shared
library is missing or a call fails, a ``RuntimeError`` is raised (the product path must fail loudly,
never route through the CPU oracle).

Build the library in-tree with ``python -c "import __graft_entry__ as g; g.build()"`` or
``make -C predict_pv_yield_b200/csrc``.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libpvb200.so")

c_void_p, c_int, c_ll, c_size_t, c_float = C.c_void_p, C.c_int, C.c_longlong, C.c_size_t, C.c_float


class Head(C.Structure):
    """Mirror of ``pvb200_head_t`` (include/pvb200.h) -- field order and types must match exactly."""

    _fields_ = [
        ("struct_size", c_size_t),
        ("B", c_int), ("F1", c_int), ("F2", c_int), ("F3", c_int), ("FO", c_int),
        ("NPV", c_int), ("NNWP", c_int), ("FNWP", c_int), ("pv_ns", c_int),
        ("K1", c_ll), ("pv_sb", c_ll), ("pv_st", c_ll),
        ("w1", c_void_p), ("b1", c_void_p), ("w2", c_void_p), ("b2", c_void_p), ("wn", c_void_p), ("bn", c_void_p),
        ("w3", c_void_p), ("b3", c_void_p), ("w4", c_void_p), ("b4", c_void_p),
        ("x", c_void_p), ("pv", c_void_p), ("nwp", c_void_p),
        ("h1", c_void_p), ("cat", c_void_p), ("h3", c_void_p), ("out", c_void_p),
        ("g_out", c_void_p), ("g_h3", c_void_p), ("g_cat", c_void_p), ("g_h1", c_void_p), ("g_x", c_void_p),
        ("dw1", c_void_p), ("db1", c_void_p), ("dw2", c_void_p), ("db2", c_void_p), ("dwn", c_void_p), ("dbn", c_void_p),
        ("dw3", c_void_p), ("db3", c_void_p), ("dw4", c_void_p), ("db4", c_void_p),
        ("workspace", c_void_p), ("workspace_bytes", c_size_t),
    ]


# name -> (restype, argtypes); every symbol include/pvb200.h declares
SIGNATURES = {
    "pvb200_abi_version": (c_int, []),
    "pvb200_last_error": (C.c_char_p, []),
    "pvb200_launch_count": (C.c_ulonglong, []),
    "pvb200_reset_launch_count": (None, []),
    "pvb200_sm_count": (c_int, []),
    "pvb200_probe_fp32_fma": (c_int, [c_void_p, c_int, C.POINTER(C.c_double), c_void_p]),
    "pvb200_probe_fp32_fma2": (c_int, [c_void_p, c_int, C.POINTER(C.c_double), c_void_p]),
    "pvb200_sat_normalise_f32": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_ll, c_void_p]),
    "pvb200_sat_normalise_bf16": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_ll, c_void_p]),
    "pvb200_conv3d_workspace_bytes": (c_size_t, [c_int, c_int]),
    "pvb200_conv3d_fwd_f32": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                      c_size_t, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "pvb200_conv3d_dgrad_f32": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t,
                                        c_int, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "pvb200_blocked_channel_groups": (c_int, [c_int]),
    "pvb200_conv3d_bf16_workspace_bytes": (c_size_t, [c_int, c_int]),
    "pvb200_nc_to_blocked_bf16": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "pvb200_blocked_to_nc_f32": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "pvb200_conv3d_fwd_bf16": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t,
                                       c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "pvb200_sat_normalise_blocked_bf16": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int,
                                                  c_int, c_void_p]),
    "pvb200_conv3d_dgrad_bf16": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t,
                                         c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "pvb200_conv3d_wgrad_bf16_gz_plane": (c_ll, [c_int, c_int]),
    "pvb200_conv3d_wgrad_bf16_workspace_bytes": (c_size_t, [c_int, c_int]),
    "pvb200_nc_to_gzw_bf16": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "pvb200_conv3d_wgrad_bf16": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t,
                                         c_int, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "pvb200_conv3d_wgrad_workspace_bytes": (c_size_t, [c_int, c_int]),
    "pvb200_conv3d_wgrad_f32": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                        c_size_t, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "pvb200_head_fwd_workspace_bytes": (c_size_t, [c_int, c_int, c_ll]),
    "pvb200_head_fwd_f32": (c_int, [C.POINTER(Head), c_void_p]),
    "pvb200_head_bwd_f32": (c_int, [C.POINTER(Head), c_void_p]),
    "pvb200_head_tail_fwd_f32": (c_int, [C.POINTER(Head), c_int, c_void_p]),
    "pvb200_head_tail_bwd_f32": (c_int, [C.POINTER(Head), c_void_p]),
    "pvb200_conv3d_fwd_f32_tpad": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t,
                                           c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "pvb200_conv3d_dgrad_f32_tpad": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t,
                                             c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "pvb200_conv3d_wgrad_f32_tpad": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                             c_size_t, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "pvb200_conv3d_fwd_bf16_tpad": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t,
                                            c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "pvb200_conv3d_dgrad_bf16_tpad": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t,
                                              c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "pvb200_conv3d_wgrad_bf16_tpad": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t,
                                              c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "pvb200_blocked4_channel_groups": (c_int, [c_int]),
    "pvb200_conv3d_tf32x3_workspace_bytes": (c_size_t, [c_int, c_int]),
    "pvb200_nc_to_blocked_f32": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    "pvb200_blocked_f32_to_nc": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "pvb200_sat_normalise_blocked_f32": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int,
                                                 c_int, c_void_p, c_void_p]),
    "pvb200_conv3d_fwd_tf32x3": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t,
                                         c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p,
                                         c_void_p]),
    "pvb200_conv3d_dgrad_tf32x3": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t,
                                           c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "pvb200_conv3d_wgrad_bf16x3_supported": (c_int, [c_int, c_int, c_int, c_int]),
    "pvb200_conv3d_wgrad_bf16x3_workspace_bytes": (c_size_t, []),
    "pvb200_conv3d_wgrad_bf16x3": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_size_t,
                                           c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "pvb200_conv3d_wgrad_f16x2": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t,
                                          c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "pvb200_absmax_f32": (c_int, [c_void_p, c_ll, c_void_p, c_void_p]),
    "pvb200_conv3d_wgrad_bf16_rows_supported": (c_int, [c_int, c_int, c_int, c_int]),
    "pvb200_conv3d_wgrad_bf16_rows_workspace_bytes": (c_size_t, []),
    "pvb200_conv3d_wgrad_bf16_rows": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_size_t,
                                              c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "pvb200_conv3d_fwd_f32_pad": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t,
                                          c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "pvb200_conv3d_dgrad_f32_pad": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t,
                                            c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "pvb200_conv3d_wgrad_f32_pad": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                            c_size_t, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "pvb200_maxpool3d_fwd_f32": (c_int, [c_void_p, c_void_p, c_void_p, c_ll, c_int, c_int, c_int, c_void_p]),
    "pvb200_maxpool3d_bwd_f32": (c_int, [c_void_p, c_void_p, c_void_p, c_ll, c_int, c_int, c_int, c_void_p]),
    "pvb200_linear_workspace_bytes": (c_size_t, [c_int, c_int, c_ll]),
    "pvb200_linear_fwd_f32": (c_int, [c_void_p, c_ll, c_void_p, c_void_p, c_void_p, c_ll, c_int, c_ll, c_int, c_int, c_void_p,
                                      c_size_t, c_void_p]),
    "pvb200_linear_bwd_f32": (c_int, [c_void_p, c_ll, c_void_p, c_void_p, c_ll, c_void_p, c_ll, c_void_p, c_ll, c_int, c_void_p,
                                      c_void_p, c_int, c_ll, c_int, c_void_p, c_size_t, c_void_p]),
    "pvb200_linear_finish_f32": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_ll, c_int, c_int, c_int, c_void_p]),
    "pvb200_linear_gpre_f32": (c_int, [c_void_p, c_ll, c_void_p, c_ll, c_void_p, c_void_p, c_int, c_int, c_void_p]),
    "pvb200_embedding_fwd_f32": (c_int, [c_void_p, c_void_p, c_void_p, c_ll, c_int, c_int, c_int, c_void_p]),
    "pvb200_embedding_bwd_f32": (c_int, [c_void_p, c_ll, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p]),
    "pvb200_history_flatten_f32": (c_int, [c_void_p, c_ll, c_ll, c_void_p, c_ll, c_int, c_int, c_int, c_void_p]),
    "pvb200_reserve_sms": (c_int, [c_int]),
    "pvb200_set_dynamic_tiles": (c_int, [c_int]),
    "pvb200_fc1_bf16_shadow_bytes": (c_size_t, [c_int, c_int, c_int, c_int]),
    "pvb200_fc1_make_shadow_bf16": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "pvb200_adam_fc1_shadow": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int,
                                       c_float, c_float, c_float, c_float, c_int, c_float, c_void_p]),
    "pvb200_adam_fc1_shadow_rows": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int,
                                            c_int, c_int, c_float, c_float, c_float, c_float, c_int, c_float, c_void_p]),
    "pvb200_fc1_shadow_from_shards": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "pvb200_fc1_fwd_bf16_splits": (c_int, []),
    "pvb200_fc1_fwd_bf16": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "pvb200_fc1_dgrad_bf16": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int,
                                      c_int, c_void_p]),
    "pvb200_fc1_wgrad_bf16": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "pvb200_l1_loss_fwd_f32": (c_int, [c_void_p, c_void_p, c_ll, c_ll, c_void_p, c_void_p, c_int, c_int, c_void_p]),
    "pvb200_l1_loss_bwd_f32": (c_int, [c_void_p, c_void_p, c_ll, c_ll, c_void_p, c_void_p, c_int, c_int, c_void_p]),
    "pvb200_validation_results_f32": (c_int, [c_void_p, c_void_p, c_ll, c_ll, c_void_p, c_ll, c_ll, c_void_p, c_void_p, c_int,
                                              c_int, c_void_p]),
    "pvb200_adam_step_f32": (c_int, [c_int, C.POINTER(c_void_p), C.POINTER(c_void_p), C.POINTER(c_void_p),
                                     C.POINTER(c_void_p), C.POINTER(c_ll), c_float, c_float, c_float, c_float, c_int,
                                     c_float, c_void_p]),
}

_lib: Optional[C.CDLL] = None


def load() -> C.CDLL:
    """Load ``libpvb200.so`` and declare every signature.  Raises if the library is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} not found: the CUDA extension is not built (run __graft_entry__.build()). "
            "predict_pv_yield_b200 has no CPU fallback."
        )
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    if lib.pvb200_abi_version() != 1:
        raise RuntimeError("libpvb200.so ABI version mismatch")
    _lib = lib
    return lib


def check(rc: int, what: str = "") -> None:
    """Turn a non-zero status into a RuntimeError carrying ``pvb200_last_error()``."""
    if rc != 0:
        msg = load().pvb200_last_error().decode("utf-8", "replace")
        raise RuntimeError(f"libpvb200 {what} failed (status {rc}): {msg}")


def launch_count() -> int:
    return int(load().pvb200_launch_count())


def reset_launch_count() -> None:
    load().pvb200_reset_launch_count()
