"""ctypes binding of libmmad_b200.so (the C-ABI in include/mmad_b200.h).

There is no fallback: if the library is missing and cannot be built, or a call
fails, this raises.  Nothing here imports oracle/.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, c_char_p, c_int32, c_int64, c_uint8, c_uint32, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "csrc", "libmmad_b200.so")
_lib = None


class MmadError(RuntimeError):
    pass


def lib_path() -> str:
    return _LIB_PATH


def load() -> ctypes.CDLL:
    """Load (building first if the .so is absent) and declare every entry point."""
    global _lib
    if _lib is not None:
        return _lib
    from . import build as _build

    try:
        _build.build()                                     # digest check; recompiles only when csrc/ or include/ changed
    except RuntimeError:
        if not os.path.exists(_LIB_PATH):                  # no nvcc on this box and nothing prebuilt
            raise
    if not os.path.exists(_LIB_PATH):
        raise MmadError(f"CUDA library not found at {_LIB_PATH}; run python -m multimodal_ad_b200.build")
    # MMAD_LIB: load another build of the same ABI instead (A/B measurements of kernel changes; tools/ only)
    lib = ctypes.CDLL(os.environ.get("MMAD_LIB") or _LIB_PATH)

    lib.mmad_last_error.restype = c_char_p
    lib.mmad_last_error.argtypes = []
    lib.mmad_abi_version.restype = ctypes.c_int
    lib.mmad_launch_count.restype = c_int64
    lib.mmad_executed_mma_flops.restype = c_int64
    lib.mmad_executed_mma_flops.argtypes = []

    P = c_void_p
    lib.mmad_roi_plan_create.argtypes = [P, c_int64, c_int32, POINTER(P)]
    lib.mmad_roi_plan_create_ex.argtypes = [P, c_int64, c_int32, c_int32, c_int32, c_int32, c_int32, POINTER(P)]
    lib.mmad_roi_plan_destroy.argtypes = [P]
    lib.mmad_roi_plan_counts.argtypes = [P, P]
    lib.mmad_roi_plan_counts_dev.argtypes = [P, POINTER(P)]
    lib.mmad_roi_pool_f32.argtypes = [P, P, c_int64, P, P, P, P]
    lib.mmad_roi_pool_host_f32.argtypes = [P, P, c_int64, P, P, P]
    lib.mmad_roi_pool_mean_backward_f32.argtypes = [P, P, c_int64, P, P]
    lib.mmad_roi_pool_algorithmic_bytes.argtypes = [P, c_int64]
    lib.mmad_roi_pool_algorithmic_bytes.restype = c_int64
    lib.mmad_roi_plan_programme.argtypes = [P, P, POINTER(c_int64), P, POINTER(c_int32), POINTER(c_int32),
                                            POINTER(c_int64), POINTER(c_int32)]
    lib.mmad_roi_plan_binding.argtypes = [P, c_int64, c_int32, POINTER(c_int32), POINTER(c_int32), POINTER(c_int32),
                                          P, P, P, P, P, P, P]
    I = ctypes.c_int
    lib.mmad_conv3d_fwd_bf16.argtypes = [P, P, P, P] + [I] * 10 + [P]
    lib.mmad_conv3d_fwd_bf16.restype = I
    lib.mmad_conv3d_stats_partials.argtypes = [I] * 9
    lib.mmad_conv3d_stats_partials.restype = I
    F, D_, L = ctypes.c_float, ctypes.c_double, c_int64
    sigs = {
        "mmad_conv3d_wgrad_bf16": [P, P, P] + [I] * 10 + [P],
        "mmad_wgrad_reduce": [P, I, P, I, I, I, P],
        "mmad_wgrad_reduce_ex": [P, I, P, I, I, I, I, I, P],
        "mmad_conv3d_fwd_ex_bf16": [P, P, P, L, P, P, P, I, P, P] + [I] * 11 + [P],
        "mmad_conv3d_wgrad_ex_bf16": [P, L, P, P] + [I] * 10 + [P],
        "mmad_conv3d_fwd_head_bf16": [P] * 9 + [I] * 9 + [P],
        "mmad_convtranspose3d_prep_weights": [P, P, I, I, P],
        "mmad_convtranspose3d_k2s2_fwd_bf16": [P, P, P, P, L, I, I, I, I, I, I, P],
        "mmad_conv3d_c1_blocks": [I, I, I, I],
        "mmad_conv3d_c1_wgrad_blocks": [I, I, I, I],
        "mmad_conv3d_c1_fwd": [P, P, P, P, P, P] + [I] * 8 + [P],
        "mmad_conv3d_c1_wgrad": [P, P, P] + [I] * 7 + [P],
        "mmad_maxpool3d_k2_fwd": [P, L, P, P, I, I, I, I, I, P],
        "mmad_maxpool3d_k2_bwd": [P, P, P, I, I, I, I, I, P],
        "mmad_head1x1_fwd": [P, P, P, P] + [I] * 9 + [P],
        "mmad_head1x1_bwd_blocks": [],
        "mmad_head1x1_bwd": [P, P, P, P, P, P, P] + [I] * 9 + [P],
        "mmad_bn_apply_ex": [P, P, P, P, P, P, I, P, L, P, L, I, P],
        "mmad_roi_pool_ndhwc_f32": [P, P, L, I, I, I, I, I, I, I, P, P],
        "mmad_conv3d_prep_weights": [P, P, P, I, I, I, P],
        "mmad_conv3d_prep_weights_s2": [P, P, I, I, P],
        "mmad_conv3d_prep_weights_batched": [I, P, P, P, P, P, P, P],
        "mmad_conv3d_dgrad_s2_bf16": [P, P, P, I, I, I, I, I, I, P],
        "mmad_stem_s2d_elems": [I, I, I, I],
        "mmad_stem_s2d_pack": [P, P, I, I, I, I, P],
        "mmad_stem_s2d_prep_weights": [P, P, P],
        "mmad_stem_s2d_stats_partials": [I, I, I, I],
        "mmad_stem_s2d_fwd": [P, P, P, P, I, I, I, I, P],
        "mmad_stem_s2d_wgrad_workspace": [I, I, I, I, P],
        "mmad_stem_s2d_wgrad": [P, P, P, I, I, I, I, P],
        "mmad_stem_s2d_wgrad_reduce": [P, I, P, P],
        "mmad_bn_finalize": [P, I, I, D_, P, P, F, F, P, P, P, P, P, P, P],
        "mmad_bn_eval_params": [I, P, P, P, P, F, P, P, P, P, P],
        "mmad_bn_apply": [P, P, P, P, P, P, I, P, P, L, I, P],
        "mmad_bn_bwd_partials": [L],
        "mmad_bn_bwd_reduce": [P, P, P, P, P, P, P, P, P, P, P, L, I, P],
        "mmad_bn_bwd_finalize": [P, I, I, D_, P, P, P, I, P, P, P, P],
        "mmad_bn_bwd_apply": [P, P, P, P, L, I, P],
        "mmad_bn_bwd_apply_ex": [P, P, P, P, P, P, L, I, P],
        "mmad_maxpool3d_fwd": [P, P, P, I, I, I, I, I, P],
        "mmad_maxpool3d_bwd": [P, P, P, I, I, I, I, I, P],
        "mmad_upsample_zero2": [P, P] + [I] * 8 + [P],
        "mmad_stem_bn_relu_maxpool_fwd": [P, P, P, P, P, I, I, I, I, I, P],

        "mmad_ncs_f32_to_nsc_bf16": [P, P, I, I, L, P],
    }
    for name, args in sigs.items():
        if os.environ.get("MMAD_LIB") and not hasattr(lib, name):
            continue                                       # an older build loaded for an A/B measurement
        getattr(lib, name).argtypes = args
        getattr(lib, name).restype = I
    lib.mmad_conv3d_wgrad_workspace.argtypes = [I] * 10 + [POINTER(ctypes.c_int)]
    lib.mmad_conv3d_wgrad_workspace.restype = c_int64
    lib.mmad_stem_s2d_elems.restype = c_int64
    lib.mmad_stem_s2d_wgrad_workspace.argtypes = [I, I, I, I, POINTER(ctypes.c_int)]
    lib.mmad_stem_s2d_wgrad_workspace.restype = c_int64
    for name in ("mmad_roi_plan_create", "mmad_roi_plan_create_ex", "mmad_roi_plan_destroy", "mmad_roi_plan_counts",
                 "mmad_roi_plan_counts_dev", "mmad_roi_pool_f32", "mmad_roi_pool_host_f32",
                 "mmad_roi_pool_mean_backward_f32", "mmad_roi_plan_programme", "mmad_roi_plan_binding"):
        getattr(lib, name).restype = ctypes.c_int
    _lib = lib
    return lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = load().mmad_last_error().decode("utf-8", "replace")
        raise MmadError(f"{what} failed ({rc}): {msg}")


def launch_count() -> int:
    return int(load().mmad_launch_count())


def executed_mma_flops() -> int:
    """Tensor-core FLOPs the library's convolution kernels have issued on the current device since load (synchronises)."""
    return int(load().mmad_executed_mma_flops())
