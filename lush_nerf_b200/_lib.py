"""ctypes binding of liblushnerf_b200.so (the C ABI declared in include/lushnerf_b200.h).

The product path has NO fallback: if the shared library is missing, or a tensor is not a contiguous
fp32 CUDA tensor, the call raises.  Build the library with `python -m lush_nerf_b200.build`.
"""
import ctypes
import os
from ctypes import POINTER, Structure, c_char_p, c_double, c_float, c_int, c_int32, c_int64, c_size_t, c_void_p

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("LUSH_LIB_PATH") or os.path.join(_HERE, "liblushnerf_b200.so")   # override: A/B runs of two builds


class MlpDesc(Structure):
    """Mirror of `struct lush_mlp_desc`."""
    _fields_ = [("D", c_int32), ("W", c_int32), ("L_pts", c_int32), ("L_views", c_int32),
                ("skip", c_int32), ("out_mode", c_int32)]


class AdamSeg(Structure):
    """Mirror of `struct lush_adam_seg`."""
    _fields_ = [("param", c_void_p), ("grad", c_void_p), ("exp_avg", c_void_p), ("exp_avg_sq", c_void_p),
                ("numel", c_int64), ("step_index", c_int32), ("lr_index", c_int32)]


class DenseLayer(Structure):
    """Mirror of `struct lush_dense_layer`."""
    _fields_ = [("w", c_void_p), ("b", c_void_p), ("gw", c_void_p), ("gb", c_void_p),
                ("in0_off", c_int32), ("in0_dim", c_int32), ("in1_off", c_int32), ("in1_dim", c_int32),
                ("out_off", c_int32), ("out_dim", c_int32), ("act", c_int32)]


MOTION_MAX_LAYERS = 24


class MotionNet(Structure):
    """Mirror of `struct lush_motion_net`."""
    _fields_ = [("n_layers", c_int32), ("row_width", c_int32), ("embed_dim", c_int32), ("num_motion", c_int32),
                ("r_off", c_int32), ("v_off", c_int32), ("w_off", c_int32), ("rv_window", c_float),
                ("layers", DenseLayer * MOTION_MAX_LAYERS)]


_P = c_void_p
_SIGNATURES = {
    "lush_abi_version": (c_int, []),
    "lush_last_error": (c_char_p, []),
    "lush_launch_count": (c_int64, []),
    "lush_ray_prep_fwd": (c_int, [_P, _P, c_int64, c_int, c_int, c_double, c_float, c_float, c_int, _P]),
    "lush_ray_prep_bwd": (c_int, [_P, _P, _P, c_int64, c_int, c_int, c_double, c_int, _P]),
    "lush_ray_gen_prep": (c_int, [_P, c_int, c_int, c_int, c_double, c_double, c_double, c_double, c_float, c_float,
                                  c_int, c_int64, c_int64, _P, _P]),
    "lush_ray_gen_gather": (c_int, [_P, c_int, c_int, c_int, _P, _P, c_int, c_int, c_double, c_double, c_double,
                                    c_double, c_int64, _P, _P, _P, _P, _P, _P, _P]),
    "lush_stratified_z": (c_int, [_P, c_int, _P, _P, _P, c_int64, c_int, c_int, _P]),
    "lush_composite_fwd": (c_int, [_P, _P, _P, c_int, _P, c_float, c_int, _P, _P, _P, _P, _P, _P, c_int64,
                                   c_int, _P]),
    "lush_composite_bwd": (c_int, [_P, _P, _P, c_int, _P, c_float, c_int, _P, _P, _P, _P, _P, _P, _P,
                                   c_int64, c_int, _P]),
    "lush_sample_pdf": (c_int, [_P, _P, _P, _P, c_int64, _P, _P, c_int64, c_int, c_int, _P]),
    "lush_sample_merge": (c_int, [_P, _P, _P, _P, c_int64, _P, _P, _P, c_int64, c_int, c_int, _P]),
    "lush_mlp_param_count": (c_int64, [POINTER(MlpDesc)]),
    "lush_mlp_impl_supported": (c_int, [POINTER(MlpDesc), c_int]),
    "lush_mlp_workspace_bytes": (c_size_t, [POINTER(MlpDesc), c_int64, c_int, c_int, c_int]),
    "lush_mlp_stash_bytes": (c_size_t, [POINTER(MlpDesc), c_int64, c_int, c_int]),
    "lush_mlp_fwd": (c_int, [POINTER(MlpDesc), _P, _P, c_int, _P, _P, _P, _P, c_size_t, c_int64, c_int,
                             c_int, _P]),
    "lush_mlp_bwd": (c_int, [POINTER(MlpDesc), _P, _P, c_int, _P, _P, _P, _P, _P, _P, c_size_t, c_int64,
                             c_int, c_int, _P]),
    "lush_rbk_warp_fwd": (c_int, [_P, _P, _P, _P, c_int64, c_int, _P]),
    "lush_rbk_warp_bwd": (c_int, [_P, _P, _P, _P, _P, _P, _P, c_int64, c_int, _P]),
    "lush_blend_tonemap_fwd": (c_int, [_P, _P, _P, _P, _P, _P, c_int64, c_int, c_int, _P]),
    "lush_blend_tonemap_bwd": (c_int, [_P, _P, _P, _P, _P, _P, _P, _P, _P, c_int64, c_int, c_int, _P]),
    "lush_rbk_motion_fwd": (c_int, [POINTER(MotionNet), _P, c_int, _P, c_int64, _P, _P, _P, _P, _P, _P]),
    "lush_rbk_motion_bwd": (c_int, [POINTER(MotionNet), c_int, _P, c_int64, _P, _P, _P, _P, _P, _P]),
    "lush_blur_loss": (c_int, [_P, _P, _P, c_int64, _P, _P, _P, _P]),
    "lush_adam_step": (c_int, [POINTER(AdamSeg), c_int, _P, _P, _P, c_double, c_double, c_double, _P]),
    "lush_tc_selftest": (c_int, [_P, _P, _P, c_int, c_int, c_int, c_int, c_int, _P]),
    "lush_tc_microbench": (c_int, [_P, c_int, c_int, c_int, c_int, c_int, _P]),
    "lush_tc_selftest_pair": (c_int, [_P, _P, _P, c_int, c_int, _P]),
    "lush_tc_bcast_bench": (c_int, [_P, c_int, c_int, c_int, c_int, c_int, _P, _P]),
    "lush_debug_read_chain_trace": (c_int, [_P, c_int]),
    "lush_debug_read_chain_trace_bwd": (c_int, [_P, c_int]),
}

_lib = None
ABI_VERSION = 2
# bits of the device-side error word (include/lushnerf_b200.h)
FLAG_BAD_IMAGE_INDEX, FLAG_BAD_VIEW_INDEX, FLAG_BAD_PIXEL_INDEX = 1 << 10, 1 << 11, 1 << 12


def exported_symbols():
    """Names every include/lushnerf_b200.h declaration must resolve to (used by the CPU tests)."""
    return sorted(_SIGNATURES)


def lib():
    """Load the shared library (once).  Raises if it has not been built -- there is no fallback."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} not found: the CUDA extension is not built. Run "
                "`python -m lush_nerf_b200.build` (nvcc, sm_100a). lush_nerf_b200 has no CPU/eager fallback.")
        handle = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(handle, name)  # AttributeError if the library does not export it
            fn.restype = res
            fn.argtypes = args
        if handle.lush_abi_version() != ABI_VERSION:
            raise RuntimeError("liblushnerf_b200.so ABI version mismatch; rebuild")
        _lib = handle
    return _lib


def check(rc: int, what: str):
    if rc != 0:
        msg = lib().lush_last_error().decode(errors="replace")
        raise RuntimeError(f"{what} failed (code {rc}): {msg}")


def launch_count() -> int:
    return int(lib().lush_launch_count())


def ptr(t, name="tensor", dtype=torch.float32):
    """Device pointer of a contiguous CUDA tensor (None -> NULL)."""
    if t is None:
        return None
    if not t.is_cuda:
        raise RuntimeError(f"{name}: lush_nerf_b200 kernels need CUDA tensors (got {t.device}); "
                           "there is no CPU fallback")
    if t.dtype != dtype:
        raise RuntimeError(f"{name}: expected dtype {dtype}, got {t.dtype}")
    if not t.is_contiguous():
        raise RuntimeError(f"{name}: expected a contiguous tensor")
    return c_void_p(t.data_ptr())


def stream():
    return c_void_p(torch.cuda.current_stream().cuda_stream)
