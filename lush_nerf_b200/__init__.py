"""lush_nerf_b200 -- B200-native (sm_100a) implementation of LuSh-NeRF's per-ray volumetric rendering
hot path behind the reference's own call surface.

    from lush_nerf_b200.models.lushnerf import NeRFAll, RBK
    from lush_nerf_b200.utils.run_lushnerf_helpers import sample_pdf, get_embedder, NeRF, ...

The numerics live in hand-written CUDA kernels (csrc/) exported through a C ABI
(include/lushnerf_b200.h, loaded with ctypes).  There is no CPU or eager fallback: importing is free,
but calling any op without the built library or without CUDA tensors raises.
"""
__version__ = "0.1.0"

from . import _lib, ops  # noqa: F401


def __getattr__(name):
    # lazy: the model modules import torch.nn and are only needed by users of the drop-in surface
    if name in ("NeRFAll", "RBK", "Rigid_Blurring_Kernel", "View_Embedding"):
        from .models import lushnerf
        return getattr(lushnerf, name)
    if name in ("NeRF", "NeRF_Noise", "sample_pdf", "get_embedder", "ToneMapping", "get_rays", "ndc_rays"):
        from .utils import run_lushnerf_helpers
        return getattr(run_lushnerf_helpers, name)
    raise AttributeError(name)
