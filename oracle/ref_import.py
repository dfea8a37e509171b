"""Import the UNMODIFIED reference (quzefan/LuSh-NeRF) from /root/reference.

TEST INFRASTRUCTURE ONLY.  This module exists so that (a) oracle/validate_against_reference.py can
pin the torch-fp32 restatement in oracle/lush_oracle.py against the reference's own code and
(b) tests/golden/make_golden.py can generate the committed golden vectors.  /root/reference does
not exist on the GPU box, so nothing in the `-m gpu` tests, smoke() or bench.py imports this file.

The reference does `import imageio` at models/lushnerf.py:8 (unused on the hot path); the module is
absent in this image, so an empty stub is registered first.
"""
import contextlib
import io
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("LUSH_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "models", "lushnerf.py"))


def import_reference():
    """Returns (models.lushnerf, utils.run_lushnerf_helpers, utils.rigid_warping) of the reference."""
    if not reference_available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    if "imageio" not in sys.modules:
        try:
            import imageio  # noqa: F401
        except Exception:
            sys.modules["imageio"] = types.ModuleType("imageio")
    # The reference uses top-level package names `models` and `utils`; make sure ours do not shadow.
    for name in ("models", "models.lushnerf", "models.mam", "utils", "utils.run_lushnerf_helpers",
                 "utils.rigid_warping"):
        mod = sys.modules.get(name)
        if mod is not None and REFERENCE_ROOT not in (getattr(mod, "__file__", "") or ""):
            del sys.modules[name]
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    with contextlib.redirect_stdout(io.StringIO()):
        import models.lushnerf as ref_model
        import utils.run_lushnerf_helpers as ref_helpers
        import utils.rigid_warping as ref_warp
    return ref_model, ref_helpers, ref_warp


def poster_args(**overrides):
    """argparse.Namespace carrying the fields NeRFAll.__init__ / raw2outputs read
    (reference models/lushnerf.py:183-231,331-333), at configs/poster_lushnerf values."""
    ns = types.SimpleNamespace(
        blur_model_type="dpnerf", multires=10, i_embed=0, use_viewdirs=True, multires_views=4,
        N_importance=64, N_samples=64, netdepth=8, netwidth=256, netdepth_fine=8, netwidth_fine=256,
        rgb_activate="sigmoid", sigma_activate="relu", tone_mapping_type="gamma",
        render_rmnearplane=80, num_images=20,
    )
    for k, v in overrides.items():
        setattr(ns, k, v)
    return ns


def build_reference_model(args=None, num_img=20, with_rbk=True):
    """Instantiate reference RBK + NeRFAll exactly as run_lushnerf.py:332-345 does."""
    ref_model, _, _ = import_reference()
    args = args or poster_args()
    rbk = None
    with contextlib.redirect_stdout(io.StringIO()):
        if with_rbk:
            rbk = ref_model.RBK(num_img=num_img, view_embed_ch=64, D_rbk=4, W_rbk=64, skips_rbk=[4],
                                num_motion_rbk=4, D_rbk_r=1, W_rbk_r=32, output_ch_rbk_r=3,
                                D_rbk_v=1, W_rbk_v=32, output_ch_rbk_v=3, D_rbk_w=1, W_rbk_w=32,
                                rbk_se_rv_window=0.1, use_dpnerf=True, rbk_use_origin=True,
                                near=0.0, far=1.0, ndc=True)
        model = ref_model.NeRFAll(args, rbk)
    return model
