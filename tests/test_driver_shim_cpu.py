"""CPU: the pieces of the drop-in demonstration that need no GPU (tools/run_reference_driver.py, tools/driver_shim):
the alias modules export what the reference driver's star imports expect (run_lushnerf.py:9-11 -- including torch / np /
nn, which the driver never imports itself), the configargparse stand-in reads the reference's config format, and the
synthetic scene is valid LLFF input (poses_bounds.npy rows of 17 values whose pose columns survive the loader's axis
reorder, data_utils/load_llff.py:342)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SHIM = os.path.join(ROOT, "tools", "driver_shim")


def _with_shim_path():
    saved = list(sys.path)
    gone = {m: sys.modules.pop(m) for m in list(sys.modules)
            if m.split(".")[0] in ("models", "utils", "configargparse", "imageio", "tensorboardX")}
    sys.path[:0] = [SHIM, os.path.join(SHIM, "stubs")]
    return saved, gone


def _restore(saved, gone):
    sys.path[:] = saved
    for m in list(sys.modules):
        if m.split(".")[0] in ("models", "utils", "configargparse", "imageio", "tensorboardX"):
            del sys.modules[m]
    sys.modules.update(gone)


def test_alias_modules_export_the_driver_namespace():
    saved, gone = _with_shim_path()
    try:
        ns = {}
        exec("from models.lushnerf import *\nfrom utils.run_lushnerf_helpers import *\n"
             "from utils.metrics import compute_img_metric\n", ns)
        for name in ("torch", "np", "nn", "NeRFAll", "RBK", "get_rays_np", "HALF_PIX", "to8b", "img2mse", "img2l1",
                     "mse2psnr", "get_masks_for_images", "compute_mean_with_confidence", "smart_load_state_dict",
                     "sample_pdf", "compute_img_metric"):
            assert name in ns, name
        import lush_nerf_b200.models.lushnerf as ours
        assert ns["NeRFAll"] is ours.NeRFAll
    finally:
        _restore(saved, gone)


def test_configargparse_stand_in_reads_the_reference_config_format(tmp_path):
    saved, gone = _with_shim_path()
    try:
        import configargparse
        cfg = tmp_path / "cfg"
        cfg.write_text("expname = poster\nN_rand = 1024\nscaleup-gamma = 1.0\nuse_viewdirs = True\nwhite_bkgd = False\n"
                       "raw_noise_std = 1e0\n\nuse_dpnerf\n# comment\nkernel_start_iter = 1200\n")
        p = configargparse.ArgumentParser()
        p.add_argument("--config", is_config_file=True)
        p.add_argument("--expname", type=str, required=True)
        p.add_argument("--N_rand", type=int, default=4096)
        p.add_argument("--scaleup-gamma", type=float, default=0.8)
        p.add_argument("--use_viewdirs", action="store_true")
        p.add_argument("--white_bkgd", action="store_true")
        p.add_argument("--raw_noise_std", type=float, default=0.0)
        p.add_argument("--use_dpnerf", action="store_true")
        p.add_argument("--kernel_start_iter", type=int, default=0)
        p.add_argument("--render_only", action="store_true")
        a = p.parse_args(["--config", str(cfg), "--kernel_start_iter", "5", "--render_only"])
        assert (a.expname, a.N_rand, a.scaleup_gamma, a.raw_noise_std) == ("poster", 1024, 1.0, 1.0)
        assert a.use_viewdirs and a.use_dpnerf and a.render_only and not a.white_bkgd
        assert a.kernel_start_iter == 5                  # the command line overrides the file
    finally:
        _restore(saved, gone)


def test_synthetic_scene_is_llff_input(tmp_path):
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    try:
        import run_reference_driver as T
    finally:
        sys.path.pop(0)
    from lush_nerf_b200.synthetic import synthetic_poses
    T.write_scene(str(tmp_path), V=4, H=40, W=60, focal=50.0)
    arr = np.load(tmp_path / "poses_bounds.npy")
    assert arr.shape == (4, 17)
    poses = arr[:, :-2].reshape(-1, 3, 5)
    # the loader's reorder (load_llff.py:342): (col1, -col0, col2, ...) must give back the camera-to-world columns
    rec = np.concatenate([poses[:, :, 1:2], -poses[:, :, 0:1], poses[:, :, 2:]], 2)
    assert np.allclose(rec[:, :, :4], synthetic_poses(4, 0), atol=1e-6)
    assert np.allclose(rec[:, :, 4], [40, 60, 50.0])
    assert sorted(os.listdir(tmp_path / "images_1")) == [f"{i:03d}.png" for i in range(4)]
    import cv2
    assert cv2.imread(str(tmp_path / "images_1" / "000.png")).shape == (40, 60, 3)
