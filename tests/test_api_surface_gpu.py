"""GPU: public methods of the drop-in surface that no other test calls directly, against goldens produced by the
unmodified reference (tests/golden/make_golden_api.py): NeRF / NeRF_Noise .forward on pre-embedded inputs,
NeRFAll.mlpforward / mlpforward_noise, Rigid_Blurring_Kernel.rbk_weighted_sum, SE3Field.warp vs the closed forms,
and the defined behaviour for image indices outside the embedding table."""
import os

import numpy as np
import pytest
import torch

from lush_nerf_b200 import _lib, ops
from tests.test_model_gpu import build, close

pytestmark = pytest.mark.gpu
DEV = "cuda"
HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def gapi():
    d = np.load(os.path.join(HERE, "golden", "golden_api.npz"))
    return {k: torch.from_numpy(np.asarray(d[k])) for k in d.files}


def _embedded(m, x, d):
    return torch.cat([m.embed_fn(x), m.embeddirs_fn(d)], -1)


@pytest.mark.parametrize("impl,tol", [(0, 3e-5), (1, 4e-3)])
def test_network_forward_on_embedded_inputs(gapi, det_state, impl, tol):
    """helpers:394-423 / :483-512: net(x) with x = [gamma(p) | gamma(d)].  The point is read back from the encoding's
    first columns and goes through the fused kernels (no torch-op network in the package)."""
    m = build(det_state, impl=impl).eval()
    x, d = gapi["net/x"].to(DEV), gapi["net/d"].to(DEV)
    emb = _embedded(m, x, d)
    assert emb.shape == (50, 90)
    n0 = _lib.launch_count()
    with torch.no_grad():
        out_c = m.mlp_coarse(emb)
        out_n = m.mlp_noise_coarse(emb)
    assert _lib.launch_count() > n0
    scale = gapi["net/coarse"].abs().amax(0).clamp_min(1.0)
    err = ((out_c.cpu() - gapi["net/coarse"]).abs().amax(0) / scale).max().item()
    assert out_c.shape == (50, 4) and err <= tol, err
    close(out_n, gapi["net/noise"], 3e-5, "NeRF_Noise.forward")          # the noise net is always exact fp32
    with pytest.raises(ValueError):
        m.mlp_coarse(emb[:, :60])


def test_mlpforward_and_mlpforward_noise(gapi, det_state):
    m = build(det_state, impl=0).eval()
    pts, vd = gapi["mfn/pts"].to(DEV), gapi["mfn/viewdirs"].to(DEV)
    with torch.no_grad():
        out_n = m.mlpforward_noise(pts, vd, m.mlp_noise_coarse, True, True)
        out = m.mlpforward(pts, vd, m.mlp_coarse, True, True)
    assert out_n.shape == (9, 3) and out.shape == (9, 24, 4)
    close(out_n, gapi["mfn/out"], 3e-5, "mlpforward_noise")
    scale = gapi["mf/out"].abs().amax().item()
    assert (out.cpu() - gapi["mf/out"]).abs().max().item() <= 3e-5 * max(scale, 1.0)
    with pytest.raises(IndexError):                                      # sample index 16 needs N_samples >= 17 (:271)
        m.mlpforward_noise(pts[:, :16], vd, m.mlp_noise_coarse, True, True)


def test_rbk_weighted_sum_vs_reference(gapi, det_state):
    m = build(det_state).eval()
    extras = {k: gapi[f"blend/extras/{k}"].to(DEV) for k in ("rgb0", "raw", "z_std")}
    extras["tag"] = 0                                                    # non-tensor entries pass through
    rgb, depth, acc, ex = m.mlp_rbk.rbk_weighted_sum(gapi["blend/rgb"].to(DEV), gapi["blend/depth"].to(DEV),
                                                      gapi["blend/acc"].to(DEV), extras, gapi["blend/ccw"].to(DEV))
    close(rgb, gapi["blend/out/rgb"], 1e-6, "rgb")
    close(depth, gapi["blend/out/depth"], 1e-6, "depth")
    close(acc, gapi["blend/out/acc"], 1e-6, "acc")
    for k in ("rgb0", "raw", "z_std"):
        close(ex[k], gapi[f"blend/out/extras/{k}"], 1e-6, k)
    assert ex["tag"] == 0
    # the hot path's fused blend kernel computes the same blend (gamma off)
    _, pure, _ = ops.blend_tonemap(gapi["blend/rgb"].to(DEV), gapi["blend/ccw"].to(DEV), None, 5, False)
    close(pure, gapi["blend/out/rgb"], 1e-6, "blend kernel")


def test_se3field_warp_matches_closed_form_matrices(gapi):
    """SE3Field.warp (one kernel) == applying Rigid_body.exp_se3's matrix, the way utils/rigid_warping.py:20-45 does:
    theta = |rot| + 1e-10, screw axis = (rot, trans) / theta."""
    from lush_nerf_b200.utils.rigid_warping import Rigid_body, SE3Field
    g = torch.Generator().manual_seed(2)
    pts = torch.randn(9, 3, generator=g).to(DEV)
    rot = (gapi["rigid/w"] * gapi["rigid/theta"][:, None]).to(DEV)
    trans = gapi["rigid/v"].to(DEV)
    out = SE3Field().warp(pts, rot, trans)
    rb = Rigid_body()
    theta = torch.linalg.norm(rot, dim=-1) + 1e-10
    X = rb.exp_se3(torch.cat([rot / theta[:, None], trans / theta[:, None]], -1), theta)
    ref = rb.from_homogenous((X @ rb.to_homogenous(pts)[..., None])[..., 0])
    close(out, ref, 2e-5, "warp")


def test_bad_image_index_is_defined_and_reported(det_state, capsys, monkeypatch):
    """nn.Embedding raises on an index outside the table (models/lushnerf.py:32,119).  Here: with an error word the
    ray gets the identity kernel (never uninitialised memory) and the guard raises; without one the wrapper validates."""
    monkeypatch.setenv("LUSH_NUMERIC_GUARD", "sync")
    m = build(det_state).train()
    rbk = m.mlp_rbk
    rays = torch.randn(6, 3, 2, device=DEV)
    idx = torch.tensor([[0], [3], [20], [-1], [19], [7]], device=DEV)
    with pytest.raises(IndexError):
        rbk(rays, {"images_idx": idx})                                   # no error word: host-side validation
    flags = torch.zeros(1, dtype=torch.int32, device=DEV)
    rbk.__dict__["err_flags"] = flags
    try:
        new_rays, ccw = rbk(rays, {"images_idx": idx})
    finally:
        rbk.__dict__.pop("err_flags")
    assert int(flags.item()) & _lib.FLAG_BAD_IMAGE_INDEX
    assert torch.isfinite(new_rays).all() and torch.isfinite(ccw).all()
    for bad in (2, 3):                                                   # identity kernel: all 5 rays = the original ray
        assert torch.equal(ccw[bad], torch.tensor([1., 0., 0., 0., 0.], device=DEV))
        nr = new_rays.reshape(6, 5, 3, 2)[bad]
        assert torch.equal(nr[0], rays[bad])                             # slot 0 is the ray itself
        assert torch.allclose(nr, rays[bad].expand(5, 3, 2), atol=1e-5)  # (o + d) - o with r = v = 0: rounding only
    assert not torch.equal(ccw[1], ccw[2])
    # through NeRFAll.forward the model's guard word collects the bit and the (synchronous) guard raises
    H, W, K = 40, 60, np.array([[50., 0, 30.], [0, 50., 20.], [0, 0, 1]])
    from lush_nerf_b200.synthetic import synthetic_poses
    from oracle import lush_oracle as O
    ro, rd = O.get_rays(H, W, K, torch.from_numpy(synthetic_poses(1)[0]))
    r6 = torch.stack([ro, rd], -1).reshape(-1, 3, 2)[::400][:6].contiguous().to(DEV)
    with pytest.raises(IndexError):
        m(H, W, K, chunk=4096, rays=r6, rays_info={"images_idx": idx}, retraw=True, force_naive=False,
          allkernel=False, kernel_pixel=None, perturb=1.0, N_importance=16, N_samples=24, use_viewdirs=True,
          white_bkgd=False, raw_noise_std=1.0, inference=False, near=0.0, far=1.0)


def test_uncovered_blur_kernel_topology_raises():
    """No eager path: a motion network the fused kernel cannot express is an error, not a silent torch evaluation."""
    from lush_nerf_b200.models.lushnerf import RBK
    rbk = RBK(num_img=5, view_embed_ch=16, D_rbk=30, W_rbk=16, skips_rbk=[4], num_motion_rbk=2, D_rbk_r=1, W_rbk_r=16,
              output_ch_rbk_r=3, D_rbk_v=1, W_rbk_v=16, output_ch_rbk_v=3, D_rbk_w=1, W_rbk_w=16, rbk_se_rv_window=0.1,
              use_dpnerf=True, rbk_use_origin=True, near=0.0, far=1.0, ndc=True).RBK.to(DEV)
    with pytest.raises(NotImplementedError):
        rbk.motion_params({"images_idx": torch.zeros(4, 1, dtype=torch.long, device=DEV)})
