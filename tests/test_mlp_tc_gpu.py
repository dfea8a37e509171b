"""GPU parity of the tcgen05 (fp16-operand / fp32-accumulate) MLP path.

Stated tolerance (north_star: "fp32/TF32 tolerance, e.g. max-abs 1e-3 on rgb"): fp16 operands carry the same
11-bit significand as TF32, so
  * raw network outputs: <= 4e-3 of the output scale vs the exact-fp32 SIMT path on the same inputs
  * rendered rgb_map / rgb0 (random-init-like nets, sigma head x30): max-abs <= 1e-3, depth <= 1e-2
"""
import numpy as np
import pytest
import torch

from lush_nerf_b200 import _lib, ops
from oracle import lush_oracle as O
from tests.test_kernels_gpu import mlp_case
from tests.test_model_gpu import build

pytestmark = pytest.mark.gpu
DEV = "cuda"


def flat_of(nets, name, D=8, W=256):
    return torch.cat([nets[name][n].reshape(-1) for n, _ in O.mlp_param_shapes(D, W, 63, 27, 4)]).to(DEV)


def run_fwd(flat, rb, z, impl, desc):
    views, off = [], 0
    for _, shp in O.mlp_param_shapes(desc.D, desc.W, 63, 27, desc.skip):
        n = int(np.prod(shp))
        views.append(flat[off:off + n].view(shp))
        off += n
    with torch.no_grad():
        return ops.mlp_apply(rb, z, desc, impl, flat, views)


@pytest.mark.parametrize("R,S", [(1, 1), (2, 64), (3, 100), (64, 24), (301, 64), (40, 192)])
def test_tc_forward_matches_fp32_simt(det_state, R, S):
    nets, _ = det_state
    desc = _lib.MlpDesc(8, 256, 10, 4, 4, 0)
    assert ops.mlp_impl_supported(desc, 1)
    rb, z = mlp_case("coarse", R, S, 5 + R)
    rb, z = rb.to(DEV), z.to(DEV)
    flat = flat_of(nets, "coarse")
    ref = run_fwd(flat, rb, z, 0, desc)
    out = run_fwd(flat, rb, z, 1, desc)
    torch.cuda.synchronize()
    assert torch.isfinite(out).all()
    scale = ref.abs().amax(dim=(0, 1)).clamp_min(1.0)
    err = ((out - ref).abs().amax(dim=(0, 1)) / scale).max().item()
    assert err <= 4e-3, f"tc vs fp32 raw outputs: {err:.3e} of scale"


def test_tc_forward_other_topologies(det_state):
    """D and skip are runtime parameters of the fused kernel: D=4 no-skip and D=6 skip=2."""
    for D, skip in ((4, -1), (6, 2), (8, 4)):
        desc = _lib.MlpDesc(D, 256, 10, 4, skip if skip >= 0 else 99, 0)
        p = O.det_mlp_params(D, 256, 63, 27, skip, 50 + D)
        flat = torch.cat([p[n].reshape(-1) for n, _ in O.mlp_param_shapes(D, 256, 63, 27, skip)]).to(DEV)
        rb, z = mlp_case("coarse", 37, 32, D)
        rb, z = rb.to(DEV), z.to(DEV)
        views, off = [], 0
        for _, shp in O.mlp_param_shapes(D, 256, 63, 27, skip):
            n = int(np.prod(shp)); views.append(flat[off:off + n].view(shp)); off += n
        with torch.no_grad():
            ref = ops.mlp_apply(rb, z, desc, 0, flat, views)
            out = ops.mlp_apply(rb, z, desc, 1, flat, views)
        err = ((out - ref).abs().amax(dim=(0, 1)) / ref.abs().amax(dim=(0, 1)).clamp_min(1.0)).max().item()
        assert err <= 4e-3, (D, skip, err)


def test_tc_render_rays_eval_within_north_star_tolerance(golden, det_state):
    """End to end through render_rays (eval, det sampling): tcgen05 path vs the reference's golden outputs."""
    m = build(det_state, impl=1).eval()
    H, W, f = golden["rays/HWf"].tolist()
    rb = ops.ray_prep(golden["rays/rays"].to(DEV), int(H), int(W), f)
    with torch.no_grad():
        ret, _ = m.render_rays(rb, 24, retraw=True, perturb=0.0, N_importance=16, raw_noise_std=0.0, inference=True)
    for k, tol in (("rgb_map", 1e-3), ("rgb0", 1e-3), ("acc_map", 1e-3), ("depth_map", 1e-2), ("depth0", 1e-2)):
        d = (ret[k].cpu() - golden[f"rr/eval/{k}"]).abs().max().item()
        assert d <= tol, f"{k}: {d:.3e} > {tol}"


def test_tc_render_larger_vs_fp32_path(det_state):
    """4096 rays x (64+64): tcgen05 vs fp32 SIMT through the whole eval render path."""
    g = torch.Generator().manual_seed(0)
    H, W, f = 400, 600, 500.0
    K = np.array([[f, 0, 0.5 * W], [0, f, 0.5 * H], [0, 0, 1]])
    c2w = torch.tensor([[0.99, 0.02, -0.1, 0.1], [-0.02, 0.999, 0.03, -0.05], [0.1, -0.03, 0.99, 0.2]])
    ro, rd = O.get_rays(H, W, K, c2w)
    rays = torch.stack([ro, rd], -1).reshape(-1, 3, 2)[torch.randperm(H * W, generator=g)[:4096]].contiguous().to(DEV)
    rb = ops.ray_prep(rays, H, W, f)
    outs = {}
    for impl in (0, 1):
        m = build(det_state, impl=impl).eval()
        with torch.no_grad():
            outs[impl], _ = m.render_rays(rb, 64, perturb=0.0, N_importance=64, raw_noise_std=0.0, inference=True)
    for k, tol in (("rgb_map", 1e-3), ("rgb0", 1e-3), ("depth0", 1e-2)):
        d = (outs[1][k] - outs[0][k]).abs().max().item()
        assert d <= tol, f"{k}: {d:.3e} > {tol}"
