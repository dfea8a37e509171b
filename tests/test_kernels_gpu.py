"""GPU parity tests: every CUDA kernel, called through the C ABI (ctypes), against the CPU oracle and
the golden vectors generated from the reference.  Tolerances are written next to each check:
  * sample_pdf indices / samples (given the same cdf): bit-exact
  * fp32 scan / elementwise kernels: 2e-6 abs on O(1) quantities (different association order)
  * fp32 SIMT MLP: 2e-5 abs on O(1..30) raw outputs (FFMA accumulation order vs MKL)
"""
import numpy as np
import os

import pytest
import torch

from lush_nerf_b200 import _lib, ops
from oracle import lush_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda"
NAMES = ("rgb", "density", "acc", "weights", "depth")


def close(a, b, tol, what=""):
    a, b = torch.as_tensor(a).detach().cpu().double(), torch.as_tensor(b).detach().cpu().double()
    assert a.shape == b.shape, (what, a.shape, b.shape)
    if a.numel():
        d = (a - b).abs().max().item()
        assert d <= tol, f"{what}: max|diff| {d:.3e} > {tol:g}"


def rel_l2(a, b):
    a, b = a.detach().cpu().double(), b.detach().cpu().double()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


# ------------------------------------------------------------------------------------------------ composite
def peaky_case(R, S, seed):
    g = torch.Generator().manual_seed(seed)
    raw = torch.randn(R, S, 4, generator=g)
    raw[..., 3] = torch.relu(3 * torch.randn(R, S, generator=g)) * 5
    z = torch.linspace(0, 1, S).expand(R, S) + torch.rand(R, S, generator=g) * (0.5 / S)
    rd = torch.randn(R, 3, generator=g)
    noise = torch.randn(R, S - 1, generator=g)
    return raw, z.contiguous(), rd, noise


@pytest.mark.parametrize("mode,thresh", [("train", -1.0), ("eval", 80 / 128)])
def test_composite_kat1(golden, mode, thresh):
    out = ops.composite(golden["kat1/raw"].to(DEV), golden["kat1/z"].to(DEV), golden["kat1/rays_d"].to(DEV), None, thresh)
    for n, t in zip(NAMES, out):
        close(t, golden[f"kat1/{mode}/{n}"], 1e-6, n)


def test_composite_golden_fwd_bwd(golden):
    raw = golden["r2o/raw"].to(DEV).requires_grad_(True)
    rd = golden["r2o/rays_d"].to(DEV).requires_grad_(True)
    outs = ops.composite(raw, golden["r2o/z"].to(DEV), rd, golden["r2o/noise"].to(DEV))
    for n, t in zip(NAMES, outs):
        close(t, golden[f"r2o/train/{n}"], 2e-6, n)
    sum((o * golden[f"r2o/cot/{n}"].to(DEV)).sum() for n, o in zip(NAMES, outs)).backward()
    close(raw.grad, golden["r2o/train/g_raw"], 2e-5, "g_raw")
    close(rd.grad, golden["r2o/train/g_rays_d"], 2e-5, "g_rays_d")
    outs = ops.composite(golden["r2o/raw"].to(DEV), golden["r2o/z"].to(DEV), golden["r2o/rays_d"].to(DEV), None, 80 / 128)
    for n, t in zip(NAMES, outs):
        close(t, golden[f"r2o/eval/{n}"], 2e-6, n)


@pytest.mark.parametrize("R,S", [(1, 2), (3, 17), (37, 32), (64, 33), (50, 64), (33, 128), (9, 192), (7, 256), (5, 300)])
def test_composite_vs_oracle_shapes(R, S):
    raw, z, rd, noise = peaky_case(R, S, 100 + S)
    cot = [torch.randn(R, 3), torch.randn(R, S - 1), torch.randn(R), torch.randn(R, S), torch.randn(R)]
    raw_o, rd_o = raw.clone().requires_grad_(True), rd.clone().requires_grad_(True)
    ref = O.raw2outputs(raw_o, z, rd_o, noise)
    sum((o * c).sum() for o, c in zip(ref, cot)).backward()
    raw_c, rd_c = raw.to(DEV).requires_grad_(True), rd.to(DEV).requires_grad_(True)
    out = ops.composite(raw_c, z.to(DEV), rd_c, noise.to(DEV))
    for n, a, b in zip(NAMES, out, ref):
        close(a, b, 3e-6 * max(1.0, b.abs().max().item()), n)
    sum((o * c.to(DEV)).sum() for o, c in zip(out, cot)).backward()
    assert rel_l2(raw_c.grad, raw_o.grad) < 2e-5, ("g_raw", rel_l2(raw_c.grad, raw_o.grad))
    assert rel_l2(rd_c.grad, rd_o.grad) < 2e-5, ("g_rays_d", rel_l2(rd_c.grad, rd_o.grad), rd_c.grad, rd_o.grad)


def test_composite_strided_rays_d_white_bkgd_and_flags():
    R, S = 40, 64
    raw, z, rd, noise = peaky_case(R, S, 7)
    rb = torch.zeros(R, 11)
    rb[:, 3:6] = rd
    rb_c = rb.to(DEV)
    flags = torch.zeros(1, dtype=torch.int32, device=DEV)
    out = ops.composite(raw.to(DEV), z.to(DEV), rb_c[:, 3:6], None, -1.0, True, flags)
    ref = O.raw2outputs(raw, z, rd, None, -1.0, True)
    for n, a, b in zip(NAMES, out, ref):
        close(a, b, 3e-6, n)
    assert int(flags.item()) == 0
    raw2 = raw.clone()
    raw2[3, 5, 0] = float("nan")
    raw2[4, 6, 3] = float("inf")
    ops.composite(raw2.to(DEV), z.to(DEV), rb_c[:, 3:6], None, -1.0, False, flags)
    bits = int(flags.item())
    assert bits & 1, "rgb_map nan not flagged"
    assert bits >> 8 & 1 and bits >> 8 & 2, "raw nan/inf not flagged"


def test_composite_saturated_sample_gradient_is_finite_and_exact():
    # a sample with alpha == 1 in the middle of the ray (1-alpha == 0): the division-free backward must
    # still match autograd's cumprod backward
    R, S = 4, 16
    raw, z, rd, _ = peaky_case(R, S, 3)
    raw[:, 5, 3] = 1e6
    raw_o = raw.clone().requires_grad_(True)
    ref = O.raw2outputs(raw_o, z, rd)
    (ref[0].sum() + ref[4].sum()).backward()
    raw_c = raw.to(DEV).requires_grad_(True)
    out = ops.composite(raw_c, z.to(DEV), rd.to(DEV))
    (out[0].sum() + out[4].sum()).backward()
    assert torch.isfinite(raw_c.grad).all()
    close(raw_c.grad, raw_o.grad, 2e-6, "g_raw saturated")


def test_composite_properties_full_size():
    """BASELINE stress size (65 536 rays x 128): weights sum to 1 (last alpha is 1), rgb in [0,1],
    depth within [z_0, z_last], linearity of rgb_map in a per-ray colour shift is exact in expectation."""
    R, S = 65536, 128
    g = torch.Generator(device=DEV).manual_seed(0)
    raw = torch.randn(R, S, 4, device=DEV, generator=g)
    raw[..., 3] = torch.relu(3 * torch.randn(R, S, device=DEV, generator=g)) * 5
    z = (torch.linspace(0, 1, S, device=DEV).expand(R, S) + torch.rand(R, S, device=DEV, generator=g) * (0.5 / S)).contiguous()
    rd = torch.randn(R, 3, device=DEV, generator=g)
    rgb, dens, acc, w, depth = ops.composite(raw, z, rd)
    assert (acc - 1).abs().max() < 2e-5
    assert (w >= 0).all() and (rgb >= -1e-6).all() and (rgb <= 1 + 1e-5).all()
    assert (depth >= z[:, 0] - 1e-5).all() and (depth <= z[:, -1] + 1e-5).all()
    assert torch.equal(dens, torch.relu(raw[:, :-1, 3]))
    # spot-check 256 rays against the oracle
    idx = torch.arange(0, R, R // 256)
    ref = O.raw2outputs(raw[idx].cpu(), z[idx].cpu(), rd[idx].cpu())
    close(rgb[idx], ref[0], 3e-6, "rgb")
    close(w[idx], ref[3], 3e-6, "weights")


# ------------------------------------------------------------------------------------------------ sampler
@pytest.mark.parametrize("ci", range(4))
def test_sample_pdf_strict_bit_exact_vs_reference(golden, ci):
    bins, w, cdf, u = (golden[f"spdf{ci}/{k}"].to(DEV) for k in ("bins", "weights", "cdf", "u"))
    N = u.shape[1]
    s, inds = ops.sample_pdf_cuda(bins, w, u, cdf=cdf, return_inds=True)
    assert torch.equal(inds.cpu(), golden[f"spdf{ci}/inds_rand"])
    assert torch.equal(s.cpu(), golden[f"spdf{ci}/rand"])
    ud = torch.linspace(0., 1., N).to(DEV)
    s, inds = ops.sample_pdf_cuda(bins, w, ud, cdf=cdf, return_inds=True)
    assert torch.equal(inds.cpu(), golden[f"spdf{ci}/inds_det"])
    assert torch.equal(s.cpu(), golden[f"spdf{ci}/det"])


@pytest.mark.parametrize("R,nb,N", [(64, 63, 64), (33, 31, 32), (16, 127, 128), (8, 255, 256), (5, 9, 7), (3, 2, 5), (7, 40, 100)])
@pytest.mark.parametrize("det", [True, False])
def test_sample_pdf_fused_cdf_bit_exact_vs_oracle_kernel_order(R, nb, N, det):
    g = torch.Generator().manual_seed(nb * 1000 + N)
    bins = torch.sort(torch.rand(R, nb, generator=g), -1)[0]
    w = torch.relu(torch.randn(R, nb - 1, generator=g)) ** 3
    u = torch.linspace(0., 1., N) if det else torch.rand(R, N, generator=g)
    s_ref, inds_ref, _ = O.sample_pdf(bins, w, u, order="kernel")
    s, inds = ops.sample_pdf_cuda(bins.to(DEV), w.to(DEV), u.to(DEV), return_inds=True)
    assert torch.equal(inds.cpu(), inds_ref)
    assert torch.equal(s.cpu(), s_ref)
    # and it is a rounding-level change w.r.t. the reference association order: every sample is accounted for by
    # the rule of SURVEY.md 7.3-4 (tests/test_sampler_exact_gpu.py::classify_against_torch_order)
    from tests.test_sampler_exact_gpu import classify_against_torch_order
    s_t, inds_t, cdf_t = O.sample_pdf(bins, w, u, order="torch")
    rep = classify_against_torch_order(bins, u, s.cpu(), inds.cpu(), s_t, inds_t, cdf_t)
    assert rep["unexplained"] == 0, rep
    assert (s.cpu() >= bins[:, :1] - 1e-6).all() and (s.cpu() <= bins[:, -1:] + 1e-6).all()


def test_sample_pdf_public_signature(golden):
    from lush_nerf_b200.utils.run_lushnerf_helpers import sample_pdf
    s = sample_pdf(golden["kat2/bins"].to(DEV), golden["kat2/weights"].to(DEV), 6, det=True)
    close(s, golden["kat2/det"], 1e-6, "kat2 det")
    s = sample_pdf(golden["kat2/bins"].to(DEV), golden["kat2/weights"].to(DEV), 6, det=False, pytest=True)
    close(s, golden["kat2/pytest"], 1e-6, "kat2 pytest")
    torch.manual_seed(1)
    a = sample_pdf(golden["spdf0/bins"].to(DEV), golden["spdf0/weights"].to(DEV), 64, det=False)
    torch.manual_seed(1)
    u = torch.rand([64, 64], device=DEV)
    b, _, _ = O.sample_pdf(golden["spdf0/bins"], golden["spdf0/weights"], u.cpu(), order="kernel")
    assert torch.equal(a.cpu(), b)


@pytest.mark.parametrize("R,S,N", [(50, 64, 64), (33, 32, 32), (9, 24, 16), (4, 128, 128), (3, 256, 256), (5, 17, 40)])
@pytest.mark.parametrize("det", [True, False])
def test_sample_merge_vs_oracle(R, S, N, det):
    g = torch.Generator().manual_seed(S * 7 + N)
    z = torch.sort(torch.rand(R, S, generator=g), -1)[0]
    w = torch.relu(torch.randn(R, S, generator=g)) ** 2
    u = torch.linspace(0., 1., N) if det else torch.rand(R, N, generator=g)
    z_mid = 0.5 * (z[:, 1:] + z[:, :-1])
    s_ref, _, _ = O.sample_pdf(z_mid, w[:, 1:-1], u, order="kernel")
    zf_ref = torch.sort(torch.cat([z, s_ref], -1), -1)[0]
    zf, zs, zstd = ops.sample_merge(z.to(DEV), w.to(DEV), u.to(DEV), strict=False)   # in-kernel cdf order
    assert torch.equal(zs.cpu(), s_ref)
    assert torch.equal(zf.cpu(), zf_ref)          # sorting is exact
    close(zstd, torch.std(s_ref, dim=-1, unbiased=False), 1e-6, "z_std")


def test_sample_merge_sortedness_full_size():
    R, S, N = 65536, 128, 128
    g = torch.Generator(device=DEV).manual_seed(0)
    z = torch.sort(torch.rand(R, S, device=DEV, generator=g), -1)[0]
    w = torch.relu(torch.randn(R, S, device=DEV, generator=g)) ** 3
    u = torch.rand(R, N, device=DEV, generator=g)
    zf, zs, _ = ops.sample_merge(z, w, u)
    assert (zf[:, 1:] >= zf[:, :-1]).all()
    assert torch.equal(torch.sort(torch.cat([z, zs], -1), -1)[0], zf)   # same multiset as torch.sort
    zmid_lo, zmid_hi = 0.5 * (z[:, 0] + z[:, 1]), 0.5 * (z[:, -2] + z[:, -1])
    assert (zs >= zmid_lo[:, None] - 1e-6).all() and (zs <= zmid_hi[:, None] + 1e-6).all()


# ------------------------------------------------------------------------------------------------ rays
def test_ray_prep_fwd_bwd(golden):
    H, W, f = golden["rays/HWf"].tolist()
    rays = golden["rays/rays"]
    r_o = rays.clone().requires_grad_(True)
    ref = O.ray_prep(r_o, int(H), int(W), f)
    cot = torch.randn(ref.shape)
    (ref * cot).sum().backward()
    r_c = rays.to(DEV).requires_grad_(True)
    out = ops.ray_prep(r_c, int(H), int(W), f)
    close(out[:, 0:3], golden["rays/ndc_o"], 1e-6, "ndc o vs reference")
    close(out[:, 3:6], golden["rays/ndc_d"], 1e-6, "ndc d vs reference")
    close(out, ref, 1e-6, "ray_batch")
    (out * cot.to(DEV)).sum().backward()
    assert rel_l2(r_c.grad, r_o.grad) < 1e-5
    # no-NDC branch
    out = ops.ray_prep(rays.to(DEV), int(H), int(W), f, 2.0, 6.0, False)
    close(out, O.ray_prep(rays, int(H), int(W), f, 2.0, 6.0, False), 1e-6, "no ndc")


@pytest.mark.parametrize("ndc", [True, False])
def test_ray_gen_prep_vs_oracle_get_rays(ndc):
    """Rays generated from the pose inside the preparation kernel vs the oracle's get_rays (helpers:517-528) +
    ray_prep; whole image, a pixel sub-range, a [4,4] pose.  2e-6: the NDC divisions amplify the last-bit
    differences of the three-term direction sum."""
    from lush_nerf_b200.synthetic import synthetic_poses
    H, W, f = 40, 60, 50.0
    K = np.array([[f, 0, 0.5 * W], [0, 1.1 * f, 0.45 * H], [0, 0, 1]])
    c2w = torch.from_numpy(synthetic_poses(3, seed=1)[2]).float()
    ro, rd = O.get_rays(H, W, K, c2w)
    near, far = (0.0, 1.0) if ndc else (2.0, 6.0)
    ref = O.ray_prep(torch.stack([ro, rd], -1).reshape(-1, 3, 2), H, W, f, near, far, ndc)
    out = ops.ray_gen_prep(c2w.to(DEV), H, W, K, near, far, ndc)
    close(out, ref, 2e-6, "ray_batch from pose")
    sub = ops.ray_gen_prep(c2w.to(DEV), H, W, K, near, far, ndc, pix0=777, n=1000)
    assert torch.equal(sub, out[777:1777])
    c44 = torch.eye(4)
    c44[:3] = c2w
    assert torch.equal(ops.ray_gen_prep(c44.to(DEV), H, W, K, near, far, ndc), out)
    with pytest.raises(RuntimeError):
        ops.ray_gen_prep(c2w.to(DEV), H, W, K, near, far, ndc, pix0=H * W - 5, n=6)


def test_ray_gen_gather_vs_oracle_ray_table():
    """Training-batch rays of random (view, pixel) pairs generated on the device vs the oracle's get_rays table
    (helpers:517-528) indexed the way the reference's batching does (run_lushnerf.py:538-616)."""
    from lush_nerf_b200.synthetic import synthetic_poses
    H, W, f = 40, 60, 50.0
    K = np.array([[f, 0, 0.5 * W], [0, 1.1 * f, 0.45 * H], [0, 0, 1]])
    poses = torch.from_numpy(synthetic_poses(5, seed=2)).float()
    table = torch.stack([torch.stack(O.get_rays(H, W, K, poses[v]), -1).reshape(-1, 3, 2) for v in range(5)])
    g = torch.Generator().manual_seed(4)
    view = torch.randint(0, 5, (777,), generator=g)
    pix = torch.randint(0, H * W, (777,), generator=g)
    out = ops.ray_gen_gather(poses.to(DEV), view.to(DEV), pix.to(DEV), H, W, K)
    close(out, table[view, pix], 1e-6, "rays")
    assert ops.ray_gen_gather(poses.to(DEV), view[:0].to(DEV), pix[:0].to(DEV), H, W, K).shape == (0, 3, 2)
    p44 = torch.eye(4).repeat(5, 1, 1)
    p44[:, :3] = poses
    assert torch.equal(ops.ray_gen_gather(p44.to(DEV), view.to(DEV), pix.to(DEV), H, W, K), out)
    # the LLFF [V, 3, 5] pose block (load_llff.py keeps the hwf column): row stride 5, view stride 15
    p35 = torch.cat([poses, torch.full((5, 3, 1), 7.0)], -1)
    assert torch.equal(ops.ray_gen_gather(p35.to(DEV), view.to(DEV), pix.to(DEV), H, W, K), out)
    # target colours and mask bits of the same pixels gathered from device-resident tables in the same launch
    images = torch.rand(5, H, W, 3, generator=g)
    masks = torch.rand(5, H, W, generator=g) < 0.5
    rays2, rgb, mk = ops.ray_gen_gather(poses.to(DEV), view.to(DEV), pix.to(DEV), H, W, K, images=images.to(DEV),
                                        masks=masks.to(DEV))
    assert torch.equal(rays2, out)
    assert torch.equal(rgb.cpu(), images.reshape(5, -1, 3)[view, pix])
    assert mk.dtype == torch.bool and torch.equal(mk.cpu(), masks.reshape(5, -1)[view, pix])


def test_ray_gen_gather_out_of_range_indices_are_clamped_and_flagged():
    """A view / pixel index outside its range must neither read out of bounds nor pass silently (the reference's
    fancy indexing raises): the ray of the clamped index comes back, the error word carries the bit, validate=True
    raises IndexError."""
    from lush_nerf_b200 import _lib
    from lush_nerf_b200.synthetic import synthetic_poses
    H, W, f = 40, 60, 50.0
    K = np.array([[f, 0, 0.5 * W], [0, f, 0.5 * H], [0, 0, 1]])
    poses = torch.from_numpy(synthetic_poses(3, seed=2)).float().to(DEV)
    view = torch.tensor([0, 5, -1, 2], device=DEV)
    pix = torch.tensor([10, 20, 30, H * W + 3], device=DEV)
    flags = torch.zeros(1, dtype=torch.int32, device=DEV)
    out = ops.ray_gen_gather(poses, view, pix, H, W, K, err_flags=flags)
    assert int(flags.item()) == _lib.FLAG_BAD_VIEW_INDEX | _lib.FLAG_BAD_PIXEL_INDEX
    ok = ops.ray_gen_gather(poses, torch.tensor([0, 2, 0, 2], device=DEV), torch.tensor([10, 20, 30, H * W - 1], device=DEV),
                            H, W, K)
    assert torch.equal(out, ok)
    with pytest.raises(IndexError):
        ops.ray_gen_gather(poses, view, pix, H, W, K, validate=True)
    assert ops.ray_gen_gather(poses, view[:1], pix[:1], H, W, K, validate=True).shape == (1, 3, 2)


@pytest.mark.parametrize("perturb", [False, True])
@pytest.mark.parametrize("lindisp", [False, True])
def test_stratified_z(perturb, lindisp):
    R, S = 37, 64
    rb = torch.zeros(R, 11)
    rb[:, 6] = 2.0 if lindisp else 0.0
    rb[:, 7] = 6.0 if lindisp else 1.0
    t_rand = torch.rand(R, S) if perturb else None
    ref = O.stratified_z(rb[:, 6:7], rb[:, 7:8], S, t_rand, lindisp)
    out = ops.stratified_z(rb.to(DEV), torch.linspace(0., 1., S).to(DEV), None if t_rand is None else t_rand.to(DEV), lindisp)
    close(out, ref, 0.0 if not lindisp else 5e-7, "z")


def test_rbk_warp_fwd_bwd(golden, det_state):
    _, rbk = det_state
    rays = golden["fwd/rays"]
    N = rays.shape[0]
    g = torch.Generator().manual_seed(2)
    r = (torch.randn(N, 3, 4, generator=g) * 0.05)
    v = (torch.randn(N, 3, 4, generator=g) * 0.05)
    r[0] = 0.0  # zero rotation: theta = 1e-10 path
    ro, rr, vv = rays.clone().requires_grad_(True), r.clone().requires_grad_(True), v.clone().requires_grad_(True)
    ref = O.rbk_warp(ro, rr, vv)
    cot = torch.randn(ref.shape, generator=g)
    (ref * cot).sum().backward()
    rc, rrc, vvc = rays.to(DEV).requires_grad_(True), r.to(DEV).requires_grad_(True), v.to(DEV).requires_grad_(True)
    out = ops.rbk_warp(rc, rrc, vvc)
    close(out, ref, 2e-6, "new_rays")
    (out * cot.to(DEV)).sum().backward()
    assert rel_l2(rc.grad, ro.grad) < 1e-5
    assert rel_l2(rrc.grad[1:], rr.grad[1:]) < 2e-4
    assert rel_l2(vvc.grad, vv.grad) < 2e-5
    assert torch.isfinite(rrc.grad).all()


def test_rbk_module_matches_reference_golden(golden, det_state):
    from tests.test_abi_cpu import make_model
    nets, rbk = det_state
    m = make_model()
    m.load_state_dict(O.reference_state_dict(nets, rbk), strict=False)
    m = m.to(DEV)
    new, ccw = m.mlp_rbk(golden["fwd/rays"].to(DEV), {"images_idx": golden["fwd/images_idx"].to(DEV)})
    close(new, golden["rbk/new_rays"], 3e-6, "new_rays")
    close(ccw, golden["rbk/ccw"], 2e-6, "ccw")


def motion_mlp_torch(rbk, e):
    """The blur kernel's motion network with torch ops on per-ray embeddings, as the reference evaluates it
    (models/lushnerf.py:118-147): e [rows, input_ch] -> (r, v, ccw).  Test-side definition of csrc/motion.cu."""
    import torch.nn.functional as F
    h = e
    for i, l in enumerate(rbk.view_embed_linears):
        h = F.relu(l(h))
        if i in rbk.skips:
            h = torch.cat([e, h], -1)
    hr, hv, hw = h, h, h
    for l in rbk.r_branch:
        hr = F.relu(l(hr))
    for l in rbk.v_branch:
        hv = F.relu(l(hv))
    for l in rbk.w_branch:
        hw = F.relu(l(hw))
    r = rbk.r_linear(hr) * rbk.rv_window
    v = rbk.v_linear(hv) * rbk.rv_window
    w = torch.sigmoid(rbk.w_linear(hw))
    return r, v, w / (torch.sum(w, dim=-1, keepdims=True) + 1e-10)


@pytest.mark.parametrize("cfg", [
    dict(num_img=20, N=1024, D=4, W=64, skips=[4], Db=1, Wb=32, M=4, C=64),     # poster configuration
    dict(num_img=3, N=1, D=2, W=48, skips=[0], Db=2, Wb=24, M=2, C=16),         # skip concat, deeper branches
    dict(num_img=37, N=777, D=5, W=96, skips=[2], Db=3, Wb=40, M=6, C=33),      # 5 CTAs, ragged last tile, odd widths
])
def test_rbk_motion_fused_matches_torch_layers(cfg):
    """Fused per-image motion network (csrc/motion.cu) vs the same layers applied with torch ops to the per-ray
    embeddings as the reference does (models/lushnerf.py:118-147): outputs and every parameter gradient."""
    from lush_nerf_b200.models.lushnerf import RBK
    torch.manual_seed(cfg["num_img"])
    rbk = RBK(num_img=cfg["num_img"], view_embed_ch=cfg["C"], D_rbk=cfg["D"], W_rbk=cfg["W"], skips_rbk=cfg["skips"],
              num_motion_rbk=cfg["M"], D_rbk_r=cfg["Db"], W_rbk_r=cfg["Wb"], output_ch_rbk_r=3, D_rbk_v=cfg["Db"],
              W_rbk_v=cfg["Wb"], output_ch_rbk_v=3, D_rbk_w=cfg["Db"], W_rbk_w=cfg["Wb"], rbk_se_rv_window=0.1,
              use_dpnerf=True, rbk_use_origin=True, near=0.0, far=1.0, ndc=True).RBK.to(DEV)
    with torch.no_grad():                                   # the identity-warp init would hide the r / v heads
        rbk.r_linear.weight.normal_(0, 0.3)
        rbk.v_linear.weight.normal_(0, 0.3)
    assert rbk._motion_plan() is not None
    g = torch.Generator().manual_seed(1)
    idx = torch.randint(0, cfg["num_img"], (cfg["N"], 1), generator=g).to(DEV)
    cots = None
    results = []
    for fused in (True, False):
        rbk.zero_grad(set_to_none=True)
        out = rbk.motion_params({"images_idx": idx}) if fused else \
            motion_mlp_torch(rbk, rbk.view_embedding_layer(idx.squeeze(-1)))
        if cots is None:
            cots = [torch.randn(o.shape, generator=g).to(DEV) for o in out]
        sum((o * c).sum() for o, c in zip(out, cots)).backward()
        results.append(([o.detach().clone() for o in out],
                        {n: p.grad.detach().clone() for n, p in rbk.named_parameters()}))
    (of, gf), (ot, gt) = results
    for a, b, n in zip(of, ot, ("r", "v", "ccw")):
        close(a, b, 2e-6, n)
    assert set(gf) == set(gt)
    for n in gt:
        assert rel_l2(gf[n], gt[n]) < 2e-5, (n, rel_l2(gf[n], gt[n]))


def test_new_entry_points_edge_cases():
    """Empty / degenerate inputs of the round's new entry points: no rays for the motion network, a pixel range of
    length 0, an optimiser step with no gradients, mismatched loss shapes, CPU tensors (no fallback)."""
    from lush_nerf_b200.models.lushnerf import RBK
    from lush_nerf_b200.optim import Adam
    rbk = RBK(num_img=5, view_embed_ch=16, D_rbk=2, W_rbk=32, skips_rbk=[4], num_motion_rbk=2, D_rbk_r=1, W_rbk_r=16,
              output_ch_rbk_r=3, D_rbk_v=1, W_rbk_v=16, output_ch_rbk_v=3, D_rbk_w=1, W_rbk_w=16, rbk_se_rv_window=0.1,
              use_dpnerf=True, rbk_use_origin=True, near=0.0, far=1.0, ndc=True).RBK.to(DEV)
    r, v, w = rbk.motion_params({"images_idx": torch.zeros(0, 1, dtype=torch.long, device=DEV)})
    assert r.shape == (0, 6) and v.shape == (0, 6) and w.shape == (0, 3)
    r, v, w = rbk.motion_params({"images_idx": torch.full((7, 1), 4, dtype=torch.long, device=DEV)})
    assert torch.isfinite(r).all() and torch.allclose(w.sum(-1), torch.ones(7, device=DEV), atol=1e-6)
    assert torch.equal(r[0], r[6])                                   # same image -> same motion
    K = np.array([[50.0, 0, 30.0], [0, 50.0, 20.0], [0, 0, 1]])
    assert ops.ray_gen_prep(torch.eye(4, device=DEV), 40, 60, K, n=0).shape == (0, 11)
    p = torch.nn.Parameter(torch.ones(3, device=DEV))
    opt = Adam([p], lr=0.1)
    opt.step()                                                       # no gradient anywhere: a no-op
    assert torch.equal(p.detach().cpu(), torch.ones(3)) and len(opt.state_dict()["state"]) == 0
    with pytest.raises(RuntimeError):
        Adam([torch.nn.Parameter(torch.ones(3))], lr=0.1)            # CPU parameter: no fallback
    with pytest.raises(RuntimeError):
        ops.blur_loss(torch.zeros(4, 3), torch.zeros(4, 3), torch.zeros(4, 3))
    with pytest.raises(RuntimeError):
        ops.ray_gen_prep(torch.eye(4), 40, 60, K)


def test_blend_tonemap_fwd_bwd():
    N, K = 33, 5
    g = torch.Generator().manual_seed(0)
    rgb = torch.rand(N * K, 3, generator=g)
    ccw = torch.softmax(torch.randn(N, K, generator=g), -1)
    nz = torch.randn(N, 3, generator=g)
    a, b, c = (t.clone().requires_grad_(True) for t in (rgb, ccw, nz))
    ref = O.blend_tonemap(a, b, c, K)
    cot = [torch.randn(N, 3, generator=g) for _ in range(3)]
    sum((o * t).sum() for o, t in zip(ref, cot)).backward()
    ac, bc, cc = (t.to(DEV).requires_grad_(True) for t in (rgb, ccw, nz))
    out = ops.blend_tonemap(ac, bc, cc, K, True)
    for o, r_, n in zip(out, ref, ("blur", "pure", "noise")):
        close(o, r_, 2e-6, n)
    sum((o * t.to(DEV)).sum() for o, t in zip(out, cot)).backward()
    assert rel_l2(ac.grad, a.grad) < 1e-5 and rel_l2(bc.grad, b.grad) < 1e-5 and rel_l2(cc.grad, c.grad) < 1e-5
    # K = 1 without noise = plain tone mapping
    out = ops.blend_tonemap(rgb.to(DEV), None, None, 1, True)
    close(out[1], rgb ** (1 / 2.2), 2e-6, "tm")


@pytest.mark.parametrize("N", [1, 33, 1024, 40000])
def test_blur_loss_value_and_gradient(N):
    """Fused loss kernel (one CTA for N*3 <= 32768, atomics above) vs the oracle's img2mse/img2l1 sum."""
    g = torch.Generator().manual_seed(N)
    tgt = torch.rand(N, 3, generator=g)
    blur = (tgt + 0.2 * torch.randn(N, 3, generator=g)).requires_grad_(True)
    blur0 = (tgt + 0.3 * torch.randn(N, 3, generator=g)).requires_grad_(True)
    blur.data[0] = tgt[0]                      # exact zeros of the residual: sign(0) = 0 like torch.abs' backward
    ref = O.blur_loss(blur, blur0, tgt)
    (ref * 1.7).backward()
    bc, b0c = (t.detach().to(DEV).requires_grad_(True) for t in (blur, blur0))
    out = ops.blur_loss(bc, b0c, tgt.to(DEV))
    assert out.dim() == 0
    assert abs(out.item() - ref.item()) <= 2e-6 * max(1.0, abs(ref.item()))
    (out * 1.7).backward()
    assert rel_l2(bc.grad, blur.grad) < 1e-6 and rel_l2(b0c.grad, blur0.grad) < 1e-6
    assert bc.grad[0].abs().max().item() == 0.0
    if N * 3 <= 32768:                          # single-CTA path is bit-reproducible
        assert ops.blur_loss(bc, b0c, tgt.to(DEV)).item() == out.item()
    with pytest.raises(ValueError):
        ops.blur_loss(bc, b0c[:-1] if N > 1 else b0c[:0], tgt.to(DEV))


# ------------------------------------------------------------------------------------------------ MLP (fp32 SIMT)
def mlp_case(net_name, R, S, seed):
    g = torch.Generator().manual_seed(seed)
    rb = torch.zeros(R, 11)
    rb[:, 0:3] = torch.rand(R, 3, generator=g) * 2 - 1
    rb[:, 3:6] = torch.randn(R, 3, generator=g) * 0.5
    v = torch.randn(R, 3, generator=g)
    rb[:, 8:11] = v / v.norm(dim=-1, keepdim=True)
    rb[:, 7] = 1
    z = torch.sort(torch.rand(R, S, generator=g), -1)[0]
    return rb, z


@pytest.mark.parametrize("net_name,D,W,mode", [("coarse", 8, 256, 0), ("noise", 4, 128, 1)])
@pytest.mark.parametrize("R,S", [(5, 7), (64, 24)])
def test_mlp_simt_fwd_bwd_vs_oracle(det_state, net_name, D, W, mode, R, S):
    nets, _ = det_state
    rb, z = mlp_case(net_name, R, S, 5)
    p = {k: t.clone().requires_grad_(True) for k, t in nets[net_name].items()}
    rb_o = rb.clone().requires_grad_(True)
    ref = O.mlp_on_rays(p, rb_o, z, D, W, 10, 4, 4, mode)
    cot = torch.randn(ref.shape)
    (ref * cot).sum().backward()
    desc = _lib.MlpDesc(D, W, 10, 4, 4, mode)
    params = [t.clone().to(DEV) for _, t in ((n, nets[net_name][n]) for n, _ in O.mlp_param_shapes(D, W, 63, 27, 4))]
    flat = torch.cat([t.reshape(-1) for t in params])
    views, off = [], 0
    for t in params:
        views.append(flat[off:off + t.numel()].view(t.shape).requires_grad_(True))
        off += t.numel()
    rb_c = rb.to(DEV).requires_grad_(True)
    out = ops.mlp_apply(rb_c, z.to(DEV), desc, 0, flat, views)
    close(out, ref, 3e-5 * max(1.0, ref.abs().max().item()), "mlp out")
    # inference path (no stash, chunked workspace) gives the same numbers
    with torch.no_grad():
        out2 = ops.mlp_apply(rb.to(DEV), z.to(DEV), desc, 0, flat, [v.detach() for v in views])
    close(out2, out, 1e-6, "inference == training forward")
    (out * cot.to(DEV)).sum().backward()
    for (name, _), v in zip(O.mlp_param_shapes(D, W, 63, 27, 4), views):
        gref = p[name].grad
        if gref is None:  # noise MLP: alpha_linear never receives a gradient
            assert v.grad is None or float(v.grad.abs().max()) == 0.0
            continue
        assert rel_l2(v.grad, gref) < 2e-5, name
    g_c, g_o = rb_c.grad.cpu(), rb_o.grad
    assert rel_l2(g_c[:, 0:6], g_o[:, 0:6]) < 5e-5
    assert rel_l2(g_c[:, 8:11], g_o[:, 8:11]) < 5e-5


def test_mlp_golden_outputs_through_public_mlpforward(golden, det_state):
    from tests.test_abi_cpu import make_model
    nets, rbk = det_state
    m = make_model()
    m.load_state_dict(O.reference_state_dict(nets, rbk), strict=False)
    m = m.to(DEV)
    m.mlp_coarse.mlp_impl = 0
    pts, views = golden["mlp/pts"].to(DEV), golden["mlp/views"].to(DEV)
    out = m.mlpforward(pts[:, None, :], views, m.mlp_coarse)
    close(out[:, 0, :], golden["mlp/coarse/out"], 5e-5, "NeRF.forward golden")


def test_small_mlp_big_row_tiles_bit_identical(det_state):
    """From 4096 rows on (a render chunk) the fused noise-MLP forward runs 32 rows per CTA instead of 8 (csrc/mlp_small.cu,
    SM_BIG_ROWS): same FMA order per output, so the result must equal the small-tile kernel's bit for bit -- and the
    oracle's on a slice."""
    nets, _ = det_state
    D, W, mode = 4, 128, 1
    R = 4096 + 37                      # a ragged last tile too
    rb, z = mlp_case("noise", R, 1, 11)
    desc = _lib.MlpDesc(D, W, 10, 4, 4, mode)
    shapes = O.mlp_param_shapes(D, W, 63, 27, 4)
    flat = torch.cat([nets["noise"][n].reshape(-1) for n, _ in shapes]).to(DEV)
    with torch.no_grad():
        big = ops.mlp_apply(rb.to(DEV), z.to(DEV), desc, 0, flat, [])
        parts = [ops.mlp_apply(rb[a:a + 2048].to(DEV), z[a:a + 2048].to(DEV), desc, 0, flat, []) for a in range(0, R, 2048)]
    assert torch.equal(big, torch.cat(parts, 0))
    ref = O.mlp_on_rays(nets["noise"], rb[:64], z[:64], D, W, 10, 4, 4, mode)
    close(big[:64], ref, 3e-5 * max(1.0, ref.abs().max().item()), "small mlp (32-row tiles) vs oracle")


@pytest.mark.parametrize("R,S", [(1024, 1), (77, 3)])
def test_small_mlp_fused_backward_vs_oracle(det_state, R, S):
    """Noise MLP with detached inputs (the model's call pattern): fused forward AND fused backward kernels
    (csrc/mlp_small.cu) against the oracle, and against the layer-by-layer GEMM path (LUSH_SMALL_MLP=0)."""
    import os
    nets, _ = det_state
    D, W, mode = 4, 128, 1
    rb, z = mlp_case("noise", R, S, 9)
    p = {k: t.clone().requires_grad_(True) for k, t in nets["noise"].items()}
    ref = O.mlp_on_rays(p, rb, z, D, W, 10, 4, 4, mode)
    cot = torch.randn(ref.shape, generator=torch.Generator().manual_seed(1))
    (ref * cot).sum().backward()
    desc = _lib.MlpDesc(D, W, 10, 4, 4, mode)
    shapes = O.mlp_param_shapes(D, W, 63, 27, 4)
    grads = {}
    for small in ("1", "0"):
        os.environ["LUSH_SMALL_MLP"] = small
        flat = torch.cat([nets["noise"][n].reshape(-1) for n, _ in shapes]).to(DEV)
        views, off = [], 0
        for _, shp in shapes:
            n = int(np.prod(shp)); views.append(flat[off:off + n].view(shp).requires_grad_(True)); off += n
        out = ops.mlp_apply(rb.to(DEV), z.to(DEV), desc, 0, flat, views)
        close(out, ref, 3e-5 * max(1.0, ref.abs().max().item()), "small mlp out")
        (out * cot.to(DEV)).sum().backward()
        grads[small] = [v.grad for v in views]
    os.environ.pop("LUSH_SMALL_MLP", None)
    for (name, _), g1, g0 in zip(shapes, grads["1"], grads["0"]):
        gref = p[name].grad
        if gref is None:
            assert g1 is None or float(g1.abs().max()) == 0.0
            continue
        tol = 2e-5 * max(1.0, gref.abs().max().item()) * max(1.0, (R * S) ** 0.5 / 8)
        close(g1, gref, tol, f"fused grad {name}")
        close(g1, g0, tol, f"fused vs gemm path {name}")
