"""GPU parity of the tcgen05 (fp16-operand / fp32-accumulate) MLP path.

Stated tolerance (north_star: "fp32/TF32 tolerance, e.g. max-abs 1e-3 on rgb"): fp16 operands carry the same
11-bit significand as TF32, so
  * raw network outputs: <= 4e-3 of the output scale vs the exact-fp32 SIMT path on the same inputs
  * rendered rgb_map / rgb0 (random-init-like nets, sigma head x30): max-abs <= 1e-3, depth <= 1e-2
"""
import ctypes

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


def test_no_grad_forward_runs_the_inference_kernel(det_state):
    """Under torch.no_grad() (rendering) the parameters still `require_grad`: the forward must nevertheless run the
    inference kernel -- no 4.7 KB/sample activation stash allocated or written -- and return the training-mode values
    bit for bit (ops._MLP.forward: the CALLER's grad mode decides, ctx.needs_input_grad does not know about no_grad)."""
    nets, _ = det_state
    desc = _lib.MlpDesc(8, 256, 10, 4, 4, 0)
    R, S = 2048, 64
    rb, z = mlp_case("coarse", R, S, 3)
    rb, z = rb.to(DEV), z.to(DEV)
    flat = flat_of(nets, "coarse")
    views, off = [], 0
    for _, shp in O.mlp_param_shapes(desc.D, desc.W, 63, 27, desc.skip):
        n = int(np.prod(shp))
        views.append(flat[off:off + n].view(shp).requires_grad_(True))
        off += n
    stash_bytes = _lib.lib().lush_mlp_stash_bytes(ctypes.byref(desc), R, S, 1)
    assert stash_bytes > 500 * 2 ** 20
    out_train = ops.mlp_apply(rb, z, desc, 1, flat, views)          # grad mode on: stash kept for backward
    assert out_train.requires_grad
    ref = out_train.detach().clone()
    del out_train
    torch.cuda.synchronize()
    torch.cuda.empty_cache()
    torch.cuda.reset_peak_memory_stats()
    base = torch.cuda.memory_allocated()
    with torch.no_grad():
        out = ops.mlp_apply(rb, z, desc, 1, flat, views)
    torch.cuda.synchronize()
    assert not out.requires_grad
    assert torch.cuda.max_memory_allocated() - base < stash_bytes // 4, "a stash was allocated under no_grad"
    assert torch.equal(out, ref)


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


def _views_of(flat, D=8, skip=4, grad=False):
    views, off = [], 0
    for _, shp in O.mlp_param_shapes(D, 256, 63, 27, skip):
        n = int(np.prod(shp))
        v = flat[off:off + n].view(shp)
        views.append(v.requires_grad_(True) if grad else v)
        off += n
    return views


def _rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def _run_train(nets_coarse, rb, z, cot, impl, need_in, desc):
    flat = torch.cat([nets_coarse[n].reshape(-1) for n, _ in O.mlp_param_shapes(8, 256, 63, 27, 4)]).to(DEV)
    views = _views_of(flat, grad=True)
    rbc = rb.to(DEV).requires_grad_(need_in)
    out = ops.mlp_apply(rbc, z.to(DEV), desc, impl, flat, views)
    (out * cot.to(DEV)).sum().backward()
    torch.cuda.synchronize()
    return out.detach(), [v.grad for v in views], rbc.grad


@pytest.mark.parametrize("R,S,need_in", [(3, 50, True), (64, 24, True), (300, 64, False), (129, 128, True)])
def test_tc_backward_exact_in_the_all_active_regime(det_state, R, S, need_in):
    """Kernel verification without ReLU-gate ambiguity: biases shifted so that every unit is active, the
    network is then linear in its weights and the tcgen05 forward+dgrad+wgrad must match the exact-fp32 SIMT
    path up to fp16 operand rounding (ten successive fp16 roundings of dZ reach the first layer): 8e-3 relative
    L2 per parameter tensor (2e-3 typical), 1e-2 for ray gradients."""
    nets, _ = det_state
    desc = _lib.MlpDesc(8, 256, 10, 4, 4, 0)
    assert ops.mlp_impl_supported(desc, 1, training=True)
    net = {k: v.clone() for k, v in nets["coarse"].items()}
    for k in net:
        if k.endswith("bias") and (k.startswith("pts_linears") or k.startswith("views_linears")):
            net[k] = net[k] + 8.0
        if k.endswith("weight") and k.startswith("pts_linears"):
            net[k] = net[k] * 0.25     # keep activations O(10) so fp16 range is not the issue under test
    rb, z = mlp_case("coarse", R, S, 11 + R)
    g = torch.Generator().manual_seed(R)
    cot = torch.randn(R, S, 4, generator=g) * torch.rand(R, S, 1, generator=g) ** 4   # wide dynamic range
    r0 = _run_train(net, rb, z, cot, 0, need_in, desc)
    r1 = _run_train(net, rb, z, cot, 1, need_in, desc)
    scale = r0[0].abs().amax(dim=(0, 1)).clamp_min(1.0)
    assert ((r1[0] - r0[0]).abs().amax(dim=(0, 1)) / scale).max().item() <= 4e-3
    for (name, _), g1, g0 in zip(O.mlp_param_shapes(8, 256, 63, 27, 4), r1[1], r0[1]):
        assert g1 is not None and torch.isfinite(g1).all(), name
        assert _rel(g1, g0) <= 1.5e-2, (name, _rel(g1, g0))
    if need_in:
        assert _rel(r1[2][:, 0:6], r0[2][:, 0:6]) <= 1e-2, ("d o,d", _rel(r1[2][:, 0:6], r0[2][:, 0:6]))
        assert _rel(r1[2][:, 8:11], r0[2][:, 8:11]) <= 1e-2, ("d viewdirs", _rel(r1[2][:, 8:11], r0[2][:, 8:11]))


@pytest.mark.parametrize("R,S,need_in", [(64, 24, True), (300, 64, True), (1024, 128, False)])
def test_tc_backward_vs_fp32_general(det_state, R, S, need_in):
    """General (ReLU) regime vs exact fp32.  A reduced-precision forward (fp16 or TF32 operands alike) moves
    pre-activations by ~3e-4, which flips the ReLU gate of the ~0.1% of units that sit that close to zero; each
    flip adds/removes one full gradient entry, so per-tensor relative L2 deviations of 2-4% are inherent
    (tests/diagnostics/diag_tc_emul.py shows tensors without flips agree to 1e-4).  Stated tolerance: 8e-2 relative L2 and
    cosine similarity >= 0.997 per parameter tensor; heads (no gate in between) 2e-3."""
    nets, _ = det_state
    desc = _lib.MlpDesc(8, 256, 10, 4, 4, 0)
    rb, z = mlp_case("coarse", R, S, 11 + R)
    g = torch.Generator().manual_seed(R)
    cot = torch.randn(R, S, 4, generator=g) * torch.rand(R, S, 1, generator=g) ** 4
    r0 = _run_train(nets["coarse"], rb, z, cot, 0, need_in, desc)
    r1 = _run_train(nets["coarse"], rb, z, cot, 1, need_in, desc)
    for (name, _), g1, g0 in zip(O.mlp_param_shapes(8, 256, 63, 27, 4), r1[1], r0[1]):
        tol = 2e-3 if name.startswith(("alpha_linear", "rgb_linear")) else 8e-2
        assert _rel(g1, g0) <= tol, (name, _rel(g1, g0))
        cos = torch.nn.functional.cosine_similarity(g1.reshape(1, -1).double(), g0.reshape(1, -1).double()).item()
        assert cos >= 0.997, (name, cos)
    if need_in:
        assert _rel(r1[2][:, 0:6], r0[2][:, 0:6]) <= 8e-2
        assert _rel(r1[2][:, 8:11], r0[2][:, 8:11]) <= 8e-2


def test_tc_training_step_vs_reference_golden(golden, det_state):
    """NeRFAll.forward (RBK mode) on the tcgen05 path vs the reference's golden outputs and loss."""
    from tests.test_model_gpu import KW, inject, K_of
    m = build(det_state, impl=1).train()
    H, W, K = K_of(golden)
    inject(m, [golden[f"fwd/rbk/{k}"] for k in ("t_rand", "noise_c", "u", "noise_f")])
    rays = golden["fwd/rays"].to(DEV).requires_grad_(True)
    out = m(H, W, K, chunk=1024, rays=rays, rays_info={"images_idx": golden["fwd/images_idx"].to(DEV)},
            force_naive=False, allkernel=True, kernel_pixel=golden["fwd/kernel_pixel"].to(DEV), **KW)
    for i in (0, 1, 5, 6):
        d = (out[i].cpu() - golden[f"fwd/rbk/out{i}"]).abs().max().item()
        assert d <= 1e-3, f"out{i}: {d:.3e}"
    tgt = golden["fwd/target"].to(DEV)
    loss = ((out[0] - tgt) ** 2).mean() + ((out[1] - tgt) ** 2).mean() + (out[0] - tgt).abs().mean() + (out[1] - tgt).abs().mean()
    assert abs(loss.item() - golden["fwd/rbk/loss"].item()) <= 1e-3
    loss.backward()
    worst = 0.0
    for name, p in m.named_parameters():
        if name.startswith(("blur_kernel_net", "dbk_view_embedding")) or p.grad is None:
            continue
        dig = golden[f"fwd/rbk/grad/{name}/digest"].double()
        if dig[1].item() < 1e-9:
            continue
        dev = abs(p.grad.double().norm().item() - dig[1].item()) / dig[1].item()
        worst = max(worst, dev)
        assert dev <= 8e-2, (name, dev)
    gr = golden["fwd/rbk/g_rays"]
    assert ((rays.grad.cpu() - gr).norm() / gr.norm()).item() <= 1e-1
    print(f"tc training step: worst |grad| norm deviation vs reference {worst:.2e}")


def test_tc_many_tiles_per_cta_forward_and_backward():
    """Sizes with >10 tiles per persistent CTA (barrier-parity bookkeeping across tiles) and a ragged last tile."""
    from oracle import lush_oracle as O
    R, S = 1500, 127                                  # 1489 tiles of 128 samples on 148 CTAs, last tile partial
    nets = O.det_nets()
    shapes = O.mlp_param_shapes(8, 256, 63, 27, 4)
    flat = torch.cat([nets["coarse"][n].reshape(-1) for n, _ in shapes]).cuda()
    desc = _lib.MlpDesc(8, 256, 10, 4, 4, 0)
    g = torch.Generator(device="cuda").manual_seed(3)
    rb = torch.zeros(R, 11, device="cuda")
    rb[:, :3] = torch.rand(R, 3, device="cuda", generator=g) * 2 - 1
    rb[:, 3:6] = torch.randn(R, 3, device="cuda", generator=g) * 0.3
    rb[:, 8:11] = torch.nn.functional.normalize(torch.randn(R, 3, device="cuda", generator=g), dim=-1)
    z = torch.rand(R, S, device="cuda", generator=g).sort(-1)[0]
    gout = torch.randn(R, S, 4, device="cuda", generator=g)
    res = []
    for impl in (0, 1):
        views, off = [], 0
        fl = flat.clone()
        for _, shp in shapes:
            n = int(np.prod(shp)); views.append(fl[off:off + n].view(shp).requires_grad_(True)); off += n
        rbg = rb.clone().requires_grad_(True)
        out = ops.mlp_apply(rbg, z, desc, impl, fl, views)
        out.backward(gout)
        res.append((out.detach(), [v.grad for v in views], rbg.grad))
    scale = res[0][0].abs().amax().item()
    assert (res[1][0] - res[0][0]).abs().max().item() / scale <= 4e-3
    for (name, _), g1, g0 in zip(shapes, res[1][1], res[0][1]):
        tol = 2e-3 if name.startswith(("alpha_linear", "rgb_linear")) else 8e-2
        assert _rel(g1, g0) <= tol, (name, _rel(g1, g0))
    assert _rel(res[1][2][:, 0:6], res[0][2][:, 0:6]) <= 8e-2


@pytest.mark.parametrize("D,skip,need_in", [(4, -1, True), (6, 2, True), (6, 2, False), (2, -1, True)])
def test_tc_backward_other_topologies(D, skip, need_in):
    """The dgrad chain and the wgrad job list are built from (D, skip) at run time; compare with the fp32 path."""
    shapes = O.mlp_param_shapes(D, 256, 63, 27, skip)
    p = O.det_mlp_params(D, 256, 63, 27, skip, 70 + D)
    desc = _lib.MlpDesc(D, 256, 10, 4, skip if skip >= 0 else 99, 0)
    rb, z = mlp_case("coarse", 300, 48, D)
    gout = torch.randn(300, 48, 4, generator=torch.Generator().manual_seed(D)).to(DEV)
    res = []
    for impl in (0, 1):
        flat = torch.cat([p[n].reshape(-1) for n, _ in shapes]).to(DEV)
        views, off = [], 0
        for _, shp in shapes:
            n = int(np.prod(shp)); views.append(flat[off:off + n].view(shp).requires_grad_(True)); off += n
        rbg = rb.to(DEV).requires_grad_(need_in)
        out = ops.mlp_apply(rbg, z.to(DEV), desc, impl, flat, views)
        out.backward(gout)
        res.append((out.detach(), [v.grad for v in views], rbg.grad if need_in else None))
    scale = res[0][0].abs().amax().item()
    assert (res[1][0] - res[0][0]).abs().max().item() / scale <= 4e-3
    for (name, _), g1, g0 in zip(shapes, res[1][1], res[0][1]):
        tol = 2e-3 if name.startswith(("alpha_linear", "rgb_linear")) else 8e-2
        assert _rel(g1, g0) <= tol, (D, skip, name, _rel(g1, g0))
    if need_in:
        assert _rel(res[1][2][:, 0:6], res[0][2][:, 0:6]) <= 8e-2


def test_tc_full_size_row_independence_and_gradient_additivity():
    """BASELINE configs[1] size (5120 rays x 128 fine samples = 655 360 samples, 35 tiles per CTA), no oracle in the
    loop: (i) rows are independent -- the batch evaluated in reversed ray order gives the reversed outputs BIT FOR BIT
    (same per-row arithmetic whatever tile, CTA or accumulator buffer a row lands in); (ii) parameter gradients are
    additive over the batch -- the gradient of the whole batch equals the sum of the gradients of its two halves up to
    fp16 rounding under (different) loss scales and fp32 atomic ordering."""
    R, S = 5120, 128
    nets = O.det_nets()
    shapes = O.mlp_param_shapes(8, 256, 63, 27, 4)
    flat = torch.cat([nets["fine"][n].reshape(-1) for n, _ in shapes]).cuda()
    desc = _lib.MlpDesc(8, 256, 10, 4, 4, 0)
    g = torch.Generator(device="cuda").manual_seed(9)
    rb = torch.zeros(R, 11, device="cuda")
    rb[:, :3] = torch.rand(R, 3, device="cuda", generator=g) * 2 - 1
    rb[:, 3:6] = torch.randn(R, 3, device="cuda", generator=g) * 0.3
    rb[:, 8:11] = torch.nn.functional.normalize(torch.randn(R, 3, device="cuda", generator=g), dim=-1)
    z = torch.rand(R, S, device="cuda", generator=g).sort(-1)[0]
    gout = torch.randn(R, S, 4, device="cuda", generator=g)

    def run(rows, with_grad):
        views, off = [], 0
        fl = flat.clone()
        for _, shp in shapes:
            n = int(np.prod(shp)); views.append(fl[off:off + n].view(shp).requires_grad_(with_grad)); off += n
        out = ops.mlp_apply(rb[rows].contiguous(), z[rows].contiguous(), desc, 1, fl, views)
        if with_grad:
            out.backward(gout[rows].contiguous())
        return out.detach(), [v.grad for v in views]

    everything = torch.arange(R, device="cuda")
    with torch.no_grad():
        fwd, _ = run(everything, False)
        rev, _ = run(everything.flip(0), False)
    assert torch.isfinite(fwd).all()
    assert torch.equal(rev.flip(0), fwd)
    out_t, g_all = run(everything, True)
    assert torch.equal(out_t, fwd)                       # the training kernel (stash written) computes the same values
    _, g_a = run(everything[:R // 2], True)
    _, g_b = run(everything[R // 2:], True)
    for (name, _), ga, gb, gw in zip(shapes, g_a, g_b, g_all):
        assert _rel(ga + gb, gw) <= 5e-3, (name, _rel(ga + gb, gw))
