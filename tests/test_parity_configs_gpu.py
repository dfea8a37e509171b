"""Parity of the BENCHMARKED path (tcgen05, fp16 operands / fp32 accumulate) on BASELINE.json's own configurations,
against the CPU oracle on identical injected draws -- direction-sensitive gradient checks, not norms.

Stated tolerances (north_star: "fp32/TF32 tolerance, e.g. max-abs 1e-3 on rgb"):
  * rendered colours (rgb_map, rgb0, acc, tone-mapped blur outputs): max-abs <= 1e-3; coarse depth <= 1e-2; FINE depth
    mean-abs <= 5e-3 and max-abs <= 8e-2: a 1e-3 change of a coarse weight can move a fine sample across a bin when u
    sits on a cdf entry, which moves depth (a sum of w_i z_i) by that weight x the bin width, while rgb (the network's
    colour barely changes over one bin) does not notice (per-sample tensors of the fine pass are not
    comparable end to end)
  * parameter / ray gradients per tensor: cosine similarity >= 0.997 AND relative L2 <= 8e-2 against the fp32 oracle
    (a 10-bit-operand forward moves pre-activations by ~3e-4 and flips the ReLU gate of the ~0.1 % of units that close
    to zero; each flip adds or removes one full gradient entry: 2-4 % per tensor is inherent to fp16 AND TF32 operands)
  * the same kernels against a torch emulation that rounds operands to fp16 at the same places (this isolates kernel
    correctness from the precision class): forward <= 5e-4 of scale; head gradients (no ReLU gate between them and the
    output) <= 5e-4 relative L2; gated tensors <= 1.5e-2 and cosine >= 0.9999 -- what is left there is fp32 summation
    order, which still flips the gate of the ~3e-5 of units whose pre-activation is within 1e-6 of zero, and
    sqrt(flipped / active) ~ 5e-3 whatever the batch size
  * fp16 range: activations pushed beyond 65504 overflow to inf / NaN in `raw` and raise the guard flag -- never a
    silently clamped finite result.
"""
import math
import warnings

import numpy as np
import pytest
import torch

from lush_nerf_b200 import _lib, ops
from oracle import lush_oracle as O
from tests.test_kernels_gpu import mlp_case
from tests.test_model_gpu import build, inject

pytestmark = pytest.mark.gpu
DEV = "cuda"
SHAPES = O.mlp_param_shapes(8, 256, 63, 27, 4)


def _cos(a, b):
    a, b = a.detach().double().cpu().reshape(-1), b.detach().double().cpu().reshape(-1)
    return (a @ b / (a.norm() * b.norm()).clamp_min(1e-300)).item()


def _rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def _grad_check(named_pairs, cos_min=0.997, rel_max=8e-2, what=""):
    worst_c, worst_r = 1.0, 0.0
    for name, g, ref in named_pairs:
        assert g is not None and torch.isfinite(g).all(), (what, name)
        if ref.norm().item() < 1e-12:
            continue
        c, r = _cos(g, ref), _rel(g, ref)
        worst_c, worst_r = min(worst_c, c), max(worst_r, r)
        assert c >= cos_min, (what, name, "cosine", c)
        assert r <= rel_max, (what, name, "rel-L2", r)
    return worst_c, worst_r


def _rays(n, H, W, f, seed):
    g = torch.Generator().manual_seed(seed)
    K = np.array([[f, 0, 0.5 * W], [0, f, 0.5 * H], [0, 0, 1]])
    c2w = torch.tensor([[0.99, 0.02, -0.1, 0.1], [-0.02, 0.999, 0.03, -0.05], [0.1, -0.03, 0.99, 0.2]])
    ro, rd = O.get_rays(H, W, K, c2w)
    pick = torch.randperm(H * W, generator=g)[:n]
    return torch.stack([ro, rd], -1).reshape(-1, 3, 2)[pick].contiguous(), K


# ------------------------------------------------------------------------------------------------------------
# BASELINE configs[0]: render_rays fwd+bwd, N_rand=256, N_samples=32, N_importance=32, perturb=1, raw_noise_std=1,
# loss = MSE(rgb_map) + MSE(rgb0)   (SURVEY.md 8d "Config 1")
# ------------------------------------------------------------------------------------------------------------
def test_config0_render_rays_fwd_bwd_tc_vs_oracle(det_state):
    nets, _ = det_state
    R, S, N = 256, 32, 32
    H, W, f = 400, 600, 500.0
    rays, _K = _rays(R, H, W, f, 1)
    g = torch.Generator().manual_seed(2)
    draws = [torch.rand(R, S, generator=g), torch.randn(R, S - 1, generator=g), torch.rand(R, N, generator=g),
             torch.randn(R, S + N - 1, generator=g)]

    # loss = MSE(rgb_map) + MSE(rgb0) as SURVEY.md 8d words it (the maps' own mean squares, like smoke()).  Against a
    # random target the per-sample sigma gradients would carry random signs and the heads' gradients -- plain sums over
    # all 8 192 coarse samples -- would cancel to ~1/90 of their terms, turning a 1e-3 forward difference into a 10-80 %
    # "error" of a sum that is itself noise (observed: coarse alpha_linear.bias 0.84, .weight 0.24 with every other
    # tensor below 5e-2).
    # ---- oracle (fp32, CPU), in-kernel cdf association order so that both sides place the same fine samples
    pn = {k: {n: t.clone().requires_grad_(True) for n, t in d.items()} for k, d in nets.items()}
    rb_o = O.ray_prep(rays, H, W, f).requires_grad_(True)
    ret_o, _ = O.render_rays(pn, O.POSTER_CFG, rb_o, N_samples=S, N_importance=N, t_rand=draws[0], noise_c=draws[1],
                             u=draws[2], noise_f=draws[3], with_noise_mlp=False, sample_order="kernel")
    loss_o = (ret_o["rgb_map"] ** 2).mean() + (ret_o["rgb0"] ** 2).mean()
    loss_o.backward()

    # ---- CUDA path, tcgen05
    m = build(det_state, impl=1).train()
    assert m.mlp_coarse.impl(True) == 1 and m.mlp_fine.impl(True) == 1
    inject(m, draws)
    rb = ops.ray_prep(rays.to(DEV), H, W, f).detach().requires_grad_(True)
    n0 = _lib.launch_count()
    ret = m.render_rays_nonoise(rb, S, retraw=True, perturb=1.0, N_importance=N, raw_noise_std=1.0)
    loss = (ret["rgb_map"] ** 2).mean() + (ret["rgb0"] ** 2).mean()
    loss.backward()
    torch.cuda.synchronize()
    assert _lib.launch_count() > n0
    bad = []
    for k, tol in (("rgb_map", 1e-3), ("rgb0", 1e-3), ("acc_map", 1e-3), ("acc0", 1e-3), ("depth0", 1e-2), ("depth_map", 8e-2)):
        d = (ret[k].detach().cpu() - ret_o[k].detach()).abs().max().item()
        if d > tol:
            bad.append(f"{k}: max-abs {d:.3e} > {tol}")
    dmean = (ret["depth_map"].detach().cpu() - ret_o["depth_map"].detach()).abs().mean().item()
    if dmean > 5e-3:
        bad.append(f"depth_map: mean-abs {dmean:.3e} > 5e-3")
    if abs(loss.item() - loss_o.item()) > 1e-3:
        bad.append(f"loss {loss.item():.6f} vs {loss_o.item():.6f}")
    pairs = []
    for net, mod in (("coarse", m.mlp_coarse), ("fine", m.mlp_fine)):
        sd = dict(mod.named_parameters())
        pairs += [(f"{net}.{n}", sd[n].grad, pn[net][n].grad) for n, _ in SHAPES]
    # ray gradients: d o, d d (columns 0..5) and d viewdirs (8..10); near / far carry none
    pairs += [("rays.d(o,d)", rb.grad[:, 0:6], rb_o.grad[:, 0:6]), ("rays.d(viewdirs)", rb.grad[:, 8:11], rb_o.grad[:, 8:11])]
    wc, wr = 1.0, 0.0
    for name, g_, ref in pairs:
        assert g_ is not None and torch.isfinite(g_).all(), name
        if ref.norm().item() < 1e-12:
            continue
        c, r = _cos(g_, ref), _rel(g_, ref)
        wc, wr = min(wc, c), max(wr, r)
        if c < 0.997 or r > 8e-2:
            bad.append(f"grad {name}: cosine {c:.5f}, rel-L2 {r:.3e}")
    gc, gr = wc, wr
    assert not bad, "config0 tc vs oracle:\n  " + "\n  ".join(bad)
    print(f"config0 tc vs oracle: worst cosine {min(wc, gc):.5f}, worst rel-L2 {max(wr, gr):.3e}")


# ------------------------------------------------------------------------------------------------------------
# BASELINE configs[1] at reduced pixel count (the CPU oracle keeps ~20 KB of fp32 activations per sample):
# NeRFAll.forward, RBK mode, 205 px x K=5 = 1025 rays x (64 + 64), blur loss, backward
# ------------------------------------------------------------------------------------------------------------
def test_config1_nerfall_rbk_tc_vs_oracle(det_state):
    from lush_nerf_b200.synthetic import build_model, poster_args
    nets, rbk = det_state
    NPX, S, N, Kr = 205, 64, 64, 5
    H, W, f = 400, 600, 500.0
    rays, K = _rays(NPX, H, W, f, 3)
    R = NPX * Kr
    g = torch.Generator().manual_seed(4)
    draws = [torch.rand(R, S, generator=g), torch.randn(R, S - 1, generator=g), torch.rand(R, N, generator=g),
             torch.randn(R, S + N - 1, generator=g)]
    target = torch.rand(NPX, 3, generator=g)
    idx = torch.randint(0, 20, (NPX, 1), generator=g)
    mask = torch.rand(NPX, generator=g) < 0.5

    pn = {k: {n: t.clone().requires_grad_(True) for n, t in d.items()} for k, d in nets.items()}
    pr = {n: t.clone().requires_grad_(True) for n, t in rbk.items()}
    out_o, _ = O.nerfall_train_forward(pn, pr, O.POSTER_CFG, rays, idx, H, W, f, N_samples=S, N_importance=N,
                                       allkernel=True, kernel_pixel=mask, t_rand=draws[0], noise_c=draws[1], u=draws[2],
                                       noise_f=draws[3], sample_order="kernel")
    loss_o = O.blur_loss(out_o[0], out_o[1], target)
    loss_o.backward()

    m = build_model(poster_args())
    m.load_state_dict(O.reference_state_dict(nets, rbk), strict=False)
    m = m.to(DEV).train()
    for net in (m.mlp_coarse, m.mlp_fine):
        net.mlp_impl = 1
    inject(m, draws)
    out = m(H, W, K, chunk=1024 * 32, rays=rays.to(DEV), rays_info={"images_idx": idx.to(DEV)}, retraw=True,
            force_naive=False, allkernel=True, kernel_pixel=mask.to(DEV), perturb=1.0, N_importance=N, N_samples=S,
            use_viewdirs=True, white_bkgd=False, raw_noise_std=1.0, inference=False, near=0.0, far=1.0)
    loss = ops.blur_loss(out[0], out[1], target.to(DEV))
    loss.backward()
    torch.cuda.synchronize()
    for i in (0, 1, 3, 5, 6):
        d = (out[i].detach().cpu() - out_o[i].detach()).abs().max().item()
        assert d <= 1e-3, f"out[{i}]: {d:.3e}"
    assert abs(loss.item() - loss_o.item()) <= 1e-3
    pairs = []
    for net, mod in (("coarse", m.mlp_coarse), ("fine", m.mlp_fine)):
        sd = dict(mod.named_parameters())
        pairs += [(f"{net}.{n}", sd[n].grad, pn[net][n].grad) for n, _ in SHAPES]
    wc, wr = _grad_check(pairs, what="config1 scene nets")
    # the noise MLP (exact fp32 kernels) and the blur-kernel network (fp32 kernels fed by tc-path ray gradients)
    sdn = dict(m.mlp_noise_coarse.named_parameters())
    noise_pairs = [(f"noise.{n}", sdn[n].grad, pn["noise"][n].grad) for n in pn["noise"] if pn["noise"][n].grad is not None
                   and sdn[n].grad is not None]
    _grad_check(noise_pairs, cos_min=0.9999, rel_max=2e-3, what="config1 noise net")
    sdr = dict(m.mlp_rbk.named_parameters())
    rbk_pairs = [(f"rbk.{n}", sdr[n].grad, pr[n].grad) for n in pr if n in sdr and pr[n].grad is not None]
    assert len(rbk_pairs) >= 10
    rc, rr = _grad_check(rbk_pairs, cos_min=0.99, rel_max=1.5e-1, what="config1 blur-kernel net")
    print(f"config1 tc vs oracle: scene nets worst cosine {wc:.5f} rel-L2 {wr:.3e}; blur-kernel net {rc:.5f} / {rr:.3e}")


# ------------------------------------------------------------------------------------------------------------
# the kernels against a torch emulation with the SAME fp16 operand roundings (promoted from tests/diagnostics)
# ------------------------------------------------------------------------------------------------------------
def _h16(t):
    return t.half().float()


def _emulate(p, rb, z, cot):
    """fp16 operands (weights, biases riding in the MMA, activations, loss-scaled dZ), fp32 accumulation, heads in fp32
    from the un-rounded activations -- the arithmetic of csrc/mlp_tc.cu / mlp_tc_bwd.cu written with torch ops."""
    R, S = z.shape
    o, d, v = rb[:, 0:3], rb[:, 3:6], rb[:, 8:11]
    x = (o[:, None, :] + d[:, None, :] * z[:, :, None]).reshape(-1, 3)
    e = _h16(O.embed(x, 10))
    ev = _h16(O.embed(v[:, None, :].expand(R, S, 3).reshape(-1, 3), 4))
    Wh = {k: _h16(t) for k, t in p.items()}
    hs, masks = [], []
    h = e
    for i in range(8):
        zz = h @ Wh[f"pts_linears.{i}.weight"].t() + Wh[f"pts_linears.{i}.bias"]
        a32 = torch.relu(zz)
        masks.append(zz > 0)
        a = _h16(a32)
        hs.append(a)
        h = torch.cat([e, a], -1) if i == 4 else a
    alpha = a32 @ p["alpha_linear.weight"].t() + p["alpha_linear.bias"]
    feat = _h16(hs[7] @ Wh["feature_linear.weight"].t() + Wh["feature_linear.bias"])
    zv = torch.cat([feat, ev], -1) @ Wh["views_linears.0.weight"].t() + Wh["views_linears.0.bias"]
    hv32 = torch.relu(zv)
    mv = zv > 0
    rgb = hv32 @ p["rgb_linear.weight"].t() + p["rgb_linear.bias"]
    out = torch.cat([rgb, alpha], -1)
    g = cot.reshape(-1, 4)
    scale = 2.0 ** math.floor(math.log2(32.0 / g.abs().max().item()))
    gs = _h16(g * scale)
    G = {"rgb_linear.weight": g[:, :3].t() @ _h16(hv32), "rgb_linear.bias": g[:, :3].sum(0),
         "alpha_linear.weight": g[:, 3:4].t() @ hs[7], "alpha_linear.bias": g[:, 3].sum().reshape(1)}
    dzv = _h16((gs[:, :3] @ Wh["rgb_linear.weight"]) * mv)
    G["views_linears.0.weight"] = dzv.t() @ torch.cat([feat, ev], -1) / scale
    G["views_linears.0.bias"] = dzv.sum(0) / scale
    gf = _h16(dzv @ Wh["views_linears.0.weight"][:, :256])
    G["feature_linear.weight"] = gf.t() @ hs[7] / scale
    G["feature_linear.bias"] = gf.sum(0) / scale
    dz = _h16((gf @ Wh["feature_linear.weight"] + gs[:, 3:4] * Wh["alpha_linear.weight"]) * masks[7])
    for l in range(7, -1, -1):
        xin = e if l == 0 else (torch.cat([e, hs[l - 1]], -1) if l == 5 else hs[l - 1])
        G[f"pts_linears.{l}.weight"] = dz.t() @ xin / scale
        G[f"pts_linears.{l}.bias"] = dz.sum(0) / scale
        if l > 0:
            Wl = Wh[f"pts_linears.{l}.weight"]
            dz = _h16((dz @ (Wl[:, 63:] if l == 5 else Wl)) * masks[l - 1])
    return out, G


@pytest.mark.parametrize("R,S,wide", [(64, 24, False), (300, 64, True), (1024, 128, True)])
def test_tc_kernels_vs_fp16_rounding_emulation(det_state, R, S, wide):
    """Same operand roundings on both sides: what is left is fp32 summation order (and the handful of ReLU gates that
    order can still flip).  Tolerances: see the module docstring."""
    nets, _ = det_state
    prev = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        desc = _lib.MlpDesc(8, 256, 10, 4, 4, 0)
        rb, z = mlp_case("coarse", R, S, 11 + R)
        g = torch.Generator().manual_seed(R)
        cot = torch.randn(R, S, 4, generator=g)
        if wide:
            cot = cot * torch.rand(R, S, 1, generator=g) ** 4
        p = {k: t.to(DEV) for k, t in nets["coarse"].items()}
        out_e, G = _emulate(p, rb.to(DEV), z.to(DEV), cot.to(DEV))
        flat = torch.cat([nets["coarse"][n].reshape(-1) for n, _ in SHAPES]).to(DEV)
        views, off = [], 0
        for _, shp in SHAPES:
            n = int(np.prod(shp))
            views.append(flat[off:off + n].view(shp).requires_grad_(True))
            off += n
        out = ops.mlp_apply(rb.to(DEV), z.to(DEV), desc, 1, flat, views)
        (out * cot.to(DEV)).sum().backward()
        torch.cuda.synchronize()
    finally:
        torch.backends.cuda.matmul.allow_tf32 = prev
    scale = out_e.abs().amax(dim=0).clamp_min(1.0)
    err = ((out.reshape(-1, 4) - out_e).abs().amax(dim=0) / scale).max().item()
    assert err <= 5e-4, f"forward vs emulation: {err:.3e} of scale"
    worst = 0.0
    for (name, _), v in zip(SHAPES, views):
        r = _rel(v.grad, G[name].reshape(v.shape))
        worst = max(worst, r)
        head = name.startswith(("alpha_linear", "rgb_linear"))
        assert r <= (5e-4 if head else 1.5e-2), (name, r)
        assert _cos(v.grad, G[name].reshape(v.shape)) >= 0.9999, name
    print(f"tc vs fp16 emulation R={R} S={S}: forward {err:.2e} of scale, worst gradient rel-L2 {worst:.2e}")


# ------------------------------------------------------------------------------------------------------------
# fp16 range: the one place fp16 operands differ from TF32
# ------------------------------------------------------------------------------------------------------------
def test_tc_fp16_overflow_is_reported_not_clamped(det_state, capsys, monkeypatch):
    """Activations pushed past 65504 (first-layer weights x 3e4): fp32 is fine there, the fp16-operand path is not.
    The result must carry inf / NaN in `raw`, and the compositing guard must print the reference's numerical-error line
    (models/lushnerf.py:474-478) -- never a finite, silently clamped picture."""
    monkeypatch.setenv("LUSH_NUMERIC_GUARD", "sync")
    nets, rbk = det_state
    big = {k: {n: t.clone() for n, t in d.items()} for k, d in nets.items()}
    for k in ("coarse", "fine"):
        big[k]["pts_linears.0.weight"] = big[k]["pts_linears.0.weight"] * 3e4
        big[k]["pts_linears.1.weight"] = big[k]["pts_linears.1.weight"] * 30.0
    H, W, f = 40, 60, 50.0
    rays, _K = _rays(64, H, W, f, 5)
    for impl, expect_overflow in ((0, False), (1, True)):
        m = build((big, rbk), impl=impl).eval()
        rb = ops.ray_prep(rays.to(DEV), H, W, f)
        capsys.readouterr()
        with torch.no_grad():
            ret, _ = m.render_rays(rb, 24, retraw=True, perturb=0.0, N_importance=16, raw_noise_std=0.0, inference=True)
        torch.cuda.synchronize()
        printed = capsys.readouterr().out
        finite = bool(torch.isfinite(ret["raw"]).all().item())
        if expect_overflow:
            assert not finite, "fp16 overflow was clamped to a finite value instead of being propagated"
            assert "[Numerical Error] raw contains" in printed and "65504" in printed, printed
        else:
            assert finite, "the exact-fp32 path must represent these activations"
            assert "Numerical Error" not in printed, printed


def test_tc_activations_up_to_fp16_max_are_exact_class(det_state):
    """Just inside the range (activations ~1e4): same 4e-3-of-scale agreement with fp32 as at O(1)."""
    nets, _ = det_state
    net = {k: v.clone() for k, v in nets["coarse"].items()}
    net["pts_linears.0.weight"] = net["pts_linears.0.weight"] * 2e3
    for i in range(1, 8):
        net[f"pts_linears.{i}.weight"] = net[f"pts_linears.{i}.weight"] * 0.5
    desc = _lib.MlpDesc(8, 256, 10, 4, 4, 0)
    rb, z = mlp_case("coarse", 64, 32, 3)
    flat = torch.cat([net[n].reshape(-1) for n, _ in SHAPES]).to(DEV)
    views, off = [], 0
    for _, shp in SHAPES:
        n = int(np.prod(shp))
        views.append(flat[off:off + n].view(shp))
        off += n
    with torch.no_grad():
        ref = ops.mlp_apply(rb.to(DEV), z.to(DEV), desc, 0, flat, views)
        out = ops.mlp_apply(rb.to(DEV), z.to(DEV), desc, 1, flat, views)
    assert torch.isfinite(out).all()
    scale = ref.abs().amax(dim=(0, 1)).clamp_min(1.0)
    assert ((out - ref).abs().amax(dim=(0, 1)) / scale).max().item() <= 4e-3


def test_simt_fallback_is_announced_once():
    """A scene network the tensor-core kernels do not cover (W = 128) runs on the fp32 SIMT kernels: one RuntimeWarning
    per topology, not silence (and not a warning per call)."""
    from lush_nerf_b200.utils import run_lushnerf_helpers as Hh
    Hh._FALLBACK_WARNED.clear()
    net = Hh.NeRF(D=4, W=128, input_ch=63, input_ch_views=27, output_ch=4, skips=[2], use_viewdirs=True).to(DEV)
    net.mlp_impl = 1
    with warnings.catch_warnings(record=True) as rec:
        warnings.simplefilter("always")
        assert net.impl(False) == 0
        assert net.impl(False) == 0
    msgs = [w for w in rec if issubclass(w.category, RuntimeWarning) and "SIMT" in str(w.message)]
    assert len(msgs) == 1, [str(w.message) for w in rec]
    big = Hh.NeRF(D=8, W=256, input_ch=63, input_ch_views=27, output_ch=4, skips=[4], use_viewdirs=True).to(DEV)
    with warnings.catch_warnings(record=True) as rec:
        warnings.simplefilter("always")
        assert big.impl(True) == 1
    assert not rec
