"""CPU: the oracle (oracle/lush_oracle.py) against the golden vectors generated from the reference
itself (tests/golden/make_golden.py).  Tolerances: bit-exact where the oracle issues the same torch ops
on the same machine type; 2e-6 elsewhere to absorb BLAS/vector-width differences between hosts."""
import numpy as np
import pytest
import torch

from oracle import lush_oracle as O

TOL = 2e-6
NAMES = ("rgb", "density", "acc", "weights", "depth")


def close(a, b, tol=TOL):
    a, b = torch.as_tensor(a), torch.as_tensor(b)
    assert a.shape == b.shape, (a.shape, b.shape)
    if a.numel():
        d = (a.double() - b.double()).abs().max().item()
        assert d <= tol, f"max|diff| {d:.3e} > {tol:g}"


def test_kat_values_from_survey(golden):
    # the known-answer numbers printed in SURVEY.md 8c
    close(golden["kat1/train/weights"], torch.tensor([[0.39346933, 0.38340050, 0.08779488, 0.13533530]]), 1e-7)
    close(golden["kat1/train/rgb"], torch.tensor([[0.63209468, 0.41815633, 0.51689953]]), 1e-7)
    close(golden["kat1/eval/weights"], torch.tensor([[0, 0, 0.39346933, 0.60653067]]), 1e-7)
    close(golden["kat2/det"], torch.tensor([[0.1, 0.53332877, 0.59999698, 0.66666520, 0.76666576, 0.89999998]]), 1e-7)
    assert golden["kat4/param_counts"].tolist() == [595844, 595844, 84548, 25117, 1301353]


@pytest.mark.parametrize("mode,thresh", [("train", -1.0), ("eval", 80 / 128)])
def test_kat1_raw2outputs(golden, mode, thresh):
    out = O.raw2outputs(golden["kat1/raw"], golden["kat1/z"], golden["kat1/rays_d"], None, thresh)
    for n, t in zip(NAMES, out):
        close(t, golden[f"kat1/{mode}/{n}"], 1e-7)


def test_kat2_sample_pdf(golden):
    s, _, _ = O.sample_pdf(golden["kat2/bins"], golden["kat2/weights"], torch.linspace(0., 1., 6))
    close(s, golden["kat2/det"], 0.0)
    np.random.seed(0)
    u = torch.tensor(np.random.rand(1, 6), dtype=torch.float32)
    s, _, _ = O.sample_pdf(golden["kat2/bins"], golden["kat2/weights"], u)
    close(s, golden["kat2/pytest"], 0.0)


def test_kat3_embed(golden):
    close(O.embed(torch.tensor([[0.1, 0.2, 0.3]]), 10), golden["kat3/embed"], 1e-7)


def test_raw2outputs_with_grads(golden):
    raw = golden["r2o/raw"].clone().requires_grad_(True)
    rd = golden["r2o/rays_d"].clone().requires_grad_(True)
    outs = O.raw2outputs(raw, golden["r2o/z"], rd, golden["r2o/noise"])
    for n, t in zip(NAMES, outs):
        close(t, golden[f"r2o/train/{n}"])
    sum((o * golden[f"r2o/cot/{n}"]).sum() for n, o in zip(NAMES, outs)).backward()
    close(raw.grad, golden["r2o/train/g_raw"], 1e-5)
    close(rd.grad, golden["r2o/train/g_rays_d"], 1e-5)
    outs = O.raw2outputs(golden["r2o/raw"], golden["r2o/z"], golden["r2o/rays_d"], None, 80 / 128)
    for n, t in zip(NAMES, outs):
        close(t, golden[f"r2o/eval/{n}"])


@pytest.mark.parametrize("ci", range(4))
def test_sample_pdf_bit_exact(golden, ci):
    bins, w, u = golden[f"spdf{ci}/bins"], golden[f"spdf{ci}/weights"], golden[f"spdf{ci}/u"]
    N = u.shape[1]
    s, inds, cdf = O.sample_pdf(bins, w, torch.linspace(0., 1., N))
    assert torch.equal(inds, golden[f"spdf{ci}/inds_det"])
    close(s, golden[f"spdf{ci}/det"], 0.0)
    s, inds, cdf = O.sample_pdf(bins, w, u)
    assert torch.equal(inds, golden[f"spdf{ci}/inds_rand"])
    close(s, golden[f"spdf{ci}/rand"], 0.0)
    # strict mode: consuming the reference cdf reproduces the reference exactly
    s2, inds2, _ = O.sample_pdf(bins, w, u, cdf=golden[f"spdf{ci}/cdf"])
    assert torch.equal(inds2, golden[f"spdf{ci}/inds_rand"]) and torch.equal(s2, golden[f"spdf{ci}/rand"])


@pytest.mark.parametrize("ci", range(4))
def test_sample_pdf_kernel_order_is_a_rounding_level_change(golden, ci):
    """The fixed association order used in-kernel moves cdf entries by <= 4 ulp; indices may flip only
    where u sits within that distance of a cdf entry and the sample moves by <= 1e-5 there."""
    bins, w, u = golden[f"spdf{ci}/bins"], golden[f"spdf{ci}/weights"], golden[f"spdf{ci}/u"]
    s, inds, cdf = O.sample_pdf(bins, w, u, order="kernel")
    ref_cdf = golden[f"spdf{ci}/cdf"]
    assert (cdf - ref_cdf).abs().max() <= 5e-7
    flip = inds != golden[f"spdf{ci}/inds_rand"]
    if flip.any():
        near = torch.gather(ref_cdf, 1, torch.minimum(inds, golden[f"spdf{ci}/inds_rand"]).clamp(max=ref_cdf.shape[1] - 1))
        assert ((u - near).abs()[flip] <= 1e-6).all()
    close(s, golden[f"spdf{ci}/rand"], 2e-5)


def test_mlp_forward_and_grads(golden, det_state):
    nets, _ = det_state
    x, v = golden["mlp/pts"], golden["mlp/views"]
    xe = torch.cat([O.embed(x, 10), O.embed(v, 4)], -1).requires_grad_(True)
    for name, (D, W, mode) in (("coarse", (8, 256, 0)), ("noise", (4, 128, 1))):
        p = {k: t.clone().requires_grad_(True) for k, t in nets[name].items()}
        out = O.mlp_forward(xe, p, D, W, 63, 27, 4, mode)
        close(out, golden[f"mlp/{name}/out"])
        c = O.det_uniform(tuple(out.shape), 700)
        (out * c).sum().backward()
        for j, (pn, _) in enumerate(O.mlp_param_shapes(D, W, 63, 27, 4)):
            g = p[pn].grad if p[pn].grad is not None else torch.zeros_like(p[pn])
            close(g.reshape(-1)[:8], golden[f"mlp/{name}/grad/{pn}/head"], 1e-5)
            dig = golden[f"mlp/{name}/grad/{pn}/digest"]
            assert abs(g.double().norm().item() - dig[1].item()) <= 1e-5 * max(1.0, dig[1].item())


def test_ray_prep(golden):
    H, W, f = golden["rays/HWf"].tolist()
    rb = O.ray_prep(golden["rays/rays"], int(H), int(W), f)
    close(rb[:, 0:3], golden["rays/ndc_o"], 1e-7)
    close(rb[:, 3:6], golden["rays/ndc_d"], 1e-7)
    close(rb[:, 8:11], golden["rays/viewdirs"], 1e-7)
    assert torch.all(rb[:, 6] == 0) and torch.all(rb[:, 7] == 1)


KEYS = ("rgb_map", "depth_map", "acc_map", "density_map", "raw", "rgb0", "depth0", "acc0", "density0", "z_std")


def test_render_rays_eval_and_train(golden, det_state):
    nets, _ = det_state
    H, W, f = golden["rays/HWf"].tolist()
    rb = O.ray_prep(golden["rays/rays"], int(H), int(W), f)
    with torch.no_grad():
        ret, retn = O.render_rays(nets, O.POSTER_CFG, rb, N_samples=24, N_importance=16,
                                  u=torch.linspace(0, 1, 16), rm_near_thresh=80 / 128)
    for k in KEYS:
        close(ret[k], golden[f"rr/eval/{k}"], 5e-5 if k in ("raw", "density_map", "density0") else 1e-5)
    close(retn["rgb_map"], golden["rr/eval/noise_rgb"], 1e-5)
    with torch.no_grad():
        ret, _ = O.render_rays(nets, O.POSTER_CFG, rb, N_samples=24, N_importance=16,
                               t_rand=golden["rr/train/t_rand"], noise_c=golden["rr/train/noise_c"],
                               noise_f=golden["rr/train/noise_f"], u=golden["rr/train/u"])
    for k in KEYS:
        close(ret[k], golden[f"rr/train/{k}"], 5e-5 if k in ("raw", "density_map", "density0") else 1e-5)


def test_rbk_forward(golden, det_state):
    _, rbk = det_state
    new, ccw = O.rbk_forward(rbk, golden["fwd/rays"], golden["fwd/images_idx"])
    close(new, golden["rbk/new_rays"], 1e-6)
    close(ccw, golden["rbk/ccw"], 1e-6)


@pytest.mark.parametrize("mode", ["rbk", "naive"])
def test_nerfall_train_forward(golden, det_state, mode):
    nets, rbk = det_state
    H, W, f = golden["rays/HWf"].tolist()
    pn = {k: {n: t.clone().requires_grad_(True) for n, t in d.items()} for k, d in nets.items()}
    pr = {n: t.clone().requires_grad_(True) for n, t in rbk.items()}
    rays = golden["fwd/rays"].clone().requires_grad_(True)
    out, _ = O.nerfall_train_forward(
        pn, pr, O.POSTER_CFG, rays, golden["fwd/images_idx"], int(H), int(W), f, N_samples=24, N_importance=16,
        force_naive=(mode == "naive"), allkernel=True, kernel_pixel=golden["fwd/kernel_pixel"],
        t_rand=golden[f"fwd/{mode}/t_rand"], noise_c=golden[f"fwd/{mode}/noise_c"],
        noise_f=golden[f"fwd/{mode}/noise_f"], u=golden[f"fwd/{mode}/u"])
    for i in (0, 1, 3) + ((5, 6) if mode == "rbk" else ()):
        close(out[i], golden[f"fwd/{mode}/out{i}"], 1e-5)
    tgt = golden["fwd/target"]
    loss = ((out[0] - tgt) ** 2).mean() + ((out[1] - tgt) ** 2).mean() + (out[0] - tgt).abs().mean() \
        + (out[1] - tgt).abs().mean()
    close(loss, golden[f"fwd/{mode}/loss"], 1e-5)
    loss.backward()
    if mode == "rbk":
        close(rays.grad, golden["fwd/rbk/g_rays"], 1e-5)
    for net, prefix in (("coarse", "mlp_coarse"), ("fine", "mlp_fine")):
        for pname in ("pts_linears.0.weight", "pts_linears.5.weight", "rgb_linear.weight", "alpha_linear.bias"):
            g = pn[net][pname].grad
            dig = golden[f"fwd/{mode}/grad/{prefix}.{pname}/digest"]
            assert abs(g.double().norm().item() - dig[1].item()) <= 2e-4 * max(1e-6, dig[1].item()) + 1e-9
