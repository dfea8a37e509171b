"""Generate tests/golden/golden_v1.npz by running the UNMODIFIED reference (/root/reference) on CPU.

    python -m tests.golden.make_golden

The reference has no golden vectors of its own, so these fixtures are the pin: inputs + reference outputs
(+ selected reference autograd gradients).  Network weights are NOT stored: they are regenerated
bit-exactly from oracle.lush_oracle.det_nets / det_rbk_params (counter-based integer hash).
Random draws of the reference (torch.rand / randn under a seed) are replayed and stored as inputs.
"""
import contextlib
import io
import os

import numpy as np
import torch

from oracle import lush_oracle as O
from oracle.ref_import import build_reference_model, import_reference, poster_args

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden_v1.npz")
G = {}


def put(name, t):
    G[name] = t.detach().cpu().numpy() if torch.is_tensor(t) else np.asarray(t)


def quiet():
    return contextlib.redirect_stdout(io.StringIO())


def grad_digest(prefix, named_grads):
    """Per-tensor digest of a gradient: sum, L2 norm, projection on a fixed pseudo-random vector,
    first 8 entries (keeps the fixture small while pinning every element)."""
    for j, (name, g) in enumerate(named_grads):
        g = g.detach().reshape(-1).double()
        probe = O.det_uniform((g.numel(),), 4242 + j).double()
        put(f"{prefix}/{name}/digest", torch.stack([g.sum(), g.norm(), (g * probe).sum()]))
        put(f"{prefix}/{name}/head", g[:8].float())


def main():
    ref_model, ref_helpers, _ = import_reference()
    cfg = O.POSTER_CFG
    nets = O.det_nets(cfg, tag=1, sigma_gain=30.0)
    rbk = O.det_rbk_params(tag=7)
    Sc, Nf = 24, 16
    model = build_reference_model(poster_args(N_importance=Nf, N_samples=Sc))
    model.load_state_dict(O.reference_state_dict(nets, rbk), strict=False)

    # ---------------- KATs of SURVEY.md 8c (recomputed from the reference here)
    raw = torch.tensor([[[0, 0, 0, 1], [1, -1, .5, 2], [2, 0, -2, .5], [.3, .2, .1, 9]]], dtype=torch.float32)
    z = torch.tensor([[0, .25, .5, 1]], dtype=torch.float32)
    rd = torch.tensor([[0, 0, 2]], dtype=torch.float32)
    put("kat1/raw", raw); put("kat1/z", z); put("kat1/rays_d", rd)
    model.train()
    for n, t in zip(("rgb", "density", "acc", "weights", "depth"), model.raw2outputs(raw, z, rd, 0)):
        put(f"kat1/train/{n}", t)
    model.eval()
    for n, t in zip(("rgb", "density", "acc", "weights", "depth"), model.raw2outputs(raw, z, rd, 0)):
        put(f"kat1/eval/{n}", t)
    model.train()
    bins = torch.tensor([[.1, .3, .5, .7, .9]]); w = torch.tensor([[.1, 0, .6, .3]])
    put("kat2/bins", bins); put("kat2/weights", w)
    put("kat2/det", ref_helpers.sample_pdf(bins, w, 6, det=True))
    put("kat2/pytest", ref_helpers.sample_pdf(bins, w, 6, det=False, pytest=True))
    fn, _ = ref_helpers.get_embedder(10, 0)
    put("kat3/embed", fn(torch.tensor([[0.1, 0.2, 0.3]])))
    put("kat4/param_counts", np.array([sum(p.numel() for p in model.mlp_coarse.parameters()),
                                       sum(p.numel() for p in model.mlp_fine.parameters()),
                                       sum(p.numel() for p in model.mlp_noise_coarse.parameters()),
                                       sum(p.numel() for p in model.mlp_rbk.parameters()),
                                       sum(p.numel() for p in model.parameters())]))

    # ---------------- raw2outputs, random, with gradients
    R, S = 16, 40
    raw = O.det_uniform((R, S, 4), 21, -3, 3); raw[..., 3] *= 5
    z = torch.sort(O.det_uniform((R, S), 22, 0, 1), -1)[0]
    rd = O.det_uniform((R, 3), 23, -1, 1)
    torch.manual_seed(5)
    noise = torch.randn(R, S - 1)
    cot = {n: O.det_uniform(s, 600 + i) for i, (n, s) in enumerate(
        (("rgb", (R, 3)), ("density", (R, S - 1)), ("acc", (R,)), ("weights", (R, S)), ("depth", (R,))))}
    put("r2o/raw", raw); put("r2o/z", z); put("r2o/rays_d", rd); put("r2o/noise", noise)
    for n, c in cot.items():
        put(f"r2o/cot/{n}", c)
    raw_g, rd_g = raw.clone().requires_grad_(True), rd.clone().requires_grad_(True)
    torch.manual_seed(5)
    outs = model.raw2outputs(raw_g, z, rd_g, raw_noise_std=1.0)
    loss = sum((o * cot[n]).sum() for n, o in zip(("rgb", "density", "acc", "weights", "depth"), outs))
    loss.backward()
    for n, t in zip(("rgb", "density", "acc", "weights", "depth"), outs):
        put(f"r2o/train/{n}", t)
    put("r2o/train/g_raw", raw_g.grad); put("r2o/train/g_rays_d", rd_g.grad)
    model.eval()
    for n, t in zip(("rgb", "density", "acc", "weights", "depth"), model.raw2outputs(raw, z, rd, 0.0)):
        put(f"r2o/eval/{n}", t)
    model.train()

    # ---------------- sample_pdf
    for ci, (Rr, nb, N) in enumerate(((64, 63, 64), (33, 31, 32), (8, 255, 256), (5, 9, 7))):
        bins = torch.sort(O.det_uniform((Rr, nb), 31 + ci, 0, 1), -1)[0]
        w = torch.relu(O.det_uniform((Rr, nb - 1), 41 + ci, -1, 1)) ** 3
        put(f"spdf{ci}/bins", bins); put(f"spdf{ci}/weights", w)
        put(f"spdf{ci}/det", ref_helpers.sample_pdf(bins, w, N, det=True))
        torch.manual_seed(9)
        put(f"spdf{ci}/rand", ref_helpers.sample_pdf(bins, w, N, det=False))
        torch.manual_seed(9)
        u = torch.rand(Rr, N)
        put(f"spdf{ci}/u", u)
        # indices with the reference's own op sequence (helpers:568-571,593)
        ww = w + 1e-5
        cdf = torch.cumsum(ww / torch.sum(ww, -1, keepdim=True), -1)
        cdf = torch.cat([torch.zeros_like(cdf[..., :1]), cdf], -1)
        put(f"spdf{ci}/cdf", cdf)
        put(f"spdf{ci}/inds_rand", torch.searchsorted(cdf, u.contiguous(), right=True))
        ud = torch.linspace(0., 1., steps=N).expand(Rr, N).contiguous()
        put(f"spdf{ci}/inds_det", torch.searchsorted(cdf, ud, right=True))

    # ---------------- MLPs on embedded inputs (fwd + grads)
    x = O.det_uniform((64, 3), 11, -1.5, 1.5)
    v = O.det_uniform((64, 3), 12, -1, 1)
    v = v / v.norm(dim=-1, keepdim=True)
    put("mlp/pts", x); put("mlp/views", v)
    fnv, _ = ref_helpers.get_embedder(4, 0)
    xe = torch.cat([fn(x), fnv(v)], -1).requires_grad_(True)
    for name, net in (("coarse", model.mlp_coarse), ("noise", model.mlp_noise_coarse)):
        model.zero_grad()
        out = net(xe)
        put(f"mlp/{name}/out", out)
        c = O.det_uniform(tuple(out.shape), 700)
        (out * c).sum().backward()
        grad_digest(f"mlp/{name}/grad", [(n, p.grad if p.grad is not None else torch.zeros_like(p))
                                          for n, p in net.named_parameters()])
    put("mlp/g_embedded_sum", xe.grad.sum(0))

    # ---------------- rays: get_rays + ray prep
    H, W, focal = 40, 60, 50.0
    K = np.array([[focal, 0, 0.5 * W], [0, focal, 0.5 * H], [0, 0, 1]])
    c2w = torch.tensor([[0.99, 0.02, -0.1, 0.1], [-0.02, 0.999, 0.03, -0.05], [0.1, -0.03, 0.99, 0.2]])
    put("rays/HWf", np.array([H, W, focal])); put("rays/c2w", c2w)
    ro, rd_ = ref_helpers.get_rays(H, W, K, c2w)
    no, nd = ref_helpers.ndc_rays(H, W, K[0][0], 1.0, ro, rd_)
    rays_all = torch.stack([ro, rd_], -1).reshape(-1, 3, 2)
    sel = torch.arange(0, rays_all.shape[0], rays_all.shape[0] // 24)[:24]
    rays = rays_all[sel]
    put("rays/rays", rays)
    put("rays/ndc_o", no.reshape(-1, 3)[sel]); put("rays/ndc_d", nd.reshape(-1, 3)[sel])
    put("rays/viewdirs", (rd_ / torch.norm(rd_, dim=-1, keepdim=True)).reshape(-1, 3)[sel])

    # ---------------- render_rays: eval (det) and train (seeded)
    rb = O.ray_prep(rays, H, W, K[0][0])
    keys = ("rgb_map", "depth_map", "acc_map", "density_map", "raw", "rgb0", "depth0", "acc0", "density0", "z_std")
    model.eval()
    with torch.no_grad(), quiet():
        ret, retn = model.render_rays(rb, Sc, retraw=True, perturb=0.0, N_importance=Nf, raw_noise_std=0.0,
                                      inference=True)
    for k in keys:
        put(f"rr/eval/{k}", ret[k])
    put("rr/eval/noise_rgb", retn["rgb_map"])
    model.train()
    torch.manual_seed(3)
    with quiet():
        ret, retn = model.render_rays(rb, Sc, retraw=True, perturb=1.0, N_importance=Nf, raw_noise_std=1.0)
    torch.manual_seed(3)
    put("rr/train/t_rand", torch.rand(24, Sc)); put("rr/train/noise_c", torch.randn(24, Sc - 1))
    put("rr/train/u", torch.rand(24, Nf)); put("rr/train/noise_f", torch.randn(24, Sc + Nf - 1))
    for k in keys:
        put(f"rr/train/{k}", ret[k])

    # ---------------- RBK + full NeRFAll.forward (training): outputs and gradients
    N = 12
    rays_n = rays_all[torch.arange(0, rays_all.shape[0], rays_all.shape[0] // N)[:N]].clone()
    idx = (torch.arange(N) % 20).reshape(N, 1)
    put("fwd/rays", rays_n); put("fwd/images_idx", idx)
    new, ccw = model.mlp_rbk(rays_n, {"images_idx": idx})
    put("rbk/new_rays", new); put("rbk/ccw", ccw)
    kp = torch.arange(N) % 2 == 0
    put("fwd/kernel_pixel", kp)
    tgt = O.det_uniform((N, 3), 77, 0, 1)
    put("fwd/target", tgt)
    for mode in ("rbk", "naive"):
        model.zero_grad()
        rays_in = rays_n.clone().requires_grad_(True)
        torch.manual_seed(4)
        with quiet():
            out = model(H, W, K, chunk=1024, rays=rays_in, rays_info={"images_idx": idx}, retraw=True,
                        force_naive=(mode == "naive"), allkernel=True, kernel_pixel=kp, perturb=1.0,
                        N_importance=Nf, N_samples=Sc, use_viewdirs=True, white_bkgd=False, raw_noise_std=1.0,
                        inference=False, near=0.0, far=1.0)
        Kr = 5 if mode == "rbk" else 1
        torch.manual_seed(4)
        put(f"fwd/{mode}/t_rand", torch.rand(N * Kr, Sc)); put(f"fwd/{mode}/noise_c", torch.randn(N * Kr, Sc - 1))
        put(f"fwd/{mode}/u", torch.rand(N * Kr, Nf)); put(f"fwd/{mode}/noise_f", torch.randn(N * Kr, Sc + Nf - 1))
        for i in (0, 1, 3) + ((5, 6) if mode == "rbk" else ()):
            put(f"fwd/{mode}/out{i}", out[i])
        loss = ((out[0] - tgt) ** 2).mean() + ((out[1] - tgt) ** 2).mean() \
            + (out[0] - tgt).abs().mean() + (out[1] - tgt).abs().mean()
        loss.backward()
        put(f"fwd/{mode}/loss", loss)
        named = [(n, p.grad if p.grad is not None else torch.zeros_like(p)) for n, p in model.named_parameters()
                 if not n.startswith(("blur_kernel_net", "dbk_view_embedding"))]
        grad_digest(f"fwd/{mode}/grad", named)
        if mode == "rbk":
            put("fwd/rbk/g_rays", rays_in.grad)

    # eval branch of NeRFAll.forward on a tiny image
    Hs, Ws = 6, 8
    Ks = np.array([[10.0, 0, 0.5 * Ws], [0, 10.0, 0.5 * Hs], [0, 0, 1]])
    poses = torch.stack([c2w, c2w + 0.01], 0)
    put("eval/HW", np.array([Hs, Ws])); put("eval/K", Ks); put("eval/poses", poses)
    model.eval()
    with torch.no_grad(), quiet():
        rgbs, noise_img, depths = model(Hs, Ws, Ks, chunk=32, poses=poses, render_kwargs=dict(
            perturb=False, N_importance=Nf, N_samples=Sc, use_viewdirs=True, white_bkgd=False, raw_noise_std=0.,
            inference=True, near=0., far=1.))
    put("eval/rgbs", rgbs); put("eval/noise", noise_img); put("eval/depths", depths)

    np.savez_compressed(OUT, **G)
    print(f"wrote {OUT}: {len(G)} arrays, {os.path.getsize(OUT) / 1024:.1f} KiB")


if __name__ == "__main__":
    main()
