"""Pin oracle/lush_oracle.py against the UNMODIFIED reference (runs only where /root/reference exists).

    python -m oracle.validate_against_reference

Prints one line per check with the max-abs difference and exits non-zero on any failure.
Bit-exact checks: embed (identity/power-of-two scaling), sample_pdf samples + indices, z merge.
Everything else must agree to <= 2e-6 (same torch ops in a different grouping).
"""
import contextlib
import io
import sys

import numpy as np
import torch

from oracle import lush_oracle as O
from oracle.ref_import import build_reference_model, import_reference, poster_args

FAILS = []


def check(name, a, b, tol=0.0):
    a, b = torch.as_tensor(a).detach(), torch.as_tensor(b).detach()
    if a.shape != b.shape:
        FAILS.append(name)
        print(f"FAIL {name}: shape {tuple(a.shape)} vs {tuple(b.shape)}")
        return
    d = (a.double() - b.double()).abs().max().item() if a.numel() else 0.0
    ok = d <= tol
    print(f"{'ok  ' if ok else 'FAIL'} {name}: max|diff| = {d:.3e} (tol {tol:g})")
    if not ok:
        FAILS.append(name)


def main():
    torch.manual_seed(0)
    ref_model, ref_helpers, ref_warp = import_reference()
    cfg = O.POSTER_CFG
    nets = O.det_nets(cfg, tag=1, sigma_gain=30.0)
    rbk = O.det_rbk_params(tag=7)
    model = build_reference_model(poster_args(N_importance=16, N_samples=24))
    missing = model.load_state_dict(O.reference_state_dict(nets, rbk), strict=False)
    assert not missing.unexpected_keys, missing.unexpected_keys

    # ---- a-1 embed
    x = O.det_uniform((50, 3), 11, -1.5, 1.5)
    fn, od = ref_helpers.get_embedder(10, 0)
    check("embed L=10", O.embed(x, 10), fn(x), 0.0)
    fn4, _ = ref_helpers.get_embedder(4, 0)
    check("embed L=4", O.embed(x, 4), fn4(x), 0.0)

    # ---- a-2 / a-3 MLPs
    xe = torch.cat([O.embed(x, 10), O.embed(x / x.norm(dim=-1, keepdim=True), 4)], -1)
    check("NeRF.forward", O.mlp_forward(xe, nets["coarse"], 8, 256, 63, 27, 4), model.mlp_coarse(xe), 2e-6)
    check("NeRF_Noise.forward", O.mlp_forward(xe, nets["noise"], 4, 128, 63, 27, 4, out_mode=1),
          model.mlp_noise_coarse(xe), 2e-6)

    # ---- a-6 raw2outputs, train (noise by seed replay) and eval (near-plane mask)
    R, S = 16, 40
    raw = O.det_uniform((R, S, 4), 21, -3, 3)
    raw[..., 3] *= 5
    z = torch.sort(O.det_uniform((R, S), 22, 0, 1), -1)[0]
    rd = O.det_uniform((R, 3), 23, -1, 1)
    model.train()
    torch.manual_seed(5)
    ref_out = model.raw2outputs(raw, z, rd, raw_noise_std=1.0)
    torch.manual_seed(5)
    noise = torch.randn(R, S - 1) * 1.0
    mine = O.raw2outputs(raw, z, rd, noise)
    for n, a, b in zip(("rgb", "density", "acc", "weights", "depth"), mine, ref_out):
        check(f"raw2outputs train {n}", a, b, 1e-6)
    model.eval()
    ref_out = model.raw2outputs(raw, z, rd, raw_noise_std=0.0)
    mine = O.raw2outputs(raw, z, rd, None, rm_near_thresh=80 / 128)
    for n, a, b in zip(("rgb", "density", "acc", "weights", "depth"), mine, ref_out):
        check(f"raw2outputs eval {n}", a, b, 1e-6)
    model.train()

    # ---- a-7 sample_pdf: det and random, bit exact (incl. indices recomputed with the same torch ops)
    for (Rr, nb, N) in ((64, 63, 64), (33, 31, 32), (8, 255, 256), (5, 9, 7)):
        bins = torch.sort(O.det_uniform((Rr, nb), 31, 0, 1), -1)[0]
        w = torch.relu(O.det_uniform((Rr, nb - 1), 32, -1, 1)) ** 3
        ref_s = ref_helpers.sample_pdf(bins, w, N, det=True)
        s, inds, cdf = O.sample_pdf(bins, w, torch.linspace(0.0, 1.0, N))
        check(f"sample_pdf det nb={nb} N={N}", s, ref_s, 0.0)
        torch.manual_seed(9)
        ref_s = ref_helpers.sample_pdf(bins, w, N, det=False)
        torch.manual_seed(9)
        u = torch.rand(Rr, N)
        s, inds, cdf = O.sample_pdf(bins, w, u)
        check(f"sample_pdf rand nb={nb} N={N}", s, ref_s, 0.0)
        # kernel-order cdf: indices may differ only where u sits within rounding of a cdf entry
        s2, inds2, cdf2 = O.sample_pdf(bins, w, u, order="kernel")
        flips = (inds != inds2).sum().item()
        print(f"     kernel-order cdf: max|cdf diff| {float((cdf - cdf2).abs().max()):.2e}, index flips "
              f"{flips}/{inds.numel()}, max|sample diff| {float((s - s2).abs().max()):.2e}")

    # ---- a-9 ray prep: get_rays + viewdirs + ndc
    H, W, focal = 40, 60, 50.0
    K = np.array([[focal, 0, 0.5 * W], [0, focal, 0.5 * H], [0, 0, 1]])
    c2w = torch.tensor([[0.99, 0.02, -0.1, 0.1], [-0.02, 0.999, 0.03, -0.05], [0.1, -0.03, 0.99, 0.2]])
    ro, rd_ = ref_helpers.get_rays(H, W, K, c2w)
    mo, md = O.get_rays(H, W, K, c2w)
    check("get_rays o", mo, ro, 0.0)
    check("get_rays d", md, rd_, 0.0)
    no, nd = ref_helpers.ndc_rays(H, W, K[0][0], 1.0, ro, rd_)
    rays = torch.stack([ro, rd_], -1).reshape(-1, 3, 2)
    rb = O.ray_prep(rays, H, W, K[0][0])
    check("ray_prep ndc o", rb[:, 0:3], no.reshape(-1, 3), 0.0)
    check("ray_prep ndc d", rb[:, 3:6], nd.reshape(-1, 3), 0.0)

    # ---- a-8 render_rays (det) and with perturb / noise (seed replay of the four draws, 7.3-6)
    Rr, Sc, Nf = 24, 24, 16
    rays_s = rays[torch.arange(0, rays.shape[0], rays.shape[0] // Rr)[:Rr]]
    rb = O.ray_prep(rays_s, H, W, K[0][0])
    model.eval()
    with torch.no_grad(), contextlib.redirect_stdout(io.StringIO()):
        ref_ret, ref_noise = model.render_rays(rb, Sc, retraw=True, perturb=0.0, N_importance=Nf,
                                               raw_noise_std=0.0, inference=True)
        ret, retn = O.render_rays(nets, cfg, rb, N_samples=Sc, N_importance=Nf, u=torch.linspace(0, 1, Nf),
                                  rm_near_thresh=80 / 128)
    for k in ("rgb_map", "depth_map", "acc_map", "density_map", "raw", "rgb0", "depth0", "acc0", "density0",
              "z_std"):
        check(f"render_rays eval {k}", ret[k], ref_ret[k], 5e-6)
    check("render_rays eval noise rgb", retn["rgb_map"], ref_noise["rgb_map"], 2e-6)
    model.train()
    torch.manual_seed(3)
    with contextlib.redirect_stdout(io.StringIO()):
        ref_ret, ref_noise = model.render_rays(rb, Sc, retraw=True, perturb=1.0, N_importance=Nf,
                                               raw_noise_std=1.0)
    torch.manual_seed(3)
    t_rand = torch.rand(Rr, Sc)
    noise_c = torch.randn(Rr, Sc - 1)
    u = torch.rand(Rr, Nf)
    noise_f = torch.randn(Rr, Sc + Nf - 1)
    ret, retn = O.render_rays(nets, cfg, rb, N_samples=Sc, N_importance=Nf, t_rand=t_rand, noise_c=noise_c,
                              noise_f=noise_f, u=u)
    for k in ("rgb_map", "depth_map", "acc_map", "density_map", "raw", "rgb0", "depth0", "acc0", "density0",
              "z_std"):
        check(f"render_rays train {k}", ret[k], ref_ret[k], 5e-6)

    # ---- a-10 RBK
    N = 12
    rays_n = rays[torch.arange(0, rays.shape[0], rays.shape[0] // N)[:N]].clone()
    idx = (torch.arange(N) % 20).reshape(N, 1)
    ref_new, ref_w = model.mlp_rbk(rays_n, {"images_idx": idx})
    new, w = O.rbk_forward(rbk, rays_n, idx)
    check("rbk new_rays", new, ref_new, 2e-6)
    check("rbk ccw", w, ref_w, 1e-6)

    # ---- a-12 NeRFAll.forward training, RBK and naive, outputs + gradients
    for mode in ("rbk", "naive"):
        model.zero_grad()
        rays_in = rays_n.clone().requires_grad_(True)
        kp = torch.arange(N) % 2 == 0
        torch.manual_seed(4)
        with contextlib.redirect_stdout(io.StringIO()):
            out = model(H, W, K, chunk=1024, rays=rays_in, rays_info={"images_idx": idx}, retraw=True,
                        force_naive=(mode == "naive"), allkernel=True, kernel_pixel=kp, perturb=1.0,
                        N_importance=Nf, N_samples=Sc, use_viewdirs=True, white_bkgd=False,
                        raw_noise_std=1.0, inference=False, near=0.0, far=1.0)
        Kr = 5 if mode == "rbk" else 1
        torch.manual_seed(4)
        t_rand = torch.rand(N * Kr, Sc)
        noise_c = torch.randn(N * Kr, Sc - 1)
        u = torch.rand(N * Kr, Nf)
        noise_f = torch.randn(N * Kr, Sc + Nf - 1)
        pn = {k: {n: t.clone().requires_grad_(True) for n, t in d.items()} for k, d in nets.items()}
        pr = {n: t.clone().requires_grad_(True) for n, t in rbk.items()}
        rays_o = rays_n.clone().requires_grad_(True)
        mine, _ = O.nerfall_train_forward(pn, pr, cfg, rays_o, idx, H, W, K[0][0], N_samples=Sc,
                                          N_importance=Nf, force_naive=(mode == "naive"), allkernel=True,
                                          kernel_pixel=kp, t_rand=t_rand, noise_c=noise_c, noise_f=noise_f, u=u)
        for i in (0, 1, 3):
            check(f"NeRFAll.forward[{mode}] out{i}", mine[i], out[i], 5e-6)
        tgt = O.det_uniform((N, 3), 77, 0, 1)
        loss_ref = ((out[0] - tgt) ** 2).mean() + ((out[1] - tgt) ** 2).mean()
        loss_ref.backward()
        loss = ((mine[0] - tgt) ** 2).mean() + ((mine[1] - tgt) ** 2).mean()
        loss.backward()
        check(f"NeRFAll.forward[{mode}] d coarse.pts0.w", pn["coarse"]["pts_linears.0.weight"].grad,
              model.mlp_coarse.pts_linears[0].weight.grad, 1e-6)
        check(f"NeRFAll.forward[{mode}] d fine.rgb.w", pn["fine"]["rgb_linear.weight"].grad,
              model.mlp_fine.rgb_linear.weight.grad, 1e-6)
        if mode == "rbk":
            check("NeRFAll.forward[rbk] d rays", rays_o.grad, rays_in.grad, 1e-6)
            check("NeRFAll.forward[rbk] d rbk.r_linear.w", pr["r_linear.weight"].grad,
                  model.mlp_rbk.r_linear.weight.grad, 2e-6)
            check("NeRFAll.forward[rbk] d noise.rgb.w", pn["noise"]["rgb_linear.weight"].grad,
                  model.mlp_noise_coarse.rgb_linear.weight.grad, 1e-6)

    print(f"\n{len(FAILS)} failure(s)" + (": " + ", ".join(FAILS) if FAILS else ""))
    return 1 if FAILS else 0


if __name__ == "__main__":
    sys.exit(main())
