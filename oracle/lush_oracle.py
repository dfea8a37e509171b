"""CPU oracle for the LuSh-NeRF per-ray rendering hot path -- TEST INFRASTRUCTURE ONLY.

A torch-fp32 restatement (pure functions, CPU tensors, autograd-differentiable) of the reference
algorithm, each function citing the reference file:line it follows (paths relative to the
quzefan/LuSh-NeRF root).  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs may import this module; the product path (lush_nerf_b200/) never does.

Pinning: the reference has no tests or golden vectors of its own, so this oracle
is pinned by executing the reference itself in the build container:
  * oracle/validate_against_reference.py compares every function here with the imported reference on
    seeded inputs (bit-exact for sample_pdf indices, <= a few ulp elsewhere), and
  * tests/golden/make_golden.py stores reference outputs as fixtures that tests/test_oracle_golden.py
    replays on any machine (the reference tree does not travel to the GPU box).
Random draws (stratified jitter, sigma noise, sampling u) are explicit inputs so that the oracle, the
reference (same seed) and the CUDA path consume identical numbers.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch
import torch.nn.functional as F

# ----------------------------------------------------------------------------------------------
# deterministic tensors (bit-reproducible everywhere; used for weights of golden cases)
# ----------------------------------------------------------------------------------------------

def det_uniform(shape, tag: int, lo: float = -1.0, hi: float = 1.0) -> torch.Tensor:
    """Counter-based uniform numbers in [lo, hi): splitmix64 on (tag, index) -> 24 random bits."""
    n = int(np.prod(shape)) if len(shape) else 1
    with np.errstate(over="ignore"):
        idx = np.arange(n, dtype=np.uint64) + np.uint64(tag) * np.uint64(0x9E3779B97F4A7C15)
        x = idx + np.uint64(0x9E3779B97F4A7C15)
        x = (x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        x = (x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        x = x ^ (x >> np.uint64(31))
    u = (x >> np.uint64(40)).astype(np.float64) / float(1 << 24)
    out = (lo + (hi - lo) * u).astype(np.float32)
    return torch.from_numpy(out.reshape(shape))


# ----------------------------------------------------------------------------------------------
# network topology / parameters
# ----------------------------------------------------------------------------------------------

def mlp_param_shapes(D: int, W: int, in_ch: int, in_ch_v: int, skip: int) -> List[Tuple[str, Tuple[int, ...]]]:
    """state_dict order of NeRF / NeRF_Noise (utils/run_lushnerf_helpers.py:379-390, 470-479)."""
    shapes = []
    for i in range(D):
        fan = in_ch if i == 0 else (W + in_ch if i == skip + 1 else W)
        shapes.append((f"pts_linears.{i}.weight", (W, fan)))
        shapes.append((f"pts_linears.{i}.bias", (W,)))
    shapes += [("views_linears.0.weight", (W // 2, W + in_ch_v)), ("views_linears.0.bias", (W // 2,)),
               ("feature_linear.weight", (W, W)), ("feature_linear.bias", (W,)),
               ("alpha_linear.weight", (1, W)), ("alpha_linear.bias", (1,)),
               ("rgb_linear.weight", (3, W // 2)), ("rgb_linear.bias", (3,))]
    return shapes


def det_mlp_params(D: int, W: int, in_ch: int, in_ch_v: int, skip: int, tag: int,
                   gain: float = 1.0) -> Dict[str, torch.Tensor]:
    """Deterministic nn.Linear-like init: U(-gain/sqrt(fan_in), gain/sqrt(fan_in))."""
    out = {}
    for j, (name, shape) in enumerate(mlp_param_shapes(D, W, in_ch, in_ch_v, skip)):
        fan_in = shape[1] if len(shape) == 2 else None
        if fan_in is None:  # bias: use the matching weight's fan-in
            fan_in = out[name.replace("bias", "weight")].shape[1]
        b = gain / math.sqrt(fan_in)
        out[name] = det_uniform(shape, tag * 1000 + j, -b, b)
    return out


def flatten_params(params: Dict[str, torch.Tensor], D, W, in_ch, in_ch_v, skip) -> torch.Tensor:
    return torch.cat([params[n].reshape(-1) for n, _ in mlp_param_shapes(D, W, in_ch, in_ch_v, skip)])


# ----------------------------------------------------------------------------------------------
# a-1 positional encoding  (utils/run_lushnerf_helpers.py:311-361)
# ----------------------------------------------------------------------------------------------

def embed(x: torch.Tensor, L: int) -> torch.Tensor:
    """[x, sin(2^0 x), cos(2^0 x), ..., sin(2^(L-1) x), cos(2^(L-1) x)] along the last axis."""
    parts = [x]
    for l in range(L):
        f = float(2.0 ** l)  # freq_bands = 2**linspace(0, L-1, L): exact powers of two (helpers:324)
        parts.append(torch.sin(x * f))
        parts.append(torch.cos(x * f))
    return torch.cat(parts, -1)


# ----------------------------------------------------------------------------------------------
# a-2 / a-3 scene + noise MLP  (utils/run_lushnerf_helpers.py:394-423, 483-512)
# ----------------------------------------------------------------------------------------------

def mlp_forward(x: torch.Tensor, p: Dict[str, torch.Tensor], D: int, W: int, in_ch: int, in_ch_v: int,
                skip: int, out_mode: int = 0) -> torch.Tensor:
    """x = [embedded pts | embedded views] -> [rgb(3), sigma] (out_mode 0) or rgb(3) (out_mode 1)."""
    e_pts, e_views = x[..., :in_ch], x[..., in_ch:in_ch + in_ch_v]
    h = e_pts
    for i in range(D):
        h = F.relu(F.linear(h, p[f"pts_linears.{i}.weight"], p[f"pts_linears.{i}.bias"]))
        if i == skip:
            h = torch.cat([e_pts, h], -1)
    alpha = F.linear(h, p["alpha_linear.weight"], p["alpha_linear.bias"])
    feat = F.linear(h, p["feature_linear.weight"], p["feature_linear.bias"])
    hv = F.relu(F.linear(torch.cat([feat, e_views], -1), p["views_linears.0.weight"],
                         p["views_linears.0.bias"]))
    rgb = F.linear(hv, p["rgb_linear.weight"], p["rgb_linear.bias"])
    return torch.cat([rgb, alpha], -1) if out_mode == 0 else rgb


def mlp_on_rays(p, ray_batch: torch.Tensor, z: torch.Tensor, D, W, L_pts, L_views, skip, out_mode=0):
    """NeRFAll.mlpforward (models/lushnerf.py:234-266) with the point generation of :414 / :441:
    pts = o + d*z, views expanded to every sample.  ray_batch [R,11], z [R,S] -> [R,S,out]."""
    o, d, v = ray_batch[:, 0:3], ray_batch[:, 3:6], ray_batch[:, 8:11]
    pts = o[:, None, :] + d[:, None, :] * z[:, :, None]
    R, S = z.shape
    e = embed(pts.reshape(-1, 3), L_pts)
    ev = embed(v[:, None, :].expand(R, S, 3).reshape(-1, 3), L_views)
    out = mlp_forward(torch.cat([e, ev], -1), p, D, W, 3 + 6 * L_pts, 3 + 6 * L_views, skip, out_mode)
    return out.reshape(R, S, -1)


# ----------------------------------------------------------------------------------------------
# a-6 alpha compositing  (models/lushnerf.py:296-352)
# ----------------------------------------------------------------------------------------------

def raw2outputs(raw, z, rays_d, sigma_noise=None, rm_near_thresh: float = -1.0, white_bkgd=False):
    """Returns (rgb_map [R,3], density [R,S-1], acc [R], weights [R,S], depth [R]).
    sigma_noise [R,S-1] is randn*raw_noise_std drawn by the caller (:322).  rm_near_thresh >= 0
    applies the eval-only near-plane mask z_{i+1} > thresh (:331-335, thresh = render_rmnearplane/128)."""
    dists = (z[:, 1:] - z[:, :-1]) * torch.norm(rays_d[:, None, :], dim=-1)          # :314-317
    rgb = torch.sigmoid(raw[..., :3])                                                  # :319
    pre = raw[:, :-1, 3] if sigma_noise is None else raw[:, :-1, 3] + sigma_noise      # :330
    density = torch.relu(pre)
    if rm_near_thresh >= 0:
        density = (z[:, 1:] > rm_near_thresh).type_as(density) * density               # :331-335
    alpha = 1.0 - torch.exp(-density * dists)                                          # :337
    alpha = torch.cat([alpha, torch.ones_like(alpha[:, :1])], -1)                      # :338
    # (1 + 1e-10) is a Python double that rounds to 1.0f when it meets the fp32 tensor      :341
    trans = torch.cumprod(torch.cat([torch.ones_like(alpha[:, :1]), 1.0 - alpha], -1), -1)[:, :-1]
    weights = alpha * trans
    rgb_map = (weights[..., None] * rgb).sum(-2)                                       # :343
    depth = (weights * z).sum(-1)                                                      # :344
    acc = weights.sum(-1)                                                              # :347
    if white_bkgd:
        rgb_map = rgb_map + (1.0 - acc[..., None])                                     # :349-350
    return rgb_map, density, acc, weights, depth


# ----------------------------------------------------------------------------------------------
# a-7 hierarchical sampling  (utils/run_lushnerf_helpers.py:566-609)
# ----------------------------------------------------------------------------------------------

def _cdf_torch(weights: torch.Tensor) -> torch.Tensor:
    w = weights + 1e-5                                                                 # :568
    pdf = w / torch.sum(w, -1, keepdim=True)                                           # :569
    cdf = torch.cumsum(pdf, -1)                                                        # :570
    return torch.cat([torch.zeros_like(cdf[..., :1]), cdf], -1)                        # :571


def _cdf_kernel_order(weights: torch.Tensor) -> torch.Tensor:
    """Same arithmetic as _cdf_torch but with the FIXED association order of the CUDA kernel
    (csrc/sampler.cu::warp_cdf_build): 32 lanes each own E = ceil(n/32) consecutive elements,
    lane partials left-to-right, xor-butterfly total, Kogge-Stone scan of lane totals.  fp32 only."""
    w = weights.detach().cpu().numpy().astype(np.float32)
    R, n = w.shape
    E = (n + 31) // 32
    wp = np.zeros((R, 32 * E), np.float32)
    wp[:, :n] = w + np.float32(1e-5)
    valid = np.zeros(32 * E, bool)
    valid[:n] = True
    wp = wp.reshape(R, 32, E)
    vm = valid.reshape(32, E)
    part = np.zeros((R, 32), np.float32)
    for e in range(E):
        part = np.where(vm[None, :, e], (part + wp[:, :, e]).astype(np.float32), part)
    tot = part.copy()
    lanes = np.arange(32)
    for o in (16, 8, 4, 2, 1):
        tot = (tot + tot[:, lanes ^ o]).astype(np.float32)
    total = tot[:, :1]  # identical in every lane
    run = np.zeros((R, 32), np.float32)
    runs = np.zeros((R, 32, E), np.float32)
    for e in range(E):
        pdf = (wp[:, :, e] / total).astype(np.float32)
        run = np.where(vm[None, :, e], (run + pdf).astype(np.float32), run)
        runs[:, :, e] = run
    v = run.copy()
    for o in (1, 2, 4, 8, 16):
        sh = np.zeros_like(v)
        sh[:, o:] = v[:, :-o]
        v = np.where(lanes[None, :] >= o, (v + sh).astype(np.float32), v)
    off = np.zeros_like(v)
    off[:, 1:] = v[:, :-1]
    cdf_body = (off[:, :, None] + runs).astype(np.float32).reshape(R, 32 * E)[:, :n]
    cdf = np.concatenate([np.zeros((R, 1), np.float32), cdf_body], 1)
    return torch.from_numpy(cdf)


def sample_pdf(bins, weights, u, order: str = "torch", cdf: Optional[torch.Tensor] = None):
    """bins [R,nb], weights [R,nb-1], u [R,N] (or [N], broadcast) -> (samples, inds, cdf).
    order="torch": cdf by torch.sum/cumsum exactly as the reference; order="kernel": fixed fp32
    association order of the CUDA kernel; cdf=...: strict mode, search/gather/lerp only."""
    if cdf is None:
        cdf = _cdf_torch(weights) if order == "torch" else _cdf_kernel_order(weights)
    if u.dim() == 1:
        u = u.expand(list(cdf.shape[:-1]) + [u.shape[0]])
    u = u.contiguous()
    inds = torch.searchsorted(cdf, u, right=True)                                      # :593
    below = torch.clamp(inds - 1, min=0)                                               # :594
    above = torch.clamp(inds, max=cdf.shape[-1] - 1)                                   # :595
    c0, c1 = torch.gather(cdf, 1, below), torch.gather(cdf, 1, above)                  # :601
    b0, b1 = torch.gather(bins, 1, below), torch.gather(bins, 1, above)                # :602
    denom = c1 - c0
    denom = torch.where(denom < 1e-5, torch.ones_like(denom), denom)                   # :604-605
    t = (u - c0) / denom
    return b0 + t * (b1 - b0), inds, cdf                                               # :606-607


# ----------------------------------------------------------------------------------------------
# a-9 ray preparation  (models/lushnerf.py:706-729; utils/run_lushnerf_helpers.py:517-562)
# ----------------------------------------------------------------------------------------------

def get_rays(H: int, W: int, K, c2w: torch.Tensor):
    """helpers:517-528 (HALF_PIX = 0.5 pixel-centre convention)."""
    j, i = torch.meshgrid(torch.linspace(0, H - 1, H), torch.linspace(0, W - 1, W), indexing="ij")
    dirs = torch.stack([(i + (0.5 - K[0][2])) / K[0][0], -(j + (0.5 - K[1][2])) / K[1][1],
                        -torch.ones_like(i)], -1)
    rays_d = torch.sum(dirs[..., None, :] * c2w[:3, :3], -1)
    rays_o = c2w[:3, -1].expand(rays_d.shape)
    return rays_o, rays_d


def ndc_rays(H, W, focal, near, rays_o, rays_d):
    """helpers:542-562."""
    t = -(near + rays_o[..., 2]) / rays_d[..., 2]
    rays_o = rays_o + t[..., None] * rays_d
    cx, cy = -1.0 / (W / (2.0 * focal)), -1.0 / (H / (2.0 * focal))
    o0 = cx * rays_o[..., 0] / rays_o[..., 2]
    o1 = cy * rays_o[..., 1] / rays_o[..., 2]
    o2 = 1.0 + 2.0 * near / rays_o[..., 2]
    d0 = cx * (rays_d[..., 0] / rays_d[..., 2] - rays_o[..., 0] / rays_o[..., 2])
    d1 = cy * (rays_d[..., 1] / rays_d[..., 2] - rays_o[..., 1] / rays_o[..., 2])
    d2 = -2.0 * near / rays_o[..., 2]
    return torch.stack([o0, o1, o2], -1), torch.stack([d0, d1, d2], -1)


def ray_prep(rays: torch.Tensor, H, W, focal, near=0.0, far=1.0, ndc=True) -> torch.Tensor:
    """rays [R,3,2] -> ray_batch [R,11] = (o, d, near, far, viewdirs)  (models/lushnerf.py:706-729)."""
    o, d = rays[..., 0], rays[..., 1]
    v = d / torch.norm(d, dim=-1, keepdim=True)
    if ndc:
        o, d = ndc_rays(H, W, focal, 1.0, o, d)
    nf = torch.ones_like(d[..., :1])
    return torch.cat([o, d, near * nf, far * nf, v], -1)


def stratified_z(near, far, S: int, t_rand: Optional[torch.Tensor] = None, lindisp=False):
    """models/lushnerf.py:389-412.  near/far [R,1]; t_rand [R,S] uniform draws or None."""
    t = torch.linspace(0.0, 1.0, steps=S)
    z = near * (1.0 - t) + far * t if not lindisp else 1.0 / (1.0 / near * (1.0 - t) + 1.0 / far * t)
    z = z.expand(near.shape[0], S)
    if t_rand is not None:
        mids = 0.5 * (z[..., 1:] + z[..., :-1])
        upper = torch.cat([mids, z[..., -1:]], -1)
        lower = torch.cat([z[..., :1], mids], -1)
        z = lower + (upper - lower) * t_rand
    return z


# ----------------------------------------------------------------------------------------------
# a-8 render_rays  (models/lushnerf.py:354-479 / 481-583 / 585-617)
# ----------------------------------------------------------------------------------------------

def render_rays(nets: Dict[str, Dict[str, torch.Tensor]], cfg: dict, ray_batch, *, N_samples, N_importance,
                t_rand=None, noise_c=None, noise_f=None, u=None, rm_near_thresh=-1.0, white_bkgd=False,
                with_noise_mlp=True, lindisp=False, sample_order="torch"):
    """nets: {"coarse","fine","noise"} -> param dicts.  cfg: D, W, L_pts, L_views, skip.
    t_rand [R,S] / None (perturb), noise_c [R,S-1] & noise_f [R,S+N-1] (sigma noise, already scaled)
    / None, u [N] (det: torch.linspace(0,1,N)) or [R,N] (random).  Returns (ret, ret_noise) dicts."""
    D, W, Lp, Lv, skip = cfg["D"], cfg["W"], cfg["L_pts"], cfg["L_views"], cfg["skip"]
    near, far = ray_batch[:, 6:7], ray_batch[:, 7:8]
    rays_d = ray_batch[:, 3:6]
    z_plain = stratified_z(near, far, N_samples, None, lindisp)
    z = stratified_z(near, far, N_samples, t_rand, lindisp)
    raw = mlp_on_rays(nets["coarse"], ray_batch, z, D, W, Lp, Lv, skip)
    ret_noise = {}
    if with_noise_mlp:
        # one point per ray: sample index 16 of the UN-jittered z, inputs detached (:271, :422)
        rb = ray_batch.detach()
        nz = mlp_on_rays(nets["noise"], rb, z_plain[:, 16:17].detach(), cfg["D"] // 2, cfg["W"] // 2, Lp, Lv,
                         skip, out_mode=1)
        ret_noise["rgb_map"] = nz[:, 0, :]
    rgb0, dens0, acc0, w0, depth0 = raw2outputs(raw, z, rays_d, noise_c, rm_near_thresh, white_bkgd)
    ret = {}
    if N_importance > 0:
        z_mid = 0.5 * (z[:, 1:] + z[:, :-1])                                           # :435
        z_samples, _, _ = sample_pdf(z_mid, w0[:, 1:-1].detach(), u, order=sample_order)  # :436
        z_samples = z_samples.detach()
        z_f, _ = torch.sort(torch.cat([z, z_samples], -1), -1)                        # :440
        fine = nets["fine"] if nets.get("fine") is not None else nets["coarse"]
        raw_f = mlp_on_rays(fine, ray_batch, z_f, D, W, Lp, Lv, skip)
        rgb, dens, acc, w, depth = raw2outputs(raw_f, z_f, rays_d, noise_f, rm_near_thresh, white_bkgd)
        ret.update(rgb_map=rgb, depth_map=depth, acc_map=acc, density_map=dens, raw=raw_f, rgb0=rgb0,
                   depth0=depth0, acc0=acc0, density0=dens0,
                   z_std=torch.std(z_samples, dim=-1, unbiased=False), z_vals=z_f, z_samples=z_samples,
                   weights=w, weights0=w0)
    else:
        ret.update(rgb_map=rgb0, depth_map=depth0, acc_map=acc0, density_map=dens0, raw=raw, weights=w0)
    return ret, ret_noise


# ----------------------------------------------------------------------------------------------
# a-10 rigid-warp blur kernel  (models/lushnerf.py:75-98,118-153; utils/rigid_warping.py:20-140)
# ----------------------------------------------------------------------------------------------

def se3_warp(pts, rot, trans):
    """utils/rigid_warping.py:20-45 (exp map, Modern Robotics 3.51/3.88) applied to points [N,3]."""
    theta = torch.linalg.norm(rot, dim=-1) + 1.0e-10
    w = rot / theta[:, None]
    v = trans / theta[:, None]
    zero = torch.zeros_like(w[:, 0])
    Wm = torch.stack([torch.stack([zero, -w[:, 2], w[:, 1]], -1),
                      torch.stack([w[:, 2], zero, -w[:, 0]], -1),
                      torch.stack([-w[:, 1], w[:, 0], zero], -1)], -2)                 # skew :93-112
    th = theta[:, None, None]
    eye = torch.eye(3)[None]
    W2 = Wm @ Wm
    Rm = eye + torch.sin(th) * Wm + (1.0 - torch.cos(th)) * W2                         # :76-91
    p = (th * eye + (1.0 - torch.cos(th)) * Wm + (th - torch.sin(th)) * W2) @ v[..., None]  # :72-73
    return (Rm @ pts[..., None] + p)[..., 0]


def rbk_warp(rays, r, v):
    """rays [N,3,2]; r, v [N,3,M] -> new_rays [N,M+1,3,2], slot 0 = original ray (lushnerf.py:75-98)."""
    o, d = rays[..., 0], rays[..., 1]
    end = o + d
    out = [rays]
    for m in range(r.shape[-1]):
        wo = se3_warp(o, r[:, :, m], v[:, :, m])
        we = se3_warp(end, r[:, :, m], v[:, :, m])
        out.append(torch.stack([wo, we - wo], -1))
    return torch.stack(out, 1)


def rbk_forward(p: Dict[str, torch.Tensor], rays, images_idx, num_motion=4, rv_window=0.1, depth=4):
    """Rigid_Blurring_Kernel.forward (lushnerf.py:118-153), poster config (no skip ever triggers).
    p uses the state_dict names below `mlp_rbk.` / `blur_kernel_net.RBK.`."""
    h = F.embedding(images_idx.reshape(-1), p["view_embedding_layer.view_embed_layer.weight"])
    for i in range(depth):
        h = F.relu(F.linear(h, p[f"view_embed_linears.{i}.weight"], p[f"view_embed_linears.{i}.bias"]))
    hr = F.relu(F.linear(h, p["r_branch.0.weight"], p["r_branch.0.bias"]))
    hv = F.relu(F.linear(h, p["v_branch.0.weight"], p["v_branch.0.bias"]))
    hw = F.relu(F.linear(h, p["w_branch.0.weight"], p["w_branch.0.bias"]))
    r = F.linear(hr, p["r_linear.weight"], p["r_linear.bias"]) * rv_window
    v = F.linear(hv, p["v_linear.weight"], p["v_linear.bias"]) * rv_window
    w = torch.sigmoid(F.linear(hw, p["w_linear.weight"], p["w_linear.bias"]))
    w = w / (torch.sum(w, -1, keepdim=True) + 1e-10)
    N = rays.shape[0]
    new_rays = rbk_warp(rays, r.reshape(N, 3, num_motion), v.reshape(N, 3, num_motion))
    return new_rays.reshape(-1, 3, 2), w


def rbk_param_shapes(num_img=20, embed=64, width=64, depth=4, bw=32, num_motion=4):
    s = [("view_embedding_layer.view_embed_layer.weight", (num_img, embed))]
    for i in range(depth):
        s += [(f"view_embed_linears.{i}.weight", (width, embed if i == 0 else width)),
              (f"view_embed_linears.{i}.bias", (width,))]
    for b, oc in (("r", 3 * num_motion), ("v", 3 * num_motion), ("w", num_motion + 1)):
        s += [(f"{b}_branch.0.weight", (bw, width)), (f"{b}_branch.0.bias", (bw,)),
              (f"{b}_linear.weight", (oc, bw)), (f"{b}_linear.bias", (oc,))]
    return s


def det_rbk_params(tag: int, head_scale: float = 0.3, **kw) -> Dict[str, torch.Tensor]:
    """Deterministic RBK weights.  The reference initialises the r/v heads to +-1e-5 (lushnerf.py:62-68),
    which would make the warp a no-op at fp32 resolution; head_scale makes the test meaningful."""
    out = {}
    for j, (name, shape) in enumerate(rbk_param_shapes(**kw)):
        if name.endswith("view_embed_layer.weight"):
            out[name] = det_uniform(shape, tag * 1000 + j, -1.0, 1.0)
            continue
        fan_in = shape[1] if len(shape) == 2 else out[name.replace("bias", "weight")].shape[1]
        b = 1.0 / math.sqrt(fan_in)
        if name.startswith(("r_linear", "v_linear")):
            b *= head_scale
        out[name] = det_uniform(shape, tag * 1000 + j, -b, b)
    return out


# ----------------------------------------------------------------------------------------------
# a-11 K-ray blend + noise + tone map  (models/lushnerf.py:100-116, 649-654; helpers:164-174)
# ----------------------------------------------------------------------------------------------

def tonemap(x, kind="gamma"):
    return x ** (1.0 / 2.2) if kind == "gamma" else x


def blend_tonemap(rgb, ccw, noise_raw, K: int, kind="gamma"):
    """rgb [N*K,3], ccw [N,K] or None -> (tm(blend+noise), tm(blend), noise_rgb)."""
    blended = rgb if ccw is None else torch.sum(rgb.reshape(-1, K, 3) * ccw[..., None], dim=1)
    nz = 0.1 * torch.sigmoid(noise_raw) if noise_raw is not None else None
    blur = tonemap(blended + nz, kind) if nz is not None else tonemap(blended, kind)
    return blur, tonemap(blended, kind), nz


def blur_loss(rgb_blur, rgb0_blur, target):
    """Stage-1/2 training loss (run_lushnerf.py:652-661 with utils/run_lushnerf_helpers.py:9-10):
    0.5*img2mse + 0.5*img2l1 of the fine and of the coarse blurred colours."""
    mse = lambda x, y: torch.mean((x - y) ** 2)
    l1 = lambda x, y: torch.mean(torch.abs(x - y))
    return (0.5 * mse(rgb_blur, target) + 0.5 * l1(rgb_blur, target)
            + 0.5 * mse(rgb0_blur, target) + 0.5 * l1(rgb0_blur, target))


# ----------------------------------------------------------------------------------------------
# a-12 NeRFAll.forward, training branches  (models/lushnerf.py:619-662)
# ----------------------------------------------------------------------------------------------

def nerfall_train_forward(nets, rbk_params, cfg, rays, images_idx, H, W, focal, *, N_samples, N_importance,
                          force_naive=False, allkernel=False, kernel_pixel=None, t_rand=None, noise_c=None,
                          noise_f=None, u=None, sample_order="torch"):
    """Returns the reference's 7-tuple plus a dict of intermediates.
    RBK mode (lushnerf.py:636-654): warp -> (optional detach mask) -> scene render of 5N rays ->
    noise MLP on the N original rays -> blend -> tone map.  Naive mode (:657-662)."""
    if force_naive or rbk_params is None:
        rb = ray_prep(rays, H, W, focal)
        ret, ret_noise = render_rays(nets, cfg, rb, N_samples=N_samples, N_importance=N_importance,
                                     t_rand=t_rand, noise_c=noise_c, noise_f=noise_f, u=u,
                                     with_noise_mlp=True, sample_order=sample_order)
        nz = 0.1 * torch.sigmoid(ret_noise["rgb_map"])
        return (tonemap(ret["rgb_map"]), tonemap(ret["rgb0"]), {}, nz, nz, {}, {}), ret
    new_rays, ccw = rbk_forward(rbk_params, rays, images_idx)
    K = ccw.shape[-1]
    if allkernel:
        m = kernel_pixel.repeat_interleave(K)[:, None, None]
        new_rays = torch.where(m, new_rays, new_rays.detach())                          # :641-643
    rb = ray_prep(new_rays, H, W, focal)
    ret, _ = render_rays(nets, cfg, rb, N_samples=N_samples, N_importance=N_importance, t_rand=t_rand,
                         noise_c=noise_c, noise_f=noise_f, u=u, with_noise_mlp=False,
                         sample_order=sample_order)
    rb0 = ray_prep(rays, H, W, focal).detach()
    z16 = stratified_z(rb0[:, 6:7], rb0[:, 7:8], N_samples)[:, 16:17]
    nraw = mlp_on_rays(nets["noise"], rb0, z16, cfg["D"] // 2, cfg["W"] // 2, cfg["L_pts"], cfg["L_views"],
                       cfg["skip"], out_mode=1)[:, 0, :]
    blur, pure, nz = blend_tonemap(ret["rgb_map"], ccw, nraw, K)
    blur0, pure0, _ = blend_tonemap(ret["rgb0"], ccw, nraw, K)
    ret["ccw"] = ccw
    ret["new_rays"] = new_rays
    return (blur, blur0, {}, nz, nz, pure, pure0), ret


# ----------------------------------------------------------------------------------------------
# deterministic model state shared by the reference, the oracle and the CUDA path in tests
# ----------------------------------------------------------------------------------------------

POSTER_CFG = dict(D=8, W=256, L_pts=10, L_views=4, skip=4)


def det_nets(cfg=POSTER_CFG, tag: int = 1, sigma_gain: float = 1.0):
    """{"coarse","fine","noise"} deterministic parameter dicts.  sigma_gain scales alpha_linear so the
    densities are peaky like a trained scene (SURVEY.md 8d uses x30)."""
    ic, icv = 3 + 6 * cfg["L_pts"], 3 + 6 * cfg["L_views"]
    nets = {
        "coarse": det_mlp_params(cfg["D"], cfg["W"], ic, icv, cfg["skip"], tag * 10 + 1),
        "fine": det_mlp_params(cfg["D"], cfg["W"], ic, icv, cfg["skip"], tag * 10 + 2),
        "noise": det_mlp_params(cfg["D"] // 2, cfg["W"] // 2, ic, icv, cfg["skip"], tag * 10 + 3),
    }
    for k in ("coarse", "fine"):
        nets[k]["alpha_linear.weight"] = nets[k]["alpha_linear.weight"] * sigma_gain
    return nets


def reference_state_dict(nets, rbk_params=None):
    """Keys as NeRFAll.state_dict() names them (checkpoint row)."""
    sd = {}
    for net, prefix in (("coarse", "mlp_coarse."), ("fine", "mlp_fine."), ("noise", "mlp_noise_coarse.")):
        if nets.get(net) is not None:
            sd.update({prefix + k: v for k, v in nets[net].items()})
    if rbk_params is not None:
        sd.update({"mlp_rbk." + k: v for k, v in rbk_params.items()})
    return sd
