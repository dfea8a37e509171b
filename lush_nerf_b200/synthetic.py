"""Synthetic LLFF-shaped scene: V views of HxW, focal f, forward-facing poses on a small
arc, NDC near/far = 0/1, uniform-random targets.  Stands in for load_llff_data's output contract
(reference data_utils/load_llff.py:337-407) -- there is no dataset in this environment."""
from __future__ import annotations

import types

import numpy as np
import torch

from .utils.run_lushnerf_helpers import get_rays_np


def poster_args(**overrides):
    """Fields NeRFAll reads (reference models/lushnerf.py:183-231,331-333) at configs/poster_lushnerf values."""
    ns = types.SimpleNamespace(
        blur_model_type="dpnerf", multires=10, i_embed=0, use_viewdirs=True, multires_views=4, N_importance=64,
        N_samples=64, netdepth=8, netwidth=256, netdepth_fine=8, netwidth_fine=256, rgb_activate="sigmoid",
        sigma_activate="relu", tone_mapping_type="gamma", render_rmnearplane=80, num_images=20,
        N_rand=1024, chunk=1024 * 32, lrate=5e-4, raw_noise_std=1.0, perturb=1.0)
    for k, v in overrides.items():
        setattr(ns, k, v)
    return ns


def build_model(args=None, num_img=20, with_rbk=True):
    """RBK + NeRFAll exactly as run_lushnerf.py:332-345 builds them (poster config)."""
    from .models.lushnerf import RBK, NeRFAll
    args = args or poster_args()
    rbk = None
    if with_rbk:
        rbk = RBK(num_img=num_img, view_embed_ch=64, D_rbk=4, W_rbk=64, skips_rbk=[4], num_motion_rbk=4,
                  D_rbk_r=1, W_rbk_r=32, output_ch_rbk_r=3, D_rbk_v=1, W_rbk_v=32, output_ch_rbk_v=3,
                  D_rbk_w=1, W_rbk_w=32, rbk_se_rv_window=0.1, use_dpnerf=True, rbk_use_origin=True,
                  near=0.0, far=1.0, ndc=True)
    return NeRFAll(args, rbk)


def synthetic_poses(V=20, seed=0):
    """[V,3,4] camera-to-world: small rotations/translations around identity, LLFF axis convention."""
    rng = np.random.RandomState(seed)
    poses = []
    for i in range(V):
        ang = (i / max(V - 1, 1) - 0.5) * 0.2
        jit = rng.uniform(-0.02, 0.02, 3)
        ry = np.array([[np.cos(ang + jit[0]), 0, np.sin(ang + jit[0])], [0, 1, 0],
                       [-np.sin(ang + jit[0]), 0, np.cos(ang + jit[0])]])
        rx = np.array([[1, 0, 0], [0, np.cos(jit[1]), -np.sin(jit[1])], [0, np.sin(jit[1]), np.cos(jit[1])]])
        t = np.array([np.sin(ang) * 0.5, jit[2], 0.05 * np.cos(ang)])
        poses.append(np.concatenate([ry @ rx, t[:, None]], 1))
    return np.stack(poses).astype(np.float32)


def spiral_poses(n=120, radius_scale=2.0, seed=0):
    """Render path standing in for render_path_spiral (load_llff.py:234) with render_radius_scale."""
    out = []
    for i in range(n):
        th = 2 * np.pi * i / n
        c = np.array([0.1 * radius_scale * np.cos(th), 0.1 * radius_scale * np.sin(th), 0.05 * np.sin(0.5 * th)])
        zax = np.array([0, 0, 1.0]) + 0.3 * np.array([c[0], c[1], 0])
        zax /= np.linalg.norm(zax)
        xax = np.cross([0, 1.0, 0], zax); xax /= np.linalg.norm(xax)
        yax = np.cross(zax, xax)
        out.append(np.stack([xax, yax, zax, c], 1))
    return np.stack(out).astype(np.float32)


class SyntheticLLFF:
    """Ray table as run_lushnerf.py:569-596 builds it: rays [N,2,3], rgbsf [N,3], fq_mask [N],
    images_idx [N,1] for the training views (llffhold=8 -> 17 of 20 views)."""

    def __init__(self, H=400, W=600, focal=500.0, V=20, llffhold=8, seed=0, device="cpu", max_views=None):
        self.H, self.W, self.focal = H, W, focal
        self.K = np.array([[focal, 0, 0.5 * W], [0, focal, 0.5 * H], [0, 0, 1]])
        poses = synthetic_poses(V, seed)
        i_test = np.arange(V)[::llffhold]
        i_train = np.array([i for i in range(V) if i not in i_test])
        if max_views:
            i_train = i_train[:max_views]
        self.poses, self.i_train, self.num_img = poses, i_train, V
        rays = np.stack([np.stack(get_rays_np(H, W, self.K, p), 0) for p in poses[i_train]], 0)  # [v,2,H,W,3]
        rays = np.transpose(rays, [0, 2, 3, 1, 4]).reshape(-1, 2, 3).astype(np.float32)
        g = torch.Generator().manual_seed(seed)
        n = rays.shape[0]
        perm = torch.randperm(n, generator=g)
        self.rays = torch.from_numpy(rays)[perm].to(device)
        self.rgbsf = torch.rand(n, 3, generator=g)[perm].to(device)
        self.fq_mask = (torch.rand(n, generator=g) < 0.5)[perm].to(device)
        idx = torch.from_numpy(np.repeat(i_train, H * W).astype(np.int64)).reshape(-1, 1)
        self.images_idx = idx[perm].to(device)
        self.n = n

    def batch(self, i, N_rand):
        s = (i * N_rand) % max(self.n - N_rand, 1)
        sl = slice(s, s + N_rand)
        return dict(rays=self.rays[sl].permute(0, 2, 1), rgbsf=self.rgbsf[sl], fq_mask=self.fq_mask[sl],
                    images_idx=self.images_idx[sl])


def scene_mlp_for_tools(device="cuda", seed=0, requires_grad=False):
    """(flat, views, desc) of a randomly initialised 8x256 scene MLP (reference initialisers) for the profiling /
    benchmarking tools under tools/: the flat fp32 arena, the parameter views in state_dict order, the MlpDesc."""
    import torch
    from .utils.run_lushnerf_helpers import NeRF
    torch.manual_seed(seed)
    net = NeRF(D=8, W=256, input_ch=63, input_ch_views=27, output_ch=4, skips=[4], use_viewdirs=True).to(device)
    flat = net.flat_params()
    views = list(net._ordered_params())
    for v in views:
        v.requires_grad_(requires_grad)
    return flat, views, net._desc
