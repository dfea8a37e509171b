#!/usr/bin/env python
"""Statistics of one rendered frame (eval path) -- a diagnostic for all-black / NaN frames."""
import os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lush_nerf_b200.synthetic import build_model, poster_args, spiral_poses

def stats(name, t):
    t = t.detach().float()
    print(f"{name:14s} shape {tuple(t.shape)} finite {bool(torch.isfinite(t).all())} nan {int(torch.isnan(t).sum())} "
          f"min {t[torch.isfinite(t)].min().item() if torch.isfinite(t).any() else None} max {t[torch.isfinite(t)].max().item() if torch.isfinite(t).any() else None} mean {t[torch.isfinite(t)].mean().item() if torch.isfinite(t).any() else None}")

def main():
    os.environ.setdefault("LUSH_VERBOSE", "0")
    os.environ["LUSH_NUMERIC_GUARD"] = "sync"
    dev = torch.device("cuda")
    torch.manual_seed(0)
    m = build_model(poster_args()).to(dev).eval()
    if len(sys.argv) > 1:
        ck = torch.load(sys.argv[1], map_location="cpu")
        sd = {k[len("module."):] if k.startswith("module.") else k: v for k, v in ck["network_state_dict"].items()}
        print("load:", m.load_state_dict(sd, strict=False))
    H, W, f = 400, 600, 500.0
    K = np.array([[f, 0, 0.5 * W], [0, f, 0.5 * H], [0, 0, 1]])
    poses = torch.from_numpy(spiral_poses(3, radius_scale=2.0)).to(dev)
    rk = dict(perturb=False, N_importance=64, N_samples=64, use_viewdirs=True, white_bkgd=False, raw_noise_std=0.0,
              inference=True, near=0.0, far=1.0)
    with torch.no_grad():
        for impl in (1, 0):
            for net in (m.mlp_coarse, m.mlp_fine):
                net.mlp_impl = impl
            rgbs, nz, depth = m(H, W, K, chunk=1024 * 32, poses=poses[:1], render_kwargs=dict(rk))
            print("impl", impl)
            stats("rgbs(tm)", rgbs); stats("noise(tm)", nz); stats("depth", depth)
            rb = __import__("lush_nerf_b200").ops.ray_gen_prep(poses[0][:3, :4], H, W, K)
            ret, rn = m.render_rays(rb[:32768], 64, retraw=True, perturb=0.0, N_importance=64, raw_noise_std=0.0, inference=True)
            for k in ("rgb_map", "acc_map", "depth_map", "raw", "rgb0", "acc0", "z_std"):
                stats(k, ret[k])

if __name__ == "__main__":
    main()
