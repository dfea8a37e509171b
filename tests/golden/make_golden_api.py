"""Generate tests/golden/golden_api.npz from the UNMODIFIED reference: the small API functions around the hot path
that golden_v1.npz does not cover -- Rigid_body.exp_so3 / exp_se3 / rp_to_se3 (utils/rigid_warping.py:53-134),
frequency masks (utils/run_lushnerf_helpers.py:630-662), compute_mean_with_confidence (:665-688),
rbk_weighted_sum (models/lushnerf.py:100-116), NeRF / NeRF_Noise .forward on pre-embedded inputs (helpers:394-423,
:483-512) and NeRFAll.mlpforward_noise (models/lushnerf.py:268-293).

    python -m tests.golden.make_golden_api
"""
import os

import numpy as np
import torch

from oracle import lush_oracle as O
from oracle.ref_import import build_reference_model, import_reference, poster_args

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden_api.npz")


def mask_images(seed=0, V=3, H=48, W=64):
    """Smooth random uint8 images (bilinear-upsampled noise + a little pixel noise): deterministic, no cv2."""
    rng = np.random.RandomState(seed)
    out = []
    for _ in range(V):
        base = rng.rand(H // 8 + 2, W // 8 + 2, 3)
        yy = np.linspace(0, base.shape[0] - 1.001, H)
        xx = np.linspace(0, base.shape[1] - 1.001, W)
        y0, x0 = yy.astype(int), xx.astype(int)
        fy, fx = (yy - y0)[:, None, None], (xx - x0)[None, :, None]
        im = (base[y0][:, x0] * (1 - fy) * (1 - fx) + base[y0 + 1][:, x0] * fy * (1 - fx)
              + base[y0][:, x0 + 1] * (1 - fy) * fx + base[y0 + 1][:, x0 + 1] * fy * fx)
        im = im + 0.08 * rng.rand(H, W, 3)
        out.append((np.clip(im, 0, 1) * 255).astype(np.uint8))
    return np.stack(out)


def main():
    ref_model, ref_helpers, ref_warp = import_reference()
    g = torch.Generator().manual_seed(11)
    out = {}
    rb = ref_warp.Rigid_body()
    w = torch.nn.functional.normalize(torch.randn(9, 3, generator=g), dim=-1)
    v = torch.randn(9, 3, generator=g)
    theta = torch.rand(9, generator=g) * 2.0 + 1e-3
    out["rigid/w"], out["rigid/v"], out["rigid/theta"] = w, v, theta
    out["rigid/exp_so3"] = rb.exp_so3(w, theta)
    out["rigid/exp_se3"] = rb.exp_se3(torch.cat([w, v], -1), theta)
    out["rigid/skew"] = rb.skew(w)
    imgs = mask_images()
    for rad in (4, 10):
        out[f"masks/r{rad}"] = torch.from_numpy(ref_helpers.get_masks_for_images(imgs, radius=rad))
    A, c = torch.rand(7, 32, 3, generator=g), torch.rand(7, 32, generator=g)
    out["mean/A"], out["mean/c"] = A, c
    out["mean/out"] = ref_helpers.compute_mean_with_confidence(A, c, 0.4)
    c0 = c.clone()
    c0[:, :5] = 0.0                                   # pixels no view is confident about -> 0
    out["mean/out_some_empty"] = ref_helpers.compute_mean_with_confidence(A, c0, 0.4)

    nets = O.det_nets(O.POSTER_CFG, tag=1, sigma_gain=30.0)
    rbk = O.det_rbk_params(tag=7)
    model = build_reference_model(poster_args(N_importance=16, N_samples=24))
    model.load_state_dict(O.reference_state_dict(nets, rbk), strict=False)
    model.eval()
    N, K = 6, 5
    rgb, depth, acc = torch.rand(N * K, 3, generator=g), torch.rand(N * K, generator=g), torch.rand(N * K, generator=g)
    extras = {"rgb0": torch.rand(N * K, 3, generator=g), "raw": torch.randn(N * K, 7, 4, generator=g),
              "z_std": torch.rand(N * K, generator=g)}
    ccw = torch.softmax(torch.randn(N, K, generator=g), -1)
    for k, t in (("rgb", rgb), ("depth", depth), ("acc", acc), ("ccw", ccw)):
        out[f"blend/{k}"] = t
    for k, t in extras.items():
        out[f"blend/extras/{k}"] = t
    b_rgb, b_depth, b_acc, b_ex = model.mlp_rbk.rbk_weighted_sum(rgb, depth, acc, dict(extras), ccw)
    out["blend/out/rgb"], out["blend/out/depth"], out["blend/out/acc"] = b_rgb, b_depth, b_acc
    for k, t in b_ex.items():
        out[f"blend/out/extras/{k}"] = t
    # networks on pre-embedded inputs and the noise branch's own entry point
    x = torch.rand(50, 3, generator=g) * 2 - 1
    d = torch.nn.functional.normalize(torch.randn(50, 3, generator=g), dim=-1)
    emb = torch.cat([model.embed_fn(x), model.embeddirs_fn(d)], -1)
    out["net/x"], out["net/d"] = x, d
    with torch.no_grad():
        out["net/coarse"] = model.mlp_coarse(emb)
        out["net/noise"] = model.mlp_noise_coarse(emb)
        pts = torch.rand(9, 24, 3, generator=g) * 2 - 1
        vd = torch.nn.functional.normalize(torch.randn(9, 3, generator=g), dim=-1)
        out["mfn/pts"], out["mfn/viewdirs"] = pts, vd
        out["mfn/out"] = model.mlpforward_noise(pts, vd, model.mlp_noise_coarse, True, True)
        out["mf/out"] = model.mlpforward(pts, vd, model.mlp_coarse, True, True)
    np.savez_compressed(OUT, **{k: np.asarray(t.detach().cpu().numpy() if torch.is_tensor(t) else t) for k, t in out.items()})
    print(f"wrote {OUT}: {len(out)} arrays, {os.path.getsize(OUT)} bytes")


if __name__ == "__main__":
    main()
