"""Inputs shared by tests/golden/make_golden_stage3.py (runs the reference, build container only) and the GPU test
of the stage-3 consistency path (tests/test_model_gpu.py): constants, the deterministic stand-in for the match
table, the poses.  No reference import here: this module travels to the GPU box."""
import os

import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden_stage3.npz")
H, W, FOCAL, V, NPIX = 40, 60, 50.0, 3, 640 * 1120
SEED = 11
KW = dict(perturb=0.0, N_importance=16, N_samples=24, use_viewdirs=True, white_bkgd=False, raw_noise_std=0.0,
          inference=False, near=0.0, far=1.0)


def stage3_tables():
    """Deterministic stand-in for the GIM/DKM match table: columns 2:4 of align_matrix[a, i, p] are the (x, y) pixel of
    view i matched to pixel p of anchor view a (some outside the image: the reference clamps), align_mask the
    certainty."""
    p = torch.arange(NPIX, dtype=torch.int64)
    am = torch.zeros(V, V, NPIX, 4)
    mk = torch.zeros(V, V, NPIX)
    for a in range(V):
        for i in range(V):
            am[a, i, :, 2] = ((p * 7 + i * 13 + a * 31) % (W + 6) - 3).float() + 0.25
            am[a, i, :, 3] = ((p * 11 + i * 17 + a * 5) % (H + 4) - 2).float() + 0.75
            mk[a, i] = ((p * 3 + i + a) % 10).float() / 10.0
    return am, mk


def stage3_poses():
    from lush_nerf_b200.synthetic import synthetic_poses
    return torch.from_numpy(synthetic_poses(V, seed=3)).float()


