"""CPU: the host-side remnants of the reference's call surface that the driver star-imports
(run_lushnerf.py:9-11) -- frequency masks, confidence-masked mean, SE(3) closed forms -- against goldens produced by
the unmodified reference (tests/golden/make_golden_api.py)."""
import os

import numpy as np
import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def gapi():
    d = np.load(os.path.join(HERE, "golden", "golden_api.npz"))
    return {k: torch.from_numpy(np.asarray(d[k])) for k in d.files}


def test_frequency_masks_vs_reference(gapi):
    """numpy-FFT restatement of helpers:630-662 (cv2.dft upstream).  The mask is a threshold of a float image: pixels
    whose value sits within rounding of the threshold may differ -- at most 0.5 % of them."""
    from lush_nerf_b200.utils.run_lushnerf_helpers import frequency_mask, get_masks_for_images
    from tests.golden.make_golden_api import mask_images
    imgs = mask_images()
    for rad in (4, 10):
        got = get_masks_for_images(imgs, radius=rad)
        want = gapi[f"masks/r{rad}"].numpy()
        assert got.shape == want.shape == imgs.shape[:-1] and got.dtype == np.uint8
        assert set(np.unique(got)) <= {0, 1}
        assert (got != want).mean() <= 0.005, (rad, (got != want).mean())
        assert 0.05 < got.mean() < 0.999                          # not a trivial mask
    assert frequency_mask(imgs[0], 4).shape == imgs.shape[1:3]


def test_compute_mean_with_confidence_vs_reference(gapi):
    from lush_nerf_b200.utils.run_lushnerf_helpers import compute_mean_with_confidence
    A, c = gapi["mean/A"], gapi["mean/c"]
    assert torch.allclose(compute_mean_with_confidence(A, c, 0.4), gapi["mean/out"], atol=1e-7)
    c0 = c.clone()
    c0[:, :5] = 0.0
    out = compute_mean_with_confidence(A, c0, 0.4)
    assert torch.allclose(out, gapi["mean/out_some_empty"], atol=1e-7)
    assert (out[:5] == 0).all()


def test_rigid_body_closed_forms_vs_reference(gapi):
    from lush_nerf_b200.utils.rigid_warping import Rigid_body
    rb = Rigid_body()
    w, v, th = gapi["rigid/w"], gapi["rigid/v"], gapi["rigid/theta"]
    assert torch.equal(rb.skew(w), gapi["rigid/skew"])
    assert torch.allclose(rb.exp_so3(w, th), gapi["rigid/exp_so3"], atol=1e-6)
    X = rb.exp_se3(torch.cat([w, v], -1), th)
    assert X.shape == (9, 4, 4)
    assert torch.allclose(X, gapi["rigid/exp_se3"], atol=1e-6)
    R = X[:, :3, :3]
    assert torch.allclose(R @ R.transpose(1, 2), torch.eye(3).expand(9, 3, 3), atol=1e-5)   # a rotation
    assert torch.equal(X[:, 3], torch.tensor([0., 0., 0., 1.]).expand(9, 4))
    assert torch.equal(rb.from_homogenous(rb.to_homogenous(v)), v)


def test_guard_flag_bits_are_disjoint():
    from lush_nerf_b200 import _lib
    from lush_nerf_b200.models.lushnerf import _GUARD_KEYS
    numeric = 0
    for _, sh in _GUARD_KEYS:
        numeric |= 3 << sh
    idx = _lib.FLAG_BAD_IMAGE_INDEX | _lib.FLAG_BAD_VIEW_INDEX | _lib.FLAG_BAD_PIXEL_INDEX
    assert numeric & idx == 0 and bin(idx).count("1") == 3
    hdr = open(os.path.join(os.path.dirname(HERE), "include", "lushnerf_b200.h")).read()
    for name, bit in (("LUSH_FLAG_BAD_IMAGE_INDEX", 10), ("LUSH_FLAG_BAD_VIEW_INDEX", 11), ("LUSH_FLAG_BAD_PIXEL_INDEX", 12)):
        assert f"#define {name} (1 << {bit})" in hdr
