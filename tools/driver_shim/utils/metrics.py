"""Stand-in for the reference's utils/metrics.py (skimage / lpips are not in this image; eval-only image metrics are out
of scope): mse and psnr are computed, ssim / lpips return 0."""
import torch


def compute_img_metric(im1t, im2t, metric="mse", margin=0, mask=None):
    a, b = torch.as_tensor(im1t).float(), torch.as_tensor(im2t).float()
    mse = torch.mean((a - b) ** 2).item()
    if metric == "mse":
        return mse
    if metric == "psnr":
        return float(-10.0 * torch.log10(torch.tensor(max(mse, 1e-12))))
    return 0.0
