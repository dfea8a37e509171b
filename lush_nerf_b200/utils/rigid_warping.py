"""SE(3) rigid warp used by the rigid blurring kernel -- drop-in for the reference's
utils/rigid_warping.py (SE3Field :7-45, Rigid_body :49-140), backed by the fused sm_100a kernel
csrc/rays.cu::rbk_warp_{fwd,bwd}_kernel (exp map + point transform in one launch, closed-form backward).
"""
import torch

from .. import ops


class Rigid_body:
    """Closed-form SE(3) helpers with the reference's names and argument meaning (utils/rigid_warping.py:49-140).
    Small tensor utilities for callers that want the matrices themselves; the hot path never builds them -- it calls
    SE3Field.warp / ops.rbk_warp, where exp map and point transform are one kernel."""

    @staticmethod
    def skew(w):
        """[..., 3] -> [..., 3, 3] cross-product matrix, W @ v == w x v  (reference :93-118)."""
        z = torch.zeros_like(w[..., 0])
        return torch.stack([torch.stack([z, -w[..., 2], w[..., 1]], -1),
                            torch.stack([w[..., 2], z, -w[..., 0]], -1),
                            torch.stack([-w[..., 1], w[..., 0], z], -1)], -2)

    def exp_so3(self, w, theta):
        """Rodrigues: unit axis w [N,3], angle theta [N] -> R [N,3,3] = I + sin(t) W + (1 - cos(t)) W^2  (:76-91)."""
        W = self.skew(w)
        t = theta[..., None, None]
        eye = torch.eye(3, dtype=w.dtype, device=w.device)
        return eye + torch.sin(t) * W + (1.0 - torch.cos(t)) * (W @ W)

    def exp_se3(self, S, theta):
        """Screw axis S = (w, v) [N,6], magnitude theta [N] -> homogeneous transform [N,4,4]  (:53-74):
        R = exp_so3(w, theta), p = (theta I + (1 - cos t) W + (t - sin t) W^2) v."""
        w, v = torch.split(S, [3, 3], dim=-1)
        W = self.skew(w)
        R = self.exp_so3(w, theta)
        t = theta[..., None, None]
        eye = torch.eye(3, dtype=S.dtype, device=S.device)
        p = (t * eye + (1.0 - torch.cos(t)) * W + (t - torch.sin(t)) * (W @ W)) @ v[..., None]
        return self.rp_to_se3(R, p)

    @staticmethod
    def rp_to_se3(R, p):
        """R [N,3,3], p [N,3,1] -> [N,4,4]  (:120-134)."""
        top = torch.cat([R, p], -1)
        bottom = torch.zeros_like(top[..., :1, :])
        bottom[..., 0, 3] = 1.0
        return torch.cat([top, bottom], dim=-2)

    @staticmethod
    def to_homogenous(v):
        return torch.cat([v, torch.ones_like(v[..., :1])], -1)

    @staticmethod
    def from_homogenous(v):
        return v[..., :3] / v[..., -1:]


class SE3Field:
    def __init__(self):
        self.rigid_body = Rigid_body()

    def warp(self, pts, rot, trans):
        """pts, rot, trans: [N,3] -> warped pts [N,3]  (reference :20-45).
        Implemented with the ray-warp kernel: a ray (o=pts, d=0) warped by one motion."""
        rays = torch.stack([pts, torch.zeros_like(pts)], -1)
        out = ops.rbk_warp(rays, rot[..., None], trans[..., None])
        return out[:, 1, :, 0]
