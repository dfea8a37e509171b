"""SE(3) rigid warp used by the rigid blurring kernel -- drop-in for the reference's
utils/rigid_warping.py (SE3Field :7-45, Rigid_body :49-140), backed by the fused sm_100a kernel
csrc/rays.cu::rbk_warp_{fwd,bwd}_kernel (exp map + point transform in one launch, closed-form backward).
"""
import torch

from .. import ops


class Rigid_body:
    """Closed-form helpers kept for API compatibility (reference utils/rigid_warping.py:49-140).
    They are small host-visible utilities; the hot path calls SE3Field.warp / ops.rbk_warp."""

    @staticmethod
    def skew(w):
        z = torch.zeros_like(w[..., 0])
        return torch.stack([torch.stack([z, -w[..., 2], w[..., 1]], -1),
                            torch.stack([w[..., 2], z, -w[..., 0]], -1),
                            torch.stack([-w[..., 1], w[..., 0], z], -1)], -2)

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
