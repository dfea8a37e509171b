"""Generate tests/golden/golden_stage3.npz by running the UNMODIFIED reference's stage-3 consistency path
(NeRFAll.forward(consist_loss=True) -> Render_Aligned_Pixel, models/lushnerf.py:664-668, :949-988) on CPU.

    python -m tests.golden.make_golden_stage3

The reference calls `.cuda()` on its index / result tensors (:964, :969-970); for a CPU run the method is patched
to the identity for the duration of the call -- nothing else is touched.  The match table (align_matrix
[V, V, 640*1120, 4], align_mask [V, V, 640*1120]; the 640 x 1120 size is hard-coded upstream) is NOT stored: it is
rebuilt from `stage3_tables` below by the test.  Python's and numpy's generators are seeded (anchor view, 32
pixel samples), perturb = 0 and raw_noise_std = 0 so that no torch draws are consumed.
"""
import contextlib
import io
import os
import random

import numpy as np
import torch

from oracle import lush_oracle as O
from oracle.ref_import import build_reference_model, import_reference, poster_args

from tests.golden.stage3_inputs import FOCAL, GOLDEN as OUT, H, KW, SEED, W, stage3_poses, stage3_tables


def main():
    import_reference()
    nets = O.det_nets(O.POSTER_CFG, tag=1, sigma_gain=30.0)
    rbk = O.det_rbk_params(tag=7)
    model = build_reference_model(poster_args(N_importance=16, N_samples=24))
    model.load_state_dict(O.reference_state_dict(nets, rbk), strict=False)
    model.train()
    K = np.array([[FOCAL, 0, 0.5 * W], [0, FOCAL, 0.5 * H], [0, 0, 1]])
    am, mk = stage3_tables()
    poses = stage3_poses()
    random.seed(SEED)
    np.random.seed(SEED)
    orig_cuda = torch.Tensor.cuda
    torch.Tensor.cuda = lambda self, *a, **k: self
    try:
        with contextlib.redirect_stdout(io.StringIO()):
            rgb_align, certain = model(H, W, K, chunk=1024, poses=poses, rays_info=None, consist_loss=True,
                                       Align_matrix=am, Align_mask=mk, render_kwargs=dict(KW))
    finally:
        torch.Tensor.cuda = orig_cuda
    np.savez_compressed(OUT, rgb_align=rgb_align.detach().numpy(), certain=certain.detach().numpy(),
                        seed=np.array(SEED))
    print("wrote", OUT, rgb_align.shape, certain.shape, float(rgb_align.abs().mean()))


if __name__ == "__main__":
    main()
