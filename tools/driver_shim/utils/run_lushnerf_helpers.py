"""`utils.run_lushnerf_helpers` as the reference driver imports it (run_lushnerf.py:11, star import)."""
from lush_nerf_b200.utils.run_lushnerf_helpers import *  # noqa: F401,F403
from lush_nerf_b200.utils.run_lushnerf_helpers import (HALF_PIX, compute_mean_with_confidence, get_masks_for_images,  # noqa: F401
                                                        get_rays, get_rays_np, img2l1, img2mse, mse2psnr, ndc_rays,
                                                        sample_pdf, smart_load_state_dict, to8b)
import numpy as np                                        # noqa: F401
import torch                                              # noqa: F401
import torch.nn as nn                                     # noqa: F401
