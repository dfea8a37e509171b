"""`models.lushnerf` as the reference driver imports it (run_lushnerf.py:9 `from models.lushnerf import *`): every
public name of the B200-native drop-in, including the `torch` / `np` / `nn` names the driver picks up through the star
import (it has no imports of its own for them)."""
from lush_nerf_b200.models.lushnerf import *            # noqa: F401,F403
from lush_nerf_b200.models.lushnerf import NeRFAll, RBK, Rigid_Blurring_Kernel, View_Embedding  # noqa: F401
import numpy as np                                       # noqa: F401
import torch                                             # noqa: F401
import torch.nn as nn                                    # noqa: F401
import torch.nn.functional as F                          # noqa: F401
