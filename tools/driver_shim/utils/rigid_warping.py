from lush_nerf_b200.utils.rigid_warping import *         # noqa: F401,F403
