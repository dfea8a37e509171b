"""Small fixed workload for ncu: one tcgen05 forward launch (R=8192 rays x 128 samples) after two warm-ups."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from lush_nerf_b200 import _lib, ops
from lush_nerf_b200.synthetic import scene_mlp_for_tools
DEV = "cuda"
mode = sys.argv[1] if len(sys.argv) > 1 else "fwd"
R, S = 8192, 128
flat, views, desc = scene_mlp_for_tools(DEV, requires_grad=(mode != "fwd"))
rb = torch.zeros(R, 11, device=DEV)
rb[:, :3] = torch.rand(R, 3, device=DEV) * 2 - 1
rb[:, 3:6] = torch.randn(R, 3, device=DEV) * 0.3
rb[:, 8:11] = torch.nn.functional.normalize(torch.randn(R, 3, device=DEV), dim=-1)
z = torch.rand(R, S, device=DEV).sort(-1)[0]
for _ in range(3):
    if mode == "fwd":
        with torch.no_grad():
            out = ops.mlp_apply(rb, z, desc, 1, flat, views)
    else:
        rbg = rb.clone().requires_grad_(True)
        out = ops.mlp_apply(rbg, z, desc, 1, flat, views)
        out.sum().backward()
torch.cuda.synchronize()
print("done", float(out.abs().mean()))
