#!/usr/bin/env python
"""Run the tcgen05 forward once on R x S samples and compare with the fp32 SIMT path (debug helper)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from lush_nerf_b200 import _lib, ops
from lush_nerf_b200.synthetic import scene_mlp_for_tools
R, S = int(sys.argv[1]), int(sys.argv[2])
DEV = "cuda"
flat, views, desc = scene_mlp_for_tools(DEV)
g = torch.Generator(device=DEV).manual_seed(1)
rb = torch.zeros(R, 11, device=DEV); rb[:, :3] = torch.rand(R, 3, device=DEV, generator=g) * 2 - 1
rb[:, 3:6] = torch.randn(R, 3, device=DEV, generator=g) * 0.3
rb[:, 8:11] = torch.nn.functional.normalize(torch.randn(R, 3, device=DEV, generator=g), dim=-1)
z = torch.rand(R, S, device=DEV, generator=g).sort(-1)[0]
with torch.no_grad():
    a = ops.mlp_apply(rb, z, desc, 1, flat, views)
    torch.cuda.synchronize()
    b = ops.mlp_apply(rb, z, desc, 0, flat, views)
    torch.cuda.synchronize()
print(R, S, "max abs diff", float((a - b).abs().max()), "scale", float(b.abs().max()))
