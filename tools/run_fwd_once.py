#!/usr/bin/env python
"""Run the tcgen05 forward once on R x S samples and compare with the fp32 SIMT path (debug helper)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from lush_nerf_b200 import _lib, ops
from oracle import lush_oracle as O
R, S = int(sys.argv[1]), int(sys.argv[2])
DEV = "cuda"
nets = O.det_nets()
flat = torch.cat([nets["coarse"][n].reshape(-1) for n, _ in O.mlp_param_shapes(8, 256, 63, 27, 4)]).to(DEV)
views, off = [], 0
for _, shp in O.mlp_param_shapes(8, 256, 63, 27, 4):
    n = int(np.prod(shp)); views.append(flat[off:off + n].view(shp)); off += n
desc = _lib.MlpDesc(8, 256, 10, 4, 4, 0)
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
