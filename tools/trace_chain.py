#!/usr/bin/env python
"""Timeline of one tile of the fused forward chain kernel (CTA 0, third tile): LUSH_CHAIN_DEBUG=2 python tools/trace_chain.py"""
import ctypes, os, sys
os.environ["LUSH_CHAIN_DEBUG"] = os.environ.get("LUSH_CHAIN_DEBUG", "2")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from lush_nerf_b200 import _lib, ops
from lush_nerf_b200.synthetic import scene_mlp_for_tools
DEV = "cuda"
R, S = 8192, 128
flat, views, desc = scene_mlp_for_tools(DEV)
rb = torch.zeros(R, 11, device=DEV); rb[:, :3] = torch.rand(R, 3, device=DEV); rb[:, 3:6] = torch.randn(R, 3, device=DEV) * 0.3
rb[:, 8:11] = torch.nn.functional.normalize(torch.randn(R, 3, device=DEV), dim=-1)
z = torch.rand(R, S, device=DEV).sort(-1)[0]
with torch.no_grad():
    for _ in range(2):
        ops.mlp_apply(rb, z, desc, 1, flat, views)
torch.cuda.synchronize()
n = 12 * 16
buf = (ctypes.c_longlong * n)()
_lib.check(_lib.lib().lush_debug_read_chain_trace(buf, n), "trace")
t = np.array(buf[:], dtype=np.int64).reshape(12, 16)
t0 = t[0][t[0] > 0].min()
names = ["i0.start", "i0.end", "i1.start", "i1.end", "w0.drdy", "w0.rel", "w0.pub0", "w0.pub1", "w1.drdy", "w1.rel", "w1.pub0", "w1.pub1"]
print("step " + " ".join(f"{x:>10s}" for x in names))
for s in range(11):
    print(f"{s:4d} " + " ".join(f"{(t[s][k] - t0) if t[s][k] > 0 else -1:10d}" for k in range(12)))
