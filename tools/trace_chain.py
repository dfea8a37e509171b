#!/usr/bin/env python
"""Timeline of one tile of the fused chain kernels (CTA 0, third tile).  Needs a library built with the probes:
    LUSH_CHAIN_TRACE=1 python -m lush_nerf_b200.build --force
    python tools/trace_chain.py [infer|train]       # train: forward with stash + dgrad
"""
import ctypes, os, sys
os.environ["LUSH_CHAIN_DEBUG"] = os.environ.get("LUSH_CHAIN_DEBUG", "2")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from lush_nerf_b200 import _lib, ops
from lush_nerf_b200.synthetic import scene_mlp_for_tools
DEV = "cuda"
mode = sys.argv[1] if len(sys.argv) > 1 else "infer"
R, S = 8192, 128
flat, views, desc = scene_mlp_for_tools(DEV, requires_grad=(mode == "train"))
rb = torch.zeros(R, 11, device=DEV); rb[:, :3] = torch.rand(R, 3, device=DEV); rb[:, 3:6] = torch.randn(R, 3, device=DEV) * 0.3
rb[:, 8:11] = torch.nn.functional.normalize(torch.randn(R, 3, device=DEV), dim=-1)
z = torch.rand(R, S, device=DEV).sort(-1)[0]
for _ in range(2):
    if mode == "train":
        out = ops.mlp_apply(rb, z, desc, 1, flat, views)
        out.backward(torch.randn_like(out))
    else:
        with torch.no_grad():
            ops.mlp_apply(rb, z, desc, 1, flat, views)
torch.cuda.synchronize()
names = ["i0.start", "i0.end", "i1.start", "i1.end", "w0.drdy", "w0.rel", "pub0", "pub1", "w1.drdy", "w1.rel", "pub2", "pub3",
         "L.reach", "L.full", "L.aready", "L.issued"]


def show(fn, title):
    n = 14 * 16
    buf = (ctypes.c_longlong * n)()
    _lib.check(fn(buf, n), "trace")
    t = np.array(buf[:], dtype=np.int64).reshape(14, 16)
    if not (t > 0).any():
        print(title, ": no trace recorded (library built without LUSH_CHAIN_TRACE=1?)")
        return
    t0 = t[t > 0].min()
    print(title)
    print("step " + " ".join(f"{x:>10s}" for x in names))
    for s in range(14):
        if (t[s] > 0).any():
            print(f"{s:4d} " + " ".join(f"{(t[s][k] - t0) if t[s][k] > 0 else -1:10d}" for k in range(16)))


show(_lib.lib().lush_debug_read_chain_trace, f"forward chain ({mode})")
if mode == "train":
    show(_lib.lib().lush_debug_read_chain_trace_bwd, "dgrad chain")
