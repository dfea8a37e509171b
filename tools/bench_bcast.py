#!/usr/bin/env python
"""Weight-delivery rate of the chain kernels' access pattern: all CTAs stream the same L2-resident image through a
3-stage ring, unicast vs cluster multicast (lush_tc_bcast_bench)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from lush_nerf_b200 import _lib
from ctypes import c_void_p
L = _lib.lib()
src = torch.randint(0, 255, (40 * 32768,), dtype=torch.uint8, device="cuda")
for block in (32768, 16384):
    nblocks = src.numel() // block
    for cluster, grid in ((1, 148), (2, 148), (4, 144), (1, 74), (1, 144)):
        out = torch.zeros(grid, dtype=torch.int64, device="cuda")
        for rep in range(2):
            _lib.check(L.lush_tc_bcast_bench(c_void_p(src.data_ptr()), nblocks, block, 20, cluster, grid,
                                             c_void_p(out.data_ptr()), _lib.stream()), "bcast")
            torch.cuda.synchronize()
        cyc = out.float()
        total = nblocks * 20 * block
        print(f"block {block:6d} cluster {cluster} grid {grid:3d}: cycles mean {cyc.mean():10.0f} max {cyc.max():10.0f}  "
              f"-> {total / cyc.mean():6.1f} B/clk per SM delivered, {total * grid / cyc.max():8.0f} B/clk chip")
