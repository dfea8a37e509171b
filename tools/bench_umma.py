#!/usr/bin/env python
"""tcgen05.mma issue / dependency micro-benchmark on one SM (cycles per MMA, M=128 K=16, blocks of 8 per elect)."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from lush_nerf_b200 import _lib
L = _lib.lib()
out = torch.zeros(4, dtype=torch.int64, device="cuda")
for N, nacc, a_tmem in ((256, 1, 1), (128, 2, 1), (64, 1, 1)):
    for flags in (0, 7, 8 + 7, 16 + 8 + 7):
        iters = 512
        _lib.check(L.lush_tc_microbench(ctypes.c_void_p(out.data_ptr()), N, nacc, a_tmem, iters, flags, None), "mb")
        torch.cuda.synchronize()
        a, b = out.tolist()[:2]
        print(f"N={N:3d} nacc={nacc} A={'tmem' if a_tmem else 'smem'} flags={flags} (1 commit, 2 fence, 4 wait per 8 MMAs; 8 busy neighbours, 16 issuer = warp 15): "
              f"issue {a/iters:6.1f} cyc/mma, complete {b/iters:6.1f} cyc/mma (floor {128*N/256:.0f})")
