#!/usr/bin/env python
"""One launch of each HBM-bound kernel of the stress sweep (65 536 rays, S = N = 128, LLFF-like peaky densities) for
    ncu --set full --clock-control none --import-source on -k regex:"composite|sample_" -o gpurun_out/r2_ncu_stress \\
        python tools/ncu_stress_target.py
Summarise with tools/ncu_top.py."""
import os
import sys
from ctypes import c_void_p

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lush_nerf_b200 import _lib, ops  # noqa: E402

R, S, dev = 65536, int(sys.argv[1]) if len(sys.argv) > 1 else 128, "cuda"
g = torch.Generator(device=dev).manual_seed(0)
raw = torch.randn(R, S, 4, device=dev, generator=g)
raw[..., 3] = torch.relu(3 * torch.randn(R, S, device=dev, generator=g)) * 5
z = (torch.linspace(0, 1, S, device=dev).expand(R, S) + torch.rand(R, S, device=dev, generator=g) * (0.5 / S)).contiguous()
rd = torch.randn(R, 3, device=dev, generator=g)
nz = torch.randn(R, S - 1, device=dev, generator=g)
flush = torch.empty(512 * 1024 * 1024, dtype=torch.uint8, device=dev)
with torch.no_grad():
    for _ in range(2):                                    # the second round is the one to look at (first: cold code)
        flush.zero_()
        out = ops.composite(raw, z, rd, nz)
        g_rgb, g_raw = torch.randn(R, 3, device=dev, generator=g), torch.empty_like(raw)
        flush.zero_()
        _lib.check(_lib.lib().lush_composite_bwd(_lib.ptr(raw), _lib.ptr(z), c_void_p(rd.data_ptr()), 3, _lib.ptr(nz), -1.0, 0,
                                                 _lib.ptr(g_rgb), None, None, None, None, _lib.ptr(g_raw), None, R, S,
                                                 _lib.stream()), "lush_composite_bwd")
        u = torch.rand(R, S, device=dev, generator=g)
        flush.zero_()
        ops.sample_merge(z, out[3], u, strict=False)
        bins = (0.5 * (z[:, 1:] + z[:, :-1])).contiguous()
        wm = out[3][:, 1:-1].contiguous()
        flush.zero_()
        ops.sample_pdf_cuda(bins, wm, ops.linspace01(S, dev))
        flush.zero_()
        ops.sample_pdf_cuda(bins, wm, ops.linspace01(S, dev), cdf=ops.reference_cdf(wm))
torch.cuda.synchronize()
print("done")
