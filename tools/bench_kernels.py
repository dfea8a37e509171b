#!/usr/bin/env python
"""Per-kernel timings on one B200 (CUDA events, warm-up, inputs > L2): compositing fwd/bwd and the sampler
against the measured HBM copy bandwidth (BASELINE config 5 stress sweep), MLP forward (fp32 SIMT vs tcgen05)
against the measured bf16 tensor peak.  Prints one JSON line per measurement."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lush_nerf_b200 import _lib, ops  # noqa: E402

DEV = "cuda"
PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}


def timeit(fn, iters=20, warm=3, graph=True):
    """Device time per call.  Short kernels are launch-bound through Python (autograd Function + ctypes is
    ~40 us of host time per call), so `iters` calls are captured in ONE CUDA graph and the replay is timed."""
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if graph:
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for _ in range(iters):
                fn()
        g.replay()
        torch.cuda.synchronize()
        a.record()
        g.replay()
        b.record()
    else:
        a.record()
        for _ in range(iters):
            fn()
        b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


def stress(R=65536):
    for S in (32, 64, 128, 256):
        g = torch.Generator(device=DEV).manual_seed(0)
        raw = torch.randn(R, S, 4, device=DEV, generator=g)
        raw[..., 3] = torch.relu(3 * torch.randn(R, S, device=DEV, generator=g)) * 5
        z = (torch.linspace(0, 1, S, device=DEV).expand(R, S) + torch.rand(R, S, device=DEV, generator=g) * (0.5 / S)).contiguous()
        rd = torch.randn(R, 3, device=DEV, generator=g)
        nz = torch.randn(R, S - 1, device=DEV, generator=g)
        raw_g = raw.clone().requires_grad_(True)
        # several independent input sets so successive launches do not hit L2 (R*S*32 B = 67..537 MB each)
        ms = timeit(lambda: ops.composite(raw, z, rd, nz))
        bytes_f = R * (32 * S + 24)
        print(json.dumps({"kernel": "composite_fwd", "R": R, "S": S, "ms": ms, "GBs": bytes_f / ms / 1e6,
                          "frac_of_measured_hbm": bytes_f / ms / 1e6 / PEAK["hbm_gbs"]}))
        out = ops.composite(raw_g, z, rd, nz)
        grgb = torch.randn(R, 3, device=DEV)

        def bwd():
            raw_g.grad = None
            out[0].backward(grgb, retain_graph=True)
        ms = timeit(bwd, graph=False)   # autograd backward inside a capture needs care; see composite_bwd(graph) below
        try:
            ms_g = timeit(lambda: torch.autograd.grad(out[0], raw_g, grgb, retain_graph=True))
            ms = min(ms, ms_g)
        except Exception:
            pass
        bytes_b = R * (40 * S + 40)
        print(json.dumps({"kernel": "composite_bwd", "R": R, "S": S, "ms": ms, "GBs": bytes_b / ms / 1e6,
                          "frac_of_measured_hbm": bytes_b / ms / 1e6 / PEAK["hbm_gbs"]}))
        w = out[3].detach()
        u = torch.rand(R, S, device=DEV, generator=g)
        ms = timeit(lambda: ops.sample_merge(z, w, u))
        bytes_s = R * (4 * S + 4 * S + 4 * S + 4 * 2 * S + 4 * S)   # z, weights, u in; z_fine, z_samples out
        print(json.dumps({"kernel": "sample_merge(rand)", "R": R, "S": S, "N": S, "ms": ms, "GBs": bytes_s / ms / 1e6,
                          "frac_of_measured_hbm": bytes_s / ms / 1e6 / PEAK["hbm_gbs"]}))
        bins = 0.5 * (z[:, 1:] + z[:, :-1]).contiguous()
        wm = w[:, 1:-1].contiguous()
        ud = torch.linspace(0, 1, S).to(DEV)
        ms = timeit(lambda: ops.sample_pdf_cuda(bins, wm, ud))
        bytes_p = R * (8 * S + 4 * S - 12)
        print(json.dumps({"kernel": "sample_pdf(det)", "R": R, "S": S, "N": S, "ms": ms, "GBs": bytes_p / ms / 1e6,
                          "frac_of_measured_hbm": bytes_p / ms / 1e6 / PEAK["hbm_gbs"]}))


def mlp(R=32768, S=128):
    from lush_nerf_b200.synthetic import scene_mlp_for_tools
    flat, views, desc = scene_mlp_for_tools(DEV)
    rb = torch.zeros(R, 11, device=DEV)
    rb[:, :3] = torch.rand(R, 3, device=DEV) * 2 - 1
    rb[:, 3:6] = torch.randn(R, 3, device=DEV) * 0.3
    rb[:, 8:11] = torch.nn.functional.normalize(torch.randn(R, 3, device=DEV), dim=-1)
    z = torch.rand(R, S, device=DEV).sort(-1)[0]
    flops = 2 * 593408 * R * S
    for impl, name in ((1, "tcgen05 fp16"), (0, "fp32 SIMT")):
        if not ops.mlp_impl_supported(desc, impl):
            continue
        with torch.no_grad():
            ms = timeit(lambda: ops.mlp_apply(rb, z, desc, impl, flat, views), iters=5 if impl == 0 else 20, warm=2, graph=False)
        print(json.dumps({"kernel": f"mlp_fwd ({name})", "R": R, "S": S, "ms": ms, "TFLOPs": flops / ms / 1e9,
                          "frac_of_measured_bf16": flops / ms / 1e9 / PEAK["bf16_tflops"]}))


if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    if which in ("all", "stress"):
        stress()
    if which in ("all", "mlp"):
        mlp()
