#!/usr/bin/env python
"""Warm (L2-resident weights), graph-replayed timings of the small latency-bound launches of the training step:
noise MLP forward / forward+backward on N_rand rows, blur-kernel motion network forward / forward+backward."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from tools.bench_kernels import timeit
from lush_nerf_b200 import ops
from lush_nerf_b200.synthetic import build_model

DEV = "cuda"
torch.manual_seed(0)
m = build_model().to(DEV).train()
N = 1024
rb = torch.zeros(N, 11, device=DEV); rb[:, :3] = torch.rand(N, 3, device=DEV); rb[:, 3:6] = torch.randn(N, 3, device=DEV) * 0.3
rb[:, 8:11] = torch.nn.functional.normalize(torch.randn(N, 3, device=DEV), dim=-1)
z = torch.rand(N, 1, device=DEV)
noise = m.mlp_noise_coarse
with torch.no_grad():
    ms = timeit(lambda: noise.render(rb, z))
print(json.dumps({"kernel": "noise MLP fwd (inference)", "rows": N, "us": ms * 1e3}))
params = [p for p in noise.parameters() if p.requires_grad]


def fb():
    out = noise.render(rb, z)
    torch.autograd.grad(out.sum(), params, allow_unused=True)


ms = timeit(fb)
print(json.dumps({"kernel": "noise MLP fwd+bwd", "rows": N, "us": ms * 1e3}))
rbk = m.mlp_rbk
idx = torch.randint(0, 20, (N, 1), device=DEV)
with torch.no_grad():
    ms = timeit(lambda: rbk.motion_params({"images_idx": idx}))
print(json.dumps({"kernel": "motion fwd", "rows": N, "us": ms * 1e3}))
mp = list(rbk.parameters())


def mfb():
    r, v, w = rbk.motion_params({"images_idx": idx})
    torch.autograd.grad((r.sum() + v.sum() + (w * w).sum()), mp, allow_unused=True)


ms = timeit(mfb)
print(json.dumps({"kernel": "motion fwd+bwd (incl. torch sum glue)", "rows": N, "us": ms * 1e3}))
