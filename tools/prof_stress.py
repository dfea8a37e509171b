"""One launch each of the compositing / sampler kernels at the stress size (65 536 rays x 128 samples) for ncu."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from lush_nerf_b200 import ops
DEV = "cuda"
R, S = 65536, int(sys.argv[1]) if len(sys.argv) > 1 else 128
g = torch.Generator(device=DEV).manual_seed(0)
raw = torch.randn(R, S, 4, device=DEV, generator=g)
raw[..., 3] = torch.relu(3 * torch.randn(R, S, device=DEV, generator=g)) * 5
z = (torch.linspace(0, 1, S, device=DEV).expand(R, S) + torch.rand(R, S, device=DEV, generator=g) * (0.5 / S)).contiguous()
rd = torch.randn(R, 3, device=DEV, generator=g)
nz = torch.randn(R, S - 1, device=DEV, generator=g)
raw_g = raw.clone().requires_grad_(True)
for _ in range(2):
    out = ops.composite(raw_g, z, rd, nz)
    out[0].backward(torch.randn(R, 3, device=DEV), retain_graph=False)
    w = out[3].detach()
    u = torch.rand(R, S, device=DEV, generator=g)
    ops.sample_merge(z, w, u)
    bins = 0.5 * (z[:, 1:] + z[:, :-1]).contiguous()
    ops.sample_pdf_cuda(bins, w[:, 1:-1].contiguous(), torch.linspace(0, 1, S).to(DEV))
torch.cuda.synchronize()
print("done")
