import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from lush_nerf_b200 import _lib, ops
from oracle import lush_oracle as O
from tests.test_kernels_gpu import mlp_case
DEV = "cuda"
nets = O.det_nets(O.POSTER_CFG, 1, 30.0)
desc = _lib.MlpDesc(8, 256, 10, 4, 4, 0)
def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()
for (R, S, need_in, wide) in [(3, 50, True, True), (2, 64, False, False), (64, 24, True, False), (300, 64, True, True)]:
    rb, z = mlp_case("coarse", R, S, 11 + R)
    g = torch.Generator().manual_seed(R)
    cot = torch.randn(R, S, 4, generator=g)
    if wide: cot = cot * torch.rand(R, S, 1, generator=g) ** 4
    res = {}
    for impl in (0, 1):
        flat = torch.cat([nets["coarse"][n].reshape(-1) for n, _ in O.mlp_param_shapes(8, 256, 63, 27, 4)]).to(DEV)
        views, off = [], 0
        for _, shp in O.mlp_param_shapes(8, 256, 63, 27, 4):
            n = int(np.prod(shp)); views.append(flat[off:off + n].view(shp).requires_grad_(True)); off += n
        rbc = rb.to(DEV).requires_grad_(need_in)
        out = ops.mlp_apply(rbc, z.to(DEV), desc, impl, flat, views)
        (out * cot.to(DEV)).sum().backward()
        torch.cuda.synchronize()
        res[impl] = (out.detach(), [v.grad for v in views], rbc.grad)
    print(f"--- R={R} S={S} need_in={need_in} wide={wide}: out rel {rel(res[1][0], res[0][0]):.2e}")
    for (name, _), g1, g0 in zip(O.mlp_param_shapes(8, 256, 63, 27, 4), res[1][1], res[0][1]):
        print(f"   {name:28s} rel {rel(g1, g0):.3e}   |g0| {g0.norm().item():.3e}")
    if need_in:
        print(f"   d(o,d) rel {rel(res[1][2][:, :6], res[0][2][:, :6]):.3e}   d(viewdirs) rel {rel(res[1][2][:, 8:], res[0][2][:, 8:]):.3e}")
