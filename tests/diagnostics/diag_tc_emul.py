"""tc backward vs a torch emulation with the SAME fp16 operand roundings (weights, activations, dZ under the
same loss scale) and fp32 accumulation.  Agreement here at ~1e-3 isolates kernel correctness from the
inherent ReLU-mask flips that separate any reduced-precision forward from the fp32 one."""
import sys, os, math
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
import torch.nn.functional as F
from lush_nerf_b200 import _lib, ops
from oracle import lush_oracle as O
from tests.test_kernels_gpu import mlp_case
DEV = "cuda"
torch.backends.cuda.matmul.allow_tf32 = False
h16 = lambda t: t.half().float()

def emulate(p, rb, z, cot):
    R, S = z.shape
    o, d, v = rb[:, 0:3], rb[:, 3:6], rb[:, 8:11]
    x = (o[:, None, :] + d[:, None, :] * z[:, :, None]).reshape(-1, 3)
    e = h16(O.embed(x, 10)); ev = h16(O.embed(v[:, None, :].expand(R, S, 3).reshape(-1, 3), 4))
    W = {k: h16(t) for k, t in p.items()}
    hs, masks = [], []
    h = e
    for i in range(8):
        zz = h @ W[f"pts_linears.{i}.weight"].t() + p[f"pts_linears.{i}.bias"]
        a = torch.relu(zz); masks.append(zz > 0)
        a32 = a
        a = h16(a); hs.append(a)
        h = torch.cat([e, a], -1) if i == 4 else a
    alpha = a32 @ p["alpha_linear.weight"].t() + p["alpha_linear.bias"]
    feat = h16(hs[7] @ W["feature_linear.weight"].t() + p["feature_linear.bias"])
    zv = torch.cat([feat, ev], -1) @ W["views_linears.0.weight"].t() + p["views_linears.0.bias"]
    hv32 = torch.relu(zv); mv = zv > 0
    rgb = hv32 @ p["rgb_linear.weight"].t() + p["rgb_linear.bias"]
    out = torch.cat([rgb, alpha], -1)
    g = cot.reshape(-1, 4)
    scale = 2.0 ** math.floor(math.log2(32.0 / g.abs().max().item()))
    gs = g * scale
    G = {}
    G["rgb_linear.weight"] = g[:, :3].t() @ h16(hv32); G["rgb_linear.bias"] = g[:, :3].sum(0)
    G["alpha_linear.weight"] = (g[:, 3:4].t() @ hs[7]); G["alpha_linear.bias"] = g[:, 3].sum().reshape(1)
    dzv = h16((gs[:, :3] @ p["rgb_linear.weight"]) * mv)
    G["views_linears.0.weight"] = dzv.t() @ torch.cat([feat, ev], -1) / scale; G["views_linears.0.bias"] = dzv.sum(0) / scale
    gf = h16(dzv @ W["views_linears.0.weight"][:, :256])
    G["feature_linear.weight"] = gf.t() @ hs[7] / scale; G["feature_linear.bias"] = gf.sum(0) / scale
    dz = h16((gf @ W["feature_linear.weight"] + gs[:, 3:4] * p["alpha_linear.weight"]) * masks[7])
    for l in range(7, -1, -1):
        xin = e if l == 0 else (torch.cat([e, hs[l - 1]], -1) if l == 5 else hs[l - 1])
        G[f"pts_linears.{l}.weight"] = dz.t() @ xin / scale; G[f"pts_linears.{l}.bias"] = dz.sum(0) / scale
        if l > 0:
            Wl = W[f"pts_linears.{l}.weight"]
            Wm = Wl[:, 63:] if l == 5 else Wl
            dz = h16((dz @ Wm) * masks[l - 1])
    # pad: layer 0/5 enc columns are 63 wide in torch layout already
    return out, G

def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()

nets = O.det_nets(O.POSTER_CFG, 1, 30.0)
desc = _lib.MlpDesc(8, 256, 10, 4, 4, 0)
for (R, S, wide) in [(3, 50, True), (64, 24, False), (300, 64, True), (1024, 128, True)]:
    rb, z = mlp_case("coarse", R, S, 11 + R)
    g = torch.Generator().manual_seed(R)
    cot = torch.randn(R, S, 4, generator=g)
    if wide: cot = cot * torch.rand(R, S, 1, generator=g) ** 4
    p = {k: t.to(DEV) for k, t in nets["coarse"].items()}
    out_e, G = emulate(p, rb.to(DEV), z.to(DEV), cot.to(DEV))
    flat = torch.cat([nets["coarse"][n].reshape(-1) for n, _ in O.mlp_param_shapes(8, 256, 63, 27, 4)]).to(DEV)
    views, off = [], 0
    for _, shp in O.mlp_param_shapes(8, 256, 63, 27, 4):
        n = int(np.prod(shp)); views.append(flat[off:off + n].view(shp).requires_grad_(True)); off += n
    out = ops.mlp_apply(rb.to(DEV), z.to(DEV), desc, 1, flat, views)
    (out * cot.to(DEV)).sum().backward()
    torch.cuda.synchronize()
    print(f"--- R={R} S={S} wide={wide}: out vs emulation rel {rel(out.reshape(-1, 4), out_e):.2e}")
    for (name, _), v in zip(O.mlp_param_shapes(8, 256, 63, 27, 4), views):
        print(f"   {name:28s} rel vs emulation {rel(v.grad, G[name].reshape(v.shape)):.3e}")
