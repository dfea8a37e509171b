"""torch.autograd.Function wrappers over the C ABI (include/lushnerf_b200.h).

PyTorch is plumbing here: it owns device memory (caching allocator), the current stream and the autograd
tape.  Every numerical step of the hot path is one of the hand-written sm_100a kernels in csrc/.
"""
from __future__ import annotations

import os
from ctypes import byref, c_void_p
from typing import Optional

import torch

from . import _lib
from ._lib import MlpDesc, check, lib, ptr, stream

# MLP implementation: 1 = tcgen05 fp16-operand / fp32-accumulate (default when the topology is covered),
# 0 = exact fp32 SIMT (strict-parity mode).  LUSH_MLP_IMPL=simt|tc overrides the default.
_IMPL_ENV = {"simt": 0, "fp32": 0, "0": 0, "tc": 1, "tcgen05": 1, "1": 1}


def default_mlp_impl() -> int:
    return _IMPL_ENV.get(os.environ.get("LUSH_MLP_IMPL", "tc").lower(), 1)


def _f32c(t: Optional[torch.Tensor]) -> Optional[torch.Tensor]:
    if t is None:
        return None
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


def _empty(shape, like):
    return torch.empty(shape, dtype=torch.float32, device=like.device)


# ------------------------------------------------------------------------------------------------
# ray preparation (reference models/lushnerf.py:706-729, helpers:542-562)
# ------------------------------------------------------------------------------------------------
class _RayPrep(torch.autograd.Function):
    @staticmethod
    def forward(ctx, rays, H, W, focal, near, far, ndc):
        rays = _f32c(rays)
        R = rays.shape[0]
        out = _empty((R, 11), rays)
        check(lib().lush_ray_prep_fwd(ptr(rays, "rays"), ptr(out), R, int(H), int(W), float(focal),
                                      float(near), float(far), int(bool(ndc)), stream()), "lush_ray_prep_fwd")
        ctx.save_for_backward(rays)
        ctx.meta = (int(H), int(W), float(focal), int(bool(ndc)))
        return out

    @staticmethod
    def backward(ctx, g):
        (rays,) = ctx.saved_tensors
        H, W, focal, ndc = ctx.meta
        g = _f32c(g)
        g_rays = torch.empty_like(rays)
        check(lib().lush_ray_prep_bwd(ptr(rays), ptr(g), ptr(g_rays), rays.shape[0], H, W, focal, ndc,
                                      stream()), "lush_ray_prep_bwd")
        return g_rays, None, None, None, None, None, None


def ray_prep(rays: torch.Tensor, H: int, W: int, focal: float, near=0.0, far=1.0, ndc=True) -> torch.Tensor:
    """rays [..., 3, 2] -> ray_batch [R, 11] = (o, d, near, far, viewdirs)."""
    return _RayPrep.apply(rays.reshape(-1, 3, 2), H, W, focal, near, far, ndc)


def ray_gen_prep(c2w: torch.Tensor, H: int, W: int, K, near=0.0, far=1.0, ndc=True, pix0: int = 0, n: int = None):
    """ray_batch [n, 11] of pixels [pix0, pix0 + n) of the pose c2w ([3,4] or [4,4]): get_rays + ray_prep in one
    launch, nothing of size [H, W, 3, 2] is materialised (no gradient: poses are constants of the path)."""
    if not c2w.is_cuda:
        raise RuntimeError("lush_nerf_b200.ray_gen_prep needs a CUDA pose; there is no CPU fallback")
    c2w = _f32c(c2w.detach())
    if c2w.dim() != 2 or c2w.shape[0] < 3 or c2w.shape[1] != 4:
        raise ValueError(f"ray_gen_prep: c2w must be [3,4] or [4,4], got {tuple(c2w.shape)}")
    n = H * W - pix0 if n is None else n
    out = _empty((n, 11), c2w)
    check(lib().lush_ray_gen_prep(ptr(c2w, "c2w"), 4, int(H), int(W), float(K[0][0]), float(K[1][1]), float(K[0][2]),
                                  float(K[1][2]), float(near), float(far), int(bool(ndc)), int(pix0), int(n), ptr(out),
                                  stream()), "lush_ray_gen_prep")
    return out


def ray_gen_gather(poses: torch.Tensor, view_idx: torch.Tensor, pix_idx: torch.Tensor, H: int, W: int, K,
                   images: Optional[torch.Tensor] = None, masks: Optional[torch.Tensor] = None,
                   err_flags: Optional[torch.Tensor] = None, validate: bool = False):
    """A training batch staged on the device.  rays [n, 3, 2] of the (view, pixel) pairs generated from the poses
    ([V,3,4], [V,4,4] or the LLFF [V,3,5] block) instead of looked up in a pre-expanded ray table; with `images`
    ([V,H,W,3] fp32) and / or `masks` ([V,H,W] bool / uint8) the pixels' target colours and mask bits come back too:
    returns rays, or (rays, rgb, mask) when a table is given (None for the missing one).  No gradient (constants).
    An index outside its range is clamped and flagged in `err_flags` (int32[1], e.g. the model's guard word);
    validate=True reads the flag back (one host sync) and raises IndexError like the reference's fancy indexing would."""
    if not poses.is_cuda:
        raise RuntimeError("lush_nerf_b200.ray_gen_gather needs CUDA tensors; there is no CPU fallback")
    poses = _f32c(poses.detach())
    if poses.dim() != 3 or poses.shape[1] < 3 or poses.shape[2] < 4:
        raise ValueError(f"ray_gen_gather: poses must be [V,3,4], [V,4,4] or [V,3,5], got {tuple(poses.shape)}")
    view_idx, pix_idx = view_idx.reshape(-1).contiguous(), pix_idx.reshape(-1).contiguous()
    if view_idx.numel() != pix_idx.numel():
        raise ValueError("ray_gen_gather: view_idx and pix_idx differ in length")
    n, V = view_idx.numel(), poses.shape[0]
    out = _empty((n, 3, 2), poses)
    rgb = mask = None
    if images is not None:
        images = _f32c(images)
        if images.numel() != V * H * W * 3:
            raise ValueError(f"ray_gen_gather: images must hold [V={V}, {H}, {W}, 3] floats, got {tuple(images.shape)}")
        rgb = _empty((n, 3), poses)
    if masks is not None:
        masks = masks.contiguous()
        masks = masks.view(torch.uint8) if masks.dtype == torch.bool else masks
        if masks.dtype != torch.uint8 or masks.numel() != V * H * W:
            raise ValueError(f"ray_gen_gather: masks must be [V={V}, {H}, {W}] bool / uint8")
        mask = torch.empty(n, dtype=torch.uint8, device=poses.device)
    if validate and err_flags is None:
        err_flags = torch.zeros(1, dtype=torch.int32, device=poses.device)
    check(lib().lush_ray_gen_gather(
        ptr(poses, "poses"), V, poses.shape[1] * poses.shape[2], poses.shape[2], ptr(view_idx, "view_idx", torch.int64),
        ptr(pix_idx, "pix_idx", torch.int64), int(H), int(W), float(K[0][0]), float(K[1][1]), float(K[0][2]),
        float(K[1][2]), n, ptr(out), ptr(images, "images"), ptr(rgb), ptr(masks, "masks", torch.uint8),
        ptr(mask, "mask_out", torch.uint8), ptr(err_flags, "err_flags", torch.int32), stream()), "lush_ray_gen_gather")
    if validate:
        bits = int(err_flags.item())
        if bits & (_lib.FLAG_BAD_VIEW_INDEX | _lib.FLAG_BAD_PIXEL_INDEX):
            err_flags.bitwise_and_(~(_lib.FLAG_BAD_VIEW_INDEX | _lib.FLAG_BAD_PIXEL_INDEX))
            what = " and ".join(w for w, b in (("view index", _lib.FLAG_BAD_VIEW_INDEX), ("pixel index", _lib.FLAG_BAD_PIXEL_INDEX))
                                if bits & b)
            raise IndexError(f"ray_gen_gather: {what} out of range (V={V}, H*W={H * W})")
    if images is None and masks is None:
        return out
    return out, rgb, (mask.view(torch.bool) if mask is not None else None)


# ------------------------------------------------------------------------------------------------
# coarse sample placement (reference models/lushnerf.py:389-412)
# ------------------------------------------------------------------------------------------------
_LINSPACE_CACHE = {}


def linspace01(n: int, device) -> torch.Tensor:
    """torch.linspace(0, 1, n) evaluated on the CPU (the oracle's device) and cached on `device`."""
    key = (int(n), str(device))
    t = _LINSPACE_CACHE.get(key)
    if t is None:
        t = torch.linspace(0.0, 1.0, steps=int(n), device="cpu").to(device)   # explicit: the driver sets a CUDA default type
        _LINSPACE_CACHE[key] = t
    return t


@torch.no_grad()
def stratified_z(ray_batch: torch.Tensor, t_vals: torch.Tensor, t_rand: Optional[torch.Tensor] = None,
                 lindisp: bool = False) -> torch.Tensor:
    """z [R, S]; near/far are read from columns 6, 7 of ray_batch.  No gradient (near/far are constants)."""
    rb = _f32c(ray_batch.detach())
    R, S = rb.shape[0], t_vals.shape[0]
    z = _empty((R, S), rb)
    check(lib().lush_stratified_z(ptr(rb), rb.shape[1], ptr(_f32c(t_vals)), ptr(_f32c(t_rand)), ptr(z), R, S,
                                  int(bool(lindisp)), stream()), "lush_stratified_z")
    return z


# ------------------------------------------------------------------------------------------------
# alpha compositing (reference models/lushnerf.py:296-352)
# ------------------------------------------------------------------------------------------------
class _Composite(torch.autograd.Function):
    @staticmethod
    def forward(ctx, raw, z, rays_d, sigma_noise, thresh, white_bkgd, flags):
        raw, z, sigma_noise = _f32c(raw), _f32c(z), _f32c(sigma_noise)
        if rays_d.dtype != torch.float32 or rays_d.stride(-1) != 1:
            rays_d = rays_d.float().contiguous()
        R, S = z.shape
        d_stride = rays_d.stride(0) if R > 1 else max(3, rays_d.stride(0))
        rgb, dens = _empty((R, 3), raw), _empty((R, max(S - 1, 0)), raw)
        acc, w, depth = _empty((R,), raw), _empty((R, S), raw), _empty((R,), raw)
        fptr = c_void_p(flags.data_ptr()) if flags is not None else None
        check(lib().lush_composite_fwd(ptr(raw, "raw"), ptr(z, "z"), c_void_p(rays_d.data_ptr()), d_stride,
                                       ptr(sigma_noise), float(thresh), int(bool(white_bkgd)), ptr(rgb),
                                       ptr(dens), ptr(acc), ptr(w), ptr(depth), fptr, R, S, stream()),
              "lush_composite_fwd")
        ctx.save_for_backward(raw, z, rays_d, sigma_noise)
        ctx.meta = (float(thresh), int(bool(white_bkgd)), d_stride)
        ctx.set_materialize_grads(False)     # unused outputs arrive as None (NULL for the kernel), not as zero fills
        return rgb, dens, acc, w, depth

    @staticmethod
    def backward(ctx, g_rgb, g_dens, g_acc, g_w, g_depth):
        raw, z, rays_d, sigma_noise = ctx.saved_tensors
        thresh, white, d_stride = ctx.meta
        R, S = z.shape
        g_raw = torch.empty_like(raw)
        need_d = ctx.needs_input_grad[2]
        g_d = _empty((R, 3), raw) if need_d else None
        check(lib().lush_composite_bwd(ptr(raw), ptr(z), c_void_p(rays_d.data_ptr()), d_stride,
                                       ptr(sigma_noise), thresh, white, ptr(_f32c(g_rgb)), ptr(_f32c(g_dens)),
                                       ptr(_f32c(g_acc)), ptr(_f32c(g_w)), ptr(_f32c(g_depth)), ptr(g_raw),
                                       ptr(g_d), R, S, stream()), "lush_composite_bwd")
        return g_raw, None, g_d, None, None, None, None


def composite(raw, z, rays_d, sigma_noise=None, rm_near_thresh: float = -1.0, white_bkgd=False, flags=None):
    """raw2outputs: returns (rgb_map, density, acc, weights, depth).  rays_d may be a strided [R,3] view
    (e.g. ray_batch[:, 3:6]).  z carries no gradient (the reference's z is built from constants)."""
    if not raw.is_cuda:
        raise RuntimeError("lush_nerf_b200.composite needs CUDA tensors; there is no CPU fallback")
    return _Composite.apply(raw, z.detach(), rays_d, None if sigma_noise is None else sigma_noise.detach(),
                            rm_near_thresh, white_bkgd, flags)


# ------------------------------------------------------------------------------------------------
# hierarchical sampling (reference helpers:566-609; models/lushnerf.py:435-440,465)
# ------------------------------------------------------------------------------------------------
def sample_strict(det: bool) -> bool:
    """Which cdf feeds the search kernel.  LUSH_SAMPLE_STRICT = det (default) | 1 | 0:
      strict -> the cdf is built by the reference's own op sequence (helpers:568-571: + 1e-5, / sum, torch.cumsum, cat)
                with torch ops ON THIS DEVICE, the kernel only searches / gathers / interpolates: indices and samples
                equal what the reference's searchsorted produces on the same GPU, bit for bit;
      fused  -> the cdf is built inside the kernel in a fixed warp-scan order (no extra launches); cdf entries move by
                <= 4 ulp against torch's order, which can flip an index only where u touches a cdf entry.
    Default: strict for deterministic sampling (rendering, perturb = 0: where 'bit-exact indices' is defined), fused for
    random u (training: the draw itself differs from run to run)."""
    mode = os.environ.get("LUSH_SAMPLE_STRICT", "det").lower()
    if mode in ("1", "true", "always", "on"):
        return True
    if mode in ("0", "false", "never", "off"):
        return False
    return bool(det)


@torch.no_grad()
def reference_cdf(weights: torch.Tensor) -> torch.Tensor:
    """cdf [R, nb] of weights [R, nb-1], op for op as the reference builds it (helpers:568-571)."""
    w = weights + 1e-5
    pdf = w / torch.sum(w, -1, keepdim=True)
    cdf = torch.cumsum(pdf, -1)
    return torch.cat([torch.zeros_like(cdf[..., :1]), cdf], -1).contiguous()


@torch.no_grad()
def sample_pdf_cuda(bins, weights, u, cdf: Optional[torch.Tensor] = None, return_inds=False):
    bins, weights, u = _f32c(bins), _f32c(weights), _f32c(u)
    R, nb = bins.shape
    N = u.shape[-1]
    u_stride = 0 if u.dim() == 1 else N
    samples = _empty((R, N), bins)
    inds = torch.empty((R, N), dtype=torch.int64, device=bins.device) if return_inds else None
    check(lib().lush_sample_pdf(ptr(bins, "bins"), ptr(weights, "weights"), ptr(_f32c(cdf)), ptr(u, "u"),
                                u_stride, ptr(samples), ptr(inds, dtype=torch.int64), R, nb, N, stream()),
          "lush_sample_pdf")
    return (samples, inds) if return_inds else samples


@torch.no_grad()
def sample_merge(z, weights, u, want_std=True, strict: Optional[bool] = None):
    """Fused z_mid + sample_pdf(weights[:,1:-1]) + sort(cat(z, samples)) + z_std.  z must be ascending per row (the
    merge is by rank).  strict: see sample_strict(); None -> strict iff u is the shared deterministic [N] vector."""
    z, weights, u = _f32c(z), _f32c(weights), _f32c(u)
    R, S = z.shape
    N = u.shape[-1]
    u_stride = 0 if u.dim() == 1 else N
    strict = sample_strict(u.dim() == 1) if strict is None else strict
    cdf = reference_cdf(weights[:, 1:-1]) if strict else None
    z_fine, z_samples = _empty((R, S + N), z), _empty((R, N), z)
    z_std = _empty((R,), z) if want_std else None
    check(lib().lush_sample_merge(ptr(z, "z"), ptr(weights, "weights"), ptr(cdf, "cdf"), ptr(u, "u"), u_stride,
                                  ptr(z_fine), ptr(z_samples), ptr(z_std), R, S, N, stream()), "lush_sample_merge")
    return z_fine, z_samples, z_std


# ------------------------------------------------------------------------------------------------
# encoder + MLP (reference helpers:311-344, 394-423, 483-512; models/lushnerf.py:234-293)
# ------------------------------------------------------------------------------------------------
# bench.py sets this to a list to collect (start, end) CUDA-event pairs around every MLP launch sequence
MLP_EVENT_SINK = None
# train.ArenaGradAllReduce sets this to `fn(flat_grad_tensor)`: called from inside backward as soon as the kernels that
# fill a network's flat gradient arena are enqueued, so the arena's all-reduce can start while the rest of the
# backward pass (the other network) is still running
GRAD_ARENA_HOOK = None


def _arena_done(g_flat):
    if GRAD_ARENA_HOOK is not None:
        GRAD_ARENA_HOOK(g_flat)


class _Timed:
    """CUDA events around one C-ABI MLP call; appended to MLP_EVENT_SINK as (label, start, end) when it is a list."""

    def __init__(self, label="mlp"):
        self.label = label

    def __enter__(self):
        self.on = MLP_EVENT_SINK is not None
        if self.on:
            self.a = torch.cuda.Event(enable_timing=True)
            self.b = torch.cuda.Event(enable_timing=True)
            self.a.record()

    def __exit__(self, *exc):
        if self.on:
            self.b.record()
            MLP_EVENT_SINK.append((self.label, self.a, self.b))


def mlp_impl_supported(desc: MlpDesc, impl: int, training: bool = False) -> bool:
    caps = int(lib().lush_mlp_impl_supported(byref(desc), int(impl)))
    return bool(caps & (2 if training else 1))


class _MLP(torch.autograd.Function):
    """out[R,S,C] = net(encode(o + d*z), encode(viewdirs)).  `flat` is the flat fp32 parameter buffer
    the individual nn.Parameters are views of; the parameters themselves are passed only so that
    autograd routes their gradients (returned as views of one flat gradient buffer)."""

    @staticmethod
    def forward(ctx, ray_batch, z, desc, impl, flat, numels, track, *params):
        rb, z = _f32c(ray_batch), _f32c(z)
        R, S = z.shape
        C = 4 if desc.out_mode == 0 else 3
        out = _empty((R, S, C), rb)
        L = lib()
        # `track` = grad mode of the CALLER (inside forward() it is always off, and ctx.needs_input_grad only says which
        # inputs require grad, also under torch.no_grad()): without it a no_grad render would run the training kernel
        # and write a 4.7 KB/sample activation stash nobody reads
        training = bool(track) and any(ctx.needs_input_grad)
        stash = None
        if training:
            nb = L.lush_mlp_stash_bytes(byref(desc), R, S, impl)
            stash = torch.empty((nb + 3) // 4, dtype=torch.float32, device=rb.device)
        wb = L.lush_mlp_workspace_bytes(byref(desc), R, S, impl, 0)
        ws = torch.empty((wb + 3) // 4, dtype=torch.float32, device=rb.device) if wb else None
        with _Timed("fwd" if impl == 1 else "fwd_simt"):
            check(L.lush_mlp_fwd(byref(desc), ptr(flat, "params"), ptr(rb, "ray_batch"), rb.shape[1], ptr(z, "z"),
                                 ptr(out), ptr(stash), ptr(ws), wb, R, S, impl, stream()), "lush_mlp_fwd")
        if training:
            ctx.save_for_backward(rb, z, flat, stash)
            ctx.meta = (desc, impl, numels, [tuple(p.shape) for p in params])
        return out

    @staticmethod
    def backward(ctx, g_out):
        rb, z, flat, stash = ctx.saved_tensors
        desc, impl, numels, shapes = ctx.meta
        R, S = z.shape
        L = lib()
        g_out = _f32c(g_out)
        g_flat = torch.zeros_like(flat)
        g_rb = _empty((R, 11), rb) if ctx.needs_input_grad[0] else None
        wb = L.lush_mlp_workspace_bytes(byref(desc), R, S, impl, 1)
        ws = torch.empty((wb + 3) // 4, dtype=torch.float32, device=rb.device) if wb else None
        with _Timed("bwd" if impl == 1 else "bwd_simt"):
            check(L.lush_mlp_bwd(byref(desc), ptr(flat), ptr(rb), rb.shape[1], ptr(z), ptr(g_out), ptr(stash),
                                 ptr(g_flat), ptr(g_rb), ptr(ws), wb, R, S, impl, stream()), "lush_mlp_bwd")
        _arena_done(g_flat)
        grads, off = [], 0
        for n, shp, need in zip(numels, shapes, ctx.needs_input_grad[7:]):
            grads.append(g_flat[off:off + n].view(shp) if need else None)
            off += n
        if g_rb is not None and rb.shape[1] != 11:
            full = torch.zeros_like(rb)
            full[:, :11] = g_rb
            g_rb = full
        return (g_rb, None, None, None, None, None, None, *grads)


def mlp_apply(ray_batch, z, desc: MlpDesc, impl: int, flat: torch.Tensor, params):
    if not ray_batch.is_cuda:
        raise RuntimeError("lush_nerf_b200 MLP needs CUDA tensors; there is no CPU fallback")
    numels = [p.numel() for p in params]
    return _MLP.apply(ray_batch, z.detach(), desc, int(impl), flat, numels, torch.is_grad_enabled(), *params)


# ------------------------------------------------------------------------------------------------
# rigid-warp blur kernel (reference models/lushnerf.py:75-98; utils/rigid_warping.py:20-140)
# ------------------------------------------------------------------------------------------------
class _RbkWarp(torch.autograd.Function):
    @staticmethod
    def forward(ctx, rays, r, v):
        rays, r, v = _f32c(rays), _f32c(r), _f32c(v)
        N, M = rays.shape[0], r.shape[-1]
        out = _empty((N, M + 1, 3, 2), rays)
        check(lib().lush_rbk_warp_fwd(ptr(rays, "rays"), ptr(r, "r"), ptr(v, "v"), ptr(out), N, M, stream()),
              "lush_rbk_warp_fwd")
        ctx.save_for_backward(rays, r, v)
        return out

    @staticmethod
    def backward(ctx, g):
        rays, r, v = ctx.saved_tensors
        N, M = rays.shape[0], r.shape[-1]
        g = _f32c(g)
        g_rays, g_r, g_v = torch.empty_like(rays), torch.empty_like(r), torch.empty_like(v)
        check(lib().lush_rbk_warp_bwd(ptr(rays), ptr(r), ptr(v), ptr(g), ptr(g_rays), ptr(g_r), ptr(g_v), N, M,
                                      stream()), "lush_rbk_warp_bwd")
        return g_rays, g_r, g_v


def rbk_warp(rays, r, v):
    """rays [N,3,2]; r, v [N,3,M] -> [N, M+1, 3, 2] (slot 0 = the original ray)."""
    if not rays.is_cuda:
        raise RuntimeError("lush_nerf_b200.rbk_warp needs CUDA tensors; there is no CPU fallback")
    return _RbkWarp.apply(rays, r, v)


# ------------------------------------------------------------------------------------------------
# motion-parameter network of the blur kernel (reference models/lushnerf.py:118-147), one launch each way
# ------------------------------------------------------------------------------------------------
class MotionPlan:
    """Host description of the network for lush_rbk_motion_{fwd,bwd}: `layers` is a list of
    (weight, bias, [(col_off, dim), ...] inputs, relu) in execution order; outputs are appended to the
    activation row after the embedding columns.  heads = indices of the r, v, w head layers."""

    def __init__(self, embed_dim, num_motion, rv_window):
        self.embed_dim, self.num_motion, self.rv_window = int(embed_dim), int(num_motion), float(rv_window)
        self.width = self.embed_dim
        self.layers, self.params = [], []
        self.heads = {}

    def add(self, linear, inputs, relu):
        """inputs: list of (col_off, dim) segments (at most two, concat order); returns the output segment."""
        if len(inputs) not in (1, 2) or len(self.layers) >= _lib.MOTION_MAX_LAYERS:
            raise ValueError("motion network: unsupported topology")
        out_dim, in_dim = linear.weight.shape
        if in_dim != sum(d for _, d in inputs) or linear.bias is None:
            raise ValueError("motion network: layer shape does not match its inputs")
        seg = (self.width, out_dim)
        self.layers.append((inputs, seg, bool(relu)))
        self.params += [linear.weight, linear.bias]
        self.width += out_dim
        return seg

    def net(self, params, grads=None):
        net = _lib.MotionNet()
        net.n_layers, net.row_width, net.embed_dim = len(self.layers), self.width, self.embed_dim
        net.num_motion, net.rv_window = self.num_motion, self.rv_window
        net.r_off, net.v_off, net.w_off = (self.layers[self.heads[k]][1][0] for k in ("r", "v", "w"))
        for i, (inputs, seg, relu) in enumerate(self.layers):
            L = net.layers[i]
            L.w, L.b = ptr(params[2 * i], "weight").value, ptr(params[2 * i + 1], "bias").value
            if grads is not None:
                L.gw, L.gb = grads[2 * i].data_ptr(), grads[2 * i + 1].data_ptr()
            L.in0_off, L.in0_dim = inputs[0]
            L.in1_off, L.in1_dim = inputs[1] if len(inputs) == 2 else (0, 0)
            L.out_off, L.out_dim = seg
            L.act = int(relu)
        return net


class _RbkMotion(torch.autograd.Function):
    @staticmethod
    def forward(ctx, idx, plan, err_flags, embed, *params):
        idx = idx.reshape(-1).contiguous()
        embed = _f32c(embed)
        params = [_f32c(p) for p in params]
        N, n_img, M = idx.numel(), embed.shape[0], plan.num_motion
        rows = _empty((n_img, plan.width), embed)
        r, v, ccw = _empty((N, 3 * M), embed), _empty((N, 3 * M), embed), _empty((N, M + 1), embed)
        net = plan.net(params)
        check(lib().lush_rbk_motion_fwd(byref(net), ptr(embed, "embedding"), n_img, ptr(idx, "images_idx", torch.int64),
                                        N, ptr(rows), ptr(r), ptr(v), ptr(ccw),
                                        ptr(err_flags, "err_flags", torch.int32), stream()), "lush_rbk_motion_fwd")
        ctx.save_for_backward(idx, rows, embed, *params)
        ctx.plan = plan
        ctx.set_materialize_grads(False)
        return r, v, ccw

    @staticmethod
    def backward(ctx, g_r, g_v, g_ccw):
        idx, rows, embed, *params = ctx.saved_tensors
        plan = ctx.plan
        numels = [p.numel() for p in params]
        flat = torch.zeros(sum(numels) + embed.numel(), dtype=torch.float32, device=embed.device)  # one fill
        grads, off = [], 0
        for p, n in zip(params, numels):
            grads.append(flat[off:off + n].view(p.shape))
            off += n
        g_embed = flat[off:].view(embed.shape)
        net = plan.net(params, grads)
        check(lib().lush_rbk_motion_bwd(byref(net), embed.shape[0], ptr(idx, "images_idx", torch.int64), idx.numel(),
                                        ptr(rows), ptr(_f32c(g_r)), ptr(_f32c(g_v)), ptr(_f32c(g_ccw)),
                                        ptr(g_embed), stream()), "lush_rbk_motion_bwd")
        _arena_done(flat)
        need = ctx.needs_input_grad
        return (None, None, None, g_embed if need[3] else None,
                *[g if nd else None for g, nd in zip(grads, need[4:])])


def rbk_motion(images_idx, plan: MotionPlan, embed_weight, params, err_flags: Optional[torch.Tensor] = None):
    """(r [N,3M], v [N,3M], ccw [N,M+1]) for the rays' image indices; one kernel forward, one backward.
    An index outside [0, num_img) -- nn.Embedding raises on it in the reference -- gives that ray the identity
    kernel and sets LUSH_FLAG_BAD_IMAGE_INDEX in `err_flags` (int32[1]); without `err_flags` the indices are
    validated here with one host sync (IndexError)."""
    if not embed_weight.is_cuda:
        raise RuntimeError("lush_nerf_b200.rbk_motion needs CUDA tensors; there is no CPU fallback")
    if err_flags is None and images_idx.numel() and not torch.cuda.is_current_stream_capturing():
        lo, hi = int(images_idx.min()), int(images_idx.max())
        if lo < 0 or hi >= embed_weight.shape[0]:
            raise IndexError(f"rbk_motion: images_idx in [{lo}, {hi}] outside the embedding table [0, {embed_weight.shape[0]})")
    return _RbkMotion.apply(images_idx, plan, err_flags, embed_weight, *params)


# ------------------------------------------------------------------------------------------------
# K-ray blend + noise + tone map (reference models/lushnerf.py:100-116, 649-654; helpers:164-174)
# ------------------------------------------------------------------------------------------------
class _BlendTonemap(torch.autograd.Function):
    @staticmethod
    def forward(ctx, rgb, ccw, noise_raw, K, gamma):
        rgb, ccw, noise_raw = _f32c(rgb), _f32c(ccw), _f32c(noise_raw)
        N = rgb.shape[0] // K
        blur, pure = _empty((N, 3), rgb), _empty((N, 3), rgb)
        nz = _empty((N, 3), rgb) if noise_raw is not None else None
        check(lib().lush_blend_tonemap_fwd(ptr(rgb, "rgb"), ptr(ccw), ptr(noise_raw), ptr(blur), ptr(pure),
                                           ptr(nz), N, K, int(bool(gamma)), stream()), "lush_blend_tonemap_fwd")
        ctx.save_for_backward(rgb, ccw, noise_raw)
        ctx.meta = (N, K, int(bool(gamma)))
        ctx.set_materialize_grads(False)
        if nz is None:
            nz = torch.zeros_like(blur)
            ctx.mark_non_differentiable(nz)
        return blur, pure, nz

    @staticmethod
    def backward(ctx, g_blur, g_pure, g_nz):
        rgb, ccw, noise_raw = ctx.saved_tensors
        N, K, gamma = ctx.meta
        g_rgb = torch.empty_like(rgb)
        g_ccw = torch.empty_like(ccw) if ccw is not None else None
        g_nr = torch.empty_like(noise_raw) if noise_raw is not None else None
        check(lib().lush_blend_tonemap_bwd(ptr(rgb), ptr(ccw), ptr(noise_raw), ptr(_f32c(g_blur)),
                                           ptr(_f32c(g_pure)), ptr(_f32c(g_nz)) if noise_raw is not None else None,
                                           ptr(g_rgb), ptr(g_ccw), ptr(g_nr), N, K, gamma, stream()),
              "lush_blend_tonemap_bwd")
        return g_rgb, g_ccw, g_nr, None, None


def blend_tonemap(rgb, ccw=None, noise_raw=None, K: int = 1, gamma: bool = True):
    """Returns (tm(blend + 0.1*sigmoid(noise)), tm(blend), 0.1*sigmoid(noise))."""
    if not rgb.is_cuda:
        raise RuntimeError("lush_nerf_b200.blend_tonemap needs CUDA tensors; there is no CPU fallback")
    return _BlendTonemap.apply(rgb.reshape(-1, 3), ccw, noise_raw, int(K), gamma)


# ------------------------------------------------------------------------------------------------
# Training loss (reference run_lushnerf.py:652-661), value + gradient in one launch
# ------------------------------------------------------------------------------------------------
class _BlurLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, blur, blur0, target):
        blur, blur0, target = _f32c(blur), _f32c(blur0), _f32c(target)
        loss = _empty((), blur)
        g, g0 = torch.empty_like(blur), torch.empty_like(blur0)
        check(lib().lush_blur_loss(ptr(blur, "rgb_blur"), ptr(blur0, "rgb0_blur"), ptr(target, "target"),
                                   blur.numel(), ptr(loss), ptr(g), ptr(g0), stream()), "lush_blur_loss")
        ctx.save_for_backward(g, g0)
        return loss

    @staticmethod
    def backward(ctx, g_loss):
        g, g0 = ctx.saved_tensors
        return g * g_loss, g0 * g_loss, None


def blur_loss(rgb_blur, rgb0_blur, target):
    """0.5*MSE + 0.5*L1 of the fine and the coarse blurred colours against the target pixels."""
    if not rgb_blur.is_cuda:
        raise RuntimeError("lush_nerf_b200.blur_loss needs CUDA tensors; there is no CPU fallback")
    if rgb_blur.shape != target.shape or rgb0_blur.shape != target.shape:
        raise ValueError(f"blur_loss: shapes differ: {tuple(rgb_blur.shape)}, {tuple(rgb0_blur.shape)}, "
                         f"{tuple(target.shape)}")
    return _BlurLoss.apply(rgb_blur, rgb0_blur, target)
