"""Drop-in for the hot-path part of the reference's utils/run_lushnerf_helpers.py.

Same names, argument meaning and error behaviour as the reference (file:line cited per symbol), but the
numerics run in the sm_100a kernels behind include/lushnerf_b200.h.  Host-only preprocessing that the
reference keeps in this file (frequency masks, visualisation, MS-SSIM) is out of scope.
"""
from __future__ import annotations

import warnings

import numpy as np
import torch
import torch.nn as nn

from .. import ops
from .._lib import MlpDesc

HALF_PIX = 0.5
_FALLBACK_WARNED = set()

# ---- small host-side metrics the driver star-imports (reference helpers:9-12)
img2mse = lambda x, y: torch.mean((x - y) ** 2)
img2l1 = lambda x, y: torch.mean(torch.abs(x - y))
mse2psnr = lambda x: -10.0 * torch.log(x) / torch.log(torch.tensor([10.0], device=x.device))
to8b = lambda x: (255 * np.clip(x, 0, 1)).astype(np.uint8)


# ------------------------------------------------------------------------------------------------
# positional encoding (reference helpers:311-361)
# ------------------------------------------------------------------------------------------------
class Embedder(nn.Module):
    """gamma(x) = [x, sin(2^0 x), cos(2^0 x), ..., sin(2^(L-1) x), cos(2^(L-1) x)].
    Stand-alone calls evaluate it with torch ops (API compatibility); inside NeRFAll the encoding is
    fused into the MLP kernel and this module only carries `num_freqs` / `out_dim`."""

    def __init__(self, input_dims=3, num_freqs=10, include_input=True, **_unused):
        super().__init__()
        self.input_dims, self.num_freqs, self.include_input = input_dims, num_freqs, include_input
        self.out_dim = input_dims * ((1 if include_input else 0) + 2 * num_freqs)
        self.freq_bands = 2.0 ** torch.linspace(0.0, num_freqs - 1, steps=num_freqs) if num_freqs else torch.zeros(0)

    def forward(self, inputs):
        outs = [inputs] if self.include_input else []
        for l in range(self.num_freqs):
            f = float(2.0 ** l)
            outs += [torch.sin(inputs * f), torch.cos(inputs * f)]
        return torch.cat(outs, -1)


def get_embedder(multires, i=0, input_dim=3):
    """(embed_fn, out_dim); i == -1 -> identity (reference helpers:347-361)."""
    if i == -1:
        return nn.Identity(), 3
    e = Embedder(input_dims=input_dim, num_freqs=multires, include_input=True)
    return e, e.out_dim


# ------------------------------------------------------------------------------------------------
# scene / noise MLP (reference helpers:365-423, 456-512)
# ------------------------------------------------------------------------------------------------
class _FlatParamMLP(nn.Module):
    """nn.Linear stack whose parameters are views of ONE flat fp32 buffer (state_dict names, shapes and
    order are the reference's).  The kernels read the flat buffer; the optimiser keeps mutating the
    nn.Parameters in place.  If something re-materialises the parameters individually (.cuda(), .to(),
    DataParallel replicas) the arena is rebuilt lazily on the next forward."""

    out_mode = 0

    def __init__(self, D=8, W=256, input_ch=3, input_ch_views=3, output_ch=4, skips=(4,), use_viewdirs=False,
                 use_awp=False):
        super().__init__()
        self.D, self.W = D, W
        self.input_ch, self.input_ch_views = input_ch, input_ch_views
        self.skips = list(skips)
        self.use_viewdirs = use_viewdirs
        self.use_awp = use_awp
        if not use_viewdirs:
            raise NotImplementedError("lush_nerf_b200 builds the use_viewdirs=True topology only "
                                      "(all reference configs set use_viewdirs; NeRF_Noise requires it)")
        if len(self.skips) > 1:
            raise NotImplementedError("a single skip connection is supported")
        self.pts_linears = nn.ModuleList(
            [nn.Linear(input_ch, W)] + [nn.Linear(W, W) if i not in self.skips else nn.Linear(W + input_ch, W)
                                        for i in range(D - 1)])
        self.views_linears = nn.ModuleList([nn.Linear(input_ch_views + W, W // 2)])
        self.feature_linear = nn.Linear(W, W)
        self.alpha_linear = nn.Linear(W, 1)
        self.rgb_linear = nn.Linear(W // 2, 3)
        self._flat = None
        if (input_ch - 3) % 6 or (input_ch_views - 3) % 6:
            raise ValueError("input_ch / input_ch_views must be 3 + 6*L (positional encoding widths)")
        skip = self.skips[0] if self.skips else -1
        self._desc = MlpDesc(D, W, (input_ch - 3) // 6, (input_ch_views - 3) // 6, skip, self.out_mode)

    # -- flat arena --------------------------------------------------------------------------
    def _ordered_params(self):
        ps = []
        for l in self.pts_linears:
            ps += [l.weight, l.bias]
        for l in (self.views_linears[0], self.feature_linear, self.alpha_linear, self.rgb_linear):
            ps += [l.weight, l.bias]
        return ps

    def flat_params(self) -> torch.Tensor:
        """The flat buffer (rebuilding the arena if the parameters no longer alias it)."""
        ps = self._ordered_params()
        flat = self._flat
        ok = flat is not None and flat.device == ps[0].device
        if ok:
            off = flat.data_ptr()
            for p in ps:
                if p.data_ptr() != off or p.dtype != torch.float32:
                    ok = False
                    break
                off += p.numel() * 4
        if not ok:
            with torch.no_grad():
                flat = torch.cat([p.detach().reshape(-1).float() for p in ps])
                off = 0
                for p in ps:
                    n = p.numel()
                    p.data = flat[off:off + n].view(p.shape)
                    off += n
            self._flat = flat
        return flat

    @property
    def desc(self) -> MlpDesc:
        return self._desc

    def impl(self, training: bool = False) -> int:
        """1 = tcgen05 fp16-operand kernels when they cover this topology (and this mode), else 0 = fp32 SIMT.
        A scene network the tensor-core kernels do not cover (width != 256, ...) runs on the ~50x slower exact-fp32
        SIMT kernels: that is announced once per topology instead of happening silently."""
        want = getattr(self, "mlp_impl", None)
        want = ops.default_mlp_impl() if want is None else want
        if ops.mlp_impl_supported(self._desc, want, training):
            return want
        if want == 1 and self.out_mode == 0:
            key = (self.D, self.W, self.input_ch, self.input_ch_views, tuple(self.skips), bool(training))
            if key not in _FALLBACK_WARNED:
                _FALLBACK_WARNED.add(key)
                warnings.warn(f"lush_nerf_b200: the tcgen05 MLP kernels do not cover this topology (D={self.D}, "
                              f"W={self.W}, skips={self.skips}, training={bool(training)}); it runs on the exact-fp32 "
                              "SIMT kernels (~20 TFLOP/s instead of ~1000)", RuntimeWarning, stacklevel=3)
        return 0

    def render(self, ray_batch, z):
        """Fused encode+MLP on the samples x = o + d*z of each ray -> [R, S, C]."""
        ps = self._ordered_params()
        training = torch.is_grad_enabled() and (ray_batch.requires_grad or any(p.requires_grad for p in ps))
        return ops.mlp_apply(ray_batch, z, self._desc, self.impl(training), self.flat_params(), ps)

    def forward(self, x, force_naive=True, inference=False):
        """Reference call surface (helpers:394 / :483): x = [gamma(p) | gamma(d)] -> [M, 4] (NeRF) or [M, 3]
        (NeRF_Noise).  The fused kernels consume points, not encodings, and gamma() carries its argument in its first
        three columns (include_input=True, helpers:325-326): p = x[:, 0:3], d = x[:, in_ch:in_ch+3] are read back and
        the points go through the same fused encode+MLP kernels as everything else.  x must therefore BE the encoding
        of those columns -- the only way the reference ever calls a network (models/lushnerf.py:238-257)."""
        if x.shape[-1] != self.input_ch + self.input_ch_views:
            raise ValueError(f"expected {self.input_ch + self.input_ch_views} input channels, got {x.shape[-1]}")
        lead = x.shape[:-1]
        x2 = x.reshape(-1, x.shape[-1])
        pts, dirs = x2[:, 0:3], x2[:, self.input_ch:self.input_ch + 3]
        rb = torch.cat([pts, torch.zeros_like(pts), torch.zeros_like(pts[:, :2]), dirs], -1)
        out = self.render(rb, torch.zeros_like(pts[:, :1]))
        return out.reshape(*lead, out.shape[-1])


class NeRF(_FlatParamMLP):
    """8x256 scene MLP -> [rgb(3), sigma] (reference helpers:365-423)."""
    out_mode = 0


class NeRF_Noise(_FlatParamMLP):
    """Half-size noise MLP -> rgb(3); alpha_linear exists but never receives a gradient
    (reference helpers:456-512)."""
    out_mode = 1


# ------------------------------------------------------------------------------------------------
# tone mapping (reference helpers:134-183)
# ------------------------------------------------------------------------------------------------
class ToneMapping(nn.Module):
    def __init__(self, map_type: str):
        super().__init__()
        assert map_type in ['none', 'gamma', 'learn', 'ycbcr', 'split_linear']
        if map_type in ('learn', 'split_linear', 'ycbcr'):
            raise NotImplementedError(f"tone_mapping_type={map_type!r}: only 'none' and 'gamma' are on the "
                                      "hot path of the shipped configs (configs/*_lushnerf: gamma)")
        self.map_type = map_type

    def forward(self, x):
        if self.map_type == 'none':
            return x
        shape = x.shape
        _, pure, _ = ops.blend_tonemap(x.reshape(-1, 3), None, None, 1, True)
        return pure.reshape(shape)


# ------------------------------------------------------------------------------------------------
# ray helpers (reference helpers:517-562)
# ------------------------------------------------------------------------------------------------
def get_rays(H, W, K, c2w):
    dev = c2w.device
    j, i = torch.meshgrid(torch.linspace(0, H - 1, H, device=dev), torch.linspace(0, W - 1, W, device=dev),
                          indexing="ij")
    dirs = torch.stack([(i + (HALF_PIX - K[0][2])) / K[0][0], -(j + (HALF_PIX - K[1][2])) / K[1][1],
                        -torch.ones_like(i)], -1)
    rays_d = torch.sum(dirs[..., None, :] * c2w[:3, :3], -1)
    rays_o = c2w[:3, -1].expand(rays_d.shape)
    return rays_o, rays_d


def get_rays_np(H, W, K, c2w):
    i, j = np.meshgrid(np.arange(W, dtype=np.float32), np.arange(H, dtype=np.float32), indexing='xy')
    dirs = np.stack([(i + (HALF_PIX - K[0][2])) / K[0][0], -(j + (HALF_PIX - K[1][2])) / K[1][1],
                     -np.ones_like(i)], -1)
    rays_d = np.sum(dirs[..., np.newaxis, :] * c2w[:3, :3], -1)
    rays_o = np.broadcast_to(c2w[:3, -1], np.shape(rays_d))
    return rays_o, rays_d


def ndc_rays(H, W, focal, near, rays_o, rays_d):
    """NDC warp as a stand-alone call (the fused version lives in ops.ray_prep)."""
    t = -(near + rays_o[..., 2]) / rays_d[..., 2]
    rays_o = rays_o + t[..., None] * rays_d
    cx, cy = -1.0 / (W / (2.0 * focal)), -1.0 / (H / (2.0 * focal))
    ox, oy, oz = rays_o[..., 0], rays_o[..., 1], rays_o[..., 2]
    o = torch.stack([cx * ox / oz, cy * oy / oz, 1.0 + 2.0 * near / oz], -1)
    d = torch.stack([cx * (rays_d[..., 0] / rays_d[..., 2] - ox / oz),
                     cy * (rays_d[..., 1] / rays_d[..., 2] - oy / oz), -2.0 * near / oz], -1)
    return o, d


# ------------------------------------------------------------------------------------------------
# hierarchical sampling (reference helpers:566-609)
# ------------------------------------------------------------------------------------------------
def sample_pdf(bins, weights, N_samples, det=False, pytest=False):
    """Same signature and random-number consumption as the reference: det -> u = linspace(0,1,N),
    else one torch.rand([..., N]); pytest -> numpy seed 0.  Search / gather / lerp run in the warp-level
    kernel (csrc/sampler.cu).  det=True (and LUSH_SAMPLE_STRICT=1) feed the kernel the cdf built by the reference's
    own torch ops on this device (ops.reference_cdf): indices and samples are then bit-identical to the reference's
    searchsorted path on the same GPU; otherwise the cdf is built inside the kernel (ops.sample_strict)."""
    lead = list(bins.shape[:-1])
    b2 = bins.reshape(-1, bins.shape[-1])
    w2 = weights.reshape(-1, weights.shape[-1])
    if pytest:
        np.random.seed(0)
        if det:
            u = torch.tensor(np.linspace(0.0, 1.0, N_samples), dtype=torch.float32, device=bins.device)
        else:
            u = torch.tensor(np.random.rand(*(lead + [N_samples])), dtype=torch.float32,
                             device=bins.device).reshape(-1, N_samples)
    elif det:
        u = ops.linspace01(N_samples, bins.device)
    else:
        u = torch.rand(lead + [N_samples], device=bins.device).reshape(-1, N_samples)
    cdf = ops.reference_cdf(w2) if ops.sample_strict(det) else None
    return ops.sample_pdf_cuda(b2, w2, u, cdf=cdf).reshape(lead + [N_samples])


# ------------------------------------------------------------------------------------------------
# checkpoint loading (reference helpers:612-628)
# ------------------------------------------------------------------------------------------------
def smart_load_state_dict(model: nn.Module, state_dict: dict):
    def strip(k):
        return k[len("module."):] if k.startswith("module.") else k

    if "network_fn_state_dict" in state_dict:
        sd = {"mlp_coarse." + strip(k): v for k, v in state_dict["network_fn_state_dict"].items()}
        sd.update({"mlp_fine." + strip(k): v for k, v in state_dict["network_fine_state_dict"].items()})
    elif "network_state_dict" in state_dict:
        sd = {strip(k): v for k, v in state_dict["network_state_dict"].items()}
    else:
        sd = dict(state_dict)
    if isinstance(model, nn.DataParallel):
        sd = {"module." + k: v for k, v in sd.items()}
    model.load_state_dict(sd, strict=False)


# ------------------------------------------------------------------------------------------------
# frequency masks (reference helpers:630-662): host-side, once per run -- which pixels receive blur-kernel ray
# gradients in stage 2 (run_lushnerf.py:283, :627-628).  numpy FFT instead of the reference's cv2.dft.
# ------------------------------------------------------------------------------------------------
def frequency_mask(image, radius=50):
    """image [H,W,3] uint8 -> [H,W] {0,1}: centred low-pass window of +-radius frequencies, magnitude of the filtered
    image normalised to 0..255 and thresholded at 48."""
    img = np.asarray(image)
    gray = (0.299 * img[..., 0].astype(np.float64) + 0.587 * img[..., 1] + 0.114 * img[..., 2])
    if img.dtype == np.uint8:
        gray = np.floor(gray + 0.5)            # cv2.cvtColor rounds to uint8 before the float conversion
    spec = np.fft.fftshift(np.fft.fft2(gray.astype(np.float32)))
    rows, cols = gray.shape
    crow, ccol = rows // 2, cols // 2
    keep = np.zeros((rows, cols), bool)
    keep[crow - radius:crow + radius, ccol - radius:ccol + radius] = True     # (a negative start wraps, as upstream)
    low = np.abs(np.fft.ifft2(np.fft.ifftshift(spec * keep)))     # cv2.idft is unscaled; the normalisation below removes it
    span = low.max() - low.min()
    low = (low - low.min()) / (span if span > 0 else 1.0) * 255
    return (low > 48).astype(np.float32)


def get_masks_for_images(images, radius=50):
    """images [V,H,W,3] uint8 -> masks [V,H,W] uint8."""
    masks = np.zeros(images.shape[:-1], dtype=np.uint8)
    for i, image in enumerate(images):
        masks[i] = frequency_mask(image, radius)
    return masks


def compute_mean_with_confidence(A, certainty, threshold=0.2):
    """Confidence-masked mean over views (reference helpers:665-688): A [V,P,3], certainty [V,P]."""
    mask = (certainty >= threshold).unsqueeze(-1).to(A.dtype)
    total = (A * mask).sum(0)
    count = mask.sum(0)
    return torch.where(count > 0, total / count.clamp_min(1.0), torch.zeros_like(total))
