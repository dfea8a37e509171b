"""`sample_pdf` exactness on the PUBLIC path (helpers.sample_pdf, NeRFAll's fused sample_merge).

north_star: "sample_pdf indices with det=True must be bit-exact".  The reference's indices come from
`torch.searchsorted(cdf, u, right=True)` on a cdf built by `torch.sum` / `torch.cumsum`, whose fp32 association order
depends on the device that runs them (CPU cumsum accumulates in double, CUDA cumsum is a parallel fp32
scan).  Two statements are tested:

  strict (default for det=True): the cdf is built by the reference's own op sequence with torch ops on the GPU
      (ops.reference_cdf) and the kernel searches / gathers / interpolates -> indices AND samples bit-identical to the
      reference's op sequence executed on the same device (here: the oracle restatement of helpers:566-609 run on CUDA
      tensors), and bit-identical to the reference's CPU goldens when handed the golden cdf.
  fused (in-kernel cdf, default for random u): every sample is accounted for by the rule of SURVEY.md 7.3-4 --
      an index may differ from the torch-order result only where u is within 4 ulp of the cdf entries that separate the
      two indices; with equal indices
      the sample moves by at most the forward error of (u - c0) / denom under a 4-ulp change of the cdf entries;
      where the interpolation formula is discontinuous (denom crossing 1e-5, helpers:604-605) the sample stays in the
      closed bin.  `classify_against_torch_order` returns the counts; `unexplained` must be 0.
"""
import numpy as np
import pytest
import torch

from oracle import lush_oracle as O

DEV = "cuda"
ULP1 = float(np.spacing(np.float32(1.0)))     # ulp of values in [1, 2): upper bound of the ulp of any cdf entry


def classify_against_torch_order(bins, u, s, inds, s_t, inds_t, cdf_t, k_ulp=4.0):
    """All arguments on the CPU.  Returns counts {equal, value_moved, index_flip, denom_switch, unexplained}."""
    R, N = s.shape
    nb = bins.shape[1]
    uu = (u.expand(R, N) if u.dim() == 1 else u).double()
    cdf = cdf_t.double()
    tol_c = k_ulp * ULP1
    same_idx = inds == inds_t
    below = (inds_t - 1).clamp(min=0)
    above = inds_t.clamp(max=nb - 1)
    c0, c1 = torch.gather(cdf, 1, below), torch.gather(cdf, 1, above)
    b0, b1 = torch.gather(bins.double(), 1, below), torch.gather(bins.double(), 1, above)
    denom = c1 - c0
    width = (b1 - b0).abs()
    diff = (s.double() - s_t.double()).abs()
    near_switch = (denom - 1e-5).abs() <= 2 * tol_c                 # the `denom < 1e-5 -> 1` branch may differ
    eff = torch.where(denom < 1e-5, torch.ones_like(denom), denom)
    # forward error of b0 + (u - c0) / denom * (b1 - b0) when c0, c1 each move by <= tol_c (u - c0 <= denom)
    bound = (3 * tol_c / eff) * width + 4e-7
    ok_value = same_idx & (diff <= bound)
    ok_switch = same_idx & near_switch & (diff <= width + 1e-6)
    # index flips: u within tol_c of the cdf entry that separates the two indices
    # (several when the separating entries are equal: zero-weight bins give runs of identical cdf values; the cdf is
    # monotone, so it is enough that the first and the last entry of the run cdf[lo .. hi-1] touch u)
    lo, hi = torch.minimum(inds, inds_t), torch.maximum(inds, inds_t)
    adjacent = hi > lo
    edge_lo = torch.gather(cdf, 1, lo.clamp(max=nb - 1))
    edge_hi = torch.gather(cdf, 1, (hi - 1).clamp(min=0, max=nb - 1))
    touches = ((uu - edge_lo).abs() <= tol_c) & ((uu - edge_hi).abs() <= tol_c)
    # at a flip both branches interpolate to the shared bin edge (within the neighbouring bins' widths when one of them
    # is degenerate, denom < 1e-5)
    e_lo, e_hi = torch.gather(bins.double(), 1, (lo - 1).clamp(min=0)), torch.gather(bins.double(), 1, hi.clamp(max=nb - 1))
    ok_flip = (~same_idx) & adjacent & touches & (s.double() >= e_lo - 1e-6) & (s.double() <= e_hi + 1e-6)
    equal = same_idx & (diff == 0)
    explained = ok_value | ok_switch | ok_flip
    return {"n": int(s.numel()), "equal": int(equal.sum()), "value_moved": int((ok_value & ~equal).sum()),
            "index_flip": int(ok_flip.sum()), "denom_switch": int((ok_switch & ~ok_value).sum()),
            "unexplained": int((~explained).sum()), "max_diff": float(diff.max())}


def _peaky_case(R, S, seed, device):
    """LLFF-like peaky weights: relu(3 randn) compositing of a jittered z row, done with torch on `device`."""
    g = torch.Generator(device=device).manual_seed(seed)
    t = torch.linspace(0, 1, S, device=device)
    mids = 0.5 * (t[1:] + t[:-1])
    lower, upper = torch.cat([t[:1], mids]), torch.cat([mids, t[-1:]])
    z = lower + (upper - lower) * torch.rand(R, S, device=device, generator=g)
    sigma = torch.relu(3.0 * torch.randn(R, S, device=device, generator=g))
    alpha = 1 - torch.exp(-sigma * 5.0 / S)
    alpha[:, -1] = 1.0
    T = torch.cumprod(torch.cat([torch.ones(R, 1, device=device), 1 - alpha + 1e-10], -1), -1)[:, :-1]
    return z, alpha * T


@pytest.mark.gpu
@pytest.mark.parametrize("R,S,N", [(4096, 64, 64), (512, 32, 32), (256, 128, 128), (64, 256, 256), (33, 24, 16)])
def test_public_sample_pdf_det_is_bit_exact_vs_reference_ops_on_this_device(R, S, N):
    from lush_nerf_b200 import ops
    from lush_nerf_b200.utils.run_lushnerf_helpers import sample_pdf
    z, w = _peaky_case(R, S, S * 13 + N, DEV)
    z_mid = 0.5 * (z[:, 1:] + z[:, :-1])
    s_ref, inds_ref, cdf_ref = O.sample_pdf(z_mid, w[:, 1:-1], torch.linspace(0., 1., N, device=DEV), order="torch")
    assert s_ref.is_cuda                                             # the reference's ops, on this GPU
    s = sample_pdf(z_mid, w[:, 1:-1], N, det=True)
    assert torch.equal(s, s_ref)
    s2, inds = ops.sample_pdf_cuda(z_mid, w[:, 1:-1], ops.linspace01(N, DEV), cdf=ops.reference_cdf(w[:, 1:-1]),
                                   return_inds=True)
    assert torch.equal(inds, inds_ref) and torch.equal(s2, s_ref)
    assert torch.equal(ops.reference_cdf(w[:, 1:-1]), cdf_ref)
    # the fused fine-sample placement of NeRFAll._render takes the same strict route for deterministic u
    zf, zs, _ = ops.sample_merge(z, w, ops.linspace01(N, DEV))
    assert torch.equal(zs, s_ref)
    assert torch.equal(zf, torch.sort(torch.cat([z, s_ref], -1), -1)[0])


@pytest.mark.gpu
def test_strict_mode_switch(monkeypatch):
    from lush_nerf_b200 import ops
    monkeypatch.delenv("LUSH_SAMPLE_STRICT", raising=False)
    assert ops.sample_strict(True) and not ops.sample_strict(False)
    monkeypatch.setenv("LUSH_SAMPLE_STRICT", "1")
    assert ops.sample_strict(False)
    monkeypatch.setenv("LUSH_SAMPLE_STRICT", "0")
    assert not ops.sample_strict(True)
    z, w = _peaky_case(300, 64, 5, DEV)
    u = torch.rand(300, 64, device=DEV)
    a = ops.sample_merge(z, w, u, strict=True)[1]
    b, _, _ = O.sample_pdf(0.5 * (z[:, 1:] + z[:, :-1]), w[:, 1:-1], u, order="torch")
    assert torch.equal(a, b)                                         # random u, strict on request: still bit-exact


@pytest.mark.gpu
def test_fused_cdf_det_frame_flip_count():
    """One 400 x 600 deterministic frame (240 000 rays, 64 + 64): in-kernel cdf order vs the reference's ops on this
    device.  Every sample must be accounted for by the 7.3-4 rule; the counts are the parity report's numbers."""
    from lush_nerf_b200 import ops
    R, S, N = 240000, 64, 64
    z, w = _peaky_case(R, S, 7, DEV)
    z_mid = 0.5 * (z[:, 1:] + z[:, :-1])
    u = ops.linspace01(N, DEV)
    s_t, inds_t, cdf_t = O.sample_pdf(z_mid, w[:, 1:-1], u, order="torch")
    s, inds = ops.sample_pdf_cuda(z_mid, w[:, 1:-1], u, return_inds=True)          # fused: in-kernel cdf
    rep = classify_against_torch_order(z_mid.cpu(), u.cpu(), s.cpu(), inds.cpu(), s_t.cpu(), inds_t.cpu(), cdf_t.cpu())
    last = int((inds[:, -1] != inds_t[:, -1]).sum())
    print(f"fused-cdf det frame, {R} rays x {N} samples: {rep}; flips in the u = 1.0 column: {last}")
    assert rep["unexplained"] == 0, rep
    assert rep["index_flip"] <= 0.02 * rep["n"], rep
    s_strict = ops.sample_pdf_cuda(z_mid, w[:, 1:-1], u, cdf=ops.reference_cdf(w[:, 1:-1]))
    assert torch.equal(s_strict, s_t)


def test_classifier_flags_real_differences():
    """The rule must not be vacuous: a sample moved by a whole bin, or an index off by two, is `unexplained`."""
    g = torch.Generator().manual_seed(0)
    bins = torch.sort(torch.rand(8, 31, generator=g), -1)[0]
    w = torch.rand(8, 30, generator=g)
    u = torch.rand(8, 16, generator=g)
    s_t, inds_t, cdf_t = O.sample_pdf(bins, w, u, order="torch")
    ok = classify_against_torch_order(bins, u, s_t, inds_t, s_t, inds_t, cdf_t)
    assert ok["unexplained"] == 0 and ok["equal"] == ok["n"]
    bad_s = s_t.clone()
    bad_s[0, 0] += 0.05
    assert classify_against_torch_order(bins, u, bad_s, inds_t, s_t, inds_t, cdf_t)["unexplained"] == 1
    bad_i = inds_t.clone()
    bad_i[1, 1] = (bad_i[1, 1] + 2) % 31
    assert classify_against_torch_order(bins, u, s_t, bad_i, s_t, inds_t, cdf_t)["unexplained"] == 1
    s_k, inds_k, _ = O.sample_pdf(bins, w, u, order="kernel")        # the kernel's association order, emulated on the CPU
    rep = classify_against_torch_order(bins, u, s_k, inds_k, s_t, inds_t, cdf_t)
    assert rep["unexplained"] == 0, rep
