"""CPU: the C-ABI library loads, exports every symbol include/lushnerf_b200.h declares, validates its
arguments without touching a GPU, and the host-side mirror of the reference interface (state_dict
names, flat parameter arena, error behaviour) holds.  No compute calls here."""
import ctypes
import os
import re
import types

import numpy as np
import pytest
import torch

from lush_nerf_b200 import _lib
from oracle import lush_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def poster_args(**kw):
    ns = types.SimpleNamespace(
        blur_model_type="dpnerf", multires=10, i_embed=0, use_viewdirs=True, multires_views=4, N_importance=64,
        N_samples=64, netdepth=8, netwidth=256, netdepth_fine=8, netwidth_fine=256, rgb_activate="sigmoid",
        sigma_activate="relu", tone_mapping_type="gamma", render_rmnearplane=80, num_images=20)
    for k, v in kw.items():
        setattr(ns, k, v)
    return ns


def make_model(with_rbk=True, **kw):
    from lush_nerf_b200.models.lushnerf import RBK, NeRFAll
    rbk = None
    if with_rbk:
        rbk = RBK(num_img=20, view_embed_ch=64, D_rbk=4, W_rbk=64, skips_rbk=[4], num_motion_rbk=4, D_rbk_r=1,
                  W_rbk_r=32, output_ch_rbk_r=3, D_rbk_v=1, W_rbk_v=32, output_ch_rbk_v=3, D_rbk_w=1, W_rbk_w=32,
                  rbk_se_rv_window=0.1, use_dpnerf=True, rbk_use_origin=True, near=0.0, far=1.0, ndc=True)
    return NeRFAll(poster_args(**kw), rbk)


def test_header_symbols_all_exported():
    hdr = open(os.path.join(ROOT, "include", "lushnerf_b200.h")).read()
    declared = set(re.findall(r"\b(lush_[a-z0-9_]+)\s*\(", hdr))
    declared.discard("lush_mlp_desc")
    assert declared, "no declarations parsed"
    handle = ctypes.CDLL(_lib.LIB_PATH)
    for name in sorted(declared):
        assert hasattr(handle, name), f"{name} declared in the header but not exported"
    assert set(_lib.exported_symbols()) == declared
    assert _lib.lib().lush_abi_version() == 2


def test_argument_validation_without_gpu():
    L = _lib.lib()
    rc = L.lush_composite_fwd(None, None, None, 3, None, -1.0, 0, None, None, None, None, None, None, 4, 8, None)
    assert rc == -1 and b"null pointer" in L.lush_last_error()
    rc = L.lush_sample_pdf(None, None, None, None, 0, None, None, 4, 8, 8, None)
    assert rc == -1
    bad = _lib.MlpDesc(0, 256, 10, 4, 4, 0)
    assert L.lush_mlp_param_count(bad) == -1
    assert L.lush_mlp_param_count(_lib.MlpDesc(8, 256, 10, 4, 7, 0)) == -1  # skip == D-1: reference crashes too
    with pytest.raises(RuntimeError, match="failed"):
        _lib.check(rc, "lush_sample_pdf")


def test_param_counts_match_reference():
    L = _lib.lib()
    assert L.lush_mlp_param_count(_lib.MlpDesc(8, 256, 10, 4, 4, 0)) == 595844
    assert L.lush_mlp_param_count(_lib.MlpDesc(4, 128, 10, 4, 4, 1)) == 84548
    assert L.lush_mlp_impl_supported(_lib.MlpDesc(8, 256, 10, 4, 4, 0), 0) == 3
    assert L.lush_mlp_impl_supported(_lib.MlpDesc(8, 256, 10, 4, 4, 0), 1) & 1
    assert L.lush_mlp_impl_supported(_lib.MlpDesc(4, 128, 10, 4, 4, 1), 1) == 0   # noise MLP: SIMT only


def test_state_dict_names_shapes_and_aliases():
    m = make_model()
    sd = m.state_dict()
    assert len(sd) == 108 and sum(p.numel() for p in m.parameters()) == 1301353
    for net, (D, W) in (("mlp_coarse", (8, 256)), ("mlp_fine", (8, 256)), ("mlp_noise_coarse", (4, 128))):
        for name, shape in O.mlp_param_shapes(D, W, 63, 27, 4):
            assert tuple(sd[f"{net}.{name}"].shape) == shape
    # the RBK tensors are registered under several names
    for k in ("dbk_view_embedding.view_embed_layer.weight", "mlp_rbk.view_embedding_layer.view_embed_layer.weight",
              "blur_kernel_net.view_embed_layer.view_embed_layer.weight",
              "blur_kernel_net.RBK.view_embedding_layer.view_embed_layer.weight"):
        assert sd[k].data_ptr() == sd["dbk_view_embedding.view_embed_layer.weight"].data_ptr()
    assert m.mlp_rbk.r_linear.weight.abs().max() <= 1.0e-5  # identity-warp init (reference :62-68)


def test_flat_parameter_arena_survives_load_and_rebuilds(det_state):
    nets, rbk = det_state
    m = make_model()
    m.load_state_dict(O.reference_state_dict(nets, rbk), strict=False)
    net = m.mlp_coarse
    flat = net.flat_params()
    assert torch.equal(flat, O.flatten_params(nets["coarse"], 8, 256, 63, 27, 4))
    # in-place optimiser updates are seen through the arena
    with torch.no_grad():
        net.rgb_linear.bias.add_(1.0)
    assert torch.equal(net.flat_params()[-3:], nets["coarse"]["rgb_linear.bias"] + 1.0)
    assert net.flat_params().data_ptr() == flat.data_ptr()
    # re-materialised parameters (what .cuda()/.to() do) trigger a rebuild
    net.rgb_linear.bias.data = net.rgb_linear.bias.data.clone()
    flat2 = net.flat_params()
    assert flat2.data_ptr() != flat.data_ptr() and net.rgb_linear.bias.data_ptr() == flat2.data_ptr() + (595844 - 3) * 4


def test_smart_load_state_dict_strips_dataparallel_prefix(det_state):
    from lush_nerf_b200.utils.run_lushnerf_helpers import smart_load_state_dict
    nets, rbk = det_state
    m = make_model()
    ckpt = {"network_state_dict": {"module." + k: v for k, v in O.reference_state_dict(nets, rbk).items()}}
    smart_load_state_dict(m, ckpt)
    assert torch.equal(m.mlp_fine.alpha_linear.weight, nets["fine"]["alpha_linear.weight"])


def test_cpu_tensors_raise_no_fallback():
    m = make_model().train()
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m(4, 6, [[5., 0, 3], [0, 5., 2], [0, 0, 1]], rays=torch.randn(4, 3, 2),
          rays_info={"images_idx": torch.zeros(4, 1, dtype=torch.long)}, force_naive=True, retraw=True, perturb=1.,
          N_importance=16, N_samples=24, use_viewdirs=True, white_bkgd=False, raw_noise_std=1., inference=False,
          near=0., far=1.)
    from lush_nerf_b200.utils.run_lushnerf_helpers import sample_pdf
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        sample_pdf(torch.rand(2, 5).sort(-1)[0], torch.rand(2, 4), 6, det=True)


def test_forward_asserts_like_the_reference():
    m = make_model().train()
    with pytest.raises(AssertionError, match="specify rays"):
        m(4, 6, None, rays=None, rays_info={}, force_naive=True)
    m.eval()
    with pytest.raises(AssertionError, match="specify poses"):
        m(4, 6, None, poses=None, render_kwargs={})


def test_motion_plan_layer_program_host_side():
    """The layer program handed to lush_rbk_motion_{fwd,bwd} (ops.MotionPlan) for the poster blur kernel and for a
    topology with a skip concat -- pure host logic, no kernel call."""
    from lush_nerf_b200.models.lushnerf import RBK

    def make(D, W, skips, C):
        return RBK(num_img=7, view_embed_ch=C, D_rbk=D, W_rbk=W, skips_rbk=skips, num_motion_rbk=4, D_rbk_r=1, W_rbk_r=32,
                   output_ch_rbk_r=3, D_rbk_v=1, W_rbk_v=32, output_ch_rbk_v=3, D_rbk_w=1, W_rbk_w=32,
                   rbk_se_rv_window=0.1, use_dpnerf=True, rbk_use_origin=True, near=0.0, far=1.0, ndc=True).RBK
    plan = make(4, 64, [4], 64)._motion_plan()
    assert plan is not None and len(plan.layers) == 10 and plan.width == 64 + 4 * 64 + 3 * 32 + 12 + 12 + 5
    assert [seg for _, seg, _ in plan.layers][:4] == [(64, 64), (128, 64), (192, 64), (256, 64)]
    assert all(inp == [(256, 64)] for inp, _, _ in (plan.layers[4], plan.layers[6], plan.layers[8]))   # three branches
    assert [plan.layers[plan.heads[k]][1][1] for k in ("r", "v", "w")] == [12, 12, 5]
    assert [relu for _, _, relu in plan.layers] == [True] * 5 + [False, True, False, True, False]
    assert len(plan.params) == 20
    rbk = make(3, 48, [0], 16)
    plan = rbk._motion_plan()
    assert plan.layers[1][0] == [(0, 16), (16, 48)]        # layer 1 reads the concat [e, h] (reference :51-52, :125-126)
    assert plan.layers[1][1] == (64, 48)
    assert rbk._motion_plan() is plan                      # cached while the parameters are the same objects


def test_trainer_without_cuda_uses_torch_adam_and_cuda_only_optimizer_raises():
    from lush_nerf_b200.optim import Adam
    from lush_nerf_b200.train import Trainer
    m = make_model()
    tr = Trainer(m, 40, 60, np.eye(3), lrate=1e-3)
    assert type(tr.optimizer) is torch.optim.Adam and len(tr.optimizer.param_groups) == 2
    with pytest.raises(RuntimeError):
        tr.enable_cuda_graph({"rays": torch.zeros(1, 3, 2)})
    with pytest.raises(RuntimeError):
        Adam(list(m.parameters())[:2], lr=1e-3)
