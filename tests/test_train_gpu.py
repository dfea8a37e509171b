"""GPU: the training step (forward + loss + backward + Adam) through lush_nerf_b200.train.Trainer, eagerly and
captured in one CUDA graph.  With perturb=0 and raw_noise_std=0 the step is deterministic, so the graph replay
must reproduce the eager losses and parameters bit for bit (same kernels, same order)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

H, W, FOCAL = 40, 60, 50.0
DEV = "cuda"


def rel_l2(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def _make(seed=0):
    from lush_nerf_b200.synthetic import SyntheticLLFF, build_model, poster_args
    from lush_nerf_b200.train import Trainer
    torch.manual_seed(seed)
    np.random.seed(seed)
    model = build_model(poster_args()).cuda()
    rk = dict(perturb=0.0, N_importance=64, N_samples=64, use_viewdirs=True, white_bkgd=False, raw_noise_std=0.0,
              inference=False, near=0.0, far=1.0)
    tr = Trainer(model, H, W, np.array([[FOCAL, 0, 0.5 * W], [0, FOCAL, 0.5 * H], [0, 0, 1]]), lrate=5e-4, render_kwargs=rk)
    data = SyntheticLLFF(H, W, FOCAL, seed=0, max_views=2)
    batches = [{k: v.cuda() for k, v in data.batch(i, 128).items()} for i in range(6)]
    return tr, batches


def test_eager_steps_reduce_the_loss():
    tr, batches = _make()
    losses = [float(tr.step(batches[0]).item()) for _ in range(8)]
    assert np.isfinite(losses).all()
    assert losses[-1] < losses[0]


def test_cuda_graph_step_matches_eager_step():
    tr_e, batches = _make()
    tr_g, _ = _make()
    for a, b in zip(tr_e.model.parameters(), tr_g.model.parameters()):
        assert torch.equal(a, b)
    # enable_cuda_graph executes 3 warm-up steps on its sample batch (the capture itself executes nothing)
    for _ in range(3):
        tr_e.step(batches[0])
    tr_g.enable_cuda_graph(batches[0])
    for i in range(1, 5):
        le = float(tr_e.step(batches[i]).item())
        lg = float(tr_g.step(batches[i]).item())
        assert abs(le - lg) <= 2e-5 * max(1.0, abs(le)), (i, le, lg)   # fp32 atomics in wgrad: order-dependent
    # Adam turns order-dependent fp32-atomic noise in near-zero gradients into +-lr steps, so compare in L2
    pe = torch.cat([p.detach().reshape(-1) for p in tr_e.model.parameters()])
    pg = torch.cat([p.detach().reshape(-1) for p in tr_g.model.parameters()])
    rel = ((pe - pg).norm() / pe.norm()).item()
    assert rel <= 2e-3, rel
    assert tr_g.check_numerics() == 0
    # learning rates live in device memory: an edit between replays reaches the captured optimizer launch
    for grp in tr_g.optimizer.param_groups:
        grp["lr"] = 0.0
    tr_g.step(batches[5])
    pg2 = torch.cat([p.detach().reshape(-1) for p in tr_g.model.parameters()])
    assert torch.equal(pg, pg2)


def test_multi_tensor_adam_matches_torch_adam():
    """lush_nerf_b200.optim.Adam (one launch per 64 tensors) vs torch.optim.Adam on 70 tensors of odd sizes, two
    groups with different learning rates, a parameter that never receives a gradient, one that starts late,
    a learning-rate edit, and a state_dict round trip into a fresh optimizer."""
    from lush_nerf_b200.optim import Adam
    g = torch.Generator().manual_seed(3)
    sizes = [(256, 256), (3,), (1,), (128, 283), (7, 5, 3)] + [(17 + i,) for i in range(65)]
    base = [torch.randn(s, generator=g) for s in sizes]
    mine = [b.clone().to(DEV).requires_grad_(True) for b in base]
    ref = [b.clone().to(DEV).requires_grad_(True) for b in base]

    def make(cls, ps):
        return cls([{"params": ps[:40]}, {"params": ps[40:], "lr": 2e-3}], lr=5e-4, betas=(0.9, 0.999), eps=1e-8)
    om, orf = make(Adam, mine), make(torch.optim.Adam, ref)
    for it in range(12):
        if it == 8:                                      # checkpoint round trip through the reference's format
            sd = om.state_dict()
            om = make(Adam, mine)
            om.load_state_dict(sd)
        if it == 5:
            for o in (om, orf):
                o.param_groups[0]["lr"] = 1e-4
        for a, b in zip(mine, ref):
            a.grad, b.grad = None, None
        for k, (a, b) in enumerate(zip(mine, ref)):
            if k == 2 or (k == 4 and it < 3):            # never / late
                continue
            gr = torch.randn(a.shape, generator=g).to(DEV) * (10.0 ** ((k % 5) - 3))
            a.grad, b.grad = gr.clone(), gr.clone()
        om.step()
        orf.step()
    for k, (a, b) in enumerate(zip(mine, ref)):
        assert rel_l2(a.detach(), b.detach()) < 2e-6, k
    assert torch.equal(mine[2].detach().cpu(), base[2])
    st, st_ref = om.state_dict()["state"], orf.state_dict()["state"]
    assert set(st) == set(st_ref) and 2 not in st
    assert float(st[4]["step"]) == 9.0 and float(st[0]["step"]) == 12.0
    assert rel_l2(st[3]["exp_avg_sq"], st_ref[3]["exp_avg_sq"]) < 2e-6


def test_index_batches_equal_table_batches():
    """Trainer.set_scene + {images_idx, pix_idx} batches (rays, targets and mask bits staged on the device by ONE kernel
    from the resident poses / images / masks) must be the same training step as the reference's batch form
    {rays, rgbsf, fq_mask, images_idx} sliced from a host-side ray table (run_lushnerf.py:538-616)."""
    from lush_nerf_b200.synthetic import synthetic_poses
    from lush_nerf_b200.utils.run_lushnerf_helpers import get_rays_np
    tr_i, _ = _make()
    tr_t, _ = _make()
    V, n = 5, 96
    g = torch.Generator().manual_seed(5)
    poses = torch.from_numpy(synthetic_poses(V, 0))
    images = torch.rand(V, H, W, 3, generator=g)
    masks = torch.rand(V, H, W, generator=g) < 0.5
    K = tr_i.K
    tr_i.set_scene(poses, images, masks)
    for step in range(3):
        v = torch.randint(0, V, (n, 1), generator=g)
        p = torch.randint(0, H * W, (n,), generator=g)
        rays = torch.empty(n, 3, 2)
        for i in range(n):
            ro, rd = get_rays_np(H, W, K, poses[v[i, 0]].numpy())
            rays[i, :, 0] = torch.from_numpy(np.ascontiguousarray(ro.reshape(-1, 3)[p[i]])).float()
            rays[i, :, 1] = torch.from_numpy(rd.reshape(-1, 3)[p[i]].astype(np.float32))
        table = dict(rays=rays.cuda(), rgbsf=images.reshape(V, -1, 3)[v[:, 0], p].cuda(),
                     fq_mask=masks.reshape(V, -1)[v[:, 0], p].cuda(), images_idx=v.cuda())
        li = float(tr_i.step(dict(images_idx=v.cuda(), pix_idx=p.cuda())).item())
        lt = float(tr_t.step(table).item())
        assert abs(li - lt) <= 2e-5 * max(1.0, abs(lt)), (step, li, lt)
    assert tr_i.check_numerics() == 0
    with pytest.raises(RuntimeError):
        tr_t.step(dict(images_idx=v.cuda(), pix_idx=p.cuda()))          # no scene tables set
    # an out-of-range pixel index is clamped, flagged, and surfaces as IndexError from the guard
    bad = dict(images_idx=v.cuda(), pix_idx=(p + H * W).cuda())
    tr_i.step(bad)
    with pytest.raises(IndexError):
        tr_i.check_numerics()
