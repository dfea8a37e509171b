"""GPU: the training step (forward + loss + backward + Adam) through lush_nerf_b200.train.Trainer, eagerly and
captured in one CUDA graph.  With perturb=0 and raw_noise_std=0 the step is deterministic, so the graph replay
must reproduce the eager losses and parameters bit for bit (same kernels, same order)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

H, W, FOCAL = 40, 60, 50.0


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
