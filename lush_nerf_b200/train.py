"""Ray-sharded training step: one process per GPU, every rank renders its own slice of the global ray
batch, ONE all-reduce of the flat fp32 gradient per step over NCCL (NVLink 5 / NVSwitch), identical Adam
on every rank.  This is the B200 replacement of the reference's nn.DataParallel wrap
(run_lushnerf.py:348) plus its loss / optimiser lines (:652-685); SURVEY.md 8e.

The path shards with no data-path collective: rays are independent, the only exchange is the sum of the
1.30 M parameter gradients (5.2 MB).
"""
from __future__ import annotations

import torch
import torch.distributed as dist

from . import ops


def lush_loss(rgb_blur, rgb0_blur, target):
    """0.5*MSE + 0.5*L1 on the fine and coarse outputs (reference run_lushnerf.py:652-661, stage 1/2), written
    with torch ops: the definition the fused kernel `ops.blur_loss` (what Trainer runs) is tested against."""
    d, d0 = rgb_blur - target, rgb0_blur - target
    return 0.5 * (d * d).mean() + 0.5 * d.abs().mean() + 0.5 * (d0 * d0).mean() + 0.5 * d0.abs().mean()


def shard_rows(n: int, rank: int, world: int):
    """Rows [lo, hi) of a global batch owned by `rank` -- the same contiguous dim-0 split DataParallel's
    scatter performs.  n must divide evenly so that the mean of local means is the global mean."""
    if n % world:
        raise ValueError(f"global batch {n} is not divisible by world size {world}")
    per = n // world
    return rank * per, (rank + 1) * per


class FlatGradAllReduce:
    """Packs the gradients of `params` into one flat fp32 buffer, all-reduces it once, unpacks."""

    def __init__(self, params, group=None):
        self.params = [p for p in params if p.requires_grad]
        self.group = group
        self.numel = sum(p.numel() for p in self.params)
        self.flat = None

    def __call__(self):
        if not dist.is_initialized() or dist.get_world_size(self.group) == 1:
            return
        # parameters without a gradient (NeRF_Noise.alpha_linear) contribute zeros and receive nothing back
        with_grad = [p for p in self.params if p.grad is not None]
        pieces = [p.grad.reshape(-1) if p.grad is not None else p.new_zeros(p.numel()) for p in self.params]
        self.flat = torch.cat(pieces)                     # one multi-tensor pack kernel instead of one copy per tensor
        dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=self.group)
        self.flat.div_(dist.get_world_size(self.group))
        views, off = [], 0
        for p in self.params:
            n = p.numel()
            if p.grad is not None:
                views.append(self.flat[off:off + n].view_as(p.grad))
            off += n
        torch._foreach_copy_([p.grad for p in with_grad], views)   # one multi-tensor unpack kernel


class Trainer:
    """model: NeRFAll (training mode).  step() = forward + loss + backward + grad all-reduce + Adam."""

    def __init__(self, model, H, W, K, lrate=5e-4, render_kwargs=None, group=None, fused_adam=True):
        self.model, self.H, self.W, self.K = model, H, W, K
        seen, params = set(), []
        for p in model.parameters():
            if id(p) not in seen:
                seen.add(id(p))
                params.append(p)
        noise = list(model.mlp_noise_coarse.parameters())
        noise_ids = {id(p) for p in noise}
        base = [p for p in params if id(p) not in noise_ids]
        # two parameter groups like the reference (run_lushnerf.py:359-371)
        groups = [{"params": base}, {"params": noise, "lr": lrate}]
        if fused_adam and params[0].is_cuda:
            from .optim import Adam          # one launch for all tensors; lr and step counters on the device
            self.optimizer = Adam(groups, lr=lrate)
        else:
            self.optimizer = torch.optim.Adam(groups, lr=lrate)
        self.allreduce = FlatGradAllReduce(params, group)
        self.render_kwargs = render_kwargs or {}
        self.lrate, self.global_step = lrate, 0
        self._graph, self._static, self._static_loss = None, None, None

    def _step_body(self, batch, force_naive=False, allkernel=True, chunk=1024 * 32):
        m = self.model
        out = m(self.H, self.W, self.K, chunk=chunk, rays=batch["rays"],
                rays_info={"images_idx": batch["images_idx"]}, retraw=True, force_naive=force_naive,
                allkernel=allkernel, kernel_pixel=batch["fq_mask"], **self.render_kwargs)
        loss = ops.blur_loss(out[0], out[1], batch["rgbsf"])      # value + gradient, one launch
        self.optimizer.zero_grad(set_to_none=True)
        loss.backward()
        self.allreduce()
        self.optimizer.step()
        return loss

    def step(self, batch, force_naive=False, allkernel=True, chunk=1024 * 32):
        self.model.train()
        if self._graph is not None and not force_naive and allkernel:
            self.optimizer._sync_lr()               # learning-rate edits reach the replay through device memory
            for k, v in self._static.items():       # H2D (or D2D) into the captured input buffers
                v.copy_(batch[k], non_blocking=True)
            self._graph.replay()
            self.global_step += 1
            return self._static_loss
        loss = self._step_body(batch, force_naive, allkernel, chunk)
        self.global_step += 1
        return loss

    def enable_cuda_graph(self, sample_batch, warmup=3):
        """Capture one whole training step (forward, loss, backward, gradient all-reduce, Adam) in a CUDA graph:
        ~240 launches become one graph launch.  The batch is copied into static buffers before every replay;
        random draws (stratified jitter, sigma noise) come from the graph-registered default CUDA generator.
        The NaN/Inf guard of the model runs in deferred mode (flag word inspected by `check_numerics()`)."""
        import os
        if dist.is_initialized() and dist.get_world_size(self.allreduce.group) > 1:
            # capturing the NCCL all-reduce together with the step hung on 2 x B200 (round 1, unresolved)
            raise RuntimeError("Trainer.enable_cuda_graph supports world size 1 only")
        if self._graph is not None:
            return True
        if not hasattr(self.optimizer, "_sync_lr"):
            raise RuntimeError("Trainer.enable_cuda_graph needs the device-side optimizer (fused_adam=True)")
        dev = sample_batch["rays"].device
        os.environ["LUSH_NUMERIC_GUARD"] = "off"    # no event queries / host copies inside the captured region
        self._static = {k: v.to(dev).clone() for k, v in sample_batch.items()}
        self.model.train()
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(warmup):
                self._step_body(self._static)
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            self._static_loss = self._step_body(self._static)
        self._graph = g
        self.global_step += warmup
        return True

    def check_numerics(self):
        """Deferred NaN/Inf guard for graph mode: one device->host read of the model's flag word."""
        f = getattr(self.model, "_flags", None)
        if f is not None:
            bits = int(f.item())
            f.zero_()
            if bits:
                self.model._print_flags(bits)
            return bits
        return 0

    def set_lr(self, decay_steps=250000, decay_rate=0.1):
        lr = self.lrate * (decay_rate ** (self.global_step / decay_steps))  # reference :681-685
        for g in self.optimizer.param_groups:
            g["lr"] = lr
