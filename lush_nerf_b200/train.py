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
    """Packs the gradients of `params` into one flat fp32 buffer, all-reduces it once, unpacks.
    The three phases are separate methods so that a captured training step can keep the NCCL call outside its
    CUDA graphs (pack at the end of graph A, all-reduce eagerly, unpack at the start of graph B)."""

    def __init__(self, params, group=None):
        self.params = [p for p in params if p.requires_grad]
        self.group = group
        self.numel = sum(p.numel() for p in self.params)
        self.flat = None

    def active(self):
        return dist.is_initialized() and dist.get_world_size(self.group) > 1

    def pack(self):
        if not self.active():
            return
        # parameters without a gradient (NeRF_Noise.alpha_linear) contribute zeros and receive nothing back
        pieces = [p.grad.reshape(-1) if p.grad is not None else p.new_zeros(p.numel()) for p in self.params]
        self.flat = torch.cat(pieces)                     # one multi-tensor pack kernel instead of one copy per tensor

    def reduce(self):
        if self.active():
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=self.group)

    def unpack(self):
        if not self.active():
            return
        self.flat.div_(dist.get_world_size(self.group))
        with_grad = [p for p in self.params if p.grad is not None]
        views, off = [], 0
        for p in self.params:
            n = p.numel()
            if p.grad is not None:
                views.append(self.flat[off:off + n].view_as(p.grad))
            off += n
        torch._foreach_copy_([p.grad for p in with_grad], views)   # one multi-tensor unpack kernel

    def __call__(self):
        self.pack()
        self.reduce()
        self.unpack()


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
        self._graph, self._graph_tail, self._static, self._static_loss = None, None, None, None

    def _fwd_bwd(self, batch, force_naive=False, allkernel=True, chunk=1024 * 32):
        m = self.model
        out = m(self.H, self.W, self.K, chunk=chunk, rays=batch["rays"],
                rays_info={"images_idx": batch["images_idx"]}, retraw=True, force_naive=force_naive,
                allkernel=allkernel, kernel_pixel=batch["fq_mask"], **self.render_kwargs)
        loss = ops.blur_loss(out[0], out[1], batch["rgbsf"])      # value + gradient, one launch
        self.optimizer.zero_grad(set_to_none=True)
        loss.backward()
        self.allreduce.pack()
        return loss

    def _finish(self):
        self.allreduce.unpack()
        self.optimizer.step()

    def _step_body(self, batch, force_naive=False, allkernel=True, chunk=1024 * 32):
        loss = self._fwd_bwd(batch, force_naive, allkernel, chunk)
        self.allreduce.reduce()
        self._finish()
        return loss

    def step(self, batch, force_naive=False, allkernel=True, chunk=1024 * 32):
        self.model.train()
        if self._graph is not None and not force_naive and allkernel:
            self.optimizer._sync_lr()               # learning-rate edits reach the replay through device memory
            for k, v in self._static.items():       # H2D (or D2D) into the captured input buffers
                v.copy_(batch[k], non_blocking=True)
            self._graph.replay()
            if self._graph_tail is not None:        # multi-GPU: the NCCL all-reduce runs between the two graphs
                self.allreduce.reduce()
                self._graph_tail.replay()
            self.global_step += 1
            return self._static_loss
        loss = self._step_body(batch, force_naive, allkernel, chunk)
        self.global_step += 1
        return loss

    def enable_cuda_graph(self, sample_batch, warmup=3):
        """Capture one whole training step in CUDA graphs: ~72 kernel launches become one graph launch.
        World size 1: ONE graph (forward, loss, backward, Adam).  World size > 1: TWO graphs around an eager NCCL
        all-reduce of the flat gradient (A: forward, loss, backward, pack; B: unpack, Adam) -- capturing the
        collective itself hung on 2 x B200 in round 1, and NCCL's launch is one call either way.
        The batch is copied into static buffers before every replay; random draws (stratified jitter, sigma noise)
        come from the graph-registered default CUDA generator.  The NaN/Inf guard of the model runs in deferred
        mode (flag word inspected by `check_numerics()`)."""
        import os
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
            for _ in range(warmup):                 # every rank runs the same number of (collective) warm-up steps
                self._step_body(self._static)
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        split = self.allreduce.active()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            if split:
                self._static_loss = self._fwd_bwd(self._static)
            else:
                self._static_loss = self._step_body(self._static)
        tail = None
        if split:
            tail = torch.cuda.CUDAGraph()
            with torch.cuda.graph(tail, pool=g.pool()):   # same memory pool: reads graph A's gradients and flat buffer
                self._finish()
        self._graph, self._graph_tail = g, tail
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
