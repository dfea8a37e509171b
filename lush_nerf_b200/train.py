"""Ray-sharded training step: one process per GPU, every rank renders its own slice of the global ray
batch, the flat fp32 gradient arenas are all-reduced once per step over NCCL (NVLink 5 / NVSwitch), identical Adam
on every rank.  This is the B200 replacement of the reference's nn.DataParallel wrap
(run_lushnerf.py:348) plus its loss / optimiser lines (:652-685) and its batch staging (:538-616); SURVEY.md 8e, 8f.

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
    """Generic form: packs the gradients of `params` into one flat fp32 buffer, all-reduces it once, unpacks.
    Used for parameters whose gradients do not live in one of the kernels' flat arenas (and by the CPU / gloo tests);
    the product path is ArenaGradAllReduce below, which needs no packing at all."""

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


class ArenaGradAllReduce:
    """All-reduce (average) of the kernels' flat gradient arenas, IN PLACE and from inside the backward pass.

    Every network's backward (lush_mlp_bwd, lush_rbk_motion_bwd) writes the gradients of ALL its parameters into one
    flat fp32 buffer; the nn.Parameters' .grad are views of it.  So there is nothing to pack: as soon as the kernels
    that fill an arena are enqueued (ops.GRAD_ARENA_HOOK), its all-reduce is issued asynchronously -- NCCL's stream
    waits for that point of the compute stream and reduces the fine network's 2.4 MB while the coarse network's backward
    is still running; `finish()` makes the compute stream wait for the collectives before Adam reads the gradients.
    ReduceOp.AVG divides inside the collective.  Parameters whose gradient is not inside an arena (none on the product
    path) go through the generic pack / reduce / unpack."""

    def __init__(self, params, group=None):
        self.params = [p for p in params if p.requires_grad]
        self.group = group
        self._works, self._ranges = [], []
        self.defer = False        # True: only record the arenas inside backward, reduce them in finish() (split graphs)
        self.arenas = []

    def active(self):
        return dist.is_initialized() and dist.get_world_size(self.group) > 1

    def _on_arena(self, g_flat):
        self.arenas.append(g_flat)                         # keeps the arena alive until finish()
        self._ranges.append((g_flat.data_ptr(), g_flat.data_ptr() + g_flat.numel() * 4))
        if not self.defer:
            self._works.append(dist.all_reduce(g_flat, op=dist.ReduceOp.AVG, group=self.group, async_op=True))

    def __enter__(self):
        self._works, self._ranges, self.arenas = [], [], []
        if self.active():
            ops.GRAD_ARENA_HOOK = self._on_arena
        return self

    def __exit__(self, *exc):
        ops.GRAD_ARENA_HOOK = None

    def finish(self):
        """After backward: wait for the arena collectives; reduce whatever was not covered by one."""
        if not self.active():
            return
        if self.defer:
            self.reduce_arenas()
        for work in self._works:
            work.wait()                                    # stream-level wait: the compute stream waits for NCCL's
        rest = []
        for p in self.params:
            if p.grad is None:
                continue
            a = p.grad.data_ptr()
            if not any(lo <= a < hi for lo, hi in self._ranges):
                rest.append(p)
        if rest:
            generic = FlatGradAllReduce(rest, self.group)
            generic()
        self._works = []

    def reduce_arenas(self):
        """Blocking-on-stream all-reduce of the recorded arenas (the eager collectives between two split graphs)."""
        for a in self.arenas:
            dist.all_reduce(a, op=dist.ReduceOp.AVG, group=self.group)


class Trainer:
    """model: NeRFAll (training mode).  step() = [batch staging] + forward + loss + backward + gradient all-reduce +
    Adam.  A batch is either the reference's form {rays [N,3,2], rgbsf [N,3], fq_mask [N], images_idx [N,1]}
    (run_lushnerf.py:606-616) or the index form {images_idx [N,1], pix_idx [N]} when `set_scene` has been given the
    device-resident poses / images / frequency masks: rays, target colours and mask bits are then produced by ONE kernel
    from 16 bytes per pixel (ops.ray_gen_gather) instead of being sliced out of a pre-expanded ray table."""

    def __init__(self, model, H, W, K, lrate=5e-4, render_kwargs=None, group=None, fused_adam=True, guard_every=64):
        self.model, self.H, self.W, self.K = model, H, W, K
        self.guard_every = int(guard_every)   # replays between two (non-blocking) looks at the NaN/Inf flag word
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
        self.allreduce = ArenaGradAllReduce(params, group)
        self.render_kwargs = render_kwargs or {}
        self.lrate, self.global_step = lrate, 0
        self._graph, self._graph_tail, self._static, self._static_loss = None, None, None, None
        self._scene = None
        self.graph_mode = None                # "single" | "split" once a graph has been captured

    # ---- device-resident scene tables for index batches (reference run_lushnerf.py:538-596 keeps the EXPANDED table)
    def set_scene(self, poses, images=None, fq_masks=None):
        """poses [V,3,4] / [V,4,4] / [V,3,5] (indexed by images_idx), images [V,H,W,3] fp32, fq_masks [V,H,W] bool/uint8."""
        dev = next(self.model.parameters()).device
        self._scene = dict(poses=poses.to(dev).float().contiguous(),
                           images=None if images is None else images.to(dev).float().contiguous(),
                           masks=None if fq_masks is None else fq_masks.to(dev).contiguous())

    def _stage(self, batch):
        if "rays" in batch:
            return batch
        if self._scene is None:
            raise RuntimeError("index batches ({images_idx, pix_idx}) need Trainer.set_scene(poses, images, fq_masks)")
        sc = self._scene
        flags = self.model._guard_flags(sc["poses"].device)          # out-of-range indices: clamped + flagged
        out = ops.ray_gen_gather(sc["poses"], batch["images_idx"], batch["pix_idx"], self.H, self.W, self.K,
                                 images=sc["images"], masks=sc["masks"], err_flags=flags)
        rays, rgb, mask = out if isinstance(out, tuple) else (out, None, None)
        return dict(rays=rays, images_idx=batch["images_idx"], rgbsf=rgb if rgb is not None else batch["rgbsf"],
                    fq_mask=mask if mask is not None else batch["fq_mask"])

    def _fwd_bwd(self, batch, force_naive=False, allkernel=True, chunk=1024 * 32):
        m = self.model
        batch = self._stage(batch)
        out = m(self.H, self.W, self.K, chunk=chunk, rays=batch["rays"],
                rays_info={"images_idx": batch["images_idx"]}, retraw=True, force_naive=force_naive,
                allkernel=allkernel, kernel_pixel=batch["fq_mask"], **self.render_kwargs)
        loss = ops.blur_loss(out[0], out[1], batch["rgbsf"])      # value + gradient, one launch
        self.optimizer.zero_grad(set_to_none=True)
        with self.allreduce:                                      # arena all-reduces start from inside backward
            loss.backward()
        return loss

    def _finish(self):
        self.allreduce.finish()
        self.optimizer.step()

    def _step_body(self, batch, force_naive=False, allkernel=True, chunk=1024 * 32):
        loss = self._fwd_bwd(batch, force_naive, allkernel, chunk)
        self._finish()
        return loss

    def step(self, batch, force_naive=False, allkernel=True, chunk=1024 * 32):
        self.model.train()
        if self._graph is not None and not force_naive and allkernel:
            self.optimizer._sync_lr()               # learning-rate edits reach the replay through device memory
            for k, v in self._static.items():       # H2D (or D2D) into the captured input buffers
                v.copy_(batch[k], non_blocking=True)
            self._graph.replay()
            if self._graph_tail is not None:        # split mode: eager NCCL between the two graphs
                for a in self._split_arenas:
                    dist.all_reduce(a, op=dist.ReduceOp.AVG, group=self.allreduce.group)
                self._graph_tail.replay()
            self.global_step += 1
            if self.guard_every > 0 and self.global_step % self.guard_every == 0:
                # the replayed kernels keep OR-ing into the model's flag word; read it the lazy way (pinned copy +
                # event query, no host sync) so that a captured step reports numerical errors like an eager one
                self.model._guard_report(next(iter(self._static.values())).device)
            return self._static_loss
        loss = self._step_body(batch, force_naive, allkernel, chunk)
        self.global_step += 1
        return loss

    def enable_cuda_graph(self, sample_batch, warmup=3, mode="single"):
        """mode="single" (default): capture one whole training step -- batch staging, forward, loss, backward, the NCCL all-reduces of the gradient
        arenas (world size > 1) and Adam -- in ONE CUDA graph: ~70 kernel launches become one graph launch, and the
        collectives of the fine network overlap the coarse network's backward as parallel branches of the graph.
        The capture runs in `thread_local` error mode: NCCL's watchdog thread polls events while we capture, which the
        default (`global`) mode turns into a capture failure on one rank -- and a hang on its peers (round 1's symptom).
        The batch is copied into static buffers before every replay; random draws (stratified jitter, sigma noise) come
        from the graph-registered default CUDA generator.  The NaN/Inf guard of the model does nothing while the stream
        is capturing; `step()` looks at the flag word every `guard_every` replays (`check_numerics()` reads it
        synchronously).
        mode="split" (world size > 1 only): TWO graphs around eager collectives -- A: staging, forward, loss, backward;
        then one eager in-place all-reduce per gradient arena; B: Adam (same memory pool).  No overlap of the collectives
        with the backward, but nothing of NCCL inside a capture."""
        if self._graph is not None:
            return True
        if mode not in ("single", "split"):
            raise ValueError(f"enable_cuda_graph: mode must be 'single' or 'split', got {mode!r}")
        if not hasattr(self.optimizer, "_sync_lr"):
            raise RuntimeError("Trainer.enable_cuda_graph needs the device-side optimizer (fused_adam=True)")
        dev = next(self.model.parameters()).device
        self._static = {k: v.to(dev).clone() for k, v in sample_batch.items()}
        self.model.train()
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(warmup):                 # every rank runs the same number of (collective) warm-up steps
                self._step_body(self._static)
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        g = torch.cuda.CUDAGraph()
        split = mode == "split" and self.allreduce.active()
        if not split:
            with torch.cuda.graph(g, capture_error_mode="thread_local"):
                self._static_loss = self._step_body(self._static)
        else:
            self.allreduce.defer = True
            with torch.cuda.graph(g):
                self._static_loss = self._fwd_bwd(self._static)      # the hook only records the arenas
            self._split_arenas = list(self.allreduce.arenas)
            tail = torch.cuda.CUDAGraph()
            with torch.cuda.graph(tail, pool=g.pool()):   # same memory pool: reads graph A's gradient arenas
                self.optimizer.step()
            self._graph_tail = tail
        self._graph = g
        self.graph_mode = "split" if split else "single"
        self.global_step += warmup
        return True

    def check_numerics(self):
        """Deferred NaN/Inf guard for graph mode: one device->host read of the model's flag word."""
        f = getattr(self.model, "_flags", None)
        if f is None:
            return 0
        bits = int(f.item())
        f.zero_()
        pend = getattr(self.model, "_flags_pending", None)     # a copy the lazy guard took but has not looked at yet
        if pend is not None:
            pend[1].synchronize()
            bits |= int(pend[0].item())
            self.model._flags_pending = None
        if bits:
            self.model._print_flags(bits, self.model._overflow_hint())
        return bits

    def set_lr(self, decay_steps=250000, decay_rate=0.1):
        lr = self.lrate * (decay_rate ** (self.global_step / decay_steps))  # reference :681-685
        for g in self.optimizer.param_groups:
            g["lr"] = lr
