"""Adam for the training step, one kernel launch for all parameter tensors (csrc/optim.cu).

Same update rule, constructor arguments and state_dict layout (`step`, `exp_avg`, `exp_avg_sq` per parameter,
`step` a 0-dim fp32 device tensor as in torch's capturable mode) as the reference's
torch.optim.Adam(params=grad_vars, lr=lrate, betas=(0.9, 0.999)) (run_lushnerf.py:359-371), so optimizer
checkpoints ('optimizer_state_dict', run_lushnerf.py:690-700) move both ways.  Differences: CUDA parameters only,
no weight decay / amsgrad / maximize (the reference uses none of them); learning rates and step counters live
on the device, so a captured CUDA graph of the step follows `param_group['lr']` edits made between replays.
"""
from __future__ import annotations

import ctypes

import torch

from ._lib import AdamSeg, check, lib, stream


class Adam(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8):
        if lr < 0 or eps < 0 or not (0 <= betas[0] < 1 and 0 <= betas[1] < 1):
            raise ValueError(f"invalid Adam hyper-parameters lr={lr} betas={betas} eps={eps}")
        super().__init__(params, dict(lr=lr, betas=tuple(betas), eps=eps))
        flat = [p for g in self.param_groups for p in g["params"]]
        if not flat or any((not p.is_cuda) or p.dtype != torch.float32 for p in flat):
            raise RuntimeError("lush_nerf_b200.optim.Adam needs fp32 CUDA parameters; there is no CPU fallback")
        b = {(g["betas"], g["eps"]) for g in self.param_groups}
        if len(b) != 1:
            raise ValueError("all parameter groups must share betas and eps")
        self._dev = flat[0].device
        self._slot = {id(p): i for i, p in enumerate(flat)}
        self._steps = torch.zeros(len(flat), dtype=torch.float32, device=self._dev)
        self._lr = torch.zeros(len(self.param_groups), dtype=torch.float32, device=self._dev)
        self._lr_host = [None] * len(self.param_groups)
        self._counter = torch.zeros(1, dtype=torch.int32, device=self._dev)
        self._table, self._table_key = None, None

    def add_param_group(self, param_group):
        if hasattr(self, "_steps"):      # the device-side step / learning-rate tables are sized at construction
            raise NotImplementedError("lush_nerf_b200.optim.Adam: add_param_group after construction is not supported")
        super().add_param_group(param_group)

    def _sync_lr(self):
        """Push edited learning rates to the device (a small copy on the current stream, outside any graph)."""
        cur = [float(g["lr"]) for g in self.param_groups]
        if cur != self._lr_host:
            if torch.cuda.is_current_stream_capturing():
                if self._lr_host[0] is None:
                    raise RuntimeError("run one optimizer step (or _sync_lr) before capturing a CUDA graph")
                return                    # replays read the device copy; the edit is pushed by the next eager call
            self._lr.copy_(torch.tensor(cur, dtype=torch.float32, device="cpu"), non_blocking=False)
            self._lr_host = cur

    def _state_of(self, p):
        st = self.state[p]
        i = self._slot[id(p)]
        if "exp_avg" not in st:
            st["step"] = self._steps[i]
            st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
            st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
        elif not torch.is_tensor(st["step"]) or st["step"].device != self._dev \
                or st["step"].data_ptr() != self._steps.data_ptr() + 4 * i:   # after load_state_dict: re-home the counter
            self._steps[i] = float(st["step"])
            st["step"] = self._steps[i]
            for k in ("exp_avg", "exp_avg_sq"):
                if st[k].device != p.device or st[k].dtype != torch.float32 or not st[k].is_contiguous():
                    st[k] = st[k].to(p.device, torch.float32).contiguous()
        return st

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        self._sync_lr()
        rows = []
        for gi, group in enumerate(self.param_groups):
            for p in group["params"]:
                if p.grad is None or p.numel() == 0:
                    continue
                g = p.grad
                if g.is_sparse or g.dtype != torch.float32 or not g.is_cuda:
                    raise RuntimeError("lush_nerf_b200.optim.Adam: gradients must be dense fp32 CUDA tensors")
                if not g.is_contiguous():
                    g = p.grad = g.contiguous()
                if not p.is_contiguous():
                    raise RuntimeError("lush_nerf_b200.optim.Adam: parameters must be contiguous")
                st = self._state_of(p)
                rows.append((p.data_ptr(), g.data_ptr(), st["exp_avg"].data_ptr(), st["exp_avg_sq"].data_ptr(),
                             p.numel(), self._slot[id(p)], gi))
        if not rows:
            return loss
        key = tuple(rows)
        if key != self._table_key:
            tab = (AdamSeg * len(rows))()
            for t, r in zip(tab, rows):
                t.param, t.grad, t.exp_avg, t.exp_avg_sq, t.numel, t.step_index, t.lr_index = r
            self._table, self._table_key = tab, key
        g0 = self.param_groups[0]
        check(lib().lush_adam_step(self._table, len(rows), ctypes.c_void_p(self._lr.data_ptr()),
                                   ctypes.c_void_p(self._steps.data_ptr()),
                                   ctypes.c_void_p(self._counter.data_ptr()), g0["betas"][0], g0["betas"][1],
                                   g0["eps"], stream()), "lush_adam_step")
        return loss
