"""torch.optim.AdamW semantics on this repo's multi-tensor kernel (csrc/adamw.cu, feddat_adamw_step).

What the reference builds in ``TaskTrainer.create_optimizer`` (task_trainer.py:477-504) and steps twice per batch
(:303-308, :323-328).  Same update rule and state as ``torch.optim.AdamW(..., fused=True, capturable=True)``: per-parameter
``exp_avg`` / ``exp_avg_sq`` / ``step`` (device scalars), a device-tensor ``lr`` per group (what LambdaLR schedulers and
``GraphedTrainStep`` write into), parameters without a gradient are skipped.  One launch (+ a step-counter bump) per
``step()`` instead of one 48-70-block launch per weight-decay group.
"""
from __future__ import annotations

import torch

from .. import _lib, ops


class FusedAdamW(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2):
        defaults = dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, capturable=True, fused=True)
        super().__init__(params, defaults)
        self._lists = {}        # (ids of the parameters stepped, their gradient addresses) -> ctypes array

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        active, key = [], []
        betas, eps = None, None
        for group in self.param_groups:
            if betas is None:
                betas, eps = tuple(group["betas"]), float(group["eps"])
            elif betas != tuple(group["betas"]) or eps != float(group["eps"]):
                raise ValueError("FusedAdamW: betas / eps must be the same in every parameter group")
            lr = group["lr"]
            if not (isinstance(lr, torch.Tensor) and lr.is_cuda and lr.dtype == torch.float32):
                raise ValueError("FusedAdamW: each group's lr must be a CUDA fp32 scalar tensor")
            for p in group["params"]:
                if p.grad is None:
                    continue
                if not (p.is_cuda and p.dtype == torch.float32 and p.is_contiguous() and p.grad.dtype == torch.float32
                        and p.grad.is_contiguous()):
                    raise ValueError("FusedAdamW: parameters and gradients must be contiguous CUDA fp32 tensors")
                st = self.state[p]
                if not st:
                    st["step"] = torch.zeros((), device=p.device, dtype=torch.float32)
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                active.append((p, st, lr, float(group["weight_decay"])))
                key.append((id(p), p.grad.data_ptr(), lr.data_ptr()))
        if not active:
            return loss
        key = tuple(key)
        arr = self._lists.get(key)
        if arr is None:
            if len(self._lists) > 64:                  # eager training: gradient addresses move from step to step
                self._lists.clear()
            arr = (_lib.AdamwTensor * len(active))()
            for a, (p, st, lr, wd) in zip(arr, active):
                a.param, a.grad, a.exp_avg, a.exp_avg_sq = p.data_ptr(), p.grad.data_ptr(), st["exp_avg"].data_ptr(), st["exp_avg_sq"].data_ptr()
                a.step, a.lr, a.weight_decay, a.numel = st["step"].data_ptr(), lr.data_ptr(), wd, p.numel()
            self._lists[key] = arr
        ops.adamw_step(arr, len(active), betas[0], betas[1], eps)
        return loss
