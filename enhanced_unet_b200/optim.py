"""Optimiser step of the reference Trainer (train_eval.py:120 ``AdamW(lr, weight_decay=1e-4, betas=(.9,.999))``,
341 ``clip_grad_norm_(max_norm=1.0)``, 343 ``optimizer.step()``) through the C-ABI kernels: one fp64 sum of
squares over all gradients, then a fused clip + AdamW update per parameter tensor.

``ClippedAdamW`` is a ``torch.optim.Optimizer`` (param_groups / state / state_dict in the torch AdamW
layout), so the reference's LR schedulers (LinearLR warm-up, CosineAnnealingWarmRestarts,
train_eval.py:122-132) and checkpoint code (train_eval.py:1143-1151) work on it unchanged.
"""
from __future__ import annotations

from typing import Callable, Iterable, Optional

import torch

from .lib import call


class ClippedAdamW(torch.optim.Optimizer):
    def __init__(self, params: Iterable[torch.nn.Parameter], lr: float = 4e-3, weight_decay: float = 1e-4,
                 betas=(0.9, 0.999), eps: float = 1e-8, max_norm: float = 1.0, on_update: Optional[Callable[[], None]] = None):
        defaults = dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, max_norm=max_norm)
        super().__init__(params, defaults)
        self.on_update = on_update
        self._sq: Optional[torch.Tensor] = None

    def _all_params(self):
        for g in self.param_groups:
            for p in g["params"]:
                if p.grad is not None:
                    yield g, p

    @torch.no_grad()
    def step(self, closure=None, grad_scale: float = 1.0):
        """``grad_scale`` multiplies every gradient before clipping (1/world_size after a sum all-reduce).
        The clip uses the global L2 norm over ALL parameters with gradients, like clip_grad_norm_."""
        if closure is not None:
            raise RuntimeError("ClippedAdamW does not support closures")
        items = list(self._all_params())
        if not items:
            return None
        dev = items[0][1].device
        if not all(p.is_cuda and p.grad.is_cuda for _, p in items):
            raise RuntimeError("ClippedAdamW needs CUDA parameters and gradients (no CPU fallback)")
        if self._sq is None or self._sq.device != dev:
            self._sq = torch.zeros((), dtype=torch.float64, device=dev)
        self._sq.zero_()
        for _, p in items:
            g = p.grad if p.grad.is_contiguous() else p.grad.contiguous()
            call("eunet_sumsq", g.data_ptr(), g.numel(), self._sq.data_ptr())
        for grp, p in items:
            st = self.state[p]
            if len(st) == 0:
                st["step"] = 0
                st["exp_avg"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
            st["step"] = int(st["step"]) + 1
            g = p.grad if p.grad.is_contiguous() else p.grad.contiguous()
            b1, b2 = grp["betas"]
            call("eunet_adamw_step", p.data_ptr(), g.data_ptr(), st["exp_avg"].data_ptr(), st["exp_avg_sq"].data_ptr(),
                 p.numel(), self._sq.data_ptr(), float(grp["max_norm"]), float(grp["lr"]), float(b1), float(b2),
                 float(grp["eps"]), float(grp["weight_decay"]), st["step"], float(grad_scale))
        if self.on_update is not None:
            self.on_update()   # parameters changed behind autograd's back: drop packed-filter caches
        return None

    def grad_norm(self) -> float:
        """Global gradient norm seen by the last ``step`` (device -> host sync)."""
        return float(self._sq.sqrt()) if self._sq is not None else 0.0
