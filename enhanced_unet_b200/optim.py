"""Optimiser step of the reference Trainer (train_eval.py:120 ``AdamW(lr, weight_decay=1e-4, betas=(.9,.999))``,
341 ``clip_grad_norm_(max_norm=1.0)``, 343 ``optimizer.step()``) through the C-ABI kernels: one fp64 sum of
squares over all gradients, then a fused clip + AdamW update per parameter tensor."""
from __future__ import annotations

from typing import Callable, Iterable, Optional

import torch

from .lib import call


class ClippedAdamW:
    def __init__(self, params: Iterable[torch.nn.Parameter], lr: float = 4e-3, weight_decay: float = 1e-4,
                 betas=(0.9, 0.999), eps: float = 1e-8, max_norm: float = 1.0, on_update: Optional[Callable[[], None]] = None):
        self.params = [p for p in params]
        if not self.params or not all(p.is_cuda for p in self.params):
            raise RuntimeError("ClippedAdamW needs CUDA parameters (no CPU fallback)")
        self.m = [torch.zeros_like(p) for p in self.params]
        self.v = [torch.zeros_like(p) for p in self.params]
        self.lr, self.wd, self.betas, self.eps, self.max_norm = lr, weight_decay, betas, eps, max_norm
        self.t = 0
        self.sq = torch.zeros((), dtype=torch.float64, device=self.params[0].device)
        self.on_update = on_update

    def zero_grad(self) -> None:
        for p in self.params:
            p.grad = None

    @torch.no_grad()
    def step(self, grad_scale: float = 1.0) -> None:
        """``grad_scale`` multiplies every gradient before clipping (1/world_size after a sum all-reduce)."""
        self.t += 1
        self.sq.zero_()
        for p in self.params:
            if p.grad is None:
                raise RuntimeError("ClippedAdamW.step: parameter without gradient")
            call("eunet_sumsq", p.grad.data_ptr(), p.numel(), self.sq.data_ptr())
        for p, m, v in zip(self.params, self.m, self.v):
            call("eunet_adamw_step", p.data_ptr(), p.grad.data_ptr(), m.data_ptr(), v.data_ptr(), p.numel(), self.sq.data_ptr(),
                 self.max_norm, self.lr, self.betas[0], self.betas[1], self.eps, self.wd, self.t, grad_scale)
        if self.on_update is not None:
            self.on_update()   # parameters changed behind autograd's back: drop packed-filter caches

    def grad_norm(self) -> float:
        return float(self.sq.sqrt())
