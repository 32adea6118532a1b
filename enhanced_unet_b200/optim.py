"""Optimiser step of the reference Trainer (train_eval.py:120 ``AdamW(lr, weight_decay=1e-4, betas=(.9,.999))``,
341 ``clip_grad_norm_(max_norm=1.0)``, 343 ``optimizer.step()``) through the C-ABI kernels: one fp64 sum of
squares over all gradients, then a fused clip + AdamW update per parameter tensor.

``ClippedAdamW`` is a ``torch.optim.Optimizer`` (param_groups / state / state_dict in the torch AdamW
layout), so the reference's LR schedulers (LinearLR warm-up, CosineAnnealingWarmRestarts,
train_eval.py:122-132) and checkpoint code (train_eval.py:1143-1151) work on it unchanged.
"""
from __future__ import annotations

import ctypes
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
        # device-resident step counter / learning rate (``graph.GraphedTrainStep``): a captured step must advance on replay
        self._dev_state: Optional[dict] = None

    def enable_device_state(self) -> dict:
        """Keep the step counter and the learning rate in device memory (needed when ``step`` is captured into a CUDA
        graph).  Requires one parameter group and a common step count.  The python-side ``state[p]['step']`` keeps being
        advanced by whoever replays the graph (``sync_steps``) so that ``state_dict()`` stays in the torch AdamW layout."""
        if len(self.param_groups) != 1:
            raise RuntimeError("device-resident optimiser state needs exactly one parameter group")
        ps = [p for p in self.param_groups[0]["params"]]
        steps = {int(self.state[p]["step"]) for p in ps if len(self.state[p])}
        if len(steps) > 1:
            raise RuntimeError("device-resident optimiser state needs a common step count")
        dev = ps[0].device
        self._dev_state = {"step": torch.tensor([steps.pop() if steps else 0], dtype=torch.int32, device=dev),
                           "lr": torch.tensor([float(self.param_groups[0]["lr"])], dtype=torch.float32, device=dev),
                           "hyper": torch.zeros(4, dtype=torch.float32, device=dev), "lr_host": float(self.param_groups[0]["lr"])}
        return self._dev_state

    def push_lr(self) -> None:
        """Copy a learning rate changed by a scheduler into the device-resident value (no-op when unchanged)."""
        ds = self._dev_state
        if ds is not None and float(self.param_groups[0]["lr"]) != ds["lr_host"]:
            ds["lr_host"] = float(self.param_groups[0]["lr"])
            ds["lr"].fill_(ds["lr_host"])

    def sync_steps(self, n: int = 1) -> None:
        """Advance the python-side step counters by ``n`` replays of a captured step."""
        for g in self.param_groups:
            for p in g["params"]:
                if len(self.state[p]):
                    self.state[p]["step"] = int(self.state[p]["step"]) + n

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
        grads = [p.grad if p.grad.is_contiguous() else p.grad.contiguous() for _, p in items]
        n = len(items)
        VP, LL = ctypes.c_void_p * n, ctypes.c_longlong * n
        g_arr = VP(*[g.data_ptr() for g in grads])
        n_arr = LL(*[g.numel() for g in grads])
        call("eunet_sumsq_multi", g_arr, n_arr, n, self._sq.data_ptr())        # global norm over ALL gradients
        # one fused clip+AdamW launch per run of parameters that share hyper-parameters and step count
        runs = []
        for i, (grp, p) in enumerate(items):
            st = self.state[p]
            if len(st) == 0:
                st["step"] = 0
                st["exp_avg"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
            st["step"] = int(st["step"]) + 1
            key = (float(grp["max_norm"]), float(grp["lr"]), float(grp["betas"][0]), float(grp["betas"][1]), float(grp["eps"]),
                   float(grp["weight_decay"]), st["step"])
            if runs and runs[-1][0] == key:
                runs[-1][1].append(i)
            else:
                runs.append((key, [i]))
        hyper = None
        if self._dev_state is not None:
            if len(runs) != 1:
                raise RuntimeError("device-resident optimiser state needs all parameters in one run (same hyper-parameters / step)")
            ds = self._dev_state
            call("eunet_adamw_prepare", ds["step"].data_ptr(), ds["lr"].data_ptr(), runs[0][0][2], runs[0][0][3], ds["hyper"].data_ptr())
            hyper = ds["hyper"].data_ptr()
        for key, idxs in runs:
            k = len(idxs)
            VPk, LLk = ctypes.c_void_p * k, ctypes.c_longlong * k
            ps = [items[i][1] for i in idxs]
            call("eunet_adamw_multi", VPk(*[p.data_ptr() for p in ps]), VPk(*[grads[i].data_ptr() for i in idxs]),
                 VPk(*[self.state[p]["exp_avg"].data_ptr() for p in ps]), VPk(*[self.state[p]["exp_avg_sq"].data_ptr() for p in ps]),
                 LLk(*[p.numel() for p in ps]), k, self._sq.data_ptr(), key[0], key[1], key[2], key[3], key[4], key[5], key[6],
                 float(grad_scale), hyper)
        # the kernels wrote the parameters behind autograd's back: bump their version counters, which is what the
        # packed-filter caches (engine.PackCache) key on - the optimiser is safe on its own, ``on_update`` is an optimisation
        for _, p in items:
            torch.autograd.graph.increment_version(p)
        if self.on_update is not None:
            self.on_update()
        return None

    def grad_norm(self) -> float:
        """Global gradient norm seen by the last ``step`` (device -> host sync)."""
        return float(self._sq.sqrt()) if self._sq is not None else 0.0
