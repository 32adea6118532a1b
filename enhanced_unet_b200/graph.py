"""One full training step as ONE CUDA graph launch.

The step of the reference's training loop (train_eval.py:236-353: forward, combined loss, backward, clip + AdamW) is a
fixed schedule of ~190 kernel launches whose shapes never change; issuing them one by one leaves the GPU idle between
dependent launches and keeps a host core busy.  ``GraphedTrainStep`` captures the schedule once (``torch.cuda.graph``:
every C-ABI launch goes to the capturing stream, device memory comes from the graph's private pool, so the TMA
descriptors encoded at capture time stay valid) and replays it per step.  Everything that changes between steps lives
in device memory: the input batch (static buffers), the AdamW step counter and learning rate
(``ClippedAdamW.enable_device_state``), BatchNorm's ``num_batches_tracked`` (incremented by the finalize kernel).

Data parallel (world size > 1): the bucketed NCCL all-reduce launched from inside backward is not captured; those runs
use the eager step (``Trainer.train_step``) - the choice is explicit, never a silent fallback.
"""
from __future__ import annotations

from typing import Optional

import torch

from .ops import combined_loss
from .optim import ClippedAdamW


class GraphedTrainStep:
    def __init__(self, model: torch.nn.Module, optimizer: ClippedAdamW, batch: int, height: int, width: int,
                 warmup_steps: int = 3, example: Optional[tuple] = None):
        p0 = next(model.parameters())
        if not p0.is_cuda:
            raise RuntimeError("GraphedTrainStep needs the model on a CUDA device (no CPU fallback)")
        if getattr(model, "grad_sink", None) is not None and torch.distributed.is_initialized() and torch.distributed.get_world_size() > 1:
            raise RuntimeError("GraphedTrainStep captures a single-GPU step; data-parallel runs use the eager step")
        if height % 32 or width % 32:
            raise RuntimeError("GraphedTrainStep takes images whose sides are multiples of 32 (pad on the host as train_eval._pad32 does)")
        self.model, self.opt = model, optimizer
        self.dev = p0.device
        self.params = [p for p in model.parameters()]
        self.x = torch.zeros(batch, 3, height, width, device=self.dev, dtype=torch.float32)
        self.t = torch.zeros(batch, height, width, device=self.dev, dtype=torch.int64)
        if example is not None:                      # warm-up on real data keeps BatchNorm statistics sensible
            self.x.copy_(example[0])
            self.t.copy_(example[1])
        optimizer.enable_device_state()
        model.train()
        # warm-up on a side stream (allocator pools, lazily created optimiser state, kernel attributes), then capture
        side = torch.cuda.Stream(self.dev)
        side.wait_stream(torch.cuda.current_stream(self.dev))
        with torch.cuda.stream(side):
            for _ in range(max(1, warmup_steps)):
                self._eager()
        torch.cuda.current_stream(self.dev).wait_stream(side)
        self.warmup_steps = max(1, warmup_steps)
        from . import lib
        before = sum(lib.COUNTERS.values())
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.loss = self._eager()
        self.launches_per_step = sum(lib.COUNTERS.values()) - before      # C-ABI launches captured in one step
        # the capture itself does not execute: python-side step counters advanced by one step that never ran
        optimizer.sync_steps(-1)

    def _eager(self) -> torch.Tensor:
        for p in self.params:
            p.grad = None
        loss = combined_loss(self.model(self.x), self.t)
        loss.backward()
        self.opt.step()
        return loss.detach()

    def __call__(self, images: torch.Tensor, masks: torch.Tensor) -> torch.Tensor:
        """One optimisation step on ``images`` [B,3,H,W] fp32 / ``masks`` [B,H,W] int64 (host or device tensors; copied into
        the graph's static buffers on the current stream).  Returns the (device) loss tensor of this step - it is
        overwritten by the next call."""
        if images.shape != self.x.shape or masks.shape != self.t.shape:
            raise RuntimeError(f"GraphedTrainStep was captured for {tuple(self.x.shape)} / {tuple(self.t.shape)}, got "
                               f"{tuple(images.shape)} / {tuple(masks.shape)}")
        if images.data_ptr() != self.x.data_ptr():
            self.x.copy_(images, non_blocking=True)
        if masks.data_ptr() != self.t.data_ptr():
            self.t.copy_(masks, non_blocking=True)
        self.opt.push_lr()
        self.graph.replay()
        self.opt.sync_steps(1)
        packs = getattr(self.model, "_packs", None)
        if packs is not None:
            packs.invalidate()       # the replay re-packed and then updated the weights: any later eager pass must re-pack
        sat = getattr(self.model, "_sat", None)
        if sat is not None:
            sat.check(block=False)
            sat.publish()
        return self.loss
