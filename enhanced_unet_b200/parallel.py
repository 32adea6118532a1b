"""Data parallelism for the hot path: one process per GPU, full parameter replica, per-rank
BatchNorm statistics, ONE exchange per training step - the gradient all-reduce - and an integer count
all-reduce for metrics.  ``torch.distributed`` (NCCL over NVLink on GPUs, gloo in the CPU tests) is the
plumbing; the arithmetic on either side is the C-ABI kernels.

The reference has no distributed code at all; semantics follow what
``DistributedDataParallel`` over the reference module would do: gradients averaged over ranks, buffers
(running statistics) rank-local with rank 0 authoritative for checkpoints.
"""
from __future__ import annotations

from typing import Iterable, List, Optional, Sequence

import torch
import torch.distributed as dist


def world_size(group=None) -> int:
    return dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1


def broadcast_parameters(tensors: Iterable[torch.Tensor], src: int = 0, group=None) -> None:
    """Initial parameter (and buffer) broadcast from ``src`` so every replica starts identical."""
    if world_size(group) == 1:
        return
    for t in tensors:
        dist.broadcast(t.data if isinstance(t, torch.nn.Parameter) else t, src, group=group)


class GradientAllReduce:
    """Bucketed SUM all-reduce of ``.grad`` over flat fp32 buffers.

    Buckets follow REVERSE parameter order (the order backward produces gradients: tail, dec2, dec3, dec4,
    enc4 ... enc1) and are capped at ``bucket_bytes`` (payload is 31 MB in total, i.e. latency-bound: a few
    large buckets).  ``reduce()`` launches every bucket asynchronously and ``wait()`` copies the
    sums back into ``.grad``; averaging is folded into the optimiser (``ClippedAdamW.step(grad_scale=1/world)``).
    """

    def __init__(self, params: Sequence[torch.nn.Parameter], bucket_bytes: int = 16 << 20, group=None):
        self.params: List[torch.nn.Parameter] = list(params)
        self.group = group
        self.buckets: List[List[torch.nn.Parameter]] = []
        cur: List[torch.nn.Parameter] = []
        size = 0
        for p in reversed(self.params):
            nbytes = p.numel() * 4
            if cur and size + nbytes > bucket_bytes:
                self.buckets.append(cur)
                cur, size = [], 0
            cur.append(p)
            size += nbytes
        if cur:
            self.buckets.append(cur)
        self._flat: List[Optional[torch.Tensor]] = [None] * len(self.buckets)
        self._work = []

    def reduce(self) -> None:
        if world_size(self.group) == 1:
            return
        self._work = []
        for i, bucket in enumerate(self.buckets):
            n = sum(p.numel() for p in bucket)
            dev = bucket[0].grad.device
            if self._flat[i] is None or self._flat[i].device != dev:
                self._flat[i] = torch.empty(n, dtype=torch.float32, device=dev)
            flat = self._flat[i]
            off = 0
            for p in bucket:
                if p.grad is None:
                    raise RuntimeError("GradientAllReduce: parameter without gradient")
                flat[off:off + p.numel()].copy_(p.grad.reshape(-1))
                off += p.numel()
            self._work.append(dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group, async_op=True))

    def wait(self) -> None:
        if not self._work:
            return
        for w, bucket, flat in zip(self._work, self.buckets, self._flat):
            w.wait()
            off = 0
            for p in bucket:
                p.grad.copy_(flat[off:off + p.numel()].view_as(p.grad))
                off += p.numel()
        self._work = []


def allreduce_counts(counts: torch.Tensor, group=None) -> torch.Tensor:
    """Global confusion matrix: int64 SUM over ranks (associative, bit-exact)."""
    if world_size(group) == 1:
        return counts
    out = counts.clone()
    dist.all_reduce(out, op=dist.ReduceOp.SUM, group=group)
    return out


def shard_batch(n_items: int, rank: int, world: int):
    """Contiguous shard [lo, hi) of ``n_items`` independent images/tiles for ``rank`` (inference configs)."""
    per = (n_items + world - 1) // world
    lo = min(n_items, rank * per)
    return lo, min(n_items, lo + per)
