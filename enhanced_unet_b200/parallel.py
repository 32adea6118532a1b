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


class FlatGradBuffer:
    """ONE flat fp32 gradient buffer laid out in the order backward FINISHES gradients
    (``engine.grad_production_order()``: tail, dec2, dec3, dec4, enc4 ... enc1), cut into buckets of at most
    ``bucket_bytes`` at parameter boundaries (and after every parameter named in ``close_after``).  It is the ``engine.GradSink`` of the data-parallel path: the backward
    kernels write every gradient straight into its slice (``dst``), and ``ready`` launches the asynchronous SUM
    all-reduce of a bucket the moment its last gradient has been enqueued - NCCL runs it on its own stream behind
    an event, so the exchange overlaps the rest of backward (the level-0/1 encoder blocks, i.e. most of its time).
    ``.grad`` of every parameter is a view of the buffer; ``wait()`` joins the exchange before the optimiser step
    (averaging is folded into ``ClippedAdamW.step(grad_scale=1/world)``).

    Works on any device (gloo on CPU in the tests); holds no CUDA-specific state."""

    def __init__(self, named_shapes: Sequence, device, bucket_bytes: int = 8 << 20, group=None, close_after: Sequence[str] = ()):
        self.group = group
        self.dev = torch.device(device)
        self.names: List[str] = [n for n, _ in named_shapes]
        self.shapes = {n: tuple(sh) for n, sh in named_shapes}
        self.offsets = {}
        off = 0
        for n, sh in named_shapes:
            self.offsets[n] = off
            off += int(torch.Size(sh).numel())
        self.numel = off
        self.flat = torch.zeros(off, dtype=torch.float32, device=self.dev)
        self.grads = {n: self.flat[self.offsets[n]:self.offsets[n] + torch.Size(self.shapes[n]).numel()].view(self.shapes[n])
                      for n in self.names}
        # buckets: contiguous [lo, hi) element ranges closed by the parameter named in ``_closes``
        self.buckets: List[tuple] = []
        self._closes = {}
        lo = 0
        for i, n in enumerate(self.names):
            hi = self.offsets[n] + torch.Size(self.shapes[n]).numel()
            nxt = self.names[i + 1] if i + 1 < len(self.names) else None
            nxt_hi = (self.offsets[nxt] + torch.Size(self.shapes[nxt]).numel()) if nxt else None
            if nxt is None or (nxt_hi - lo) * 4 > bucket_bytes or n in close_after:
                self._closes[n] = len(self.buckets)
                self.buckets.append((lo, hi))
                lo = hi
        self._work: List = []
        self._launched = 0

    # Encoder levels finish ~35 % / ~15 % of backward before its end: closing a bucket after each of them sends their
    # gradients on their way then, so that only enc1's 154 KB (the last bucket) is exchanged after the last kernel.
    LEVEL_ENDS = ("model.enc3.0.bias", "model.enc2.0.bias")


    @classmethod
    def for_model(cls, model: torch.nn.Module, bucket_bytes: int = 8 << 20, group=None) -> "FlatGradBuffer":
        from . import engine
        params = dict(model.named_parameters())
        order = engine.grad_production_order()
        if set(order) != set(params):
            raise RuntimeError("FlatGradBuffer: the model's parameters do not match the hot path's production order")
        dev = next(iter(params.values())).device
        return cls([(n, params[n].shape) for n in order], dev, bucket_bytes, group, close_after=cls.LEVEL_ENDS)

    # ---- engine.GradSink protocol ----
    def dst(self, name: str, shape) -> torch.Tensor:
        t = self.grads[name]
        if tuple(t.shape) != tuple(shape):
            raise RuntimeError(f"FlatGradBuffer: gradient of {name} has shape {tuple(shape)}, expected {tuple(t.shape)}")
        return t

    def zero_dst(self, name: str, shape, pool=None) -> torch.Tensor:
        """Slice for a gradient that is exactly zero: ``begin`` zeroed the whole buffer (one fill), nothing to do."""
        return self.dst(name, shape)

    def closes(self, name: str) -> bool:
        """A bucket leaves for the all-reduce when ``name`` is ready (nothing leaves early on a single rank)."""
        return name in self._closes and world_size(self.group) > 1

    def ready(self, name: str) -> None:
        b = self._closes.get(name)
        if b is None:
            return
        if b != self._launched:
            raise RuntimeError(f"FlatGradBuffer: bucket {b} closed out of order (expected {self._launched})")
        self._launched += 1
        if world_size(self.group) == 1:
            return
        lo, hi = self.buckets[b]
        self._work.append(dist.all_reduce(self.flat[lo:hi], op=dist.ReduceOp.SUM, group=self.group, async_op=True))

    # ---- step protocol ----
    def begin(self) -> None:
        """Call before a backward pass (after the previous step's ``wait``)."""
        if self._work:
            raise RuntimeError("FlatGradBuffer.begin: the previous exchange was not waited for")
        self._launched = 0
        self.flat.zero_()      # ONE fill: the exactly-zero conv-bias gradients need no launches of their own

    def wait(self) -> None:
        """Join every bucket's all-reduce (the current stream waits on NCCL's; no host sync on CUDA)."""
        if self._launched not in (0, len(self.buckets)):
            raise RuntimeError(f"FlatGradBuffer.wait: only {self._launched} of {len(self.buckets)} buckets were produced")
        for w in self._work:
            w.wait()
        self._work = []


class GradientAllReduce:
    """Data-parallel gradient exchange for an ``EnhancedUNet``: attaches a ``FlatGradBuffer`` to the model so that
    ``loss.backward()`` writes gradients into it and launches the bucket all-reduces as it goes (overlapped with the
    remaining backward kernels).  ``wait()`` before ``optimizer.step``.  At world size 1 nothing is exchanged."""

    def __init__(self, model: torch.nn.Module, bucket_bytes: int = 8 << 20, group=None):
        self.model = model
        self.buffer = FlatGradBuffer.for_model(model, bucket_bytes, group)
        model.grad_sink = self.buffer

    def wait(self) -> None:
        self.buffer.wait()

    def detach(self) -> None:
        self.model.grad_sink = None


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
