"""Host -> device input staging for the training / inference loops (the reference feeds
``batch['images'].to(device)`` synchronously, train_eval.py:244).

``HostBatchPrefetcher`` double-buffers pinned host batches onto the GPU on a dedicated copy stream, so the H2D copy of
step i+1 runs under the kernels of step i.  torch is used for memory, streams and events only; no arithmetic."""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import torch


class HostBatchPrefetcher:
    """``submit(*host_tensors)`` starts the asynchronous copy of one batch into the next device slot;
    ``get()`` returns the oldest submitted batch as device tensors, ordered after its copy on the current stream.

    A slot is reused every ``depth`` batches: its copy is ordered (event) behind all work that was enqueued on the
    compute stream when ``submit`` was called, i.e. behind the last consumer of that slot as long as ``submit`` for
    batch i+depth-1 is called after the kernels of batch i-1 were launched (the usual loop below).

        pf.submit(x0, t0)
        for i in range(steps):
            x, t = pf.get()
            if i + 1 < steps: pf.submit(x_next, t_next)
            loss = step(x, t)
    """

    def __init__(self, device, depth: int = 2):
        self.dev = torch.device(device)
        if self.dev.type != "cuda":
            raise RuntimeError("HostBatchPrefetcher stages batches onto a CUDA device")
        self.depth = depth
        self.stream = torch.cuda.Stream(self.dev)
        self._slots: List[Optional[List[torch.Tensor]]] = [None] * depth
        self._ready: List[Optional[torch.cuda.Event]] = [None] * depth
        self._head = 0      # next slot to fill
        self._tail = 0      # next slot to hand out
        self._pending = 0

    def submit(self, *host: torch.Tensor) -> None:
        if self._pending >= self.depth:
            raise RuntimeError("HostBatchPrefetcher: every slot holds an unconsumed batch (call get() first)")
        s = self._head
        bufs = self._slots[s]
        if bufs is None or len(bufs) != len(host) or any(b.shape != h.shape or b.dtype != h.dtype for b, h in zip(bufs, host)):
            bufs = [torch.empty(h.shape, dtype=h.dtype, device=self.dev) for h in host]
            self._slots[s] = bufs
        free = torch.cuda.Event()
        free.record(torch.cuda.current_stream(self.dev))       # everything launched so far (incl. this slot's last consumer)
        self.stream.wait_event(free)
        with torch.cuda.stream(self.stream):
            for b, h in zip(bufs, host):
                b.copy_(h, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(self.stream)
        self._ready[s] = ev
        self._head = (s + 1) % self.depth
        self._pending += 1

    def get(self) -> Tuple[torch.Tensor, ...]:
        if self._pending == 0:
            raise RuntimeError("HostBatchPrefetcher.get: nothing was submitted")
        s = self._tail
        torch.cuda.current_stream(self.dev).wait_event(self._ready[s])
        self._tail = (s + 1) % self.depth
        self._pending -= 1
        return tuple(self._slots[s])


def bytes_of(tensors: Sequence[torch.Tensor]) -> int:
    return sum(t.numel() * t.element_size() for t in tensors)
