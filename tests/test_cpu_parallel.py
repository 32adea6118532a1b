"""Host-side data-parallel logic on CPU: world_size-2 gloo processes:
(b) all-reduced gradient == sum of per-rank gradients (averaged by the optimiser's grad_scale),
(c) integer metric counts summed over ranks == single-process counts, bit-exactly; plus sharding."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, tmp):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from enhanced_unet_b200 import parallel
        from enhanced_unet_b200.models import EnhancedUNet
        import oracle
        torch.manual_seed(rank)                     # different init per rank on purpose
        m = EnhancedUNet(3)
        parallel.broadcast_parameters(list(m.parameters()) + list(m.buffers()))
        ref = EnhancedUNet(3)
        torch.manual_seed(0)
        ref = EnhancedUNet(3)
        for a, b in zip(m.state_dict().values(), ref.state_dict().values()):
            assert torch.equal(a, b)               # everyone holds rank 0's parameters
        # the exchange as backward drives it: gradients are written into the flat buffer in production order and every
        # bucket's all-reduce is launched the moment its last gradient is there
        from enhanced_unet_b200 import engine
        ar = parallel.GradientAllReduce(m, bucket_bytes=4 << 20)
        buf = m.grad_sink
        assert buf is ar.buffer and buf.numel == sum(p.numel() for p in m.parameters())
        assert len(buf.buckets) >= 3 and buf.buckets[0][0] == 0 and buf.buckets[-1][1] == buf.numel
        assert all(a[1] == b[0] for a, b in zip(buf.buckets, buf.buckets[1:]))          # contiguous, in order
        assert buf.names[0].startswith("enhance.") and buf.names[-1] == "model.enc1.0.bias"   # the tail first
        # hand-over points for the deferred filter-gradient unpack (engine.backward): exactly the names that close a bucket
        closing = [n for n in buf.names if buf.closes(n)]
        assert len(closing) == len(buf.buckets) and closing[-1] == buf.names[-1]
        assert all(n in closing for n in parallel.FlatGradBuffer.LEVEL_ENDS)
        assert not engine.GradSink(torch.device("cpu")).closes(buf.names[-1])      # single-rank sink: one unpack at the end
        params = dict(m.named_parameters())
        g = torch.Generator().manual_seed(100 + rank)
        other = torch.Generator().manual_seed(100 + (1 - rank))
        want = {}
        buf.begin()
        for n in engine.grad_production_order():
            a = torch.randn(params[n].shape, generator=g)
            want[n] = a + torch.randn(params[n].shape, generator=other)
            buf.dst(n, params[n].shape).copy_(a)
            buf.ready(n)
        buf.wait()
        for n in buf.names:
            assert torch.allclose(buf.grads[n], want[n], rtol=0, atol=1e-6), n
        # a second step reuses the buffer; closing a bucket out of order is an error, not a silent wrong exchange
        buf.begin()
        try:
            buf.ready(buf.names[-1])
            raise AssertionError("out-of-order bucket was accepted")
        except RuntimeError:
            pass
        # metrics: each rank counts its shard; the int64 sum is the single-process result
        rng = np.random.default_rng(7)
        pred, gt = rng.integers(0, 3, (6, 16, 16)), rng.integers(0, 3, (6, 16, 16))
        lo, hi = parallel.shard_batch(6, rank, world)
        local = torch.from_numpy(oracle.confusion_counts(pred[lo:hi], gt[lo:hi]).sum(0))
        total = parallel.allreduce_counts(local)
        assert torch.equal(total, torch.from_numpy(oracle.confusion_counts(pred, gt).sum(0)))
        open(os.path.join(tmp, f"ok{rank}"), "w").write("ok")
    finally:
        dist.destroy_process_group()


def test_gradient_allreduce_and_counts_world2(tmp_path):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert (tmp_path / "ok0").exists() and (tmp_path / "ok1").exists()


def test_shard_batch_covers_everything():
    from enhanced_unet_b200.parallel import shard_batch
    for n in (0, 1, 7, 8, 33):
        for w in (1, 2, 4, 8):
            seen = []
            for r in range(w):
                lo, hi = shard_batch(n, r, w)
                seen += list(range(lo, hi))
            assert seen == list(range(n))
