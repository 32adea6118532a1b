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
        params = list(m.parameters())
        g = torch.Generator().manual_seed(100 + rank)
        for p in params:
            p.grad = torch.randn(p.shape, generator=g)
        mine = [p.grad.clone() for p in params]
        ar = parallel.GradientAllReduce(params, bucket_bytes=4 << 20)
        assert len(ar.buckets) >= 3 and sum(len(b) for b in ar.buckets) == len(params)
        assert ar.buckets[0][0] is params[-1]       # reverse execution order: the tail first
        ar.reduce()
        ar.wait()
        other = torch.Generator().manual_seed(100 + (1 - rank))
        for p, a in zip(params, mine):
            b = torch.randn(p.shape, generator=other)
            assert torch.allclose(p.grad, a + b, rtol=0, atol=1e-6)
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
