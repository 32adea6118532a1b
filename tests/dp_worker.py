"""Worker of tests/test_gpu_parallel.py (one process per GPU under torch.distributed.run, NCCL): data-parallel training
of the drop-in EnhancedUNet on the hardware.  Checks:
  1. the reduced gradient equals the sum of the per-rank gradients computed WITHOUT the exchange (rank 0 recomputes every
     rank's batch single-process), i.e. what the optimiser sees after ``grad_scale = 1/world`` is their mean;
  2. after several optimiser steps every rank holds bit-identical parameters (and rank-local BN running statistics that
     differ, as DistributedDataParallel over the reference's plain BatchNorm2d would have them)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def main():
    import oracle
    from enhanced_unet_b200 import parallel
    from enhanced_unet_b200.models import EnhancedUNet
    from enhanced_unet_b200.ops import combined_loss
    from enhanced_unet_b200.optim import ClippedAdamW
    world, rank, local = int(os.environ["WORLD_SIZE"]), int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    sd = oracle.make_state_dict(3)
    batches = [(oracle.make_input(2, 64, 64, 100 + r).to(dev), oracle.make_target(2, 64, 64, 200 + r).to(dev)) for r in range(world)]

    def fresh():
        m = EnhancedUNet(3)
        m.load_state_dict(sd)
        return m.to(dev).train()

    # ---- 1. reduced gradient == sum of the per-rank gradients
    m = fresh()
    ar = parallel.GradientAllReduce(m)
    x, t = batches[rank]
    combined_loss(m(x), t).backward()
    ar.wait()
    reduced = {n: p.grad.clone() for n, p in m.named_parameters()}
    if rank == 0:
        total = None
        for xr, tr in batches:
            ms = fresh()
            combined_loss(ms(xr), tr).backward()
            g = {n: p.grad.double() for n, p in ms.named_parameters()}
            total = g if total is None else {n: total[n] + g[n] for n in g}
        worst = 0.0
        for n, want in total.items():
            if float(want.abs().max()) == 0.0:
                assert float(reduced[n].abs().max()) == 0.0, n
                continue
            e = float((reduced[n].double() - want).abs().max() / want.abs().max())
            worst = max(worst, e)
            assert e <= 2e-4, (n, e)                 # fp32 atomics: summation order only
        print(f"DP_OK reduced == sum over {world} ranks, worst normalised error {worst:.2e}", flush=True)
    ar.detach()

    # ---- 2. replicas stay bit-identical through optimiser steps
    m = fresh()
    parallel.broadcast_parameters(list(m.parameters()) + list(m.buffers()))
    ar = parallel.GradientAllReduce(m)
    opt = ClippedAdamW(list(m.parameters()), lr=1e-3)
    for _ in range(4):
        opt.zero_grad(set_to_none=True)
        combined_loss(m(x), t).backward()
        ar.wait()
        opt.step(grad_scale=1.0 / world)
    m.check_numerics()
    flat = torch.cat([p.detach().reshape(-1) for p in m.parameters()]).view(torch.int32).to(torch.int64)
    sig = torch.stack([flat.sum(), (flat * (torch.arange(flat.numel(), device=dev) % 8191 + 1)).sum()])
    sigs = [torch.empty_like(sig) for _ in range(world)]
    dist.all_gather(sigs, sig)
    assert all(torch.equal(s, sigs[0]) for s in sigs), "replicas diverged"
    rm = m.model.enc1[1].running_mean.clone()
    rms = [torch.empty_like(rm) for _ in range(world)]
    dist.all_gather(rms, rm)
    if rank == 0:
        assert not torch.equal(rms[0], rms[1])       # per-rank BatchNorm statistics (different batches), as plain BatchNorm2d under DDP
        print("DP_OK parameters bit-identical on every rank after 4 steps", flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
