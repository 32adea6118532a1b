#!/usr/bin/env python
"""Headline benchmark: Enhanced-UNet training throughput (images/s) at 512x512, batch 16 per GPU, bf16
tensor-core path, on N B200s (BASELINE.json metric / configs[1], configs[2]).

  python bench.py --gpus N --steps K --warmup W            # our arm (torchrun launches N ranks)
  python bench.py --impl reference --steps K --warmup W     # CPU arm: the oracle port of the reference
                                                            # path on the box's host cores

A "step" is one full training step on synthetic bright-field tensors: forward, fused focal+dice+tversky
loss, backward, (N>1: gradient all-reduce over NCCL), global-norm clip + AdamW.  Prints ONE JSON line.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

BATCH, RES = 16, 512
METRIC = "train images/sec at 512x512 (Enhanced-UNet, batch 16/GPU)"

# conv layers of the primary body: (Cin, Cout, level) - SURVEY.md §8 layer table
CONVS = [(3, 64, 0), (64, 64, 0), (64, 128, 1), (128, 128, 1), (128, 256, 2), (256, 256, 2), (256, 512, 3), (512, 512, 3),
         (768, 256, 2), (256, 256, 2), (384, 128, 1), (128, 128, 1), (192, 64, 0), (64, 64, 0)]


def conv_flops_train(batch: int, res: int) -> float:
    """Algorithmic conv FLOPs of one training step: fwd + dgrad + wgrad (no dgrad for the first conv),
    2*M*N*K each, plus the 2Hx2W tail convs (1029.8 GFLOP per 512^2 image)."""
    total = 0.0
    for i, (ci, co, lvl) in enumerate(CONVS):
        m = batch * (res >> lvl) ** 2
        f = 2.0 * m * co * 9 * ci
        total += f * (2 if i == 0 else 3)
    m2 = batch * (2 * res) ** 2
    total += 3 * 2.0 * m2 * 64 * 27          # enhance.0 (3->64, 3x3) fwd + dgrad + wgrad
    total += 3 * 2.0 * m2 * 3 * 64 * 2       # dec1 + enhance.3 (1x1) fwd + dgrad + wgrad
    return total


def hbm_bytes_per_step(batch: int, res: int) -> dict:
    """Algorithmic HBM bytes per training step of the bandwidth-bound kernels (every logical input read
    once + every logical output written once at its storage dtype; bf16/fp16 = 2 B, fp32 = 4 B)."""
    bn_layers = [(0, 64)] * 2 + [(1, 128)] * 2 + [(2, 256)] * 2 + [(3, 512)] * 2 + [(2, 256)] * 2 + [(1, 128)] * 2 + [(0, 64)] * 2
    elems = sum(batch * (res >> l) ** 2 * c for l, c in bn_layers)
    pooled = sum(batch * (res >> l) ** 2 * c // 4 for l, c in ((0, 64), (1, 128), (2, 256)))
    ups = [(3, 512), (2, 256), (1, 128)]          # (input level, channels)
    up_bytes = sum(batch * (res >> l) ** 2 * c * 2 * 5 for l, c in ups)        # read in (2 B) + write 4x out (8 B)
    m1, m2 = batch * res * res, 4 * batch * res * res
    return {
        "eunet_bn_apply_relu": elems * 4 + pooled * 2,
        "eunet_bn_bwd_reduce": elems * 4,
        "eunet_bn_bwd_apply": elems * 6,
        "eunet_upsample2_fwd": up_bytes,
        "eunet_upsample2_bwd": up_bytes,
        "eunet_maxpool2_bwd": sum(batch * (res >> l) ** 2 * c * (2 + 4) + batch * (res >> l) ** 2 * c // 4 * 2
                                  for l, c in ((0, 64), (1, 128), (2, 256))),
        "eunet_tail_out_fwd": m2 * (128 + 16 + 12),
        "eunet_tail_bwd_reduce": m2 * (128 + 16),
        "eunet_tail_bwd_dmid": m2 * (128 + 16 + 128),
        "eunet_tail_dec1_fwd": m1 * (128 + 16),
        "eunet_tail_dec1_bwd": m1 * (16 + 128 + 128),
        "eunet_loss_fwd": m1 * (48 + 8),
        "eunet_loss_bwd": m1 * (48 + 8 + 48),
    }


class ClockSampler:
    """Samples SM clocks / throttle reasons with nvidia-smi while the timed region runs."""

    def __init__(self, index: int):
        self.index = index
        self.samples = []
        self.reasons = set()
        self.max_mhz = None
        self._stop = threading.Event()
        self._t = None

    def _run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                                     capture_output=True, text=True, timeout=5).stdout.strip().split(",")
                self.samples.append(float(out[0]))
                self.max_mhz = float(out[1])
                for n, v in zip(names, out[2:]):
                    if v.strip().lower().startswith("active"):
                        self.reasons.add(n)
            except Exception:
                pass
            self._stop.wait(0.2)

    def __enter__(self):
        self._t = threading.Thread(target=self._run, daemon=True)
        self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._t.join(timeout=6)

    def summary(self):
        return {"sm_mhz": statistics.median(self.samples) if self.samples else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


def synth_batch(batch: int, res: int, seed: int, device):
    """Synthetic single-channel bright-field plane replicated to 3 channels + a 3-class label map
    (SURVEY.md §8d): bright noisy background with darker Gaussian blobs; weak blobs = live, strong = dead."""
    g = torch.Generator(device=device).manual_seed(seed)
    n_blobs = max(1, res * res // 4096)
    yy = torch.arange(res, device=device, dtype=torch.float32).view(1, 1, res, 1)
    xx = torch.arange(res, device=device, dtype=torch.float32).view(1, 1, 1, res)
    img = 0.75 + 0.05 * torch.randn(batch, 1, res, res, device=device, generator=g)
    label = torch.zeros(batch, res, res, device=device, dtype=torch.int64)
    cy = torch.rand(batch, n_blobs, device=device, generator=g) * res
    cx = torch.rand(batch, n_blobs, device=device, generator=g) * res
    sig = 3 + 7 * torch.rand(batch, n_blobs, device=device, generator=g)
    amp = 0.15 + 0.30 * torch.rand(batch, n_blobs, device=device, generator=g)
    for k0 in range(0, n_blobs, 16):          # 16 blobs per pass keeps the generator to a few launches
        sl = slice(k0, min(n_blobs, k0 + 16))
        a_ = amp[:, sl].view(batch, -1, 1, 1)
        d2 = (yy - cy[:, sl].view(batch, -1, 1, 1)) ** 2 + (xx - cx[:, sl].view(batch, -1, 1, 1)) ** 2
        blob = a_ * torch.exp(-d2 / (2 * sig[:, sl].view(batch, -1, 1, 1) ** 2))
        img = img - blob.sum(1, keepdim=True)
        cls = torch.where(a_ < 0.3, 1, 2) * (blob > 0.5 * a_)      # 0 outside, 1 live, 2 dead (dead wins overlaps)
        label = torch.maximum(label, cls.amax(1))
    img = img.clamp_(0, 1).expand(batch, 3, res, res).contiguous()
    return img, label


# ---------------------------------------------------------------------------------------------
# CPU arm: oracle port of the reference path (the reference itself cannot travel to the GPU box)
# ---------------------------------------------------------------------------------------------
def cpu_step_factory(batch: int, res: int, threads: int):
    import oracle
    torch.set_num_threads(threads)
    sd = oracle.make_state_dict(0, randomize_bn=False)
    params = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and "running" not in k else v) for k, v in sd.items()}
    plist = [p for p in params.values() if p.requires_grad]
    opt = torch.optim.AdamW(plist, lr=4e-3, weight_decay=1e-4, betas=(0.9, 0.999))
    x, t = synth_batch(batch, res, 1234, torch.device("cpu"))

    def step():
        opt.zero_grad()
        y, nb = oracle.unet_forward(params, x, train=True)
        loss = oracle.batch_loss(y, t)
        loss.backward()
        torch.nn.utils.clip_grad_norm_(plist, 1.0)
        opt.step()
        for k, v in nb.items():
            params[k] = v
        return float(loss)

    return step


def run_cpu(steps: int, warmup: int, batch: int = 1, res: int = 512):
    threads = os.cpu_count() or 1
    step = cpu_step_factory(batch, res, threads)
    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = time.perf_counter() - t0
    # throughput is quoted per 512x512-image equivalent so that it shares the GPU arm's unit:
    # work is proportional to pixels (every layer is convolutional)
    px_ratio = (res * res) / float(RES * RES)
    return {"img_per_s_native": batch * steps / dt, "img512_per_s": batch * steps * px_ratio / dt, "cores": threads,
            "sample": f"{steps} full train steps (fwd+loss+bwd+clip+AdamW) of the oracle port, fp32, batch {batch} "
                      f"(bounded sample of the batch-16 workload), 3x{res}x{res}, {threads} threads",
            "ms_per_step": 1e3 * dt / steps}


def main_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    r = run_cpu(max(1, args.steps), max(0, args.warmup))
    line = {"impl": "reference", "metric": METRIC, "value": r["img512_per_s"], "unit": "images/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "Enhanced-UNet train step, oracle port of reference models.py/train_eval.py on host cores",
                       "batch": 1, "resolution": 512},
            "cpu_baseline": {"value": r["img512_per_s"], "unit": "images/s", "cores": r["cores"], "kind": "port",
                             "sample": r["sample"]},
            "e2e": {"value": r["img512_per_s"], "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------
def main_gpu(args):
    import torch.distributed as dist
    from enhanced_unet_b200 import lib, parallel
    from enhanced_unet_b200.optim import ClippedAdamW
    from enhanced_unet_b200.models import EnhancedUNet
    from enhanced_unet_b200.ops import combined_loss

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib.load()

    torch.manual_seed(0)
    model = EnhancedUNet(3, dtype="bf16").to(dev).train()
    params = [p for p in model.parameters()]
    parallel.broadcast_parameters(list(model.parameters()) + list(model.buffers()))
    allreduce = parallel.GradientAllReduce(model)   # gradients land in one flat buffer; buckets are exchanged DURING backward
    opt = ClippedAdamW(params, on_update=model._packs.invalidate)
    x_dev, t_dev = synth_batch(BATCH, RES, 1234 + 1000 * rank, dev)
    x_host = x_dev.cpu().pin_memory()
    t_host = t_dev.cpu().pin_memory()
    h2d = x_host.numel() * 4 + t_host.numel() * 8

    def step(x, t):
        for p in params:
            p.grad = None
        y = model(x)
        loss = combined_loss(y, t)
        loss.backward()             # launches the bucketed gradient all-reduce as gradients are finished (N > 1)
        allreduce.wait()
        opt.step(grad_scale=1.0 / world)
        return loss

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, n):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            tms = torch.tensor([ms], device=dev, dtype=torch.float64)
            dist.all_reduce(tms, op=dist.ReduceOp.MAX)
            ms = float(tms)
        return ms

    for _ in range(args.warmup):
        step(x_dev, t_dev)
    lib.COUNTERS.clear()
    clk = ClockSampler(local)
    if rank == 0:          # one nvidia-smi poller per job, not one per rank
        with clk:
            ms = timed(lambda: step(x_dev, t_dev), args.steps)
    else:
        ms = timed(lambda: step(x_dev, t_dev), args.steps)
    launches = sum(lib.COUNTERS.values())
    value = world * BATCH * args.steps / (ms / 1e3)

    # end to end through the public API with HOST buffers: every step copies its batch from pinned host memory
    # (data.HostBatchPrefetcher: copy stream, double-buffered, so the copy of step i+1 runs under step i) and reads
    # the loss back to the host.  Exactly `steps` H2D batch copies and `steps` D2H loss reads inside the timed region.
    from enhanced_unet_b200.data import HostBatchPrefetcher
    pf = HostBatchPrefetcher(dev)

    def e2e_run(n):
        pf.submit(x_host, t_host)
        for i in range(n):
            x, t = pf.get()
            if i + 1 < n:
                pf.submit(x_host, t_host)
            float(step(x, t).detach())

    e2e_run(min(2, max(1, args.warmup)))
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    e2e_run(args.steps)
    e1.record()
    barrier()
    ms_e2e = e0.elapsed_time(e1)
    if world > 1:
        tms = torch.tensor([ms_e2e], device=dev, dtype=torch.float64)
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
        ms_e2e = float(tms)
    e2e_value = world * BATCH * args.steps / (ms_e2e / 1e3)

    roof = None
    cpu = None
    # roofline of the dominant kernel (the tcgen05 implicit-GEMM conv, forward + dgrad launches):
    # algorithmic FLOPs of those launches / their CUDA-event time on the launching stream.
    # Every rank runs the instrumented steps (they contain the collective); rank 0 reports.
    lib.PROFILE = []
    for _ in range(2):
        step(x_dev, t_dev)
    prof = lib.collect_profile()
    lib.PROFILE = None
    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = peaks.get("bf16_tflops_sustained") or 1400.0
        which = "measured sustained (MEASURED_PEAKS.json)" if "bf16_tflops_sustained" in peaks else "fallback 1.4 PF sustained"
        if "eunet_conv3x3_fwd" in prof:
            fl, msk, n = prof["eunet_conv3x3_fwd"]
            ach = fl / (msk / 1e3) / 1e12
            hbm_peak = peaks.get("hbm_gbs") or 6650.0
            hb = hbm_bytes_per_step(BATCH, RES)
            hbm = {k: {"GB/s": round(hb[k] / (prof[k][1] / 2 / 1e3) / 1e9, 1), "frac": round(hb[k] / (prof[k][1] / 2 / 1e3) / 1e9 / hbm_peak, 3)}
                   for k in hb if k in prof and prof[k][1] > 0}
            wg = prof.get("eunet_conv3x3_wgrad")
            roof = {"bound": "tensor", "kernel": "conv3x3 implicit-GEMM tcgen05 kernels, forward + dgrad launches "
                                                 "(conv3x3_halo_kernel / conv3x3_fwd_tc_kernel)",
                    "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak,
                    # dram__bytes_read+write per launch, ncu --set full, mean of the 8 launches in
                    # profiles/conv_halo_r1e_ncu_summary.txt (equals the algorithmic input+output bytes of those layers +-3 %)
                    "traffic": 3.63e8, "peak_source": which,
                    "launches_per_step": n // 2, "kernel_ms_per_step": msk / 2,
                    "wgrad_tflops": (wg[0] / (wg[1] / 1e3) / 1e12) if wg else None,
                    "hbm_kernels": hbm, "hbm_peak_gbs": hbm_peak,
                    "all_kernels_ms_per_step": {k: round(v[1] / 2, 4) for k, v in prof.items()}}
        if world == 1 and not args.no_cpu:
            r = run_cpu(2, 1)
            cpu = {"value": r["img512_per_s"], "unit": "images/s", "cores": r["cores"], "kind": "port", "sample": r["sample"]}
        line = {"metric": METRIC, "value": value, "unit": "images/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
                "data": "synthetic",
                "config": {"workload": "Enhanced-UNet bf16 training, batch 16/GPU, 3x512x512 (BASELINE configs[1]/[2]): "
                                       "fwd + focal/dice/tversky loss + bwd + clip + AdamW",
                           "global_batch": BATCH * world, "resolution": RES, "parallelism": f"dp{world}",
                           "l2": "working set (>5 GB of activations per step) far exceeds the 126 MB L2; no flush needed",
                           "conv_tflop_per_step": conv_flops_train(BATCH, RES) / 1e12},
                "clocks": clk.summary(),
                "e2e": {"value": e2e_value, "unit": "images/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4},
                "gpu_launches": launches, "roofline": roof, "cpu_baseline": cpu}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the bounded CPU baseline leg")
    args = ap.parse_args()
    if args.impl == "reference":
        main_reference(args)
    else:
        main_gpu(args)


if __name__ == "__main__":
    main()
