#!/usr/bin/env python
"""Headline benchmark: Enhanced-UNet training throughput (images/s) at 512x512, batch 16 per GPU, 16-bit tensor-core
path, on N B200s (BASELINE.json metric / configs[1], configs[2]).

  python bench.py --gpus N --steps K --warmup W            # our arm (torchrun launches N ranks)
  python bench.py --impl reference --steps K --warmup W     # CPU arm: the oracle port of the reference path on the
                                                            # box's host cores (a bounded sample of the same workload)

A "step" is one full training step on synthetic bright-field tensors: forward, fused focal+dice+tversky loss, backward,
(N>1: gradient all-reduce over NCCL), global-norm clip + AdamW.  Prints ONE JSON line.  Besides the contract keys the
line carries
  roofline            the tensor-core conv group (K >= 576 layers; ALGORITHMIC FLOPs / CUDA-event time), wgrad, the
                      K = 27 convs and every bandwidth kernel against their own (HBM) roofline
  inference           BASELINE configs[3] / [4]: eval forward -> 2x2-mean softmax -> mask cascade -> confusion counts,
                      plus the confusion kernel alone at 16 B/px (int64 masks, the reference dtypes) and 2 B/px (uint8)
  stock_pytorch_b200  comparator OUTSIDE the product path: the same network as plain torch.nn on cuDNN (fp32 and
                      bf16 autocast / channels_last), same batch, same step
  dp_check            (N > 1) parameters and reduced gradients are bit-identical on every rank after the timed steps
  cpu_baseline        the oracle port on the host cores (N = 1 only)
"""
from __future__ import annotations

import argparse
import hashlib
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
CPU_SAMPLE_BATCH = 2          # images of 3x512x512 per CPU step: a bounded sample (1/8) of the batch-16 step

# conv layers of the primary body: (Cin, Cout, level) - SURVEY.md §8 layer table
CONVS = [(3, 64, 0), (64, 64, 0), (64, 128, 1), (128, 128, 1), (128, 256, 2), (256, 256, 2), (256, 512, 3), (512, 512, 3),
         (768, 256, 2), (256, 256, 2), (384, 128, 1), (128, 128, 1), (192, 64, 0), (64, 64, 0)]


def conv_flops_fwd(batch: int, res: int) -> float:
    """Algorithmic conv FLOPs of one forward pass (1,310,592 per input pixel)."""
    total = sum(2.0 * batch * (res >> lvl) ** 2 * co * 9 * ci for ci, co, lvl in CONVS)
    m2 = batch * (2 * res) ** 2
    return total + 2.0 * m2 * 64 * 27 + 2 * 2.0 * m2 * 3 * 64


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
    once + every logical output written once at its storage dtype; 16-bit = 2 B, fp32 = 4 B).  Keys are C-ABI entry
    points, `name[k27]` = the launches of that entry point on the 3-channel layers (enc1.0, enhance.0: K = 27)."""
    bn_layers = [(0, 64)] * 2 + [(1, 128)] * 2 + [(2, 256)] * 2 + [(3, 512)] * 2 + [(2, 256)] * 2 + [(1, 128)] * 2 + [(0, 64)] * 2
    elems = sum(batch * (res >> l) ** 2 * c for l, c in bn_layers)
    pooled = sum(batch * (res >> l) ** 2 * c // 4 for l, c in ((0, 64), (1, 128), (2, 256)))
    ups = [(3, 512), (2, 256), (1, 128)]          # (input level, channels): enc4.4, dec4.4, dec3.4 feed the upsample only
    up_elems = sum(batch * (res >> l) ** 2 * c for l, c in ups)
    up_bytes = up_elems * 2 * 5                   # read in (2 B) + write 4x out (8 B)
    m1, m2 = batch * res * res, 4 * batch * res * res
    return {
        "eunet_bn_apply_relu": (elems - up_elems) * 4 + pooled * 2,
        "eunet_bn_apply_relu_upsample2": up_bytes,         # training: BN apply + ReLU + upsample fused (no stored activation)
        "eunet_bn_bwd_reduce": (elems - 4 * pooled) * 4,
        "eunet_bn_bwd_apply": (elems - 4 * pooled) * 6,
        # enc1.4 / enc2.4 / enc3.4: BN backward with the max-pool gradient routed on the fly (dskip 2 + y 2 + dpool 0.5 B)
        "eunet_bn_bwd_reduce_pool": 4 * pooled * 4 + pooled * 2,
        "eunet_bn_bwd_apply_pool": 4 * pooled * 6 + pooled * 2,
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
        # K = 27 convolutions (SURVEY.md §8 layer table: enc1.0 25 / 537 MB, enhance.0 101 / 2147 MB at 2 B per element)
        "eunet_conv3x3_fwd[k27]": m1 * (3 * 2 + 64 * 2) + m2 * (3 * 2 + 64 * 2),
        "eunet_conv3x3_wgrad[k27]": m1 * (3 * 2 + 64 * 2),
        "eunet_tail_bwd_fused[k27]": m2 * (128 + 16 + 32 + 16),
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


def run_cpu(steps: int, warmup: int, batch: int = CPU_SAMPLE_BATCH, res: int = RES):
    threads = os.cpu_count() or 1
    step = cpu_step_factory(batch, res, threads)
    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = time.perf_counter() - t0
    # throughput in the GPU arm's unit (512x512 images/s): work is proportional to pixels (every layer is convolutional)
    px_ratio = (res * res) / float(RES * RES)
    return {"img_per_s_native": batch * steps / dt, "img512_per_s": batch * steps * px_ratio / dt, "cores": threads,
            "sample": f"{steps} full train steps (fwd+loss+bwd+clip+AdamW) of the oracle port of reference models.py / train_eval.py, "
                      f"fp32, {batch} images of 3x{res}x{res} per step (a bounded sample: {batch} of the {BATCH} images of the "
                      f"GPU arm's step), {threads} host threads",
            "ms_per_step": 1e3 * dt / steps}


def main_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    r = run_cpu(max(1, args.steps), max(0, args.warmup))
    line = {"impl": "reference", "metric": METRIC, "value": r["img512_per_s"], "unit": "images/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "Enhanced-UNet training, batch 16/GPU, 3x512x512 (BASELINE configs[1]/[2]): fwd + focal/dice/tversky "
                                   f"loss + bwd + clip + AdamW; CPU arm = the oracle port of the reference path on the host cores, each "
                                   f"step a bounded sample ({CPU_SAMPLE_BATCH} of the {BATCH} images) of that workload",
                       "global_batch": BATCH, "resolution": RES, "cpu_images_per_step": CPU_SAMPLE_BATCH},
            "cpu_baseline": {"value": r["img512_per_s"], "unit": "images/s", "cores": r["cores"], "kind": "port",
                             "sample": r["sample"]},
            "e2e": {"value": r["img512_per_s"], "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------
# comparator (NOT the product path): the same network as plain torch.nn on cuDNN - BASELINE.md §4.6
# ---------------------------------------------------------------------------------------------
def stock_pytorch_block(dev, steps: int = 5, warmup: int = 3):
    import torch.nn as nn
    import torch.nn.functional as F

    def block(ci, co):
        return nn.Sequential(nn.Conv2d(ci, co, 3, padding=1), nn.BatchNorm2d(co), nn.ReLU(inplace=True),
                             nn.Conv2d(co, co, 3, padding=1), nn.BatchNorm2d(co), nn.ReLU(inplace=True))

    class StockUNet(nn.Module):            # reference models.py:199-238 + 308-313, 337 (fallback body), stock modules
        def __init__(self):
            super().__init__()
            self.enc1, self.enc2, self.enc3, self.enc4 = block(3, 64), block(64, 128), block(128, 256), block(256, 512)
            self.dec4, self.dec3, self.dec2 = block(768, 256), block(384, 128), block(192, 64)
            self.dec1 = nn.Conv2d(64, 3, 1)
            self.pool = nn.MaxPool2d(2)
            self.up = nn.Upsample(scale_factor=2, mode="bilinear", align_corners=False)
            self.enhance = nn.Sequential(nn.Conv2d(3, 64, 3, padding=1), nn.BatchNorm2d(64), nn.ReLU(inplace=True), nn.Conv2d(64, 3, 1))

        def forward(self, x):
            e1 = self.enc1(x); e2 = self.enc2(self.pool(e1)); e3 = self.enc3(self.pool(e2)); e4 = self.enc4(self.pool(e3))
            d4 = self.dec4(torch.cat([self.up(e4), e3], 1))
            d3 = self.dec3(torch.cat([self.up(d4), e2], 1))
            d2 = self.dec2(torch.cat([self.up(d3), e1], 1))
            out = self.dec1(self.up(d2))
            return out + self.enhance(out)

    def loss_fn(logits, target):          # reference train_eval.py:37-60, 134-197, 306-337 in plain torch ops, batched
        z = F.avg_pool2d(logits.float(), 2)
        lp = F.log_softmax(z, 1)
        p = lp.exp()
        oh = F.one_hot(target, 3).permute(0, 3, 1, 2).float()
        w = torch.tensor([1.0, 20.0, 10.0], device=z.device)
        al = torch.tensor([1.0, 8.0, 5.0], device=z.device)
        ce = -(lp * oh).sum(1) * w[target]
        focal = (al[target] * (1 - torch.exp(-ce)) ** 5 * ce).mean((1, 2))
        inter, sp, st = (p * oh).sum((2, 3)), p.sum((2, 3)), oh.sum((2, 3))
        dice = ((torch.tensor([1.0, 15.0, 8.0], device=z.device) * (1 - (2 * inter + 1e-6) / (sp + st + 1e-6))).mean(1))
        tv = (inter + 1e-6) / (inter + 0.7 * (sp - inter) + 0.3 * (st - inter) + 1e-6)
        tvl = (torch.tensor([1.0, 12.0, 6.0], device=z.device) * (1 - tv)).mean(1)
        return (2.5 * focal + 2.5 * dice + tvl).mean()

    out = {"note": "comparator outside the product path: plain torch.nn modules on cuDNN / ATen, same architecture, batch, "
                   "resolution and step (fwd + loss + bwd + clip_grad_norm_ + AdamW); cudnn.benchmark on"}
    x, t = synth_batch(BATCH, RES, 1234, dev)
    prev = (torch.backends.cudnn.benchmark, torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32)
    torch.backends.cudnn.benchmark = True
    try:
        for name in ("fp32", "tf32", "bf16_autocast_channels_last"):
            torch.backends.cuda.matmul.allow_tf32 = torch.backends.cudnn.allow_tf32 = (name == "tf32")
            torch.manual_seed(0)
            m = StockUNet().to(dev).train()
            xi = x
            if name.startswith("bf16"):
                m = m.to(memory_format=torch.channels_last)
                xi = x.contiguous(memory_format=torch.channels_last)
            opt = torch.optim.AdamW(m.parameters(), lr=4e-3, weight_decay=1e-4, fused=True)

            def step():
                opt.zero_grad(set_to_none=True)
                with torch.autocast("cuda", dtype=torch.bfloat16, enabled=name.startswith("bf16")):
                    y = m(xi)
                loss = loss_fn(y, t)
                loss.backward()
                torch.nn.utils.clip_grad_norm_(m.parameters(), 1.0)
                opt.step()

            try:
                for _ in range(warmup):
                    step()
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(steps):
                    step()
                e1.record()
                torch.cuda.synchronize()
                ms = e0.elapsed_time(e1) / steps
                out[name] = {"ms_per_step": round(ms, 3), "images_per_s": round(BATCH / ms * 1e3, 1)}
            except Exception as e:  # noqa: BLE001 - a comparator failure must not take the product's line down
                out[name] = {"error": str(e).splitlines()[0][:200]}
            del m, opt
            torch.cuda.empty_cache()
    finally:
        torch.backends.cudnn.benchmark, torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = prev
    return out


# ---------------------------------------------------------------------------------------------
# inference configs (BASELINE configs[3] / [4]) and the metric kernels
# ---------------------------------------------------------------------------------------------
def inference_block(model, dev, world: int, rank: int, peaks: dict, reps: int = 3):
    """configs[3]: 32 x 3x1024^2 in total (sharded over the ranks); configs[4]: 8 x 3x2048^2 PER GPU.  Pipeline per batch:
    eval forward (16-bit tensor-core path, fused 2Hx2W tail) -> 2x2-mean + softmax -> probability->mask cascade ->
    confusion counts against a synthetic uint8 ground truth; inputs resident, single view (no TTA)."""
    import torch.distributed as dist
    from enhanced_unet_b200 import parallel
    from enhanced_unet_b200.ops import confusion_counts
    from enhanced_unet_b200.train_eval import Evaluator
    model.eval()
    ev = Evaluator(model, dev, "enhanced_unet", tta=False)
    peak_tf = peaks.get("bf16_tflops_sustained") or 1400.0
    hbm_peak = peaks.get("hbm_gbs") or 6650.0
    out = {}

    def timed(fn, n):
        fn()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
        if world > 1:
            tms = torch.tensor([ms], device=dev, dtype=torch.float64)
            dist.all_reduce(tms, op=dist.ReduceOp.MAX)
            ms = float(tms)
        return ms

    for key, total, res, sharded in (("config4_32x1024", 32, 1024, True), ("config5_8x2048_per_gpu", 8 * world, 2048, False)):
        lo, hi = parallel.shard_batch(total, rank, world) if sharded else (0, 8)
        b = hi - lo
        g = torch.Generator(device=dev).manual_seed(4321 + rank)
        x = torch.rand(b, 1, res, res, device=dev, generator=g).expand(-1, 3, -1, -1).contiguous()
        gt = torch.randint(0, 3, (b, res, res), device=dev, generator=g, dtype=torch.uint8)
        sub = 8 if res == 1024 else 2          # images per forward call: bounds the activation working set, whole batch timed

        def run():
            cms = []
            with torch.no_grad():
                for i in range(0, b, sub):
                    probs = ev._probs(x[i:i + sub])
                    masks = ev._convert_probs_to_mask_device(probs)
                    cms.append(confusion_counts(masks, gt[i:i + sub]))
            return torch.cat(cms)

        ms = timed(run, reps)
        cm = run()
        assert int(cm.sum()) == b * res * res
        gflop = conv_flops_fwd(total, res) / 1e9
        out[key] = {"images_per_s": round(total / ms * 1e3, 1), "ms_per_batch": round(ms, 2), "images": total, "resolution": res,
                    "images_per_forward_call": sub, "conv_tflops": round(gflop / ms, 1), "frac_of_tensor_peak": round(gflop / ms / (peak_tf * world), 3)}
        del x, gt
        torch.cuda.empty_cache()
    model.check_numerics()
    if rank == 0:
        # the metric kernel alone: 32 x 1024^2 masks, int64 / int64 (16 B per pixel, the dtypes the reference passes to
        # metrics.calculate_semantic_metrics) and uint8 / uint8 (2 B per pixel, the in-pipeline form)
        g = torch.Generator(device=dev).manual_seed(7)
        pred8 = torch.randint(0, 3, (32, 1024, 1024), device=dev, dtype=torch.uint8, generator=g)
        gt8 = torch.randint(0, 3, (32, 1024, 1024), device=dev, dtype=torch.uint8, generator=g)
        for name, p_, g_, bpp in (("uint8_2B_per_px", pred8, gt8, 2), ("int64_16B_per_px", pred8.long(), gt8.long(), 16)):
            confusion_counts(p_, g_)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10):
                cm = confusion_counts(p_, g_)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 10
            gbs = p_.numel() * bpp / (ms / 1e3) / 1e9
            out["confusion_kernel_" + name] = {"GB/s": round(gbs, 1), "frac_of_hbm_peak": round(gbs / hbm_peak, 3), "ms": round(ms, 4),
                                               "Mpx": p_.numel() / 1e6, "note": "includes the zero-fill of the 32x16 int64 counters"}
    model.train()
    return out


def traffic_from_profile():
    """dram bytes per launch of the dominant kernel from the committed ncu capture (profiles/conv_traffic.json, written by
    scripts/ncu_traffic.py).  Stale captures (kernel source changed since) are reported as such, never silently reused."""
    path = os.path.join(ROOT, "profiles", "conv_traffic.json")
    try:
        rec = json.load(open(path))
    except Exception:
        return None, "no profiles/conv_traffic.json"
    src = os.path.join(ROOT, "enhanced_unet_b200", "csrc", "conv_halo.cu")
    sha = hashlib.sha256(open(src, "rb").read()).hexdigest()[:16]
    if rec.get("conv_halo_cu_sha16") != sha:
        sys.stderr.write(f"bench: profiles/conv_traffic.json was captured for conv_halo.cu {rec.get('conv_halo_cu_sha16')} but the "
                         f"source is now {sha}: traffic reported as null (re-run scripts/ncu_traffic.py)\n")
        return None, "stale capture (conv_halo.cu changed since profiles/conv_traffic.json)"
    return rec.get("dram_bytes_per_launch_mean"), rec.get("source")


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
    model = EnhancedUNet(3, dtype=args.dtype).to(dev).train()
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

    def maxrank(ms):
        if world > 1:
            tms = torch.tensor([ms], device=dev, dtype=torch.float64)
            dist.all_reduce(tms, op=dist.ReduceOp.MAX)
            ms = float(tms)
        return ms

    def timed(fn, n):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        barrier()
        return maxrank(e0.elapsed_time(e1))

    # One GPU: the whole step is ONE CUDA-graph launch (graph.GraphedTrainStep: the public API for a fixed-shape training
    # loop).  N > 1: the eager step - the NCCL exchange launched from inside backward is not captured.
    gstep = None
    run_step = step
    if world == 1 and not args.no_graph:
        from enhanced_unet_b200.graph import GraphedTrainStep
        gstep = GraphedTrainStep(model, opt, BATCH, RES, RES, warmup_steps=2, example=(x_dev, t_dev))
        run_step = gstep
    for _ in range(args.warmup):
        run_step(x_dev, t_dev)
    lib.COUNTERS.clear()
    clk = ClockSampler(local)
    if rank == 0:          # one nvidia-smi poller per job, not one per rank
        with clk:
            ms = timed(lambda: run_step(x_dev, t_dev), args.steps)
    else:
        ms = timed(lambda: run_step(x_dev, t_dev), args.steps)
    launches = gstep.launches_per_step * args.steps if gstep is not None else sum(lib.COUNTERS.values())
    value = world * BATCH * args.steps / (ms / 1e3)
    model.check_numerics()

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
            float(run_step(x, t).detach())

    e2e_run(min(2, max(1, args.warmup)))
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    e2e_run(args.steps)
    e1.record()
    barrier()
    ms_e2e = maxrank(e0.elapsed_time(e1))
    e2e_value = world * BATCH * args.steps / (ms_e2e / 1e3)

    long_run = None
    if world > 1 and not args.quick:
        # a 0.5 s timed region is short for an N-GPU number: the same loop over 60 steps (reported beside, not instead of, K)
        ms60 = timed(lambda: step(x_dev, t_dev), 60)
        long_run = {"steps": 60, "ms_per_step": ms60 / 60, "images_per_s": world * BATCH * 60 / (ms60 / 1e3)}

    # data-parallel correctness on the hardware: after all those steps every rank must hold bit-identical parameters, and
    # the reduced gradient of the last step must be bit-identical on every rank (NCCL all-reduce delivers the same bits)
    dp_check = None
    if world > 1:
        flat_p = torch.cat([p.detach().reshape(-1) for p in params]).view(torch.int32).to(torch.int64)
        flat_g = allreduce.buffer.flat.view(torch.int32).to(torch.int64)
        sums = torch.stack([flat_p.sum(), (flat_p * (torch.arange(flat_p.numel(), device=dev) % 8191 + 1)).sum(),
                            flat_g.sum(), (flat_g * (torch.arange(flat_g.numel(), device=dev) % 8191 + 1)).sum()])
        allsums = [torch.empty_like(sums) for _ in range(world)]
        dist.all_gather(allsums, sums)
        same_p = all(torch.equal(a[:2], allsums[0][:2]) for a in allsums)
        same_g = all(torch.equal(a[2:], allsums[0][2:]) for a in allsums)
        dp_check = {"params_bit_identical_across_ranks": bool(same_p), "reduced_grads_bit_identical_across_ranks": bool(same_g),
                    "ranks": world, "buckets": len(allreduce.buffer.buckets),
                    "exposed_bucket_bytes": 4 * (allreduce.buffer.buckets[-1][1] - allreduce.buffer.buckets[-1][0])}
        if not (same_p and same_g):
            raise RuntimeError(f"data-parallel replicas diverged: {dp_check}")

    roof = None
    cpu = None
    # per-kernel profile: every C-ABI call bracketed by CUDA events on the launching stream, two steps.
    # Every rank runs the instrumented steps (they contain the collective); rank 0 reports.
    lib.PROFILE = []
    for _ in range(2):
        step(x_dev, t_dev)
    prof = lib.collect_profile(by_tag=True)
    lib.PROFILE = None
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    if rank == 0:
        peak = peaks.get("bf16_tflops_sustained") or 1400.0
        which = "measured sustained (MEASURED_PEAKS.json)" if "bf16_tflops_sustained" in peaks else "fallback 1.4 PF sustained"
        hbm_peak = peaks.get("hbm_gbs") or 6650.0

        def grp(name, tag):
            return prof.get((name, tag), (0.0, 0.0, 0))

        fl, msk, n = grp("eunet_conv3x3_fwd", "tc")
        if msk > 0:
            ach = fl / (msk / 1e3) / 1e12
            wfl, wms, wn = grp("eunet_conv3x3_wgrad", "tc")
            allc_fl = sum(v[0] for v in prof.values())
            allc_ms = sum(v[1] for k, v in prof.items() if v[0] > 0)
            hb = hbm_bytes_per_step(BATCH, RES)
            by_name = {}
            for (name, tag), v in prof.items():
                key = f"{name}[{tag}]" if tag == "k27" else name
                f0, m0, n0 = by_name.get(key, (0.0, 0.0, 0))
                by_name[key] = (f0 + v[0], m0 + v[1], n0 + v[2])
            hbm = {}
            for k_, b_ in hb.items():
                if k_ in by_name and by_name[k_][1] > 0:
                    gbs = b_ / (by_name[k_][1] / 2 / 1e3) / 1e9
                    hbm[k_] = {"GB/s": round(gbs, 1), "frac": round(gbs / hbm_peak, 3), "ms_per_step": round(by_name[k_][1] / 2, 4)}
            traffic, traffic_src = traffic_from_profile()
            roof = {"bound": "tensor",
                    "kernel": "tcgen05 implicit-GEMM 3x3 conv, forward + dgrad launches of the K >= 576 layers "
                              "(conv3x3_halo_kernel / conv3x3_fwd_tc_kernel); ALGORITHMIC FLOPs 2*M*N*K (SURVEY.md §8 layer table)",
                    "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak, "traffic": traffic,
                    "traffic_source": traffic_src, "peak_source": which,
                    "launches_per_step": n // 2, "kernel_ms_per_step": msk / 2,
                    "wgrad": {"tflops": (wfl / (wms / 1e3) / 1e12) if wms > 0 else None, "frac": (wfl / (wms / 1e3) / 1e12 / peak) if wms > 0 else None,
                              "launches_per_step": wn // 2, "kernel_ms_per_step": wms / 2},
                    "all_conv_launches": {"tflops": allc_fl / (allc_ms / 1e3) / 1e12, "frac": allc_fl / (allc_ms / 1e3) / 1e12 / peak,
                                          "kernel_ms_per_step": allc_ms / 2},
                    "whole_step": {"conv_tflop_per_step": conv_flops_train(BATCH, RES) / 1e12,
                                   "tflops": conv_flops_train(BATCH, RES) / 1e12 / (ms / args.steps / 1e3),
                                   "frac": conv_flops_train(BATCH, RES) / 1e12 / (ms / args.steps / 1e3) / peak},
                    "hbm_kernels": hbm, "hbm_peak_gbs": hbm_peak,
                    "all_kernels_ms_per_step": {k: round(v[1] / 2, 4) for k, v in sorted(by_name.items(), key=lambda kv: -kv[1][1])}}
    # the optimiser states / saved activations of training are not needed any more
    del pf
    torch.cuda.empty_cache()
    infer = None
    if not args.quick:
        infer = inference_block(model, dev, world, rank, peaks)
    stock = None
    if rank == 0 and world == 1 and not args.quick and not args.no_stock:
        del model, opt, allreduce, params
        torch.cuda.empty_cache()
        stock = stock_pytorch_block(dev)
    if rank == 0:
        if world == 1 and not args.no_cpu and not args.quick:
            r = run_cpu(3, 1)
            cpu = {"value": r["img512_per_s"], "unit": "images/s", "cores": r["cores"], "kind": "port", "sample": r["sample"]}
        dtype_label = {"fp16": "fp16 (16-bit tcgen05 kind::f16 operands: fp16 activations / filters / gradients, fp32 accumulate and "
                               "statistics; same tensor-core rate and bytes as bf16, 3 more mantissa bits)",
                       "bf16": "bf16", "fp32": "fp32"}[args.dtype]
        line = {"metric": METRIC, "value": value, "unit": "images/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": dtype_label,
                "data": "synthetic",
                "config": {"workload": "Enhanced-UNet 16-bit tensor-core training, batch 16/GPU, 3x512x512 (BASELINE configs[1]/[2]): "
                                       "fwd + focal/dice/tversky loss + bwd + clip + AdamW",
                           "global_batch": BATCH * world, "resolution": RES, "parallelism": f"dp{world}",
                           "l2": "working set (>5 GB of activations per step) far exceeds the 126 MB L2; no flush needed",
                           "launch": ("one CUDA graph per step (graph.GraphedTrainStep), %d kernels-launching C-ABI calls captured" % gstep.launches_per_step)
                                     if gstep is not None else "eager: one C-ABI call per kernel",
                           "conv_tflop_per_step": conv_flops_train(BATCH, RES) / 1e12},
                "clocks": clk.summary(),
                "e2e": {"value": e2e_value, "unit": "images/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4},
                "gpu_launches": launches, "roofline": roof, "cpu_baseline": cpu, "inference": infer, "stock_pytorch_b200": stock,
                "dp_check": dp_check, "long_run": long_run}
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
    ap.add_argument("--dtype", default="fp16", choices=["fp16", "bf16", "fp32"], help="compute mode of the product path")
    ap.add_argument("--no-cpu", action="store_true", help="skip the bounded CPU baseline leg")
    ap.add_argument("--no-stock", action="store_true", help="skip the stock-PyTorch comparator block")
    ap.add_argument("--no-graph", action="store_true", help="one GPU: launch the step kernel by kernel instead of as one CUDA graph")
    ap.add_argument("--quick", action="store_true", help="training line only (no inference / comparator / CPU blocks): profiling runs")
    args = ap.parse_args()
    if args.impl == "reference":
        main_reference(args)
    else:
        main_gpu(args)


if __name__ == "__main__":
    main()
