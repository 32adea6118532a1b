#!/usr/bin/env python
"""Inference throughput at the BASELINE high-res configs (configs[3] / [4]): eval-mode bf16 forward -> 2x2-mean + softmax ->
probability->mask cascade -> confusion counts, CUDA events, inputs resident.  A/B of the fused 2Hx2W tail epilogue.
Usage: python scripts/infer_bench.py [--batch 8] [--res 1024] [--reps 3]"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from enhanced_unet_b200 import engine
from enhanced_unet_b200.models import EnhancedUNet
from enhanced_unet_b200.ops import confusion_counts
from enhanced_unet_b200.train_eval import Evaluator

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=8)
ap.add_argument("--res", type=int, default=1024)
ap.add_argument("--reps", type=int, default=3)
args = ap.parse_args()
torch.manual_seed(0)
m = EnhancedUNet(3).cuda().eval()
ev = Evaluator(m, "cuda", "enhanced_unet")
g = torch.Generator(device="cuda").manual_seed(1)
x = torch.rand(args.batch, 1, args.res, args.res, device="cuda", generator=g).expand(-1, 3, -1, -1).contiguous()
gt = torch.randint(0, 3, (args.batch, args.res, args.res), device="cuda", generator=g, dtype=torch.uint8)


def run():
    probs = ev._probs(x)
    masks = ev._convert_probs_to_mask_device(probs)
    return confusion_counts(masks, gt)


for fused in (True, False):
    engine.FUSED_TAIL = fused
    with torch.no_grad():
        run(); run()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.reps):
            cm = run()
        e1.record()
        torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.reps
    print(f"fused_tail={fused}: batch {args.batch} x {args.res}^2: {ms:.2f} ms -> {args.batch / ms * 1e3:.1f} images/s "
          f"(peak mem {torch.cuda.max_memory_allocated() / 2**30:.1f} GiB, counts sum {int(cm.sum())})")
engine.FUSED_TAIL = None
