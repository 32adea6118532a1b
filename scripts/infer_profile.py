#!/usr/bin/env python
"""Per-entry-point time of the inference pipeline (BASELINE configs[3]: 1024^2 tiles; eval forward -> softmax -> mask cascade ->
confusion counts), CUDA events around every C-ABI call.  Usage: python scripts/infer_profile.py [--batch 8] [--res 1024]"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from enhanced_unet_b200 import lib
from enhanced_unet_b200.models import EnhancedUNet
from enhanced_unet_b200.ops import confusion_counts
from enhanced_unet_b200.train_eval import Evaluator

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=8)
ap.add_argument("--res", type=int, default=1024)
args = ap.parse_args()
torch.manual_seed(0)
m = EnhancedUNet(3).cuda().eval()
ev = Evaluator(m, "cuda", "enhanced_unet", tta=False)
g = torch.Generator(device="cuda").manual_seed(1)
x = torch.rand(args.batch, 1, args.res, args.res, device="cuda", generator=g).expand(-1, 3, -1, -1).contiguous()
gt = torch.randint(0, 3, (args.batch, args.res, args.res), device="cuda", generator=g, dtype=torch.uint8)


def run():
    with torch.no_grad():
        probs = ev._probs(x)
        masks = ev._convert_probs_to_mask_device(probs)
        return confusion_counts(masks, gt)


run(); run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3):
    run()
e1.record()
torch.cuda.synchronize()
total = e0.elapsed_time(e1) / 3
lib.PROFILE = []
run()
prof = lib.collect_profile(by_tag=True)
lib.PROFILE = None
print(f"batch {args.batch} x {args.res}^2: {total:.2f} ms per batch = {args.batch / total * 1e3:.1f} images/s; per entry point (ms, launches):")
for (name, tag), (fl, ms, n) in sorted(prof.items(), key=lambda kv: -kv[1][1]):
    print(f"  {name + ('[' + tag + ']' if tag else ''):36s} {ms:8.3f} {n:4d}" + (f"   {fl / ms / 1e9:7.0f} TFLOP/s" if fl else ""))
print(f"  sum {sum(v[1] for v in prof.values()):.3f} ms")
