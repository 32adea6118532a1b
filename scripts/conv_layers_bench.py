#!/usr/bin/env python
"""Per-layer timing of the conv kernels at the BASELINE config-2 shapes (batch 16, 512x512): forward, dgrad, wgrad,
CUDA events; A/B over a library option (default: the CTA-pair kernel, `cta_pair` = 0 / 1 / 2).
Usage: python scripts/conv_layers_bench.py [--batch 16] [--res 512] [--option cta_pair --values 0,1,2]"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from enhanced_unet_b200 import lib

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=16)
ap.add_argument("--res", type=int, default=512)
ap.add_argument("--reps", type=int, default=5)
ap.add_argument("--only", default="", help="comma-separated layer names")
ap.add_argument("--option", default="cta_pair")
ap.add_argument("--values", default="0,1,2")
ap.add_argument("--default", type=int, default=1)
args = ap.parse_args()
B, R = args.batch, args.res
vals = [int(v) for v in args.values.split(",")]
LAYERS = [("enc1.0", 16, 64, 0), ("enc1.3", 64, 64, 0), ("enc2.0", 64, 128, 1), ("enc2.3", 128, 128, 1), ("enc3.0", 128, 256, 2),
          ("enc3.3", 256, 256, 2), ("enc4.0", 256, 512, 3), ("enc4.3", 512, 512, 3), ("dec4.0", 768, 256, 2), ("dec4.3", 256, 256, 2),
          ("dec3.0", 384, 128, 1), ("dec3.3", 128, 128, 1), ("dec2.0", 192, 64, 0), ("dec2.3", 64, 64, 0), ("enh.0", 16, 64, -1)]
dt, code = torch.float16, lib.F16


def timeit(fn):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / args.reps


tot = {}
print(f"{'layer':8} {'pass':6} {'GFLOP':>8} | " + " | ".join(f"{args.option}={v}: ms   TF/s" for v in vals))
for name, cin, cout, lvl in LAYERS:
    if args.only and name not in args.only.split(","):
        continue
    h = (R >> lvl) if lvl >= 0 else 2 * R
    M = B * h * h
    real_cin = 3 if cin == 16 else cin
    x = torch.randn(M, cin, device="cuda").to(dt)
    dy = torch.randn(M, cout, device="cuda").to(dt)
    w = torch.randn(cout, 9, cin, device="cuda").to(dt)
    wt = torch.randn(cin, 9, cout, device="cuda").to(dt)
    y = torch.empty(M, cout, device="cuda", dtype=torch.float16)
    dx = torch.empty(M, cin, device="cuda", dtype=dt)
    dw = torch.zeros(cout, 9, cin, device="cuda")
    stats = torch.zeros(2 * cout, device="cuda", dtype=torch.float64)
    fl = 2.0 * M * cout * 9 * real_cin
    passes = [("fwd", lambda: lib.call("eunet_conv3x3_fwd", x.data_ptr(), cin, w.data_ptr(), y.data_ptr(), cout, code, B, h, h, cin, cout,
                                         stats.data_ptr(), None, None, 0, 1, None))]
    if name != "enc1.0":
        passes.append(("dgrad", lambda: lib.call("eunet_conv3x3_fwd", dy.data_ptr(), cout, wt.data_ptr(), dx.data_ptr(), cin, code, B, h, h,
                                                 cout, cin, None, None, None, 0, 0, None)))
    passes.append(("wgrad", lambda: lib.call("eunet_conv3x3_wgrad", x.data_ptr(), cin, dy.data_ptr(), cout, dw.data_ptr(), code, B, h, h,
                                             cin, cout)))
    for pname, fn in passes:
        res = []
        for v in vals:
            lib.set_option(args.option, v)
            res.append(timeit(fn))
        lib.set_option(args.option, args.default)
        for v, r in zip(vals, res):
            tot[(pname, v)] = tot.get((pname, v), 0.0) + r
        print(f"{name:8} {pname:6} {fl / 1e9:8.1f} | " + " | ".join(f"{r:12.3f} {fl / r / 1e9:7.0f}" for r in res))
    del x, dy, y, dx
for v in vals:
    print(f"totals {args.option}={v} (ms):", {p: round(t, 3) for (p, vv), t in tot.items() if vv == v},
          "sum", round(sum(t for (p, vv), t in tot.items() if vv == v), 3))
