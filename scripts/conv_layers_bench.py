#!/usr/bin/env python
"""Per-layer timing of the conv kernels at the BASELINE config-2 shapes (batch 16, 512x512): forward, dgrad, wgrad,
CUDA events, halo kernels on / off.  Usage: python scripts/conv_layers_bench.py [--batch 16] [--res 512]"""
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
ap.add_argument("--halo", type=int, default=1, help="value of the conv_halo option for the first column")
args = ap.parse_args()
B, R = args.batch, args.res
LAYERS = [("enc1.0", 16, 64, 0), ("enc1.3", 64, 64, 0), ("enc2.0", 64, 128, 1), ("enc2.3", 128, 128, 1), ("enc3.0", 128, 256, 2),
          ("enc3.3", 256, 256, 2), ("enc4.0", 256, 512, 3), ("enc4.3", 512, 512, 3), ("dec4.0", 768, 256, 2), ("dec4.3", 256, 256, 2),
          ("dec3.0", 384, 128, 1), ("dec3.3", 128, 128, 1), ("dec2.0", 192, 64, 0), ("dec2.3", 64, 64, 0), ("enh.0", 16, 64, -1)]
bf = torch.bfloat16


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
print(f"{'layer':8} {'pass':6} {'GFLOP':>8} | {'halo ms':>8} {'TF/s':>7} | {'pertap ms':>9} {'TF/s':>7}")
for name, cin, cout, lvl in LAYERS:
    if args.only and name not in args.only.split(","):
        continue
    h = (R >> lvl) if lvl >= 0 else 2 * R
    M = B * h * h
    real_cin = 3 if cin == 16 else cin
    x = torch.randn(M, cin, device="cuda").to(bf)
    dy = torch.randn(M, cout, device="cuda").to(bf)
    w = torch.randn(cout, 9, cin, device="cuda").to(bf)
    wt = torch.randn(cin, 9, cout, device="cuda").to(bf)
    y = torch.empty(M, cout, device="cuda", dtype=torch.float16)
    dx = torch.empty(M, cin, device="cuda", dtype=bf)
    dw = torch.zeros(cout, 9, cin, device="cuda")
    stats = torch.zeros(2 * cout, device="cuda", dtype=torch.float64)
    fl = 2.0 * M * cout * 9 * real_cin
    passes = [("fwd", lambda: lib.call("eunet_conv3x3_fwd", x.data_ptr(), cin, w.data_ptr(), y.data_ptr(), cout, 1, B, h, h, cin, cout,
                                         stats.data_ptr(), None, None, 0, 1))]
    if name != "enc1.0":
        passes.append(("dgrad", lambda: lib.call("eunet_conv3x3_fwd", dy.data_ptr(), cout, wt.data_ptr(), dx.data_ptr(), cin, 1, B, h, h,
                                                 cout, cin, None, None, None, 0, 0)))
    passes.append(("wgrad", lambda: lib.call("eunet_conv3x3_wgrad", x.data_ptr(), cin, dy.data_ptr(), cout, dw.data_ptr(), 1, B, h, h,
                                             cin, cout)))
    for pname, fn in passes:
        res = []
        for halo in (args.halo, 0):
            lib.set_option("conv_halo", halo)
            res.append(timeit(fn))
        lib.set_option("conv_halo", 1)
        tot[pname] = tot.get(pname, 0.0) + res[0]
        print(f"{name:8} {pname:6} {fl / 1e9:8.1f} | {res[0]:8.3f} {fl / res[0] / 1e9:7.0f} | {res[1]:9.3f} {fl / res[1] / 1e9:7.0f}")
    del x, dy, y, dx
print("totals (halo path, ms):", {k: round(v, 3) for k, v in tot.items()}, "sum", round(sum(tot.values()), 3))
