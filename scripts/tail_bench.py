#!/usr/bin/env python
"""Timing of the 2Hx2W tail backward kernels at the BASELINE config-2 shape (batch 16, 1024x1024 output grid), CUDA events."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from enhanced_unet_b200 import lib

B, H2 = 16, 1024
M = B * H2 * H2
dt, code = torch.float16, lib.F16
g = torch.Generator(device="cuda").manual_seed(0)
mid = (torch.randn(M, 64, device="cuda", generator=g) * 1.5 + 0.3).to(torch.float16)
dout4 = torch.randn(M, 4, device="cuda", generator=g)
d1p = torch.zeros(M, 16, dtype=dt, device="cuda")
d1p[:, :3] = torch.randn(M, 3, device="cuda", generator=g).to(dt)
w0 = torch.randn(64, 3, 3, 3, device="cuda", generator=g) / 5
wflip = torch.empty(16, 9, 64, dtype=dt, device="cuda")
lib.call("eunet_pack_weight3x3", w0.data_ptr(), wflip.data_ptr(), code, 64, 3, 64, 16, 1)
mean = torch.full((64,), 0.3, device="cuda")
invstd = torch.full((64,), 1 / 1.5, device="cuda")
scale, shift = invstd.clone(), (-mean * invstd).contiguous()
w3 = (torch.randn(3, 64, device="cuda", generator=g) / 8).contiguous()
acc = torch.zeros(328, dtype=torch.float64, device="cuda")
dx = torch.empty(M, 4, device="cuda")
dw = torch.zeros(64, 9, 16, device="cuda")


def timeit(fn, reps=5):
    fn(); fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


t_red = timeit(lambda: lib.call("eunet_tail_bwd_reduce", dout4.data_ptr(), mid.data_ptr(), code, scale.data_ptr(), shift.data_ptr(),
                                mean.data_ptr(), invstd.data_ptr(), w3.data_ptr(), acc.data_ptr(), B, H2 // 2, H2 // 2))
t_fused = timeit(lambda: lib.call("eunet_tail_bwd_fused", dout4.data_ptr(), mid.data_ptr(), d1p.data_ptr(), wflip.data_ptr(), scale.data_ptr(),
                                  shift.data_ptr(), mean.data_ptr(), invstd.data_ptr(), w3.data_ptr(), acc.data_ptr(), dx.data_ptr(),
                                  dw.data_ptr(), code, B, H2, H2))
for dbg in (16, 1, 2, 4, 8, 6, 7, 15, 9, 14):
    lib.set_option("tail_dbg", dbg)
    t = timeit(lambda: lib.call("eunet_tail_bwd_fused", dout4.data_ptr(), mid.data_ptr(), d1p.data_ptr(), wflip.data_ptr(), scale.data_ptr(),
                                shift.data_ptr(), mean.data_ptr(), invstd.data_ptr(), w3.data_ptr(), acc.data_ptr(), dx.data_ptr(),
                                dw.data_ptr(), code, B, H2, H2))
    print(f"  tail_dbg={dbg:2d} (16 fp32 transform, 1 no transform, 2 no wgrad MMA, 4 no U^T MMA, 8 no drain): {t:.3f} ms")
lib.set_option("tail_dbg", 0)
gb = M * (128 + 16 + 32 + 16) / 1e9
print(f"tail_bwd_reduce {t_red:.3f} ms; tail_bwd_fused {t_fused:.3f} ms = {gb / t_fused * 1e3:.0f} GB/s of algorithmic traffic")
