"""Python driver for the UMMA/TMA probe kernel (csrc/probe.cu): builds the parameter blob, tensor maps are
encoded by the library.  Descriptor encoders mirror csrc/tc_common.cuh so the tests pin those encodings."""
from __future__ import annotations

import ctypes as C
import struct
from typing import List, Sequence, Tuple

import numpy as np
import torch

from enhanced_unet_b200 import lib

SW_NONE, SW128, SW64, SW32 = 0, 2, 4, 6


def smem_desc(addr: int, lbo: int, sbo: int, layout: int, base_offset: int = 0) -> int:
    return (((addr >> 4) & 0x3FFF) | (((lbo >> 4) & 0x3FFF) << 16) | (((sbo >> 4) & 0x3FFF) << 32) | (1 << 46)
            | ((base_offset & 7) << 49) | ((layout & 7) << 61))


def idesc_bf16(m: int, n: int, a_mn: int, b_mn: int) -> int:
    return (1 << 4) | (1 << 7) | (1 << 10) | (a_mn << 15) | (b_mn << 16) | ((n >> 3) << 17) | ((m >> 4) << 24)


def run_probe(a: torch.Tensor, a_box: Tuple[int, int], a_sw: int, b: torch.Tensor, b_box: Tuple[int, int], b_sw: int,
              x4: torch.Tensor, x_box: Sequence[int], x_sw: int, loads: List[Tuple[int, Sequence[int], int]], tx_bytes: int,
              mmas: List[Tuple[int, int, int, int, int]], ncols: int, smem_bytes: int, dump_bytes: int):
    """a, b: 2-D bf16 CUDA tensors [rows, cols]; x4: bf16 NHWC [B,H,W,C].  loads: (map, coords, smem_off);
    mmas: (adesc, bdesc, idesc, accumulate, tmem_col).  Returns (tmem fp32 [128, ncols], smem uint8 [dump_bytes])."""
    assert a.dtype == torch.bfloat16 and b.dtype == torch.bfloat16 and x4.dtype == torch.bfloat16
    blob = struct.pack("<iiIii3i", len(loads), len(mmas), tx_bytes, ncols, dump_bytes, 0, 0, 0)
    for i in range(24):
        if i < len(loads):
            m, c, off = loads[i]
            c = list(c) + [0] * (4 - len(c))
            blob += struct.pack("<i4iI", m, *c, off)
        else:
            blob += struct.pack("<i4iI", 0, 0, 0, 0, 0, 0)
    for i in range(32):
        if i < len(mmas):
            ad, bd, idc, acc, col = mmas[i]
            blob += struct.pack("<QQIIII", ad, bd, idc, acc, col, 0)
        else:
            blob += struct.pack("<QQIIII", 0, 0, 0, 0, 0, 0)
    assert len(blob) == 32 + 24 * 24 + 32 * 32
    out_t = torch.zeros(128, ncols, device="cuda", dtype=torch.float32)
    out_s = torch.zeros(max(dump_bytes, 16), device="cuda", dtype=torch.uint8)
    Bx, Hx, Wx, Cx = x4.shape
    xd = (C.c_int * 4)(Cx, Wx, Hx, Bx)
    xb = (C.c_int * 4)(*x_box)
    buf = C.create_string_buffer(blob, len(blob))
    lib.call("eunet_probe_umma", a.data_ptr(), a.shape[0], a.shape[1], a_box[0], a_box[1], a_sw,
             b.data_ptr(), b.shape[0], b.shape[1], b_box[0], b_box[1], b_sw,
             x4.data_ptr(), C.cast(xd, C.c_void_p), C.cast(xb, C.c_void_p), x_sw,
             C.cast(buf, C.c_void_p), len(blob), out_t.data_ptr(), out_s.data_ptr(), smem_bytes)
    torch.cuda.synchronize()
    return out_t.cpu().numpy(), out_s.cpu().numpy()


def rand_bf16(*shape, seed=0, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    t = (torch.randn(*shape, generator=g) * scale).to(torch.bfloat16)
    return t


def f32(t: torch.Tensor) -> np.ndarray:
    return t.float().cpu().numpy()
