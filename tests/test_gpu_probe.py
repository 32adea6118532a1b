"""Pins the tcgen05 shared-memory / instruction descriptor encodings and the TMA box layouts used by
csrc/conv_tc.cu against numpy matmuls, through the probe kernel (csrc/probe.cu).  Each experiment also
records alternatives in gpurun_out/probe_report.json so a failed hypothesis tells us what WOULD work."""
import json
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

REPORT = {}


def _save():
    out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
    os.makedirs(out, exist_ok=True)
    with open(os.path.join(out, "probe_report.json"), "w") as f:
        json.dump(REPORT, f, indent=1, sort_keys=True)


def _err(got, want):
    return float(np.abs(got - want).max() / (np.abs(want).max() + 1e-9))


@pytest.fixture(scope="module")
def pu():
    import importlib.util
    here = os.path.dirname(os.path.abspath(__file__))
    spec = importlib.util.spec_from_file_location("eunet_probe_util", os.path.join(here, "probe_util.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def _dummy_x():
    return torch.zeros(1, 8, 8, 64, device="cuda", dtype=torch.bfloat16)


def test_kmajor_sw128(pu):
    """Forward conv operand layout: rows of 64 bf16 (128 B), 128B swizzle, SBO = 1024, K advance = +32 B."""
    a = pu.rand_bf16(128, 64, seed=1).cuda()
    b = pu.rand_bf16(64, 64, seed=2).cuda()
    want = pu.f32(a) @ pu.f32(b).T
    loads = [(0, (0, 0), 0), (1, (0, 0), 16384)]
    res = {}
    for lbo in (16, 0, 1024):
        mmas = [(pu.smem_desc(0 + 32 * j, lbo, 1024, pu.SW128), pu.smem_desc(16384 + 32 * j, lbo, 1024, pu.SW128),
                 pu.idesc_bf16(128, 64, 0, 0), int(j > 0), 0) for j in range(4)]
        t, s = pu.run_probe(a, (128, 64), 128, b, (64, 64), 128, _dummy_x(), (64, 8, 8, 1), 128, loads, 16384 + 8192, mmas, 64,
                            32768, 32768)
        res[f"lbo{lbo}"] = _err(t[:, :64], want)
    # the TMA 128B swizzle: 16-byte chunk c of row r lands at chunk (c ^ (r & 7))
    raw = s[:16384].view(np.uint16).reshape(128, 8, 8)
    src = a.cpu().view(torch.int16).numpy().view(np.uint16).reshape(128, 8, 8)
    unsw = np.stack([raw[r, [c ^ (r & 7) for c in range(8)]] for r in range(128)])
    res["tma_swizzle_ok"] = bool((unsw == src).all())
    REPORT["kmajor_sw128"] = res
    _save()
    assert res["lbo16"] < 1e-5, res
    assert res["tma_swizzle_ok"]


def test_kmajor_sw32(pu):
    """16-channel chunks (first layer / enhance head): rows of 16 bf16 (32 B), 32B swizzle, SBO = 256."""
    a = pu.rand_bf16(128, 16, seed=3).cuda()
    b = pu.rand_bf16(64, 16, seed=4).cuda()
    want = pu.f32(a) @ pu.f32(b).T
    loads = [(0, (0, 0), 0), (1, (0, 0), 4096)]
    res = {}
    for sbo in (256, 512):
        mmas = [(pu.smem_desc(0, 16, sbo, pu.SW32), pu.smem_desc(4096, 16, sbo, pu.SW32), pu.idesc_bf16(128, 64, 0, 0), 0, 0)]
        t, _ = pu.run_probe(a, (128, 16), 32, b, (64, 16), 32, _dummy_x(), (64, 8, 8, 1), 128, loads, 4096 + 2048, mmas, 64, 8192,
                            8192)
        res[f"sbo{sbo}"] = _err(t[:, :64], want)
    REPORT["kmajor_sw32"] = res
    _save()
    assert res["sbo256"] < 1e-5, res


def test_mnmajor_sw128(pu):
    """wgrad operand layout: tiles [64 pixels (K)][64 channels (MN)] exactly as TMA writes NHWC boxes,
    consumed MN-major.  A' = 2 blocks (128 channels), B' = 3 blocks (192), LBO = block stride 8192,
    SBO = 8-pixel group stride 1024, K advance = 16 rows = 2048 B."""
    at = pu.rand_bf16(64, 128, seed=5).cuda()     # [k, m]
    bt = pu.rand_bf16(64, 192, seed=6).cuda()     # [k, n]
    want = pu.f32(at).T @ pu.f32(bt)
    loads = [(0, (0, 0), 0), (0, (64, 0), 8192)] + [(1, (64 * i, 0), 16384 + 8192 * i) for i in range(3)]
    res = {}
    for name, (lbo, sbo) in {"lbo8192_sbo1024": (8192, 1024), "lbo1024_sbo8192": (1024, 8192)}.items():
        mmas = [(pu.smem_desc(0 + 2048 * j, lbo, sbo, pu.SW128), pu.smem_desc(16384 + 2048 * j, lbo, sbo, pu.SW128),
                 pu.idesc_bf16(128, 192, 1, 1), int(j > 0), 0) for j in range(4)]
        t, _ = pu.run_probe(at, (64, 64), 128, bt, (64, 64), 128, _dummy_x(), (64, 8, 8, 1), 128, loads, 5 * 8192, mmas, 256, 65536,
                            16)
        res[name] = _err(t[:, :192], want)
    REPORT["mnmajor_sw128"] = res
    _save()
    assert res["lbo8192_sbo1024"] < 1e-5, res


def test_mnmajor_sw32_b_operand(pu):
    """wgrad with 16-channel input chunks: B' = 9 tap tiles [64 px][16 ch] (32B swizzle), N = 144."""
    at = pu.rand_bf16(64, 128, seed=7).cuda()
    bt = pu.rand_bf16(64, 144, seed=8).cuda()
    want = pu.f32(at).T @ pu.f32(bt)
    loads = [(0, (0, 0), 0), (0, (64, 0), 8192)] + [(1, (16 * i, 0), 16384 + 2048 * i) for i in range(9)]
    res = {}
    for name, (lbo, sbo, kadv) in {"lbo2048_sbo256_k512": (2048, 256, 512), "lbo256_sbo2048_k512": (256, 2048, 512)}.items():
        mmas = [(pu.smem_desc(0 + 2048 * j, 8192, 1024, pu.SW128), pu.smem_desc(16384 + kadv * j, lbo, sbo, pu.SW32),
                 pu.idesc_bf16(128, 144, 1, 1), int(j > 0), 0) for j in range(4)]
        t, _ = pu.run_probe(at, (64, 64), 128, bt, (64, 16), 32, _dummy_x(), (64, 8, 8, 1), 128, loads, 16384 + 9 * 2048, mmas, 256,
                            65536, 16)
        res[name] = _err(t[:, :144], want)
    REPORT["mnmajor_sw32"] = res
    _save()
    assert res["lbo2048_sbo256_k512"] < 1e-5, res


def test_tma_4d_box_zero_fill_and_row_order(pu):
    """The conv A operand: one 4-D box {64 ch, bw, bh, bb} at a shifted origin; out-of-bounds pixels must
    read as zero and box row r must be pixel (bb, y, x) = (r / (bh*bw), (r / bw) % bh, r % bw)."""
    B, H, W, Cc = 2, 6, 10, 64
    x = pu.rand_bf16(B, H, W, Cc, seed=9).cuda()
    eye = torch.eye(64, dtype=torch.bfloat16).cuda()
    bw, bh, bb = 16, 4, 2
    x0, y0, b0 = -1, -1, 0     # tap (0,0) of the tile at the image origin
    loads = [(2, (0, x0, y0, b0), 0), (1, (0, 0), 16384)]
    mmas = [(pu.smem_desc(32 * j, 16, 1024, pu.SW128), pu.smem_desc(16384 + 32 * j, 16, 1024, pu.SW128),
             pu.idesc_bf16(128, 64, 0, 0), int(j > 0), 0) for j in range(4)]
    dummy = torch.zeros(128, 64, device="cuda", dtype=torch.bfloat16)
    t, _ = pu.run_probe(dummy, (128, 64), 128, eye, (64, 64), 128, x, (64, bw, bh, bb), 128, loads, 16384 + 8192, mmas, 64, 32768,
                        16)
    xf = pu.f32(x)
    want = np.zeros((128, 64), np.float32)
    for r in range(128):
        xx, yy, b_ = r % bw, (r // bw) % bh, r // (bw * bh)
        gx, gy, gb = x0 + xx, y0 + yy, b0 + b_
        if 0 <= gx < W and 0 <= gy < H and 0 <= gb < B:
            want[r] = xf[gb, gy, gx]
    err = _err(t[:, :64], want)
    REPORT["tma_4d"] = {"err": err}
    _save()
    assert err < 1e-6, err


def test_row_shifted_descriptor_start(pu):
    """Exploratory (feeds the halo-reuse optimisation): does a K-major SW128 A operand whose start address
    is shifted by r rows (r*128 B) read rows r..r+127?  Variants: base_offset = 0 / r.  Reported, not required."""
    a = pu.rand_bf16(256, 64, seed=10).cuda()
    b = pu.rand_bf16(64, 64, seed=11).cuda()
    loads = [(0, (0, 0), 0), (0, (0, 128), 16384), (1, (0, 0), 32768)]
    res = {}
    for r in (1, 2, 3, 8, 16, 19):
        want = pu.f32(a)[r:r + 128] @ pu.f32(b).T
        for bo_name, bo in (("bo0", 0), ("bor", r & 7)):
            mmas = [(pu.smem_desc(128 * r + 32 * j, 16, 1024, pu.SW128, bo), pu.smem_desc(32768 + 32 * j, 16, 1024, pu.SW128),
                     pu.idesc_bf16(128, 64, 0, 0), int(j > 0), 0) for j in range(4)]
            t, _ = pu.run_probe(a, (128, 64), 128, b, (64, 64), 128, _dummy_x(), (64, 8, 8, 1), 128, loads, 2 * 16384 + 8192, mmas,
                                64, 65536, 16)
            res[f"shift{r}_{bo_name}"] = _err(t[:, :64], want)
    REPORT["row_shift"] = res
    _save()
    assert res["shift8_bo0"] < 1e-5 and res["shift16_bo0"] < 1e-5, res   # whole swizzle atoms must work


def _halo_rows(dy, dx, bw_halo=10, th=16, tw=8):
    """row index inside an (th+2) x bw_halo halo tile of GEMM row r = ty*tw + tx for tap (dy, dx)"""
    return [(ty + dy) * bw_halo + (tx + dx) for ty in range(th) for tx in range(tw)]


@pytest.mark.parametrize("swz,kc", [(128, 64), (32, 16)])
def test_halo_tile_taps_by_descriptor_shift_kmajor(pu, swz, kc):
    """Halo-reuse scheme of the v2 forward kernel: ONE 18x10-pixel halo tile in smem serves all nine taps of a
    16x8 output tile.  GEMM row r = (ty, tx) reads halo row (ty+dy)*10 + (tx+dx): the eight pixels of an output
    row are one 8-row group, consecutive output rows are 10 halo rows apart -> SBO = 10 rows, start = (dy*10+dx) rows."""
    row_b = kc * 2
    a = pu.rand_bf16(180, kc, seed=20).cuda()
    b = pu.rand_bf16(64, kc, seed=21).cuda()
    layout = pu.SW128 if swz == 128 else pu.SW32
    a_bytes = ((180 * row_b + 1023) // 1024) * 1024
    loads = [(0, (0, 0), 0), (1, (0, 0), a_bytes)]
    res = {}
    for dy in range(3):
        for dx in range(3):
            want = pu.f32(a)[_halo_rows(dy, dx)] @ pu.f32(b).T
            mmas = [(pu.smem_desc((dy * 10 + dx) * row_b + 32 * j, 16, 10 * row_b, layout),
                     pu.smem_desc(a_bytes + 32 * j, 16, 8 * row_b, layout), pu.idesc_bf16(128, 64, 0, 0), int(j > 0), 0)
                    for j in range(kc // 16)]
            t, _ = pu.run_probe(a, (180, kc), swz, b, (64, kc), swz, _dummy_x(), (64, 8, 8, 1), 128, loads,
                                180 * row_b + 64 * row_b, mmas, 64, 65536, 16)
            res[f"dy{dy}dx{dx}"] = _err(t[:, :64], want)
    REPORT[f"halo_kmajor_sw{swz}"] = res
    _save()
    assert max(res.values()) < 1e-5, res


def test_halo_tile_taps_by_descriptor_shift_mnmajor(pu):
    """Same scheme for wgrad's B' operand (MN-major: K = pixels): K=16 per MMA = two output rows of 8 pixels =
    two 8-row groups 10 halo rows apart (SBO = 1280 B); K advance per MMA = 2 output rows = 20 halo rows."""
    xh = pu.rand_bf16(180, 64, seed=22).cuda()        # halo tile [180 px][64 ci]
    dyt = pu.rand_bf16(128, 128, seed=23).cuda()      # dY tile [128 px (16x8)][128 co]
    a_off, b_off = 0, 32768
    loads = [(0, (0, 0), a_off), (0, (64, 0), a_off + 16384), (1, (0, 0), b_off)]
    res = {}
    for dy in range(3):
        for dx in range(3):
            want = pu.f32(dyt).T @ pu.f32(xh)[_halo_rows(dy, dx)]          # [128 co, 64 ci]
            mmas = [(pu.smem_desc(a_off + 2048 * j, 16384, 1024, pu.SW128),
                     pu.smem_desc(b_off + ((dy + 2 * j) * 10 + dx) * 128, 0, 1280, pu.SW128),
                     pu.idesc_bf16(128, 64, 1, 1), int(j > 0), 0) for j in range(8)]
            t, _ = pu.run_probe(dyt, (128, 64), 128, xh, (180, 64), 128, _dummy_x(), (64, 8, 8, 1), 128, loads,
                                2 * 16384 + 180 * 128, mmas, 64, 65536, 16)
            res[f"dy{dy}dx{dx}"] = _err(t[:, :64], want)
    REPORT["halo_mnmajor_sw128"] = res
    _save()
    assert max(res.values()) < 1e-5, res


def test_wgrad_v2_tap_pair_stacked_in_m(pu):
    """wgrad v2 operand scheme: A' = the X halo tile consumed MN-major with M = 128 = TWO taps x 64 input channels,
    the second "64-channel block" being the same tile at another tap shift (LBO = byte distance between the two tap
    origins, e.g. 128 B); B' = the dense dY tile [128 px][64 co]; K = 16 pixels per MMA = two output rows."""
    xh = pu.rand_bf16(180, 64, seed=30).cuda()
    dyt = pu.rand_bf16(128, 64, seed=31).cuda()
    x_off, d_off = 0, 24576
    loads = [(0, (0, 0), x_off), (1, (0, 0), d_off)]
    res = {}
    pairs = [((0, 0), (0, 1)), ((0, 2), (1, 0)), ((1, 1), (1, 2)), ((2, 0), (2, 1)), ((2, 2), (2, 2))]
    for (t1, t2) in pairs:
        off1, off2 = (t1[0] * 10 + t1[1]) * 128, (t2[0] * 10 + t2[1]) * 128
        want = np.concatenate([pu.f32(xh)[_halo_rows(*t1)].T @ pu.f32(dyt), pu.f32(xh)[_halo_rows(*t2)].T @ pu.f32(dyt)])
        mmas = [(pu.smem_desc(x_off + off1 + 2 * j * 10 * 128, off2 - off1, 1280, pu.SW128),
                 pu.smem_desc(d_off + 2048 * j, 0, 1024, pu.SW128), pu.idesc_bf16(128, 64, 1, 1), int(j > 0), 0) for j in range(8)]
        t, _ = pu.run_probe(xh, (180, 64), 128, dyt, (128, 64), 128, _dummy_x(), (64, 8, 8, 1), 128, loads, 180 * 128 + 16384, mmas,
                            64, 65536, 16)
        res[f"{t1}{t2}"] = _err(t[:, :64], want)
    REPORT["wgrad_v2_pairs"] = res
    _save()
    assert max(res.values()) < 1e-5, res

