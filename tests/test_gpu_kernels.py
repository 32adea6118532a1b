"""Per-kernel parity through the C ABI (libeunet_b200.so via ctypes) against plain torch fp32 references
and the CPU oracle.  Integer work is bit-exact; fp32-mode kernels <= 1e-4 (normalised max error, the
north-star tolerance); bf16 kernels are compared against the same op evaluated on the bf16-rounded
inputs with a tolerance that only covers accumulation-order / output-rounding differences."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

DT = {"fp32": torch.float32, "bf16": torch.bfloat16, "fp16": torch.float16}
TOL = {"fp32": 1e-5, "bf16": 6e-3, "fp16": 8e-4}   # output rounding: bf16 2^-9, fp16 2^-12 relative
MODES = ["fp32", "bf16", "fp16"]


@pytest.fixture(scope="module")
def k():
    from enhanced_unet_b200 import lib
    lib.load()
    return lib


def nhwc(x: torch.Tensor, dt, ld=None, off=0):
    """[B,C,H,W] fp32 (cpu) -> CUDA [M, C] view (of an [M, ld] buffer when ld is given)."""
    B, C, H, W = x.shape
    flat = x.permute(0, 2, 3, 1).reshape(-1, C)
    if ld is None:
        return flat.to(dt).cuda().contiguous()
    buf = torch.full((flat.shape[0], ld), 7.0, dtype=dt, device="cuda")
    buf[:, off:off + C] = flat.to(dt).cuda()
    return buf[:, off:off + C]


def nchw(t: torch.Tensor, B, H, W):
    return t.float().cpu().reshape(B, H, W, -1).permute(0, 3, 1, 2).contiguous()


def nerr(got: torch.Tensor, want: torch.Tensor) -> float:
    return float((got.double() - want.double()).abs().max() / (want.double().abs().max() + 1e-12))


def rnd(dtn, t):   # what the kernel actually reads
    return t.to(DT[dtn]).float()


# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("dtype", [torch.int64, torch.int32, torch.uint8])
@pytest.mark.parametrize("shape", [(1, 64, 64), (3, 37, 41), (2, 1, 7), (5, 128, 130), (1, 1, 1)])
def test_confusion_counts_bit_exact(k, dtype, shape):
    import oracle
    from enhanced_unet_b200.ops import confusion_counts
    rng = np.random.default_rng(hash((str(dtype), shape)) % (2 ** 31))
    pred = rng.integers(0, 3, shape)
    gt = rng.integers(0, 3, shape)
    gt[rng.random(shape) < 0.05] = 255      # ignore label
    pred[rng.random(shape) < 0.01] = 7
    want = oracle.confusion_counts(pred, gt)
    got = confusion_counts(torch.from_numpy(pred).to(dtype).cuda(), torch.from_numpy(gt).to(dtype).cuda())
    assert got.dtype == torch.int64
    assert np.array_equal(got.cpu().numpy(), want)


def test_confusion_counts_large_checksum(k):
    """Full inference-config size (32 x 1024^2, uint8): the counts of every image must sum to the pixel
    count and match a torch bincount."""
    from enhanced_unet_b200.ops import confusion_counts
    g = torch.Generator(device="cuda").manual_seed(5)
    pred = torch.randint(0, 3, (32, 1024, 1024), device="cuda", dtype=torch.uint8, generator=g)
    gt = torch.randint(0, 4, (32, 1024, 1024), device="cuda", dtype=torch.uint8, generator=g)
    got = confusion_counts(pred, gt)
    assert torch.equal(got.sum((1, 2)), torch.full((32,), 1024 * 1024, device="cuda"))
    code = (gt.clamp(max=3).long() * 4 + pred.long()).reshape(32, -1)
    want = torch.stack([torch.bincount(code[i], minlength=16) for i in range(32)]).reshape(32, 4, 4)
    assert torch.equal(got, want)


# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("dtn", MODES)
@pytest.mark.parametrize("shape", [(2, 64, 8, 12), (1, 128, 6, 6), (3, 16, 2, 2)])
def test_maxpool_fwd_bwd(k, dtn, shape):
    B, C, H, W = shape
    dt = DT[dtn]
    g = torch.Generator().manual_seed(1)
    x = torch.randn(shape, generator=g).relu()            # many exact ties at 0, like post-ReLU activations
    xr = rnd(dtn, x).requires_grad_(True)
    want = F.max_pool2d(xr, 2)
    dp = torch.randn(want.shape, generator=g)
    want.backward(rnd(dtn, dp))
    xd = nhwc(x, dt, ld=C + 16, off=8)
    out = torch.empty(B * H * W // 4, C, dtype=dt, device="cuda")
    k.call("eunet_maxpool2_fwd", xd.data_ptr(), xd.stride(0), out.data_ptr(), C, k.dtype_code(dt), B, H, W, C)
    assert torch.equal(nchw(out, B, H // 2, W // 2), want.detach())
    dpd = nhwc(dp, dt)
    dx = torch.full((B * H * W, C), 3.0, dtype=dt, device="cuda")
    k.call("eunet_maxpool2_bwd", dpd.data_ptr(), C, xd.data_ptr(), xd.stride(0), dx.data_ptr(), C, 0, k.dtype_code(dt), B, H, W, C)
    assert torch.equal(nchw(dx, B, H, W), xr.grad)        # first-max tie-break of ATen
    base = torch.randn(B * H * W, C, generator=g).to(dt).cuda()
    dx2 = base.clone()
    k.call("eunet_maxpool2_bwd", dpd.data_ptr(), C, xd.data_ptr(), xd.stride(0), dx2.data_ptr(), C, 1, k.dtype_code(dt), B, H, W, C)
    want2 = (base.float() + dx.float()).to(dt)
    assert torch.equal(dx2, want2)


@pytest.mark.parametrize("dtn", MODES)
@pytest.mark.parametrize("shape", [(2, 64, 5, 7), (1, 16, 1, 1), (1, 8, 1, 4), (2, 128, 8, 8), (2, 64, 12, 20), (1, 128, 6, 9), (3, 192, 4, 8)])
def test_upsample_fwd_bwd(k, dtn, shape):
    B, C, H, W = shape
    dt = DT[dtn]
    g = torch.Generator().manual_seed(2)
    x = torch.randn(shape, generator=g)
    xr = rnd(dtn, x).requires_grad_(True)
    want = F.interpolate(xr, scale_factor=2, mode="bilinear", align_corners=False)
    do = torch.randn(want.shape, generator=g)
    want.backward(rnd(dtn, do))
    xd = nhwc(x, dt)
    out = torch.empty(B * 4 * H * W, C + 8, dtype=dt, device="cuda")[:, 8:]
    k.call("eunet_upsample2_fwd", xd.data_ptr(), C, out.data_ptr(), out.stride(0), k.dtype_code(dt), B, H, W, C)
    assert nerr(nchw(out, B, 2 * H, 2 * W), want.detach()) < TOL[dtn]
    dod = nhwc(do, dt)
    dx = torch.empty(B * H * W, C, dtype=dt, device="cuda")
    k.call("eunet_upsample2_bwd", dod.data_ptr(), C, dx.data_ptr(), C, k.dtype_code(dt), B, H, W, C)
    assert nerr(nchw(dx, B, H, W), xr.grad) < TOL[dtn]


@pytest.mark.parametrize("dtn", MODES)
@pytest.mark.parametrize("shape", [(2, 64, 5, 7), (1, 16, 1, 1), (2, 128, 8, 8), (1, 512, 6, 9), (2, 256, 12, 20)])
def test_bn_apply_relu_upsample_fused(k, dtn, shape):
    """BN apply + ReLU + bilinear x2 upsample in one pass over the RAW conv output (the activation of enc4 / dec4 / dec3 is
    consumed by the upsample only and never stored in training) against F.interpolate(relu(y * scale + shift)); the output is
    a channel slice of a wider (concat) buffer."""
    B, C, H, W = shape
    dt = DT[dtn]
    raw_dt = k.raw_dtype(dt)
    g = torch.Generator().manual_seed(C + H)
    y = torch.randn(shape, generator=g) * 1.5 + 0.2
    sc, sh = torch.rand(C, generator=g) + 0.5, torch.randn(C, generator=g) * 0.5
    yr = y.to(raw_dt).float()
    want = F.interpolate(F.relu(yr * sc[None, :, None, None] + sh[None, :, None, None]), scale_factor=2, mode="bilinear", align_corners=False)
    yd = nhwc(y, raw_dt)
    buf = torch.full((B * 4 * H * W, C + 16), 3.0, dtype=dt, device="cuda")
    out = buf[:, 8:8 + C]
    scd, shd = sc.cuda(), sh.cuda()
    k.call("eunet_bn_apply_relu_upsample2", yd.data_ptr(), C, out.data_ptr(), out.stride(0), k.dtype_code(dt), B, H, W, C, scd.data_ptr(),
           shd.data_ptr())
    torch.cuda.synchronize()
    assert nerr(nchw(out, B, 2 * H, 2 * W), want) < TOL[dtn]
    assert torch.all(buf[:, :8] == 3.0) and torch.all(buf[:, 8 + C:] == 3.0)


@pytest.mark.parametrize("dtn", MODES)
@pytest.mark.parametrize("C", [64, 128, 512])
def test_bn_train_apply_pool_and_backward(k, dtn, C):
    """conv output y -> batch stats -> finalize (+running stats) -> apply+ReLU(+pool) and the full backward,
    against torch BatchNorm2d + ReLU (+max_pool2d) in fp32 on the same (rounded) y."""
    B, H, W = 2, 6, 8
    dt = DT[dtn]
    g = torch.Generator().manual_seed(3)
    y = torch.randn(B, C, H, W, generator=g) * 1.7 + 0.3
    bias = torch.randn(C, generator=g) * 0.1
    gamma, beta = torch.rand(C, generator=g) + 0.5, torch.randn(C, generator=g) * 0.2
    rm0, rv0 = torch.randn(C, generator=g) * 0.1, torch.rand(C, generator=g) + 0.5
    yr = y.to(k.raw_dtype(dt)).float()       # raw conv outputs are stored in the RAW dtype (fp16 in bf16 mode)
    # reference: BN sees conv output + bias
    bn = torch.nn.BatchNorm2d(C)
    with torch.no_grad():
        bn.weight.copy_(gamma); bn.bias.copy_(beta); bn.running_mean.copy_(rm0); bn.running_var.copy_(rv0)
    yin = (yr + bias[None, :, None, None]).requires_grad_(True)
    act = F.relu(bn(yin))
    pooled = F.max_pool2d(act, 2)
    dact = torch.randn(act.shape, generator=g)
    dpool = torch.randn(pooled.shape, generator=g)
    (act * rnd(dtn, dact)).sum().backward(retain_graph=True)
    g_act_only = yin.grad.clone()
    # kernels
    M = B * H * W
    raw_dt = k.raw_dtype(dt)
    yd = nhwc(y, raw_dt)
    stats = torch.stack([yd.double().sum(0), (yd.double() ** 2).sum(0)]).reshape(-1).contiguous()
    dev = lambda t: t.clone().cuda()
    gm, bt, bs, rm, rv = dev(gamma), dev(beta), dev(bias), dev(rm0), dev(rv0)
    nbt = torch.zeros((), dtype=torch.int64, device="cuda")
    scale, shift, mean, invstd = (torch.empty(C, device="cuda") for _ in range(4))
    k.call("eunet_bn_finalize", stats.data_ptr(), M, gm.data_ptr(), bt.data_ptr(), bs.data_ptr(), rm.data_ptr(), rv.data_ptr(),
           nbt.data_ptr(), 0.1, 1e-5, scale.data_ptr(), shift.data_ptr(), mean.data_ptr(), invstd.data_ptr(), C)
    assert int(nbt) == 1
    assert nerr(rm.cpu(), bn.running_mean) < 1e-5 and nerr(rv.cpu(), bn.running_var) < 1e-5
    out = torch.empty(M, C + 8, dtype=dt, device="cuda")[:, :C]
    pl = torch.empty(M // 4, C, dtype=dt, device="cuda")
    k.call("eunet_bn_apply_relu", yd.data_ptr(), C, out.data_ptr(), out.stride(0), pl.data_ptr(), C, k.dtype_code(dt), B, H, W, C,
           scale.data_ptr(), shift.data_ptr())
    assert nerr(nchw(out, B, H, W), act.detach()) < TOL[dtn]
    assert torch.equal(nchw(pl, B, H // 2, W // 2), F.max_pool2d(nchw(out, B, H, W), 2))
    out2 = torch.empty(M, C, dtype=dt, device="cuda")
    k.call("eunet_bn_apply_relu", yd.data_ptr(), C, out2.data_ptr(), C, None, 0, k.dtype_code(dt), B, H, W, C,
           scale.data_ptr(), shift.data_ptr())
    assert torch.equal(out2, out.contiguous())
    # backward
    dad = nhwc(dact, dt)
    sums = torch.zeros(2 * C, dtype=torch.float64, device="cuda")
    k.call("eunet_bn_bwd_reduce", dad.data_ptr(), C, yd.data_ptr(), C, k.dtype_code(dt), M, C, scale.data_ptr(), shift.data_ptr(),
           mean.data_ptr(), invstd.data_ptr(), sums.data_ptr())
    dy = torch.empty(M, C, dtype=dt, device="cuda")
    dg, db = torch.empty(C, device="cuda"), torch.empty(C, device="cuda")
    k.call("eunet_bn_bwd_apply", dad.data_ptr(), C, yd.data_ptr(), C, dy.data_ptr(), C, k.dtype_code(dt), M, C, scale.data_ptr(),
           shift.data_ptr(), mean.data_ptr(), invstd.data_ptr(), sums.data_ptr(), dg.data_ptr(), db.data_ptr(), None)
    assert nerr(nchw(dy, B, H, W), g_act_only) < 2 * TOL[dtn]
    assert nerr(dg.cpu(), bn.weight.grad) < 1e-4 and nerr(db.cpu(), bn.bias.grad) < 1e-4
    # eval fold: conv epilogue affine == BN(eval)(y + bias)
    bn.eval()
    s2, h2 = torch.empty(C, device="cuda"), torch.empty(C, device="cuda")
    k.call("eunet_bn_fold_eval", gm.data_ptr(), bt.data_ptr(), bs.data_ptr(), rm.data_ptr(), rv.data_ptr(), 1e-5, s2.data_ptr(),
           h2.data_ptr(), C)
    with torch.no_grad():
        want_eval = bn(yr + bias[None, :, None, None])
    got_eval = yr * s2.cpu()[None, :, None, None] + h2.cpu()[None, :, None, None]
    assert nerr(got_eval, want_eval) < 1e-5


@pytest.mark.parametrize("dtn", MODES)
@pytest.mark.parametrize("case", [(2, 64, 8, 12), (1, 128, 6, 6), (2, 256, 4, 8), (3, 64, 2, 2)])
def test_bn_backward_with_fused_maxpool_routing(k, dtn, case):
    """bn_bwd_reduce_pool / bn_bwd_apply_pool: BN + ReLU backward of an encoder output whose gradient is
    dskip + maxpool-backward(dpool), with the routing (first maximum of the fp32 activations, ATen tie-break) done inside the
    two passes - against torch: max_pool2d indices of the activations, max_unpool2d, then autograd through
    relu(batch_norm(y)).  The skip gradient is a channel slice of a wider buffer."""
    B, C, H, W = case
    dt = DT[dtn]
    raw_dt = k.raw_dtype(dt)
    g = torch.Generator().manual_seed(C + H)
    y = torch.randn(B, C, H, W, generator=g) * 1.3 + 0.2
    gamma, beta = torch.rand(C, generator=g) + 0.5, torch.randn(C, generator=g) * 0.3
    dskip = rnd(dtn, torch.randn(B, C, H, W, generator=g))
    dpool = rnd(dtn, torch.randn(B, C, H // 2, W // 2, generator=g))
    yr = y.to(raw_dt).float().requires_grad_(True)
    gm, bt = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    act = F.relu(F.batch_norm(yr, None, None, gm, bt, True, 0.1, 1e-5))
    _, idx = F.max_pool2d(act.detach(), 2, return_indices=True)      # first maximum of the fp32 activations (see PoolWindow::grads)
    routed = F.max_unpool2d(dpool, idx, 2, output_size=(H, W))
    (act * (dskip + routed)).sum().backward()
    M = B * H * W
    yd = nhwc(y, raw_dt)
    mean = yr.detach().mean((0, 2, 3))
    invstd = 1.0 / torch.sqrt(yr.detach().var((0, 2, 3), unbiased=False) + 1e-5)
    scale, shift = (gamma * invstd).cuda(), (beta - mean * gamma * invstd).cuda()
    mean_d, invstd_d = mean.cuda(), invstd.cuda()
    dsd = nhwc(dskip, dt, ld=C + 16, off=8)
    dpd = nhwc(dpool, dt)
    sums = torch.zeros(2 * C, dtype=torch.float64, device="cuda")
    code = k.dtype_code(dt)
    k.call("eunet_bn_bwd_reduce_pool", dsd.data_ptr(), dsd.stride(0), dpd.data_ptr(), C, yd.data_ptr(), C, code, B, H, W, C,
           scale.data_ptr(), shift.data_ptr(), mean_d.data_ptr(), invstd_d.data_ptr(), sums.data_ptr())
    dy = torch.empty(M, C, dtype=dt, device="cuda")
    dg, db = torch.empty(C, device="cuda"), torch.empty(C, device="cuda")
    k.call("eunet_bn_bwd_apply_pool", dsd.data_ptr(), dsd.stride(0), dpd.data_ptr(), C, yd.data_ptr(), C, dy.data_ptr(), C, code, B, H, W,
           C, scale.data_ptr(), shift.data_ptr(), mean_d.data_ptr(), invstd_d.data_ptr(), sums.data_ptr(), dg.data_ptr(), db.data_ptr(),
           None)
    torch.cuda.synchronize()
    assert nerr(dg.cpu(), gm.grad) < 1e-4 and nerr(db.cpu(), bt.grad) < 1e-4
    assert nerr(nchw(dy, B, H, W), yr.grad) < 2 * TOL[dtn]


# ---------------------------------------------------------------------------------------------
CONV_CASES = [
    # B, H, W, Cin, Cout
    (2, 16, 16, 64, 64),
    (1, 8, 24, 128, 128),
    (2, 8, 8, 256, 256),
    (1, 4, 4, 512, 512),      # two N tiles of 256
    (1, 8, 8, 192, 64),       # concat-sized Cin
    (2, 12, 20, 16, 64),      # 16-channel chunks (first layer / enhance.0), ragged tile
    (1, 16, 16, 64, 16),      # dgrad of enhance.0 (N = 16)
    (3, 2, 2, 64, 128),       # tiny spatial extent: batch folded into the pixel tile
    (1, 40, 24, 64, 64),      # overhanging tiles
    (4, 128, 128, 64, 64),    # > 148 work items: persistent loop + double-buffered TMEM of the halo kernel
    (2, 48, 40, 192, 64),     # halo kernel, streamed filters, 3 channel chunks
    (1, 64, 64, 128, 128),    # halo kernel BN = 128
    (2, 80, 72, 16, 64),      # halo kernel, 16-channel chunks, MT = 4
    (1, 72, 64, 64, 16),      # halo kernel, N = 16 (dgrad of enhance.0)
    (2, 32, 32, 64, 192),     # three N tiles sharing the pixel blocks (dgrad of dec2.0)
]


def _conv_ref(x, w):
    return F.conv2d(x, w, None, padding=1)


@pytest.mark.parametrize("dtn", ["fp32", "bf16", "bf16-pertap", "fp16", "fp16-pertap"])
@pytest.mark.parametrize("case", CONV_CASES)
def test_conv3x3_fwd_stats_and_epilogue(k, dtn, case):
    B, H, W, Cin, Cout = case
    k.set_option("conv_halo", 0 if dtn.endswith("-pertap") else 2)     # both tensor-core kernels are covered (2: halo wherever it applies)
    dtn = dtn.split("-")[0]
    dt = DT[dtn]
    g = torch.Generator().manual_seed(hash(case) % 1000)
    x = torch.randn(B, Cin, H, W, generator=g)
    w = torch.randn(Cout, Cin, 3, 3, generator=g) / (3 * Cin ** 0.5)
    want = _conv_ref(rnd(dtn, x), rnd(dtn, w))
    xd = nhwc(x, dt, ld=Cin + 16, off=16)
    wp = torch.empty(Cout, 9, Cin, dtype=dt, device="cuda")
    k.call("eunet_pack_weight3x3", w.cuda().data_ptr(), wp.data_ptr(), k.dtype_code(dt), Cout, Cin, Cout, Cin, 0)
    assert torch.equal(wp.float().cpu(), rnd(dtn, w).permute(0, 2, 3, 1).reshape(Cout, 9, Cin))
    M = B * H * W
    ybuf = torch.full((M, Cout + 8), 5.0, dtype=dt, device="cuda")
    y = ybuf[:, 8:]
    stats = torch.zeros(2 * Cout, dtype=torch.float64, device="cuda")
    k.call("eunet_conv3x3_fwd", xd.data_ptr(), xd.stride(0), wp.data_ptr(), y.data_ptr(), y.stride(0), k.dtype_code(dt), B, H, W,
           Cin, Cout, stats.data_ptr(), None, None, 0, 0, None)
    torch.cuda.synchronize()
    assert nerr(nchw(y, B, H, W), want) < TOL[dtn]
    assert torch.all(ybuf[:, :8] == 5.0)                       # neighbouring channels of the wider buffer untouched
    s = stats.cpu()
    assert nerr(s[:Cout], want.double().sum((0, 2, 3))) < 1e-4
    assert nerr(s[Cout:], (want.double() ** 2).sum((0, 2, 3))) < 1e-4
    # fused eval epilogue: affine + ReLU
    sc, sh = (torch.rand(Cout, generator=g) + 0.5).cuda(), torch.randn(Cout, generator=g).cuda()
    y2 = torch.empty(M, Cout, dtype=dt, device="cuda")
    k.call("eunet_conv3x3_fwd", xd.data_ptr(), xd.stride(0), wp.data_ptr(), y2.data_ptr(), Cout, k.dtype_code(dt), B, H, W,
           Cin, Cout, None, sc.data_ptr(), sh.data_ptr(), 1, 0, None)
    # raw (pre-BN) output dtype: fp16 in bf16 mode (fp32 mode: identical to the activation dtype)
    raw_dt = k.raw_dtype(dt)
    y3 = torch.empty(M, Cout, dtype=raw_dt, device="cuda")
    k.call("eunet_conv3x3_fwd", xd.data_ptr(), xd.stride(0), wp.data_ptr(), y3.data_ptr(), Cout, k.dtype_code(dt), B, H, W,
           Cin, Cout, None, None, None, 0, 1, None)
    assert nerr(nchw(y3, B, H, W), want) < (1e-5 if dtn == "fp32" else 1e-3)
    k.set_option("conv_halo", 1)
    want2 = F.relu(want * sc.cpu()[None, :, None, None] + sh.cpu()[None, :, None, None])
    assert nerr(nchw(y2, B, H, W), want2) < TOL[dtn]


def test_unpack_wgrad_multi_matches_single(k):
    """eunet_unpack_wgrad3x3_multi (every filter gradient of a pass in one launch) == eunet_unpack_wgrad3x3 per tensor,
    including the hi/lo input split of the first layer and the gradient scale."""
    import ctypes
    g = torch.Generator(device="cuda").manual_seed(3)
    shapes = [(64, 3, 16, 1), (64, 3, 16, 0), (64, 64, 64, 0), (128, 192, 192, 0), (512, 256, 256, 0), (16, 48, 48, 0)]
    gs = torch.tensor([8.0, 0.125, 0.0, 0.0], device="cuda")
    srcs = [torch.randn(co, 9, cip, device="cuda", generator=g) for co, ci, cip, _ in shapes]
    want = []
    for (co, ci, cip, hilo), src in zip(shapes, srcs):
        d = torch.empty(co, ci, 3, 3, device="cuda")
        k.call("eunet_unpack_wgrad3x3", src.data_ptr(), d.data_ptr(), co, ci, cip, hilo, gs.data_ptr())
        want.append(d)
    got = [torch.full((co, ci, 3, 3), 7.0, device="cuda") for co, ci, _, _ in shapes]
    n = len(shapes)
    VP, IA = ctypes.c_void_p * n, ctypes.c_int * n
    k.call("eunet_unpack_wgrad3x3_multi", VP(*[s_.data_ptr() for s_ in srcs]), VP(*[d.data_ptr() for d in got]),
           IA(*[s_[0] for s_ in shapes]), IA(*[s_[1] for s_ in shapes]), IA(*[s_[2] for s_ in shapes]), IA(*[s_[3] for s_ in shapes]),
           n, gs.data_ptr())
    torch.cuda.synchronize()
    for a_, b_ in zip(got, want):
        assert torch.equal(a_, b_)
    ref = srcs[2].reshape(64, 3, 3, 64).permute(0, 3, 1, 2) * 0.125
    assert torch.equal(want[2], ref.contiguous())


@pytest.mark.parametrize("opt", [("bn192", 1, (2, 32, 40, 64, 192)), ("bn192", 2, (1, 24, 40, 128, 384)),
                                 ("a_ahead", 0, (2, 48, 40, 192, 64)), ("a_ahead", 0, (1, 64, 64, 128, 128))])
def test_conv3x3_alternative_configurations(k, opt):
    """The A/B switches of the halo kernel compute the same convolution: N = 192 tiles (off by default, measured slower) and
    the activation-tile prefetch order (a_ahead = 0: tiles requested after the previous chunk's last filter tap)."""
    name, value, (B, H, W, Cin, Cout) = opt
    g = torch.Generator().manual_seed(Cin + Cout)
    x = torch.randn(B, Cin, H, W, generator=g)
    w = torch.randn(Cout, Cin, 3, 3, generator=g) / (3 * Cin ** 0.5)
    want = _conv_ref(rnd("fp16", x), rnd("fp16", w))
    xd = nhwc(x, torch.float16)
    wp = torch.empty(Cout, 9, Cin, dtype=torch.float16, device="cuda")
    k.call("eunet_pack_weight3x3", w.cuda().data_ptr(), wp.data_ptr(), k.F16, Cout, Cin, Cout, Cin, 0)
    ys = []
    for v in (value, 1 - value if name == "a_ahead" else 0):
        k.set_option(name, v)
        try:
            y = torch.empty(B * H * W, Cout, dtype=torch.float16, device="cuda")
            stats = torch.zeros(2 * Cout, dtype=torch.float64, device="cuda")
            k.call("eunet_conv3x3_fwd", xd.data_ptr(), Cin, wp.data_ptr(), y.data_ptr(), Cout, k.F16, B, H, W, Cin, Cout,
                   stats.data_ptr(), None, None, 0, 0, None)
            torch.cuda.synchronize()
        finally:
            k.set_option(name, {"bn192": 0, "a_ahead": 1}[name])
        assert nerr(nchw(y, B, H, W), want) < TOL["fp16"]
        assert nerr(stats[:Cout].cpu(), want.double().sum((0, 2, 3))) < 1e-4
        ys.append(y)
    assert nerr(ys[0].float(), ys[1].float()) < 1e-3          # same products, at most a different accumulation order


@pytest.mark.parametrize("dtn", ["fp32", "bf16", "bf16-pertap", "fp16", "fp16-pertap"])
@pytest.mark.parametrize("case", CONV_CASES)
def test_conv3x3_dgrad_and_wgrad(k, dtn, case):
    B, H, W, Cin, Cout = case
    k.set_option("conv_halo", 0 if dtn.endswith("-pertap") else 2)     # 2 = halo kernels for every shape they cover
    dtn = dtn.split("-")[0]
    dt = DT[dtn]
    g = torch.Generator().manual_seed(hash(case) % 1000 + 1)
    x = rnd(dtn, torch.randn(B, Cin, H, W, generator=g)).requires_grad_(True)
    w = rnd(dtn, torch.randn(Cout, Cin, 3, 3, generator=g) / (3 * Cin ** 0.5)).requires_grad_(True)
    dy = rnd(dtn, torch.randn(B, Cout, H, W, generator=g))
    _conv_ref(x, w).backward(dy)
    xd, dyd = nhwc(x.detach(), dt), nhwc(dy, dt, ld=Cout + 8, off=0)
    M = B * H * W
    # dgrad = forward conv over dY with the flipped / transposed pack
    wpT = torch.empty(Cin, 9, Cout, dtype=dt, device="cuda")
    k.call("eunet_pack_weight3x3", w.detach().cuda().data_ptr(), wpT.data_ptr(), k.dtype_code(dt), Cout, Cin, Cout, Cin, 1)
    dx = torch.empty(M, Cin, dtype=dt, device="cuda")
    k.call("eunet_conv3x3_fwd", dyd.data_ptr(), dyd.stride(0), wpT.data_ptr(), dx.data_ptr(), Cin, k.dtype_code(dt), B, H, W,
           Cout, Cin, None, None, None, 0, 0, None)
    assert nerr(nchw(dx, B, H, W), x.grad) < TOL[dtn]
    if dtn != "fp32" and Cout % 64:
        return   # the bf16 wgrad kernel takes dY in 64-channel boxes (every layer of the model has Cout >= 64)
    dwp = torch.zeros(Cout, 9, Cin, dtype=torch.float32, device="cuda")
    k.call("eunet_conv3x3_wgrad", xd.data_ptr(), Cin, dyd.data_ptr(), dyd.stride(0), dwp.data_ptr(), k.dtype_code(dt), B, H, W,
           Cin, Cout)
    dw = torch.empty(Cout, Cin, 3, 3, device="cuda")
    k.call("eunet_unpack_wgrad3x3", dwp.data_ptr(), dw.data_ptr(), Cout, Cin, Cin, 0, None)
    k.set_option("conv_halo", 1)
    assert nerr(dw.cpu(), w.grad) < 2e-5     # fp32 accumulation in both modes (inputs are identical)


@pytest.mark.parametrize("dtn", ["bf16", "fp16"])
@pytest.mark.parametrize("case", [(1, 8, 24, 128, 128), (2, 48, 40, 192, 64), (1, 64, 64, 128, 128), (2, 32, 32, 64, 384), (4, 128, 128, 64, 64),
                                  (1, 40, 24, 64, 64)])
def test_conv3x3_cta_pair_kernel(k, dtn, case):
    """The CTA-pair (tcgen05 cta_group::2, M = 256 across two SMs) halo kernel of csrc/conv_halo2.cu - an option, off by
    default - against F.conv2d on the rounded operands: raw output + batch statistics, and the affine + ReLU epilogue.
    Odd item counts (the last pair holds one valid item), ragged tiles, several N tiles."""
    B, H, W, Cin, Cout = case
    dt = DT[dtn]
    g = torch.Generator().manual_seed(hash(case) % 1000 + 7)
    x = torch.randn(B, Cin, H, W, generator=g)
    w = torch.randn(Cout, Cin, 3, 3, generator=g) / (3 * Cin ** 0.5)
    want = _conv_ref(rnd(dtn, x), rnd(dtn, w))
    xd = nhwc(x, dt, ld=Cin + 16, off=16)
    wp = torch.empty(Cout, 9, Cin, dtype=dt, device="cuda")
    k.call("eunet_pack_weight3x3", w.cuda().data_ptr(), wp.data_ptr(), k.dtype_code(dt), Cout, Cin, Cout, Cin, 0)
    M = B * H * W
    y = torch.empty(M, Cout, dtype=k.raw_dtype(dt), device="cuda")
    y2 = torch.empty(M, Cout, dtype=dt, device="cuda")
    stats = torch.zeros(2 * Cout, dtype=torch.float64, device="cuda")
    sc, sh = (torch.rand(Cout, generator=g) + 0.5).cuda(), torch.randn(Cout, generator=g).cuda()
    k.set_option("cta_pair", 2)
    try:
        k.call("eunet_conv3x3_fwd", xd.data_ptr(), xd.stride(0), wp.data_ptr(), y.data_ptr(), Cout, k.dtype_code(dt), B, H, W, Cin, Cout,
               stats.data_ptr(), None, None, 0, 1, None)
        k.call("eunet_conv3x3_fwd", xd.data_ptr(), xd.stride(0), wp.data_ptr(), y2.data_ptr(), Cout, k.dtype_code(dt), B, H, W, Cin, Cout,
               None, sc.data_ptr(), sh.data_ptr(), 1, 0, None)
        torch.cuda.synchronize()
    finally:
        k.set_option("cta_pair", 0)
    assert nerr(nchw(y, B, H, W), want) < 1e-3
    s = stats.cpu()
    assert nerr(s[:Cout], want.double().sum((0, 2, 3))) < 1e-4 and nerr(s[Cout:], (want.double() ** 2).sum((0, 2, 3))) < 1e-4
    want2 = F.relu(want * sc.cpu()[None, :, None, None] + sh.cpu()[None, :, None, None])
    assert nerr(nchw(y2, B, H, W), want2) < TOL[dtn]


@pytest.mark.parametrize("dtn", ["bf16", "fp16"])
@pytest.mark.parametrize("shape", [(2, 16, 24), (1, 40, 20), (3, 8, 8), (2, 128, 136), (1, 250, 64)])
def test_conv3x3_dgrad_few_channels(k, shape, dtn):
    """Transposed tcgen05 dgrad for a conv with 3 real input channels (enhance.0): fp32 [pixels][4] output against
    autograd of F.conv2d on the bf16-rounded operands (fp32 accumulation in both: only summation order differs)."""
    B, H, W = shape
    g = torch.Generator().manual_seed(B * 1000 + H)
    x = torch.randn(B, 3, H, W, generator=g, requires_grad=True)
    dt = DT[dtn]
    w = rnd(dtn, torch.randn(64, 3, 3, 3, generator=g) / 5.0)
    dy = rnd(dtn, torch.randn(B, 64, H, W, generator=g))
    F.conv2d(x, w, None, padding=1).backward(dy)
    dyd = nhwc(dy, dt, ld=64 + 8, off=8)
    wpT = torch.empty(16, 9, 64, dtype=dt, device="cuda")
    k.call("eunet_pack_weight3x3", w.cuda().data_ptr(), wpT.data_ptr(), k.dtype_code(dt), 64, 3, 64, 16, 1)
    dx4 = torch.full((B * H * W, 4), 9.0, device="cuda")
    k.call("eunet_conv3x3_dgrad_few", dyd.data_ptr(), dyd.stride(0), wpT.data_ptr(), dx4.data_ptr(), k.dtype_code(dt), B, H, W, 64, 16)
    torch.cuda.synchronize()
    assert torch.all(dx4[:, 3] == 0)
    assert nerr(nchw(dx4[:, :3], B, H, W), x.grad) < 2e-5


@pytest.mark.parametrize("dtn", MODES)
def test_pack_weight_multi_matches_single(k, dtn):
    import ctypes
    dt = DT[dtn]
    g = torch.Generator().manual_seed(11)
    entries = [(64, 3, 2), (64, 64, 0), (64, 64, 1), (128, 64, 0), (256, 768, 1), (64, 3, 1), (64, 192, 0)]   # (Co, Ci, mode)
    ws, outs, want = [], [], []
    for co, ci, mode in entries:
        w = torch.randn(co, ci, 3, 3, generator=g).cuda()
        cop, cip = (co + 15) // 16 * 16, (ci + 15) // 16 * 16
        rows, inner = (cip, cop) if mode == 1 else (cop, cip)
        single = torch.empty(rows, 9, inner, dtype=dt, device="cuda")
        k.call("eunet_pack_weight3x3", w.data_ptr(), single.data_ptr(), k.dtype_code(dt), co, ci, cop, cip, mode)
        ws.append(w); want.append(single)
        outs.append(torch.full((rows, 9, inner), 3.0, dtype=dt, device="cuda"))
    n = len(entries)
    VP, IA = ctypes.c_void_p * n, ctypes.c_int * n
    pad = lambda c: (c + 15) // 16 * 16
    k.call("eunet_pack_weight3x3_multi", VP(*[w.data_ptr() for w in ws]), VP(*[o.data_ptr() for o in outs]),
           IA(*[e[0] for e in entries]), IA(*[e[1] for e in entries]), IA(*[pad(e[0]) for e in entries]),
           IA(*[pad(e[1]) for e in entries]), IA(*[e[2] for e in entries]), n, k.dtype_code(dt))
    for o, wnt, e in zip(outs, want, entries):
        assert torch.equal(o, wnt), e


@pytest.mark.parametrize("dtn", ["bf16", "fp16"])
@pytest.mark.parametrize("shape", [(2, 16, 24), (1, 72, 40), (3, 8, 8), (2, 136, 128)])
@pytest.mark.parametrize("store_mid", [True, False])
def test_conv3x3_tail_fwd_fused_epilogue(k, shape, store_mid, dtn):
    """Fused 2Hx2W tail: conv3x3 (3 -> 64, tcgen05) + BN affine + ReLU + 1x1 (64 -> 3) + residual + bias from the TMEM
    accumulators, against the same chain in torch fp32 on the bf16-rounded conv operands; plus the statistics-only pass."""
    B, H, W = shape
    g = torch.Generator().manual_seed(B * 100 + H)
    d1 = torch.randn(B, 3, H, W, generator=g)
    w0 = torch.randn(64, 3, 3, 3, generator=g) / 5.0
    sc, sh = torch.rand(64, generator=g) + 0.5, torch.randn(64, generator=g) * 0.5
    w3, b3 = torch.randn(3, 64, generator=g) / 8.0, torch.randn(3, generator=g)
    dt, code = DT[dtn], k.dtype_code(DT[dtn])
    mid = _conv_ref(rnd(dtn, d1), rnd(dtn, w0))
    want = d1 + b3[None, :, None, None] + torch.einsum("kc,bchw->bkhw", w3, F.relu(mid * sc[None, :, None, None] + sh[None, :, None, None]))
    M = B * H * W
    x16 = torch.zeros(M, 16, dtype=dt, device="cuda")
    x16[:, :3] = d1.permute(0, 2, 3, 1).reshape(M, 3).to(dt).cuda()
    d14 = torch.zeros(M, 4, device="cuda")
    d14[:, :3] = d1.permute(0, 2, 3, 1).reshape(M, 3).cuda()
    wp = torch.empty(64, 9, 16, dtype=dt, device="cuda")
    k.call("eunet_pack_weight3x3", w0.cuda().data_ptr(), wp.data_ptr(), code, 64, 3, 64, 16, 0)
    out = torch.full((B, 3, H, W), 7.0, device="cuda")
    midt = torch.full((M, 64), 3.0, dtype=torch.float16, device="cuda") if store_mid else None
    scd, shd, w3d, b3d = sc.cuda(), sh.cuda(), w3.cuda().contiguous(), b3.cuda()     # keep the device copies alive
    k.call("eunet_conv3x3_tail_fwd", x16.data_ptr(), wp.data_ptr(), midt.data_ptr() if store_mid else None, scd.data_ptr(),
           shd.data_ptr(), w3d.data_ptr(), b3d.data_ptr(), d14.data_ptr(), out.data_ptr(), code, B, H, W)
    torch.cuda.synchronize()
    assert nerr(out.cpu(), want) < 2e-5
    if store_mid:
        assert nerr(nchw(midt, B, H, W), mid) < 1e-3            # raw fp16 storage of the fp32 accumulators
    # statistics-only pass (y == NULL): the same sums as the storing call
    s1 = torch.zeros(128, dtype=torch.float64, device="cuda")
    k.call("eunet_conv3x3_fwd", x16.data_ptr(), 16, wp.data_ptr(), None, 64, code, B, H, W, 16, 64, s1.data_ptr(), None, None, 0, 1, None)
    assert nerr(s1[:64].cpu(), mid.double().sum((0, 2, 3))) < 1e-4
    assert nerr(s1[64:].cpu(), (mid.double() ** 2).sum((0, 2, 3))) < 1e-4


@pytest.mark.parametrize("dtn", ["bf16", "fp16"])
@pytest.mark.parametrize("shape", [(2, 16, 24), (1, 72, 40), (3, 8, 8), (2, 144, 136)])
def test_tail_bwd_fused_matches_separate_kernels(k, shape, dtn):
    """eunet_tail_bwd_fused (BN/ReLU backward formed on chip + wgrad + 3-channel transposed dgrad of enhance.0 from one staged
    tile) against the three separate kernels it replaces (tail_bwd_dmid -> conv3x3_wgrad / conv3x3_dgrad_few), which are
    themselves pinned against torch."""
    B, H2, W2 = shape
    M = B * H2 * W2
    dt, code = DT[dtn], k.dtype_code(DT[dtn])
    g = torch.Generator(device="cuda").manual_seed(B * 7 + H2)
    mid = (torch.randn(M, 64, device="cuda", generator=g) * 1.5 + 0.3).to(torch.float16)
    dout4 = torch.randn(M, 4, device="cuda", generator=g)
    dout4[:, 3] = 0
    d1p = torch.zeros(M, 16, dtype=dt, device="cuda")
    d1p[:, :3] = torch.randn(M, 3, device="cuda", generator=g).to(dt)
    w0 = torch.randn(64, 3, 3, 3, device="cuda", generator=g) / 5.0
    wflip = torch.empty(16, 9, 64, dtype=dt, device="cuda")
    k.call("eunet_pack_weight3x3", w0.data_ptr(), wflip.data_ptr(), code, 64, 3, 64, 16, 1)
    mean = mid.float().mean(0)
    invstd = 1.0 / torch.sqrt(mid.float().var(0, unbiased=False) + 1e-5)
    gamma, beta = torch.rand(64, device="cuda", generator=g) + 0.5, torch.randn(64, device="cuda", generator=g) * 0.3
    scale = (gamma * invstd).contiguous()
    shift = (beta - mean * scale).contiguous()
    w3 = (torch.randn(3, 64, device="cuda", generator=g) / 8.0).contiguous()
    acc = torch.zeros(328, dtype=torch.float64, device="cuda")
    # geometry arguments of the tail kernels are the HALF resolution (H, W) of a 2H x 2W grid
    assert H2 % 2 == 0 and W2 % 2 == 0
    k.call("eunet_tail_bwd_reduce", dout4.data_ptr(), mid.data_ptr(), code, scale.data_ptr(), shift.data_ptr(), mean.data_ptr(),
           invstd.data_ptr(), w3.data_ptr(), acc.data_ptr(), B, H2 // 2, W2 // 2)
    dmid = torch.empty(M, 64, dtype=dt, device="cuda")
    k.call("eunet_tail_bwd_dmid", dout4.data_ptr(), mid.data_ptr(), dmid.data_ptr(), code, scale.data_ptr(), shift.data_ptr(),
           mean.data_ptr(), invstd.data_ptr(), w3.data_ptr(), acc.data_ptr(), B, H2 // 2, W2 // 2)
    dw_ref = torch.zeros(64, 9, 16, device="cuda")
    k.call("eunet_conv3x3_wgrad", d1p.data_ptr(), 16, dmid.data_ptr(), 64, dw_ref.data_ptr(), code, B, H2, W2, 16, 64)
    dx_ref = torch.empty(M, 4, device="cuda")
    k.call("eunet_conv3x3_dgrad_few", dmid.data_ptr(), 64, wflip.data_ptr(), dx_ref.data_ptr(), code, B, H2, W2, 64, 16)
    dw = torch.zeros(64, 9, 16, device="cuda")
    dx = torch.full((M, 4), 5.0, device="cuda")
    k.call("eunet_tail_bwd_fused", dout4.data_ptr(), mid.data_ptr(), d1p.data_ptr(), wflip.data_ptr(), scale.data_ptr(),
           shift.data_ptr(), mean.data_ptr(), invstd.data_ptr(), w3.data_ptr(), acc.data_ptr(), dx.data_ptr(), dw.data_ptr(), code,
           B, H2, W2)
    torch.cuda.synchronize()
    # bf16 operands: the dmid formed on chip is bit-identical to tail_bwd_dmid's, only fp32 summation order differs.
    # fp16 operands: the fused kernel forms dmid in packed half2 arithmetic (exact ReLU mask, addends rounded to fp16):
    # measured 5e-4 .. 6e-4 against the fp32-then-round dmid of the separate kernel
    tol = 2e-5 if dtn == "bf16" else 2e-3
    assert nerr(dx, dx_ref) < tol, nerr(dx, dx_ref)
    assert nerr(dw, dw_ref) < tol, nerr(dw, dw_ref)
    if dtn == "fp16":
        k.set_option("tail_dbg", 16)          # fp32 transform arithmetic: bit-identical dmid again
        try:
            dw.zero_()
            k.call("eunet_tail_bwd_fused", dout4.data_ptr(), mid.data_ptr(), d1p.data_ptr(), wflip.data_ptr(), scale.data_ptr(),
                   shift.data_ptr(), mean.data_ptr(), invstd.data_ptr(), w3.data_ptr(), acc.data_ptr(), dx.data_ptr(), dw.data_ptr(),
                   code, B, H2, W2)
            torch.cuda.synchronize()
        finally:
            k.set_option("tail_dbg", 0)
        assert nerr(dx, dx_ref) < 2e-5 and nerr(dw, dw_ref) < 2e-5, (nerr(dx, dx_ref), nerr(dw, dw_ref))


@pytest.mark.parametrize("dtn", ["bf16", "fp16", "fp32"])
@pytest.mark.parametrize("shape", [(2, 8, 12), (1, 36, 20), (2, 72, 68), (3, 4, 4)])
def test_tail_backward_against_autograd(k, shape, dtn):
    """INDEPENDENT oracle for the backward of the 2Hx2W tail (reference models.py:236, 308-313, 337 under loss.backward()):
    torch autograd of   d1 = up(z);  out = d1 + conv1x1(relu(bn_train(conv3x3(d1; W0))); W3) + b3   in float64, against
    eunet_tail_pack3 -> tail_bwd_reduce -> {tail_bwd_fused | tail_bwd_dmid + conv3x3_wgrad + conv3x3_dgrad_few} ->
    tail_up_bwd, and eunet_tail_up_fwd for the forward leg.  shape = (B, H, W) of the HALF-resolution grid.

    The reference evaluates the BatchNorm on the stored (RAW-dtype) conv output so that both sides differentiate the same
    function; what is left is summation order plus the 16-bit rounding of the on-chip gradient dmid."""
    B, H, W = shape
    H2, W2 = 2 * H, 2 * W
    M1, M2 = B * H * W, B * H2 * W2
    dt, code = DT[dtn], k.dtype_code(DT[dtn])
    raw_dt = k.raw_dtype(dt)
    tc = dtn != "fp32"
    g = torch.Generator().manual_seed(B * 131 + H)
    z = torch.randn(B, 3, H, W, generator=g)
    w0 = rnd(dtn, torch.randn(64, 3, 3, 3, generator=g) / 5.0)
    gamma, beta = torch.rand(64, generator=g) + 0.5, torch.randn(64, generator=g) * 0.3
    w3, b3 = torch.randn(3, 64, generator=g) / 8.0, torch.randn(3, generator=g)
    dout = torch.randn(B, 3, H2, W2, generator=g)

    # ---- forward leg on the device: z4 -> d1p (activation dtype, 16 channels) + d14 (fp32)
    z4 = torch.zeros(M1, 4, device="cuda")
    z4[:, :3] = z.permute(0, 2, 3, 1).reshape(M1, 3).cuda()
    d1p = torch.full((M2, 16), 5.0, dtype=dt, device="cuda")
    d14 = torch.full((M2, 4), 5.0, device="cuda")
    k.call("eunet_tail_up_fwd", z4.data_ptr(), d1p.data_ptr(), d14.data_ptr(), code, B, H, W)
    zr = z.double().requires_grad_(True)
    d1 = F.interpolate(zr, scale_factor=2, mode="bilinear", align_corners=False)
    assert nerr(nchw(d14[:, :3], B, H2, W2), d1.detach()) < 1e-6
    assert nerr(nchw(d1p[:, :3], B, H2, W2), d1.detach()) < TOL[dtn] and float(d1p[:, 3:].float().abs().max()) == 0.0

    # ---- float64 autograd reference: the conv reads the ROUNDED d1 (d1p), the BN reads the RAW-dtype conv output
    d1_conv = nchw(d1p[:, :3], B, H2, W2).double()
    d1_conv_leaf = d1_conv.clone().requires_grad_(True)
    w0r = w0.double().requires_grad_(True)
    mid = F.conv2d(d1_conv_leaf, w0r, None, padding=1)
    mid_raw = mid.detach().to(raw_dt)                                   # what the forward kernel stores
    mid_leaf = mid_raw.double().requires_grad_(True)
    gm, bt = gamma.double().requires_grad_(True), beta.double().requires_grad_(True)
    w3r, b3r = w3.double().requires_grad_(True), b3.double().requires_grad_(True)
    a = F.relu(F.batch_norm(mid_leaf, None, None, gm, bt, True, 0.1, 1e-5))
    enh = torch.einsum("kc,bchw->bkhw", w3r, a) + b3r[None, :, None, None]
    (enh * dout.double()).sum().backward()
    dmid_ref = mid_leaf.grad                                            # BN + ReLU + 1x1 backward
    dmid_q = dmid_ref.to(dt).double() if tc else dmid_ref               # the tensor-core path rounds dmid to 16 bits
    mid.backward(dmid_q)
    dw0_ref, dd1_conv_ref = w0r.grad, d1_conv_leaf.grad
    d1.backward(dd1_conv_ref + dout.double())                           # conv path + residual path into the upsample adjoint
    dz_ref = zr.grad

    # ---- device chain
    mean = mid_raw.double().mean((0, 2, 3))
    var = mid_raw.double().var((0, 2, 3), unbiased=False)
    invstd = 1.0 / torch.sqrt(var + 1e-5)
    scale, shift = (gamma.double() * invstd).float().cuda(), (beta.double() - mean * gamma.double() * invstd).float().cuda()
    mean_d, invstd_d = mean.float().cuda(), invstd.float().cuda()
    midd = nhwc(mid_raw.float(), raw_dt)
    doutd = dout.cuda()
    dout4 = torch.full((M2, 4), 5.0, device="cuda")
    k.call("eunet_tail_pack3", doutd.data_ptr(), dout4.data_ptr(), B, H2, W2, None)
    assert torch.equal(nchw(dout4[:, :3], B, H2, W2), dout) and torch.all(dout4[:, 3] == 0)
    w3d = w3.cuda().contiguous()
    acc = torch.zeros(328, dtype=torch.float64, device="cuda")
    k.call("eunet_tail_bwd_reduce", dout4.data_ptr(), midd.data_ptr(), code, scale.data_ptr(), shift.data_ptr(), mean_d.data_ptr(),
           invstd_d.data_ptr(), w3d.data_ptr(), acc.data_ptr(), B, H, W)
    accc = acc.cpu()
    assert nerr(accc[0:64], bt.grad) < 1e-4, "dbeta"
    assert nerr(accc[64:128], gm.grad) < 1e-4, "dgamma"
    assert nerr(accc[128:320].reshape(3, 64), w3r.grad) < 1e-4, "dW3"
    assert nerr(accc[320:323], b3r.grad) < 1e-5, "db3"
    dmid = torch.empty(M2, 64, dtype=dt, device="cuda")
    k.call("eunet_tail_bwd_dmid", dout4.data_ptr(), midd.data_ptr(), dmid.data_ptr(), code, scale.data_ptr(), shift.data_ptr(),
           mean_d.data_ptr(), invstd_d.data_ptr(), w3d.data_ptr(), acc.data_ptr(), B, H, W)
    assert nerr(nchw(dmid, B, H2, W2), dmid_ref) < (2 * TOL[dtn] if tc else 2e-5), "dmid"
    wflip = torch.empty(16, 9, 64, dtype=dt, device="cuda")
    k.call("eunet_pack_weight3x3", w0.cuda().data_ptr(), wflip.data_ptr(), code, 64, 3, 64, 16, 1)
    tol_g = 3e-3 if tc else 2e-5     # a 1-ulp difference in the 16-bit rounding of single dmid elements is 2^-9 / 2^-12 of that element
    paths = ["separate"] + (["fused"] if tc and H2 >= 8 and W2 >= 8 else [])
    for path in paths:
        dwp = torch.zeros(64, 9, 16, device="cuda")
        if path == "fused":
            dd1 = torch.full((M2, 4), 5.0, device="cuda")
            k.call("eunet_tail_bwd_fused", dout4.data_ptr(), midd.data_ptr(), d1p.data_ptr(), wflip.data_ptr(), scale.data_ptr(),
                   shift.data_ptr(), mean_d.data_ptr(), invstd_d.data_ptr(), w3d.data_ptr(), acc.data_ptr(), dd1.data_ptr(),
                   dwp.data_ptr(), code, B, H2, W2)
            dd1_code, dd1_stride = k.F32, 4
        else:
            k.call("eunet_conv3x3_wgrad", d1p.data_ptr(), 16, dmid.data_ptr(), 64, dwp.data_ptr(), code, B, H2, W2, 16, 64)
            if tc and H2 >= 8 and W2 >= 8:
                dd1 = torch.full((M2, 4), 5.0, device="cuda")
                k.call("eunet_conv3x3_dgrad_few", dmid.data_ptr(), 64, wflip.data_ptr(), dd1.data_ptr(), code, B, H2, W2, 64, 16)
                dd1_code, dd1_stride = k.F32, 4
            else:
                dd1 = torch.empty(M2, 16, dtype=dt, device="cuda")
                k.call("eunet_conv3x3_fwd", dmid.data_ptr(), 64, wflip.data_ptr(), dd1.data_ptr(), 16, code, B, H2, W2, 64, 16, None, None,
                       None, 0, 0, None)
                dd1_code, dd1_stride = code, 16
        dw0 = torch.empty(64, 3, 3, 3, device="cuda")
        k.call("eunet_unpack_wgrad3x3", dwp.data_ptr(), dw0.data_ptr(), 64, 3, 16, 0, None)
        assert nerr(dw0.cpu(), dw0_ref) < tol_g, (path, "dW0", nerr(dw0.cpu(), dw0_ref))
        got_dd1 = nchw(dd1[:, :3], B, H2, W2)
        assert nerr(got_dd1, dd1_conv_ref) < (max(tol_g, TOL[dtn]) if dd1_stride == 16 else tol_g), (path, "dd1", nerr(got_dd1, dd1_conv_ref))
        dz4 = torch.full((M1, 4), 5.0, device="cuda")
        k.call("eunet_tail_up_bwd", dd1.data_ptr(), dd1_code, dd1_stride, doutd.data_ptr(), dz4.data_ptr(), B, H, W, None)
        assert nerr(nchw(dz4[:, :3], B, H, W), dz_ref) < max(tol_g, TOL[dtn] if dd1_stride == 16 else 0), (path, "dz")


def test_grad_scale_plumbing(k):
    """fp16 mode: eunet_grad_scale picks S = 2^floor(log2(target / max|g|)); tail_pack3 / tail_up_bwd apply S, unpack_wgrad /
    cast_f64_f32 / bn_bwd_apply remove it - all exact powers of two."""
    g = torch.Generator(device="cuda").manual_seed(3)
    for amp in (3e-6, 1.0, 700.0):
        dout = (torch.randn(2, 3, 8, 8, device="cuda", generator=g) * amp).contiguous()
        gs = torch.full((4,), 9.0, device="cuda")
        k.call("eunet_grad_scale", dout.data_ptr(), dout.numel(), 16.0, gs.data_ptr())
        S, inv = float(gs[0]), float(gs[1])
        m = float(dout.abs().max())
        assert S * inv == 1.0 and np.log2(S) == np.floor(np.log2(S)) and 8.0 <= S * m <= 16.0 * (1 + 1e-6), (amp, S, m)
        d4 = torch.empty(2 * 64, 4, device="cuda")
        k.call("eunet_tail_pack3", dout.data_ptr(), d4.data_ptr(), 2, 8, 8, gs.data_ptr())
        assert torch.equal(d4[:, :3].reshape(2, 8, 8, 3).permute(0, 3, 1, 2), dout * S)
        src = torch.randn(100, device="cuda", generator=g).double()
        dst = torch.empty(100, device="cuda")
        k.call("eunet_cast_f64_f32", src.data_ptr(), dst.data_ptr(), 100, gs.data_ptr())
        assert torch.equal(dst, (src * inv).float())
    z = torch.zeros(64, device="cuda")
    gs = torch.full((4,), 9.0, device="cuda")
    k.call("eunet_grad_scale", z.data_ptr(), 64, 16.0, gs.data_ptr())
    assert float(gs[0]) == 1.0 and float(gs[1]) == 1.0          # all-zero gradient: unscaled


def test_fp16_saturation_is_reported(k):
    """|y| > 65504 in an fp16 conv output is clamped by cvt.rn.satfinite - and REPORTED through `amax` (both conv kernels)."""
    B, H, W, C = 1, 16, 16, 64
    x = torch.full((B * H * W, C), 30.0, dtype=torch.float16, device="cuda")
    w = torch.full((C, C, 3, 3), 4.0)
    wp = torch.empty(C, 9, C, dtype=torch.float16, device="cuda")
    k.call("eunet_pack_weight3x3", w.cuda().data_ptr(), wp.data_ptr(), k.F16, C, C, C, C, 0)
    for halo in (2, 0):
        k.set_option("conv_halo", halo)
        try:
            y = torch.empty(B * H * W, C, dtype=torch.float16, device="cuda")
            amax = torch.zeros(1, device="cuda")
            k.call("eunet_conv3x3_fwd", x.data_ptr(), C, wp.data_ptr(), y.data_ptr(), C, k.F16, B, H, W, C, C, None, None, None, 0, 1,
                   amax.data_ptr())
            torch.cuda.synchronize()
        finally:
            k.set_option("conv_halo", 1)
        assert float(y.float().abs().max()) == 65504.0              # clamped, not inf
        assert float(amax) == 30.0 * 4.0 * 9 * 64                   # 69120: the true magnitude
        ok = torch.zeros(1, device="cuda")
        xs = (x.float() * 0.01).half()
        k.call("eunet_conv3x3_fwd", xs.data_ptr(), C, wp.data_ptr(), y.data_ptr(), C, k.F16, B, H, W, C, C, None, None, None, 0, 1,
               ok.data_ptr())
        assert float(ok) == 0.0                                     # in range: nothing written


@pytest.mark.parametrize("variant", ["tma", "thread_per_pixel"])
@pytest.mark.parametrize("shape", [(2, 8, 12), (1, 36, 20), (3, 64, 64), (1, 5, 13)])
def test_tail_out_fwd(k, shape, variant):
    """out = d1 + b3 + W3 . relu(mid * scale + shift) over the raw fp16 conv output (training-mode tail forward): the
    TMA-pipelined persistent kernel (chunk-per-warp) and the thread-per-pixel kernel against torch fp32 on the same fp16 data.
    shape = (B, H, W) of the HALF-resolution grid; the tensors live at 2H x 2W (ragged last tile for most shapes)."""
    B, H, W = shape
    M = B * 4 * H * W
    g = torch.Generator(device="cuda").manual_seed(H * 31 + W)
    mid = (torch.randn(M, 64, device="cuda", generator=g) * 2).to(torch.float16)
    d14 = torch.randn(M, 4, device="cuda", generator=g)
    sc, sh = torch.rand(64, device="cuda", generator=g) + 0.5, torch.randn(64, device="cuda", generator=g)
    w3, b3 = (torch.randn(3, 64, device="cuda", generator=g) / 8).contiguous(), torch.randn(3, device="cuda", generator=g)
    out = torch.full((B, 3, 2 * H, 2 * W), 9.0, device="cuda")
    k.set_option("tail_out_tma", 1 if variant == "tma" else 0)
    try:
        k.call("eunet_tail_out_fwd", d14.data_ptr(), mid.data_ptr(), k.BF16, sc.data_ptr(), sh.data_ptr(), w3.data_ptr(), b3.data_ptr(),
               out.data_ptr(), B, H, W)
        torch.cuda.synchronize()
    finally:
        k.set_option("tail_out_tma", 1)
    a = torch.relu(mid.float() * sc + sh)
    want = (d14[:, :3] + b3 + a @ w3.t()).reshape(B, 2 * H, 2 * W, 3).permute(0, 3, 1, 2)
    assert nerr(out, want) < 1e-5


@pytest.mark.parametrize("case", [(64, 5000, 64, 0), (128, 3333, 192, 64), (512, 1100, 768, 256), (256, 1024, 256, 0)])
def test_bn_bwd_reduce_on_channel_slices(k, case):
    """bn_bwd_reduce (masked sums, predicated accumulation, one resident wave) against a torch fp64 evaluation of sum g' /
    sum g' xhat; dact is a channel slice of a wider buffer (the concat gradient buffers), M is not a multiple of anything."""
    C, M, ld, off = case
    g = torch.Generator(device="cuda").manual_seed(C + M)
    y = (torch.randn(M, C, device="cuda", generator=g) * 1.3 + 0.4).to(torch.float16)
    buf = torch.randn(M, ld, device="cuda", generator=g).to(torch.bfloat16)
    dact = buf[:, off:off + C]
    mean = y.float().mean(0)
    invstd = 1.0 / torch.sqrt(y.float().var(0, unbiased=False) + 1e-5)
    scale = ((torch.rand(C, device="cuda", generator=g) + 0.5) * invstd).contiguous()
    shift = (torch.randn(C, device="cuda", generator=g) * 0.3 - mean * scale).contiguous()
    sums = torch.zeros(2 * C, dtype=torch.float64, device="cuda")
    k.call("eunet_bn_bwd_reduce", dact.data_ptr(), ld, y.data_ptr(), C, k.BF16, M, C, scale.data_ptr(), shift.data_ptr(),
           mean.data_ptr(), invstd.data_ptr(), sums.data_ptr())
    torch.cuda.synchronize()
    yd, dd = y.double(), dact.double()
    mask = (yd * scale.double() + shift.double()) > 0
    gp = torch.where(mask, dd, torch.zeros_like(dd))
    want = torch.cat([gp.sum(0), (gp * (yd - mean.double()) * invstd.double()).sum(0)])
    assert nerr(sums, want) < 2e-5, nerr(sums, want)


@pytest.mark.parametrize("dtn", ["bf16", "fp16"])
@pytest.mark.parametrize("case", [(5000, 64, 0), (777, 192, 128), (300, 64, 0)])
def test_tail_dec1_fwd(k, case, dtn):
    """z = dec1(d2) (1x1, 64 -> 3) on bf16 rows that may be a channel slice of a wider buffer: TMA-pipelined kernel for
    M >= 512 pixels, the 8-lanes-per-pixel kernel below that; against torch fp32 on the same bf16 data."""
    M, ld, off = case
    g = torch.Generator(device="cuda").manual_seed(M)
    buf = torch.randn(M, ld, device="cuda", generator=g).to(DT[dtn])
    d2 = buf[:, off:off + 64]
    w1, b1 = (torch.randn(3, 64, device="cuda", generator=g) / 8).contiguous(), torch.randn(3, device="cuda", generator=g)
    z4 = torch.full((M, 4), 7.0, device="cuda")
    k.call("eunet_tail_dec1_fwd", d2.data_ptr(), ld, k.dtype_code(DT[dtn]), w1.data_ptr(), b1.data_ptr(), z4.data_ptr(), M)
    torch.cuda.synchronize()
    want = d2.float() @ w1.t() + b1
    assert nerr(z4[:, :3], want) < 1e-5 and torch.all(z4[:, 3] == 0)


@pytest.mark.parametrize("dtn", ["bf16", "fp16"])
@pytest.mark.parametrize("case", [(5000, 64, 0, 64), (777, 192, 128, 72), (300, 64, 0, 64)])
def test_tail_dec1_bwd(k, case, dtn):
    """dec1 backward (dd2 = dz W1, dW1 = dz^T d2, db1 = sum dz) on strided bf16 rows: TMA-pipelined kernel with a TMA tensor
    store for M >= 512 pixels, the 8-lanes-per-pixel kernel below that; against torch on the same data."""
    M, ld, off, ldo = case
    g = torch.Generator(device="cuda").manual_seed(M + 1)
    buf = torch.randn(M, ld, device="cuda", generator=g).to(DT[dtn])
    d2 = buf[:, off:off + 64]
    dz4 = torch.randn(M, 4, device="cuda", generator=g)
    dz4[:, 3] = 0
    w1 = (torch.randn(3, 64, device="cuda", generator=g) / 8).contiguous()
    obuf = torch.full((M, ldo), 3.0, dtype=DT[dtn], device="cuda")
    dd2 = obuf[:, :64]
    acc = torch.zeros(200, dtype=torch.float64, device="cuda")
    k.call("eunet_tail_dec1_bwd", dz4.data_ptr(), d2.data_ptr(), ld, dd2.data_ptr(), ldo, k.dtype_code(DT[dtn]), w1.data_ptr(), acc.data_ptr(), M)
    torch.cuda.synchronize()
    assert nerr(dd2.float(), dz4[:, :3] @ w1) < TOL[dtn]                  # output rounding
    if ldo > 64:
        assert torch.all(obuf[:, 64:] == 3.0)                             # neighbouring channels untouched
    assert nerr(acc[:192].reshape(3, 64), dz4[:, :3].double().t() @ d2.double()) < 1e-5
    assert nerr(acc[192:195], dz4[:, :3].double().sum(0)) < 1e-5


@pytest.mark.parametrize("dtn", ["bf16", "fp16"])
@pytest.mark.parametrize("case", [(64, 5000, 64, 192, 128), (128, 3333, 128, 384, 256), (512, 1100, 512, 512, 0)])
def test_bn_apply_relu_tma_on_channel_slices(k, case, dtn):
    """TMA load -> transform -> TMA store bn_apply_relu against the cp.async ring kernel (bit-identical) and torch; the output
    is a channel slice of a wider buffer (skip slices of the concat buffers), M is not a multiple of the 128-pixel tile."""
    C, M, ldy, ldo, off = case
    g = torch.Generator(device="cuda").manual_seed(C + M)
    y = (torch.randn(M, ldy, device="cuda", generator=g) * 1.5).to(torch.float16)
    sc, sh = torch.rand(C, device="cuda", generator=g) + 0.5, torch.randn(C, device="cuda", generator=g)
    outs = {}
    for name, opt in (("tma", 1), ("ring", 0)):
        buf = torch.full((M, ldo), 3.0, dtype=DT[dtn], device="cuda")
        out = buf[:, off:off + C]
        k.set_option("bn_tma", opt)
        try:
            k.call("eunet_bn_apply_relu", y.data_ptr(), ldy, out.data_ptr(), ldo, None, 0, k.dtype_code(DT[dtn]), 1, 1, M, C, sc.data_ptr(), sh.data_ptr())
            torch.cuda.synchronize()
        finally:
            k.set_option("bn_tma", 1)
        outs[name] = buf
    assert torch.equal(outs["tma"], outs["ring"])
    want = torch.relu(y[:, :C].float() * sc + sh)
    assert nerr(outs["tma"][:, off:off + C].float(), want) < TOL[dtn]
    if ldo > C:
        rest = torch.cat([outs["tma"][:, :off], outs["tma"][:, off + C:]], 1)
        assert torch.all(rest == 3.0)


@pytest.mark.parametrize("dtn", ["bf16", "fp16"])
def test_pack_input_and_padded_weights(k, dtn):
    dt, code = DT[dtn], k.dtype_code(DT[dtn])
    g = torch.Generator().manual_seed(4)
    x = torch.rand(2, 3, 8, 8, generator=g)
    out = torch.empty(2 * 64, 16, dtype=dt, device="cuda")
    k.call("eunet_pack_input_nchw", x.cuda().data_ptr(), out.data_ptr(), code, 2, 3, 8, 8, 16, 0)
    want = torch.zeros(2, 16, 8, 8); want[:, :3] = x.to(dt).float()
    assert torch.equal(nchw(out, 2, 8, 8), want)
    # hi/lo split of the 3-channel input + {w_hi, w_hi, w_lo} filters: the bf16 conv then reproduces the fp32 conv to ~2^-16
    k.call("eunet_pack_input_nchw", x.cuda().data_ptr(), out.data_ptr(), code, 2, 3, 8, 8, 16, 1)
    hi = x.to(dt).float()
    got = nchw(out, 2, 8, 8)
    assert torch.equal(got[:, 0:3], hi) and torch.equal(got[:, 6:9], hi) and torch.equal(got[:, 3:6], (x - hi).to(dt).float())
    assert float(got[:, 9:].abs().max()) == 0.0
    w1 = torch.randn(64, 3, 3, 3, generator=g) / 5
    wp2 = torch.empty(64, 9, 16, dtype=dt, device="cuda")
    k.call("eunet_pack_weight3x3", w1.cuda().data_ptr(), wp2.data_ptr(), code, 64, 3, 64, 16, 2)
    y = torch.empty(2 * 64, 64, dtype=torch.float16, device="cuda")
    k.call("eunet_conv3x3_fwd", out.data_ptr(), 16, wp2.data_ptr(), y.data_ptr(), 64, code, 2, 8, 8, 16, 64, None, None, None, 0, 1, None)
    ref = F.conv2d(x, w1, None, padding=1)
    assert nerr(nchw(y, 2, 8, 8), ref) < 1e-3            # fp16 output rounding only (plain bf16 operands: ~4e-3)
    dy = torch.randn(2, 64, 8, 8, generator=g).to(dt)
    dyd = nhwc(dy.float(), dt)
    dwp = torch.zeros(64, 9, 16, dtype=torch.float32, device="cuda")
    k.call("eunet_conv3x3_wgrad", out.data_ptr(), 16, dyd.data_ptr(), 64, dwp.data_ptr(), code, 2, 8, 8, 16, 64)
    dw = torch.empty(64, 3, 3, 3, device="cuda")
    k.call("eunet_unpack_wgrad3x3", dwp.data_ptr(), dw.data_ptr(), 64, 3, 16, 1, None)
    xr = x.clone().requires_grad_(False)
    wr = w1.clone().requires_grad_(True)
    F.conv2d(xr, wr, None, padding=1).backward(dy.float())
    assert nerr(dw.cpu(), wr.grad) < 1e-4                  # x_hi + x_lo carries ~16 bits of the input
    w = torch.randn(64, 3, 3, 3, generator=g)
    wp = torch.empty(64, 9, 16, dtype=torch.float32, device="cuda")
    k.call("eunet_pack_weight3x3", w.cuda().data_ptr(), wp.data_ptr(), k.F32, 64, 3, 64, 16, 0)
    ref = torch.zeros(64, 9, 16); ref[:, :, :3] = w.permute(0, 2, 3, 1).reshape(64, 9, 3)
    assert torch.equal(wp.cpu(), ref)
    wpT = torch.empty(16, 9, 64, dtype=torch.float32, device="cuda")
    k.call("eunet_pack_weight3x3", w.cuda().data_ptr(), wpT.data_ptr(), k.F32, 64, 3, 64, 16, 1)
    refT = torch.zeros(16, 9, 64); refT[:3] = w.flip(2, 3).permute(1, 2, 3, 0).reshape(3, 9, 64)
    assert torch.equal(wpT.cpu(), refT)


# ---------------------------------------------------------------------------------------------
def test_loss_matches_reference_fixture_and_oracle(k, golden_dir):
    import os
    import oracle
    from enhanced_unet_b200.ops import combined_loss
    gold = np.load(os.path.join(golden_dir, "loss.npz"))
    for name in "abc":
        logits = torch.from_numpy(gold[f"{name}/logits"]).cuda().requires_grad_(True)
        t = torch.from_numpy(gold[f"{name}/target"]).cuda()
        loss = combined_loss(logits, t)
        ref = float(gold[f"{name}/loss"])
        assert abs(loss.item() - ref) <= 1e-5 * abs(ref), (name, loss.item(), ref)
        (loss * 1.0).backward()
        rg = torch.from_numpy(gold[f"{name}/grad"])
        assert nerr(logits.grad.cpu(), rg) < 2e-5, name
    # same-resolution logits (scale 1) against the oracle
    g = torch.Generator().manual_seed(12)
    lg = torch.randn(2, 3, 24, 40, generator=g) * 2
    tg = torch.randint(0, 3, (2, 24, 40), generator=g)
    lr = lg.clone().requires_grad_(True)
    want = oracle.batch_loss(lr, tg)
    want.backward()
    lc = lg.cuda().requires_grad_(True)
    got = combined_loss(lc, tg.cuda())
    (got * 3.0).backward()
    assert abs(got.item() - want.item()) <= 1e-5 * abs(want.item())
    assert nerr(lc.grad.cpu(), 3.0 * lr.grad) < 2e-5


def test_adamw_and_clip_against_torch(k):
    g = torch.Generator().manual_seed(13)
    n = 10007
    p0, g0 = torch.randn(n, generator=g), torch.randn(n, generator=g) * 3
    pr = p0.clone().requires_grad_(True)
    opt = torch.optim.AdamW([pr], lr=4e-3, weight_decay=1e-4, betas=(0.9, 0.999))
    p, m, v = p0.clone().cuda(), torch.zeros(n, device="cuda"), torch.zeros(n, device="cuda")
    for step in range(1, 4):
        gi = g0 * step
        pr.grad = gi.clone()
        torch.nn.utils.clip_grad_norm_([pr], 1.0)
        opt.step()
        gd = gi.cuda()
        sq = torch.zeros((), dtype=torch.float64, device="cuda")
        k.call("eunet_sumsq", gd.data_ptr(), n, sq.data_ptr())
        assert abs(sq.item() - float((gi.double() ** 2).sum())) < 1e-6 * sq.item()
        k.call("eunet_adamw_step", p.data_ptr(), gd.data_ptr(), m.data_ptr(), v.data_ptr(), n, sq.data_ptr(), 1.0, 4e-3, 0.9,
               0.999, 1e-8, 1e-4, step, 1.0)
        assert nerr(p.cpu(), pr.detach()) < 1e-5
