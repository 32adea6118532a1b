"""Host-side schedule of the Enhanced-UNet hot path over the C-ABI kernels (libeunet_b200.so).

Mirrors ``BasicUNet.forward`` (reference models.py:227-238) + the enhance head (308-313, 337) and its
autograd backward, but laid out B200-first:

* activations are NHWC in HBM (bf16 by default, fp32 in "fp32 mode"); ``torch.cat([up, skip])`` is never
  materialised - producers write straight into channel slices of one concat buffer per level;
* conv -> BN(train) is "conv with fused fp64 batch statistics" + a tiny finalize + one apply/ReLU pass
  (fused with the 2x2 max-pool in the encoder); eval-mode BN is folded into the conv epilogue;
* the 2Hx2W tail computes dec1 (1x1) BEFORE the last upsample (they commute) so no 64-channel tensor
  is materialised at 2Hx2W except the enhance conv's own output;
* dgrad = the same conv kernel over dY with flipped/transposed packed filters; wgrad has its own kernel.

torch is used for device memory (caching allocator) and the current stream only.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Tuple

import torch

from . import lib
from .lib import call, ptr

BLOCKS = [  # (prefix, Cin, Cout) - reference models.py:203-211
    ("model.enc1", 3, 64), ("model.enc2", 64, 128), ("model.enc3", 128, 256), ("model.enc4", 256, 512),
    ("model.dec4", 768, 256), ("model.dec3", 384, 128), ("model.dec2", 192, 64),
]
BN_MOMENTUM = 0.1
BN_EPS = 1e-5
FUSED_TAIL_BWD = True  # bf16 mode: dmid + wgrad + dgrad of enhance.0 in one kernel (False: the three separate kernels)
FUSED_TAIL = None      # None: fused 2Hx2W tail epilogue in eval mode only; True / False force it (benchmarks, tests)


def _pad16(c: int) -> int:
    return (c + 15) // 16 * 16


class _Ctx:
    """Per-call helpers bound to a device / activation dtype."""

    def __init__(self, device: torch.device, act_dtype: torch.dtype):
        self.dev = device
        self.dt = act_dtype
        self.code = lib.dtype_code(act_dtype)
        self.raw = lib.raw_dtype(act_dtype)   # pre-BN conv outputs: fp16 in the 16-bit modes, fp32 in fp32 mode
        self.tc = act_dtype in (torch.bfloat16, torch.float16)    # tensor-core (tcgen05) convolutions
        self.amax: Optional[torch.Tensor] = None    # fp16 saturation monitor (1 float, sticky max of |y| beyond the fp16 range)
        self.gs: Optional[torch.Tensor] = None      # fp16 mode, backward: {S, 1/S} gradient scale (eunet_grad_scale)

    def empty(self, *shape, dtype=None):
        return torch.empty(*shape, device=self.dev, dtype=dtype or self.dt)

    def zeros(self, *shape, dtype=None):
        return torch.zeros(*shape, device=self.dev, dtype=dtype or self.dt)


def _ld(t: torch.Tensor) -> int:
    assert t.dim() == 2 and t.stride(1) == 1, "activation operands are [pixels, channels] views with unit channel stride"
    return t.stride(0)


# ---------------------------------------------------------------------------------------------
# packed-filter cache (invalidated by the parameter's in-place version counter)
# ---------------------------------------------------------------------------------------------
class PackCache:
    def __init__(self):
        self._c: Dict[Tuple[str, int, int], Tuple[int, int, torch.Tensor]] = {}

    @staticmethod
    def _key(cx: _Ctx, name: str, flip: bool, hilo: bool):
        return (name, 2 if hilo else int(flip), cx.code)

    def _hit(self, key, w: torch.Tensor) -> Optional[torch.Tensor]:
        hit = self._c.get(key)
        if hit is not None and hit[0] == (w._version, w.data_ptr()) and hit[2].device == w.device:
            return hit[2]
        return None

    def get(self, cx: _Ctx, name: str, w: torch.Tensor, flip: bool, hilo: bool = False) -> torch.Tensor:
        key = self._key(cx, name, flip, hilo)
        out = self._hit(key, w)
        if out is not None:
            return out
        co, ci = w.shape[0], w.shape[1]
        cop, cip = _pad16(co), _pad16(ci)
        rows, inner = (cip, cop) if flip else (cop, cip)
        out = cx.empty(rows, 9, inner)
        call("eunet_pack_weight3x3", ptr(w), ptr(out), cx.code, co, ci, cop, cip, 2 if hilo else int(flip))
        self._c[key] = ((w._version, w.data_ptr()), 0, out)
        return out

    def prepare(self, cx: _Ctx, sd: Dict[str, torch.Tensor], specs) -> None:
        """Pack every stale filter of ``specs`` = [(conv name, flip, hilo)] in ONE launch (after an optimiser step all
        30 operands of a training step are stale).  Later ``get`` calls hit the cache."""
        import ctypes
        todo = []
        for name, flip, hilo in specs:
            w = sd[name + ".weight"]
            key = self._key(cx, name, flip, hilo)
            if self._hit(key, w) is None:
                co, ci = w.shape[0], w.shape[1]
                cop, cip = _pad16(co), _pad16(ci)
                rows, inner = (cip, cop) if flip else (cop, cip)
                old = self._c.get(key)
                out = old[2] if old is not None and old[2].shape == (rows, 9, inner) and old[2].device == w.device else cx.empty(rows, 9, inner)
                todo.append((key, w, out, co, ci, cop, cip, 2 if hilo else int(flip)))
        for i in range(0, len(todo), 32):
            part = todo[i:i + 32]
            n = len(part)
            VP, IA = ctypes.c_void_p * n, ctypes.c_int * n
            call("eunet_pack_weight3x3_multi", VP(*[t[1].data_ptr() for t in part]), VP(*[t[2].data_ptr() for t in part]),
                 IA(*[t[3] for t in part]), IA(*[t[4] for t in part]), IA(*[t[5] for t in part]), IA(*[t[6] for t in part]),
                 IA(*[t[7] for t in part]), n, cx.code)
        for key, w, out, *_ in todo:
            self._c[key] = ((w._version, w.data_ptr()), 0, out)

    def invalidate(self):
        """Parameters were updated in place behind autograd's back (optimiser kernels): every packed copy is stale, the
        buffers are kept for the next ``prepare``."""
        self._c = {k: (None, 0, v[2]) for k, v in self._c.items()}

    def clear(self):
        self._c.clear()


def pack_specs(cx: _Ctx, train: bool):
    """(conv, flip, hilo) of every packed filter a forward (+ backward when ``train``) pass uses."""
    specs = []
    for prefix, _, _ in BLOCKS:
        first = prefix == "model.enc1"
        specs.append((prefix + ".0", False, first and _first_layer_split(cx)))
        specs.append((prefix + ".3", False, False))
        if train:
            if not first:
                specs.append((prefix + ".0", True, False))
            specs.append((prefix + ".3", True, False))
    specs.append(("enhance.0", False, False))
    if train:
        specs.append(("enhance.0", True, False))
    return specs


def conv3x3(cx: _Ctx, x: torch.Tensor, wp: torch.Tensor, y: torch.Tensor, B: int, H: int, W: int, cin: int, cout: int,
            stats: Optional[torch.Tensor] = None, scale: Optional[torch.Tensor] = None,
            shift: Optional[torch.Tensor] = None, relu: bool = False, cin_real: int = 0, cout_real: int = 0) -> None:
    """``cin`` / ``cout`` are the (16-padded) channel counts the kernel runs on; ``cin_real`` (when the input is a padded
    3-channel tensor) (``cout_real`` for a dgrad onto 3 channels) is what the ALGORITHMIC FLOP count of the launch uses (SURVEY.md §8d:
    K = 9 * 3 = 27); such launches are tagged "k27" (HBM-roofline kernels), the others "tc" (tensor-core roofline)."""
    out_raw = int(y.dtype == cx.raw and cx.raw != cx.dt)
    assert y.dtype in (cx.dt, cx.raw) and x.dtype == cx.dt
    call("eunet_conv3x3_fwd", ptr(x), _ld(x), ptr(wp), ptr(y), _ld(y), cx.code, B, H, W, cin, cout, ptr(stats), ptr(scale),
         ptr(shift), int(relu), out_raw, ptr(cx.amax), flops=2.0 * B * H * W * (cout_real or cout) * 9 * (cin_real or cin),
         tag="k27" if (cin_real or cout_real) else "tc")


class _ZeroPool:
    """One zero-filled workspace (a single fill launch) that hands out the accumulators of a pass: packed wgrad
    accumulators (fp32), BatchNorm statistics / backward sums (fp64), exact-zero bias gradients (fp32)."""

    def __init__(self, cx: _Ctx, numel: int, dtype: torch.dtype = torch.float32):
        self.buf = cx.zeros(numel, dtype=dtype)
        self.off = 0

    def take(self, *shape) -> torch.Tensor:
        n = 1
        for d in shape:
            n *= d
        if self.off + n > self.buf.numel():
            raise RuntimeError("zero-filled workspace exhausted")
        t = self.buf[self.off:self.off + n].view(*shape)
        self.off += (n + 63) // 64 * 64
        return t


STATS_NUMEL = 2 * (2 * sum(c for _, _, c in BLOCKS) + 64) + 64 * len(BLOCKS) * 2 + 1024   # fp64 accumulators of one pass (+ padding)
ZERO_BIAS_NUMEL = 2 * sum(c for _, _, c in BLOCKS) + 64 + 64 * (2 * len(BLOCKS) + 1)


def wgrad_workspace_numel() -> int:
    n = 64 * 9 * 16 + 64                                       # enhance.0
    for _, cin, cout in BLOCKS:
        n += cout * 9 * _pad16(cin) + cout * 9 * cout + 128
    return n


def conv3x3_wgrad(cx: _Ctx, x: torch.Tensor, dy: torch.Tensor, B: int, H: int, W: int, cin: int, cout: int,
                  pool: Optional[_ZeroPool] = None, cin_real: int = 0) -> torch.Tensor:
    dwp = pool.take(cout, 9, cin) if pool is not None else cx.zeros(cout, 9, cin, dtype=torch.float32)
    call("eunet_conv3x3_wgrad", ptr(x), _ld(x), ptr(dy), _ld(dy), ptr(dwp), cx.code, B, H, W, cin, cout,
         flops=2.0 * B * H * W * cout * 9 * (cin_real or cin), tag="k27" if cin_real else "tc")
    return dwp


def unpack_wgrad(dwp: torch.Tensor, co: int, ci: int, hilo: bool = False, out: Optional[torch.Tensor] = None,
                 gscale: Optional[torch.Tensor] = None) -> torch.Tensor:
    dw = out if out is not None else torch.empty(co, ci, 3, 3, device=dwp.device, dtype=torch.float32)
    assert dw.is_contiguous() and dw.numel() == co * ci * 9 and dw.dtype == torch.float32
    call("eunet_unpack_wgrad3x3", ptr(dwp), ptr(dw), co, ci, dwp.shape[2], int(hilo), ptr(gscale))
    return dw


def unpack_wgrad_multi(items, gscale: Optional[torch.Tensor]) -> None:
    """``items`` = [(packed gradient, destination, co, ci, hilo)]: every filter gradient of a pass in ONE launch."""
    import ctypes
    for i in range(0, len(items), 32):
        part = items[i:i + 32]
        n = len(part)
        VP, IA = ctypes.c_void_p * n, ctypes.c_int * n
        for dwp, dw, co, ci, _ in part:
            assert dw.is_contiguous() and dw.numel() == co * ci * 9 and dw.dtype == torch.float32
        call("eunet_unpack_wgrad3x3_multi", VP(*[t[0].data_ptr() for t in part]), VP(*[t[1].data_ptr() for t in part]),
             IA(*[t[2] for t in part]), IA(*[t[3] for t in part]), IA(*[t[0].shape[2] for t in part]),
             IA(*[int(t[4]) for t in part]), n, ptr(gscale))


class GradSink:
    """Where ``backward`` puts parameter gradients.  The default allocates one fp32 tensor per parameter; the
    data-parallel path passes a ``parallel.FlatGradBuffer`` so that gradients land in ONE flat buffer laid out in
    production order and finished buckets can be all-reduced while the rest of backward still runs."""

    def closes(self, name: str) -> bool:
        """True when ``ready(name)`` hands gradients to a consumer that runs before backward returns (a data-parallel
        bucket leaving).  The filter-gradient unpacks are deferred into ONE multi-tensor launch per such hand-over - a
        single launch at the end of backward here, instead of fifteen."""
        return False

    def __init__(self, device: torch.device):
        self.dev = device
        self.grads: Dict[str, torch.Tensor] = {}

    def dst(self, name: str, shape) -> torch.Tensor:
        t = torch.empty(shape, device=self.dev, dtype=torch.float32)
        self.grads[name] = t
        return t

    def zero_dst(self, name: str, shape, pool: "_ZeroPool") -> torch.Tensor:
        """Destination of a gradient that is EXACTLY zero (conv bias in front of a train-mode BN): a slice of the pass's
        zero-filled pool, so that fifteen such gradients cost no fill launches."""
        t = pool.take(*shape)
        self.grads[name] = t
        return t

    def ready(self, name: str) -> None:
        pass


# order in which ``backward`` finishes parameter gradients (tail first, enc1 last): the layout of flat gradient buffers
def grad_production_order() -> List[str]:
    names = ["enhance.1.bias", "enhance.1.weight", "enhance.3.weight", "enhance.3.bias", "enhance.0.weight", "enhance.0.bias",
             "model.dec1.weight", "model.dec1.bias"]
    for blk in ("dec2", "dec3", "dec4", "enc4", "enc3", "enc2", "enc1"):
        for sub in ("4.weight", "4.bias", "3.weight", "3.bias", "1.weight", "1.bias", "0.weight", "0.bias"):
            names.append(f"model.{blk}.{sub}")
    return names


def _first_layer_split(cx: _Ctx) -> bool:
    """16-bit modes: the 3-channel network input and the first filters are split into hi + lo parts that ride in the
    otherwise zero padding channels (free: K stays 16), removing the largest single source of train-mode logit error."""
    return cx.tc


class _BNSaved:
    __slots__ = ("y", "scale", "shift", "mean", "invstd")

    def __init__(self, y, scale, shift, mean, invstd):
        self.y, self.scale, self.shift, self.mean, self.invstd = y, scale, shift, mean, invstd


def _conv_bn_train(cx: _Ctx, packs: PackCache, sd: Dict[str, torch.Tensor], conv: str, bn: str, x: torch.Tensor, B, H, W, cin_p,
                   cout, out: Optional[torch.Tensor], pooled: Optional[torch.Tensor], hilo: bool = False, stats=None,
                   up_into: Optional[torch.Tensor] = None) -> _BNSaved:
    """conv (+ batch statistics) -> finalize -> BN apply + ReLU into ``out`` (+ 2x2 max-pool into ``pooled``), or - when the
    activation is consumed by the x2 upsample only (``up_into`` = the concat slice at twice the resolution) - BN apply + ReLU +
    bilinear upsample in one pass, without ever storing the activation."""
    M = B * H * W
    y = cx.empty(M, cout, dtype=cx.raw)
    if stats is None:
        stats = cx.zeros(2 * cout, dtype=torch.float64)
    w = sd[conv + ".weight"]
    conv3x3(cx, x, packs.get(cx, conv, w, False, hilo), y, B, H, W, cin_p, cout, stats=stats,
            cin_real=w.shape[1] if w.shape[1] < 16 else 0)
    f32 = torch.float32
    scale, shift, mean, invstd = (cx.empty(cout, dtype=f32) for _ in range(4))
    call("eunet_bn_finalize", ptr(stats), M, ptr(sd[bn + ".weight"]), ptr(sd[bn + ".bias"]), ptr(sd.get(conv + ".bias")),
         ptr(sd[bn + ".running_mean"]), ptr(sd[bn + ".running_var"]), ptr(sd[bn + ".num_batches_tracked"]), BN_MOMENTUM, BN_EPS,
         ptr(scale), ptr(shift), ptr(mean), ptr(invstd), cout)
    if up_into is not None:
        call("eunet_bn_apply_relu_upsample2", ptr(y), _ld(y), ptr(up_into), _ld(up_into), cx.code, B, H, W, cout, ptr(scale), ptr(shift))
    else:
        call("eunet_bn_apply_relu", ptr(y), _ld(y), ptr(out), _ld(out), ptr(pooled), _ld(pooled) if pooled is not None else 0,
             cx.code, B, H, W, cout, ptr(scale), ptr(shift))
    return _BNSaved(y, scale, shift, mean, invstd)


def _conv_bn_eval(cx: _Ctx, packs: PackCache, sd, conv: str, bn: str, x, B, H, W, cin_p, cout, out, hilo: bool = False) -> None:
    f32 = torch.float32
    scale, shift = cx.empty(cout, dtype=f32), cx.empty(cout, dtype=f32)
    call("eunet_bn_fold_eval", ptr(sd[bn + ".weight"]), ptr(sd[bn + ".bias"]), ptr(sd[conv + ".bias"]),
         ptr(sd[bn + ".running_mean"]), ptr(sd[bn + ".running_var"]), BN_EPS, ptr(scale), ptr(shift), cout)
    w = sd[conv + ".weight"]
    conv3x3(cx, x, packs.get(cx, conv, w, False, hilo), out, B, H, W, cin_p, cout, scale=scale, shift=shift,
            relu=True, cin_real=w.shape[1] if w.shape[1] < 16 else 0)


class Saved:
    """Everything the backward pass needs (activations stay resident in HBM between fwd and bwd)."""

    def __init__(self):
        self.B = self.H = self.W = 0
        self.x16 = None
        self.bn: Dict[str, _BNSaved] = {}
        self.act: Dict[str, torch.Tensor] = {}


def _check_input(x: torch.Tensor) -> Tuple[int, int, int]:
    if x.dim() != 4 or x.shape[1] != 3:
        raise RuntimeError(f"EnhancedUNet expects input [B,3,H,W], got {tuple(x.shape)}")
    B, _, H, W = x.shape
    if H % 8 or W % 8 or H == 0 or W == 0 or B == 0:
        raise RuntimeError(f"EnhancedUNet needs H and W to be non-zero multiples of 8 (got {H}x{W})")
    if not x.is_cuda:
        raise RuntimeError("EnhancedUNet (B200) runs on CUDA tensors only; there is no CPU fallback")
    return B, H, W


def forward(sd: Dict[str, torch.Tensor], x: torch.Tensor, train: bool, act_dtype: torch.dtype, packs: PackCache,
            want_saved: bool, amax: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, Optional[Saved]]:
    """Returns logits [B,3,2H,2W] fp32 (NCHW) and, in train mode, the saved state for ``backward``.
    ``amax``: one device float; fp16 conv outputs beyond the fp16 range leave their largest magnitude there."""
    B, H, W = _check_input(x)
    cx = _Ctx(x.device, act_dtype)
    cx.amax = amax
    if x.dtype != torch.float32 or not x.is_contiguous():
        x = x.contiguous().float()
    zp64 = _ZeroPool(cx, STATS_NUMEL, torch.float64) if train else None
    sv = Saved()
    sv.B, sv.H, sv.W = B, H, W
    dims = [(H, W), (H // 2, W // 2), (H // 4, W // 4), (H // 8, W // 8)]
    Ms = [B * h * w for h, w in dims]

    packs.prepare(cx, sd, pack_specs(cx, train and want_saved))
    x16 = cx.empty(Ms[0], 16)
    hilo = _first_layer_split(cx)
    call("eunet_pack_input_nchw", ptr(x), ptr(x16), cx.code, B, 3, H, W, 16, int(hilo))
    sv.x16 = x16

    cat2 = cx.empty(Ms[0], 192)   # [up(d3) | e1]
    cat3 = cx.empty(Ms[1], 384)   # [up(d4) | e2]
    cat4 = cx.empty(Ms[2], 768)   # [up(e4) | e3]
    e1, e2, e3 = cat2[:, 128:192], cat3[:, 256:384], cat4[:, 512:768]
    p1, p2, p3 = cx.empty(Ms[1], 64), cx.empty(Ms[2], 128), cx.empty(Ms[3], 256)
    # e4 / d4 / d3 feed the upsample only: materialised in eval mode (conv epilogue output), never in training
    e4, d4, d3 = (None, None, None) if train else (cx.empty(Ms[3], 512), cx.empty(Ms[2], 256), cx.empty(Ms[1], 128))
    d2 = cx.empty(Ms[0], 64)

    def block(prefix, xin, lvl, cin_p, cout, out, pooled=None, up_into=None):
        """``up_into``: the block's activation is consumed by the x2 upsample only -> in training it is never stored (fused BN
        apply + ReLU + upsample); ``out`` is then only used in eval mode (conv epilogue -> ``out`` -> upsample)."""
        h, w = dims[lvl]
        mid = cx.empty(Ms[lvl], cout)
        first = hilo and prefix == "model.enc1"
        if train:
            sv.bn[prefix + ".1"] = _conv_bn_train(cx, packs, sd, prefix + ".0", prefix + ".1", xin, B, h, w, cin_p, cout, mid, None,
                                                  hilo=first, stats=zp64.take(2 * cout))
            sv.bn[prefix + ".4"] = _conv_bn_train(cx, packs, sd, prefix + ".3", prefix + ".4", mid, B, h, w, cout, cout, out, pooled,
                                                  stats=zp64.take(2 * cout), up_into=up_into)
            sv.act[prefix + ".in"] = xin
            sv.act[prefix + ".mid"] = mid
        else:
            _conv_bn_eval(cx, packs, sd, prefix + ".0", prefix + ".1", xin, B, h, w, cin_p, cout, mid, hilo=first)
            _conv_bn_eval(cx, packs, sd, prefix + ".3", prefix + ".4", mid, B, h, w, cout, cout, out)
            if pooled is not None:
                call("eunet_maxpool2_fwd", ptr(out), _ld(out), ptr(pooled), _ld(pooled), cx.code, B, h, w, cout)
            if up_into is not None:
                call("eunet_upsample2_fwd", ptr(out), _ld(out), ptr(up_into), _ld(up_into), cx.code, B, h, w, cout)

    block("model.enc1", x16, 0, 16, 64, e1, p1)
    block("model.enc2", p1, 1, 64, 128, e2, p2)
    block("model.enc3", p2, 2, 128, 256, e3, p3)
    block("model.enc4", p3, 3, 256, 512, e4, up_into=cat4[:, 0:512])
    block("model.dec4", cat4, 2, 768, 256, d4, up_into=cat3[:, 0:256])
    block("model.dec3", cat3, 1, 384, 128, d3, up_into=cat2[:, 0:128])
    block("model.dec2", cat2, 0, 192, 64, d2)

    # ---- tail: z = dec1(d2) @HxW, d1 = up(z), mid = enhance.0(d1), out = d1 + enhance.3(relu(bn(mid))) ----
    f32 = torch.float32
    M1, M2x = Ms[0], 4 * Ms[0]
    w1 = sd["model.dec1.weight"].reshape(3, 64)
    w3 = sd["enhance.3.weight"].reshape(3, 64)
    z4 = cx.empty(M1, 4, dtype=f32)
    call("eunet_tail_dec1_fwd", ptr(d2), _ld(d2), cx.code, ptr(w1), ptr(sd["model.dec1.bias"]), ptr(z4), M1)
    d1p = cx.empty(M2x, 16)
    d14 = cx.empty(M2x, 4, dtype=f32)
    call("eunet_tail_up_fwd", ptr(z4), ptr(d1p), ptr(d14), cx.code, B, H, W)
    out = torch.empty(B, 3, 2 * H, 2 * W, device=x.device, dtype=f32)
    wp0 = packs.get(cx, "enhance.0", sd["enhance.0.weight"], False)
    # tensor-core path: BN + ReLU + enhance.3 + residual in the conv epilogue.  Inference only: in training the 64-channel
    # tensor has to be stored for the backward pass anyway and the separate bandwidth pass is faster than the heavier
    # epilogue (measured: 1.53 ms fused + 0.25 ms statistics pass against 0.49 + 0.79 ms).
    fused = cx.tc and ((not train) if FUSED_TAIL is None else bool(FUSED_TAIL))
    if train:
        stats = zp64.take(128)
        midt = cx.empty(M2x, 64, dtype=cx.raw) if (want_saved or not fused) else None
        if fused:   # pass 1: batch statistics only (nothing stored); pass 2 below recomputes the tiles
            call("eunet_conv3x3_fwd", ptr(d1p), 16, ptr(wp0), None, 64, cx.code, B, 2 * H, 2 * W, 16, 64, ptr(stats), None, None, 0, 1,
                 None, flops=2.0 * M2x * 64 * 27, tag="k27")
        else:
            conv3x3(cx, d1p, wp0, midt, B, 2 * H, 2 * W, 16, 64, stats=stats, cin_real=3)
        scale, shift, mean, invstd = (cx.empty(64, dtype=f32) for _ in range(4))
        call("eunet_bn_finalize", ptr(stats), M2x, ptr(sd["enhance.1.weight"]), ptr(sd["enhance.1.bias"]),
             ptr(sd["enhance.0.bias"]), ptr(sd["enhance.1.running_mean"]), ptr(sd["enhance.1.running_var"]),
             ptr(sd["enhance.1.num_batches_tracked"]), BN_MOMENTUM, BN_EPS, ptr(scale), ptr(shift), ptr(mean), ptr(invstd), 64)
        sv.bn["enhance.1"] = _BNSaved(midt, scale, shift, mean, invstd)
    else:
        scale, shift = cx.empty(64, dtype=f32), cx.empty(64, dtype=f32)
        call("eunet_bn_fold_eval", ptr(sd["enhance.1.weight"]), ptr(sd["enhance.1.bias"]), ptr(sd["enhance.0.bias"]),
             ptr(sd["enhance.1.running_mean"]), ptr(sd["enhance.1.running_var"]), BN_EPS, ptr(scale), ptr(shift), 64)
        midt = None
    if fused:
        call("eunet_conv3x3_tail_fwd", ptr(d1p), ptr(wp0), ptr(midt), ptr(scale), ptr(shift), ptr(w3), ptr(sd["enhance.3.bias"]),
             ptr(d14), ptr(out), cx.code, B, 2 * H, 2 * W, flops=2.0 * M2x * 64 * 27, tag="k27")
    else:
        if not train:   # fp32 mode, eval: conv with the folded affine + ReLU, then the 1x1 + residual pass
            midt = cx.empty(M2x, 64, dtype=cx.raw)
            conv3x3(cx, d1p, wp0, midt, B, 2 * H, 2 * W, 16, 64, scale=scale, shift=shift, relu=True, cin_real=3)
            scale, shift = torch.ones(64, device=x.device, dtype=f32), torch.zeros(64, device=x.device, dtype=f32)
        call("eunet_tail_out_fwd", ptr(d14), ptr(midt), cx.code, ptr(scale), ptr(shift), ptr(w3), ptr(sd["enhance.3.bias"]), ptr(out),
             B, H, W)
    if train and want_saved:
        sv.act.update(dict(cat2=cat2, cat3=cat3, cat4=cat4, d2=d2, d1p=d1p, z4=z4))
        return out, sv
    return out, None


GRAD_SCALE_TARGET = 16.0   # fp16 mode: the largest |dLoss/dlogit| is scaled to [8, 16]: 2^12 of headroom, 2^28 below


def backward(sd: Dict[str, torch.Tensor], sv: Saved, dout: torch.Tensor, act_dtype: torch.dtype, packs: PackCache,
             sink: Optional[GradSink] = None, amax: Optional[torch.Tensor] = None) -> Dict[str, torch.Tensor]:
    """Gradients (fp32, parameter layout) for every parameter of the state_dict, written through ``sink`` in
    ``grad_production_order()``."""
    B, H, W = sv.B, sv.H, sv.W
    cx = _Ctx(dout.device, act_dtype)
    cx.amax = amax
    f32, f64 = torch.float32, torch.float64
    dims = [(H, W), (H // 2, W // 2), (H // 4, W // 4), (H // 8, W // 8)]
    Ms = [B * h * w for h, w in dims]
    if sink is None:
        sink = GradSink(dout.device)
    grads = sink.grads
    pool = _ZeroPool(cx, wgrad_workspace_numel())
    zp64 = _ZeroPool(cx, STATS_NUMEL, f64)
    zbias = _ZeroPool(cx, ZERO_BIAS_NUMEL)
    if dout.dtype != f32 or not dout.is_contiguous():
        dout = dout.contiguous().float()
    if cx.dt == torch.float16:
        # tcgen05 takes both MMA operands in one 16-bit format, so gradients are fp16 too; ONE power-of-two scale per pass
        # (backward is linear in dout) keeps them in range: applied where dout is read, removed where fp32 gradients leave
        cx.gs = cx.empty(4, dtype=f32)
        call("eunet_grad_scale", ptr(dout), dout.numel(), GRAD_SCALE_TARGET, ptr(cx.gs))
    gs = cx.gs

    def cast64(name: str, src: torch.Tensor, shape) -> None:
        dst = sink.dst(name, shape)
        call("eunet_cast_f64_f32", ptr(src), ptr(dst), dst.numel(), ptr(gs))
        ready(name)

    def zero_bias(name: str, c: int) -> None:      # conv bias in front of a train-mode BN: exact zero gradient
        sink.zero_dst(name, (c,), zbias)
        ready(name)

    deferred = []          # filter gradients waiting for the next multi-tensor unpack launch

    def ready(name: str) -> None:
        if deferred and sink.closes(name):
            unpack_wgrad_multi(deferred, gs)
            deferred.clear()
        sink.ready(name)

    def wgrad_into(name: str, dwp: torch.Tensor, co: int, ci: int, hilo: bool = False) -> None:
        deferred.append((dwp, sink.dst(name, (co, ci, 3, 3)), co, ci, hilo))
        ready(name)

    # ---- tail ----
    M1, M2x = Ms[0], 4 * Ms[0]
    bn = sv.bn["enhance.1"]
    w1 = sd["model.dec1.weight"].reshape(3, 64)
    w3 = sd["enhance.3.weight"].reshape(3, 64)
    acc = zp64.take(328)
    dout4 = cx.empty(M2x, 4, dtype=f32)
    call("eunet_tail_pack3", ptr(dout), ptr(dout4), B, 2 * H, 2 * W, ptr(gs))
    call("eunet_tail_bwd_reduce", ptr(dout4), ptr(bn.y), cx.code, ptr(bn.scale), ptr(bn.shift), ptr(bn.mean), ptr(bn.invstd),
         ptr(w3), ptr(acc), B, H, W)
    cast64("enhance.1.bias", acc[0:64], (64,))
    cast64("enhance.1.weight", acc[64:128], (64,))
    cast64("enhance.3.weight", acc[128:320], (3, 64, 1, 1))
    cast64("enhance.3.bias", acc[320:323], (3,))
    d1p = sv.act["d1p"]
    dz4 = cx.empty(M1, 4, dtype=f32)
    wflip = packs.get(cx, "enhance.0", sd["enhance.0.weight"], True)
    tc_path = cx.tc and 2 * H >= 8 and 2 * W >= 8
    if tc_path and FUSED_TAIL_BWD:
        # ONE kernel: BN/ReLU backward on chip, wgrad and the 3-channel transposed dgrad from the same staged tile
        dwp = pool.take(64, 9, 16)
        dd1 = cx.empty(M2x, 4, dtype=f32)
        call("eunet_tail_bwd_fused", ptr(dout4), ptr(bn.y), ptr(d1p), ptr(wflip), ptr(bn.scale), ptr(bn.shift), ptr(bn.mean),
             ptr(bn.invstd), ptr(w3), ptr(acc), ptr(dd1), ptr(dwp), cx.code, B, 2 * H, 2 * W, flops=2 * 2.0 * M2x * 64 * 27,
             tag="k27")
        wgrad_into("enhance.0.weight", dwp, 64, 3)
        zero_bias("enhance.0.bias", 64)                                        # cancelled by train-mode BN
        call("eunet_tail_up_bwd", ptr(dd1), lib.F32, 4, ptr(dout), ptr(dz4), B, H, W, ptr(gs))
    else:
        dmid = cx.empty(M2x, 64)
        call("eunet_tail_bwd_dmid", ptr(dout4), ptr(bn.y), ptr(dmid), cx.code, ptr(bn.scale), ptr(bn.shift), ptr(bn.mean),
             ptr(bn.invstd), ptr(w3), ptr(acc), B, H, W)
        dwp = conv3x3_wgrad(cx, d1p, dmid, B, 2 * H, 2 * W, 16, 64, pool, cin_real=3)
        wgrad_into("enhance.0.weight", dwp, 64, 3)
        zero_bias("enhance.0.bias", 64)                                        # cancelled by train-mode BN
        if tc_path:
            # 3 real gradient channels: transposed dgrad (every dmid row read once), fp32 [pixels][4] output
            dd1 = cx.empty(M2x, 4, dtype=f32)
            call("eunet_conv3x3_dgrad_few", ptr(dmid), _ld(dmid), ptr(wflip), ptr(dd1), cx.code, B, 2 * H, 2 * W, 64, 16,
                 flops=2.0 * M2x * 64 * 27, tag="k27")
            call("eunet_tail_up_bwd", ptr(dd1), lib.F32, 4, ptr(dout), ptr(dz4), B, H, W, ptr(gs))
        else:
            dd1p = cx.empty(M2x, 16)
            conv3x3(cx, dmid, wflip, dd1p, B, 2 * H, 2 * W, 64, 16, cout_real=3)
            call("eunet_tail_up_bwd", ptr(dd1p), cx.code, 16, ptr(dout), ptr(dz4), B, H, W, ptr(gs))
    acc2 = zp64.take(200)
    d2 = sv.act["d2"]
    dd2 = cx.empty(M1, 64)
    call("eunet_tail_dec1_bwd", ptr(dz4), ptr(d2), _ld(d2), ptr(dd2), _ld(dd2), cx.code, ptr(w1), ptr(acc2), M1)
    cast64("model.dec1.weight", acc2[0:192], (3, 64, 1, 1))
    cast64("model.dec1.bias", acc2[192:195], (3,))

    def bn_bwd(name: str, dact: torch.Tensor, M: int, C: int, dpool: Optional[torch.Tensor] = None, hw=None) -> torch.Tensor:
        """BN + ReLU backward of ``name``.  ``dpool``: the activation also fed a 2x2 max-pool whose gradient this is; the
        two passes then form dact + route(dpool) on the fly (no max-pool backward pass over the skip slice)."""
        s = sv.bn[name]
        sums = zp64.take(2 * C)
        dy = cx.empty(M, C)
        dg, db = sink.dst(name + ".weight", (C,)), sink.dst(name + ".bias", (C,))
        if dpool is not None:
            h, w = hw
            call("eunet_bn_bwd_reduce_pool", ptr(dact), _ld(dact), ptr(dpool), _ld(dpool), ptr(s.y), _ld(s.y), cx.code, B, h, w, C,
                 ptr(s.scale), ptr(s.shift), ptr(s.mean), ptr(s.invstd), ptr(sums))
            call("eunet_bn_bwd_apply_pool", ptr(dact), _ld(dact), ptr(dpool), _ld(dpool), ptr(s.y), _ld(s.y), ptr(dy), _ld(dy), cx.code,
                 B, h, w, C, ptr(s.scale), ptr(s.shift), ptr(s.mean), ptr(s.invstd), ptr(sums), ptr(dg), ptr(db), ptr(gs))
            ready(name + ".weight")
            ready(name + ".bias")
            return dy
        call("eunet_bn_bwd_reduce", ptr(dact), _ld(dact), ptr(s.y), _ld(s.y), cx.code, M, C, ptr(s.scale), ptr(s.shift),
             ptr(s.mean), ptr(s.invstd), ptr(sums))
        call("eunet_bn_bwd_apply", ptr(dact), _ld(dact), ptr(s.y), _ld(s.y), ptr(dy), _ld(dy), cx.code, M, C, ptr(s.scale),
             ptr(s.shift), ptr(s.mean), ptr(s.invstd), ptr(sums), ptr(dg), ptr(db), ptr(gs))
        ready(name + ".weight")
        ready(name + ".bias")
        return dy

    def block_bwd(prefix: str, dact: torch.Tensor, lvl: int, cin: int, cout: int, need_dx: bool,
                  dpool: Optional[torch.Tensor] = None) -> Optional[torch.Tensor]:
        h, w = dims[lvl]
        M = Ms[lvl]
        cin_p = _pad16(cin)
        xin, mid = sv.act[prefix + ".in"], sv.act[prefix + ".mid"]
        dy_b = bn_bwd(prefix + ".4", dact, M, cout, dpool=dpool, hw=(h, w))
        wgrad_into(prefix + ".3.weight", conv3x3_wgrad(cx, mid, dy_b, B, h, w, cout, cout, pool), cout, cout)
        zero_bias(prefix + ".3.bias", cout)
        dmid_act = cx.empty(M, cout)
        conv3x3(cx, dy_b, packs.get(cx, prefix + ".3", sd[prefix + ".3.weight"], True), dmid_act, B, h, w, cout, cout)
        dy_a = bn_bwd(prefix + ".1", dmid_act, M, cout)
        wgrad_into(prefix + ".0.weight", conv3x3_wgrad(cx, xin, dy_a, B, h, w, cin_p, cout, pool, cin_real=cin if cin < 16 else 0), cout, cin,
                   hilo=(prefix == "model.enc1" and _first_layer_split(cx)))
        zero_bias(prefix + ".0.bias", cout)
        if not need_dx:
            return None
        dx = cx.empty(M, cin_p)
        conv3x3(cx, dy_a, packs.get(cx, prefix + ".0", sd[prefix + ".0.weight"], True), dx, B, h, w, cout, cin_p)
        return dx

    def up_bwd(dsrc: torch.Tensor, lvl_in: int, c: int) -> torch.Tensor:
        h, w = dims[lvl_in]
        dx = cx.empty(Ms[lvl_in], c)
        call("eunet_upsample2_bwd", ptr(dsrc), _ld(dsrc), ptr(dx), _ld(dx), cx.code, B, h, w, c)
        return dx

    dcat2 = block_bwd("model.dec2", dd2, 0, 192, 64, True)
    dd3 = up_bwd(dcat2[:, 0:128], 1, 128)
    dcat3 = block_bwd("model.dec3", dd3, 1, 384, 128, True)
    dd4 = up_bwd(dcat3[:, 0:256], 2, 256)
    dcat4 = block_bwd("model.dec4", dd4, 2, 768, 256, True)
    de4 = up_bwd(dcat4[:, 0:512], 3, 512)
    # encoder outputs fed the decoder (skip slice of the concat gradient) AND the max-pool (dp*): both gradients are summed
    # inside the BN backward passes of the block's last layer
    dp3 = block_bwd("model.enc4", de4, 3, 256, 512, True)
    dp2 = block_bwd("model.enc3", dcat4[:, 512:768], 2, 128, 256, True, dpool=dp3)
    dp1 = block_bwd("model.enc2", dcat3[:, 256:384], 1, 64, 128, True, dpool=dp2)
    block_bwd("model.enc1", dcat2[:, 128:192], 0, 3, 64, False, dpool=dp1)
    if deferred:
        unpack_wgrad_multi(deferred, gs)
        deferred.clear()
    return grads


def _fold_bn_host(sd, prefix: str):
    """Eval-mode BN -> (scale, shift) on the host (tiny vectors, used for the kernel-parameter block)."""
    g, b = sd[prefix + ".weight"].detach().float().cpu(), sd[prefix + ".bias"].detach().float().cpu()
    rm, rv = sd[prefix + ".running_mean"].float().cpu(), sd[prefix + ".running_var"].float().cpu()
    sc = g / torch.sqrt(rv + BN_EPS)
    return sc, b - rm * sc


_FUSION_BLOB_KEYS = ("attention_gate.0.weight", "attention_gate.1.weight", "attention_gate.1.bias", "attention_gate.1.running_mean",
                     "attention_gate.1.running_var", "attention_gate.3.weight", "attention_gate.4.weight", "attention_gate.4.bias",
                     "attention_gate.4.running_mean", "attention_gate.4.running_var", "fusion_residual.weight", "fusion_residual.bias")


def fusion_forward(sd: Dict[str, torch.Tensor], out_main: torch.Tensor, out_aux: torch.Tensor, act_dtype: torch.dtype,
                   packs: PackCache, blob_cache: Optional[dict] = None) -> torch.Tensor:
    """Reference models.py:320-328 in eval mode.  Returns [B,3,H,W] fp32.  The 219-float kernel-parameter block of the gate
    (folded BN included) is built on the host; ``blob_cache`` keeps it until one of its source tensors changes (version
    counters), so that steady-state inference does no device->host copies."""
    import ctypes
    B, _, H, W = out_main.shape
    cx = _Ctx(out_main.device, act_dtype)
    f32 = torch.float32
    M = B * H * W
    key = tuple((sd[k]._version, sd[k].data_ptr()) for k in _FUSION_BLOB_KEYS)
    blob = blob_cache.get("blob") if blob_cache is not None and blob_cache.get("key") == key else None
    if blob is None:
        s1, h1 = _fold_bn_host(sd, "attention_gate.1")
        s4, h4 = _fold_bn_host(sd, "attention_gate.4")
        blob = torch.cat([sd["attention_gate.0.weight"].detach().float().cpu().reshape(-1), s1, h1,
                          sd["attention_gate.3.weight"].detach().float().cpu().reshape(-1), s4, h4,
                          sd["fusion_residual.weight"].detach().float().cpu().reshape(-1),
                          sd["fusion_residual.bias"].detach().float().cpu().reshape(-1)]).contiguous()
        if blob_cache is not None:
            blob_cache["key"], blob_cache["blob"] = key, blob
    assert blob.numel() == 219
    fg16 = cx.empty(M, 16)
    res4 = cx.empty(M, 4, dtype=f32)
    call("eunet_fusion_gate_fwd", ptr(out_main.contiguous().float()), ptr(out_aux.contiguous().float()),
         ctypes.c_void_p(blob.data_ptr()), ptr(fg16), cx.code, ptr(res4), B, H, W)
    h = fg16
    cin_p = 16
    for conv, bn, cout in (("fusion_head.0", "fusion_head.1", 256), ("fusion_head.4", "fusion_head.5", 128),
                           ("fusion_head.8", "fusion_head.9", 64)):
        scale, shift = cx.empty(cout, dtype=f32), cx.empty(cout, dtype=f32)
        call("eunet_bn_fold_eval", ptr(sd[bn + ".weight"]), ptr(sd[bn + ".bias"]), None, ptr(sd[bn + ".running_mean"]),
             ptr(sd[bn + ".running_var"]), BN_EPS, ptr(scale), ptr(shift), cout)
        nxt = cx.empty(M, cout)
        conv3x3(cx, h, packs.get(cx, conv, sd[conv + ".weight"], False), nxt, B, H, W, cin_p, cout, scale=scale, shift=shift,
                relu=True)
        h, cin_p = nxt, cout
    z4 = cx.empty(M, 4, dtype=f32)
    w11 = sd["fusion_head.11.weight"].reshape(3, 64)
    call("eunet_tail_dec1_fwd", ptr(h), _ld(h), cx.code, ptr(w11), ptr(sd["fusion_head.11.bias"]), ptr(z4), M)
    out = torch.empty(B, 3, H, W, device=out_main.device, dtype=f32)
    call("eunet_fusion_out_fwd", ptr(z4), ptr(res4), ptr(out), B, H, W)
    return out


# ---------------------------------------------------------------------------------------------
# fusion blocks of the smp body (reference models.py:276-302, 320-328) in TRAINING mode + backward
# ---------------------------------------------------------------------------------------------
FUSION_HEAD = (("fusion_head.0", "fusion_head.1", 6, 256, 0), ("fusion_head.4", "fusion_head.5", 256, 128, 1),
               ("fusion_head.8", "fusion_head.9", 128, 64, None))     # (conv, bn, Cin, Cout, index of the Dropout2d factor after it)


def _gate_blob(sd: Dict[str, torch.Tensor]) -> torch.Tensor:
    """HOST blob of the small gate / residual filters (kernel parameter block): w0[3][6][9], w3[6][3], wr[3][6], br[3]."""
    blob = torch.cat([sd["attention_gate.0.weight"].detach().float().cpu().reshape(-1),
                      sd["attention_gate.3.weight"].detach().float().cpu().reshape(-1),
                      sd["fusion_residual.weight"].detach().float().cpu().reshape(-1),
                      sd["fusion_residual.bias"].detach().float().cpu().reshape(-1)]).contiguous()
    assert blob.numel() == 201
    return blob


class FusionSaved:
    def __init__(self):
        self.B = self.H = self.W = 0
        self.t: Dict[str, torch.Tensor] = {}
        self.bn: Dict[str, _BNSaved] = {}
        self.scales = None


def fusion_train_forward(sd: Dict[str, torch.Tensor], out_main: torch.Tensor, out_aux: torch.Tensor, act_dtype: torch.dtype,
                         packs: PackCache, scales, amax: Optional[torch.Tensor] = None):
    """Training-mode forward of the fusion blocks; ``scales`` = (s1 [B,256], s2 [B,128]) fp32 Dropout2d factors
    (keep / (1 - p)).  Returns (out [B,3,H,W] fp32, saved state for ``fusion_backward``)."""
    import ctypes
    B, _, H, W = out_main.shape
    cx = _Ctx(out_main.device, act_dtype)
    cx.amax = amax
    f32 = torch.float32
    M = B * H * W
    main, aux = out_main.contiguous().float(), out_aux.contiguous().float()
    blob = _gate_blob(sd)
    bp = ctypes.c_void_p(blob.data_ptr())
    zp64 = _ZeroPool(cx, 2 * (3 + 6 + 256 + 128 + 64) + 64 * 8, torch.float64)
    sv = FusionSaved()
    sv.B, sv.H, sv.W, sv.scales = B, H, W, scales

    def finalize(stats, bn, C):
        c = cx.empty(4 * C, dtype=f32)        # {scale, shift, mean, invstd}
        call("eunet_bn_finalize", ptr(stats), M, ptr(sd[bn + ".weight"]), ptr(sd[bn + ".bias"]), None, ptr(sd[bn + ".running_mean"]),
             ptr(sd[bn + ".running_var"]), ptr(sd[bn + ".num_batches_tracked"]), BN_MOMENTUM, BN_EPS, ptr(c[0:C]), ptr(c[C:2 * C]),
             ptr(c[2 * C:3 * C]), ptr(c[3 * C:4 * C]), C)
        return c

    a1 = cx.empty(M, 4, dtype=f32)
    st1 = zp64.take(6)
    call("eunet_fusion_gate_conv_fwd", ptr(main), ptr(aux), bp, ptr(a1), ptr(st1), B, H, W)
    bn1 = finalize(st1, "attention_gate.1", 3)
    a2 = cx.empty(M, 8, dtype=f32)
    st2 = zp64.take(12)
    call("eunet_fusion_gate_mid_fwd", ptr(a1), ptr(bn1[0:3]), ptr(bn1[3:6]), bp, ptr(a2), ptr(st2), M)
    bn2 = finalize(st2, "attention_gate.4", 6)
    fg16 = cx.empty(M, 16)
    res4 = cx.empty(M, 4, dtype=f32)
    call("eunet_fusion_gate_apply_fwd", ptr(main), ptr(aux), ptr(a2), ptr(bn2[0:6]), ptr(bn2[6:12]), bp, ptr(fg16), cx.code, ptr(res4),
         B, H, W)
    sv.t.update(main=main, aux=aux, a1=a1, a2=a2, bn1=bn1, bn2=bn2)
    h, cin_p = fg16, 16
    for conv, bn, cin, cout, drop in FUSION_HEAD:
        nxt = cx.empty(M, cout)
        sv.bn[bn] = _conv_bn_train(cx, packs, sd, conv, bn, h, B, H, W, cin_p, cout, nxt, None, stats=zp64.take(2 * cout))
        if drop is not None:
            sc = scales[drop].contiguous().float()
            assert sc.shape == (B, cout)
            call("eunet_channel_scale", ptr(nxt), _ld(nxt), ptr(sc), cx.code, B, H * W, cout)
        sv.t[conv + ".in"] = h
        h, cin_p = nxt, cout
    sv.t["h3"] = h
    z4 = cx.empty(M, 4, dtype=f32)
    w11 = sd["fusion_head.11.weight"].reshape(3, 64)
    call("eunet_tail_dec1_fwd", ptr(h), _ld(h), cx.code, ptr(w11), ptr(sd["fusion_head.11.bias"]), ptr(z4), M)
    out = torch.empty(B, 3, H, W, device=main.device, dtype=f32)
    call("eunet_fusion_out_fwd", ptr(z4), ptr(res4), ptr(out), B, H, W)
    return out, sv


def fusion_backward(sd: Dict[str, torch.Tensor], sv: FusionSaved, dout: torch.Tensor, act_dtype: torch.dtype, packs: PackCache,
                    amax: Optional[torch.Tensor] = None):
    """Returns (dmain, daux, {parameter name: fp32 gradient}) for ``fusion_train_forward``."""
    import ctypes
    B, H, W = sv.B, sv.H, sv.W
    cx = _Ctx(dout.device, act_dtype)
    cx.amax = amax
    f32, f64 = torch.float32, torch.float64
    M = B * H * W
    dout = dout.contiguous().float()
    if cx.dt == torch.float16:
        cx.gs = cx.empty(4, dtype=f32)
        call("eunet_grad_scale", ptr(dout), dout.numel(), GRAD_SCALE_TARGET, ptr(cx.gs))
    gs = cx.gs
    grads: Dict[str, torch.Tensor] = {}
    zp64 = _ZeroPool(cx, 2 * (256 + 128 + 64) + 200 + 219 + 64 * 8, f64)
    pool = _ZeroPool(cx, 256 * 9 * 16 + 128 * 9 * 256 + 64 * 9 * 128 + 256)

    def cast64(name, src, shape):
        dst = torch.empty(shape, device=dout.device, dtype=f32)
        call("eunet_cast_f64_f32", ptr(src), ptr(dst), dst.numel(), ptr(gs))
        grads[name] = dst

    dout4 = cx.empty(M, 4, dtype=f32)
    call("eunet_tail_pack3", ptr(dout), ptr(dout4), B, H, W, ptr(gs))
    h3 = sv.t["h3"]
    acc2 = zp64.take(200)
    dact = cx.empty(M, 64)
    w11 = sd["fusion_head.11.weight"].reshape(3, 64)
    call("eunet_tail_dec1_bwd", ptr(dout4), ptr(h3), _ld(h3), ptr(dact), _ld(dact), cx.code, ptr(w11), ptr(acc2), M)
    cast64("fusion_head.11.weight", acc2[0:192], (3, 64, 1, 1))
    cast64("fusion_head.11.bias", acc2[192:195], (3,))
    for conv, bn, cin, cout, drop in reversed(FUSION_HEAD):
        cin_p = _pad16(cin)
        if drop is not None:     # gradient through Dropout2d: the same per-(sample, channel) factor
            sc = sv.scales[drop].contiguous().float()
            call("eunet_channel_scale", ptr(dact), _ld(dact), ptr(sc), cx.code, B, H * W, cout)
        s = sv.bn[bn]
        sums = zp64.take(2 * cout)
        call("eunet_bn_bwd_reduce", ptr(dact), _ld(dact), ptr(s.y), _ld(s.y), cx.code, M, cout, ptr(s.scale), ptr(s.shift), ptr(s.mean),
             ptr(s.invstd), ptr(sums))
        dy = cx.empty(M, cout)
        dg, db = torch.empty(cout, device=dout.device, dtype=f32), torch.empty(cout, device=dout.device, dtype=f32)
        call("eunet_bn_bwd_apply", ptr(dact), _ld(dact), ptr(s.y), _ld(s.y), ptr(dy), _ld(dy), cx.code, M, cout, ptr(s.scale),
             ptr(s.shift), ptr(s.mean), ptr(s.invstd), ptr(sums), ptr(dg), ptr(db), ptr(gs))
        grads[bn + ".weight"], grads[bn + ".bias"] = dg, db
        xin = sv.t[conv + ".in"]
        dwp = conv3x3_wgrad(cx, xin, dy, B, H, W, cin_p, cout, pool, cin_real=cin if cin < 16 else 0)
        grads[conv + ".weight"] = unpack_wgrad(dwp, cout, cin, gscale=gs)
        dx = cx.empty(M, cin_p)
        conv3x3(cx, dy, packs.get(cx, conv, sd[conv + ".weight"], True), dx, B, H, W, cout, cin_p, cout_real=cin if cin < 16 else 0)
        dact = dx
    blob = _gate_blob(sd)
    acc = zp64.take(219)
    dz1, dz2 = cx.empty(M, 4, dtype=f32), cx.empty(M, 8, dtype=f32)
    dmain, daux = torch.empty(B, 3, H, W, device=dout.device, dtype=f32), torch.empty(B, 3, H, W, device=dout.device, dtype=f32)
    t = sv.t
    call("eunet_fusion_gate_bwd", ptr(t["main"]), ptr(t["aux"]), ptr(t["a1"]), ptr(t["a2"]), ptr(t["bn1"]), ptr(t["bn2"]), ptr(dact),
         cx.code, ptr(dout4), ctypes.c_void_p(blob.data_ptr()), ptr(dz1), ptr(dz2), ptr(acc), ptr(dmain), ptr(daux), ptr(gs), B, H, W)
    cast64("attention_gate.4.bias", acc[0:6], (6,))
    cast64("attention_gate.4.weight", acc[6:12], (6,))
    cast64("fusion_residual.weight", acc[12:30], (3, 6, 1, 1))
    cast64("fusion_residual.bias", acc[30:33], (3,))
    cast64("attention_gate.1.bias", acc[33:36], (3,))
    cast64("attention_gate.1.weight", acc[36:39], (3,))
    cast64("attention_gate.3.weight", acc[39:57], (6, 3, 1, 1))
    cast64("attention_gate.0.weight", acc[57:219], (3, 6, 3, 3))
    return dmain, daux, grads
