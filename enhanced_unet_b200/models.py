"""Drop-in for the hot-path classes of the reference ``models.py``.

``EnhancedUNet(num_classes=3)`` keeps the reference constructor, ``forward(x)`` signature,
``get_aux_outputs()``, attribute ``num_classes`` and - bit for bit - the 109-key ``state_dict`` of the
reference's fallback body (``UNet(num_classes).model`` = BasicUNet, reference models.py:199-238, plus the
``enhance`` head, 308-313), so reference checkpoints load unchanged and the reference Trainer / optimizer
/ clip code works on its ``nn.Parameter``s.  The arithmetic runs in libeunet_b200.so (sm_100a); the torch
modules below only HOLD parameters (same construction order => same default init for a given seed).

There is no CPU path: calling ``forward`` on a non-CUDA tensor raises.
"""
from __future__ import annotations

from typing import Dict, Optional

import torch
import torch.nn as nn

from . import engine

_DTYPES = {"fp16": torch.float16, "float16": torch.float16, "bf16": torch.bfloat16, "bfloat16": torch.bfloat16,
           "fp32": torch.float32, "float32": torch.float32}
DEFAULT_DTYPE = "fp16"
FP16_MAX = 65504.0


class _SaturationMonitor:
    """fp16 tensors saturate at +-65504 (``cvt.rn.satfinite``) instead of overflowing to inf.  That must not be silent:
    the conv epilogues leave the largest magnitude they had to clamp in one device float; it is copied to pinned host
    memory after every pass without a sync and inspected at the start of the next pass (or by ``check()``)."""

    def __init__(self):
        self.dev = None
        self.host = None
        self.event = None
        self.device = None

    def buffer(self, device: torch.device) -> torch.Tensor:
        if self.dev is None or self.device != device:
            self.device = device
            self.dev = torch.zeros(1, device=device, dtype=torch.float32)
            self.host = torch.zeros(1, dtype=torch.float32).pin_memory()
            self.event = None
        return self.dev

    def publish(self) -> None:
        if torch.cuda.is_current_stream_capturing():
            return          # inside a CUDA-graph capture: graph.GraphedTrainStep publishes after each replay
        if self.dev is not None:
            self.host.copy_(self.dev, non_blocking=True)
            self.event = torch.cuda.Event()
            self.event.record()

    def check(self, block: bool) -> None:
        if self.event is None or torch.cuda.is_current_stream_capturing():
            return
        if block:
            self.event.synchronize()
        elif not self.event.query():
            return
        v = float(self.host[0])
        if v > FP16_MAX:
            self.dev.zero_()
            self.host.zero_()
            raise RuntimeError(f"EnhancedUNet (fp16 mode): convolution outputs exceeded the fp16 range (max |y| = {v:.4g} > "
                               f"{FP16_MAX:.0f}) and were clamped; results of that pass are invalid. Use dtype='bf16' or 'fp32' "
                               "for weights / inputs of this magnitude.")


def _conv_block(in_ch: int, out_ch: int) -> nn.Sequential:
    # parameter container with the reference's module indices (0 conv, 1 bn, 2 relu, 3 conv, 4 bn, 5 relu)
    return nn.Sequential(
        nn.Conv2d(in_ch, out_ch, 3, padding=1), nn.BatchNorm2d(out_ch), nn.ReLU(inplace=True),
        nn.Conv2d(out_ch, out_ch, 3, padding=1), nn.BatchNorm2d(out_ch), nn.ReLU(inplace=True))


class _BasicUNetParams(nn.Module):
    """Holds the parameters of the reference BasicUNet (models.py:200-215) under the same names."""

    def __init__(self, num_classes: int):
        super().__init__()
        self.enc1 = _conv_block(3, 64)
        self.enc2 = _conv_block(64, 128)
        self.enc3 = _conv_block(128, 256)
        self.enc4 = _conv_block(256, 512)
        self.dec4 = _conv_block(512 + 256, 256)
        self.dec3 = _conv_block(256 + 128, 128)
        self.dec2 = _conv_block(128 + 64, 64)
        self.dec1 = nn.Conv2d(64, num_classes, 1)

    def forward(self, x):  # pragma: no cover - never used as a compute path
        raise RuntimeError("BasicUNet parameters are executed by EnhancedUNet.forward (CUDA kernels), not directly")


class _UNetFunction(torch.autograd.Function):
    """Whole-network autograd node: forward and backward are fixed kernel schedules (engine.py)."""

    @staticmethod
    def forward(ctx, module: "EnhancedUNet", x: torch.Tensor, *params: torch.Tensor):
        sd = module._tensor_dict()
        out, saved = engine.forward(sd, x, True, module.act_dtype, module._packs, want_saved=True, amax=module._amax_buffer(x.device))
        module._sat.publish()
        ctx.module = module
        ctx.saved_state = saved
        ctx.n_params = len(params)
        return out

    @staticmethod
    def backward(ctx, dout: torch.Tensor):
        module = ctx.module
        sv = ctx.saved_state
        if sv is None:
            raise RuntimeError("EnhancedUNet backward called twice (saved activations were released)")
        sd = module._tensor_dict()
        sink = module.grad_sink
        if sink is not None:
            sink.begin()
        grads = engine.backward(sd, sv, dout, module.act_dtype, module._packs, sink=sink, amax=module._amax_buffer(dout.device))
        module._sat.publish()
        ctx.saved_state = None
        names = module._param_names
        if sink is None:
            return (None, None) + tuple(grads[n] for n in names)
        # data-parallel path: .grad IS the slice of the flat buffer the (possibly still running) all-reduce works on;
        # handing the views to autograd would let AccumulateGrad clone them before the exchange has finished
        params = dict(module.named_parameters())
        for n in names:
            p = params[n]
            if p.grad is not None:
                raise RuntimeError("EnhancedUNet with a grad_sink (data-parallel) needs gradients cleared before backward "
                                   "(optimizer.zero_grad(set_to_none=True)); accumulation across backward passes is not "
                                   "supported on this path")
            p.grad = grads[n]
        return (None, None) + (None,) * len(names)


class EnhancedUNet(nn.Module):
    """Reference models.py:246-343 (fallback body).  Extra keyword ``dtype``: 'fp16' (default) / 'bf16' - 16-bit tensors,
    tcgen05 tensor-core convolutions, fp32 accumulation / statistics - or 'fp32' (CUDA-core fp32 mode).  fp16's 11
    mantissa bits keep train-mode (batch-statistics) BatchNorm inside the 2e-2 logit tolerance where bf16's 8 do not;
    its range is guarded (gradient scale, loud saturation check)."""

    def __init__(self, num_classes: int = 3, dtype: str = DEFAULT_DTYPE):
        super().__init__()
        if num_classes != 3:
            raise ValueError("the B200 hot path implements the reference configuration num_classes=3")
        if dtype not in _DTYPES:
            raise ValueError(f"dtype must be one of {sorted(_DTYPES)}")
        self.num_classes = num_classes
        self.act_dtype = _DTYPES[dtype]
        self.model = _BasicUNetParams(num_classes)
        self.enhance = nn.Sequential(
            nn.Conv2d(num_classes, 64, 3, padding=1), nn.BatchNorm2d(64), nn.ReLU(inplace=True), nn.Conv2d(64, num_classes, 1))
        self._aux_outputs = None
        self._packs = engine.PackCache()
        self._sat = _SaturationMonitor()
        self.grad_sink = None   # parallel.FlatGradBuffer when gradients are exchanged across ranks (parallel.GradientAllReduce)
        self._param_names = [n for n, _ in self.named_parameters()]

    # -- helpers ---------------------------------------------------------------------------------
    def _tensor_dict(self) -> Dict[str, torch.Tensor]:
        d: Dict[str, torch.Tensor] = dict(self.named_parameters())
        d.update(dict(self.named_buffers()))
        return d

    def _amax_buffer(self, device: torch.device) -> Optional[torch.Tensor]:
        return self._sat.buffer(device) if self.act_dtype == torch.float16 else None

    def check_numerics(self) -> None:
        """Raise if any fp16 tensor of the passes launched so far had to be clamped (waits for the device)."""
        self._sat.check(block=True)

    def set_compute_dtype(self, dtype: str) -> "EnhancedUNet":
        self.act_dtype = _DTYPES[dtype]
        self._packs.clear()
        return self

    def _apply(self, fn, *args, **kwargs):
        r = super()._apply(fn, *args, **kwargs)
        self._packs.clear()   # parameters may have moved
        return r

    # -- reference surface -----------------------------------------------------------------------
    def forward(self, x: torch.Tensor) -> torch.Tensor:
        self._aux_outputs = None
        p0 = next(self.parameters())
        if not p0.is_cuda:
            raise RuntimeError("EnhancedUNet (B200) parameters must live on a CUDA device: call .to('cuda'); no CPU fallback")
        if x.device != p0.device:
            raise RuntimeError(f"input on {x.device} but parameters on {p0.device}")
        self._sat.check(block=False)     # an earlier pass clamped fp16 values: loud, one pass late at most, no sync
        need_grad = torch.is_grad_enabled() and (x.requires_grad or any(p.requires_grad for p in self.parameters()))
        if need_grad:
            if not self.training:
                raise NotImplementedError("gradients through eval-mode (running-statistics) BatchNorm are not implemented; "
                                          "call .train() for training or torch.no_grad() for inference")
            params = [p for _, p in self.named_parameters()]
            return _UNetFunction.apply(self, x, *params)
        with torch.no_grad():
            out, _ = engine.forward(self._tensor_dict(), x, self.training, self.act_dtype, self._packs, want_saved=False,
                                    amax=self._amax_buffer(x.device))
            self._sat.publish()
        return out

    def get_aux_outputs(self) -> Optional[Dict[str, torch.Tensor]]:
        """Reference models.py:341-343; always None in the fallback body."""
        return getattr(self, "_aux_outputs", None)


def get_model(model_name: str, num_classes: int = 3, device: str = "cuda", train_mode: bool = False, data_dir: str = None,
              max_size: int = 640, dtype: str = DEFAULT_DTYPE) -> nn.Module:
    """Reference models.py:590-624.  Same signature (the last three reference kwargs are ignored there
    too); does NOT move the model to ``device`` (the caller does, train_eval.py:1079)."""
    print(f"Initializing model: {model_name}")
    if model_name == "enhanced_unet":
        model = EnhancedUNet(num_classes=num_classes, dtype=dtype)
    elif model_name in ("segnet", "unet", "fcn", "pspnet", "linknet"):
        raise NotImplementedError(f"model '{model_name}' is outside the B200 hot path (only 'enhanced_unet' is accelerated)")
    else:
        raise ValueError(f"Unknown model: {model_name}")
    print(f"Model {model_name} initialized")
    return model


class _FusionFunction(torch.autograd.Function):
    """Training-mode fusion blocks as one autograd node (fixed kernel schedules in engine.py)."""

    @staticmethod
    def forward(ctx, module: "FusionHead", out_main, out_aux, s1, s2, *params):
        sd = dict(module.named_parameters())
        sd.update(dict(module.named_buffers()))
        out, saved = engine.fusion_train_forward(sd, out_main, out_aux, module.act_dtype, module._packs, (s1, s2),
                                                 amax=module._amax_buffer(out_main.device))
        module._sat.publish()
        ctx.module, ctx.saved_state = module, saved
        return out

    @staticmethod
    def backward(ctx, dout):
        module, sv = ctx.module, ctx.saved_state
        if sv is None:
            raise RuntimeError("FusionHead backward called twice (saved activations were released)")
        sd = dict(module.named_parameters())
        sd.update(dict(module.named_buffers()))
        dmain, daux, grads = engine.fusion_backward(sd, sv, dout, module.act_dtype, module._packs, amax=module._amax_buffer(dout.device))
        module._sat.publish()
        ctx.saved_state = None
        return (None, dmain, daux, None, None) + tuple(grads[n] for n, _ in module.named_parameters())


class FusionHead(nn.Module):
    """The in-file fusion blocks of the reference's smp body (models.py:276-302, 320-328): attention gate,
    fusion head and residual 1x1 over ``cat([out_main, out_aux])``.  Same sub-module names / indices as the
    reference, so the ``attention_gate.*``, ``fusion_head.*`` and ``fusion_residual.*`` entries of a reference
    checkpoint load directly.  The two smp branches that feed it (UnetPlusPlus / DeepLabV3Plus) are third-party
    code outside this repo's scope; this head runs standalone, forward and backward, in eval and in training mode.

    Training mode: batch-statistics BatchNorm (running statistics updated), Dropout2d (models.py:290, 294) as
    per-(sample, channel) factors keep / (1 - p).  ``forward(out_main, out_aux, dropout_scales=(s1 [B,256], s2 [B,128]))``
    takes the draw from the caller (parity tests hand over the reference's own draw); without it the factors are drawn
    with torch's device generator ([B, C] tensors - bookkeeping, not arithmetic of the path)."""

    def __init__(self, num_classes: int = 3, dtype: str = DEFAULT_DTYPE):
        super().__init__()
        if num_classes != 3:
            raise ValueError("the B200 hot path implements the reference configuration num_classes=3")
        self.num_classes = num_classes
        self.act_dtype = _DTYPES[dtype]
        fc = num_classes * 2
        self.attention_gate = nn.Sequential(
            nn.Conv2d(fc, fc // 2, kernel_size=3, padding=1, bias=False), nn.BatchNorm2d(fc // 2), nn.GELU(),
            nn.Conv2d(fc // 2, fc, kernel_size=1, bias=False), nn.BatchNorm2d(fc), nn.Sigmoid())
        self.fusion_head = nn.Sequential(
            nn.Conv2d(num_classes * 2, 256, kernel_size=3, padding=1, bias=False), nn.BatchNorm2d(256), nn.ReLU(inplace=True),
            nn.Dropout2d(0.2),
            nn.Conv2d(256, 128, kernel_size=3, padding=1, bias=False), nn.BatchNorm2d(128), nn.ReLU(inplace=True),
            nn.Dropout2d(0.15),
            nn.Conv2d(128, 64, kernel_size=3, padding=1, bias=False), nn.BatchNorm2d(64), nn.ReLU(inplace=True),
            nn.Conv2d(64, num_classes, kernel_size=1))
        self.fusion_residual = nn.Conv2d(num_classes * 2, num_classes, kernel_size=1)
        self._packs = engine.PackCache()
        self._sat = _SaturationMonitor()
        self._blob_cache: dict = {}

    def _apply(self, fn, *args, **kwargs):
        r = super()._apply(fn, *args, **kwargs)
        self._packs.clear()
        return r

    def _amax_buffer(self, device: torch.device) -> Optional[torch.Tensor]:
        return self._sat.buffer(device) if self.act_dtype == torch.float16 else None

    def check_numerics(self) -> None:
        self._sat.check(block=True)

    def forward(self, out_main: torch.Tensor, out_aux: torch.Tensor, dropout_scales=None) -> torch.Tensor:
        if not (out_main.is_cuda and out_aux.is_cuda):
            raise RuntimeError("FusionHead (B200) runs on CUDA tensors only; there is no CPU fallback")
        if out_main.shape != out_aux.shape or out_main.dim() != 4 or out_main.shape[1] != 3:
            raise RuntimeError(f"FusionHead expects two [B,3,H,W] tensors, got {tuple(out_main.shape)} / {tuple(out_aux.shape)}")
        self._sat.check(block=False)
        if self.training:
            b, dev = out_main.shape[0], out_main.device
            if dropout_scales is None:
                dropout_scales = tuple(torch.bernoulli(torch.full((b, c), 1.0 - p, device=dev)) / (1.0 - p) for c, p in ((256, 0.2), (128, 0.15)))
            s1, s2 = (s.to(dev, torch.float32).contiguous() for s in dropout_scales)
            if s1.shape != (b, 256) or s2.shape != (b, 128):
                raise RuntimeError(f"dropout_scales must be ([B,256], [B,128]), got {tuple(s1.shape)} / {tuple(s2.shape)}")
            return _FusionFunction.apply(self, out_main, out_aux, s1, s2, *[p for _, p in self.named_parameters()])
        with torch.no_grad():
            sd = dict(self.named_parameters())
            sd.update(dict(self.named_buffers()))
            return engine.fusion_forward(sd, out_main, out_aux, self.act_dtype, self._packs, blob_cache=self._blob_cache)
