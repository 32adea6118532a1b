"""Thin host wrappers (torch tensors in / out) over single C-ABI kernels that callers use directly:
the fused training loss and the integer confusion counts."""
from __future__ import annotations

from typing import Optional, Tuple

import torch

from .lib import call, ptr


def confusion_counts(pred: torch.Tensor, gt: torch.Tensor) -> torch.Tensor:
    """Per-image 4x4 int64 counts ``CM[i, g, p]`` (class 3 = any value outside {0,1,2}).

    ``pred`` / ``gt``: CUDA integer tensors [N, ...] of identical shape and dtype uint8 / int32 / int64
    (the reference passes int64 masks, metrics.py:29; uint8 is the in-pipeline form).  Bit-exact."""
    if not (pred.is_cuda and gt.is_cuda):
        raise RuntimeError("confusion_counts needs CUDA tensors (no CPU fallback)")
    if pred.shape != gt.shape:
        raise RuntimeError(f"pred {tuple(pred.shape)} and gt {tuple(gt.shape)} differ in shape")
    if pred.dtype != gt.dtype:
        raise RuntimeError(f"pred ({pred.dtype}) and gt ({gt.dtype}) must share a dtype")
    sizes = {torch.uint8: 1, torch.int32: 4, torch.int64: 8}
    if pred.dtype not in sizes:
        raise RuntimeError(f"unsupported mask dtype {pred.dtype}")
    n = pred.shape[0] if pred.dim() > 0 else 1
    pred = pred.contiguous()
    gt = gt.contiguous()
    px = pred.numel() // n if n else 0
    out = torch.zeros(n, 4, 4, device=pred.device, dtype=torch.int64)
    done = 0
    while done < n:   # the kernel takes at most 65535 images per launch (gridDim.y)
        m = min(65535, n - done)
        call("eunet_confusion4x4", pred[done:].data_ptr(), gt[done:].data_ptr(), sizes[pred.dtype], m, px, out[done:].data_ptr())
        done += m
    return out


class _CombinedLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits: torch.Tensor, target: torch.Tensor):
        B, C, Hl, Wl = logits.shape
        H, W = target.shape[1:]
        if C != 3:
            raise RuntimeError("combined_loss expects 3-class logits")
        if (Hl, Wl) == (2 * H, 2 * W):
            scale = 2
        elif (Hl, Wl) == (H, W):
            scale = 1
        else:
            raise RuntimeError(f"logits {Hl}x{Wl} must be 1x or 2x the mask size {H}x{W}")
        logits = logits.contiguous().float()
        target = target.contiguous().long()
        dev = logits.device
        partial = torch.empty(B, 10, device=dev, dtype=torch.float64)
        coef = torch.empty(B, 8, device=dev, dtype=torch.float64)
        loss = torch.empty((), device=dev, dtype=torch.float32)
        per = torch.empty(B, device=dev, dtype=torch.float32)
        call("eunet_loss_fwd", ptr(logits), ptr(target), B, H, W, scale, ptr(partial), ptr(loss), ptr(per), ptr(coef))
        ctx.save_for_backward(logits, target, coef)
        ctx.geom = (B, H, W, scale)
        ctx.per_sample = per
        return loss

    @staticmethod
    def backward(ctx, grad_out: torch.Tensor):
        logits, target, coef = ctx.saved_tensors
        B, H, W, scale = ctx.geom
        dlogits = torch.empty_like(logits)
        go = grad_out.contiguous().float()
        call("eunet_loss_bwd", ptr(logits), ptr(target), B, H, W, scale, ptr(coef), ptr(go), ptr(dlogits))
        return dlogits, None


def combined_loss(logits: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
    """Batch loss of the reference Trainer for 'enhanced_unet' (train_eval.py:183-197, 261-337):
    mean over samples of 2.5*focal + 2.5*dice + 1.0*tversky, logits [B,3,2H,2W] (or [B,3,H,W]) fp32,
    target [B,H,W] int64 in {0,1,2}.  Differentiable w.r.t. ``logits``."""
    if not logits.is_cuda:
        raise RuntimeError("combined_loss needs CUDA tensors (no CPU fallback)")
    return _CombinedLoss.apply(logits, target)


def pair_intersections(a: torch.Tensor, b: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """All pairwise intersection counts of two sets of membership masks (reference metrics.py:95-101 evaluates
    ``calculate_iou`` per pair on the host).  ``a`` [P, ...], ``b`` [G, ...]: CUDA uint8/bool masks of the same
    per-mask shape (non-zero = member).  Returns int64 (inter [P,G], area_a [P], area_b [G]); bit-exact."""
    if not (a.is_cuda and b.is_cuda):
        raise RuntimeError("pair_intersections needs CUDA tensors (no CPU fallback)")
    if a.shape[1:] != b.shape[1:]:
        raise RuntimeError(f"mask shapes differ: {tuple(a.shape[1:])} vs {tuple(b.shape[1:])}")
    P, G = a.shape[0], b.shape[0]
    hw = a[0].numel() if P else (b[0].numel() if G else 0)
    words = (hw + 31) // 32
    dev = a.device
    out = torch.zeros(P, G, device=dev, dtype=torch.int64)
    areas = []
    bits = []
    for m, n in ((a, P), (b, G)):
        m = (m != 0).to(torch.uint8).contiguous() if m.dtype != torch.uint8 else m.contiguous()
        bt = torch.empty(n, words, device=dev, dtype=torch.int32)
        ar = torch.zeros(n, device=dev, dtype=torch.int64)
        done = 0
        while done < n:
            k = min(65535, n - done)
            call("eunet_pack_mask_bits", m[done:].data_ptr(), k, hw, bt[done:].data_ptr(), ar[done:].data_ptr())
            done += k
        bits.append(bt)
        areas.append(ar)
    if P and G:
        call("eunet_pair_intersections", bits[0].data_ptr(), P, bits[1].data_ptr(), G, words, out.data_ptr())
    return out, areas[0], areas[1]
