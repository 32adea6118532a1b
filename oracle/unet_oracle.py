"""CPU restatement (torch fp32, functional) of the Enhanced-UNet hot path.  TEST INFRASTRUCTURE ONLY.

What is restated, with the reference lines each piece follows (paths are into /root/reference):

* ``unet_forward``   - ``EnhancedUNet.forward`` fallback body (models.py:334-339) =
                       ``BasicUNet.forward`` (models.py:227-238) + ``out + enhance(out)``
                       (models.py:308-313, 337).  ``_conv_block`` is models.py:217-225.
* ``combined_loss``  - ``Trainer._compute_combined_loss`` (train_eval.py:183-197) =
                       2.5*FocalLoss (train_eval.py:37-60, gamma=5, alpha=[1,8,5], CE weights
                       [1,20,10], train_eval.py:74-79) + 2.5*dice_loss (train_eval.py:134-157)
                       + 1.0*tversky_loss (train_eval.py:159-181), after the 2x-down bilinear
                       resize of the logits (train_eval.py:306-310) which is an exact 2x2 mean.
* ``batch_loss``     - the per-sample loop + ``/ batch_size`` of ``Trainer.train_epoch``
                       (train_eval.py:261-337).
* ``fusion_forward`` - the in-file fusion blocks of the smp body (models.py:276-302, 320-328),
                       eval mode (Dropout2d is the identity), testable standalone.

The restatement is functional (explicit parameter dict, no nn.Module) so it shares no code with the
product's drop-in ``nn.Module``.  It is pinned against the imported reference by
``oracle/make_golden.py`` -> ``tests/golden/*.npz`` and ``tests/test_oracle_golden.py``.
"""
from __future__ import annotations

import math
from typing import Dict, List, Tuple

import torch
import torch.nn.functional as F

BLOCKS: List[Tuple[str, int, int]] = [
    # (state_dict prefix, Cin, Cout) in construction order, models.py:203-211
    ("model.enc1", 3, 64),
    ("model.enc2", 64, 128),
    ("model.enc3", 128, 256),
    ("model.enc4", 256, 512),
    ("model.dec4", 512 + 256, 256),
    ("model.dec3", 256 + 128, 128),
    ("model.dec2", 128 + 64, 64),
]


def _specs(num_classes: int = 3):
    params, buffers = [], []
    for prefix, cin, cout in BLOCKS:
        for conv_i, bn_i, ci in ((0, 1, cin), (3, 4, cout)):
            params.append((f"{prefix}.{conv_i}.weight", (cout, ci, 3, 3)))
            params.append((f"{prefix}.{conv_i}.bias", (cout,)))
            params.append((f"{prefix}.{bn_i}.weight", (cout,)))
            params.append((f"{prefix}.{bn_i}.bias", (cout,)))
            buffers.append((f"{prefix}.{bn_i}.running_mean", (cout,)))
            buffers.append((f"{prefix}.{bn_i}.running_var", (cout,)))
            buffers.append((f"{prefix}.{bn_i}.num_batches_tracked", ()))
    params.append(("model.dec1.weight", (num_classes, 64, 1, 1)))
    params.append(("model.dec1.bias", (num_classes,)))
    params.append(("enhance.0.weight", (64, num_classes, 3, 3)))
    params.append(("enhance.0.bias", (64,)))
    params.append(("enhance.1.weight", (64,)))
    params.append(("enhance.1.bias", (64,)))
    buffers.append(("enhance.1.running_mean", (64,)))
    buffers.append(("enhance.1.running_var", (64,)))
    buffers.append(("enhance.1.num_batches_tracked", ()))
    params.append(("enhance.3.weight", (num_classes, 64, 1, 1)))
    params.append(("enhance.3.bias", (num_classes,)))
    return params, buffers


PARAM_SPECS, BUFFER_SPECS = _specs()


def make_state_dict(seed: int = 0, randomize_bn: bool = True) -> Dict[str, torch.Tensor]:
    """Deterministic fp32 state_dict with the reference's 109 keys/shapes.

    Conv weights/biases follow the nn.Conv2d default scale (kaiming-uniform with a=sqrt(5), i.e.
    U(-1/sqrt(fan_in), 1/sqrt(fan_in))), drawn from our own seeded CPU generator so fixtures do not
    depend on the construction order of the reference module.  With ``randomize_bn`` the BN affine
    parameters and running statistics are made non-trivial so eval-mode parity exercises them.
    """
    g = torch.Generator().manual_seed(1000 + seed)
    sd: Dict[str, torch.Tensor] = {}
    for name, shape in PARAM_SPECS:
        if len(shape) == 4:
            fan_in = shape[1] * shape[2] * shape[3]
            bound = 1.0 / math.sqrt(fan_in)
            sd[name] = (torch.rand(shape, generator=g) * 2 - 1) * bound
        elif name.split(".")[-2] in ("0", "3", "dec1") and name.endswith("bias"):
            # conv bias: fan_in of the matching weight
            w = sd[name[: -len("bias")] + "weight"]
            bound = 1.0 / math.sqrt(w.shape[1] * w.shape[2] * w.shape[3])
            sd[name] = (torch.rand(shape, generator=g) * 2 - 1) * bound
        elif name.endswith("weight"):  # BN gamma
            sd[name] = 0.5 + torch.rand(shape, generator=g) if randomize_bn else torch.ones(shape)
        else:  # BN beta
            sd[name] = (torch.rand(shape, generator=g) - 0.5) * 0.4 if randomize_bn else torch.zeros(shape)
    for name, shape in BUFFER_SPECS:
        if name.endswith("running_mean"):
            sd[name] = (torch.rand(shape, generator=g) - 0.5) * 0.2 if randomize_bn else torch.zeros(shape)
        elif name.endswith("running_var"):
            sd[name] = 0.5 + torch.rand(shape, generator=g) if randomize_bn else torch.ones(shape)
        else:
            sd[name] = torch.tensor(0, dtype=torch.int64)
    # order the dict like the reference's state_dict (params and buffers interleaved per module)
    ordered: Dict[str, torch.Tensor] = {}
    def _mod(k):
        return k.rsplit(".", 1)[0]
    mods: List[str] = []
    for name, _ in PARAM_SPECS:
        if _mod(name) not in mods:
            mods.append(_mod(name))
    for m in mods:
        for suffix in ("weight", "bias", "running_mean", "running_var", "num_batches_tracked"):
            k = f"{m}.{suffix}"
            if k in sd:
                ordered[k] = sd[k]
    assert len(ordered) == len(sd)
    return ordered


def make_input(batch: int, h: int, w: int, seed: int = 1) -> torch.Tensor:
    """Gray plane in [0,1] replicated to 3 channels (what ``Image.convert('RGB')`` gives for a gray
    JPEG, dataset.py:139).  Shape [B,3,H,W] fp32 contiguous."""
    g = torch.Generator().manual_seed(seed)
    return torch.rand(batch, 1, h, w, generator=g).expand(batch, 3, h, w).contiguous()


def make_target(batch: int, h: int, w: int, seed: int = 2) -> torch.Tensor:
    g = torch.Generator().manual_seed(seed)
    return torch.randint(0, 3, (batch, h, w), generator=g, dtype=torch.int64)


def _bn(x, sd, prefix, train, new_buffers, momentum=0.1, eps=1e-5):
    """nn.BatchNorm2d (models.py:220,223,310): train -> batch statistics (biased var for the
    normalisation, unbiased var into running_var, momentum 0.1, num_batches_tracked += 1)."""
    rm = sd[f"{prefix}.running_mean"].clone()
    rv = sd[f"{prefix}.running_var"].clone()
    y = F.batch_norm(x, rm, rv, sd[f"{prefix}.weight"], sd[f"{prefix}.bias"], train, momentum, eps)
    if train:
        new_buffers[f"{prefix}.running_mean"] = rm
        new_buffers[f"{prefix}.running_var"] = rv
        new_buffers[f"{prefix}.num_batches_tracked"] = sd[f"{prefix}.num_batches_tracked"] + 1
    return y


def _conv_block(x, sd, prefix, train, nb):
    """``BasicUNet._conv_block`` (models.py:217-225): conv3x3+bias -> BN -> ReLU, twice."""
    x = F.conv2d(x, sd[f"{prefix}.0.weight"], sd[f"{prefix}.0.bias"], padding=1)
    x = F.relu(_bn(x, sd, f"{prefix}.1", train, nb))
    x = F.conv2d(x, sd[f"{prefix}.3.weight"], sd[f"{prefix}.3.bias"], padding=1)
    x = F.relu(_bn(x, sd, f"{prefix}.4", train, nb))
    return x


def _up(x):
    """nn.Upsample(scale_factor=2, mode='bilinear', align_corners=False) (models.py:215)."""
    return F.interpolate(x, scale_factor=2, mode="bilinear", align_corners=False)


def unet_forward(sd: Dict[str, torch.Tensor], x: torch.Tensor, train: bool):
    """Returns (logits [B,3,2H,2W], new_buffers dict).  models.py:227-238 then 337."""
    nb: Dict[str, torch.Tensor] = {}
    e1 = _conv_block(x, sd, "model.enc1", train, nb)
    e2 = _conv_block(F.max_pool2d(e1, 2), sd, "model.enc2", train, nb)
    e3 = _conv_block(F.max_pool2d(e2, 2), sd, "model.enc3", train, nb)
    e4 = _conv_block(F.max_pool2d(e3, 2), sd, "model.enc4", train, nb)
    d4 = _conv_block(torch.cat([_up(e4), e3], dim=1), sd, "model.dec4", train, nb)
    d3 = _conv_block(torch.cat([_up(d4), e2], dim=1), sd, "model.dec3", train, nb)
    d2 = _conv_block(torch.cat([_up(d3), e1], dim=1), sd, "model.dec2", train, nb)
    d1 = F.conv2d(_up(d2), sd["model.dec1.weight"], sd["model.dec1.bias"])
    # enhance head (models.py:308-313) with residual add (models.py:337)
    t = F.conv2d(d1, sd["enhance.0.weight"], sd["enhance.0.bias"], padding=1)
    t = F.relu(_bn(t, sd, "enhance.1", train, nb))
    t = F.conv2d(t, sd["enhance.3.weight"], sd["enhance.3.bias"])
    return d1 + t, nb


# ---------------------------------------------------------------------------------------------
# loss (train_eval.py)
# ---------------------------------------------------------------------------------------------
CE_W = (1.0, 20.0, 10.0)      # train_eval.py:76
FOCAL_ALPHA = (1.0, 8.0, 5.0)  # train_eval.py:77
FOCAL_GAMMA = 5.0              # train_eval.py:81
DICE_W = (1.0, 15.0, 8.0)      # train_eval.py:140
TV_W = (1.0, 12.0, 6.0)        # train_eval.py:164
TV_ALPHA = 0.7                 # train_eval.py:159
W_FOCAL, W_DICE, W_TV = 2.5, 2.5, 1.0  # train_eval.py:83-86 (enhanced_unet)


def combined_loss(logits_hw: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
    """One sample.  ``logits_hw`` [3,H,W] (already resized to the mask size), ``target`` [H,W] int64."""
    lp = F.log_softmax(logits_hw, dim=0)
    p = lp.exp()
    oh = F.one_hot(target, 3).permute(2, 0, 1).to(lp.dtype)
    w = torch.tensor(CE_W, dtype=lp.dtype)
    al = torch.tensor(FOCAL_ALPHA, dtype=lp.dtype)
    ce = -(lp * oh).sum(0) * w[target]                      # weighted CE, reduction='none'
    pt = torch.exp(-ce)                                     # NB: includes the class weight
    focal = (al[target] * (1 - pt) ** FOCAL_GAMMA * ce).mean()
    inter = (p * oh).sum((1, 2))
    sp = p.sum((1, 2))
    st = oh.sum((1, 2))
    dice = (2 * inter + 1e-6) / (sp + st + 1e-6)
    dice_l = (torch.tensor(DICE_W, dtype=lp.dtype) * (1 - dice)).mean()
    fp = sp - inter
    fn = st - inter
    tv = (inter + 1e-6) / (inter + TV_ALPHA * fp + (1 - TV_ALPHA) * fn + 1e-6)
    tv_l = (torch.tensor(TV_W, dtype=lp.dtype) * (1 - tv)).mean()
    return W_FOCAL * focal + W_DICE * dice_l + W_TV * tv_l


def batch_loss(logits: torch.Tensor, targets: torch.Tensor) -> torch.Tensor:
    """``logits`` [B,3,2H,2W] straight from the model, ``targets`` [B,H,W] int64.
    train_eval.py:261-337: per-sample bilinear resize to the mask size (== 2x2 mean), combined
    loss, sum over samples, divide by batch size."""
    b = logits.shape[0]
    h, w = targets.shape[1:]
    total = logits.new_zeros(())
    for i in range(b):
        li = logits[i]
        if li.shape[1:] != targets[i].shape:
            li = F.interpolate(li.unsqueeze(0), size=(h, w), mode="bilinear", align_corners=False).squeeze(0)
        total = total + combined_loss(li, targets[i])
    return total / b


# ---------------------------------------------------------------------------------------------
# fusion blocks of the smp body (secondary path, models.py:276-302, 320-328), eval mode
# ---------------------------------------------------------------------------------------------
FUSION_PARAM_SPECS = [
    ("attention_gate.0.weight", (3, 6, 3, 3)), ("attention_gate.1.weight", (3,)), ("attention_gate.1.bias", (3,)),
    ("attention_gate.3.weight", (6, 3, 1, 1)), ("attention_gate.4.weight", (6,)), ("attention_gate.4.bias", (6,)),
    ("fusion_head.0.weight", (256, 6, 3, 3)), ("fusion_head.1.weight", (256,)), ("fusion_head.1.bias", (256,)),
    ("fusion_head.4.weight", (128, 256, 3, 3)), ("fusion_head.5.weight", (128,)), ("fusion_head.5.bias", (128,)),
    ("fusion_head.8.weight", (64, 128, 3, 3)), ("fusion_head.9.weight", (64,)), ("fusion_head.9.bias", (64,)),
    ("fusion_head.11.weight", (3, 64, 1, 1)), ("fusion_head.11.bias", (3,)),
    ("fusion_residual.weight", (3, 6, 1, 1)), ("fusion_residual.bias", (3,)),
]
FUSION_BN = ["attention_gate.1", "attention_gate.4", "fusion_head.1", "fusion_head.5", "fusion_head.9"]


def make_fusion_state_dict(seed: int = 0) -> Dict[str, torch.Tensor]:
    g = torch.Generator().manual_seed(2000 + seed)
    sd: Dict[str, torch.Tensor] = {}
    for name, shape in FUSION_PARAM_SPECS:
        if len(shape) == 4:
            bound = 1.0 / math.sqrt(shape[1] * shape[2] * shape[3])
            sd[name] = (torch.rand(shape, generator=g) * 2 - 1) * bound
        elif name.rsplit(".", 1)[0] in FUSION_BN:
            sd[name] = 0.5 + torch.rand(shape, generator=g) if name.endswith("weight") else (torch.rand(shape, generator=g) - 0.5) * 0.4
        else:
            sd[name] = (torch.rand(shape, generator=g) * 2 - 1) * 0.1
    for bn in FUSION_BN:
        c = sd[f"{bn}.weight"].shape
        sd[f"{bn}.running_mean"] = (torch.rand(c, generator=g) - 0.5) * 0.2
        sd[f"{bn}.running_var"] = 0.5 + torch.rand(c, generator=g)
        sd[f"{bn}.num_batches_tracked"] = torch.tensor(0, dtype=torch.int64)
    return sd


def fusion_forward(sd: Dict[str, torch.Tensor], out_main: torch.Tensor, out_aux: torch.Tensor, train: bool = False,
                   dropout_scales=None):
    """models.py:320-328: cat -> attention gate -> multiply -> fusion head + residual.

    Eval mode (default): BatchNorm on running statistics, Dropout2d is the identity; returns the output tensor.
    ``train=True``: batch statistics (running statistics / num_batches_tracked updated as nn.BatchNorm2d does) and
    Dropout2d (models.py:290, 294) applied as a per-(sample, channel) factor ``dropout_scales = (s1 [B,256], s2 [B,128])``
    with entries ``keep / (1 - p)`` - the draw is the caller's (the reference takes it from torch's global RNG); returns
    ``(output, new_buffers)``."""
    nb: Dict[str, torch.Tensor] = {}

    def bn(x, p):
        if not train:
            return F.batch_norm(x, sd[f"{p}.running_mean"], sd[f"{p}.running_var"], sd[f"{p}.weight"], sd[f"{p}.bias"], False, 0.1, 1e-5)
        return _bn(x, sd, p, True, nb)

    f = torch.cat([out_main, out_aux], dim=1)
    a = F.conv2d(f, sd["attention_gate.0.weight"], None, padding=1)
    a = F.gelu(bn(a, "attention_gate.1"))
    a = F.conv2d(a, sd["attention_gate.3.weight"], None)
    a = torch.sigmoid(bn(a, "attention_gate.4"))
    f = f * a
    h = F.relu(bn(F.conv2d(f, sd["fusion_head.0.weight"], None, padding=1), "fusion_head.1"))
    if train:
        h = h * dropout_scales[0][:, :, None, None]
    h = F.relu(bn(F.conv2d(h, sd["fusion_head.4.weight"], None, padding=1), "fusion_head.5"))
    if train:
        h = h * dropout_scales[1][:, :, None, None]
    h = F.relu(bn(F.conv2d(h, sd["fusion_head.8.weight"], None, padding=1), "fusion_head.9"))
    h = F.conv2d(h, sd["fusion_head.11.weight"], sd["fusion_head.11.bias"])
    y = h + F.conv2d(f, sd["fusion_residual.weight"], sd["fusion_residual.bias"])
    return (y, nb) if train else y
