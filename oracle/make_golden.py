"""Generate tests/golden/*.npz by running the UNMODIFIED reference (/root/reference) on seeded inputs.

Run in the build container only:  ``python -m oracle.make_golden``.
The fixtures pin the oracle (tests/test_oracle_golden.py) and, through it, the CUDA path.
Parameters and inputs come from ``oracle.make_state_dict`` / ``make_input`` / ``make_target`` (our own
seeded generators), are loaded into the reference module with ``load_state_dict(strict=True)`` and
pushed through the reference's own ``forward`` / ``Trainer._compute_combined_loss`` /
``metrics.calculate_semantic_metrics`` / ``Evaluator._convert_probs_to_mask``.
"""
from __future__ import annotations

import os

import numpy as np
import torch

from . import ref_import
from .unet_oracle import (make_state_dict, make_input, make_target, make_fusion_state_dict)

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def _ref_model(ref_models, sd):
    torch.manual_seed(0)
    m = ref_models.EnhancedUNet(3)
    m.load_state_dict(sd, strict=True)
    return m


def _trainer(ref_te, model):
    return ref_te.Trainer(model, torch.device("cpu"), "enhanced_unet", total_epochs=50)


def model_case(ref_models, ref_te, name, batch, h, w, pseed, xseed, tseed):
    sd = make_state_dict(pseed)
    x = make_input(batch, h, w, xseed)
    t = make_target(batch, h, w, tseed)
    out = {}
    # eval forward
    m = _ref_model(ref_models, sd).eval()
    with torch.no_grad():
        y_eval = m(x)
    # train forward + reference loss + backward
    m = _ref_model(ref_models, sd).train()
    tr = _trainer(ref_te, m)
    y = m(x)
    loss = 0.0
    for i in range(batch):
        oi = torch.nn.functional.interpolate(y[i].unsqueeze(0), size=(h, w), mode="bilinear",
                                             align_corners=False).squeeze(0)   # train_eval.py:306-310
        loss = loss + tr._compute_combined_loss(oi, t[i])
    loss = loss / batch
    loss.backward()
    out["meta"] = np.array([batch, h, w, pseed, xseed, tseed], dtype=np.int64)
    out["logits_eval"] = y_eval.numpy()
    out["logits_train"] = y.detach().numpy()
    out["loss"] = np.array(loss.item(), dtype=np.float64)
    new_sd = m.state_dict()
    for k, v in new_sd.items():
        if "running" in k or "num_batches" in k:
            out["buf/" + k] = v.numpy()
    for k, p in m.named_parameters():
        g = p.grad
        out["gnorm/" + k] = np.array(g.double().norm().item())
        out["gabsmax/" + k] = np.array(g.abs().max().item())
        flat = g.flatten()
        idx = torch.linspace(0, flat.numel() - 1, steps=min(64, flat.numel())).long()
        out["gidx/" + k] = idx.numpy()
        out["gval/" + k] = flat[idx].numpy()
        if g.numel() <= 4096 or k in ("model.enc1.0.weight", "enhance.0.weight", "model.dec1.weight", "enhance.3.weight"):
            out["gfull/" + k] = g.numpy()
    # PyTorch's OWN bf16 path on the unmodified reference (torch.autocast): the yardstick for what bf16
    # arithmetic can deliver on this network (train-mode BN makes it ~20x more sensitive than eval mode)
    def nerr(a, b):
        return float((a.double() - b.double()).abs().max() / b.double().abs().max())
    m2 = _ref_model(ref_models, sd).eval()
    with torch.no_grad(), torch.autocast("cpu", dtype=torch.bfloat16):
        ya = m2(x).float()
    out["autocast/logits_eval_err"] = np.array(nerr(ya, y_eval))
    m2 = _ref_model(ref_models, sd).train()
    tr2 = _trainer(ref_te, m2)
    with torch.autocast("cpu", dtype=torch.bfloat16):
        ya = m2(x)
    out["autocast/logits_train_err"] = np.array(nerr(ya.float(), y.detach()))
    la = 0.0
    for i in range(batch):
        oi = torch.nn.functional.interpolate(ya[i].float().unsqueeze(0), size=(h, w), mode="bilinear",
                                             align_corners=False).squeeze(0)
        la = la + tr2._compute_combined_loss(oi, t[i])
    (la / batch).backward()
    out["autocast/loss"] = np.array(float(la / batch))
    ref_grads = dict(m.named_parameters())
    for k, p in m2.named_parameters():
        a, b = p.grad.double().flatten(), ref_grads[k].grad.double().flatten()
        out["autocast/gcos/" + k] = np.array(float((a * b).sum() / (a.norm() * b.norm() + 1e-30)))
        out["autocast/grel/" + k] = np.array(float((a - b).norm() / (b.norm() + 1e-30)))
    np.savez_compressed(os.path.join(OUT, f"model_{name}.npz"), **out)
    print(f"model_{name}: loss={loss.item():.6f} ysum={y.sum().item():.4f} autocast eval/train err "
          f"{float(out['autocast/logits_eval_err']):.3e}/{float(out['autocast/logits_train_err']):.3e}")


def loss_case(ref_te, ref_models):
    m = _ref_model(ref_models, make_state_dict(0))
    tr = _trainer(ref_te, m)
    g = torch.Generator().manual_seed(77)
    out = {}
    for name, (b, h, w, scale) in {"a": (2, 40, 56, 3.0), "b": (1, 8, 8, 0.5), "c": (3, 16, 24, 10.0)}.items():
        logits = (torch.randn(b, 3, 2 * h, 2 * w, generator=g) * scale).requires_grad_(True)
        t = torch.randint(0, 3, (b, h, w), generator=g)
        if name == "b":
            t[:] = 0          # single-class image: dice/tversky smoothing branch
        loss = 0.0
        per = []
        for i in range(b):
            oi = torch.nn.functional.interpolate(logits[i].unsqueeze(0), size=(h, w), mode="bilinear",
                                                 align_corners=False).squeeze(0)
            li = tr._compute_combined_loss(oi, t[i])
            per.append(li.item())
            loss = loss + li
        loss = loss / b
        loss.backward()
        out[f"{name}/logits"] = logits.detach().numpy()
        out[f"{name}/target"] = t.numpy()
        out[f"{name}/loss"] = np.array(loss.item())
        out[f"{name}/per_sample"] = np.array(per)
        out[f"{name}/grad"] = logits.grad.numpy()
    np.savez_compressed(os.path.join(OUT, "loss.npz"), **out)
    print("loss:", {k: float(v) for k, v in out.items() if k.endswith("/loss")})


def metrics_case(ref_metrics):
    out = {}
    rng = np.random.default_rng(1234)            # KAT-1 of SURVEY.md §8c
    cases = {"kat1": (rng.integers(0, 3, (64, 64)), rng.integers(0, 3, (64, 64)))}
    z = np.zeros((8, 8), dtype=np.int64)
    one_dead = z.copy(); one_dead[3, 4] = 2
    cases["kat2_allbg"] = (z, z.copy())
    cases["kat2_onedead"] = (z.copy(), one_dead)
    rng = np.random.default_rng(99)
    cases["skewed"] = (rng.choice(3, (96, 80), p=[.9, .07, .03]), rng.choice(3, (96, 80), p=[.85, .1, .05]))
    gt_ign = rng.integers(0, 3, (50, 70)); gt_ign[rng.random((50, 70)) < 0.1] = 255
    cases["ignore255"] = (rng.integers(0, 3, (50, 70)), gt_ign)
    cases["ragged_1x7"] = (rng.integers(0, 3, (1, 7)), rng.integers(0, 3, (1, 7)))
    cases["no_live"] = (rng.choice([0, 2], (33, 31)), rng.choice([0, 2], (33, 31)))
    keys = None
    for name, (pred, gt) in cases.items():
        pred = pred.astype(np.int64); gt = gt.astype(np.int64)
        m = ref_metrics.calculate_semantic_metrics(pred, gt)
        keys = list(m.keys())
        out[f"{name}/pred"] = pred
        out[f"{name}/gt"] = gt
        out[f"{name}/values"] = np.array([float(m[k]) for k in keys], dtype=np.float64)
    out["keys"] = np.array(keys)
    np.savez_compressed(os.path.join(OUT, "metrics.npz"), **out)
    print("metrics kat1:", dict(zip(keys, out["kat1/values"])))


def mask_case(ref_te, ref_models):
    ev = ref_te.Evaluator(_ref_model(ref_models, make_state_dict(0)), torch.device("cpu"), "enhanced_unet")
    g = torch.Generator().manual_seed(5)
    out = {}
    specs = {"uniform": (torch.tensor([1., 1., 1.]), 3.0), "live_heavy": (torch.tensor([0., 2.5, 0.]), 2.0),
             "dead_mid": (torch.tensor([1.0, 0., 1.2]), 3.0), "dead_heavy": (torch.tensor([0., 0., 2.0]), 3.0),
             "dead_extreme": (torch.tensor([-1., -1., 3.0]), 2.0), "flat": (torch.tensor([0., 0., 0.]), 0.3)}
    for name, (bias, scale) in specs.items():
        logits = torch.randn(3, 48, 40, generator=g) * scale + bias[:, None, None]
        probs = torch.softmax(logits, dim=0)
        mask = ev._convert_probs_to_mask(probs.clone())
        out[f"{name}/probs"] = probs.numpy()
        out[f"{name}/mask"] = np.asarray(mask, dtype=np.int64)
        print("mask", name, np.bincount(mask.ravel(), minlength=3))
    np.savez_compressed(os.path.join(OUT, "mask.npz"), **out)


def fusion_case(ref_models):
    """Re-instantiate the three nn.Sequential fusion blocks exactly as models.py:277-302 builds
    them (they are plain torch.nn; the smp branches that feed them are not installable)."""
    import torch.nn as nn
    nc = 3
    fc = nc * 2
    gate = nn.Sequential(nn.Conv2d(fc, fc // 2, 3, padding=1, bias=False), nn.BatchNorm2d(fc // 2), nn.GELU(),
                         nn.Conv2d(fc // 2, fc, 1, bias=False), nn.BatchNorm2d(fc), nn.Sigmoid())
    head = nn.Sequential(nn.Conv2d(nc * 2, 256, 3, padding=1, bias=False), nn.BatchNorm2d(256), nn.ReLU(inplace=True),
                         nn.Dropout2d(0.2), nn.Conv2d(256, 128, 3, padding=1, bias=False), nn.BatchNorm2d(128),
                         nn.ReLU(inplace=True), nn.Dropout2d(0.15), nn.Conv2d(128, 64, 3, padding=1, bias=False),
                         nn.BatchNorm2d(64), nn.ReLU(inplace=True), nn.Conv2d(64, nc, 1))
    res = nn.Conv2d(nc * 2, nc, 1)
    sd = make_fusion_state_dict(0)
    gate.load_state_dict({k[len("attention_gate."):]: v for k, v in sd.items() if k.startswith("attention_gate.")})
    head.load_state_dict({k[len("fusion_head."):]: v for k, v in sd.items() if k.startswith("fusion_head.")})
    res.load_state_dict({k[len("fusion_residual."):]: v for k, v in sd.items() if k.startswith("fusion_residual.")})
    gate.eval(); head.eval(); res.eval()
    g = torch.Generator().manual_seed(11)
    a = torch.randn(2, 3, 32, 40, generator=g)
    b = torch.randn(2, 3, 32, 40, generator=g)
    with torch.no_grad():
        f = torch.cat([a, b], 1)
        f = f * gate(f)
        y = head(f) + res(f)
    np.savez_compressed(os.path.join(OUT, "fusion.npz"), main=a.numpy(), aux=b.numpy(), out=y.numpy())
    print("fusion ysum", y.sum().item())


def fusion_train_case(ref_models):
    """The same three blocks in TRAINING mode (batch-statistics BatchNorm, Dropout2d drawing from torch's global RNG) and
    their autograd backward.  The per-(sample, channel) Dropout2d factors the reference drew are read back through forward
    hooks, so that the oracle / the CUDA path can be handed the identical draw."""
    import torch.nn as nn
    nc, fc = 3, 6
    gate = nn.Sequential(nn.Conv2d(fc, fc // 2, 3, padding=1, bias=False), nn.BatchNorm2d(fc // 2), nn.GELU(),
                         nn.Conv2d(fc // 2, fc, 1, bias=False), nn.BatchNorm2d(fc), nn.Sigmoid())
    head = nn.Sequential(nn.Conv2d(nc * 2, 256, 3, padding=1, bias=False), nn.BatchNorm2d(256), nn.ReLU(inplace=True),
                         nn.Dropout2d(0.2), nn.Conv2d(256, 128, 3, padding=1, bias=False), nn.BatchNorm2d(128),
                         nn.ReLU(inplace=True), nn.Dropout2d(0.15), nn.Conv2d(128, 64, 3, padding=1, bias=False),
                         nn.BatchNorm2d(64), nn.ReLU(inplace=True), nn.Conv2d(64, nc, 1))
    res = nn.Conv2d(nc * 2, nc, 1)
    sd = make_fusion_state_dict(0)
    gate.load_state_dict({k[len("attention_gate."):]: v for k, v in sd.items() if k.startswith("attention_gate.")})
    head.load_state_dict({k[len("fusion_head."):]: v for k, v in sd.items() if k.startswith("fusion_head.")})
    res.load_state_dict({k[len("fusion_residual."):]: v for k, v in sd.items() if k.startswith("fusion_residual.")})
    gate.train(); head.train(); res.train()
    scales = {}

    def hook(name):
        def fn(mod, inp, out):
            x = inp[0].detach()
            idx = x.abs().flatten(2).argmax(2, keepdim=True)             # a non-zero element of every (b, c) plane
            xi, oi = x.flatten(2).gather(2, idx), out.detach().flatten(2).gather(2, idx)
            scales[name] = torch.where(xi != 0, oi / xi, torch.zeros_like(xi)).squeeze(2)
        return fn

    head[3].register_forward_hook(hook("s1"))
    head[7].register_forward_hook(hook("s2"))
    g = torch.Generator().manual_seed(21)
    a = torch.randn(2, 3, 32, 40, generator=g).requires_grad_(True)
    b = torch.randn(2, 3, 32, 40, generator=g).requires_grad_(True)
    dout = torch.randn(2, 3, 32, 40, generator=g)
    torch.manual_seed(5)
    f = torch.cat([a, b], 1)
    f = f * gate(f)
    y = head(f) + res(f)
    (y * dout).sum().backward()
    out = dict(main=a.detach().numpy(), aux=b.detach().numpy(), dout=dout.numpy(), out=y.detach().numpy(),
               dmain=a.grad.numpy(), daux=b.grad.numpy(), s1=scales["s1"].numpy(), s2=scales["s2"].numpy())
    for prefix, mod in (("attention_gate", gate), ("fusion_head", head), ("fusion_residual", res)):
        for n, p in mod.named_parameters():
            out[f"grad/{prefix}.{n}"] = p.grad.numpy()
        for n, bf in mod.named_buffers():
            out[f"buf/{prefix}.{n}"] = bf.numpy()
    for key, keep in (("s1", 1 / 0.8), ("s2", 1 / 0.85)):
        assert np.all(np.isclose(out[key], 0.0) | np.isclose(out[key], keep, rtol=1e-5)), key
    np.savez_compressed(os.path.join(OUT, "fusion_train.npz"), **out)
    print("fusion_train ysum", y.sum().item(), "kept", float((out["s1"] > 0).mean()), float((out["s2"] > 0).mean()))


def instances_case(ref_metrics):
    """metrics.calculate_instance_metrics (metrics.py:61-194) on seeded synthetic instance sets."""
    from .metrics_oracle import INSTANCE_CASES, make_instance_case
    out = {}
    for name, spec in INSTANCE_CASES.items():
        pm, pl, ps, gm, gl = make_instance_case(*spec)
        m = ref_metrics.calculate_instance_metrics(pm, pl, ps, gm, gl)
        keys = sorted(m.keys())
        out[f"{name}/keys"] = np.array(keys)
        out[f"{name}/values"] = np.array([float(m[k]) for k in keys], dtype=np.float64)
        print("instances", name, {k: round(float(m[k]), 4) for k in keys})
    np.savez_compressed(os.path.join(OUT, "instances.npz"), **out)


def tta_case(ref_te, ref_models):
    """Evaluator._run_tta_inference (train_eval.py:419-453): 5 views incl. the 0.75x / 1.25x bilinear rescales."""
    model = _ref_model(ref_models, make_state_dict(0)).eval()
    ev = ref_te.Evaluator(model, torch.device("cpu"), "enhanced_unet")
    assert ev.enable_tta
    out = {}
    for name, (h, w, seed) in {"48x40": (48, 40, 21), "64x64": (64, 64, 22)}.items():
        img = make_input(1, h, w, seed)[0]
        with torch.no_grad():
            base = ev._run_model_single(img)
            tta = ev._run_tta_inference(img)
        out[f"{name}/meta"] = np.array([h, w, seed])
        out[f"{name}/base"] = base.numpy()
        out[f"{name}/tta"] = tta.numpy()
        print("tta", name, float(tta.sum()), float((tta - base).abs().max()))
    np.savez_compressed(os.path.join(OUT, "tta.npz"), **out)


def preprocess_case(ref_te, ref_models):
    """Evaluator._prepare_image_tensor (train_eval.py:365-395): the cv2 CLAHE + sharpening step in front of the model."""
    model = _ref_model(ref_models, make_state_dict(0)).eval()
    ev = ref_te.Evaluator(model, torch.device("cpu"), "enhanced_unet")
    out = {}
    for name, (h, w, seed) in {"96x80": (96, 80, 31), "64x64": (64, 64, 32)}.items():
        g = torch.Generator().manual_seed(seed)
        yy, xx = torch.meshgrid(torch.arange(h, dtype=torch.float32), torch.arange(w, dtype=torch.float32), indexing="ij")
        img = 0.7 + 0.05 * torch.randn(1, h, w, generator=g) - 0.3 * torch.exp(-((yy - h / 2) ** 2 + (xx - w / 3) ** 2) / 60.0)
        img = img.clamp(0, 1).expand(3, h, w).contiguous()
        out[f"{name}/image"] = img.numpy()
        out[f"{name}/prepared"] = ev._prepare_image_tensor(img).numpy()
        print("preprocess", name, float(out[f"{name}/prepared"].sum()))
    np.savez_compressed(os.path.join(OUT, "preprocess.npz"), **out)


def main():
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    ref_models, ref_metrics, ref_te = ref_import.load()
    import sys
    only = set(sys.argv[1:])          # e.g. ``python -m oracle.make_golden instances tta`` regenerates just those fixtures

    def want(name):
        return not only or name in only

    if want("model"):
        model_case(ref_models, ref_te, "b2_32x32", 2, 32, 32, 0, 1, 2)
        model_case(ref_models, ref_te, "b1_16x24", 1, 16, 24, 3, 4, 5)
        model_case(ref_models, ref_te, "b2_64x64", 2, 64, 64, 6, 7, 8)
    if want("loss"):
        loss_case(ref_te, ref_models)
    if want("metrics"):
        metrics_case(ref_metrics)
    if want("mask"):
        mask_case(ref_te, ref_models)
    if want("fusion"):
        fusion_case(ref_models)
    if want("fusion_train"):
        fusion_train_case(ref_models)
    if want("instances"):
        instances_case(ref_metrics)
    if want("tta"):
        tta_case(ref_te, ref_models)
    if want("preprocess"):
        preprocess_case(ref_te, ref_models)


if __name__ == "__main__":
    main()
