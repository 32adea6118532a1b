"""Diagnostics on the GPU box: (1) does tcgen05 kind::f16 accept MIXED operand formats (A = fp16, B = bf16)?
(2) parity of the CUDA path against the CPU oracle at BASELINE-scale shapes (logits, masks, gradients).
Writes gpurun_out/diag_parity.json.  Test infrastructure: uses oracle/ as the checker."""
import importlib.util
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402

import oracle  # noqa: E402
import bench  # noqa: E402
from enhanced_unet_b200.models import EnhancedUNet  # noqa: E402
from enhanced_unet_b200.ops import combined_loss  # noqa: E402
from enhanced_unet_b200.optim import ClippedAdamW  # noqa: E402

OUT = {}


def save():
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(OUT, open(os.path.join(ROOT, "gpurun_out", "diag_parity.json"), "w"), indent=1)


def probe_mixed(only=None):
    spec = importlib.util.spec_from_file_location("pu", os.path.join(ROOT, "tests", "probe_util.py"))
    pu = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(pu)
    g = torch.Generator().manual_seed(5)
    a32 = torch.randn(128, 64, generator=g)
    b32 = torch.randn(64, 64, generator=g)
    res = {}
    for an, bn in (("bf16", "bf16"), ("f16", "f16"), ("f16", "bf16"), ("bf16", "f16")):
        if only is not None and only != f"{an},{bn}":
            continue
        dt = {"bf16": torch.bfloat16, "f16": torch.float16}
        a, b = a32.to(dt[an]), b32.to(dt[bn])
        want = a.float().numpy() @ b.float().numpy().T
        af, bf = (0 if an == "f16" else 1), (0 if bn == "f16" else 1)
        idesc = (1 << 4) | (af << 7) | (bf << 10) | ((64 >> 3) << 17) | ((128 >> 4) << 24)
        loads = [(0, (0, 0), 0), (1, (0, 0), 16384)]
        mmas = [(pu.smem_desc(32 * j, 16, 1024, pu.SW128), pu.smem_desc(16384 + 32 * j, 16, 1024, pu.SW128), idesc, int(j > 0), 0)
                for j in range(4)]
        x = torch.zeros(1, 8, 8, 64, device="cuda", dtype=torch.bfloat16)
        try:
            t, _ = pu.run_probe(a.view(torch.bfloat16).cuda(), (128, 64), 128, b.view(torch.bfloat16).cuda(), (64, 64), 128, x,
                                (64, 8, 8, 1), 128, loads, 16384 + 8192, mmas, 64, 32768, 32768)
            res[f"A={an},B={bn}"] = float(np.abs(t[:, :64] - want).max() / np.abs(want).max())
        except Exception as e:  # noqa: BLE001
            res[f"A={an},B={bn}"] = f"ERROR {e}"
    print("mixed-format MMA:", res, flush=True)
    return res


def grad_table(model, gref):
    rows = {}
    A, Bv = [], []
    for k, p in model.named_parameters():
        if (k.endswith(".0.bias") or k.endswith(".3.bias")) and "enhance.3" not in k:
            continue
        a, b = p.grad.detach().cpu().flatten().double(), gref[k].flatten().double()
        A.append(a)
        Bv.append(b)
        rows[k] = (float((a - b).norm() / b.norm()), float((a @ b) / (a.norm() * b.norm() + 1e-300)))
    a, b = torch.cat(A), torch.cat(Bv)
    worst = sorted(rows.items(), key=lambda kv: -kv[1][0])[:6]
    return {"global_rel": float((a - b).norm() / b.norm()), "global_cos": float((a @ b) / (a.norm() * b.norm())),
            "worst_rel": worst, "min_cos": min(v[1] for v in rows.values())}


def mask_stats(y, yref):
    pr, p = F.avg_pool2d(yref, 2), F.avg_pool2d(y, 2)
    agree = (p.argmax(1) == pr.argmax(1))
    top2 = pr.topk(2, dim=1).values
    margin = top2[:, 0] - top2[:, 1]
    scale = float(yref.abs().max())
    out = {"agree": float(agree.float().mean())}
    for tol in (2e-2, 1e-2, 5e-3):
        dec = margin > 2 * tol * scale
        out[f"decidable_frac@{tol}"] = float(dec.float().mean())
        out[f"agree_on_decidable@{tol}"] = float(agree[dec].float().mean()) if dec.any() else None
    return out


def parity_case(name, sd, x, t, train, dtypes=("bf16",)):
    t0 = time.time()
    if train:
        params = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and "running" not in k else v) for k, v in sd.items()}
        yref, _ = oracle.unet_forward(params, x, train=True)
        lref = oracle.batch_loss(yref, t)
        lref.backward()
        gref = {k: v.grad for k, v in params.items() if getattr(v, "grad", None) is not None}
        yref = yref.detach()
    else:
        with torch.no_grad():
            yref, _ = oracle.unet_forward(sd, x, train=False)
    res = {"oracle_s": time.time() - t0}
    for dtype in dtypes:
        m = EnhancedUNet(3, dtype=dtype)
        m.load_state_dict(sd)
        m = m.cuda()
        m.train(train)
        if train:
            y = m(x.cuda())
            loss = combined_loss(y, t.cuda())
            loss.backward()
            r = {"loss": float(loss), "loss_ref": float(lref), **grad_table(m, gref)}
        else:
            with torch.no_grad():
                y = m(x.cuda())
            r = {}
        y = y.detach().cpu()
        r["logits_err"] = float((y - yref).abs().max() / yref.abs().max())
        r["logits_rms"] = float((y - yref).pow(2).mean().sqrt() / yref.abs().max())
        r["mask"] = mask_stats(y, yref)
        res[dtype] = r
        del m
        torch.cuda.empty_cache()
    OUT[name] = res
    print(name, json.dumps(res), flush=True)
    save()


def trained_state(steps=60, res=256, batch=8):
    """Briefly train the bf16 model on the synthetic bright-field task: weights with decided (non-tied) predictions."""
    torch.manual_seed(0)
    m = EnhancedUNet(3, dtype="bf16").cuda().train()
    opt = ClippedAdamW(list(m.parameters()), lr=1e-3, on_update=m._packs.invalidate)
    losses = []
    for i in range(steps):
        x, t = bench.synth_batch(batch, res, 77 + i, torch.device("cuda"))
        for p in m.parameters():
            p.grad = None
        loss = combined_loss(m(x), t)
        loss.backward()
        opt.step()
        losses.append(float(loss))
    OUT["trained_losses"] = losses[::10] + [losses[-1]]
    return {k: v.detach().cpu().clone() for k, v in m.state_dict().items()}


def main():
    torch.set_num_threads(os.cpu_count() or 8)
    import subprocess
    res = {}
    for combo in ("bf16,bf16", "f16,f16", "f16,bf16", "bf16,f16"):   # one process each: an illegal-instruction fault poisons the context
        r = subprocess.run([sys.executable, os.path.abspath(__file__), "--probe", combo], capture_output=True, text=True)
        res[combo] = (r.stdout.strip().splitlines() or ["?"])[-1] if r.returncode == 0 else f"rc={r.returncode}: {r.stderr.strip().splitlines()[-1] if r.stderr.strip() else ''}"
    OUT["mixed_format_mma"] = res
    print(res, flush=True)
    save()
    sd = oracle.make_state_dict(0, randomize_bn=True)
    x, t = bench.synth_batch(2, 256, 1234, torch.device("cpu"))
    parity_case("train_b2_256_synth", sd, x, t, True, ("bf16", "fp32"))
    parity_case("train_b2_256_rand", sd, oracle.make_input(2, 256, 256, 1), oracle.make_target(2, 256, 256, 2), True)
    x5, t5 = bench.synth_batch(2, 512, 1234, torch.device("cpu"))
    parity_case("train_b2_512_synth", sd, x5, t5, True)
    x1, _ = bench.synth_batch(1, 1024, 1234, torch.device("cpu"))
    parity_case("eval_b1_1024_synth", sd, x1, None, False)
    sdt = trained_state()
    parity_case("trained_train_b2_256", sdt, x, t, True)
    parity_case("trained_eval_b2_256", sdt, x, t, False)
    parity_case("trained_eval_b1_1024", sdt, x1, None, False)
    save()


if __name__ == "__main__":
    if len(sys.argv) > 2 and sys.argv[1] == "--probe":
        probe_mixed(sys.argv[2])
    else:
        main()
