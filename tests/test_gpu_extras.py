"""GPU parity for the callers either side of the model: the metrics drop-in (reference fixtures, bit-exact),
the inference post-processing kernels (softmax/resize + probability->mask cascade) and the optimiser class."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_metrics_dropin_bit_exact_against_reference_fixture(golden_dir):
    from enhanced_unet_b200 import metrics
    g = np.load(os.path.join(golden_dir, "metrics.npz"))
    keys = [str(k) for k in g["keys"]]
    names = sorted({k.split("/")[0] for k in g.files if "/" in k})
    for name in names:
        pred, gt = g[f"{name}/pred"], g[f"{name}/gt"]
        m = metrics.calculate_semantic_metrics(pred, gt)
        assert list(m.keys()) == keys
        assert np.array_equal(np.array([float(m[k]) for k in keys]), g[f"{name}/values"]), name   # bit-exact float64
        m2 = metrics.calculate_semantic_metrics(torch.from_numpy(pred).cuda(), torch.from_numpy(gt).cuda().to(torch.int64))
        assert all(float(m[k]) == float(m2[k]) for k in keys)
    # return types of the edge rules (python float 1.0 on the empty-union branch, numpy float64 otherwise)
    z = np.zeros((8, 8), np.int64)
    m = metrics.calculate_semantic_metrics(z, z)
    assert m["sem_live_iou"] == 1.0 and isinstance(m["sem_live_iou"], float)
    assert isinstance(metrics.calculate_semantic_metrics(g["kat1/pred"], g["kat1/gt"])["sem_live_iou"], np.floating)
    # binary helpers (metrics.py:12-26) on instance-style masks
    import oracle
    rng = np.random.default_rng(3)
    a, b = (rng.random((40, 50)) < 0.3).astype(np.uint8), (rng.random((40, 50)) < 0.4).astype(np.uint8)
    assert metrics.calculate_iou(a, b) == oracle.calculate_iou(a, b)
    assert metrics.calculate_dice(a, b) == oracle.calculate_dice(a, b)
    e = np.zeros((4, 4), np.uint8)
    assert metrics.calculate_iou(e, e) == 1.0 and metrics.calculate_dice(e, e) == 1.0
    # batched form and the 3x3 confusion matrix with ignore label
    pb, gb = rng.integers(0, 3, (5, 33, 17)), rng.integers(0, 3, (5, 33, 17))
    per = metrics.batch_semantic_metrics(pb, gb)
    for i in range(5):
        w = oracle.calculate_semantic_metrics(pb[i], gb[i])
        assert all(float(per[i][k]) == float(w[k]) for k in w)
    gi = gb.copy(); gi[rng.random(gi.shape) < 0.1] = 255
    cm = metrics.confusion_matrix_3x3(pb, gi)
    keep = gi != 255
    want = np.bincount(gi[keep] * 3 + pb[keep], minlength=9).reshape(3, 3)
    assert np.array_equal(cm, want)


def test_mask_cascade_matches_reference_fixture(golden_dir):
    from enhanced_unet_b200.lib import call
    g = np.load(os.path.join(golden_dir, "mask.npz"))
    names = sorted({k.split("/")[0] for k in g.files})
    probs = torch.from_numpy(np.stack([g[f"{n}/probs"] for n in names])).cuda().contiguous()     # [N,3,H,W]
    N, _, H, W = probs.shape
    mask = torch.empty(N, H, W, dtype=torch.uint8, device="cuda")
    counts = torch.empty(N, 2, dtype=torch.int32, device="cuda")
    call("eunet_probs_to_mask", probs.data_ptr(), mask.data_ptr(), counts.data_ptr(), N, H, W)
    for i, n in enumerate(names):
        assert np.array_equal(mask[i].cpu().numpy().astype(np.int64), g[f"{n}/mask"]), n      # bit-exact vs the reference


def test_softmax_probs_matches_torch():
    from enhanced_unet_b200.lib import call
    g = torch.Generator().manual_seed(1)
    logits = (torch.randn(3, 3, 48, 64, generator=g) * 3).cuda()
    want = torch.softmax(torch.nn.functional.interpolate(logits, size=(24, 32), mode="bilinear", align_corners=False), dim=1)
    probs = torch.empty(3, 3, 24, 32, device="cuda")
    call("eunet_softmax_probs", logits.data_ptr(), probs.data_ptr(), 3, 24, 32, 2)
    assert float((probs - want).abs().max()) < 2e-6
    probs1 = torch.empty_like(logits)
    call("eunet_softmax_probs", logits.data_ptr(), probs1.data_ptr(), 3, 48, 64, 1)
    assert float((probs1 - torch.softmax(logits, 1)).abs().max()) < 2e-6


def test_clipped_adamw_matches_torch_adamw_with_clip():
    from enhanced_unet_b200.optim import ClippedAdamW
    g = torch.Generator().manual_seed(2)
    shapes = [(64, 3, 3, 3), (64,), (128, 64, 3, 3), (3, 64, 1, 1)]
    ref = [torch.nn.Parameter(torch.randn(s, generator=g)) for s in shapes]
    ours = [torch.nn.Parameter(p.detach().clone().cuda()) for p in ref]
    o_ref = torch.optim.AdamW(ref, lr=4e-3, weight_decay=1e-4, betas=(0.9, 0.999))
    bumped = []
    o = ClippedAdamW(ours, lr=4e-3, weight_decay=1e-4, betas=(0.9, 0.999), max_norm=1.0, on_update=lambda: bumped.append(1))
    sched = torch.optim.lr_scheduler.LinearLR(o, start_factor=0.001, end_factor=1.0, total_iters=5)       # reference warm-up
    sched_ref = torch.optim.lr_scheduler.LinearLR(o_ref, start_factor=0.001, end_factor=1.0, total_iters=5)
    for it in range(4):
        for p, q in zip(ref, ours):
            p.grad = torch.randn(p.shape, generator=g) * (it + 1)
            q.grad = p.grad.clone().cuda()
        torch.nn.utils.clip_grad_norm_(ref, 1.0)
        o_ref.step(); o.step()
        sched.step(); sched_ref.step()
        for p, q in zip(ref, ours):
            assert float((q.detach().cpu() - p.detach()).abs().max() / p.detach().abs().max()) < 1e-5
    assert len(bumped) == 4
    sd = o.state_dict()
    assert set(sd["state"][0].keys()) == {"step", "exp_avg", "exp_avg_sq"}


@pytest.mark.parametrize("dtype,tol", [("fp32", 1e-4), ("fp16", 2e-2), ("bf16", 2e-2)])
def test_fusion_head_matches_reference_fixture(golden_dir, dtype, tol):
    """Secondary path a10: the in-file fusion blocks of the smp body, eval mode, against the fixture produced by
    re-instantiating the reference's own nn.Sequential blocks (oracle/make_golden.py:fusion_case)."""
    from oracle.unet_oracle import make_fusion_state_dict
    from enhanced_unet_b200.models import FusionHead
    g = np.load(os.path.join(golden_dir, "fusion.npz"))
    m = FusionHead(3, dtype=dtype)
    m.load_state_dict(make_fusion_state_dict(0), strict=True)
    m = m.cuda().eval()
    y = m(torch.from_numpy(g["main"]).cuda(), torch.from_numpy(g["aux"]).cuda())
    ref = torch.from_numpy(g["out"])
    err = float((y.cpu() - ref).abs().max() / ref.abs().max())
    print(f"[fusion {dtype}] err {err:.3e}")
    assert y.shape == ref.shape and err <= tol, err


@pytest.mark.parametrize("dtype,tol,gtol", [("fp32", 1e-4, 2e-3), ("fp16", 2e-2, 0.12)])
def test_fusion_head_training_forward_backward_matches_reference_fixture(golden_dir, dtype, tol, gtol):
    """a10 in TRAINING mode (batch-statistics BatchNorm, the reference's own Dropout2d draw) with its backward: output,
    gradients w.r.t. both inputs and every parameter, updated BN buffers - against the fixture produced by the reference's
    nn.Sequential blocks + autograd (oracle/make_golden.py:fusion_train_case)."""
    from oracle.unet_oracle import make_fusion_state_dict
    from enhanced_unet_b200.models import FusionHead
    g = np.load(os.path.join(golden_dir, "fusion_train.npz"))
    m = FusionHead(3, dtype=dtype)
    m.load_state_dict(make_fusion_state_dict(0), strict=True)
    m = m.cuda().train()
    a = torch.from_numpy(g["main"]).cuda().requires_grad_(True)
    b = torch.from_numpy(g["aux"]).cuda().requires_grad_(True)
    y = m(a, b, dropout_scales=(torch.from_numpy(g["s1"]), torch.from_numpy(g["s2"])))
    ref = torch.from_numpy(g["out"])
    err = float((y.detach().cpu() - ref).abs().max() / ref.abs().max())
    (y * torch.from_numpy(g["dout"]).cuda()).sum().backward()
    m.check_numerics()

    def rel(got, want):
        got, want = got.detach().cpu().double().flatten(), torch.as_tensor(want).double().flatten()
        return float((got - want).norm() / (want.norm() + 1e-30))

    worst = {"dmain": rel(a.grad, g["dmain"]), "daux": rel(b.grad, g["daux"])}
    for n, p in m.named_parameters():
        assert p.grad is not None and p.grad.shape == p.shape, n
        worst[n] = rel(p.grad, g["grad/" + n])
    print(f"[fusion train {dtype}] out err {err:.3e}; worst gradient relL2 {max(worst.values()):.3e} ({max(worst, key=worst.get)})")
    assert err <= tol, err
    assert all(v <= gtol for v in worst.values()), {k: v for k, v in worst.items() if v > gtol}
    sd = m.state_dict()
    for k in g.files:
        if k.startswith("buf/"):
            want = torch.from_numpy(g[k]).double()
            e = float((sd[k[4:]].cpu().double() - want).abs().max() / max(1.0, float(want.abs().max())))
            assert e <= (1e-4 if dtype == "fp32" else 5e-3), (k, e)
    # without a supplied draw the module draws its own Dropout2d factors: runs, finite, different from the fixture's
    y2 = m(a.detach(), b.detach())
    assert torch.isfinite(y2).all()


def test_trainer_and_evaluator_entry_points(tmp_path, monkeypatch):
    """The reference's training / inference entry points on reference-format batches: loss decreases over a few
    steps on a fixed batch, ragged masks take the per-sample path, predict_semantic_mask / evaluate return the
    reference's types, and the checkpoint round-trips (train_eval.py:66, 236, 570, 852, 1143-1151)."""
    from enhanced_unet_b200.models import get_model
    from enhanced_unet_b200.train_eval import Evaluator, SyntheticCellBatches, Trainer, evaluate_model, train_model
    monkeypatch.chdir(tmp_path)
    torch.manual_seed(0)
    model = get_model("enhanced_unet", num_classes=3, device="cuda").to("cuda")
    tr = Trainer(model, "cuda", "enhanced_unet", total_epochs=50)
    assert tr.warmup_epochs == 5 and abs(tr.optimizer.param_groups[0]["lr"] - 4e-6) < 1e-12      # SURVEY KAT-5
    batch = next(iter(SyntheticCellBatches(1, 2, 96, seed=3)))                                    # 96 -> reflect-padded to 96
    tr.warmup_scheduler.step()
    for g in tr.optimizer.param_groups:
        g["lr"] = 1e-3
    losses = [tr.train_epoch([batch]) for _ in range(6)]
    assert all(np.isfinite(losses)) and losses[-1] < losses[0], losses
    # ragged masks (different sizes) -> per-sample loss path; 72x72 image is reflect-padded to 96
    rb = {"images": torch.rand(2, 3, 72, 72), "batch_items": [{"semantic_mask": torch.randint(0, 3, (72, 72))},
                                                             {"semantic_mask": torch.randint(0, 3, (72, 72))}]}
    assert np.isfinite(tr.train_epoch([rb]))
    ev = Evaluator(model, "cuda", "enhanced_unet")
    mask = ev.predict_semantic_mask(batch["images"][0])
    assert isinstance(mask, np.ndarray) and mask.dtype == np.int64 and mask.shape == (96, 96) and set(np.unique(mask)) <= {0, 1, 2}
    probs = ev._run_model_single(torch.rand(3, 72, 80))
    assert probs.shape == (3, 72, 80) and float((probs.sum(0) - 1).abs().max()) < 1e-5
    res = ev.evaluate(SyntheticCellBatches(2, 2, 64, seed=5))
    assert set(res) == {"sem_background_iou", "sem_background_dice", "sem_live_iou", "sem_live_dice", "sem_dead_iou",
                        "sem_dead_dice", "sem_mean_iou", "sem_mean_iou_all", "sem_mean_dice"}
    ckpt = train_model("enhanced_unet", None, "cuda", num_epochs=3, train_batches=SyntheticCellBatches(2, 2, 64, seed=1),
                       val_batches=SyntheticCellBatches(1, 2, 64, seed=2))
    saved = torch.load(ckpt, map_location="cpu", weights_only=False)
    assert set(saved) >= {"epoch", "model_state_dict", "optimizer_state_dict", "scheduler_state_dict", "best_miou", "history"}
    assert len(saved["model_state_dict"]) == 109
    out = evaluate_model("enhanced_unet", None, "cuda", ckpt, batches=SyntheticCellBatches(1, 2, 64, seed=2))
    assert 0.0 <= out["sem_mean_iou_all"] <= 1.0


def test_instance_metrics_dropin_bit_exact(golden_dir):
    """calculate_instance_metrics (metrics.py:61-194) through the bit-plane pairwise-intersection kernels: float64-identical
    to the reference fixture and to the CPU oracle; the pairwise IoU matrix equals calculate_iou pair by pair."""
    import oracle
    from enhanced_unet_b200 import metrics
    g = np.load(os.path.join(golden_dir, "instances.npz"))
    for name, spec in oracle.INSTANCE_CASES.items():
        args = oracle.make_instance_case(*spec)
        m = metrics.calculate_instance_metrics(*args)
        keys = sorted(m.keys())
        assert keys == [str(k) for k in g[f"{name}/keys"]], name
        assert np.array_equal(np.array([float(m[k]) for k in keys]), g[f"{name}/values"]), name
    pm, _, _, gm, _ = oracle.make_instance_case(31, 37, 45, 12, 17)          # ragged size: 1665 pixels, not a multiple of 32
    pm.append(np.zeros((37, 45), np.uint8)); gm.append(np.zeros((37, 45), np.uint8))   # empty masks: the union == 0 rule
    iou = metrics.pairwise_iou(pm, gm)
    want = np.array([[float(oracle.calculate_iou(a, b)) for b in gm] for a in pm])
    assert np.array_equal(iou, want)
    # counts at a full-size shape: checksum of the intersection matrix against torch integer arithmetic
    from enhanced_unet_b200.ops import pair_intersections
    gen = torch.Generator(device="cuda").manual_seed(5)
    a = (torch.rand(40, 1024, 1024, device="cuda", generator=gen) < 0.3).to(torch.uint8)
    b = (torch.rand(24, 1024, 1024, device="cuda", generator=gen) < 0.2).to(torch.uint8)
    inter, aa, ab = pair_intersections(a, b)
    assert torch.equal(aa, a.flatten(1).long().sum(1)) and torch.equal(ab, b.flatten(1).long().sum(1))
    want = (a.flatten(1).float() @ b.flatten(1).float().t()).long()         # exact: counts < 2^24
    assert torch.equal(inter, want)


@pytest.mark.parametrize("case", [(2, 3, 48, 40, 0.75), (1, 3, 48, 40, 1.25), (2, 3, 36, 30, (48, 40)), (1, 2, 7, 5, (13, 11)),
                                  (1, 1, 64, 64, (32, 32)), (1, 3, 1, 1, (4, 3))])
def test_resize_bilinear_matches_torch(case):
    from enhanced_unet_b200.train_eval import Evaluator
    b, c, h, w, how = case
    x = torch.randn(b, c, h, w, device="cuda", generator=torch.Generator(device="cuda").manual_seed(1))
    if isinstance(how, tuple):
        got = Evaluator._resize(x, size=how)
        want = torch.nn.functional.interpolate(x, size=how, mode="bilinear", align_corners=False)
    else:
        got = Evaluator._resize(x, scale_factor=how)
        want = torch.nn.functional.interpolate(x, scale_factor=how, mode="bilinear", align_corners=False)
    assert got.shape == want.shape
    assert float((got - want).abs().max()) <= 2e-6


@pytest.mark.parametrize("dtype,tol", [("fp32", 2e-5), ("fp16", 1e-3), ("bf16", 4e-3)])
def test_tta_inference_matches_reference_fixture(golden_dir, dtype, tol):
    """Evaluator._run_tta_inference (train_eval.py:419-453): base + 2 flips + 0.75x / 1.25x views, against the reference."""
    import oracle
    from enhanced_unet_b200.models import EnhancedUNet
    from enhanced_unet_b200.train_eval import Evaluator
    g = np.load(os.path.join(golden_dir, "tta.npz"))
    m = EnhancedUNet(3, dtype=dtype)
    m.load_state_dict(oracle.make_state_dict(0), strict=True)
    ev = Evaluator(m.cuda().eval(), "cuda", "enhanced_unet", tta=True)
    for name in ("48x40", "64x64"):
        h, w, seed = [int(v) for v in g[f"{name}/meta"]]
        img = oracle.make_input(1, h, w, seed)[0].cuda()
        base = ev._run_model_single(img).cpu().numpy()
        tta = ev._run_tta_inference(img).cpu().numpy()
        assert tta.shape == (3, h, w)
        eb, et = np.abs(base - g[f"{name}/base"]).max(), np.abs(tta - g[f"{name}/tta"]).max()
        effect = np.abs(g[f"{name}/tta"] - g[f"{name}/base"]).max()
        print(f"[tta {dtype} {name}] base err {eb:.2e}, tta err {et:.2e} (tta-vs-base effect {effect:.2e})")
        assert eb <= tol and et <= tol
        if dtype == "fp32":
            assert et < 0.05 * effect          # the check resolves the augmentation itself, not just the base forward


def test_host_batch_prefetcher_round_trips():
    """data.HostBatchPrefetcher: double-buffered pinned-host -> device staging returns every submitted batch intact and in
    order, also when a consumer kernel is still reading the slot that is about to be refilled."""
    from enhanced_unet_b200.data import HostBatchPrefetcher
    pf = HostBatchPrefetcher("cuda")
    g = torch.Generator().manual_seed(3)
    batches = [(torch.rand(4, 3, 64, 64, generator=g).pin_memory(), torch.randint(0, 3, (4, 64, 64), generator=g).pin_memory())
               for _ in range(5)]
    sums = []
    pf.submit(*batches[0])
    for i in range(5):
        x, t = pf.get()
        if i + 1 < 5:
            pf.submit(*batches[i + 1])
        big = x.double()
        for _ in range(20):                  # keep the compute stream busy on this slot while the next copy is in flight
            big = big * 1.0000001
        sums.append((x.clone(), t.clone()))
    torch.cuda.synchronize()
    for (x, t), (hx, ht) in zip(sums, batches):
        assert torch.equal(x.cpu(), hx) and torch.equal(t.cpu(), ht)
    with pytest.raises(RuntimeError):
        pf.get()


def test_five_step_trajectory_matches_reference_optimiser():
    """Multi-step parity of the FULL training step (reference train_eval.py:236-353: forward -> combined loss -> backward ->
    clip_grad_norm_(1.0) -> AdamW -> BatchNorm running statistics), fp32 mode, five steps of ``Trainer.train_step`` against
    the oracle port driven by ``torch.optim.AdamW`` + ``clip_grad_norm_``.

    What can and cannot agree: AdamW's first steps are sign-like (update = lr * m / (sqrt(v) + eps)), so an element whose
    gradient is smaller than the fp32 summation-order difference of two correct implementations can move by 2 * lr in
    opposite directions; parameters are therefore compared by relative L2 and by the cosine of the 5-step update, the
    loss of every step (which sees all of it) at 1e-3.  The conv biases in front of a train-mode BatchNorm are excluded:
    their true gradient is zero, the reference random-walks them on fp32 rounding noise (+-lr per step), this path keeps
    them still - they cancel in training and enter eval only through running_mean, which is compared without them."""
    import re
    import oracle
    from enhanced_unet_b200.models import EnhancedUNet
    from enhanced_unet_b200.train_eval import Trainer
    pre_bn_bias = re.compile(r"^(model\.(enc|dec)[1234]\.(0|3)|enhance\.0)\.bias$")
    lr, steps = 2e-4, 5
    sd = oracle.make_state_dict(5, randomize_bn=False)
    xs = [oracle.make_input(2, 64, 64, 30 + i) for i in range(steps)]
    ts = [oracle.make_target(2, 64, 64, 40 + i) for i in range(steps)]
    # ---- reference side (CPU oracle + torch optimiser)
    params = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and "running" not in k else v.clone()) for k, v in sd.items()}
    plist = [p for p in params.values() if p.requires_grad]
    opt = torch.optim.AdamW(plist, lr=lr, weight_decay=1e-4, betas=(0.9, 0.999))
    ref_losses = []
    bn_of = lambda conv: re.sub(r"\.(\d)$", lambda mo: f".{int(mo.group(1)) + 1}", conv)      # model.enc1.0 -> model.enc1.1
    convs = [n[:-5] for n in sd if pre_bn_bias.match(n)]
    bias_ref = {c: [] for c in convs}
    bias_got = {c: [] for c in convs}
    for x, t in zip(xs, ts):
        opt.zero_grad()
        for c in convs:
            bias_ref[c].append(params[c + ".bias"].detach().clone())
        y, nb = oracle.unet_forward(params, x, train=True)
        loss = oracle.batch_loss(y, t)
        loss.backward()
        torch.nn.utils.clip_grad_norm_(plist, 1.0)
        opt.step()
        params.update(nb)
        ref_losses.append(float(loss))
    # ---- this repo
    m = EnhancedUNet(3, dtype="fp32")
    m.load_state_dict(sd)
    m = m.cuda().train()
    tr = Trainer(m, "cuda", "enhanced_unet", total_epochs=50)
    for g in tr.optimizer.param_groups:
        g["lr"] = lr
    losses = []
    for x, t in zip(xs, ts):
        for c in convs:
            bias_got[c].append(dict(m.named_parameters())[c + ".bias"].detach().cpu().clone())
        losses.append(float(tr.train_step(x, [t[0], t[1]])))
    print(f"[trajectory] losses {['%.5f' % v for v in losses]} vs {['%.5f' % v for v in ref_losses]}")
    # step 0 sees identical weights (1e-6); from then on the two fp32 trajectories drift apart as described above
    # (measured 2e-5, then 4e-4 .. 7e-4 per step at this learning rate, run-to-run variation included)
    for k, (a, b) in enumerate(zip(losses, ref_losses)):
        assert abs(a - b) <= (1e-6 if k == 0 else 1e-4 if k == 1 else 2e-3) * abs(b), (k, a, b)
    new = m.state_dict()
    worst_p, worst_c, worst_s = 0.0, 1.0, 0.0
    for name, ref in params.items():
        got = new[name].detach().cpu()
        if name.endswith("num_batches_tracked"):
            assert int(got) == steps == int(ref), name
        elif "running_" in name:
            ref_t, got_t = ref.double(), got.double()
            if name.endswith("running_mean"):
                # running_mean = momentum average of (batch mean of the bias-free conv + conv bias); the reference random-walks
                # that bias on rounding noise (docstring), so its recorded history is taken out on both sides before comparing
                conv = [c for c in convs if bn_of(c) + ".running_mean" == name][0]
                for hist, tt in ((bias_ref[conv], ref_t), (bias_got[conv], got_t)):
                    for k, b in enumerate(hist):
                        tt -= 0.1 * 0.9 ** (steps - 1 - k) * b.double()
            e = float((got_t - ref_t).abs().max() / max(float(ref_t.abs().max()), 1e-3))
            worst_s = max(worst_s, e)
            assert e <= 5e-2, (name, e)       # measured <= 2.3e-2 (enc4.1: 128 values per channel at this size), deterministic
        elif not pre_bn_bias.match(name):
            p0 = sd[name].double()
            dr, dg = ref.detach().double() - p0, got.double() - p0
            cos = float((dr * dg).sum() / (dr.norm() * dg.norm() + 1e-300))
            worst_c = min(worst_c, cos)
            assert cos >= 0.85, (name, cos)                         # (64-element BN vectors: a handful of sign flips)
            if float(p0.norm()) > 0:                                # (BN betas start at exactly 0: only the update exists)
                rel = float((got.double() - ref.detach().double()).norm() / ref.detach().double().norm())
                worst_p = max(worst_p, rel)
                assert rel <= 2e-2, (name, rel)     # measured 5e-3 = ~10 % of the 5-step update: gradients of two correct fp32 paths differ by ~1e-3
    print(f"[trajectory] worst parameter relL2 {worst_p:.2e}, worst update cosine {worst_c:.4f}, worst running-statistic err {worst_s:.2e}")


def test_graphed_train_step_matches_eager_steps():
    """graph.GraphedTrainStep (the whole step as one CUDA-graph launch, AdamW step counter / learning rate in device memory)
    against the same steps launched kernel by kernel: losses of every step, parameters, optimiser step count and BatchNorm
    counters after four steps with a learning-rate change in between."""
    import oracle
    from enhanced_unet_b200.graph import GraphedTrainStep
    from enhanced_unet_b200.models import EnhancedUNet
    from enhanced_unet_b200.ops import combined_loss
    from enhanced_unet_b200.optim import ClippedAdamW
    sd = oracle.make_state_dict(6)
    xs = [oracle.make_input(2, 64, 64, 50 + i).cuda() for i in range(6)]
    ts = [oracle.make_target(2, 64, 64, 60 + i).cuda() for i in range(6)]
    lrs = [2e-4, 2e-4, 2e-4, 1e-4, 1e-4, 1e-4]

    def fresh():
        m = EnhancedUNet(3, dtype="fp32")        # fp32 mode: deterministic enough for a tight comparison
        m.load_state_dict(sd)
        m = m.cuda().train()
        return m, ClippedAdamW(m.parameters(), lr=lrs[0])

    m1, o1 = fresh()
    eager = []
    for x, t, lr in zip(xs, ts, lrs):
        o1.param_groups[0]["lr"] = lr
        o1.zero_grad(set_to_none=True)
        loss = combined_loss(m1(x), t)
        loss.backward()
        o1.step()
        eager.append(float(loss))
    m2, o2 = fresh()
    g = GraphedTrainStep(m2, o2, 2, 64, 64, warmup_steps=2, example=(xs[0], ts[0]))     # two warm-up steps == steps 0, 1 ... on batch 0
    # the warm-up trained on batch 0 twice; restart both sides from the same state for the comparison
    m2.load_state_dict(sd)
    for st in o2.state.values():                 # in place: the graph holds the addresses of these tensors
        st["exp_avg"].zero_(); st["exp_avg_sq"].zero_(); st["step"] = 0
    o2._dev_state["step"].zero_()
    got = []
    for x, t, lr in zip(xs, ts, lrs):
        o2.param_groups[0]["lr"] = lr
        got.append(float(g(x, t)))
    assert g.launches_per_step > 100
    # both sides are this repo: the first step must agree to fp32 summation order; after that AdamW's sign-like updates
    # amplify the atomics' run-to-run differences (two eager runs differ by as much), hence the widening tolerance
    for k, (a, b) in enumerate(zip(got, eager)):
        assert abs(a - b) <= (1e-5 if k == 0 else 1e-4 if k == 1 else 2e-3) * abs(b), (k, got, eager)
    for (n, p), (_, q) in zip(m1.state_dict().items(), m2.state_dict().items()):
        if n.endswith("num_batches_tracked"):
            assert int(p) == int(q) == 6, n
        else:
            e = float((p.double() - q.double()).norm() / (p.double().norm() + 1e-30))
            # running_mean carries the pre-BN conv bias, which has no gradient signal and random-walks under AdamW (see
            # test_five_step_trajectory...): two runs of this repo differ by 5.2e-3 .. 5.6e-3 on dec4.1
            assert e <= (2e-2 if n.endswith("running_mean") else 5e-3), (n, e)
    assert int(o2._dev_state["step"]) == 6 and all(int(o2.state[p]["step"]) == 6 for p in m2.parameters())
    # eager inference after graph replays sees the CURRENT weights (packed-filter caches were invalidated)
    m1.eval(); m2.eval()
    with torch.no_grad():
        a, b = m1(xs[0]), m2(xs[0])
    assert float((a - b).abs().max() / a.abs().max()) <= 2e-2


@pytest.mark.parametrize("shape", [(2, 3, 72, 80), (1, 3, 33, 47), (3, 1, 96, 100), (1, 3, 64, 64)])
def test_reflect_pad_matches_torch(shape):
    """train_eval._pad32 (the reflect pad to multiples of 32 in front of the model, reference train_eval.py:249-253, 400-406)
    on the pad kernel: bit-identical to F.pad(mode='reflect')."""
    from enhanced_unet_b200.train_eval import _pad32
    g = torch.Generator(device="cuda").manual_seed(shape[2])
    x = torch.rand(shape, device="cuda", generator=g)
    got, hp, wp = _pad32(x)
    h, w = shape[2:]
    assert (hp, wp) == ((32 - h % 32) % 32, (32 - w % 32) % 32)
    want = torch.nn.functional.pad(x, (0, wp, 0, hp), mode="reflect") if (hp or wp) else x
    assert got.shape == want.shape and torch.equal(got, want)
    with pytest.raises(RuntimeError):
        _pad32(torch.rand(1, 3, 8, 40, device="cuda"))          # pad 24 >= height 8: F.pad raises too
