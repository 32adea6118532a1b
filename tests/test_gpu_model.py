"""Whole-path parity of the drop-in EnhancedUNet (CUDA kernels through the C ABI) against the CPU oracle and
the fixtures generated from the unmodified reference (tests/golden/model_*.npz).

Tolerances (BASELINE.json north_star): logits max|a-b| / max|b| <= 1e-4 in fp32 mode and <= 2e-2 in the 16-bit
tensor-core mode - in EVAL AND TRAIN mode; thresholded (argmax of the 2x2-mean-resized logits) masks agree on
>= 99.9 % of pixels; metric counts are bit-exact.

Modes: "fp16" is the product's 16-bit mode (default): fp16 tensors, fp32 accumulation.  "bf16" is kept as an option;
its 8 mantissa bits do not hold 2e-2 through sixteen train-mode BatchNorms (3e-2 .. 5e-2 at every size from 32^2 to
512^2, measured; PyTorch's own bf16 autocast of the reference: 4e-2 .. 7e-2), so its train-mode check is against that
yardstick only.

Masks at RANDOM-INIT weights: the reference's own top-2 logit margin is below 1.5e-3 of the logit range on 1 % of the
pixels (near ties; measured on the oracle), so no 16-bit arithmetic can agree on 99.9 % of ALL pixels there - fp16 WEIGHT
rounding alone flips 0.11 %.  The tests therefore assert (a) >= 99.9 % of all pixels in eval mode and for (briefly)
TRAINED weights in both modes, and (b) at random init in train mode: 100 % agreement on every pixel whose reference margin
exceeds twice the logit tolerance, i.e. wherever the reference's decision is not itself inside the tolerance band."""
import os
import re

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

LOGIT_TOL = {"fp32": 1e-4, "fp16": 2e-2, "bf16": 2e-2}
MODES = ["fp32", "fp16", "bf16"]
PRE_BN_BIAS = re.compile(r"^(model\.(enc|dec)[1234]\.(0|3)|enhance\.0)\.bias$")


def nerr(got, want) -> float:
    got = torch.as_tensor(got).double().cpu()
    want = torch.as_tensor(want).double().cpu()
    return float((got - want).abs().max() / (want.abs().max() + 1e-12))


def _model(dtype, sd):
    from enhanced_unet_b200.models import EnhancedUNet
    m = EnhancedUNet(3, dtype=dtype)
    m.load_state_dict(sd, strict=True)
    return m.cuda()


def _mask(logits):
    return torch.nn.functional.avg_pool2d(logits.float().cpu(), 2).argmax(1)


def mask_agreement(y, ref, tol):
    """(agreement over all pixels, agreement over the pixels whose REFERENCE top-2 margin exceeds 2 * tol * max|ref|,
    fraction of such decidable pixels) for the base prediction argmax(2x2 mean) (train_eval.py:471)."""
    pr = torch.nn.functional.avg_pool2d(torch.as_tensor(ref).float().cpu(), 2)
    agree = _mask(y) == pr.argmax(1)
    top2 = pr.topk(2, dim=1).values
    dec = (top2[:, 0] - top2[:, 1]) > 2 * tol * float(torch.as_tensor(ref).abs().max())
    return agree.float().mean().item(), (agree[dec].float().mean().item() if dec.any() else 1.0), dec.float().mean().item()


def grad_report(model, ref_grads):
    """Per-tensor (name, rel-L2, cosine) and the global pair over all compared gradients (pre-BN conv biases excluded:
    exact zeros here, fp32 rounding noise in the reference)."""
    rows, A, R = [], [], []
    for name, p in model.named_parameters():
        assert p.grad is not None and p.grad.dtype == torch.float32 and p.grad.shape == p.shape, name
        gr = p.grad.detach().cpu().double().flatten()
        if PRE_BN_BIAS.match(name):
            assert float(gr.abs().max()) < 1e-3, name
            continue
        rf = ref_grads[name].double().flatten()
        rows.append((name, float((gr - rf).norm() / (rf.norm() + 1e-30)), float((gr * rf).sum() / (gr.norm() * rf.norm() + 1e-30))))
        A.append(gr)
        R.append(rf)
    a, r = torch.cat(A), torch.cat(R)
    return rows, float((a - r).norm() / r.norm()), float((a * r).sum() / (a.norm() * r.norm()))


@pytest.fixture(scope="module")
def trained_sd():
    """Weights with decided predictions: 60 optimisation steps of the product's default mode on the synthetic bright-field
    task (how they were obtained is irrelevant to parity - both implementations are handed the same state_dict)."""
    import bench
    from enhanced_unet_b200.models import EnhancedUNet
    from enhanced_unet_b200.ops import combined_loss
    from enhanced_unet_b200.optim import ClippedAdamW
    torch.manual_seed(0)
    m = EnhancedUNet(3).cuda().train()
    opt = ClippedAdamW(list(m.parameters()), lr=1e-3)
    for i in range(60):
        x, t = bench.synth_batch(8, 256, 77 + i, torch.device("cuda"))
        opt.zero_grad(set_to_none=True)
        combined_loss(m(x), t).backward()
        opt.step()
    m.check_numerics()
    return {k: v.detach().cpu().clone() for k, v in m.state_dict().items()}


@pytest.mark.parametrize("dtype", MODES)
@pytest.mark.parametrize("case", ["b2_32x32", "b1_16x24", "b2_64x64"])
def test_forward_eval_and_train_match_reference_fixture(golden_dir, dtype, case):
    import oracle
    g = np.load(os.path.join(golden_dir, f"model_{case}.npz"))
    b, h, w, pseed, xseed, tseed = [int(v) for v in g["meta"]]
    sd = oracle.make_state_dict(pseed)
    x = oracle.make_input(b, h, w, xseed).cuda()
    m = _model(dtype, sd).eval()
    with torch.no_grad():
        y = m(x)
    assert y.shape == (b, 3, 2 * h, 2 * w) and y.dtype == torch.float32
    ref = torch.from_numpy(g["logits_eval"])
    print(f"[{dtype} {case}] eval logits err {nerr(y, ref):.3e} (torch autocast bf16: {float(g['autocast/logits_eval_err']):.3e})")
    assert nerr(y, ref) <= LOGIT_TOL[dtype], ("eval", nerr(y, ref))
    assert (_mask(y) == _mask(ref)).float().mean().item() >= 0.999
    assert m.get_aux_outputs() is None
    # train mode: batch statistics + running-stat update
    m = _model(dtype, sd).train()
    with torch.no_grad():
        y = m(x)
    ref = torch.from_numpy(g["logits_train"])
    # train-mode (batch-statistics) BN makes the logits ~20x more sensitive to rounding than eval mode: bf16
    # rounding of the conv WEIGHTS alone moves them by 2.5e-2 and PyTorch's own bf16 autocast of the reference by
    # 4e-2..7e-2 (fixture).  fp32 mode must meet 1e-4; bf16 train mode must beat PyTorch's bf16 and stay <= 6e-2.
    ac = float(g["autocast/logits_train_err"])
    tol = min(6e-2, ac) if dtype == "bf16" else LOGIT_TOL[dtype]     # bf16: optional mode, yardstick only (module docstring)
    e = nerr(y, ref)
    raw, dec, frac = mask_agreement(y, ref, LOGIT_TOL[dtype])
    print(f"[{dtype} {case}] train logits err {e:.3e} (torch autocast bf16: {ac:.3e}); mask agreement {raw:.5f} "
          f"(on the {frac:.3f} of pixels with a decidable reference margin: {dec:.5f})")
    assert e <= tol, ("train", e, tol)
    if dtype == "fp32":
        assert raw >= 0.999
    elif dtype == "fp16":
        assert dec == 1.0, ("train masks on decidable pixels", dec)
    new_sd = m.state_dict()
    stat_tol = {"fp32": 1e-4, "fp16": 3e-3, "bf16": 2e-2}[dtype]
    for k in g.files:
        if k.startswith("buf/"):
            name = k[4:]
            if name.endswith("num_batches_tracked"):
                assert int(new_sd[name]) == int(g[k])
            else:
                assert nerr(new_sd[name], g[k]) <= stat_tol, (name, nerr(new_sd[name], g[k]))


def _oracle_grads(sd, x, t):
    import oracle
    params = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and "running" not in k else v) for k, v in sd.items()}
    y, _ = oracle.unet_forward(params, x, train=True)
    loss = oracle.batch_loss(y, t)
    loss.backward()
    return loss.item(), {k: p.grad for k, p in params.items() if p.requires_grad}


@pytest.mark.parametrize("dtype", MODES)
@pytest.mark.parametrize("case", ["b2_32x32", "b1_16x24", "b2_64x64"])
def test_loss_and_gradients_match_reference(golden_dir, dtype, case):
    """Loss + every parameter gradient after one fwd+bwd.  The gradient of this network is chaotic at these
    tiny sizes (ReLU / max-pool kinks under small-batch BatchNorm: a 1e-6 relative weight perturbation of
    the REFERENCE moves single gradient entries by 3 %), so tensors are compared by relative
    L2 error / cosine similarity; for bf16 the yardstick is PyTorch's own bf16 autocast run of the
    reference, recorded in the fixture."""
    import oracle
    from enhanced_unet_b200.ops import combined_loss
    g = np.load(os.path.join(golden_dir, f"model_{case}.npz"))
    b, h, w, pseed, xseed, tseed = [int(v) for v in g["meta"]]
    sd = oracle.make_state_dict(pseed)
    x = oracle.make_input(b, h, w, xseed)
    t = oracle.make_target(b, h, w, tseed)
    ref_loss, ref_grads = _oracle_grads(sd, x, t)
    assert abs(ref_loss - float(g["loss"])) <= 1e-5 * abs(ref_loss)          # oracle == reference fixture
    m = _model(dtype, sd).train()
    y = m(x.cuda())
    loss = combined_loss(y, t.cuda())
    loss_tol = {"fp32": 1e-4, "fp16": 2e-3}.get(dtype, max(3e-2, 2 * abs(float(g["autocast/loss"]) - ref_loss) / abs(ref_loss)))
    assert abs(loss.item() - ref_loss) <= loss_tol * abs(ref_loss), (loss.item(), ref_loss)
    loss.backward()
    m.check_numerics()
    rows, grel, gcos = grad_report(m, ref_grads)
    bad = []
    for name, rel, cos in rows:
        ac_cos, ac_rel = float(g["autocast/gcos/" + name]), float(g["autocast/grel/" + name])
        if dtype == "fp32":
            ok = rel <= 3e-2 and cos >= 0.9995
        elif dtype == "fp16":
            # random-init weights at 16^2 .. 64^2: kink-dominated (module docstring); the tight bars are asserted at
            # BASELINE-scale shapes and on trained weights in test_parity_at_baseline_shapes
            ok = rel <= 0.30 and cos >= 0.96
        else:
            ok = (1 - cos) <= 2.0 * (1 - ac_cos) + 0.02 and rel <= 2.0 * ac_rel + 0.1
        if not ok:
            bad.append((name, round(rel, 4), round(cos, 5), round(ac_rel, 4), round(ac_cos, 5)))
    worst = sorted(rows, key=lambda r: r[2])[:5]
    print(f"[{dtype} {case}] global relL2 {grel:.4f} cos {gcos:.5f}; worst (name, relL2, cos):", [(n, round(a, 4), round(c, 4)) for n, a, c in worst])
    assert not bad, (dtype, case, bad[:8])
    if dtype == "fp16":
        assert grel <= 0.18 and gcos >= 0.985, (grel, gcos)    # (B=1, 16x24: six values per channel at the bottleneck: 0.129 / 0.9917)


def test_state_dict_round_trip_and_error_paths():
    import oracle
    from enhanced_unet_b200.models import EnhancedUNet, get_model
    sd = oracle.make_state_dict(1)
    m = EnhancedUNet(3)
    assert list(m.state_dict().keys()) == list(sd.keys())
    m.load_state_dict(sd, strict=True)
    for k, v in m.state_dict().items():
        assert torch.equal(v, sd[k]) and v.dtype == sd[k].dtype and v.shape == sd[k].shape
    m = m.cuda()
    with pytest.raises(RuntimeError):
        m(torch.rand(1, 1, 32, 32, device="cuda"))       # 1 channel: the reference raises too
    with pytest.raises(RuntimeError):
        m(torch.rand(1, 3, 36, 32, device="cuda"))       # not a multiple of 8
    with pytest.raises(RuntimeError):
        m(torch.rand(1, 3, 32, 32))                      # CPU tensor: no fallback
    with pytest.raises(ValueError):
        get_model("nope")


@pytest.mark.parametrize("dtype", MODES)
def test_train_logits_do_not_depend_on_batch_replication(dtype):
    """Duplicating a sample leaves every batch statistic unchanged, so train-mode results must not move.  Guards the
    cross-CTA statistics accumulation (fp32 partials of <= 32 addends per thread, fp64 across threads / CTAs): fp32
    running sums over a persistent CTA's whole lifetime once cost 1.6e-2 on the logits.
    fp32 mode: logits agree to 1e-5.  bf16 mode: the statistics of the first BatchNorm (identical inputs in both runs)
    agree to 1e-6; deeper layers and the logits see bf16 re-rounding of activations whenever a statistic moves in its
    last fp32 bit (the grouping of fp32 partials depends on the batch size), which train-mode BN amplifies exactly as it
    amplifies any bf16 rounding (see test_forward_eval_and_train_match_reference_fixture): the logits must stay
    inside the bf16 train-mode budget."""
    import oracle
    sd = oracle.make_state_dict(2)
    x1 = oracle.make_input(1, 256, 256, 3).cuda()
    m1, m4 = _model(dtype, sd).train(), _model(dtype, sd).train()
    with torch.no_grad():
        y1 = m1(x1)
        y4 = m4(x1.expand(4, 3, 256, 256).contiguous())
    s1, s4 = m1.state_dict(), m4.state_dict()
    for key in ("model.enc1.1.running_mean", "model.enc1.1.running_var"):
        # the unbiased-variance factor n/(n-1) differs between the two batch sizes by 4e-6
        assert nerr(s4[key], s1[key]) <= (1e-5 if key.endswith("var") else 1e-6), (key, nerr(s4[key], s1[key]))
    tol = {"fp32": 1e-5, "fp16": 2e-2, "bf16": 6e-2}[dtype]
    e0, e3 = nerr(y4[0], y1[0]), nerr(y4[3], y1[0])
    print(f"[{dtype}] batch-replication logit difference {e0:.3e} / {e3:.3e}")
    assert e0 <= tol and e3 <= tol, (e0, e3)
    assert nerr(y4[3], y4[0]) <= 1e-6      # replicas inside one batch are identical


def test_full_size_properties():
    """BASELINE config-2 shape on one GPU (batch 16, 512x512, 16-bit train step): size-independent properties -
    finite logits, per-channel BN statistics consistency, gradient of a batch-replicated input equals the
    single-sample gradient structure (loss invariance under batch duplication)."""
    import oracle
    from enhanced_unet_b200.ops import combined_loss
    sd = oracle.make_state_dict(2)
    m = _model("fp16", sd).train()
    x1 = oracle.make_input(1, 512, 512, 3).cuda()
    t1 = oracle.make_target(1, 512, 512, 4).cuda()
    y1 = m(x1)
    l1 = combined_loss(y1, t1)
    l1.backward()
    g1 = {n: p.grad.clone() for n, p in m.named_parameters()}
    m.zero_grad(set_to_none=True)
    x = x1.expand(16, 3, 512, 512).contiguous()
    t = t1.expand(16, 512, 512).contiguous()
    y = m(x)
    assert y.shape == (16, 3, 1024, 1024) and torch.isfinite(y).all()
    # duplicated samples: batch statistics are identical, so logits / loss / gradients must be too
    assert nerr(y[7], y1[0]) < 2e-2
    loss = combined_loss(y, t)
    assert abs(loss.item() - l1.item()) <= 2e-2 * abs(l1.item())
    loss.backward()
    for n, p in m.named_parameters():
        if PRE_BN_BIAS.match(n):
            continue
        assert torch.isfinite(p.grad).all(), n
        assert nerr(p.grad, g1[n]) < 0.2, (n, nerr(p.grad, g1[n]))
    m.check_numerics()


@pytest.mark.parametrize("shape", [(4, 1024, 1024), (1, 2048, 2048)])
def test_inference_configs_masks_and_counts(shape):
    """BASELINE configs 4 / 5 (high-res and whole-slide inference, reduced batch): eval-mode bf16 logits stay within
    2e-2 of the fp32-mode path (itself pinned to the reference at 1e-6), thresholded masks agree on >= 99.9 % of the
    pixels, and the integer metric pipeline (logits -> probabilities -> mask cascade -> confusion counts) is
    bit-identical between a batched device run and the per-image CPU oracle."""
    import oracle
    from enhanced_unet_b200.train_eval import Evaluator
    from enhanced_unet_b200.ops import confusion_counts
    b, h, w = shape
    sd = oracle.make_state_dict(4)
    x = oracle.make_input(b, h, w, 5).cuda()
    m16, m32 = _model("fp16", sd).eval(), _model("fp32", sd).eval()
    with torch.no_grad():
        y16 = m16(x)
        y32 = m32(x)
    assert y16.shape == (b, 3, 2 * h, 2 * w)
    err = nerr(y16, y32)
    agree = (torch.nn.functional.avg_pool2d(y16, 2).argmax(1) == torch.nn.functional.avg_pool2d(y32, 2).argmax(1)).float().mean().item()
    print(f"[infer {shape}] fp16 vs fp32-mode logits err {err:.3e}, mask agreement {agree:.5f}")
    assert err <= 2e-2 and agree >= 0.999
    del y32, m32
    ev = Evaluator(m16, "cuda", "enhanced_unet", tta=False)
    probs = ev._probs(x)
    masks = ev._convert_probs_to_mask_device(probs)                      # uint8 [B,h,w]
    gt = oracle.make_target(b, h, w, 6).to(torch.uint8).cuda()
    cm = confusion_counts(masks, gt)
    assert torch.equal(cm.sum((1, 2)), torch.full((b,), h * w, device="cuda"))
    # per-image CPU oracle on the first image: cascade and counts must be bit-identical
    want_mask = oracle.convert_probs_to_mask(probs[0].cpu().numpy())
    assert np.array_equal(masks[0].cpu().numpy().astype(np.int64), want_mask)
    assert np.array_equal(cm[0].cpu().numpy(), oracle.confusion_counts(want_mask[None], gt[0].cpu().numpy()[None])[0])


def test_flat_gradient_sink_matches_autograd_path():
    """Data-parallel plumbing at world size 1: with parallel.GradientAllReduce attached, backward writes every gradient
    into ONE flat buffer (production order) and sets .grad to views of it - the same numbers as the autograd path."""
    import oracle
    from enhanced_unet_b200 import parallel
    from enhanced_unet_b200.ops import combined_loss
    sd = oracle.make_state_dict(3)
    x = oracle.make_input(2, 64, 64, 7).cuda()
    t = oracle.make_target(2, 64, 64, 8).cuda()
    m = _model("fp16", sd).train()
    combined_loss(m(x), t).backward()
    want = {n: p.grad.clone() for n, p in m.named_parameters()}
    m2 = _model("fp16", sd).train()
    ar = parallel.GradientAllReduce(m2)
    for step in range(2):                       # the buffer is reused from step to step
        m2.zero_grad(set_to_none=True)
        m2.load_state_dict(sd, strict=True)     # reset the running statistics; parameters are unchanged
        combined_loss(m2(x), t).backward()
        ar.wait()
        lo, hi = ar.buffer.flat.data_ptr(), ar.buffer.flat.data_ptr() + 4 * ar.buffer.numel
        for n, p in m2.named_parameters():
            assert p.grad is not None and lo <= p.grad.data_ptr() < hi, n
            assert nerr(p.grad, want[n]) <= 1e-4, (step, n, nerr(p.grad, want[n]))   # atomics: summation order only
    with pytest.raises(RuntimeError):           # accumulation over two backward passes is refused, not silently wrong
        combined_loss(m2(x), t).backward()


@pytest.mark.parametrize("shape", [(2, 8, 8), (2, 24, 40), (1, 40, 72)])
def test_ragged_small_shapes_against_oracle(shape):
    """Smallest legal image (8x8: 1x1 at the bottleneck; batch 2 because train-mode BatchNorm needs > 1 value per channel) and sizes that are multiples of 8 but not of the 16x8 / 64x8 pixel
    tiles: every kernel runs its overhanging-tile / TMA zero-fill / clipped-store path.  fp32 mode against the CPU oracle
    at 1e-4; bf16 mode (tensor-core kernels incl. the fused tail backward) must stay finite and inside the train budget."""
    import oracle
    from enhanced_unet_b200.ops import combined_loss
    b, h, w = shape
    sd = oracle.make_state_dict(9)
    x, t = oracle.make_input(b, h, w, 10), oracle.make_target(b, h, w, 11)
    ref_loss, ref_grads = _oracle_grads(sd, x, t)
    with torch.no_grad():
        y_eval_ref, _ = oracle.unet_forward(sd, x, train=False)
    for dtype in ("fp32", "fp16", "bf16"):
        m = _model(dtype, sd).eval()
        with torch.no_grad():
            e_eval = nerr(m(x.cuda()), y_eval_ref)
        assert e_eval <= LOGIT_TOL[dtype], (dtype, shape, e_eval)
        m.train()
        loss = combined_loss(m(x.cuda()), t.cuda())
        loss.backward()
        assert abs(loss.item() - ref_loss) <= (1e-4 if dtype == "fp32" else 8e-2) * abs(ref_loss), (dtype, loss.item(), ref_loss)
        for name, p in m.named_parameters():
            assert torch.isfinite(p.grad).all(), (dtype, name)
            if dtype == "fp32" and not PRE_BN_BIAS.match(name):
                gr, rf = p.grad.cpu().double().flatten(), ref_grads[name].double().flatten()
                rel = float((gr - rf).norm() / (rf.norm() + 1e-30))
                assert rel <= 3e-2, (name, rel)


# ---------------------------------------------------------------------------------------------
# BASELINE-scale parity against the CPU oracle: reference models.py:227-238, 334-339 and
# train_eval.py:306-338 (loss) on identical weights and inputs; the oracle takes 1 - 3 s per case on the box's host cores
# ---------------------------------------------------------------------------------------------
def _baseline_case(sd, b, res, train, dtype):
    import bench
    import oracle
    from enhanced_unet_b200.ops import combined_loss
    torch.set_num_threads(os.cpu_count() or 8)
    x, t = bench.synth_batch(b, res, 1234, torch.device("cpu"))       # the benchmark's synthetic bright-field generator
    if train:
        ref_loss, ref_grads = None, None
        params = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and "running" not in k else v) for k, v in sd.items()}
        yref, _ = oracle.unet_forward(params, x, train=True)
        lref = oracle.batch_loss(yref, t)
        lref.backward()
        ref_loss, ref_grads = float(lref), {k: p.grad for k, p in params.items() if getattr(p, "grad", None) is not None}
        yref = yref.detach()
    else:
        with torch.no_grad():
            yref, _ = oracle.unet_forward(sd, x, train=False)
    m = _model(dtype, sd)
    m.train(train)
    out = {}
    if train:
        y = m(x.cuda())
        loss = combined_loss(y, t.cuda())
        loss.backward()
        out["loss"], out["ref_loss"] = float(loss), ref_loss
        out["rows"], out["grel"], out["gcos"] = grad_report(m, ref_grads)
    else:
        with torch.no_grad():
            y = m(x.cuda())
    m.check_numerics()
    out["err"] = nerr(y, yref)
    out["mask"] = mask_agreement(y.detach(), yref, LOGIT_TOL[dtype])
    return out


@pytest.mark.parametrize("weights", ["random_init", "trained"])
@pytest.mark.parametrize("b,res", [(2, 256), (2, 512)])
def test_train_parity_at_baseline_shapes(trained_sd, weights, b, res):
    """16-bit (fp16) TRAIN-mode forward + loss + backward at B=2, 3x256^2 (BASELINE config-1 shape) and 3x512^2 (the
    headline resolution) against the CPU oracle: logits <= 2e-2 (north star, no train-mode carve-out), loss, masks, and
    every parameter gradient by relative L2 / cosine.  Gradient bars: 5e-2 / 0.995 per tensor on trained weights; at
    random init the gradient is kink-dominated (the reference moves by 4e-3 under an exact reformulation of its own
    arithmetic), there the global pair and a looser per-tensor bound are asserted."""
    import oracle
    sd = oracle.make_state_dict(0) if weights == "random_init" else trained_sd
    r = _baseline_case(sd, b, res, True, "fp16")
    raw, dec, frac = r["mask"]
    worst = sorted(r["rows"], key=lambda t: -t[1])[:4]
    print(f"[fp16 train {weights} B={b} {res}^2] logits err {r['err']:.3e}; loss {r['loss']:.6f} vs {r['ref_loss']:.6f}; masks {raw:.5f} "
          f"(decidable {frac:.3f}: {dec:.5f}); grads global relL2 {r['grel']:.4f} cos {r['gcos']:.5f}; worst {[(n, round(a, 4), round(c, 5)) for n, a, c in worst]}")
    assert r["err"] <= 2e-2, r["err"]
    assert abs(r["loss"] - r["ref_loss"]) <= 1e-3 * abs(r["ref_loss"])
    assert dec == 1.0
    if weights == "trained":
        assert raw >= 0.999, raw
        for name, rel, cos in r["rows"]:
            assert rel <= 5e-2 and cos >= 0.995, (name, rel, cos)
    else:
        assert r["grel"] <= 0.12 and r["gcos"] >= 0.993, (r["grel"], r["gcos"])
        for name, rel, cos in r["rows"]:
            assert rel <= 0.30 and cos >= 0.96, (name, rel, cos)


@pytest.mark.parametrize("weights", ["random_init", "trained"])
def test_eval_parity_at_1024(trained_sd, weights):
    """Eval-mode forward at B=1, 3x1024^2 (BASELINE config-4 resolution) against the CPU oracle: logits <= 2e-2 and
    >= 99.9 % of ALL mask pixels."""
    import oracle
    sd = oracle.make_state_dict(0) if weights == "random_init" else trained_sd
    r = _baseline_case(sd, 1, 1024, False, "fp16")
    raw, dec, frac = r["mask"]
    print(f"[fp16 eval {weights} 1x1024^2] logits err {r['err']:.3e}; masks {raw:.5f}")
    assert r["err"] <= 2e-2 and raw >= 0.999


def test_fp32_mode_parity_at_256():
    """fp32 mode at the BASELINE config-1 shape (B=2, 3x256^2, train): 1e-4 on the logits, tight gradients."""
    import oracle
    r = _baseline_case(oracle.make_state_dict(0), 2, 256, True, "fp32")
    print(f"[fp32 train B=2 256^2] logits err {r['err']:.3e}; grads global relL2 {r['grel']:.2e}")
    assert r["err"] <= 1e-4 and r["mask"][0] >= 0.999
    assert abs(r["loss"] - r["ref_loss"]) <= 1e-5 * abs(r["ref_loss"])
    for name, rel, cos in r["rows"]:
        assert rel <= 1e-2 and cos >= 0.9999, (name, rel, cos)


def test_saturation_raises_not_silent():
    """Weights that drive a conv output beyond the fp16 range: the pass is flagged and check_numerics() raises
    (the clamp used to be silent)."""
    import oracle
    sd = oracle.make_state_dict(0)
    sd = {k: (v * 1.0e6 if k == "model.enc2.0.weight" else v) for k, v in sd.items()}
    m = _model("fp16", sd).eval()
    with torch.no_grad():
        m(oracle.make_input(1, 64, 64, 1).cuda())
    with pytest.raises(RuntimeError, match="fp16 range"):
        m.check_numerics()
    m2 = _model("bf16", sd).eval()          # the bf16 mode stores activations in bf16: no flag (its RAW tensors only exist in training)
    with torch.no_grad():
        assert torch.isfinite(m2(oracle.make_input(1, 64, 64, 1).cuda())).all()
    m2.check_numerics()


def test_optimizer_updates_invalidate_packed_filters():
    """ClippedAdamW without ``on_update`` must still make the next forward see the new weights."""
    import oracle
    from enhanced_unet_b200.ops import combined_loss
    from enhanced_unet_b200.optim import ClippedAdamW
    sd = oracle.make_state_dict(0)
    x, t = oracle.make_input(2, 32, 32, 1).cuda(), oracle.make_target(2, 32, 32, 2).cuda()
    m = _model("fp16", sd).train()
    opt = ClippedAdamW(m.parameters(), lr=1e-2)          # no on_update callback
    outs = []
    for _ in range(3):
        opt.zero_grad(set_to_none=True)
        y = m(x)
        outs.append(y.detach().clone())
        combined_loss(y, t).backward()
        opt.step()
    assert nerr(outs[1], outs[0]) > 1e-3 and nerr(outs[2], outs[1]) > 1e-3
    # and the packed copies equal a fresh pack of the current weights: identical logits from a brand-new module
    m.eval()
    fresh = _model("fp16", {k: v.detach().cpu() for k, v in m.state_dict().items()}).eval()
    with torch.no_grad():
        assert torch.equal(m(x), fresh(x))


def test_reference_style_loop_with_torch_loss_and_optimizer():
    """The "two import lines changed" integration path: the drop-in module inside a plain PyTorch training loop - torch ops
    for the loss (the reference's per-sample interpolate + FocalLoss / dice / tversky are torch code), ``clip_grad_norm_`` and
    ``torch.optim.AdamW`` on its parameters (reference train_eval.py:120, 306-343).  Autograd enters the CUDA backward with an
    arbitrary ``dout``; the torch optimiser's in-place updates must reach the packed filters (version counters)."""
    import oracle
    import torch.nn.functional as F
    sd = oracle.make_state_dict(7)
    x, t = oracle.make_input(2, 64, 64, 70), oracle.make_target(2, 64, 64, 71)

    def torch_loss(logits, target):
        z = F.interpolate(logits, size=target.shape[-2:], mode="bilinear", align_corners=False)
        return F.cross_entropy(z, target, weight=torch.tensor([1.0, 20.0, 10.0], device=z.device)) + 0.1 * z.square().mean()

    # reference side: the oracle forward + the same torch loss, first-step gradients
    params = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and "running" not in k else v) for k, v in sd.items()}
    y_ref, _ = oracle.unet_forward(params, x, train=True)
    torch_loss(y_ref, t).backward()
    m = _model("fp32", sd).train()
    opt = torch.optim.AdamW(m.parameters(), lr=1e-3, weight_decay=1e-4)
    losses = []
    for step in range(4):
        opt.zero_grad()
        loss = torch_loss(m(x.cuda()), t.cuda())
        loss.backward()
        if step == 0:
            for name, p in m.named_parameters():
                if PRE_BN_BIAS.match(name):
                    continue
                rf = params[name].grad.double().flatten()
                gr = p.grad.detach().cpu().double().flatten()
                assert float((gr - rf).norm() / rf.norm()) <= 3e-2, name
        torch.nn.utils.clip_grad_norm_(m.parameters(), 1.0)
        opt.step()
        losses.append(float(loss))
    assert all(np.isfinite(losses)) and losses[-1] < losses[0], losses
    # the torch optimiser changed the weights in place: a fresh module with the same state_dict must give the same logits
    m.eval()
    fresh = _model("fp32", {k: v.detach().cpu() for k, v in m.state_dict().items()}).eval()
    with torch.no_grad():
        assert torch.equal(m(x.cuda()), fresh(x.cuda()))
