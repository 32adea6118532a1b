"""The oracle (oracle/) against the fixtures generated from the UNMODIFIED reference
(oracle/make_golden.py -> tests/golden/*.npz), and against the live reference when present."""
import os

import numpy as np
import pytest
import torch

import oracle
from oracle import ref_import
from oracle.unet_oracle import make_fusion_state_dict


import re
PRE_BN_BIAS = re.compile(r"^(model\.(enc|dec)[1234]\.(0|3)|enhance\.0)\.bias$")


def _load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name), allow_pickle=False)


@pytest.mark.parametrize("case", ["b2_32x32", "b1_16x24", "b2_64x64"])
def test_model_forward_backward_matches_reference_fixture(golden_dir, case):
    g = _load(golden_dir, f"model_{case}.npz")
    b, h, w, pseed, xseed, tseed = [int(v) for v in g["meta"]]
    sd = oracle.make_state_dict(pseed)
    x = oracle.make_input(b, h, w, xseed)
    t = oracle.make_target(b, h, w, tseed)
    with torch.no_grad():
        y_eval, nb = oracle.unet_forward(sd, x, train=False)
    assert nb == {}
    np.testing.assert_allclose(y_eval.numpy(), g["logits_eval"], rtol=0, atol=2e-5)
    params = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and "running" not in k else v)
              for k, v in sd.items()}
    y, nb = oracle.unet_forward(params, x, train=True)
    np.testing.assert_allclose(y.detach().numpy(), g["logits_train"], rtol=0, atol=2e-5)
    loss = oracle.batch_loss(y, t)
    assert abs(loss.item() - float(g["loss"])) <= 1e-5 * abs(float(g["loss"]))
    loss.backward()
    for k in g.files:
        if k.startswith("buf/"):
            np.testing.assert_allclose(nb[k[4:]].numpy(), g[k], rtol=1e-5, atol=1e-6)
        if k.startswith("gval/"):
            name = k[5:]
            grad = params[name].grad.flatten()[torch.from_numpy(g["gidx/" + name])].numpy()
            if PRE_BN_BIAS.match(name):
                # conv bias feeding a train-mode BN: gradient is exactly 0 in real arithmetic; the
                # reference value is fp32 rounding noise (SURVEY.md §7 hard part 5)
                assert np.abs(grad).max() < 1e-3 and np.abs(g[k]).max() < 1e-3
                continue
            tol = 2e-4 * float(g["gabsmax/" + name]) + 1e-6
            np.testing.assert_allclose(grad, g[k], rtol=0, atol=tol)


def test_loss_matches_reference_fixture(golden_dir):
    g = _load(golden_dir, "loss.npz")
    for name in "abc":
        logits = torch.from_numpy(g[f"{name}/logits"]).requires_grad_(True)
        t = torch.from_numpy(g[f"{name}/target"])
        loss = oracle.batch_loss(logits, t)
        assert abs(loss.item() - float(g[f"{name}/loss"])) <= 2e-6 * abs(float(g[f"{name}/loss"]))
        loss.backward()
        ref = g[f"{name}/grad"]
        np.testing.assert_allclose(logits.grad.numpy(), ref, rtol=0, atol=1e-5 * np.abs(ref).max())


def test_metrics_match_reference_fixture(golden_dir):
    g = _load(golden_dir, "metrics.npz")
    keys = [str(k) for k in g["keys"]]
    names = sorted({k.split("/")[0] for k in g.files if "/" in k})
    assert "kat1" in names and "ignore255" in names
    for name in names:
        pred, gt = g[f"{name}/pred"], g[f"{name}/gt"]
        m = oracle.calculate_semantic_metrics(pred, gt)
        vals = np.array([float(m[k]) for k in keys])
        assert np.array_equal(vals, g[f"{name}/values"]), name          # bit-exact float64
        cm = oracle.confusion_counts(pred[None], gt[None])[0]
        m2 = oracle.metrics_from_counts(cm)
        vals2 = np.array([float(m2[k]) for k in keys])
        assert np.array_equal(vals2, g[f"{name}/values"]), name


def test_kat1_confusion_matrix(golden_dir):
    g = _load(golden_dir, "metrics.npz")
    cm = oracle.confusion_counts(g["kat1/pred"][None], g["kat1/gt"][None])[0]
    assert cm[:3, :3].tolist() == [[450, 430, 480], [436, 490, 461], [440, 471, 438]]
    assert cm[3].sum() == 0 and cm[:, 3].sum() == 0


def test_mask_cascade_matches_reference_fixture(golden_dir):
    g = _load(golden_dir, "mask.npz")
    names = sorted({k.split("/")[0] for k in g.files})
    for name in names:
        got = oracle.convert_probs_to_mask(g[f"{name}/probs"])
        assert np.array_equal(got, g[f"{name}/mask"]), name


def test_fusion_blocks_match_reference_fixture(golden_dir):
    g = _load(golden_dir, "fusion.npz")
    sd = make_fusion_state_dict(0)
    with torch.no_grad():
        y = oracle.fusion_forward(sd, torch.from_numpy(g["main"]), torch.from_numpy(g["aux"]))
    np.testing.assert_allclose(y.numpy(), g["out"], rtol=0, atol=2e-5)


def test_fusion_blocks_training_match_reference_fixture(golden_dir):
    """Training mode of the fusion blocks (batch-statistics BN, the reference's own Dropout2d draw read back through hooks):
    output, input gradients, every parameter gradient and the updated BN buffers."""
    g = _load(golden_dir, "fusion_train.npz")
    sd = make_fusion_state_dict(0)
    params = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and "running" not in k else v) for k, v in sd.items()}
    a = torch.from_numpy(g["main"]).requires_grad_(True)
    b = torch.from_numpy(g["aux"]).requires_grad_(True)
    y, nb = oracle.fusion_forward(params, a, b, train=True, dropout_scales=(torch.from_numpy(g["s1"]), torch.from_numpy(g["s2"])))
    np.testing.assert_allclose(y.detach().numpy(), g["out"], rtol=0, atol=2e-5)
    (y * torch.from_numpy(g["dout"])).sum().backward()
    np.testing.assert_allclose(a.grad.numpy(), g["dmain"], rtol=0, atol=2e-5 * np.abs(g["dmain"]).max())
    np.testing.assert_allclose(b.grad.numpy(), g["daux"], rtol=0, atol=2e-5 * np.abs(g["daux"]).max())
    for k in g.files:
        if k.startswith("grad/"):
            got = params[k[5:]].grad.numpy()
            np.testing.assert_allclose(got, g[k], rtol=0, atol=1e-4 * max(1e-12, np.abs(g[k]).max()), err_msg=k)
        elif k.startswith("buf/"):
            np.testing.assert_allclose(nb[k[4:]].numpy(), g[k], rtol=0, atol=1e-5 * max(1.0, np.abs(g[k]).max()), err_msg=k)


def test_resize_identity_kat3():
    # bilinear 2x-down with align_corners=False == 2x2 mean
    x = torch.randn(2, 3, 16, 24)
    a = torch.nn.functional.interpolate(x, size=(8, 12), mode="bilinear", align_corners=False)
    assert torch.equal(a, torch.nn.functional.avg_pool2d(x, 2))


@pytest.mark.ref
def test_oracle_against_live_reference():
    ref_models, ref_metrics, ref_te = ref_import.load()
    sd = oracle.make_state_dict(9)
    x = oracle.make_input(2, 32, 48, 10)
    m = ref_models.EnhancedUNet(3)
    m.load_state_dict(sd, strict=True)
    assert list(m.state_dict().keys()) == list(sd.keys())
    m.train()
    y_ref = m(x)
    y, nb = oracle.unet_forward(sd, x, train=True)
    assert torch.allclose(y, y_ref, atol=2e-5, rtol=0)
    for k, v in nb.items():
        assert torch.allclose(v.float(), m.state_dict()[k].float(), atol=1e-6), k
    rng = np.random.default_rng(3)
    pred, gt = rng.integers(0, 3, (37, 41)), rng.integers(0, 3, (37, 41))
    a = ref_metrics.calculate_semantic_metrics(pred, gt)
    b = oracle.calculate_semantic_metrics(pred, gt)
    assert all(float(a[k]) == float(b[k]) for k in a)


def test_instance_metrics_oracle_matches_reference_fixture(golden_dir):
    """metrics.py:61-194: greedy score-ordered instance matching; values are float64-identical to the reference run."""
    g = _load(golden_dir, "instances.npz")
    for name, spec in oracle.INSTANCE_CASES.items():
        m = oracle.calculate_instance_metrics(*oracle.make_instance_case(*spec))
        keys = sorted(m.keys())
        assert keys == [str(k) for k in g[f"{name}/keys"]], name
        assert np.array_equal(np.array([float(m[k]) for k in keys]), g[f"{name}/values"]), name
    if ref_import.available():
        _, ref_metrics, _ = ref_import.load()
        args = oracle.make_instance_case(11, 40, 40, 7, 9)
        a, b = oracle.calculate_instance_metrics(*args), ref_metrics.calculate_instance_metrics(*args)
        assert sorted(a) == sorted(b) and all(float(a[k]) == float(b[k]) for k in a)
