"""CPU-only checks of the host-side logic around the kernels (no compute calls): gradient production order / flat buffer
layout, packed-filter bookkeeping, the instance-matching loop of the metrics drop-in (with the GPU pairwise-IoU step
replaced by a numpy stand-in), bench.py's FLOP accounting, and the loud failures without a GPU."""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def test_grad_production_order_covers_every_parameter_once():
    from enhanced_unet_b200 import engine
    from enhanced_unet_b200.models import EnhancedUNet
    from enhanced_unet_b200.parallel import FlatGradBuffer
    m = EnhancedUNet(3)
    names = [n for n, _ in m.named_parameters()]
    order = engine.grad_production_order()
    assert sorted(order) == sorted(names) and len(set(order)) == len(order) == 64
    buf = FlatGradBuffer.for_model(m, bucket_bytes=8 << 20)
    assert buf.numel == sum(p.numel() for p in m.parameters()) == 7_790_790
    assert buf.buckets[0][0] == 0 and buf.buckets[-1][1] == buf.numel
    assert all(a[1] == b[0] for a, b in zip(buf.buckets, buf.buckets[1:]))
    assert all((hi - lo) * 4 <= (8 << 20) or sum(1 for n in buf.names if lo <= buf.offsets[n] < hi) == 1 for lo, hi in buf.buckets)
    for n, p in m.named_parameters():
        assert buf.grads[n].shape == p.shape and buf.grads[n].is_contiguous()
    with pytest.raises(RuntimeError):
        buf.dst("model.enc1.0.weight", (1, 2, 3))
    # buckets close at the ends of encoder levels 3 and 2: only enc1's gradients (154 KB) are exchanged after the last kernel
    ends = [[n for n in buf.names if lo <= buf.offsets[n] < hi][-1] for lo, hi in buf.buckets]
    assert "model.enc3.0.bias" in ends and "model.enc2.0.bias" in ends and ends[-1] == "model.enc1.0.bias"
    lo, hi = buf.buckets[-1]
    assert all(n.startswith("model.enc1.") for n in buf.names if lo <= buf.offsets[n] < hi) and (hi - lo) * 4 < 160 * 1024


def test_pack_specs_and_wgrad_workspace():
    from enhanced_unet_b200 import engine
    cx = engine._Ctx(torch.device("cpu"), torch.float16)
    train, infer = engine.pack_specs(cx, True), engine.pack_specs(cx, False)
    assert len(infer) == 15 and len(train) == 29                                      # no dgrad operand for enc1.0
    assert ("model.enc1.0", False, True) in train                                     # hi/lo split of the first layer in the 16-bit modes
    assert ("model.enc1.0", True, False) not in train and ("enhance.0", True, False) in train
    need = 64 * 9 * 16 + sum(co * 9 * ((ci + 15) // 16 * 16) + co * 9 * co for _, ci, co in engine.BLOCKS)
    assert engine.wgrad_workspace_numel() >= need


def test_instance_matching_host_loop_matches_oracle(monkeypatch):
    import oracle
    from enhanced_unet_b200 import metrics

    def cpu_pairwise_iou(pred_masks, gt_masks):       # stand-in for the GPU bit-plane kernels
        return np.array([[float(oracle.calculate_iou(a, b)) for b in gt_masks] for a in pred_masks]).reshape(len(pred_masks), len(gt_masks))

    monkeypatch.setattr(metrics, "pairwise_iou", cpu_pairwise_iou)
    for name, spec in oracle.INSTANCE_CASES.items():
        args = oracle.make_instance_case(*spec)
        got, want = metrics.calculate_instance_metrics(*args), oracle.calculate_instance_metrics(*args)
        assert sorted(got) == sorted(want), name
        assert all(float(got[k]) == float(want[k]) for k in want), name
    args = oracle.make_instance_case(77, 40, 44, 10, 16)
    got, want = metrics.calculate_instance_metrics(*args, iou_threshold=0.5), oracle.calculate_instance_metrics(*args, iou_threshold=0.5)
    assert all(float(got[k]) == float(want[k]) for k in want)


def test_bench_flop_accounting_matches_survey():
    import bench
    assert abs(bench.conv_flops_train(16, 512) / 1e9 - 16476.6) < 0.5                 # 16,476.6 GFLOP per step
    hb = bench.hbm_bytes_per_step(16, 512)
    assert hb["eunet_tail_out_fwd"] == 16 * 1024 * 1024 * (128 + 16 + 12)


def test_no_gpu_means_loud_failure():
    from enhanced_unet_b200.data import HostBatchPrefetcher
    from enhanced_unet_b200.models import EnhancedUNet
    with pytest.raises(RuntimeError):
        HostBatchPrefetcher("cpu")
    m = EnhancedUNet(3)
    with pytest.raises(RuntimeError):
        m(torch.rand(1, 3, 32, 32))                                                    # CPU parameters / tensors: no fallback
    if not torch.cuda.is_available():
        from enhanced_unet_b200 import metrics
        with pytest.raises(RuntimeError):
            metrics.calculate_semantic_metrics(np.zeros((4, 4), np.int64), np.zeros((4, 4), np.int64))


def test_prepare_image_tensor_matches_reference_fixture(golden_dir):
    """Evaluator._prepare_image_tensor (host-side cv2 CLAHE + sharpening, reference train_eval.py:365-395) against the output of
    the reference's own method (tests/golden/preprocess.npz, oracle/make_golden.py:preprocess_case): bit-identical."""
    pytest.importorskip("cv2")
    from enhanced_unet_b200.train_eval import Evaluator
    g = np.load(os.path.join(golden_dir, "preprocess.npz"))
    ev = Evaluator(model=None, device="cpu", model_name="enhanced_unet")
    assert ev.enable_tta                                    # reference train_eval.py:363
    for name in ("96x80", "64x64"):
        got = ev._prepare_image_tensor(torch.from_numpy(g[f"{name}/image"])).numpy()
        assert got.shape == g[f"{name}/prepared"].shape and np.array_equal(got, g[f"{name}/prepared"]), name
