"""CPU restatement (numpy) of the Dice / IoU / confusion reductions.  TEST INFRASTRUCTURE ONLY.

Follows /root/reference/metrics.py:12-58 (``calculate_iou``, ``calculate_dice``,
``calculate_semantic_metrics``), the confusion-matrix producers of visualization.py:294-311 /
1484-1492, and (the "next" row) ``Evaluator._convert_probs_to_mask`` train_eval.py:455-568.
"""
from __future__ import annotations

from typing import Dict

import numpy as np

CLASS_NAMES = ("background", "live", "dead")


def calculate_iou(mask1: np.ndarray, mask2: np.ndarray):
    """metrics.py:12-18.  Empty union -> 1.0 (python float); else numpy float64 ratio."""
    inter = int(np.logical_and(mask1, mask2).sum())
    union = int(np.logical_or(mask1, mask2).sum())
    if union == 0:
        return 1.0 if inter == 0 else 0.0
    return np.int64(inter) / np.int64(union)


def calculate_dice(mask1: np.ndarray, mask2: np.ndarray):
    """metrics.py:21-26."""
    inter = int(np.logical_and(mask1, mask2).sum())
    s = int(mask1.sum()) + int(mask2.sum())
    if s == 0:
        return 1.0
    return 2 * np.int64(inter) / np.int64(s)


def calculate_semantic_metrics(pred_mask: np.ndarray, gt_mask: np.ndarray) -> Dict:
    """metrics.py:29-58, literally: per class binary masks -> IoU / Dice; three means."""
    m: Dict = {}
    for c, name in enumerate(CLASS_NAMES):
        p = (pred_mask == c).astype(np.uint8)
        g = (gt_mask == c).astype(np.uint8)
        m[f"sem_{name}_iou"] = calculate_iou(p, g)
        m[f"sem_{name}_dice"] = calculate_dice(p, g)
    mean_iou = (m["sem_background_iou"] + m["sem_live_iou"] + m["sem_dead_iou"]) / 3
    m["sem_mean_iou"] = (m["sem_live_iou"] + m["sem_dead_iou"]) / 2
    m["sem_mean_iou_all"] = mean_iou
    m["sem_mean_dice"] = (m["sem_live_dice"] + m["sem_dead_dice"]) / 2
    return m


def confusion_counts(pred: np.ndarray, gt: np.ndarray) -> np.ndarray:
    """Per-image 4x4 int64 count matrix ``CM[g, p]`` with class index 3 = "any other value"
    (e.g. the ignore label 255 of visualization.py:299-303).  ``pred``/``gt``: [N, ...] integer
    arrays; returns [N,4,4].  The 3x3 top-left block is sklearn's
    ``confusion_matrix(labels=[0,1,2])`` (visualization.py:306-311) for that image."""
    pred = np.asarray(pred)
    gt = np.asarray(gt)
    n = pred.shape[0]
    p = pred.reshape(n, -1).astype(np.int64)
    g = gt.reshape(n, -1).astype(np.int64)
    p = np.where((p >= 0) & (p <= 2), p, 3)
    g = np.where((g >= 0) & (g <= 2), g, 3)
    out = np.zeros((n, 4, 4), dtype=np.int64)
    for i in range(n):
        out[i] = np.bincount(g[i] * 4 + p[i], minlength=16).reshape(4, 4)
    return out


def metrics_from_counts(cm: np.ndarray) -> Dict:
    """The 9 keys of metrics.py:45-56 from one image's 4x4 count matrix:
    I = CM[c,c], sum(gt==c) = row c, sum(pred==c) = column c, U = row + col - I."""
    cm = np.asarray(cm, dtype=np.int64)
    m: Dict = {}
    for c, name in enumerate(CLASS_NAMES):
        inter = cm[c, c]
        g = cm[c, :].sum()
        p = cm[:, c].sum()
        union = g + p - inter
        m[f"sem_{name}_iou"] = (1.0 if union == 0 else inter / union)
        m[f"sem_{name}_dice"] = (1.0 if (g + p) == 0 else 2 * inter / (g + p))
    mean_iou = (m["sem_background_iou"] + m["sem_live_iou"] + m["sem_dead_iou"]) / 3
    m["sem_mean_iou"] = (m["sem_live_iou"] + m["sem_dead_iou"]) / 2
    m["sem_mean_iou_all"] = mean_iou
    m["sem_mean_dice"] = (m["sem_live_dice"] + m["sem_dead_dice"]) / 2
    return m


def convert_probs_to_mask(probs: np.ndarray) -> np.ndarray:
    """train_eval.py:455-568 on a [3,H,W] float32 probability map, without padding/crop.
    Order-dependent rule cascade, then the two global pixel-ratio filters.  float32 arithmetic to
    match the tensor ops of the reference (thresholds are python floats promoted to float32)."""
    probs = np.asarray(probs, dtype=np.float32)
    bg, live, dead = probs[0], probs[1], probs[2]
    f = np.float32
    pred = np.argmax(probs, axis=0).astype(np.int64)
    maxp = probs.max(axis=0)
    low = (pred == 1) & ((live < f(0.42)) | (live <= bg * f(1.15)))
    pred[low] = 0
    low = (pred == 2) & ((dead < f(0.5)) | (dead <= bg * f(1.3)) | (bg > f(0.3)) | (live > dead * f(0.9)))
    pred[low] = 0
    hi_live = (pred == 0) & (live > f(0.42)) & (live > bg * f(1.15)) & (live > dead * f(1.05))
    pred[hi_live] = 1
    hi_dead = (pred == 0) & (dead > f(0.5)) & (dead > bg * f(1.3)) & (dead > live * f(1.1)) & (bg < f(0.3)) & (~hi_live)
    pred[hi_dead] = 2
    sw = (pred == 1) & (dead > live * f(1.15)) & (dead > f(0.45))
    pred[sw] = 2
    sw = (pred == 2) & (live > dead * f(1.15)) & (live > f(0.42))
    pred[sw] = 1
    pred[maxp < f(0.3)] = 0
    h, w = pred.shape
    live_ratio = (pred == 1).sum() / (h * w)
    dead_ratio = (pred == 2).sum() / (h * w)
    if live_ratio > 0.5:
        keep = (live > f(0.5)) & (live > bg * f(1.3)) & (bg < f(0.3))
        pred[(pred == 1) & (~keep)] = 0
    if dead_ratio > 0.15:
        dm = pred == 2
        if dead_ratio > 0.4:
            keep = (dead > f(0.65)) & (dead > bg * f(1.6)) & (bg < f(0.2)) & (live < dead * f(0.7))
        elif dead_ratio > 0.25:
            keep = (dead > f(0.6)) & (dead > bg * f(1.5)) & (bg < f(0.25)) & (live < dead * f(0.8))
        else:
            keep = (dead > f(0.55)) & (dead > bg * f(1.4)) & (bg < f(0.25))
        pred[dm & (~keep)] = 0
    return pred


def calculate_instance_metrics(pred_masks, pred_labels, pred_scores, gt_masks, gt_labels, iou_threshold: float = 0.05) -> Dict:
    """metrics.py:61-194, restated literally (per class: score-ordered greedy matching on calculate_iou, mean IoU of the
    matches - or of all best IoUs when nothing matched -, precision, recall, precision*recall as "AP")."""
    metrics: Dict = {"live_iou": 0.0, "live_precision": 0.0, "live_recall": 0.0, "live_ap": 0.0,
                     "dead_iou": 0.0, "dead_precision": 0.0, "dead_recall": 0.0, "dead_ap": 0.0}
    for label, name in ((0, "live"), (1, "dead")):
        pred = [(m, s) for m, l, s in zip(pred_masks, pred_labels, pred_scores) if l == label]
        gt = [m for m, l in zip(gt_masks, gt_labels) if l == label]
        if len(gt) == 0:
            continue
        ious, all_ious, matched_gt = [], [], set()
        for pm, _score in sorted(pred, key=lambda x: x[1], reverse=True):
            best_iou, best_idx = 0.0, -1
            for i, gm in enumerate(gt):
                if i in matched_gt:
                    continue
                iou = calculate_iou(pm, gm)
                if iou > best_iou:
                    best_iou, best_idx = iou, i
            all_ious.append(best_iou)
            if best_iou >= iou_threshold and best_idx >= 0:
                ious.append(best_iou)
                matched_gt.add(best_idx)
        metrics[f"{name}_iou"] = np.mean(ious) if ious else (np.mean(all_ious) if all_ious else 0.0)
        metrics[f"{name}_precision"] = len(ious) / len(pred) if pred else 0.0
        metrics[f"{name}_recall"] = len(ious) / len(gt) if gt else 0.0
        if metrics[f"{name}_precision"] == 0.0 and metrics[f"{name}_iou"] > 0.0 and pred:
            avg = np.mean(all_ious) if all_ious else 0.0
            if not avg < 0.1:
                metrics[f"{name}_avg_iou_below_threshold"] = avg
        if pred:
            metrics[f"{name}_ap"] = metrics[f"{name}_precision"] * metrics[f"{name}_recall"]
    return metrics


def make_instance_case(seed: int, h: int, w: int, n_gt: int, n_pred: int):
    """Seeded synthetic instance set: random discs as ground truth, jittered / spurious discs as predictions."""
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:h, 0:w]

    def disc(cy, cx, r):
        return ((yy - cy) ** 2 + (xx - cx) ** 2 <= r * r).astype(np.uint8)

    gts = [(rng.uniform(0, h), rng.uniform(0, w), rng.uniform(2, 7), int(rng.integers(0, 2))) for _ in range(n_gt)]
    gt_masks = [disc(cy, cx, r) for cy, cx, r, _ in gts]
    gt_labels = [l for *_, l in gts]
    pred_masks, pred_labels, pred_scores = [], [], []
    for _ in range(n_pred):
        if gts and rng.random() < 0.7:
            cy, cx, r, l = gts[int(rng.integers(0, len(gts)))]
            cy, cx, r = cy + rng.normal(0, 1.5), cx + rng.normal(0, 1.5), max(1.0, r + rng.normal(0, 1.0))
            if rng.random() < 0.15:
                l = 1 - l
        else:
            cy, cx, r, l = rng.uniform(0, h), rng.uniform(0, w), rng.uniform(1, 6), int(rng.integers(0, 2))
        pred_masks.append(disc(cy, cx, r))
        pred_labels.append(int(l))
        pred_scores.append(float(np.round(rng.uniform(0.2, 1.0), 2)))      # rounded: ties exercise the stable sort
    return pred_masks, pred_labels, pred_scores, gt_masks, gt_labels


INSTANCE_CASES = {"mixed": (5, 48, 56, 9, 14), "no_pred": (6, 32, 32, 4, 0), "no_gt": (7, 32, 32, 0, 5),
                  "dense": (8, 64, 64, 25, 40), "tiny": (9, 8, 8, 2, 3)}
