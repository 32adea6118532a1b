"""Drop-in for the hot-path functions of the reference ``metrics.py`` (lines 12-58): same names, argument
meaning, return keys and edge rules; the counting runs in the integer confusion kernel
(``eunet_confusion4x4``), ratios are formed in float64 on the host exactly as numpy does in the
reference (``np.int64 / np.int64``), so every returned value is bit-identical.

Inputs may be numpy arrays (copied to the GPU) or CUDA tensors (used in place).  No CPU fallback.
"""
from __future__ import annotations

from typing import Dict, List, Sequence, Union

import numpy as np
import torch

from .ops import confusion_counts, pair_intersections

ArrayLike = Union[np.ndarray, torch.Tensor]
CLASS_NAMES = ("background", "live", "dead")


def _to_cuda_int(a: ArrayLike) -> torch.Tensor:
    if isinstance(a, np.ndarray):
        if a.dtype == np.bool_:
            a = a.astype(np.uint8)
        if a.dtype not in (np.uint8, np.int32, np.int64):
            a = a.astype(np.int64)
        t = torch.from_numpy(np.ascontiguousarray(a))
    else:
        t = a
        if t.dtype == torch.bool:
            t = t.to(torch.uint8)
        if t.dtype not in (torch.uint8, torch.int32, torch.int64):
            t = t.long()
    if not torch.cuda.is_available():
        raise RuntimeError("enhanced_unet_b200.metrics needs a CUDA device (no CPU fallback)")
    return t.cuda(non_blocking=True)


def _pair(a: ArrayLike, b: ArrayLike):
    ta, tb = _to_cuda_int(a), _to_cuda_int(b)
    if ta.dtype != tb.dtype:
        ta, tb = ta.long(), tb.long()
    if ta.shape != tb.shape:
        raise ValueError(f"mask shapes differ: {tuple(ta.shape)} vs {tuple(tb.shape)}")
    return ta, tb


def _binary_counts(mask1: ArrayLike, mask2: ArrayLike) -> np.ndarray:
    """Membership masks (non-zero = member, what every reference call site passes: metrics.py:39-40, 98, 154)."""
    t1, t2 = _pair(mask1, mask2)
    t1 = (t1 != 0).to(torch.uint8)
    t2 = (t2 != 0).to(torch.uint8)
    return confusion_counts(t2.reshape(1, -1), t1.reshape(1, -1))[0].cpu().numpy()   # CM[m1, m2]


def calculate_iou(mask1: ArrayLike, mask2: ArrayLike):
    """Reference metrics.py:12-18."""
    cm = _binary_counts(mask1, mask2)
    intersection = cm[1, 1]
    union = cm[1, 1] + cm[1, 0] + cm[0, 1]
    if union == 0:
        return 1.0 if intersection == 0 else 0.0
    return intersection / union


def calculate_dice(mask1: ArrayLike, mask2: ArrayLike):
    """Reference metrics.py:21-26."""
    cm = _binary_counts(mask1, mask2)
    intersection = cm[1, 1]
    s = (cm[1, 0] + cm[1, 1]) + (cm[0, 1] + cm[1, 1])
    if s == 0:
        return 1.0
    return 2 * intersection / s


def metrics_from_counts(cm: np.ndarray) -> Dict:
    """The 9 keys of reference metrics.py:45-56 from one image's 4x4 count matrix CM[gt, pred]."""
    cm = np.asarray(cm, dtype=np.int64)
    m: Dict = {}
    for c, name in enumerate(CLASS_NAMES):
        inter = cm[c, c]
        g = cm[c, :].sum()
        p = cm[:, c].sum()
        union = g + p - inter
        m[f"sem_{name}_iou"] = 1.0 if union == 0 else inter / union
        m[f"sem_{name}_dice"] = 1.0 if (g + p) == 0 else 2 * inter / (g + p)
    mean_iou = (m["sem_background_iou"] + m["sem_live_iou"] + m["sem_dead_iou"]) / 3
    mean_iou_cells = (m["sem_live_iou"] + m["sem_dead_iou"]) / 2
    mean_dice = (m["sem_live_dice"] + m["sem_dead_dice"]) / 2
    m["sem_mean_iou"] = mean_iou_cells
    m["sem_mean_iou_all"] = mean_iou
    m["sem_mean_dice"] = mean_dice
    return m


def calculate_semantic_metrics(pred_mask: ArrayLike, gt_mask: ArrayLike) -> Dict:
    """Reference metrics.py:29-58 for one image (masks of any shape, values 0/1/2; other values such as the
    ignore label 255 match no class, as in the reference)."""
    p, g = _pair(pred_mask, gt_mask)
    cm = confusion_counts(p.reshape(1, -1), g.reshape(1, -1))[0].cpu().numpy()
    return metrics_from_counts(cm)


def batch_semantic_metrics(pred_masks: ArrayLike, gt_masks: ArrayLike) -> List[Dict]:
    """Per-image metrics for a batch [N,H,W] in ONE kernel launch and ONE device->host copy of the 16*N
    counts (the reference loops over images on the host, train_eval.py:887-905)."""
    p, g = _pair(pred_masks, gt_masks)
    cm = confusion_counts(p.reshape(p.shape[0], -1), g.reshape(g.shape[0], -1)).cpu().numpy()
    return [metrics_from_counts(cm[i]) for i in range(cm.shape[0])]


def confusion_matrix_3x3(pred_masks: ArrayLike, gt_masks: ArrayLike, ignore_label: int = 255) -> np.ndarray:
    """3x3 int64 confusion matrix over a set of images, rows = gt, cols = pred: the counts of reference
    visualization.py:294-311 (drop gt == ignore_label, clip both to [0,2], sklearn confusion_matrix).
    Values other than ``ignore_label`` outside [0,2] are clipped like ``np.clip`` does there."""
    p, g = _pair(pred_masks, gt_masks)
    p = p.reshape(1, -1)
    g = g.reshape(1, -1)
    keep = g != ignore_label
    # np.clip(., 0, 2) of the reference, applied on device (plumbing, not the reduction)
    p = torch.where(keep, p.clamp(0, 2), torch.full_like(p, 3))
    g = torch.where(keep, g.clamp(0, 2), torch.full_like(g, 3))
    cm = confusion_counts(p, g)[0].cpu().numpy()
    return cm[:3, :3].copy()


def _stack_masks(masks: Sequence[ArrayLike]) -> torch.Tensor:
    if len(masks) == 0:
        return torch.zeros(0, 0, dtype=torch.uint8, device="cuda")
    ts = []
    for m in masks:
        t = torch.from_numpy(np.ascontiguousarray(m)) if isinstance(m, np.ndarray) else m
        ts.append((t != 0).to(torch.uint8).reshape(-1))
    return torch.stack(ts).cuda(non_blocking=True)


def pairwise_iou(pred_masks: Sequence[ArrayLike], gt_masks: Sequence[ArrayLike]) -> np.ndarray:
    """``calculate_iou(p, g)`` (metrics.py:12-18) for every pair, as a float64 [P,G] matrix: intersections from ONE
    bit-plane AND/popcount pass on the GPU, unions from the areas, ratios formed as numpy int64 / int64."""
    P, G = len(pred_masks), len(gt_masks)
    if P == 0 or G == 0:
        return np.zeros((P, G), dtype=np.float64)
    a, b = _stack_masks(pred_masks), _stack_masks(gt_masks)
    if a.shape[1] != b.shape[1]:
        raise ValueError("prediction and ground-truth masks differ in size")
    inter, area_a, area_b = pair_intersections(a, b)
    inter, area_a, area_b = inter.cpu().numpy(), area_a.cpu().numpy(), area_b.cpu().numpy()
    union = area_a[:, None] + area_b[None, :] - inter
    with np.errstate(divide="ignore", invalid="ignore"):
        iou = inter / union
    empty = union == 0
    iou[empty] = np.where(inter[empty] == 0, 1.0, 0.0)
    return iou


def _match_class(ious: np.ndarray, scores: Sequence[float], n_gt: int, iou_threshold: float):
    """The greedy matching loop of metrics.py:88-107 (identical for live and dead) over a precomputed IoU matrix."""
    matched_ious: List = []
    all_pred_ious: List = []
    matched_gt = set()
    order = sorted(range(len(scores)), key=lambda i: scores[i], reverse=True)      # stable, like sorted(pred, key=score)
    for pi in order:
        best_iou, best_gt = 0.0, -1
        for gi in range(n_gt):
            if gi in matched_gt:
                continue
            iou = ious[pi, gi]
            if iou > best_iou:
                best_iou, best_gt = iou, gi
        all_pred_ious.append(best_iou)
        if best_iou >= iou_threshold and best_gt >= 0:
            matched_ious.append(best_iou)
            matched_gt.add(best_gt)
    return matched_ious, all_pred_ious


def calculate_instance_metrics(pred_masks: List[np.ndarray], pred_labels: List[int], pred_scores: List[float],
                               gt_masks: List[np.ndarray], gt_labels: List[int], iou_threshold: float = 0.05) -> Dict:
    """Reference metrics.py:61-194: per class (label 0 = live, 1 = dead) greedy score-ordered matching of predicted
    to ground-truth instance masks; same keys, edge rules and float64 values.  The O(P*G) full-image IoUs of the
    reference (95-101, 151-157) are replaced by ONE pairwise intersection pass on the GPU per class."""
    metrics: Dict = {"live_iou": 0.0, "live_precision": 0.0, "live_recall": 0.0, "live_ap": 0.0,
                     "dead_iou": 0.0, "dead_precision": 0.0, "dead_recall": 0.0, "dead_ap": 0.0}
    for label, name in ((0, "live"), (1, "dead")):
        pred = [(m, s) for m, l, s in zip(pred_masks, pred_labels, pred_scores) if l == label]
        gt = [m for m, l in zip(gt_masks, gt_labels) if l == label]
        if len(gt) == 0:
            continue
        ious = pairwise_iou([m for m, _ in pred], gt)
        matched, all_ious = _match_class(ious, [s for _, s in pred], len(gt), iou_threshold)
        if matched:
            metrics[f"{name}_iou"] = np.mean(matched)
        elif all_ious:
            metrics[f"{name}_iou"] = np.mean(all_ious)
        else:
            metrics[f"{name}_iou"] = 0.0
        metrics[f"{name}_precision"] = len(matched) / len(pred) if pred else 0.0
        metrics[f"{name}_recall"] = len(matched) / len(gt) if gt else 0.0
        if metrics[f"{name}_precision"] == 0.0 and metrics[f"{name}_iou"] > 0.0 and pred:
            avg = np.mean(all_ious) if all_ious else 0.0
            if not avg < 0.1:
                metrics[f"{name}_avg_iou_below_threshold"] = avg
        if pred:
            metrics[f"{name}_ap"] = metrics[f"{name}_precision"] * metrics[f"{name}_recall"]
    return metrics
