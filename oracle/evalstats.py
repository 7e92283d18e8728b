"""ORACLE (test infrastructure only) -- CPU restatement of the label / evaluator arithmetic.

Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of
``bench.py`` may import this module.  The product (``segmentation-pipeline_b200/``) never does.

Integer work is numpy and bit-exact by construction; the float32 statistic formulas are evaluated with
torch fp32 scalars exactly as the reference does (counts are cast with ``.float()`` BEFORE the divisions,
``segmentation_pipeline/evaluators/segmentation_evaluator.py:74-90``).

PINNED: ``tests/golden/evaluator.npz`` (made by ``oracle/make_golden.py`` from the reference's own
``SegmentationEvaluator`` / ``LabelMapEvaluator`` imported in-process with a stub ``torchio`` module) is
checked in ``tests/test_oracle_eval.py``.
"""
from __future__ import annotations

from typing import Dict, Sequence

import numpy as np
import torch

STATS = ("target_volume", "prediction_volume", "TP", "FP", "TN", "FN", "dice", "jaccard", "precision", "recall")


def argmax_labels(probs: np.ndarray) -> np.ndarray:
    """``CustomArgMax.apply_transform``, transforms/custom_label_transforms.py:267:
    ``torch.argmax(data, dim=0, keepdim=True)`` -> (1, W, H, D) int64, ties resolved to the lowest index."""
    return np.argmax(probs, axis=0)[None].astype(np.int64)


def confusion_matrix(pred: np.ndarray, target: np.ndarray, num_classes: int) -> np.ndarray:
    """L x L joint histogram ``cm[t, p]`` = #voxels with target t and prediction p.  Not a reference function:
    it is the single-pass statistic from which every per-label count of
    segmentation_evaluator.py:69-77 follows (see ``counts_from_confusion``).  Values outside
    [0, num_classes) are counted in the LAST class ("other"): a voxel whose prediction is an unlisted label still is a
    false negative of its target label, exactly as ``(target == v) & ~(pred == v)`` counts it."""
    p = np.asarray(pred).reshape(-1).astype(np.int64)
    t = np.asarray(target).reshape(-1).astype(np.int64)
    p = np.where((p >= 0) & (p < num_classes), p, num_classes - 1)
    t = np.where((t >= 0) & (t < num_classes), t, num_classes - 1)
    cm = np.bincount(t * num_classes + p, minlength=num_classes * num_classes)
    return cm.reshape(num_classes, num_classes).astype(np.int64)


def label_counts(pred: np.ndarray, target: np.ndarray, label_value: int) -> Dict[str, int]:
    """segmentation_evaluator.py:69-77 literally: boolean masks, four sums."""
    p = np.asarray(pred) == label_value
    t = np.asarray(target) == label_value
    return {"TP": int((t & p).sum()), "FP": int((~t & p).sum()), "TN": int((~t & ~p).sum()),
            "FN": int((t & ~p).sum())}


def counts_from_confusion(cm: np.ndarray, label_value: int, total_voxels: int) -> Dict[str, int]:
    """TP/FP/TN/FN of one label from the joint histogram (rows = target, cols = prediction)."""
    v = label_value
    if 0 <= v < cm.shape[0]:
        tp = int(cm[v, v])
        pred_v = int(cm[:, v].sum())
        targ_v = int(cm[v, :].sum())
    else:
        tp = pred_v = targ_v = 0
    fp = pred_v - tp
    fn = targ_v - tp
    return {"TP": tp, "FP": fp, "TN": total_voxels - tp - fp - fn, "FN": fn}


def stats_from_counts(c: Dict[str, int]) -> Dict[str, float]:
    """segmentation_evaluator.py:79-90 in float32 (0/0 -> nan, x/0 -> inf, as torch)."""
    TP, FP, TN, FN = (torch.tensor(float(c[k]), dtype=torch.float32) for k in ("TP", "FP", "TN", "FN"))
    s = {
        "target_volume": TP + FN,
        "prediction_volume": TP + FP,
        "TP": TP, "FP": FP, "TN": TN, "FN": FN,
        "dice": 2 * TP / (2 * TP + FP + FN),
        "jaccard": TP / (TP + FP + FN),
        "precision": TP / (TP + FP),
        "recall": TP / (TP + FN),
    }
    return {k: v.item() for k, v in s.items()}


def segmentation_stats(pred: np.ndarray, target: np.ndarray, label_values: Dict[str, int],
                       stats: Sequence[str] = STATS) -> Dict[str, Dict[str, float]]:
    """Per-label statistics for one subject, ``{label_name: {stat: value}}``."""
    out = {}
    for name, v in label_values.items():
        full = stats_from_counts(label_counts(pred, target, v))
        out[name] = {k: full[k] for k in stats}
    return out


def label_volumes(label_map: np.ndarray, label_values: Dict[str, int]) -> Dict[str, int]:
    """label_map_evaluator.py:77-81: ``(data == v).sum()`` per label (int64)."""
    return {name: int((np.asarray(label_map) == v).sum()) for name, v in label_values.items()}


# --------------------------------------------------------------------------------------------- instance evaluation
def connected_components(mask: np.ndarray, connectivity: int = 2):
    """``skimage.morphology.label(mask, return_num=True, connectivity=connectivity)`` as used at
    evaluators/instance_segmentation_evaluator.py:107-109.  skimage is not installed here; ``scipy.ndimage.label`` with
    ``generate_binary_structure(3, connectivity)`` defines the same neighbourhood (voxels within ``connectivity``
    orthogonal steps) and numbers components in the same raster-scan order."""
    from scipy import ndimage
    labels, n = ndimage.label(np.asarray(mask) > 0, structure=ndimage.generate_binary_structure(3, connectivity))
    return labels.astype(np.int64), int(n)


def instance_overlap_histogram(pred: np.ndarray, target: np.ndarray, connectivity: int = 2):
    """instance_segmentation_evaluator.py:103-124: components of pred > 0 and target > 0, then the (N + 1, M + 1) table
    of voxel counts per (target component, predicted component) pair (the reference encodes the pair as
    target + prediction * 10**6 and counts with torch.unique)."""
    pc, m = connected_components(pred, connectivity)
    tc, n = connected_components(target, connectivity)
    hist = np.zeros((n + 1, m + 1), np.int64)
    np.add.at(hist, (tc.reshape(-1), pc.reshape(-1)), 1)
    return hist, n, m


def msseg_detection_test(overlap_histogram: torch.Tensor, min_recall=0.1, contribution_threshold=0.65,
                         min_precision=0.3) -> torch.Tensor:
    """instance_segmentation_evaluator.py:10-72, restated statement by statement (float32 table)."""
    n = overlap_histogram.shape[0] - 1
    target_volume = overlap_histogram.sum(dim=1)
    prediction_volume = overlap_histogram.sum(dim=0)
    detected = []
    for i in range(1, n + 1):
        target_tp = overlap_histogram[i, 1:].sum()
        recall = target_tp / target_volume[i]
        if recall < min_recall:
            detected.append(False)
            continue
        predicted_ids = torch.argsort(overlap_histogram[i, 1:], descending=True) + 1
        contribution_total = 0.0
        for j in predicted_ids:
            precision = overlap_histogram[i, j] / prediction_volume[j]
            if precision < min_precision:
                detected.append(False)
                break
            contribution = overlap_histogram[i, j] / target_tp
            contribution_total += contribution
            if contribution_total >= contribution_threshold:
                detected.append(True)
                break
    return torch.tensor(detected)


def instance_stats(pred: np.ndarray, target: np.ndarray, connectivity: int = 2) -> Dict[str, float]:
    """instance_segmentation_evaluator.py:126-160 for one subject."""
    hist, n, m = instance_overlap_histogram(pred, target, connectivity)
    h = torch.from_numpy(hist).to(torch.float32)
    target_detected = msseg_detection_test(h)
    prediction_detected = msseg_detection_test(h.T)
    detection_recall = target_detected.sum() / n
    detection_precision = prediction_detected.sum() / m
    detection_f1 = 2 * (detection_recall * detection_precision) / (detection_recall + detection_precision)
    TP, FP, TN, FN = h[1:, 1:].sum(), h[0, 1:].sum(), h[0, 0].sum(), h[1:, 0].sum()
    s = {'target_components': n, 'predicted_components': m, 'target_detections': target_detected.sum(),
         'predicted_detections': prediction_detected.sum(), 'detection_recall': detection_recall,
         'detection_precision': detection_precision, 'detection_f1': detection_f1, 'target_volume': TP + FN,
         'prediction_volume': TP + FP, 'TP': TP, 'FP': FP, 'TN': TN, 'FN': FN, 'dice': 2 * TP / (2 * TP + FP + FN),
         'jaccard': TP / (TP + FP + FN), 'precision': TP / (TP + FP), 'recall': TP / (TP + FN)}
    return {k: (v.item() if isinstance(v, torch.Tensor) else v) for k, v in s.items()}
