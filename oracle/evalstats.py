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
