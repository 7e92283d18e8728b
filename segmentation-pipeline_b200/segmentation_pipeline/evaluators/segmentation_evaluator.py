"""SegmentationEvaluator -- same interface and float32 statistics as the reference's
evaluators/segmentation_evaluator.py:56-102, but the voxel work is ONE pass of b200seg_confusion over the two
label maps (an L x L joint histogram) instead of 2 compares + 8 boolean passes per label on the CPU.
TP / FP / TN / FN are exact integers derived from the histogram; they are cast to float32 before the
divisions exactly like the reference's ``.sum().float()`` (:74-77)."""
from __future__ import annotations

from typing import Dict, Sequence

import torch

from .evaluator import Evaluator
from .labeled_tensor import LabeledTensor

MAX_CLASSES = 40      # histogram classes of b200seg_confusion, including the 'other' class


def _device_labels(image, device):
    data = image["data"] if not hasattr(image, "data") else image.data
    return data.to(device).contiguous()


def confusion_counts(pred: torch.Tensor, target: torch.Tensor, label_values: Dict[str, int]):
    """-> {label_name: (TP, FP, TN, FN)} as Python ints, computed on the device."""
    import b200seg
    device = pred.device if pred.is_cuda else torch.device("cuda", torch.cuda.current_device())
    # one extra class: every value that is not below max(label_values) + 1 -- unlisted or negative labels -- lands
    # in it, so FP = column sum - TP and FN = row sum - TP count them exactly like the reference's
    # (~target_label & pred_label).sum() / (target_label & ~pred_label).sum() (segmentation_evaluator.py:75,77)
    num_classes = max(int(v) for v in label_values.values()) + 2
    if min(int(v) for v in label_values.values()) < 0 or num_classes > MAX_CLASSES:
        raise NotImplementedError(f"label values must lie in [0, {MAX_CLASSES - 2}] for the device histogram, got "
                                  f"{sorted(label_values.values())}")
    pred = pred.to(device).contiguous()
    target = target.to(device).contiguous()
    if pred.dtype != target.dtype or pred.dtype not in (torch.uint8, torch.int64):
        pred, target = pred.to(torch.int64), target.to(torch.int64)
    if pred.numel() != target.numel():
        raise RuntimeError("prediction and target label maps differ in size")
    with b200seg.on_device(device):
        cm = torch.zeros((num_classes, num_classes), dtype=torch.int64, device=device)
        b200seg.confusion(pred, target, num_classes, cm)
    cm = cm.cpu()
    return {name: counts_from_cm(cm, int(value)) for name, value in label_values.items()}


def counts_from_cm(cm: torch.Tensor, value: int):
    """(TP, FP, TN, FN) of one label from the joint histogram cm[target][pred] (exact integers)."""
    total = int(cm.sum())
    tp = int(cm[value, value])
    fp = int(cm[:, value].sum()) - tp
    fn = int(cm[value, :].sum()) - tp
    return tp, fp, total - tp - fp - fn, fn


class SegmentationEvaluator(Evaluator):
    """Per-label TP/FP/TN/FN, Dice, Jaccard, precision, recall and volumes between two label maps that carry a
    ``'label_values'`` ``{name: value}`` property.  Output: ``{'subject_stats': DataFrame,
    'summary_stats': LabeledTensor}``."""

    def __init__(
            self,
            prediction_label_map_name: str,
            target_label_map_name: str,
            stats_to_output: Sequence[str] = ('target_volume', 'prediction_volume',
                                              'TP', 'FP', 'TN', 'FN', 'dice', 'precision', 'recall'),
            summary_stats_to_output: Sequence[str] = ('mean', 'std', 'min', 'max'),
    ):
        self.prediction_label_map_name = prediction_label_map_name
        self.target_label_map_name = target_label_map_name
        self.stats_to_output = stats_to_output
        self.summary_stats_to_output = summary_stats_to_output

    def __call__(self, subjects):
        label_values = subjects[0][self.prediction_label_map_name]['label_values']
        label_names = list(label_values.keys())
        subject_names = [subject['name'] for subject in subjects]
        subject_stats = LabeledTensor(dim_names=['subject', 'label', 'stat'],
                                      dim_keys=[subject_names, label_names, self.stats_to_output])
        device = torch.device("cuda", torch.cuda.current_device())
        for subject in subjects:
            pred = _device_labels(subject[self.prediction_label_map_name], device)
            target = _device_labels(subject[self.target_label_map_name], device)
            counts = confusion_counts(pred, target, label_values)
            for label_name in label_names:
                TP, FP, TN, FN = (torch.tensor(float(c), dtype=torch.float32) for c in counts[label_name])
                stats = {
                    'target_volume': TP + FN,
                    'prediction_volume': TP + FP,
                    'TP': TP, 'FP': FP, 'TN': TN, 'FN': FN,
                    'dice': 2 * TP / (2 * TP + FP + FN),
                    'jaccard': TP / (TP + FP + FN),
                    'precision': TP / (TP + FP),
                    'recall': TP / (TP + FN),
                }
                for stat_name in self.stats_to_output:
                    subject_stats[subject['name'], label_name, stat_name] = stats[stat_name].item()
        summary_stats = subject_stats.compute_summary_stats(self.summary_stats_to_output)
        return {'subject_stats': subject_stats.to_dataframe(), 'summary_stats': summary_stats}
