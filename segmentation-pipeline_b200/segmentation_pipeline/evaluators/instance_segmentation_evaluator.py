"""InstanceSegmentationEvaluator -- same constructor, statistics and output as the reference's
evaluators/instance_segmentation_evaluator.py:75-175 (the lesion-detection metric of the MSSEG / MSSEG-2 challenges),
with the voxel work on the device: connected components of ``prediction > 0`` and ``target > 0``
(``b200seg_ccl3d_*``, numbered like ``skimage.morphology.label``) and the (N + 1) x (M + 1) overlap table
(``b200seg_overlap_histogram``) that the reference builds on the CPU with skimage + ``torch.unique``.  The detection test
itself runs on the small table on the host, restated from :10-72.  No skimage import: the reference module needs it at
import time, this one does not."""
from __future__ import annotations

from typing import Any, Callable, Dict, Sequence

import torch

from .evaluator import Evaluator
from .labeled_tensor import LabeledTensor


def msseg_detection_test(overlap_histogram, min_recall=0.1, contribution_threshold=0.65, min_precision=0.3):
    """Detection test of "Objective Evaluation of Multiple Sclerosis Lesion Segmentation using a Data Management and
    Processing Infrastructure" (MSSEG 2016 / MSSEG-2 2021), reference instance_segmentation_evaluator.py:10-72.

    ``overlap_histogram``: (N + 1, M + 1), element [i, j] = voxels shared by target component i and predicted component
    j (0 = background).  Returns a boolean tensor of length N: target instance i is detected when it is covered by at
    least ``min_recall`` and the predicted instances that make up the first ``contribution_threshold`` of its overlap
    each have precision >= ``min_precision``."""
    n = overlap_histogram.shape[0] - 1
    target_volume = overlap_histogram.sum(dim=1)
    prediction_volume = overlap_histogram.sum(dim=0)
    detected = []
    for i in range(1, n + 1):
        target_tp = overlap_histogram[i, 1:].sum()
        if target_tp / target_volume[i] < min_recall:
            detected.append(False)
            continue
        order = torch.argsort(overlap_histogram[i, 1:], descending=True) + 1
        contribution_total = 0.0
        for j in order:
            if overlap_histogram[i, j] / prediction_volume[j] < min_precision:
                detected.append(False)
                break
            contribution_total += overlap_histogram[i, j] / target_tp
            if contribution_total >= contribution_threshold:
                detected.append(True)
                break
    return torch.tensor(detected)


def instance_overlap(pred_labels: torch.Tensor, target_labels: torch.Tensor, connectivity: int = 2):
    """(1, W, H, D) or (W, H, D) label maps -> (float32 (N + 1, M + 1) overlap table on the CPU, N, M), computed on
    the device."""
    import b200seg
    device = pred_labels.device if pred_labels.is_cuda else torch.device("cuda", torch.cuda.current_device())
    with b200seg.on_device(device):
        def mask(t):
            t = t.reshape(t.shape[-3:]) if t.dim() == 4 else t
            t = t.to(device)
            if t.dtype not in (torch.uint8, torch.int32, torch.int64):
                t = (t > 0).to(torch.uint8)
            return t.contiguous()
        pred_c, m = b200seg.connected_components(mask(pred_labels), connectivity)
        targ_c, n = b200seg.connected_components(mask(target_labels), connectivity)
        hist = b200seg.overlap_histogram(targ_c, pred_c, n, m)
    return hist.cpu().to(torch.float32), n, m


class InstanceSegmentationEvaluator(Evaluator):
    def __init__(
            self,
            prediction_label_map_name: str,
            target_label_map_name: str,
            stats_to_output: Sequence[str] = ('target_components', 'predicted_components',
                                              'target_detections', 'predicted_detections',
                                              'detection_recall', 'detection_precision', 'detection_f1',
                                              'target_volume', 'prediction_volume', 'TP', 'FP', 'TN', 'FN',
                                              'dice', 'jaccard', 'precision', 'recall',),
            summary_stats_to_output: Sequence[str] = ('mean', 'std', 'min', 'max', 'median', 'mode'),
            connectivity: int = 2,
            detection_test: Callable = msseg_detection_test,
            detection_test_params: Dict[str, Any] = None,
    ):
        self.prediction_label_map_name = prediction_label_map_name
        self.target_label_map_name = target_label_map_name
        self.stats_to_output = stats_to_output
        self.summary_stats_to_output = summary_stats_to_output
        self.connectivity = connectivity
        self.detection_test = detection_test
        self.detection_test_params = {} if detection_test_params is None else detection_test_params

    def __call__(self, subjects):
        subject_names = [subject['name'] for subject in subjects]
        subject_stats = LabeledTensor(dim_names=['subject', 'stat'],
                                      dim_keys=[subject_names, self.stats_to_output])
        for subject in subjects:
            pred = subject[self.prediction_label_map_name]
            targ = subject[self.target_label_map_name]
            pred_data = pred["data"] if not hasattr(pred, "data") else pred.data
            targ_data = targ["data"] if not hasattr(targ, "data") else targ.data
            overlap_histogram, N, M = instance_overlap(pred_data, targ_data, self.connectivity)
            target_detected = self.detection_test(overlap_histogram, **self.detection_test_params)
            prediction_detected = self.detection_test(overlap_histogram.T, **self.detection_test_params)
            detection_recall = target_detected.sum() / N
            detection_precision = prediction_detected.sum() / M
            detection_f1 = 2 * (detection_recall * detection_precision) / (detection_recall + detection_precision)
            TP = overlap_histogram[1:, 1:].sum()
            FP = overlap_histogram[0, 1:].sum()
            TN = overlap_histogram[0, 0].sum()
            FN = overlap_histogram[1:, 0].sum()
            stats = {
                'target_components': N,
                'predicted_components': M,
                'target_detections': target_detected.sum(),
                'predicted_detections': prediction_detected.sum(),
                'detection_recall': detection_recall,
                'detection_precision': detection_precision,
                'detection_f1': detection_f1,
                'target_volume': TP + FN,
                'prediction_volume': TP + FP,
                'TP': TP, 'FP': FP, 'TN': TN, 'FN': FN,
                'dice': 2 * TP / (2 * TP + FP + FN),
                'jaccard': TP / (TP + FP + FN),
                'precision': TP / (TP + FP),
                'recall': TP / (TP + FN),
            }
            for stat_name in self.stats_to_output:
                value = stats[stat_name]
                if isinstance(value, torch.Tensor):
                    value = value.item()
                subject_stats[subject['name'], stat_name] = value
        summary_stats = subject_stats.compute_summary_stats(self.summary_stats_to_output)
        return {'subject_stats': subject_stats.to_dataframe(), 'summary_stats': summary_stats}
