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


def _instance_detected(overlaps, own_volume, other_volumes, min_recall, contribution_threshold, min_precision) -> bool:
    """One row of the detection test: ``overlaps[j]`` = voxels this instance shares with instance j + 1 of the other
    map, ``own_volume`` its voxel count, ``other_volumes[j]`` the voxel count of that other instance."""
    covered = overlaps.sum()
    if covered / own_volume < min_recall:
        return False                                    # the instance is barely touched by the other map
    reached = 0.0
    for j in torch.argsort(overlaps, descending=True):  # largest contributor first
        if overlaps[j] / other_volumes[j] < min_precision:
            return False                                # a main contributor mostly lies outside this instance
        reached += overlaps[j] / covered
        if reached >= contribution_threshold:
            return True
    return False                                        # unreachable for covered > 0; the reference appends nothing here


def msseg_detection_test(overlap_histogram, min_recall=0.1, contribution_threshold=0.65, min_precision=0.3):
    """Detection test of "Objective Evaluation of Multiple Sclerosis Lesion Segmentation using a Data Management and
    Processing Infrastructure" (MSSEG 2016 / MSSEG-2 2021), reference instance_segmentation_evaluator.py:10-72.

    ``overlap_histogram``: (N + 1, M + 1), element [i, j] = voxels shared by target component i and predicted component
    j (0 = background).  Returns a boolean tensor of length N: target instance i is detected when it is covered by at
    least ``min_recall`` (alpha of the paper) and the predicted instances that make up the first
    ``contribution_threshold`` (gamma) of its overlap each have precision >= ``min_precision`` (1 - beta)."""
    row_volume = overlap_histogram.sum(dim=1)
    column_volume = overlap_histogram.sum(dim=0)
    flags = [_instance_detected(overlap_histogram[i, 1:], row_volume[i], column_volume[1:], min_recall,
                                contribution_threshold, min_precision)
             for i in range(1, overlap_histogram.shape[0])]
    return torch.tensor(flags)


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

    def _subject_stats(self, subject) -> Dict[str, Any]:
        """All statistics of one subject from its (N + 1) x (M + 1) overlap table (reference :131-160)."""
        images = [subject[self.prediction_label_map_name], subject[self.target_label_map_name]]
        pred_data, targ_data = (im.data if hasattr(im, "data") else im["data"] for im in images)
        table, n_target, n_pred = instance_overlap(pred_data, targ_data, self.connectivity)
        hits_target = self.detection_test(table, **self.detection_test_params).sum()
        hits_pred = self.detection_test(table.T, **self.detection_test_params).sum()
        det_recall, det_precision = hits_target / n_target, hits_pred / n_pred
        # voxel-level counts: foreground = any component
        tp, fp, tn, fn = table[1:, 1:].sum(), table[0, 1:].sum(), table[0, 0].sum(), table[1:, 0].sum()
        return {
            'target_components': n_target, 'predicted_components': n_pred,
            'target_detections': hits_target, 'predicted_detections': hits_pred,
            'detection_recall': det_recall, 'detection_precision': det_precision,
            'detection_f1': 2 * (det_recall * det_precision) / (det_recall + det_precision),
            'target_volume': tp + fn, 'prediction_volume': tp + fp,
            'TP': tp, 'FP': fp, 'TN': tn, 'FN': fn,
            'dice': 2 * tp / (2 * tp + fp + fn), 'jaccard': tp / (tp + fp + fn),
            'precision': tp / (tp + fp), 'recall': tp / (tp + fn),
        }

    def __call__(self, subjects):
        table = LabeledTensor(dim_names=['subject', 'stat'],
                              dim_keys=[[subject['name'] for subject in subjects], self.stats_to_output])
        for subject in subjects:
            stats = self._subject_stats(subject)
            for key in self.stats_to_output:
                value = stats[key]
                table[subject['name'], key] = value.item() if isinstance(value, torch.Tensor) else value
        return {'subject_stats': table.to_dataframe(),
                'summary_stats': table.compute_summary_stats(self.summary_stats_to_output)}
