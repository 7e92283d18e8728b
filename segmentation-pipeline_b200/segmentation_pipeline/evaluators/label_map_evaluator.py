"""LabelMapEvaluator -- per-label voxel counts (and optional age-curve error statistics) with the interface of
the reference's evaluators/label_map_evaluator.py:38-109; the counts come from the same device histogram pass
as the segmentation evaluator (a label map against itself puts the volumes on the diagonal)."""
from __future__ import annotations

from typing import Dict, Sequence, Union

import numpy as np
import torch

from .evaluator import Evaluator
from .labeled_tensor import LabeledTensor
from .segmentation_evaluator import confusion_counts


class LabelMapEvaluator(Evaluator):
    def __init__(
            self,
            label_map_name: str,
            curve_params: Union[Dict[str, np.ndarray], None] = None,
            curve_attribute: Union[str, None] = None,
            stats_to_output: Sequence[str] = ('volume',),
            summary_stats_to_output: Sequence[str] = ('mean', 'std', 'min', 'max'),
    ):
        self.label_map_name = label_map_name
        self.curve_params = curve_params
        self.curve_attribute = curve_attribute
        self.stats_to_output = stats_to_output
        self.summary_stats_to_output = summary_stats_to_output

        curve_stats = ['error', 'absolute_error', 'squared_error', 'percent_diff']
        if any(stat in curve_stats for stat in self.stats_to_output):
            if curve_params is None:
                raise ValueError("curve_params must be provided")
            if curve_attribute is None:
                raise ValueError("curve_attribute must be provided")
        if curve_params is not None and curve_attribute is not None:
            self.poly_func = {label: np.poly1d(param) for label, param in curve_params.items()}
        else:
            self.poly_func = None

    def __call__(self, subjects):
        label_values = subjects[0][self.label_map_name]['label_values']
        label_names = list(label_values.keys())
        subject_names = [subject['name'] for subject in subjects]
        subject_stats = LabeledTensor(dim_names=['subject', 'label', 'stat'],
                                      dim_keys=[subject_names, label_names, self.stats_to_output])
        for subject in subjects:
            image = subject[self.label_map_name]
            data = image["data"] if not hasattr(image, "data") else image.data
            counts = confusion_counts(data, data, label_values)
            for label_name in label_names:
                volume = torch.tensor(counts[label_name][0])   # TP of a map against itself = its volume (int64)
                stats = {'volume': volume}
                if self.poly_func is not None:
                    predicted = self.poly_func[label_name](subject[self.curve_attribute])
                    error = volume - predicted
                    stats.update({'error': error, 'absolute_error': abs(error), 'squared_error': error ** 2,
                                  'percent_diff': (error / predicted) * 100})
                for stat_name in self.stats_to_output:
                    subject_stats[subject['name'], label_name, stat_name] = stats[stat_name].item()
        summary_stats = subject_stats.compute_summary_stats(self.summary_stats_to_output)
        return {'subject_stats': subject_stats.to_dataframe(), 'summary_stats': summary_stats}
