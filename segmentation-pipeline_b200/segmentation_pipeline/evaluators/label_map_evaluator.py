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

        needs_curve = {'error', 'absolute_error', 'squared_error', 'percent_diff'} & set(self.stats_to_output)
        for argument, value in (("curve_params", curve_params), ("curve_attribute", curve_attribute)):
            if needs_curve and value is None:
                raise ValueError(f"{argument} must be provided")
        self.poly_func = None
        if curve_params is not None and curve_attribute is not None:
            self.poly_func = {name: np.poly1d(coefficients) for name, coefficients in curve_params.items()}

    def _label_stats(self, subject, label_name, voxels: int) -> Dict[str, torch.Tensor]:
        """Volume of one label and, when a growth curve was given, its deviation from the curve at the subject's
        ``curve_attribute`` (reference :83-99)."""
        volume = torch.tensor(voxels)                          # int64, like label.sum() in the reference
        out = {'volume': volume}
        if self.poly_func is not None:
            expected = self.poly_func[label_name](subject[self.curve_attribute])
            delta = volume - expected
            out['error'], out['absolute_error'], out['squared_error'] = delta, abs(delta), delta ** 2
            out['percent_diff'] = (delta / expected) * 100
        return out

    def __call__(self, subjects):
        label_values = subjects[0][self.label_map_name]['label_values']
        names = list(label_values.keys())
        table = LabeledTensor(dim_names=['subject', 'label', 'stat'],
                              dim_keys=[[subject['name'] for subject in subjects], names, self.stats_to_output])
        for subject in subjects:
            image = subject[self.label_map_name]
            data = image.data if hasattr(image, "data") else image["data"]
            # one device histogram pass: a map against itself puts every label's volume on the diagonal (its "TP")
            counts = confusion_counts(data, data, label_values)
            for label_name in names:
                stats = self._label_stats(subject, label_name, counts[label_name][0])
                for key in self.stats_to_output:
                    table[subject['name'], label_name, key] = stats[key].item()
        return {'subject_stats': table.to_dataframe(),
                'summary_stats': table.compute_summary_stats(self.summary_stats_to_output)}
