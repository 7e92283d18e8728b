"""Named-axis float32 tensor with summary statistics -- host-side packaging of the evaluator results, same
behaviour as the reference's evaluators/labeled_tensor.py:11-110 (float32 storage, string keys per axis,
summary stats over the first axis after dropping non-finite values)."""
from __future__ import annotations

import copy
from itertools import product
from typing import Sequence

import pandas as pd
import torch

from ..utils import as_list, is_sequence


class LabeledTensor:
    def __init__(self, dim_names: Sequence[str], dim_keys: Sequence[Sequence[str]]):
        if len(dim_names) != len(dim_keys):
            raise ValueError(f"The number of dimension names ({len(dim_names)}) "
                             f"does not match the number of dimension keys ({len(dim_keys)}")
        self.dim_names = dim_names
        self.dim_keys = dim_keys
        self.dim_key_map = [{key: index for index, key in enumerate(keys)} for keys in dim_keys]
        self.data = torch.zeros([len(keys) for keys in dim_keys])

    def parse_key(self, key):
        key = as_list(key)
        if any(k is Ellipsis for k in key):
            raise NotImplementedError("Elipsis indexing is not supported for LabeledTensors")
        for axis, k in enumerate(key):
            lookup = self.dim_key_map[axis]
            if isinstance(k, str):
                key[axis] = lookup[k]
            elif is_sequence(k):
                key[axis] = [lookup[e] if isinstance(e, str) else e for e in k]
        return tuple(key)

    def __getitem__(self, key) -> torch.Tensor:
        return self.data[self.parse_key(key)]

    def __setitem__(self, key, value):
        self.data[self.parse_key(key)] = value

    def to_dataframe(self):
        columns = {dim: [] for dim in self.dim_names[:-1]}
        columns.update({dim: [] for dim in self.dim_keys[-1]})
        for keys in product(*self.dim_keys[:-1]):
            for dim, key in zip(self.dim_names[:-1], keys):
                columns[dim].append(key)
            for dim, value in zip(self.dim_keys[-1], self[keys].tolist()):
                columns[dim].append(value)
        return pd.DataFrame(columns)

    def to_dict(self):
        nested = 0
        for keys in reversed(self.dim_keys):
            nested = {key: copy.deepcopy(nested) for key in keys}
        for key in product(*self.dim_keys):
            cursor = nested
            for k in key[:-1]:
                cursor = cursor[k]
            cursor[key[-1]] = self[key].item()
        return nested

    def compute_summary_stats(self, summary_stats_to_output):
        summary = LabeledTensor(dim_names=["summary_stat", *self.dim_names[1:]],
                                dim_keys=[summary_stats_to_output, *self.dim_keys[1:]])
        funcs = LabeledTensor.get_summary_stat_funcs()
        for keys in product(*self.dim_keys[1:]):
            values = self[(slice(None), *keys)]
            for stat in summary_stats_to_output:
                summary[(stat, *keys)] = funcs[stat](values).item()
        return summary

    @staticmethod
    def fix_tensor(x):
        x = x[x.isfinite()]
        return torch.tensor([0.]) if x.shape[0] == 0 else x

    @staticmethod
    def get_summary_stat_funcs(dim: int = 0):
        fix = LabeledTensor.fix_tensor
        return {
            'mean': lambda x: torch.mean(fix(x), dim=dim),
            'median': lambda x: torch.median(fix(x), dim=dim).values,
            'mode': lambda x: torch.mode(fix(x), dim=dim).values,
            'std': lambda x: torch.std(fix(x), dim=dim),
            'min': lambda x: torch.min(fix(x), dim=dim).values,
            'max': lambda x: torch.max(fix(x), dim=dim).values,
        }
