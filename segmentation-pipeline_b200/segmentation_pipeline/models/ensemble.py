"""Test-time-augmentation ensembles -- host-side mirror of the reference's models/ensemble.py:16-103.

The member forwards are native plans; the flips / permutations and the reduction over members are index
shuffles and small reductions on the device tensors the plans return."""
from __future__ import annotations

import itertools
from typing import Sequence

import torch
from torch import nn


def parse_strategy(strategy: str):
    strategies = ('mean', 'majority')
    if strategy not in strategies:
        raise ValueError(f"Ensembling strategy must be one of {strategies} not {strategy}")
    return strategy


def apply_strategy(predictions: Sequence[torch.Tensor], strategy: str):
    """'mean' averages member probabilities; 'majority' takes each member's argmax, the per-voxel mode over
    members, and returns it one-hot (N, C, ...), as reference ensemble.py:16-35."""
    stacked = torch.stack(list(predictions))  # (E, N, C, ...)
    if strategy == 'mean':
        return stacked.mean(dim=0)
    if strategy == 'majority':
        num_classes = stacked.shape[2]
        votes = stacked.argmax(dim=2)
        winner = torch.mode(votes, dim=0).values
        return torch.nn.functional.one_hot(winner, num_classes=num_classes).movedim(-1, 1)
    raise RuntimeError(f"Invalid prediction strategy {strategy}")


def _flip_sets(spatial_dims):
    sets = []
    for size in range(len(spatial_dims) + 1):
        sets.extend(itertools.combinations(spatial_dims, size))
    return sets


class EnsembleModels(nn.Module):
    def __init__(self, models: Sequence[nn.Module], strategy: str = 'mean'):
        super().__init__()
        self.models = nn.ModuleList(models)
        self.strategy = parse_strategy(strategy)

    def forward(self, x):
        return apply_strategy([member(x) for member in self.models], self.strategy)


class EnsembleFlips(nn.Module):
    def __init__(self, model: nn.Module, strategy: str = 'mean', spatial_dims: Sequence[int] = (2, 3, 4)):
        super().__init__()
        self.model = model
        self.strategy = parse_strategy(strategy)
        self.spatial_dims = spatial_dims
        self.flips = _flip_sets(self.spatial_dims)

    def forward(self, x):
        members = []
        for dims in self.flips:
            members.append(self.model(x.flip(dims).contiguous()).flip(dims))
        return apply_strategy(members, self.strategy)


class EnsembleOrientations(nn.Module):
    def __init__(self, model: nn.Module, strategy: str = 'mean'):
        super().__init__()
        self.model = model
        self.strategy = parse_strategy(strategy)
        spatial_dims = (2, 3, 4)
        self.permutations = list(itertools.permutations(spatial_dims))
        self.flips = _flip_sets(spatial_dims)

    def forward(self, x):
        members = []
        for permutation in self.permutations:
            inverse = tuple((torch.argsort(torch.tensor(permutation)) + 2).tolist())
            x_perm = x.permute(0, 1, *permutation)
            for dims in self.flips:
                y = self.model(x_perm.flip(dims).contiguous())
                members.append(y.flip(dims).permute(0, 1, *inverse))
        return apply_strategy(members, self.strategy)
