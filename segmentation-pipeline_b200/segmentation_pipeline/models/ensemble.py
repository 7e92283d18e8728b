"""Test-time-augmentation ensembles -- host-side mirror of the reference's models/ensemble.py:16-103.

With a native member network the whole ensemble runs on the device without materialising anything the reference
materialises: the flip / permutation of member e's input is folded into the NCDHW -> blocked packing kernel
(``b200seg_pack_ncdhw_tta``), the member runs as its native plan, and its output is un-flipped / un-permuted while it is
added to the running fp32 sum ('mean') or to the per-voxel uint8 vote counts of its argmax ('majority')
(``b200seg_tta_accumulate``); ``b200seg_tta_finalize`` scales the sum or writes the int64 one-hot of the vote winner
(smallest label on ties, like ``torch.mode``).  No flipped copies, no (E, N, C, ...) stack.  Members that are not
native modules (arbitrary ``nn.Module``) take the reference's tensor expressions."""
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


def _native(model) -> bool:
    from .components import _NativeForward, StochasticMatrix
    return isinstance(model, _NativeForward) and not isinstance(model, StochasticMatrix)


class _Fused:
    """Running reduction over ensemble members on the device (see the module docstring)."""

    def __init__(self, x: torch.Tensor, strategy: str):
        import b200seg
        b200seg.load_library()
        self.lib = b200seg
        if not x.is_cuda:
            raise RuntimeError("the b200 ensembles run on CUDA tensors only (no CPU fallback)")
        self.x = x.detach().to(torch.float32).contiguous()
        self.in_dtype = x.dtype
        self.strategy = strategy
        self.acc = self.votes = None
        self.members = 0

    def _buffers(self, channels: int):
        shape = (self.x.shape[0], channels, *self.x.shape[2:])
        if self.strategy == 'mean' and self.acc is None:
            self.acc = torch.zeros(shape, dtype=torch.float32, device=self.x.device)
        if self.strategy == 'majority' and self.votes is None:
            self.votes = torch.zeros(shape, dtype=torch.uint8, device=self.x.device)

    def add_native(self, model, perm=(0, 1, 2), flip=(False, False, False)) -> None:
        """One member = native ``model`` on x.permute(perm).flip(flip), accumulated in the original orientation."""
        from . import _engine
        if model.training:
            raise NotImplementedError("ensembles of native members need model.eval() (inference path only)")
        with self.lib.on_device(self.x):
            compiled = _engine.compiled_for(model, _engine._resolve_precision(model, self.x if self.in_dtype == torch.float32
                                                                              else self.x.new_empty(0, dtype=self.in_dtype)),
                                            self.x.device)
            n = self.x.shape[0]
            ext = tuple(self.x.shape[2 + p] for p in perm)
            ws = compiled._workspace(n, *ext)
            self.lib.pack_ncdhw_tta(self.x, perm, flip, ws["in"].view(compiled.plan.in_channels))
            y = compiled.run_blocked(n, *ext)
            self.add_output(y, perm, flip)

    def add_output(self, y: torch.Tensor, perm=(0, 1, 2), flip=(False, False, False)) -> None:
        with self.lib.on_device(self.x):
            y = y.detach().to(torch.float32).contiguous()
            self._buffers(y.shape[1])
            self.lib.tta_accumulate(y, perm, flip, self.acc, self.votes)
            self.members += 1

    def result(self) -> torch.Tensor:
        with self.lib.on_device(self.x):
            if self.strategy == 'mean':
                self.lib.tta_finalize(self.acc, None, None, self.members)
                return self.acc if self.in_dtype == torch.float32 else self.acc.to(self.in_dtype)
            onehot = torch.empty(self.votes.shape, dtype=torch.int64, device=self.x.device)
            self.lib.tta_finalize(None, self.votes, onehot, self.members)
            return onehot


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
        if x.is_cuda and len(self.models) <= 255:
            fused = _Fused(x, self.strategy)
            for member in self.models:
                if _native(member):
                    fused.add_native(member)
                else:
                    fused.add_output(member(x))
            return fused.result()
        return apply_strategy([member(x) for member in self.models], self.strategy)


class EnsembleFlips(nn.Module):
    def __init__(self, model: nn.Module, strategy: str = 'mean', spatial_dims: Sequence[int] = (2, 3, 4)):
        super().__init__()
        self.model = model
        self.strategy = parse_strategy(strategy)
        self.spatial_dims = spatial_dims
        self.flips = _flip_sets(self.spatial_dims)

    def forward(self, x):
        if _native(self.model) and x.dim() == 5 and all(d in (2, 3, 4) for d in self.spatial_dims):
            fused = _Fused(x, self.strategy)
            for dims in self.flips:
                fused.add_native(self.model, (0, 1, 2), tuple((k + 2) in dims for k in range(3)))
            return fused.result()
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
        if _native(self.model) and x.dim() == 5:
            fused = _Fused(x, self.strategy)
            for permutation in self.permutations:
                perm = tuple(p - 2 for p in permutation)
                for dims in self.flips:
                    fused.add_native(self.model, perm, tuple((k + 2) in dims for k in range(3)))
            return fused.result()
        members = []
        for permutation in self.permutations:
            inverse = tuple((torch.argsort(torch.tensor(permutation)) + 2).tolist())
            x_perm = x.permute(0, 1, *permutation)
            for dims in self.flips:
                y = self.model(x_perm.flip(dims).contiguous())
                members.append(y.flip(dims).permute(0, 1, *inverse))
        return apply_strategy(members, self.strategy)
