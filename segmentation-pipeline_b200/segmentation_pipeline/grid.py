"""Patch grid of the sliding window (host logic) -- the location table and padding rules of
``tio.GridSampler`` / ``tio.GridAggregator`` (torchio 0.18.45, the release pinned by the reference in
research/msseg2/competition/docker-requirements.txt:44) as used by ``PatchPredict`` at
segmentation_pipeline/prediction.py:132-143 of the reference.  Pure integer work on the host; the voxel work
(extraction, overlap-add, divide, crop, argmax) is done by libb200seg."""
from __future__ import annotations

import itertools
from typing import List, Optional, Sequence, Tuple, Union


def triple(value) -> Tuple[int, int, int]:
    if isinstance(value, int):
        return (value, value, value)
    value = tuple(int(v) for v in value)
    if len(value) == 1:
        return value * 3
    if len(value) != 3:
        raise ValueError(f"expected an int or 3 values, got {value}")
    return value


class PatchGrid:
    """Locations (sorted lexicographically, as torchio returns them), padded extent and per-axis coverage."""

    def __init__(self, spatial_shape: Sequence[int], patch_size, patch_overlap=(0, 0, 0),
                 padding_mode: Union[str, float, None] = None):
        self.patch_size = triple(patch_size)
        self.patch_overlap = triple(patch_overlap)
        self.padding_mode = padding_mode
        self.volume_padded = padding_mode is not None
        self.border = tuple(o // 2 for o in self.patch_overlap) if self.volume_padded else (0, 0, 0)
        self.spatial_shape = tuple(int(s) for s in spatial_shape)
        self.padded_shape = tuple(s + 2 * b for s, b in zip(self.spatial_shape, self.border))
        self._validate()
        self.axis_starts = [self._starts(s, p, o) for s, p, o in
                            zip(self.padded_shape, self.patch_size, self.patch_overlap)]
        self.locations: List[Tuple[int, ...]] = [
            (i, j, k, i + self.patch_size[0], j + self.patch_size[1], k + self.patch_size[2])
            for i, j, k in itertools.product(*self.axis_starts)]   # product of sorted lists is lexicographic

    def _validate(self) -> None:
        if any(p > s for p, s in zip(self.patch_size, self.padded_shape)):
            raise ValueError(f"Patch size {self.patch_size} cannot be larger than image size {self.padded_shape}")
        if any(o >= p for o, p in zip(self.patch_overlap, self.patch_size)):
            raise ValueError(f"Patch overlap {self.patch_overlap} must be smaller than patch size "
                             f"{self.patch_size}")
        if any(o % 2 for o in self.patch_overlap):
            raise ValueError(f"Patch overlap must be a tuple of even integers, not {self.patch_overlap}")

    @staticmethod
    def _starts(size: int, patch: int, overlap: int) -> List[int]:
        starts = list(range(0, size + 1 - patch, patch - overlap))
        if starts[-1] != size - patch:
            starts.append(size - patch)   # last patch flush with the far border
        return starts

    def axis_counts(self) -> List[List[int]]:
        """Coverage count per index along each axis; the 3-D count map is their outer product."""
        counts = []
        for size, patch, starts in zip(self.padded_shape, self.patch_size, self.axis_starts):
            c = [0] * size
            for s in starts:
                for i in range(s, s + patch):
                    c[i] += 1
            counts.append(c)
        return counts

    def hann_windows(self):
        """Per-axis window of torchio's 'hann' aggregation: ``torch.hann_window(size + 2, periodic=False)[1:-1]`` (the
        two zero end points dropped), fp32.  The 3-D window is their outer product."""
        import torch
        return [torch.hann_window(p + 2, periodic=False)[1:-1].contiguous() for p in self.patch_size]

    def axis_window_sums(self):
        """Summed window per index along each axis (fp32, patches added in ascending start order); the 3-D sum of
        windows over the product grid is the outer product of the three."""
        import torch
        sums = []
        for size, patch, starts, win in zip(self.padded_shape, self.patch_size, self.axis_starts, self.hann_windows()):
            acc = torch.zeros(size, dtype=torch.float32)
            for s0 in starts:
                acc[s0:s0 + patch] += win
            sums.append(acc)
        return sums

    @property
    def pad_mode_code(self) -> int:
        """0 none, 1 'edge' (clamp), 2 constant -- the modes b200seg_grid_extract implements."""
        if self.padding_mode is None:
            return 0
        if isinstance(self.padding_mode, (int, float)):
            return 2
        if self.padding_mode == "edge":
            return 1
        if self.padding_mode == "constant":
            return 2
        raise NotImplementedError(f"padding_mode {self.padding_mode!r} is not implemented on the device "
                                  f"('edge', 'constant' or a number are)")

    @property
    def pad_value(self) -> float:
        return float(self.padding_mode) if isinstance(self.padding_mode, (int, float)) else 0.0

    def batches(self, batch_size: int):
        for start in range(0, len(self.locations), batch_size):
            yield self.locations[start:start + batch_size]
