"""ORACLE (test infrastructure only) -- CPU restatement of torchio's grid sampling / aggregation.

Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of
``bench.py`` may import this module.  The product (``segmentation-pipeline_b200/``) never does.

What it restates
----------------
``tio.GridSampler`` and ``tio.GridAggregator`` as called by the reference at
``segmentation_pipeline/prediction.py:132`` (sampler), ``:133`` (``__getitem__`` through the DataLoader),
``:134``/``:141`` (aggregator ``add_batch``) and ``:143`` (``get_output_tensor``).

torchio is a third-party dependency that is NOT vendored in ``/root/reference`` and is not installed in
this image.  The reference pins it as ``torchio==0.18.45``
(``research/msseg2/competition/docker-requirements.txt:44``).  The algorithm below restates that release's
``torchio/data/sampler/grid.py`` (``GridSampler._pad / _parse_sizes / _get_patches_locations / __getitem__``)
and ``torchio/data/inference/aggregator.py`` (``GridAggregator.add_batch / crop_batch / get_output_tensor``).

PINNED (as far as this image allows): the reference repository holds no test or fixture for this path (SURVEY.md
section 8c) and the torchio wheel cannot be imported here, so the restatement is pinned to
``tests/golden/grid_torchio.json`` (written by ``oracle/make_grid_golden.py``): torchio's OWN unit-test fixtures for
``GridSampler`` locations and ``GridAggregator`` crop / average (restated from its test-suite and re-derived by hand in
that script) plus brute-force per-voxel cases that share no code with this module, including the centre-crop quirk of
``crop_batch``.  ``tests/test_oracle_grid.py`` checks this module against them; ``tests/test_gpu_kernels.py`` checks
the CUDA path against them and, when the box has torchio, against ``tio.GridSampler`` / ``tio.GridAggregator``
themselves.  What remains unpinned: a byte-for-byte run of torchio 0.18.45 in this image.
"""
from __future__ import annotations

from typing import Optional, Sequence, Tuple, Union

import numpy as np

Triple = Union[int, Sequence[int]]


def to_triple(value: Triple) -> np.ndarray:
    """torchio ``to_tuple(value, length=3)``: an int is broadcast to the three spatial axes."""
    if isinstance(value, (int, np.integer)):
        return np.array([int(value)] * 3, dtype=np.int64)
    value = tuple(int(v) for v in value)
    if len(value) == 1:
        value = value * 3
    if len(value) != 3:
        raise ValueError(f"expected 1 or 3 values, got {value}")
    return np.array(value, dtype=np.int64)


def parse_sizes(image_size: Sequence[int], patch_size: Triple, patch_overlap: Triple) -> None:
    """``GridSampler._parse_sizes``: same three ValueErrors, same order."""
    image_size = np.array(image_size)
    patch_size = to_triple(patch_size)
    patch_overlap = to_triple(patch_overlap)
    if np.any(patch_size > image_size):
        raise ValueError(f"Patch size {tuple(patch_size)} cannot be larger than image size {tuple(image_size)}")
    if np.any(patch_overlap >= patch_size):
        raise ValueError(f"Patch overlap {tuple(patch_overlap)} must be smaller than patch size {tuple(patch_size)}")
    if np.any(patch_overlap % 2):
        raise ValueError(f"Patch overlap must be a tuple of even integers, not {tuple(patch_overlap)}")


def axis_starts(size: int, patch: int, overlap: int) -> list:
    """Per-axis start indices: ``range(0, size + 1 - patch, patch - overlap)`` plus a final start flush with
    the far border when the stride does not land on it."""
    starts = list(range(0, size + 1 - patch, patch - overlap))
    if starts[-1] != size - patch:
        starts.append(size - patch)
    return starts


def grid_locations(image_size: Sequence[int], patch_size: Triple, patch_overlap: Triple) -> np.ndarray:
    """``GridSampler._get_patches_locations``: (n, 6) int64 ``[i0, j0, k0, i1, j1, k1]`` rows, unique and
    sorted lexicographically."""
    parse_sizes(image_size, patch_size, patch_overlap)
    patch_size = to_triple(patch_size)
    patch_overlap = to_triple(patch_overlap)
    per_axis = [axis_starts(int(s), int(p), int(o)) for s, p, o in zip(image_size, patch_size, patch_overlap)]
    ini = np.array(np.meshgrid(*per_axis)).reshape(3, -1).T
    ini = np.unique(ini, axis=0)
    fin = ini + patch_size
    locations = np.hstack((ini, fin))
    return np.array(sorted(locations.tolist()), dtype=np.int64)


def pad_volume(volume: np.ndarray, patch_overlap: Triple, padding_mode: Union[str, float, None]) -> np.ndarray:
    """``GridSampler._pad``: when ``padding_mode`` is not None every spatial axis of the (C, W, H, D) volume
    is padded by ``overlap // 2`` on both sides with ``numpy.pad``; a number means constant fill."""
    if padding_mode is None:
        return volume
    border = to_triple(patch_overlap) // 2
    widths = [(0, 0)] + [(int(b), int(b)) for b in border]
    if isinstance(padding_mode, (int, float)):
        return np.pad(volume, widths, mode="constant", constant_values=padding_mode)
    return np.pad(volume, widths, mode=padding_mode)


def extract_patches(volume: np.ndarray, locations: np.ndarray) -> np.ndarray:
    """``GridSampler.__getitem__`` for every location: crop ``[ini, fin)`` -> (n, C, p0, p1, p2)."""
    return np.stack([volume[:, i0:i1, j0:j1, k0:k1] for i0, j0, k0, i1, j1, k1 in locations])


def aggregate_average(patches: np.ndarray, locations: np.ndarray, spatial_shape: Sequence[int]
                      ) -> Tuple[np.ndarray, np.ndarray]:
    """``GridAggregator.add_batch`` in ``'average'`` mode over all patches in order; returns (sum, count).
    Both buffers have the dtype of the patches (torchio allocates them with ``dtype=batch.dtype``)."""
    out = np.zeros((patches.shape[1], *spatial_shape), dtype=patches.dtype)
    cnt = np.zeros_like(out)
    for patch, (i0, j0, k0, i1, j1, k1) in zip(patches, locations):
        out[:, i0:i1, j0:j1, k0:k1] += patch
        cnt[:, i0:i1, j0:j1, k0:k1] += 1
    return out, cnt


def aggregate_crop(patches: np.ndarray, locations: np.ndarray, spatial_shape: Sequence[int],
                   patch_overlap: Triple, volume_padded: bool) -> np.ndarray:
    """``GridAggregator.add_batch`` in ``'crop'`` mode (``crop_batch``): each patch is trimmed by
    ``overlap // 2`` on every side that is not flush with the volume border (all sides if the volume was
    padded), the trimmed block being taken from the CENTRE of the patch, and assigned."""
    border = to_triple(patch_overlap) // 2
    out = np.zeros((patches.shape[1], *spatial_shape), dtype=patches.dtype)
    patch_shape = np.array(patches.shape[2:])
    size = np.array(spatial_shape)
    for patch, loc in zip(patches, locations):
        ini = loc[:3].copy()
        fin = loc[3:].copy()
        b_ini = border.copy()
        b_fin = border.copy()
        if not volume_padded:
            b_ini[ini == 0] = 0
            b_fin[fin == size] = 0
        ini = ini + b_ini
        fin = fin - b_fin
        crop_shape = fin - ini
        left = ((patch_shape - crop_shape) / 2).astype(int)
        right = left + crop_shape
        out[:, ini[0]:fin[0], ini[1]:fin[1], ini[2]:fin[2]] = \
            patch[:, left[0]:right[0], left[1]:right[1], left[2]:right[2]]
    return out


def hann_window_3d(patch_size: Triple) -> np.ndarray:
    """torchio's ``GridAggregator._get_hann_window`` (torchio >= 0.19, data/inference/aggregator.py; restated from the
    upstream source, torchio is not in this image): per axis ``torch.hann_window(size + 2, periodic=False)[1:-1]``,
    multiplied up axis by axis in float32."""
    window = np.ones((1, 1, 1), dtype=np.float32)
    for axis, size in enumerate(to_triple(patch_size)):
        n = np.arange(size + 2, dtype=np.float64)
        w1d = (0.5 - 0.5 * np.cos(2.0 * np.pi * n / (size + 1))).astype(np.float32)[1:-1]
        shape = [1, 1, 1]
        shape[axis] = size
        window = window * w1d.reshape(shape)
    return window.astype(np.float32)


def aggregate_hann(patches: np.ndarray, locations: np.ndarray, spatial_shape: Sequence[int]
                   ) -> Tuple[np.ndarray, np.ndarray]:
    """``GridAggregator.add_batch`` in ``'hann'`` mode: ``output += patch * window``, ``avgmask += window`` per patch
    in order; returns (sum, mask).  ``get_output_tensor`` divides them like the average mode."""
    window = hann_window_3d(patches.shape[2:])
    out = np.zeros((patches.shape[1], *spatial_shape), dtype=patches.dtype)
    mask = np.zeros_like(out)
    for patch, (i0, j0, k0, i1, j1, k1) in zip(patches, locations):
        out[:, i0:i1, j0:j1, k0:k1] += patch * window
        mask[:, i0:i1, j0:j1, k0:k1] += window
    return out, mask


def finalize(out: np.ndarray, cnt: Optional[np.ndarray], patch_overlap: Triple, volume_padded: bool) -> np.ndarray:
    """``GridAggregator.get_output_tensor``: ``true_divide(out, cnt)`` in average mode, then crop the
    ``overlap // 2`` border if the sampler padded the volume."""
    res = out if cnt is None else np.true_divide(out, cnt)
    if volume_padded:
        b = to_triple(patch_overlap) // 2
        w, h, d = res.shape[1:]
        res = res[:, b[0]:w - b[0], b[1]:h - b[1], b[2]:d - b[2]]
    return res


def sliding_window(volume: np.ndarray, model_fn, patch_size: Triple, patch_overlap: Triple,
                   padding_mode: Union[str, float, None], overlap_mode: str = "average",
                   patch_batch_size: int = 16) -> np.ndarray:
    """The loop of ``PatchPredict.predict`` (``prediction.py:131-143``) for one subject.  ``model_fn`` maps a
    (B, C, p, p, p) float32 array to (B, C_out, p, p, p)."""
    padded = pad_volume(volume, patch_overlap, padding_mode)
    spatial = padded.shape[1:]
    locations = grid_locations(spatial, patch_size, patch_overlap)
    outs = []
    for b0 in range(0, len(locations), patch_batch_size):
        locs = locations[b0:b0 + patch_batch_size]
        outs.append(np.asarray(model_fn(extract_patches(padded, locs))))
    y = np.concatenate(outs)
    if overlap_mode == "average":
        out, cnt = aggregate_average(y, locations, spatial)
        return finalize(out, cnt, patch_overlap, padding_mode is not None)
    if overlap_mode == "crop":
        out = aggregate_crop(y, locations, spatial, patch_overlap, padding_mode is not None)
        return finalize(out, None, patch_overlap, padding_mode is not None)
    if overlap_mode == "hann":
        out, mask = aggregate_hann(y, locations, spatial)
        return finalize(out, mask, patch_overlap, padding_mode is not None)
    raise ValueError(f'Overlap mode must be "crop", "average" or "hann" but "{overlap_mode}" was passed')
