"""Label-map post-processing on the device -- host-side mirror of the reference's post_processing.py:5-73
(``sort_by_size``, ``unsort_by_size``, ``keep_components``, ``remove_holes``, ``remove_small_components``; called after
inference by research/msseg2/competition/ms-inference.py:47-50, research/dmri_hippo/hippo_inference.py:40-44 and
run_inference.py:206).

Same signatures and return values: numpy label volume in, numpy label volume (same dtype) plus the reference's counts
out.  A torch CUDA tensor is accepted too and then a CUDA tensor comes back (no host round trip of the volume).

What runs where.  Every per-voxel pass is a kernel of libb200seg (``csrc/ccl_kernels.cu``): union-find connected
components (by value with 26-connectivity for ``label(img)``; of the inverted mask with 6-connectivity for
``remove_small_holes``), per-label voxel counts, look-up-table relabelling, the cross-shaped grey dilation fused with
the masked assignment.  The only host work is on the PER-LABEL tables (a few hundred entries): ``np.argsort`` of the
counts -- done with numpy itself so that ties between equally large components fall exactly as in the reference -- and
the thresholding of component sizes.  There is no CPU fallback: without the CUDA library the calls raise."""
from __future__ import annotations

import numpy as np
import torch

MAX_LABEL = 1 << 22          # label values index look-up tables on the device


def _lib():
    import b200seg
    b200seg.load_library()
    return b200seg


class _Volume:
    """The label volume on the device as int32, remembering how to hand it back."""

    def __init__(self, img):
        self.numpy_dtype = None
        self.torch_dtype = None
        if isinstance(img, np.ndarray):
            if not torch.cuda.is_available():
                raise RuntimeError("segmentation_pipeline.post_processing runs on a CUDA device (no CPU fallback)")
            self.numpy_dtype = img.dtype
            host = torch.from_numpy(np.ascontiguousarray(img).astype(np.int32, copy=False))
            self.data = host.cuda()
        elif isinstance(img, torch.Tensor):
            if not img.is_cuda:
                raise RuntimeError("post_processing takes numpy arrays or CUDA tensors (no CPU fallback)")
            self.torch_dtype = img.dtype
            self.data = img.to(torch.int32).contiguous()
            if self.data.data_ptr() == img.data_ptr():
                self.data = self.data.clone()
        else:
            raise TypeError(f"expected a numpy array or a CUDA tensor, got {type(img)}")
        if self.data.dim() != 3:
            raise ValueError(f"post_processing works on (W, H, D) label volumes, got shape {tuple(self.data.shape)}")

    def back(self, data: torch.Tensor):
        if self.numpy_dtype is not None:
            return data.cpu().numpy().astype(self.numpy_dtype, copy=False)
        return data.to(self.torch_dtype)


def _value_counts(lib, data: torch.Tensor):
    """np.unique(data, return_counts=True) for a non-negative int32 device volume: (labels, counts) on the host."""
    lo, hi = (int(v) for v in torch.aminmax(data))
    if lo < 0 or hi >= MAX_LABEL:
        raise ValueError(f"label values must lie in [0, {MAX_LABEL}), got [{lo}, {hi}]")
    counts = lib.overlap_histogram(data, None, hi, 0).view(-1).cpu().numpy()
    labels = np.nonzero(counts)[0]
    return labels, counts[labels]


def _rank_lut(labels, counts, descending: bool, device):
    """-> (int32 device table label -> rank, labels in rank order, counts in rank order) -- post_processing.py:15-23."""
    ids = np.argsort(counts)
    if descending:
        ids = ids[::-1]
    labels, counts = labels[ids], counts[ids]
    lut = np.zeros(int(labels.max()) + 1 if labels.size else 1, dtype=np.int32)
    lut[labels] = np.arange(labels.size, dtype=np.int32)
    return torch.from_numpy(lut).to(device), labels, counts


def sort_by_size(img, descending=False):
    """Relabels ``img`` by the rank of each label's voxel count (post_processing.py:12-26)."""
    lib = _lib()
    vol = _Volume(img)
    with lib.on_device(vol.data):
        labels, counts = _value_counts(lib, vol.data)
        lut, labels, counts = _rank_lut(labels, counts, descending, vol.data.device)
        out = lib.relabel_lut(vol.data, lut)
        dtype = vol.numpy_dtype if vol.numpy_dtype is not None else np.int64
        return vol.back(out), labels.astype(dtype, copy=False), counts


def unsort_by_size(img, sorted_labels):
    """Inverse of ``sort_by_size`` given its ``unique_labels`` (post_processing.py:5-9)."""
    lib = _lib()
    vol = _Volume(img)
    with lib.on_device(vol.data):
        lut = torch.from_numpy(np.asarray(sorted_labels).astype(np.int32)).to(vol.data.device)
        return vol.back(lib.relabel_lut(vol.data, lut))


def keep_components(img, num, max_dilations=100):
    """Keeps the background and the ``num`` largest connected components; every other component is eaten by the labels
    dilating into it (post_processing.py:29-49).  -> (img, num_components_removed, num_elements_removed)."""
    lib = _lib()
    vol = _Volume(img)
    num = int(num)
    num_components_removed = num_elements_removed = 0
    with lib.on_device(vol.data):
        data = vol.data
        device = data.device
        for i in range(max_dilations):
            comp, n = lib.connected_components(data, connectivity=3, by_value=1)            # label(img)
            counts = lib.overlap_histogram(comp, None, n, 0).view(-1).cpu().numpy()
            labels = np.nonzero(counts)[0]
            ids = np.argsort(counts[labels])[::-1]                                          # sort_by_size(descending)
            ranked = labels[ids]
            keep_lut = np.zeros(n + 1, dtype=np.int32)
            keep_lut[ranked[:num + 1]] = 1                                                  # rank <= num
            removed = int(counts[ranked[num + 1:]].sum())
            if i == 0:
                num_elements_removed = removed
                num_components_removed = (labels.size - 1) - num                            # img_comp_sorted.max() - num
            if removed == 0:
                break
            img_labels, img_counts = _value_counts(lib, data)
            lut, sorted_labels, _ = _rank_lut(img_labels, img_counts, False, device)        # sort_by_size(img)
            keep_dev = torch.from_numpy(keep_lut).to(device)
            sorted_img = lib.relabel_lut(data, lut)
            to_dilate = lib.relabel_masked(data, lut, comp, keep_dev)                       # sorted_img * keep
            remove = lib.label_equals(lib.relabel_lut(comp, keep_dev), 0)                   # ~keep
            sorted_img = lib.dilate_where(to_dilate, remove, sorted_img)
            data = lib.relabel_lut(sorted_img, torch.from_numpy(sorted_labels.astype(np.int32)).to(device))
        return vol.back(data), num_components_removed, num_elements_removed


def _remove_holes(lib, data: torch.Tensor, hole_size, max_dilations):
    total_holes = 0
    hole_size = int(hole_size)
    for i in range(max_dilations):
        # small_holes = ~mask & remove_small_holes(mask, hole_size): the 6-connected components of (img <= 0) that
        # hold fewer than hole_size voxels
        comp, n = lib.connected_components(data, connectivity=1, by_value=2)
        counts = lib.overlap_histogram(comp, None, n, 0).view(-1).cpu().numpy()
        small = counts < hole_size if hole_size > 0 else np.zeros_like(counts, dtype=bool)
        small[0] = False
        num_holes = int(counts[small].sum())
        if i == 0:
            total_holes = num_holes
        if num_holes == 0:
            break
        holes = lib.relabel_lut(comp, torch.from_numpy(small.astype(np.int32)).to(data.device))
        data = lib.dilate_where(data, holes, data)                                          # img[holes] = dilation(img)[holes]
    return data, total_holes


def remove_holes(img, hole_size, max_dilations=100):
    """Fills background holes smaller than ``hole_size`` voxels from their surroundings (post_processing.py:52-64).
    -> (img, number of hole voxels found in the first sweep)."""
    lib = _lib()
    vol = _Volume(img)
    with lib.on_device(vol.data):
        data, total = _remove_holes(lib, vol.data, hole_size, max_dilations)
        return vol.back(data), total


def remove_small_components(img, component_size, max_dilations=100):
    """Zeroes foreground components smaller than ``component_size`` voxels (post_processing.py:67-73: the holes of the
    inverted image).  -> (img, number of voxels removed)."""
    lib = _lib()
    vol = _Volume(img)
    with lib.on_device(vol.data):
        inverted = lib.label_equals(vol.data, 0)
        holes_removed, counts = _remove_holes(lib, inverted, component_size, max_dilations)
        data = lib.mask_assign(vol.data, holes_removed, 0)
        return vol.back(data), counts
