"""Predictors -- host-side mirror of the reference's segmentation_pipeline/prediction.py (Predictor :41-54,
StandardPredict :57-102, PatchPredict :105-152, add_evaluation_labels :155-170) with the same constructor
arguments, attributes and return values, driving libb200seg instead of torchio + ATen:

  PatchPredict.predict       volume -> device once; patches are cut on the device straight into the network's
                             blocked input buffer (b200seg_grid_extract, padding never materialised); the
                             network runs as one native plan per patch batch; b200seg_overlap_add accumulates
                             in the reference's patch order (bit-identical sums); b200seg_finalize divides by
                             the separable coverage count, crops the padding and emits probabilities + labels.
  add_evaluation_labels      argmax on the device (b200seg_argmax), ties -> lowest index, int64 (1, W, H, D).
"""
from __future__ import annotations

import copy
from abc import ABC, abstractmethod
from typing import Any, Dict, Optional, Sequence, Tuple, Union

import torch
from torch import nn

from . import _tio
from .grid import PatchGrid, triple
from .models import _engine
from .models.components import _NativeForward, StochasticMatrix
from .utils import Config, collate_subjects

LABELS_KEY = "_b200_labels"   # device uint8 label map stashed next to 'y_pred' by PatchPredict

# Device-side regrouping of patches.  ``patch_batch_size`` is a MEMORY knob of the reference (the msseg2 inference script
# uses 1 to fit a 32 GB V100, research/msseg2/competition/ms-inference.py:32): patch forwards are independent and the
# aggregation runs in sorted patch order whatever the grouping, so the result does not depend on it bit for bit (tested).
# On a 180 GB B200 a native network therefore evaluates up to DEVICE_BATCH patches per launch sequence when the caller's
# batch is smaller and the activation workspace fits in a quarter of the free memory.  0 = literal patch_batch_size.
DEVICE_BATCH = [int(__import__("os").environ.get("B200SEG_DEVICE_BATCH", "48"))]


def set_device_batch(n: int) -> None:
    """Upper bound of the device-side patch regrouping of native networks (0: use ``patch_batch_size`` literally)."""
    DEVICE_BATCH[0] = int(n)



def _lib():
    import b200seg
    b200seg.load_library()
    return b200seg


def split_and_flip(x: torch.Tensor) -> torch.Tensor:
    """Sagittal halves folded into the batch, second half mirrored (reference prediction.py:16-20)."""
    halves = list(x.split(x.shape[2] // 2, dim=2))
    halves[1] = halves[1].flip(2)
    return torch.cat(halves, dim=0)


def reverse_split_and_flip(x: torch.Tensor) -> torch.Tensor:
    halves = list(x.split(x.shape[0] // 2, dim=0))
    halves[1] = halves[1].flip(2)
    return torch.cat(halves, dim=2)


def apply_stochastic_matrix(y_pred, y_prior):
    n, c = y_prior.shape[:2]
    y_pred = y_pred.reshape(n, c, c, *y_prior.shape[2:])
    return (y_pred * y_prior[:, None]).sum(dim=1)


def _require_cuda(device) -> torch.device:
    device = torch.device(device)
    if device.type != "cuda":
        raise RuntimeError("the b200 predictors run on a CUDA device only (no CPU fallback)")
    return device


class Predictor(ABC, Config):
    """Representation to get model predictions"""

    @abstractmethod
    def predict(
        self,
        model: nn.Module,
        device: torch.device,
        subjects: Sequence[Any],
        label_attributes: Optional[Dict[str, Any]] = None,
    ) -> Tuple[Sequence[Any], Dict[str, torch.Tensor]]:
        """Creates predictions for subjects and adds the predictions as an image with name 'y_pred' and
        batch with with key 'y_pred'"""
        raise NotImplementedError()


class StandardPredict(Predictor):
    """ Creates predictions on whole images"""

    def __init__(
            self,
            image_names: Sequence[str] = ("X",),
            sagittal_split: bool = False,
            refine_image: str = None,
    ):
        image_names = list(image_names)
        if refine_image is not None and refine_image not in image_names:
            image_names.append(refine_image)
        self.image_names = image_names
        self.sagittal_split = sagittal_split
        self.refine_image = refine_image

    def predict(self, model, device, subjects, label_attributes=None):
        device = _require_cuda(device)
        label_attributes = {} if label_attributes is None else label_attributes
        with _lib().on_device(device):
            # H2D: every subject's volume straight into its slot of the device batch (asynchronous when the host tensor
            # is pinned) instead of torch.stack on the host + one pageable copy (utils/utils.py:75-85)
            batch = {}
            for name in self.image_names:
                first = subjects[0][name]["data"]
                stacked = torch.empty((len(subjects), *first.shape), dtype=first.dtype, device=device)
                for i, subject in enumerate(subjects):
                    stacked[i].copy_(subject[name]["data"], non_blocking=True)
                batch[name] = stacked
            if self.sagittal_split:
                y_pred = reverse_split_and_flip(model(split_and_flip(batch['X']).contiguous()))
            else:
                y_pred = model(batch["X"])
            batch['y_pred'] = y_pred
            # D2H: ONE asynchronous copy of the whole batch into pinned memory and one synchronisation, instead of a
            # blocking .cpu() per subject (prediction.py:97); the per-subject images are views of that buffer
            y_det = y_pred.detach()
            y_host = torch.empty(y_det.shape, dtype=y_det.dtype, pin_memory=True)
            y_host.copy_(y_det, non_blocking=True)
            torch.cuda.current_stream(device).synchronize()
        out_subjects = []
        for i, subject in enumerate(subjects):
            image = _tio.make_label_map(y_host[i], **copy.deepcopy(label_attributes))
            subject.add_image(image, "y_pred")
            out_subjects.append(_tio.enforce_consistent_affine(subject, "X"))
        return out_subjects, batch


class PatchPredict(Predictor):
    """ Creates predictions on patches and aggregates"""

    def __init__(
        self,
        image_names: Sequence[str] = ("X",),
        patch_batch_size: int = 16,
        patch_size=None,
        patch_overlap=(0, 0, 0),
        padding_mode: Union[str, float, None] = None,
        overlap_mode: str = "average",
    ):
        self.image_names = image_names
        self.patch_batch_size = patch_batch_size
        self.patch_size = patch_size
        self.patch_overlap = patch_overlap
        self.padding_mode = padding_mode
        self.overlap_mode = overlap_mode

    # ------------------------------------------------------------------ device pipeline for one volume
    def predict_volume(self, model: nn.Module, volume: torch.Tensor, want_probs: bool = True,
                       want_labels: bool = True):
        """volume: fp32 (C, W, H, D) already on the device.  Returns (probs fp32 (C_out, W, H, D) or None,
        labels uint8 (W, H, D) or None), both on the device."""
        if not volume.is_cuda:
            raise RuntimeError("predict_volume expects the volume on the CUDA device")
        with _lib().on_device(volume):     # the caller's current device may be another GPU
            return self._predict_volume(model, volume, want_probs, want_labels)

    def _predict_volume(self, model, volume, want_probs, want_labels):
        lib = _lib()
        if self.overlap_mode not in ("average", "crop", "hann"):
            raise ValueError(f'Overlap mode must be "crop", "average" or "hann" but "{self.overlap_mode}" was passed')
        if not volume.is_cuda:
            raise RuntimeError("predict_volume expects the volume on the CUDA device")
        in_dtype = volume.dtype          # 'auto' precision follows the caller's dtype (a bf16 volume -> bf16 path)
        volume = volume.detach().to(torch.float32).contiguous()
        grid = PatchGrid(volume.shape[1:], self.patch_size, self.patch_overlap, self.padding_mode)
        p0, p1, p2 = grid.patch_size
        device = volume.device
        native = isinstance(model, _NativeForward) and not isinstance(model, StochasticMatrix)
        compiled = None
        if native:
            if model.training:
                raise NotImplementedError("PatchPredict needs model.eval() (inference path only)")
            precision = _engine._resolve_precision(model, volume if in_dtype == torch.float32 else volume.new_empty(0, dtype=in_dtype))
            compiled = _engine.compiled_for(model, precision, device)
            if compiled.plan.out_scale != 0:
                raise RuntimeError("PatchPredict needs a network whose output extent equals its input extent")
        batch_size = self.patch_batch_size
        if native and DEVICE_BATCH[0] > batch_size:
            per_patch = compiled.workspace_bytes(1, p0, p1, p2)
            free = torch.cuda.mem_get_info(device)[0]
            fit = max(int(0.25 * free // max(per_patch, 1)), 1)
            batch_size = max(batch_size, min(DEVICE_BATCH[0], fit, len(grid.locations)))
        out = None
        windows = [w.to(device) for w in grid.hann_windows()] if self.overlap_mode == "hann" else None
        for locations in grid.batches(batch_size):
            b = len(locations)
            if native:
                in_buf = compiled.input_buffer(b, p0, p1, p2)
                lib.grid_extract(volume, locations, grid.border, grid.pad_mode_code, grid.pad_value,
                                 in_buf.view(volume.shape[0]))
                y = compiled.run_blocked(b, p0, p1, p2)
            else:
                # arbitrary nn.Module (e.g. a TTA ensemble around native members): hand it NCDHW patches
                staging = lib.Blocked(b, (volume.shape[0] + 7) // 8, p0, p1, p2, torch.float32, device)
                lib.grid_extract(volume, locations, grid.border, grid.pad_mode_code, grid.pad_value,
                                 staging.view(volume.shape[0]))
                x = torch.empty((b, volume.shape[0], p0, p1, p2), dtype=torch.float32, device=device)
                lib.unpack_ncdhw(staging.view(volume.shape[0]), x)
                with torch.no_grad():
                    y = model(x)
                y = y.detach().to(torch.float32).contiguous()
            if out is None:
                out = torch.zeros((y.shape[1], *grid.padded_shape), dtype=torch.float32, device=device)
            if self.overlap_mode == "average":
                lib.overlap_add(out, y, locations)
            elif self.overlap_mode == "hann":
                # weighted overlap-add: patch * window, then the same owner-computes gather
                if p2 % 4:
                    raise NotImplementedError("'hann' aggregation needs a patch size whose last axis is a multiple of 4")
                lib.window_patches(y, windows, count=b)
                lib.overlap_add(out, y, locations)
            else:
                lib.overlap_crop(out, y, locations, [o // 2 for o in grid.patch_overlap], grid.volume_padded)
        counts = None
        if self.overlap_mode == "average":
            counts = [torch.tensor(c, dtype=torch.int32, device=device) for c in grid.axis_counts()]
        elif self.overlap_mode == "hann":
            lib.divide_separable(out, [v.to(device) for v in grid.axis_window_sums()])
        w, h, d = grid.spatial_shape
        probs = torch.empty((out.shape[0], w, h, d), dtype=torch.float32, device=device) if want_probs else None
        labels = torch.empty((w, h, d), dtype=torch.uint8, device=device) \
            if want_labels and out.shape[0] <= 256 else None
        lib.finalize(out, counts, grid.border, probs, None, labels)
        return probs, labels

    def predict(self, model, device, subjects, label_attributes=None):
        device = _require_cuda(device)
        label_attributes = {} if label_attributes is None else label_attributes
        out_subjects = []
        volumes_on_device = []
        pending = []
        with _lib().on_device(device):
            main = torch.cuda.current_stream(device)
            d2h = torch.cuda.Stream(device=device) if len(subjects) > 1 else main
            for subject in subjects:
                volume = subject["X"]["data"]
                volume_dev = volume.to(device, non_blocking=True)   # one H2D per subject (async when pinned)
                volumes_on_device.append(volume_dev)
                with torch.no_grad():
                    probs, labels = self.predict_volume(model, volume_dev)
                # D2H through pinned memory on a second stream: with several subjects the copy of subject i overlaps
                # the network of subject i + 1; ONE synchronisation at the end (the reference blocks per patch batch)
                probs_host = torch.empty(probs.shape, dtype=probs.dtype, pin_memory=True)
                if d2h is not main:
                    d2h.wait_stream(main)
                    probs.record_stream(d2h)
                with torch.cuda.stream(d2h):
                    probs_host.copy_(probs, non_blocking=True)
                pending.append((subject, probs_host, labels))
            d2h.synchronize()
        for subject, probs_host, labels in pending:
            image = _tio.make_label_map(probs_host, **copy.deepcopy(label_attributes))
            if labels is not None:
                # device label map + the identity of the probabilities it was computed from: add_evaluation_labels
                # uses the stash only while 'data' is still that very tensor (post-processing may replace it)
                image[LABELS_KEY] = (labels, probs_host.data_ptr(), probs_host._version)
            subject.add_image(image, "y_pred")
            out_subjects.append(_tio.enforce_consistent_affine(subject, "X"))
        # batch[name]: (S, C, W, H, D) on the device, as collate_subjects returns (utils/utils.py:75-85); the 'X'
        # volumes are already there, so they are stacked on the device instead of being uploaded a second time
        batch = {}
        for name in self.image_names:
            if name == "X":
                batch[name] = volumes_on_device[0][None] if len(subjects) == 1 else torch.stack(volumes_on_device)
            else:
                batch.update(collate_subjects(subjects, image_names=[name], device=device))
        preds = [subject["y_pred"]["data"] for subject in out_subjects]
        batch["y_pred"] = preds[0][None] if len(preds) == 1 else torch.stack(preds)
        return out_subjects, batch


def _argmax_labels(image) -> torch.Tensor:
    """(C, W, H, D) probabilities / one-hot -> (1, W, H, D) int64 on the CPU, computed on the device."""
    data = image["data"]
    if data.shape[0] == 1:
        return data.long()
    lib = _lib()
    stash = image.get(LABELS_KEY) if hasattr(image, "get") else None
    if stash is not None:
        labels, ptr, version = stash
        if data.data_ptr() == ptr and data._version == version:
            return labels.to(torch.int64).cpu()[None]
    device = data.device if data.is_cuda else torch.device("cuda", torch.cuda.current_device())
    with lib.on_device(device):
        probs = data.detach().to(device=device, dtype=torch.float32).contiguous()
        labels = torch.empty(probs.shape[1:], dtype=torch.int64, device=device)
        lib.argmax(probs, labels, None)
    return labels.cpu()[None]


_DEVICE_ARGMAX = []


def _device_argmax_class():
    """``CustomArgMax`` (reference transforms/custom_label_transforms.py:253-272) with the argmax on the device: same
    constructor, history entry and inverse, so that it can stand in for it inside an inverted transform history."""
    if not _DEVICE_ARGMAX:
        from .transforms import CustomArgMax

        class DeviceArgMax(CustomArgMax):
            def apply_transform(self, subject):
                for image in self.get_images(subject):
                    image.set_data(_argmax_labels(image))
                    image['one_hot'] = False
                    if LABELS_KEY in image:
                        del image[LABELS_KEY]
                return subject

        _DEVICE_ARGMAX.append(DeviceArgMax)
    return _DEVICE_ARGMAX[0]


def _swap_argmax(transform):
    """Replaces every CustomArgMax of a (nested) Compose by its device version."""
    from .transforms import CustomArgMax
    compose = type(transform)
    out = []
    for t in transform:
        if isinstance(t, compose):
            out.append(_swap_argmax(t))
        elif type(t) is CustomArgMax:
            out.append(_device_argmax_class()(num_classes=t.num_classes, **t.kwargs))
        else:
            out.append(t)
    return compose(out)


def _history_inverse(subject):
    """The evaluation transform of reference prediction.py:157-160: the subject's applied-transform history, filtered
    to the label-affecting transforms, inverted.  None when the subject carries no torchio history (the torchio-less
    stand-in subjects of bench.py / smoke())."""
    if not (_tio.HAVE_TORCHIO and hasattr(subject, "get_composed_history")):
        return None
    try:
        from torchio.transforms.preprocessing.label.label_transform import LabelTransform
        from .transforms import ConcatenateImages, CopyProperty, RenameProperty, filter_transform
    except ImportError:          # stand-alone package: the reference's transform classes are not present
        return None
    transform = subject.get_composed_history()
    label_transform = filter_transform(transform, include_types=[LabelTransform, CopyProperty, RenameProperty,
                                                                 ConcatenateImages])
    return _swap_argmax(label_transform.inverse(warn=False))


def add_evaluation_labels(subjects: Sequence[Any]):
    """Adds 'y_pred_eval' / 'y_eval' next to 'y_pred' / 'y' (reference prediction.py:155-170).

    With torchio subjects (the reference's callers) this is the reference's algorithm: the inverse of the
    label-transform part of the subject's history is applied to ``Subject({'y': image})`` -- remaps, merges, copies and
    renames run as the reference's own transform classes; the ``CustomArgMax`` that inverts ``CustomOneHot`` runs on
    the device (``b200seg_argmax``, or the uint8 label map ``PatchPredict`` already produced there).
    Without torchio (bench.py / smoke() stand-in subjects, which have no history) multi-channel maps are argmaxed --
    the inverse of the one-hot targets every shipped config uses (int64, (1, W, H, D), ties -> lowest index)."""
    for subject in subjects:
        evaluation_transform = _history_inverse(subject)
        for name, eval_name in (("y_pred", "y_pred_eval"), ("y", "y_eval")):
            if name not in subject:
                continue
            source = subject[name]
            if evaluation_transform is not None:
                image = evaluation_transform(_tio.Subject({'y': source})).get_first_image()
                if LABELS_KEY in image:
                    del image[LABELS_KEY]
                subject.add_image(image, eval_name)
                continue
            attributes = {k: copy.deepcopy(v) for k, v in dict(source).items()
                          if k not in ("data", "affine", LABELS_KEY, "tensor", "path", "type", "stem")}
            attributes["one_hot"] = False
            image = _tio.make_label_map(_argmax_labels(source), affine=source["affine"], **attributes)
            subject.add_image(image, eval_name)
