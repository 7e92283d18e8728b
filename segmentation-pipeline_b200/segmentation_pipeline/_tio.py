"""torchio access for the predictors.  When torchio is importable it is used as is (the reference's callers
pass real ``tio.Subject`` objects).  This image has no torchio, so a minimal stand-in with the handful of
members the hot path touches is provided; it is NOT a torchio re-implementation."""
from __future__ import annotations

import copy

import numpy as np
import torch

try:  # pragma: no cover - depends on the environment
    import torchio as tio
    HAVE_TORCHIO = True
except Exception:  # noqa: BLE001
    tio = None
    HAVE_TORCHIO = False


class _Image(dict):
    """dict-like image: ``image['data']`` / ``image.data`` (C, W, H, D) tensor, ``affine``, free attributes."""

    def __init__(self, tensor=None, affine=None, **attributes):
        super().__init__()
        self["data"] = tensor
        self["affine"] = np.eye(4) if affine is None else affine
        self.update(attributes)

    @property
    def data(self):
        return self["data"]

    @property
    def affine(self):
        return self["affine"]

    @affine.setter
    def affine(self, value):
        self["affine"] = value

    def set_data(self, tensor):
        self["data"] = tensor

    @property
    def spatial_shape(self):
        return tuple(self["data"].shape[1:])


class _ScalarImage(_Image):
    pass


class _LabelMap(_Image):
    pass


class _Subject(dict):
    def add_image(self, image, name):
        self[name] = image

    def get_images_dict(self, intensity_only=True, include=None, exclude=None):
        out = {}
        for k, v in self.items():
            if not isinstance(v, _Image):
                continue
            if intensity_only and isinstance(v, _LabelMap):
                continue
            if include is not None and k not in include:
                continue
            if exclude is not None and k in exclude:
                continue
            out[k] = v
        return out

    def get_first_image(self):
        return next(iter(self.get_images_dict(intensity_only=False).values()))

    @property
    def spatial_shape(self):
        return self.get_first_image().spatial_shape


# With torchio importable the predictors hand real tio objects to the caller (the reference's trainer, transforms and
# evaluators expect them); the stand-ins above only serve torchio-less scripts (bench.py, smoke()).
if HAVE_TORCHIO:
    Image, ScalarImage, LabelMap, Subject = tio.Image, tio.ScalarImage, tio.LabelMap, tio.Subject
else:
    Image, ScalarImage, LabelMap, Subject = _Image, _ScalarImage, _LabelMap, _Subject


def make_label_map(tensor: torch.Tensor, **attributes):
    if HAVE_TORCHIO:
        return tio.LabelMap(tensor=tensor, **attributes)
    return LabelMap(tensor=tensor, **attributes)


def enforce_consistent_affine(subject, source_image_name="X"):
    """``EnforceConsistentAffine(source_image_name)(subject)`` of the reference
    (transforms/enforce_consistent_affine.py:14-29): every other image takes the source image's affine."""
    if HAVE_TORCHIO:
        try:    # overlay mode: the reference's own transform (returns a copy and records itself in the history)
            from .transforms import EnforceConsistentAffine
            return EnforceConsistentAffine(source_image_name=source_image_name)(subject)
        except ImportError:
            pass
    if source_image_name not in subject:
        return subject
    source = subject[source_image_name]
    images = subject.get_images_dict(intensity_only=False)
    for name, image in images.items():
        if name == source_image_name:
            continue
        image.affine = copy.deepcopy(source.affine) if not HAVE_TORCHIO else source.affine
    return subject
