"""TEST INFRASTRUCTURE -- a minimal stand-in for the third-party ``torchio`` package (pinned 0.18.45 by the
reference, research/msseg2/competition/docker-requirements.txt:44), which is not installed in this image and cannot be
fetched (no network).  It exists so that tests can import and run the UNMODIFIED reference modules
(segmentation_trainer.py, transforms/, evaluators/, data_processing/ ...) against the b200 hot path
(tests/test_dropin*.py); it is put on ``sys.path`` by those tests only when the real torchio is not importable.

Scope: the members the reference touches at import time and on the prediction / evaluation path (SURVEY.md section 8
f1): Subject / Image / ScalarImage / LabelMap, Transform (history, include/exclude, copy, inverse), Compose,
LabelTransform / SpatialTransform / IntensityTransform / RandomTransform, Pad / Crop / CopyAffine, SubjectsDataset.
NOT provided: file I/O, resampling, augmentation, GridSampler / GridAggregator / Queue (the b200 PatchPredict does
not use them; calling them raises).  Behaviour follows torchio 0.18.x as recalled from its source; nothing here is
product code."""
from .constants import AFFINE, DATA, INTENSITY, LABEL, LOCATION, PATH, STEM, TYPE  # noqa: F401
from .data import GridAggregator, GridSampler, Image, LabelMap, Queue, ScalarImage, Subject, SubjectsDataset  # noqa: F401
from .transforms import (Compose, CopyAffine, Crop, IntensityTransform, LabelTransform, Pad, RandomTransform,  # noqa: F401
                         Resample, SpatialTransform, Transform)
from . import data, transforms, typing  # noqa: F401

__version__ = "0.18.45+b200shim"
