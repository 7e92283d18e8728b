"""ORACLE helper (test infrastructure only): import the reference's own Python modules in-process.

``/root/reference`` exists only in the authoring container, never on the GPU box, so nothing that runs under
``-m gpu``, ``smoke()`` or ``bench.py`` may call this at run time -- it is used by ``oracle/make_golden.py``
to produce the committed fixtures under ``tests/golden/`` and by CPU tests that are skipped when the
reference tree is absent.

``segmentation_pipeline/__init__.py:1-16`` star-imports sub-packages that need torchio / skimage /
matplotlib, none of which are installed; the technique here registers an empty parent package whose
``__path__`` points at the reference tree so that ``segmentation_pipeline.models`` (torch only) and the
evaluator arithmetic can be imported without executing that ``__init__``.  Nothing is copied.
"""
from __future__ import annotations

import importlib
import os
import sys
import types
from typing import Sequence

REFERENCE_ROOT = os.environ.get("B200SEG_REFERENCE_ROOT", "/root/reference")
_PKG = "segmentation_pipeline"


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, _PKG, "models"))


class _RefModules:
    """Context manager: temporarily binds ``segmentation_pipeline*`` in ``sys.modules`` to the reference tree
    (our own drop-in package uses the same import name), restoring the previous bindings on exit."""

    def __enter__(self):
        self._saved = {k: v for k, v in sys.modules.items() if k == _PKG or k.startswith(_PKG + ".") or
                       k == "torchio"}
        for k in self._saved:
            del sys.modules[k]
        root = types.ModuleType(_PKG)
        root.__path__ = [os.path.join(REFERENCE_ROOT, _PKG)]
        sys.modules[_PKG] = root
        utils = types.ModuleType(_PKG + ".utils")
        # the three helpers the imported modules take from segmentation_pipeline/utils/utils.py:19-36
        utils.is_sequence = lambda x: isinstance(x, Sequence) and not isinstance(x, str)
        utils.as_list = lambda x: [] if x is None else (list(x) if utils.is_sequence(x) else [x])
        utils.auto_str = lambda obj: type(obj).__name__
        sys.modules[_PKG + ".utils"] = utils
        if "torchio" not in sys.modules:
            tio = types.ModuleType("torchio")
            tio.Subject = dict
            tio.SubjectsDataset = list
            sys.modules["torchio"] = tio
            self._stub_tio = True
        else:
            self._stub_tio = False
        return self

    def load(self, name: str):
        return importlib.import_module(f"{_PKG}.{name}")

    def __exit__(self, *exc):
        for k in [k for k in sys.modules if k == _PKG or k.startswith(_PKG + ".")]:
            del sys.modules[k]
        if self._stub_tio:
            sys.modules.pop("torchio", None)
        sys.modules.update(self._saved)
        return False


def load_reference_models():
    """Returns the reference ``segmentation_pipeline.models`` module (classes stay usable after return)."""
    if not reference_available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    with _RefModules() as ctx:
        return ctx.load("models")


def load_reference_evaluators():
    """Returns (SegmentationEvaluator, LabelMapEvaluator, LabeledTensor) classes of the reference."""
    if not reference_available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    with _RefModules() as ctx:
        ev = types.ModuleType(_PKG + ".evaluators")
        ev.__path__ = [os.path.join(REFERENCE_ROOT, _PKG, "evaluators")]
        sys.modules[_PKG + ".evaluators"] = ev
        seg = ctx.load("evaluators.segmentation_evaluator")
        lab = ctx.load("evaluators.label_map_evaluator")
        lt = ctx.load("evaluators.labeled_tensor")
        return seg.SegmentationEvaluator, lab.LabelMapEvaluator, lt.LabeledTensor


def load_reference_criterion():
    if not reference_available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    with _RefModules() as ctx:
        cr = types.ModuleType(_PKG + ".criterions")
        cr.__path__ = [os.path.join(REFERENCE_ROOT, _PKG, "criterions")]
        sys.modules[_PKG + ".criterions"] = cr
        return ctx.load("criterions.hybrid_logistic_dice_loss").HybridLogisticDiceLoss
