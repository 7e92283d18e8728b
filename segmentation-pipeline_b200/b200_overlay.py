"""Drop-in installation: run the UNMODIFIED reference package with the hot path replaced by this one.

    import b200_overlay
    b200_overlay.install("/path/to/Segmentation-Pipeline")      # the checkout that contains segmentation_pipeline/
    import segmentation_pipeline                                  # the reference's own __init__.py runs
    from segmentation_pipeline import SegmentationTrainer, PatchPredict, ModularUNet, ...

Why an import hook.  Both trees are regular packages called ``segmentation_pipeline``; putting one of them ahead on
``sys.path`` hides the other completely (the reference's ``__init__.py:1-16`` star-imports ``segmentation_trainer``,
``transforms``, ``loggers`` ..., none of which this package rebuilds).  ``install`` therefore registers a meta-path
finder that resolves every ``segmentation_pipeline.*`` module FILE BY FILE:

  * the modules of the hot path come from this package (``SHADOW``): ``prediction``, ``models.*``, the two
    evaluators whose voxel work is the device confusion histogram, the criterion, and ``post_processing``;
  * every other module, and every package ``__init__``, is the reference's own file, executed unmodified -- so
    ``segmentation_trainer.py``, the transforms, ``TorchContext``, the data loaders and the ``research/*`` configs see
    exactly the namespace they were written against (``from .evaluators import *``, ``from segmentation_pipeline
    import *``);
  * modules that exist only here (``grid``, ``_tio``, ``distributed``, ``models._engine``, ``models._plan``) are
    found here.

Nothing of the reference is copied or patched; classes keep their import paths, so checkpoints that pickle
``segmentation_pipeline.models.modular_unet.ModularUNet`` (reference ``utils/torch_context.py:131``) load.
"""
from __future__ import annotations

import importlib
import importlib.abc
import importlib.util
import os
import sys
from typing import Optional

PACKAGE = "segmentation_pipeline"
HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), PACKAGE)

# modules replaced by this package (relative to ``segmentation_pipeline``)
SHADOW = frozenset({
    "prediction",
    "models",                       # the package __init__ too (same exports + set_precision / get_precision)
    "models.components", "models.modular_unet", "models.nested_residual_unet", "models.ensemble", "models.utils",
    "evaluators.segmentation_evaluator", "evaluators.label_map_evaluator",
    "evaluators.instance_segmentation_evaluator",     # also drops the reference module's import-time skimage dependency
    "criterions.hybrid_logistic_dice_loss",
    "post_processing",              # label clean-up after inference; also drops its import-time skimage dependency
})


def _locate(root: str, rel):
    """-> (file, is_package) of module ``rel`` (tuple of name parts) under ``root``, or (None, False)."""
    base = os.path.join(root, *rel)
    init = os.path.join(base, "__init__.py")
    if os.path.isfile(init):
        return init, True
    if rel and os.path.isfile(base + ".py"):
        return base + ".py", False
    return None, False


class OverlayFinder(importlib.abc.MetaPathFinder):
    def __init__(self, reference_pkg: str, overlay_pkg: str = HERE):
        self.reference_pkg = reference_pkg
        self.overlay_pkg = overlay_pkg
        self.resolved = {}                      # fullname -> file, for inspection / tests

    def find_spec(self, fullname, path=None, target=None):
        if fullname != PACKAGE and not fullname.startswith(PACKAGE + "."):
            return None
        rel = tuple(fullname.split(".")[1:])
        ref_file, ref_pkg = _locate(self.reference_pkg, rel)
        own_file, own_pkg = _locate(self.overlay_pkg, rel)
        if ".".join(rel) in SHADOW and own_file is not None:
            file, is_pkg = own_file, own_pkg
        elif ref_file is not None:
            file, is_pkg = ref_file, ref_pkg
        elif own_file is not None:
            file, is_pkg = own_file, own_pkg
        else:
            return None
        locations = None
        if is_pkg:
            # children are resolved by this finder again, so the list only has to be non-empty and truthful
            locations = [d for d in (os.path.join(self.overlay_pkg, *rel), os.path.join(self.reference_pkg, *rel))
                         if os.path.isdir(d)]
        self.resolved[fullname] = file
        return importlib.util.spec_from_file_location(fullname, file, submodule_search_locations=locations)


_installed: Optional[OverlayFinder] = None


def install(reference_root: str) -> OverlayFinder:
    """Registers the overlay for the reference checkout at ``reference_root`` (the directory that contains
    ``segmentation_pipeline/``).  Call it before anything imports ``segmentation_pipeline``; modules of that name that
    are already loaded are dropped so that the package is re-resolved through the overlay."""
    global _installed
    reference_pkg = os.path.join(os.path.abspath(reference_root), PACKAGE)
    if not os.path.isfile(os.path.join(reference_pkg, "__init__.py")):
        raise FileNotFoundError(f"{reference_pkg}/__init__.py not found: pass the reference checkout's root")
    if os.path.samefile(reference_pkg, HERE):
        raise ValueError("reference_root points at the b200 package itself")
    uninstall()
    for name in [n for n in sys.modules if n == PACKAGE or n.startswith(PACKAGE + ".")]:
        del sys.modules[name]
    _installed = OverlayFinder(reference_pkg)
    sys.meta_path.insert(0, _installed)
    importlib.invalidate_caches()
    return _installed


def uninstall() -> None:
    global _installed
    if _installed is not None:
        if _installed in sys.meta_path:
            sys.meta_path.remove(_installed)
        for name in [n for n in sys.modules if n == PACKAGE or n.startswith(PACKAGE + ".")]:
            del sys.modules[name]
        _installed = None


def installed() -> Optional[OverlayFinder]:
    return _installed
