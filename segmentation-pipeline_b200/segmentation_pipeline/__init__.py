"""B200-native drop-in for the sliding-window inference path of ``segmentation_pipeline``.

Only the hot path is rebuilt: ``models`` (same constructors / state_dict keys, native
forward), ``prediction`` (Predictor, StandardPredict, PatchPredict, add_evaluation_labels), the evaluators'
count / Dice reduction, ``CustomArgMax`` and the label post-processing (``post_processing``).  Everything is lazy: importing the package does not need
torchio, a GPU or the shared library; calling a forward without libb200seg.so raises.
"""
from . import utils  # noqa: F401

__all__ = ["models", "prediction", "evaluators", "criterions", "transforms", "utils", "post_processing"]


def __getattr__(name):
    import importlib
    if name in ("models", "prediction", "evaluators", "transforms", "grid", "criterions", "distributed", "post_processing"):
        return importlib.import_module(f"{__name__}.{name}")
    if name in ("StandardPredict", "PatchPredict"):
        return getattr(importlib.import_module(f"{__name__}.prediction"), name)
    if name in ("sort_by_size", "keep_components", "remove_holes", "remove_small_components"):
        return getattr(importlib.import_module(f"{__name__}.post_processing"), name)
    raise AttributeError(name)
