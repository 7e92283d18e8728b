"""B200-native drop-in for the sliding-window inference path of ``segmentation_pipeline``.

Only the hot path is rebuilt: ``models`` (same constructors / state_dict keys, native
forward), ``prediction`` (Predictor, StandardPredict, PatchPredict, add_evaluation_labels), the evaluators'
count / Dice reduction and ``CustomArgMax``.  Everything is lazy: importing the package does not need
torchio, a GPU or the shared library; calling a forward without libb200seg.so raises.
"""
from . import utils  # noqa: F401

__all__ = ["models", "prediction", "evaluators", "criterions", "transforms", "utils"]


def __getattr__(name):
    import importlib
    if name in ("models", "prediction", "evaluators", "transforms", "grid", "criterions", "distributed"):
        return importlib.import_module(f"{__name__}.{name}")
    if name in ("StandardPredict", "PatchPredict"):
        return getattr(importlib.import_module(f"{__name__}.prediction"), name)
    raise AttributeError(name)
