"""Drop-in proof: the UNMODIFIED reference package -- its own ``__init__.py``,
``segmentation_trainer.py``, ``SubjectFolder``, filters, transforms, data-loader factories, criterion -- imported through
``b200_overlay.install`` with ``prediction`` / ``models`` / the count evaluators served by this repo, driven for one
trainer iteration by tests/dropin_driver.py in a fresh process.

CPU part (this file, ``-m "not gpu"``): the wiring.  Every module resolves to the intended tree, the reference's
training step runs, and the validation branch reaches the b200 ``PatchPredict.predict``, which refuses a CPU device
loudly (no fallback).  GPU part: tests/test_gpu_dropin.py runs the same driver on the B200 and compares the
evaluator's output with the oracle."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def reference_root():
    """The reference checkout: /root/reference in the authoring container, else the copy ``build()`` leaves under
    baseline/_ref (git-ignored; it travels to the GPU box with the snapshot)."""
    for root in ("/root/reference", os.path.join(ROOT, "baseline", "_ref")):
        if os.path.isfile(os.path.join(root, "segmentation_pipeline", "__init__.py")):
            return root
    return None


def run_driver(device, precision="fp32", timeout=900):
    root = reference_root()
    if root is None:
        pytest.skip("no reference checkout on this box (/root/reference or baseline/_ref): drop-in run skipped")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "dropin_driver.py"), "--reference", root,
                        "--device", device, "--precision", precision], capture_output=True, text=True, timeout=timeout)
    lines = [l for l in r.stdout.splitlines() if l.startswith("DROPIN ")]
    assert lines, f"driver produced no result (rc={r.returncode}):\n{r.stdout[-2000:]}\n{r.stderr[-4000:]}"
    return json.loads(lines[-1][7:])


def test_unmodified_trainer_reaches_b200_predictor():
    res = run_driver("cpu")
    served = res["served_from"]
    assert served["segmentation_pipeline"] == "reference"                          # the reference's own __init__.py
    assert served["segmentation_pipeline.segmentation_trainer"] == "reference"
    assert served["segmentation_pipeline.data_processing.subject_folder"] == "reference"
    assert served["segmentation_pipeline.transforms.custom_label_transforms"] == "reference"
    assert served["segmentation_pipeline.utils.torch_context"] == "reference"
    assert served["segmentation_pipeline.prediction"] == "b200"
    assert served["segmentation_pipeline.models.modular_unet"] == "b200"
    assert served["segmentation_pipeline.evaluators.segmentation_evaluator"] == "b200"
    assert served["segmentation_pipeline.post_processing"] == "b200"               # remove_holes & co on the device
    # the training step ran (stub predictor + the reference's criterion), the validation branch called our predictor,
    # and on a CPU device that is an error, not a silent fallback
    err = res["error"]
    assert err["type"] == "RuntimeError" and "CUDA device only" in err["message"]
    assert err["raised_in"][-3:] == ["segmentation_trainer.py:train", "prediction.py:predict",
                                     "prediction.py:_require_cuda"]


def test_overlay_resolution_rules(tmp_path):
    """File-by-file resolution on a toy reference tree: shadowed modules come from here, package __init__ files and
    everything else from the reference, b200-only modules are still found."""
    code = r'''
import os, sys, json
sys.path.insert(0, os.path.join(%r, "segmentation-pipeline_b200"))
ref = %r
pkg = os.path.join(ref, "segmentation_pipeline")
os.makedirs(os.path.join(pkg, "models")); os.makedirs(os.path.join(pkg, "evaluators"))
open(os.path.join(pkg, "__init__.py"), "w").write("MARK = 'ref-init'\nfrom . import extra\n")
open(os.path.join(pkg, "extra.py"), "w").write("WHO = 'ref-extra'\n")
open(os.path.join(pkg, "prediction.py"), "w").write("WHO = 'ref-prediction'\n")
open(os.path.join(pkg, "models", "__init__.py"), "w").write("WHO = 'ref-models'\n")
open(os.path.join(pkg, "evaluators", "__init__.py"), "w").write("WHO = 'ref-evaluators'\n")
import b200_overlay
f = b200_overlay.install(ref)
import segmentation_pipeline as sp, segmentation_pipeline.grid, segmentation_pipeline.evaluators
out = {"init": sp.MARK, "extra": sp.extra.WHO, "grid": hasattr(sp.grid, "PatchGrid"),
       "evaluators": sp.evaluators.WHO,
       "prediction_file": f.find_spec("segmentation_pipeline.prediction").origin,
       "models_file": f.find_spec("segmentation_pipeline.models").origin}
b200_overlay.uninstall()
out["uninstalled"] = "segmentation_pipeline" not in sys.modules
print("RES " + json.dumps(out))
''' % (ROOT, str(tmp_path))
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=120)
    line = [l for l in r.stdout.splitlines() if l.startswith("RES ")]
    assert line, r.stderr[-3000:]
    out = json.loads(line[0][4:])
    assert out["init"] == "ref-init" and out["extra"] == "ref-extra" and out["evaluators"] == "ref-evaluators"
    assert out["grid"] is True and out["uninstalled"] is True
    assert "segmentation-pipeline_b200" in out["prediction_file"] and "segmentation-pipeline_b200" in out["models_file"]
