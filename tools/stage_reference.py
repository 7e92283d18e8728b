#!/usr/bin/env python
"""Developer tool: temporarily stage the UNMODIFIED reference package under baseline/_ref (git-ignored) so that ONE
gpurun call can run tests/test_gpu_dropin.py -- the reference's own segmentation_trainer.py against the b200 hot path --
on a B200 (the GPU box has no /root/reference).

    python tools/stage_reference.py            # copy /root/reference/segmentation_pipeline -> baseline/_ref/
    gpurun -- 'python -m pytest tests/test_gpu_dropin.py -m gpu -q -s'
    python tools/stage_reference.py --clean    # remove it again

The copy is never committed and is removed after the run; without it the GPU drop-in test SKIPS (and says so), and
the CPU half of the proof (tests/test_dropin.py, which reads /root/reference in the authoring container) still runs."""
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = "/root/reference/segmentation_pipeline"
DST = os.path.join(ROOT, "baseline", "_ref")

if __name__ == "__main__":
    if os.path.isdir(DST):
        shutil.rmtree(DST)
    if "--clean" in sys.argv:
        print("removed", DST)
    else:
        shutil.copytree(SRC, os.path.join(DST, "segmentation_pipeline"),
                        ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
        print("staged", SRC, "->", DST)
