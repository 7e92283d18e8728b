#!/usr/bin/env python
"""Host-side breakdown of one PatchPredict.predict call on the bench workload (config 2): H2D, device pipeline,
D2H, subject bookkeeping.  Developer tool."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "segmentation-pipeline_b200")]

import torch  # noqa: E402

import bench  # noqa: E402
from segmentation_pipeline import _tio  # noqa: E402
from segmentation_pipeline.models import set_precision  # noqa: E402
from segmentation_pipeline.prediction import PatchPredict  # noqa: E402


def main():
    set_precision("bf16")
    dev = torch.device("cuda", 0)
    model = bench.build_model().to(dev)
    pred = PatchPredict(patch_batch_size=bench.PATCH_BATCH, patch_size=96, patch_overlap=48, padding_mode="edge")
    vol_host = bench.synthetic_volume(0).pin_memory()

    def sync():
        torch.cuda.synchronize()
        return time.perf_counter()

    for it in range(5):
        t0 = sync()
        v = vol_host.to(dev, non_blocking=True)
        t1 = sync()
        with torch.no_grad():
            probs, labels = pred.predict_volume(model, v)
        t2 = sync()
        host = torch.empty(probs.shape, dtype=probs.dtype, pin_memory=True)
        t3 = sync()
        host.copy_(probs, non_blocking=True)
        t4 = sync()
        subject = _tio.Subject(X=_tio.ScalarImage(tensor=vol_host), name="e2e")
        t5 = sync()
        out, batch = pred.predict(model, dev, [subject])
        t6 = sync()
        print(f"iter {it}: h2d {1e3*(t1-t0):.1f}  device {1e3*(t2-t1):.1f}  pinned alloc {1e3*(t3-t2):.1f}  d2h {1e3*(t4-t3):.1f}  "
              f"subject {1e3*(t5-t4):.1f}  full predict() {1e3*(t6-t5):.1f} ms")


if __name__ == "__main__":
    main()
