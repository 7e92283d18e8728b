#!/usr/bin/env python
"""Times one trainer iteration (segmentation_trainer.py:162-180) of the msseg2 network on the device:
model.train() forward, HybridLogisticDiceLoss, backward, SGD step.
Usage: profile_train.py [batch] [patch] [steps] [fp32|bf16] [msseg2|nested]."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "segmentation-pipeline_b200")]

import torch  # noqa: E402

import bench  # noqa: E402
from segmentation_pipeline.criterions.hybrid_logistic_dice_loss import HybridLogisticDiceLoss  # noqa: E402


def main():
    batch = int(sys.argv[1]) if len(sys.argv) > 1 else 4
    patch = int(sys.argv[2]) if len(sys.argv) > 2 else 96
    steps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
    from segmentation_pipeline.models import set_precision
    set_precision(sys.argv[4] if len(sys.argv) > 4 else "fp32")
    nested = len(sys.argv) > 5 and sys.argv[5] == "nested"
    if nested:      # dmri_hippo-style: NestedResUNet(3 -> 2, 40 filters, dropout 0.2) on 3 x 96 x 88 x 24 volumes
        from segmentation_pipeline import models as M
        model = M.NestedResUNet(3, 2, 40, dropout_p=0.2).cuda().train()
    else:
        model = bench.build_model().cuda().train()
    opt = torch.optim.SGD(model.parameters(), lr=1e-3, momentum=0.95)
    criterion = HybridLogisticDiceLoss(logistic_class_weights=None if nested else [1, 100])
    g = torch.Generator().manual_seed(0)
    shape = (96, 88, 24) if nested else (patch, patch, patch)
    x = torch.randn(batch, 3 if nested else 2, *shape, generator=g).cuda()
    labels = (torch.rand(batch, *shape, generator=g) < 0.05).long()
    y = torch.nn.functional.one_hot(labels, 2).movedim(-1, 1).float().cuda()
    times = {"forward": 0.0, "loss": 0.0, "backward": 0.0, "step": 0.0}
    for it in range(steps + 1):
        marks = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
        marks[0].record()
        probs = model(x)
        marks[1].record()
        loss = criterion(probs, y)["loss"]
        marks[2].record()
        opt.zero_grad()
        loss.backward()
        marks[3].record()
        opt.step()
        marks[4].record()
        torch.cuda.synchronize()
        if it:                                   # first iteration = warm-up
            for k, (a, b) in zip(times, zip(marks[:-1], marks[1:])):
                times[k] += a.elapsed_time(b) / steps
        print(f"iter {it}: loss {float(loss):.6f}  peak mem {torch.cuda.max_memory_allocated() / 2**30:.1f} GiB", flush=True)
    total = sum(times.values())
    flop = 3 * 1529078.0 * 96 * 88 * 24 * batch if nested else 3 * 668.74e9 * batch * (patch / 96) ** 3
    print({k: round(v, 1) for k, v in times.items()}, f"total {total:.1f} ms/step; {flop / total / 1e9:.1f} TFLOP/s (3 x forward FLOPs)")


if __name__ == "__main__":
    main()
