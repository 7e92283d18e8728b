#!/usr/bin/env python
"""Per-layer timing of the msseg2 network on one batch of 96^3 patches (CUDA events around every op of the
plan).  Prints one line per op: name, output extent, ms, algorithmic TFLOP/s.  Developer tool."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "segmentation-pipeline_b200")]

import torch  # noqa: E402

import bench  # noqa: E402
from segmentation_pipeline.models import _engine, _plan, set_precision  # noqa: E402


def main():
    batch = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    only = sys.argv[3] if len(sys.argv) > 3 else None     # run just the ops whose name contains this (for ncu)
    only = None if only in (None, "", "-") else only
    precision = sys.argv[4] if len(sys.argv) > 4 else "bf16"      # "fp32": the CUDA-core precision path
    set_precision(precision)
    model = bench.build_model().cuda()
    compiled = _engine.compiled_for(model, precision, torch.device("cuda"))
    if only:
        compiled.calls = [c for c in compiled.calls if only in getattr(c, "name", "")]
        compiled.plan.buffers.setdefault("out", (1, 0))
    x = torch.randn(batch, 2, 96, 96, 96, device="cuda")
    with torch.no_grad():
        for _ in range(2):
            compiled.run(x)
        torch.cuda.synchronize()
        _engine.TRACE = []
        for _ in range(reps):
            compiled.run(x)
        torch.cuda.synchronize()
    trace, _engine.TRACE = _engine.TRACE, None
    per = len(trace) // reps
    total_ms, total_flop = 0.0, 0.0
    print(f"{'op':28s} {'mode':5s} {'cin':>4s} {'cout':>4s} {'ext':>12s} {'ms':>8s} {'TFLOP/s':>8s}")
    for i in range(per):
        call, (n, z, y, xx), _ = trace[i]
        ms = sum(trace[r * per + i][2][0].elapsed_time(trace[r * per + i][2][1]) for r in range(reps)) / reps
        total_ms += ms
        if isinstance(call, _engine._ConvCall):
            lvl = compiled.plan.buffers[call.src.buf][1]
            ez, ey, ex = z >> lvl, y >> lvl, xx >> lvl
            mode = call.mode if call.tc else (2 if call.transposed else (1 if call.stride == 2 else 0))
            taps = 27 if mode in (0, 3) else (64 if mode == 1 else 8)
            if mode == 1:
                ez, ey, ex = ez // 2, ey // 2, ex // 2
            if mode == 2:
                ez, ey, ex = ez * 2, ey * 2, ex * 2
            flop = 2.0 * n * ez * ey * ex * taps * call.src.c * call.cout
            total_flop += flop
            print(f"{call.name:28s} {['k3','down','up','k3t'][mode]:5s} {call.src.c:4d} {call.cout:4d} "
                  f"{f'{ez}x{ey}x{ex}':>12s} {ms:8.3f} {flop / ms / 1e9:8.1f}")
        else:
            print(f"{type(call).__name__:28s} {'':5s} {'':4s} {'':4s} {'':>12s} {ms:8.3f}")
    print(f"total {total_ms:.2f} ms per batch of {batch}; {total_flop / total_ms / 1e9:.1f} TFLOP/s (physical-channel FLOPs)")


if __name__ == "__main__":
    main()
