#!/usr/bin/env python
"""Achieved HBM bandwidth of the streaming kernels of the path (extraction, overlap-add, finalize+argmax, argmax,
confusion) at config-2 / config-3 scale.  Timed on the device inside a CUDA graph (10 launches, an L2 flush before
each -- a 256 MB READ, so that no dirty lines are left behind -- flush time subtracted; see ``timed``).  Prints one line per kernel: algorithmic bytes, time, GB/s and
the fraction of the measured copy bandwidth (MEASURED_PEAKS.json)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "segmentation-pipeline_b200")]

import torch  # noqa: E402

import b200seg  # noqa: E402
from segmentation_pipeline.grid import PatchGrid  # noqa: E402


ONCE = "--once" in sys.argv      # under ncu: one warm-up + one measured launch per kernel (ncu reports the durations)


def timed(fn, flush, reps=10):
    """ms per launch.  The kernels here last 20-200 us, less than a Python + ctypes call costs when the GPU is idle, so
    CUDA events around a single eager call would time the host.  Instead ``reps`` launches, each preceded by an L2
    flush, are captured into ONE CUDA graph; the same graph with the flushes only is timed too and
    subtracted."""
    if ONCE:
        fn()
        torch.cuda.synchronize()
        flush.zero_()
        torch.cuda.nvtx.range_push("measured")
        fn()
        torch.cuda.synchronize()
        torch.cuda.nvtx.range_pop()
        return float("nan")

    sink = torch.zeros(1, dtype=torch.int64, device=flush.device)

    def capture(with_kernel):
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for _ in range(reps):
                # flush by READING 256 MB: the lines left in L2 are clean.  (A memset leaves ~126 MB of dirty lines whose
                # write-back is then charged to the kernel under test -- ~40 us, half the time of the small kernels.)
                sink.add_(flush.view(torch.int64).sum())
                if with_kernel:
                    fn()
        return g

    fn()
    torch.cuda.synchronize()
    both, only_flush = capture(True), capture(False)
    best = []
    for g in (both, only_flush):
        t = 1e9
        for _ in range(5):
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            g.replay()
            e.record()
            torch.cuda.synchronize()
            t = min(t, s.elapsed_time(e))
        best.append(t)
    return (best[0] - best[1]) / reps


def main():
    peak = 6554.9
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        peak = json.load(open(p)).get("hbm_gbs", peak)
    dev = "cuda"
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    rows = []

    # ---- config 2 geometry: 2 x 256 x 256 x 192, patch 96, overlap 48, edge padding, 48 patches per batch
    vol = torch.randn(2, 256, 256, 192, device=dev)
    grid = PatchGrid(vol.shape[1:], 96, 48, "edge")
    locs = grid.locations[:48]
    buf = b200seg.Blocked(48, 1, 96, 96, 96, torch.bfloat16, dev)
    t = timed(lambda: b200seg.grid_extract(vol, locs, grid.border, 1, 0.0, buf.view(2)), flush)
    nbytes = 48 * 96 ** 3 * 16 + 48 * 2 * 96 ** 3 * 4       # bf16 chunk written (16 B/voxel) + fp32 gathered
    rows.append(("grid_extract (48 patches, bf16 blocked out)", nbytes, t))

    patches = torch.rand(48, 2, 96, 96, 96, device=dev)
    out = torch.zeros((2, *grid.padded_shape), device=dev)
    t = timed(lambda: b200seg.overlap_add(out, patches, locs), flush)
    # bounding box of the first 48 locations
    bb = [min(l[i] for l in locs) for i in range(3)] + [max(l[i + 3] for l in locs) for i in range(3)]
    box = (bb[3] - bb[0]) * (bb[4] - bb[1]) * (bb[5] - bb[2])
    nbytes = patches.numel() * 4 + 2 * 2 * box * 4          # every patch read once + accumulator read & written
    rows.append(("overlap_add (48 patches, fp32)", nbytes, t))

    counts = [torch.tensor(c, dtype=torch.int32, device=dev) for c in grid.axis_counts()]
    labels = torch.empty(vol.shape[1:], dtype=torch.uint8, device=dev)
    probs = torch.empty_like(vol)
    t = timed(lambda: b200seg.finalize(out, counts, grid.border, None, None, labels), flush)
    v = vol[0].numel()
    rows.append(("finalize -> uint8 labels (config 2)", 2 * v * 4 + v, t))
    t = timed(lambda: b200seg.finalize(out, counts, grid.border, probs, None, labels), flush)
    rows.append(("finalize -> probs + labels (config 2)", 2 * v * 4 + 2 * v * 4 + v, t))

    # ---- config 3 scale: 10 classes, 224^3
    p10 = torch.rand(10, 224, 224, 224, device=dev)
    lab64 = torch.empty((224, 224, 224), dtype=torch.int64, device=dev)
    lab8 = torch.empty((224, 224, 224), dtype=torch.uint8, device=dev)
    v3 = 224 ** 3
    t = timed(lambda: b200seg.argmax(p10, None, lab8), flush)
    rows.append(("argmax 10 x 224^3 -> uint8", 10 * v3 * 4 + v3, t))
    t = timed(lambda: b200seg.argmax(p10, lab64, None), flush)
    rows.append(("argmax 10 x 224^3 -> int64 (reference dtype)", 10 * v3 * 4 + 8 * v3, t))

    # ---- confusion: cohort of 64 config-3 label maps in one buffer (uint8) and one map in int64
    a = torch.randint(0, 10, (64 * v3 // 4,), dtype=torch.uint8, device=dev)    # 16 volumes worth, 180 MB each side
    b = torch.randint(0, 10, (64 * v3 // 4,), dtype=torch.uint8, device=dev)
    cm = torch.zeros((10, 10), dtype=torch.int64, device=dev)
    t = timed(lambda: b200seg.confusion(a, b, 10, cm), flush)
    rows.append(("confusion 10 classes, uint8, 180 M voxels", 2 * a.numel(), t))
    a2 = torch.randint(0, 2, (256 * 256 * 192,), dtype=torch.uint8, device=dev)
    b2 = torch.randint(0, 2, (256 * 256 * 192,), dtype=torch.uint8, device=dev)
    cm2 = torch.zeros((2, 2), dtype=torch.int64, device=dev)
    t = timed(lambda: b200seg.confusion(a2, b2, 2, cm2), flush)
    rows.append(("confusion 2 classes, uint8, config 2 (12.6 M voxels)", 2 * a2.numel(), t))
    # piecewise-constant label maps (what a segmentation looks like): blocks of 8^3 voxels share a label
    coarse = torch.randint(0, 10, (1, 1, 28, 28, 28), device=dev).float()
    sm_a = torch.nn.functional.interpolate(coarse, scale_factor=8, mode="nearest")[0, 0].to(torch.uint8).contiguous()
    sm_b = torch.roll(sm_a, shifts=(3, 2, 1), dims=(0, 1, 2)).contiguous()
    t = timed(lambda: b200seg.confusion(sm_a, sm_b, 10, cm), flush)
    rows.append(("confusion 10 classes, uint8, 224^3 piecewise-constant", 2 * sm_a.numel(), t))
    a64, b64 = a[: v3].long(), b[: v3].long()
    t = timed(lambda: b200seg.confusion(a64, b64, 10, cm), flush)
    rows.append(("confusion 10 classes, int64, 224^3", 2 * 8 * v3, t))

    if ONCE:
        print(json.dumps([{"name": n, "bytes": b} for n, b, _ in rows]))
        return
    print(f"{'kernel':52s} {'MB':>9s} {'ms':>8s} {'GB/s':>8s} {'of measured copy peak':>22s}")
    for name, nbytes, ms in rows:
        gbs = nbytes / ms / 1e6
        print(f"{name:52s} {nbytes / 1e6:9.1f} {ms:8.3f} {gbs:8.0f} {gbs / peak:21.1%}")


if __name__ == "__main__":
    main()
