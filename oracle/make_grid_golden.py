"""ORACLE (test infrastructure only) -- writes tests/golden/grid_torchio.json, the known-answer vectors that pin
oracle/grid.py (and through it the CUDA sampler / aggregator kernels) to torchio's documented behaviour.

torchio is not installed and not vendored (pinned 0.18.45 by the reference,
research/msseg2/competition/docker-requirements.txt:44), so the vectors come from two sources that do NOT go through
oracle/grid.py:

1. torchio's OWN unit-test fixtures for this path, restated from its test-suite (torchio 0.18.x):
     tests/data/sampler/test_grid_sampler.py::TestGridSampler::test_locations
         subject image (10, 20, 30), patch (5, 20, 20), overlap (2, 0, 6) -> the six location rows below
     tests/data/inference/test_aggregator.py::TestAggregator::test_overlap_crop / test_overlap_average
         image ones(1, 1, 4, 4), patch (1, 3, 3), overlap (0, 2, 2); the four patches (j0, k0) = (0,0), (0,1), (1,0),
         (1,1) are multiplied by 0, 2, 4, 6 -> crop [[0,0,2,2],[0,0,2,2],[4,4,6,6],[4,4,6,6]],
         average [[0,1,1,2],[2,3,3,4],[2,3,3,4],[4,5,5,6]]
   Each is re-derived by hand in the comments next to it, so a mis-remembered digit would show.
2. Brute-force cases computed HERE, voxel by voxel from the definition in torchio's docstrings (GridSampler:
   "padding_mode ... the volume is padded with w/2 on each side", "patch_overlap must be even";
   GridAggregator: 'crop' = "the overlapping predictions will be cropped", 'average' = "the predictions in the
   overlapping areas will be averaged with equal weights"), with no code shared with oracle/grid.py: for every output
   voxel, enumerate the patches that cover it.

Run in the authoring container:  python oracle/make_grid_golden.py
"""
from __future__ import annotations

import itertools
import json
import os

import numpy as np

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "grid_torchio.json")


def starts_by_definition(size: int, patch: int, overlap: int):
    """Walk the axis with stride patch - overlap; the last patch is shifted back so that it ends at the border."""
    out, pos = [], 0
    while pos + patch < size:
        out.append(pos)
        pos += patch - overlap
    out.append(size - patch)
    return sorted(set(out))


def brute_force_case(shape, patch, overlap, padding_mode, mode, seed):
    """Per-voxel evaluation of the sliding window with a 'model' that multiplies patch n by (n + 1)."""
    rng = np.random.default_rng(seed)
    vol = rng.integers(-8, 9, size=(1, *shape)).astype(np.float32)          # small integers: sums are exact
    border = [o // 2 if padding_mode is not None else 0 for o in overlap]
    padded_shape = [s + 2 * b for s, b in zip(shape, border)]

    def padded_value(c, idx):
        src = []
        for i, b, s in zip(idx, border, shape):
            j = i - b
            if padding_mode == "edge":
                j = min(max(j, 0), s - 1)
            elif j < 0 or j >= s:
                return float(padding_mode)                                   # constant fill
            src.append(j)
        return float(vol[(c, *src)])

    axis = [starts_by_definition(ps, p, o) for ps, p, o in zip(padded_shape, patch, overlap)]
    locations = [list(ini) + [i + p for i, p in zip(ini, patch)] for ini in itertools.product(*axis)]
    result = np.zeros((1, *shape), np.float64)
    for idx in itertools.product(*[range(s) for s in shape]):
        pidx = [i + b for i, b in zip(idx, border)]
        total, count, assigned = 0.0, 0, None
        for n, loc in enumerate(locations):
            if not all(loc[a] <= pidx[a] < loc[a + 3] for a in range(3)):
                continue
            value = padded_value(0, pidx) * (n + 1)
            total += value
            count += 1
            # crop mode: the patch keeps the voxel unless it lies in its overlap//2 margin on a side that is not the
            # border of the (unpadded) sampled volume; with padding every side is trimmed.  torchio 0.18.45 then takes
            # the kept block from the CENTRE of the patch (GridAggregator.crop_batch: left = (patch - crop) / 2), which
            # is the voxel's own position when both sides are trimmed and is shifted by overlap//4 when only one is.
            keep, src = True, []
            for a in range(3):
                t = overlap[a] // 2
                lo_trim = t if (padding_mode is not None or loc[a] != 0) else 0
                hi_trim = t if (padding_mode is not None or loc[a + 3] != padded_shape[a]) else 0
                if not (loc[a] + lo_trim <= pidx[a] < loc[a + 3] - hi_trim):
                    keep = False
                src.append(loc[a] + (pidx[a] - (loc[a] + lo_trim)) + (lo_trim + hi_trim) // 2)
            if keep:
                assigned = padded_value(0, src) * (n + 1)                    # later patches overwrite earlier ones
        result[(0, *idx)] = total / count if mode == "average" else assigned
    return {"shape": list(shape), "patch": list(patch), "overlap": list(overlap), "padding_mode": padding_mode,
            "overlap_mode": mode, "patch_gain": "patch n (sorted order) is multiplied by n + 1",
            "volume": vol[0].astype(int).tolist(), "locations": locations,
            "output": result[0].astype(np.float32).astype(float).tolist()}


def main() -> None:
    golden = {
        "source": "torchio 0.18.x unit-test fixtures (restated) + brute-force cases; see oracle/make_grid_golden.py",
        # size 10, patch 5, overlap 2: stride 3 -> 0, 3, then 6 + 5 > 10 so the last patch is flush: 5.
        # size 20, patch 20: 0.   size 30, patch 20, overlap 6: stride 14 -> 0, then flush: 10.
        "torchio_test_locations": {
            "image": [10, 20, 30], "patch": [5, 20, 20], "overlap": [2, 0, 6],
            "locations": [[0, 0, 0, 5, 20, 20], [0, 0, 10, 5, 20, 30], [3, 0, 0, 8, 20, 20], [3, 0, 10, 8, 20, 30],
                          [5, 0, 0, 10, 20, 20], [5, 0, 10, 10, 20, 30]],
        },
        # 4 x 4 image of ones, 3 x 3 patches at (0,0), (0,1), (1,0), (1,1) scaled by 0, 2, 4, 6.
        # average: voxel (r, c) is covered by the patches with j0 in {r-2..r} & {0,1}, k0 likewise, e.g. (1, 1) by all
        #          four -> (0 + 2 + 4 + 6) / 4 = 3; (0, 1) by (0,0), (0,1) -> 1; (3, 3) only by (1,1) -> 6.
        # crop:    border 1 is trimmed on the inner sides only (no padding): patch (0,0) keeps rows 0-1 x cols 0-1, ...
        "torchio_test_aggregator": {
            "image": [1, 4, 4], "patch": [1, 3, 3], "overlap": [0, 2, 2],
            "patch_values": {"0,0": 0, "0,1": 2, "1,0": 4, "1,1": 6},
            "crop": [[0, 0, 2, 2], [0, 0, 2, 2], [4, 4, 6, 6], [4, 4, 6, 6]],
            "average": [[0, 1, 1, 2], [2, 3, 3, 4], [2, 3, 3, 4], [4, 5, 5, 6]],
        },
        "brute_force": [
            brute_force_case((7, 6, 5), (4, 3, 3), (2, 2, 0), None, "average", 1),
            brute_force_case((7, 6, 5), (4, 3, 3), (2, 2, 0), "edge", "average", 2),
            brute_force_case((7, 6, 5), (4, 3, 3), (2, 2, 0), "edge", "crop", 3),
            brute_force_case((7, 6, 5), (4, 3, 3), (2, 2, 0), None, "crop", 4),
            brute_force_case((6, 5, 5), (4, 4, 3), (2, 2, 2), 3, "average", 5),
        ],
    }
    with open(OUT, "w") as f:
        json.dump(golden, f, separators=(",", ":"))
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
