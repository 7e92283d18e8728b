"""Grid sampler / aggregator restatement (oracle/grid.py): torchio's worked example, the patch-count table of
SURVEY.md section 8 and the structural properties of the algorithm.  CPU only."""
import numpy as np
import pytest

from oracle import grid
from oracle import grid as ogrid


def test_torchio_source_comment_example():
    # torchio/data/sampler/grid.py: image 10, patch 5, overlap 2 -> [[0,5],[3,8],[5,10]]
    assert grid.axis_starts(10, 5, 2) == [0, 3, 5]


@pytest.mark.parametrize("size,patch,overlap,n", [
    ((96, 96, 96), 64, 16, 8),            # config 1
    ((304, 304, 240), 96, 48, 144),       # config 2, edge padded
    ((256, 256, 192), 96, 48, 75),        # config 2, unpadded
    ((272, 272, 272), 96, 48, 125),       # config 3, padded
    ((224, 224, 224), 96, 48, 64),
])
def test_patch_counts(size, patch, overlap, n):
    loc = grid.grid_locations(size, patch, overlap)
    assert loc.shape == (n, 6) and loc.dtype == np.int64
    assert (loc[:, 3:] - loc[:, :3] == patch).all()
    assert loc.tolist() == sorted(loc.tolist())
    assert (loc[:, :3] >= 0).all() and (loc[:, 3:] <= np.array(size)).all()


def test_config2_starts():
    loc = grid.grid_locations((304, 304, 240), 96, 48)
    assert sorted(set(loc[:, 0])) == [0, 48, 96, 144, 192, 208]
    assert sorted(set(loc[:, 2])) == [0, 48, 96, 144]


def test_size_errors():
    with pytest.raises(ValueError):
        grid.grid_locations((10, 10, 10), 12, 0)
    with pytest.raises(ValueError):
        grid.grid_locations((10, 10, 10), 4, 4)
    with pytest.raises(ValueError):
        grid.grid_locations((10, 10, 10), 4, 1)


@pytest.mark.parametrize("padding_mode", [None, "edge", 0.0])
def test_average_of_identity_model_is_identity(padding_mode):
    rng = np.random.default_rng(0)
    vol = rng.standard_normal((2, 20, 17, 13)).astype(np.float32)
    out = grid.sliding_window(vol, lambda p: p, (8, 8, 6), (4, 2, 2), padding_mode, "average", 5)
    assert out.shape == vol.shape
    np.testing.assert_allclose(out, vol, rtol=1e-6, atol=1e-6)


def test_count_map_is_separable_and_covers():
    size, patch, overlap = (20, 17, 13), (8, 8, 6), (4, 2, 2)
    loc = grid.grid_locations(size, patch, overlap)
    ones = np.ones((len(loc), 1, *patch), dtype=np.float32)
    _, cnt = grid.aggregate_average(ones, loc, size)
    assert cnt.min() >= 1
    per_axis = []
    for s, p, o in zip(size, patch, overlap):
        c = np.zeros(s)
        for st in grid.axis_starts(s, p, o):
            c[st:st + p] += 1
        per_axis.append(c)
    sep = per_axis[0][:, None, None] * per_axis[1][None, :, None] * per_axis[2][None, None, :]
    np.testing.assert_array_equal(cnt[0], sep)


def test_crop_mode_padded_identity():
    rng = np.random.default_rng(1)
    vol = rng.standard_normal((1, 16, 16, 16)).astype(np.float32)
    out = grid.sliding_window(vol, lambda p: p, 8, 4, "edge", "crop", 4)
    np.testing.assert_array_equal(out, vol)


def test_edge_padding_shape():
    vol = np.arange(2 * 4 * 4 * 4, dtype=np.float32).reshape(2, 4, 4, 4)
    pad = grid.pad_volume(vol, (2, 4, 0), "edge")
    assert pad.shape == (2, 6, 8, 4)
    assert pad[0, 0, 0, 0] == vol[0, 0, 0, 0] and pad[1, -1, -1, -1] == vol[1, -1, -1, -1]


# ------------------------------------------------------------------------------------------------- pinned vectors
def _golden():
    import json
    import os
    from helpers import GOLDEN
    with open(os.path.join(GOLDEN, "grid_torchio.json")) as f:
        return json.load(f)


def test_locations_match_torchio_unit_test_fixture():
    g = _golden()["torchio_test_locations"]
    loc = ogrid.grid_locations(g["image"], g["patch"], g["overlap"])
    assert loc.tolist() == g["locations"]


@pytest.mark.parametrize("mode", ["crop", "average"])
def test_aggregator_matches_torchio_unit_test_fixture(mode):
    g = _golden()["torchio_test_aggregator"]
    vol = np.ones((1, *g["image"]), np.float32)

    def model(patches, locs=[]):
        return patches

    loc = ogrid.grid_locations(g["image"], g["patch"], g["overlap"])
    patches = ogrid.extract_patches(vol, loc)
    for i, l in enumerate(loc):
        patches[i] *= g["patch_values"][f"{l[1]},{l[2]}"]
    if mode == "average":
        out, cnt = ogrid.aggregate_average(patches, loc, g["image"])
        res = ogrid.finalize(out, cnt, g["overlap"], False)
    else:
        res = ogrid.finalize(ogrid.aggregate_crop(patches, loc, g["image"], g["overlap"], False), None, g["overlap"],
                             False)
    np.testing.assert_array_equal(res[0, 0], np.array(g[mode], np.float32))


def test_sliding_window_matches_brute_force_vectors():
    for case in _golden()["brute_force"]:
        vol = np.array(case["volume"], np.float32)[None]
        loc = ogrid.grid_locations([s + 2 * (o // 2 if case["padding_mode"] is not None else 0)
                                    for s, o in zip(case["shape"], case["overlap"])], case["patch"], case["overlap"])
        assert loc.tolist() == case["locations"]
        gains = iter(range(1, len(loc) + 1))
        res = ogrid.sliding_window(vol, lambda p: p * np.array([next(gains) for _ in p], np.float32)[:, None, None, None, None],
                                   case["patch"], case["overlap"], case["padding_mode"], case["overlap_mode"],
                                   patch_batch_size=3)
        np.testing.assert_array_equal(res[0], np.array(case["output"], np.float32))


@pytest.mark.parametrize("padding_mode", [None, "edge", 0.0])
def test_hann_mode_is_a_partition_of_unity_and_separable(padding_mode):
    """'hann' aggregation (newer torchio): an identity model reproduces the volume; the summed 3-D windows of the
    product grid equal the outer product of the per-axis sums the device path divides by."""
    import sys, os
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "segmentation-pipeline_b200"))
    from segmentation_pipeline.grid import PatchGrid
    rng = np.random.default_rng(4)
    vol = rng.normal(size=(2, 20, 24, 32)).astype(np.float32)
    out = grid.sliding_window(vol, lambda p: p, (8, 12, 16), (4, 4, 8), padding_mode, "hann", 5)
    assert out.shape == vol.shape and np.abs(out - vol).max() <= 2e-6
    pg = PatchGrid(vol.shape[1:], (8, 12, 16), (4, 4, 8), padding_mode)
    ones = np.ones((len(pg.locations), 1, 8, 12, 16), dtype=np.float32)
    _, mask = grid.aggregate_hann(ones, np.array(pg.locations), pg.padded_shape)
    s0, s1, s2 = (v.numpy() for v in pg.axis_window_sums())
    sep = (s0[:, None, None] * s1[None, :, None]) * s2[None, None, :]
    assert np.abs(mask[0] - sep).max() <= 1e-6 * sep.max()
    assert mask.min() > 0
    w = grid.hann_window_3d((8, 12, 16))
    ws = pg.hann_windows()
    w3 = ((ws[0][:, None, None] * ws[1][None, :, None]) * ws[2][None, None, :]).numpy()
    assert np.abs(w - w3).max() <= 3e-7
