"""The post-processing oracle (oracle/postproc.py) on hand-checked cases -- post_processing.py:5-73."""
import numpy as np

from oracle import postproc


def test_label_by_value_numbers_in_raster_order_and_splits_values():
    img = np.zeros((1, 4, 6), dtype=np.int32)
    img[0, 0, 0:2] = 2          # first component met in raster order
    img[0, 0, 2:4] = 1          # touches the 2s but carries another value -> its own component
    img[0, 3, 5] = 2            # a second, separate 2
    img[0, 2, 4] = 2            # diagonal neighbour of (3, 5): joined under full connectivity
    out = postproc.label_by_value(img)
    assert out[0, 0, 0] == out[0, 0, 1] == 1
    assert out[0, 0, 2] == out[0, 0, 3] == 2
    assert out[0, 2, 4] == out[0, 3, 5] == 3
    assert out.max() == 3 and (out[img == 0] == 0).all()


def test_remove_small_holes_threshold_is_strict_and_6_connected():
    mask = np.ones((1, 7, 7), dtype=bool)
    mask[0, 1, 1] = mask[0, 1, 2] = False       # hole of 2 voxels
    mask[0, 4, 4] = False
    mask[0, 5, 5] = False                       # diagonal: two holes of 1 voxel under connectivity 1
    assert postproc.remove_small_holes(mask, 2)[0, 4, 4] and postproc.remove_small_holes(mask, 2)[0, 5, 5]
    assert not postproc.remove_small_holes(mask, 2)[0, 1, 1]      # size 2 is not < 2
    assert postproc.remove_small_holes(mask, 3).all()
    assert (postproc.remove_small_holes(mask, 0) == mask).all()


def test_sort_and_unsort_round_trip():
    rng = np.random.default_rng(0)
    img = rng.choice(5, size=(6, 7, 8), p=[0.5, 0.2, 0.15, 0.1, 0.05]).astype(np.int32)
    out, labels, counts = postproc.sort_by_size(img)
    assert (np.diff(counts) >= 0).all() and labels[-1] == 0
    assert (postproc.unsort_by_size(out, labels) == img).all()


def test_remove_holes_fills_from_the_largest_neighbour_label():
    img = np.full((5, 5, 5), 1, dtype=np.int32)
    img[2, 2, 2] = 0
    img[2, 2, 3] = 3
    out, total = postproc.remove_holes(img, 64)
    assert total == 1 and out[2, 2, 2] == 3 and (out != 0).all()


def test_remove_small_components_and_keep_components():
    img = np.zeros((8, 8, 8), dtype=np.int32)
    img[1:4, 1:4, 1:4] = 1          # 27 voxels
    img[6, 6, 6] = 1                # 1 voxel
    img[6, 1, 1:3] = 2              # 2 voxels
    out, removed = postproc.remove_small_components(img, 3)
    assert removed == 3 and out[6, 6, 6] == 0 and out[6, 1, 1] == 0 and out[2, 2, 2] == 1
    kept, n_comp, n_elem = postproc.keep_components(img, 1)
    assert n_comp == 2 and n_elem == 3
    assert kept[6, 6, 6] == 0 and kept[6, 1, 1] == 0 and (kept[1:4, 1:4, 1:4] == 1).all()
