"""Device post-processing (segmentation_pipeline/post_processing.py, csrc/ccl_kernels.cu) against the scipy oracle
(oracle/postproc.py) -- bit-exact: labels, counts, tie order.  Reference: post_processing.py:5-73."""
import numpy as np
import pytest
import torch

from oracle import postproc as ref

pytestmark = pytest.mark.gpu


def lesion_map(shape, n_labels, seed, n_blobs=40, speckle=0.002, holes=0.01):
    """A label map that looks like a network's output: blobs of several labels, isolated speckle voxels, pinholes."""
    rng = np.random.default_rng(seed)
    img = np.zeros(shape, dtype=np.int32)
    grid = np.stack(np.meshgrid(*[np.arange(s) for s in shape], indexing="ij"), axis=-1)
    for _ in range(n_blobs):
        centre = rng.uniform(0, 1, 3) * np.array(shape)
        radius = rng.uniform(1.0, min(shape) / 5)
        scale = rng.uniform(0.6, 1.4, 3)
        inside = (((grid - centre) / (radius * scale)) ** 2).sum(-1) < 1
        img[inside] = rng.integers(1, n_labels + 1)
    img[rng.random(shape) < speckle] = rng.integers(1, n_labels + 1)
    img[(rng.random(shape) < holes) & (img > 0)] = 0
    return img


CASES = [((24, 20, 28), 1, 0), ((40, 33, 27), 3, 1), ((17, 64, 19), 5, 2), ((48, 48, 48), 2, 3)]


@pytest.mark.parametrize("shape,n_labels,seed", CASES)
def test_label_by_value_matches_oracle(shape, n_labels, seed):
    import b200seg
    img = lesion_map(shape, n_labels, seed)
    comp, n = b200seg.connected_components(torch.from_numpy(img).cuda(), connectivity=3, by_value=1)
    want = ref.label_by_value(img)
    assert n == want.max()
    assert (comp.cpu().numpy() == want).all()


@pytest.mark.parametrize("shape,n_labels,seed", CASES)
def test_dilate_cross_matches_grey_dilation(shape, n_labels, seed):
    import b200seg
    img = lesion_map(shape, n_labels, seed) * 7 - 3          # negative values too
    got = b200seg.dilate_cross(torch.from_numpy(img).cuda()).cpu().numpy()
    assert (got == ref.dilation(img)).all()


@pytest.mark.parametrize("shape,n_labels,seed", CASES)
@pytest.mark.parametrize("descending", [False, True])
def test_sort_by_size_round_trip(shape, n_labels, seed, descending):
    from segmentation_pipeline import post_processing as pp
    img = lesion_map(shape, n_labels, seed)
    got, labels, counts = pp.sort_by_size(img, descending=descending)
    want, wlabels, wcounts = ref.sort_by_size(img, descending=descending)
    assert got.dtype == img.dtype and (got == want).all()
    assert (labels == wlabels).all() and (counts == wcounts).all()
    assert (pp.unsort_by_size(got, labels) == img).all()


@pytest.mark.parametrize("shape,n_labels,seed", CASES)
@pytest.mark.parametrize("hole_size", [0, 1, 2, 8, 64])
def test_remove_holes(shape, n_labels, seed, hole_size):
    from segmentation_pipeline import post_processing as pp
    img = lesion_map(shape, n_labels, seed, holes=0.03)
    got, total = pp.remove_holes(img, hole_size)
    want, wtotal = ref.remove_holes(img, hole_size)
    assert total == wtotal
    assert got.dtype == img.dtype and (got == want).all()


def test_remove_holes_needs_several_dilations_and_keeps_large_holes():
    from segmentation_pipeline import post_processing as pp
    img = np.full((20, 20, 20), 2, dtype=np.int64)
    img[5:9, 5:9, 5:9] = 0              # 64 voxels: filled only for hole_size > 64, in two sweeps
    img[12:18, 12:18, 12:18] = 0        # 216 voxels: stays
    img[0, 0, 0] = 0                    # corner hole
    for hole_size in (64, 65):
        got, total = pp.remove_holes(img, hole_size)
        want, wtotal = ref.remove_holes(img, hole_size)
        assert total == wtotal and (got == want).all() and got.dtype == np.int64
    assert (got[5:9, 5:9, 5:9] == 2).all() and (got[12:18, 12:18, 12:18] == 0).all()


@pytest.mark.parametrize("shape,n_labels,seed", CASES)
@pytest.mark.parametrize("size", [1, 3, 10])
def test_remove_small_components(shape, n_labels, seed, size):
    from segmentation_pipeline import post_processing as pp
    img = lesion_map(shape, n_labels, seed, speckle=0.01)
    got, removed = pp.remove_small_components(img, size)
    want, wremoved = ref.remove_small_components(img, size)
    assert removed == wremoved and (got == want).all()


@pytest.mark.parametrize("shape,n_labels,seed", CASES)
def test_keep_components(shape, n_labels, seed):
    from segmentation_pipeline import post_processing as pp
    img = lesion_map(shape, n_labels, seed, n_blobs=12)
    for num in (n_labels, 1, 0):
        got, n_comp, n_elem = pp.keep_components(img, num)
        want, wn_comp, wn_elem = ref.keep_components(img, num)
        assert (n_comp, n_elem) == (wn_comp, wn_elem)
        assert (got == want).all()


def test_hippo_post_process_sequence():
    """research/dmri_hippo/hippo_inference.py:36-48: remove_holes(64) then keep_components(max label)."""
    from segmentation_pipeline import post_processing as pp
    img = lesion_map((48, 64, 64), 2, 11, n_blobs=10, speckle=0.004, holes=0.02)
    got, _ = pp.remove_holes(img, hole_size=64)
    got, a, b = pp.keep_components(got, got.max())
    want, _ = ref.remove_holes(img, hole_size=64)
    want, wa, wb = ref.keep_components(want, want.max())
    assert (a, b) == (wa, wb) and (got == want).all()


def test_msseg_post_process_sequence():
    """research/msseg2/competition/ms-inference.py:47-50: remove_holes(64) then remove_small_components(3)."""
    from segmentation_pipeline import post_processing as pp
    img = lesion_map((64, 64, 48), 1, 12, n_blobs=60, speckle=0.004, holes=0.02)
    got, a = pp.remove_holes(img, hole_size=64)
    got, b = pp.remove_small_components(got, 3)
    want, wa = ref.remove_holes(img, hole_size=64)
    want, wb = ref.remove_small_components(want, 3)
    assert (a, b) == (wa, wb) and (got == want).all()


def test_cuda_tensor_in_cuda_tensor_out_and_no_cpu_path():
    from segmentation_pipeline import post_processing as pp
    img = lesion_map((24, 24, 24), 2, 5)
    dev = torch.from_numpy(img).cuda().to(torch.int64)
    got, total = pp.remove_holes(dev, 8)
    assert got.is_cuda and got.dtype == torch.int64 and (dev.cpu().numpy() == img).all()      # input untouched
    want, wtotal = ref.remove_holes(img, 8)
    assert total == wtotal and (got.cpu().numpy() == want).all()
    with pytest.raises(RuntimeError):
        pp.remove_holes(torch.from_numpy(img), 8)
