"""ORACLE (test infrastructure, never on the product path): CPU restatement of the reference's label post-processing,
segmentation_pipeline/post_processing.py:5-73, used after inference by research/msseg2/competition/ms-inference.py:47-50,
research/dmri_hippo/hippo_inference.py:40-44 and run_inference.py:206.

The reference calls three scikit-image functions; scikit-image is not in this image, so they are restated from their
published behaviour on top of scipy.ndimage (which is what scikit-image itself calls for two of them):

  * ``skimage.morphology.label(img)`` = ``skimage.measure.label``: background 0, connectivity ``img.ndim`` (26 in 3-D),
    voxels connect only when they carry the SAME value, components numbered 1.. in raster order of their first voxel.
    -> ``label_by_value``: ``ndimage.label`` per distinct value, merged and renumbered by first voxel.
  * ``skimage.morphology.remove_small_holes(mask, area_threshold, connectivity=1)``: NOT(remove_small_objects(NOT mask)),
    and remove_small_objects drops components whose voxel count is ``< min_size`` (``ndi.label`` + ``bincount``).
  * ``skimage.morphology.dilation(img)``: ``ndi.grey_dilation`` with the default footprint
    ``generate_binary_structure(ndim, 1)`` (the cross) and scipy's default 'reflect' border.

Parity status: pinned against scipy (the library scikit-image delegates to), NOT against scikit-image itself -- the
component numbering of ``label`` for multi-valued images is restated from its documentation.  The functions below are
the reference's functions statement by statement, each citing its lines."""
import numpy as np
from scipy import ndimage


def label_by_value(img: np.ndarray) -> np.ndarray:
    """skimage.measure.label(img) with default arguments (background=0, connectivity=img.ndim)."""
    full = ndimage.generate_binary_structure(img.ndim, img.ndim)
    tmp = np.zeros(img.shape, dtype=np.int64)
    offset = 0
    for value in np.unique(img):
        if value == 0:
            continue
        comp, n = ndimage.label(img == value, structure=full)
        sel = comp > 0
        tmp[sel] = comp[sel] + offset
        offset += n
    # renumber 1.. in raster order of each component's first voxel
    labels, first = np.unique(tmp.ravel(), return_index=True)
    fg = labels != 0
    labels, first = labels[fg], first[fg]
    lut = np.zeros(offset + 1, dtype=np.int64)
    lut[labels[np.argsort(first)]] = np.arange(1, labels.size + 1)
    return lut[tmp]


def remove_small_holes(mask: np.ndarray, area_threshold: int) -> np.ndarray:
    """skimage.morphology.remove_small_holes(mask, area_threshold) with connectivity 1."""
    inv = np.logical_not(mask)
    if area_threshold > 0:
        comp, _ = ndimage.label(inv, structure=ndimage.generate_binary_structure(mask.ndim, 1))
        sizes = np.bincount(comp.ravel())
        too_small = sizes < area_threshold
        inv = inv.copy()
        inv[too_small[comp]] = False
    return np.logical_not(inv)


def dilation(img: np.ndarray) -> np.ndarray:
    """skimage.morphology.dilation(img) with the default (cross) footprint."""
    return ndimage.grey_dilation(img, footprint=ndimage.generate_binary_structure(img.ndim, 1))


def unsort_by_size(img, sorted_labels):
    """post_processing.py:5-9."""
    out = img.copy()
    for i in range(sorted_labels.shape[0]):
        out[img == i] = sorted_labels[i]
    return out


def sort_by_size(img, descending=False):
    """post_processing.py:12-26: relabel by rank of the voxel count (np.argsort's own tie order)."""
    out = img.copy()
    labels, counts = np.unique(img, return_counts=True)
    ids = np.argsort(counts)
    if descending:
        ids = ids[::-1]
    labels, counts = labels[ids], counts[ids]
    for i in range(ids.shape[0]):
        out[img == labels[i]] = i
    return out, labels, counts


def keep_components(img, num, max_dilations=100):
    """post_processing.py:29-49."""
    img = img.copy()
    num_components_removed = num_elements_removed = 0
    for i in range(max_dilations):
        comp = label_by_value(img)
        comp_sorted, _, _ = sort_by_size(comp, descending=True)
        keep = comp_sorted <= num
        remove = ~keep
        if i == 0:
            num_elements_removed = remove.sum()
            num_components_removed = comp_sorted.max() - num
        if remove.sum() == 0:
            break
        sorted_img, sorted_labels, _ = sort_by_size(img)
        to_dilate = sorted_img * keep
        dilated = dilation(to_dilate)
        change = (dilated != to_dilate) & remove
        sorted_img[change] = dilated[change]
        img = unsort_by_size(sorted_img, sorted_labels)
    return img, num_components_removed, num_elements_removed


def remove_holes(img, hole_size, max_dilations=100):
    """post_processing.py:52-64."""
    img = img.copy()
    total_holes = 0
    for i in range(max_dilations):
        mask = img > 0
        small_holes = ~mask & remove_small_holes(mask, hole_size)
        num_holes = small_holes.sum()
        if i == 0:
            total_holes = num_holes
        if num_holes == 0:
            break
        img[small_holes] = dilation(img)[small_holes]
    return img, total_holes


def remove_small_components(img, component_size, max_dilations=100):
    """post_processing.py:67-73."""
    img = img.copy()
    inverted = img == 0
    holes_removed, counts = remove_holes(inverted, component_size, max_dilations=max_dilations)
    img[holes_removed] = 0
    return img, counts
