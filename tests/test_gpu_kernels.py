"""GPU parity tests of the individual libb200seg kernels against the CPU oracle (through the C ABI).

Bit-exact for integer / index / byte work (extraction, overlap-add order, argmax, confusion counts); fp32
kernels within 1e-5 (max|d| / max|ref|); the bf16 tensor-core engine against an fp32 CPU convolution of the
same bf16-rounded operands within 1e-2 (output rounding to bf16 is 2^-9 relative)."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import evalstats, grid as ogrid
from helpers import rel_err

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def lib():
    import b200seg
    b200seg.load_library()
    sm, major, minor = b200seg.device_info()
    assert major == 10
    return b200seg


def dev(t):
    return t.cuda().contiguous()


def to_blocked(lib, x, dtype):
    n, c, z, y, xx = x.shape
    buf = lib.Blocked(n, (c + 7) // 8, z, y, xx, dtype, "cuda")
    lib.pack_ncdhw(dev(x), buf.view(c))
    return buf


def from_blocked(lib, buf, c):
    out = torch.empty((buf.n, c, buf.z, buf.y, buf.x), dtype=torch.float32, device="cuda")
    lib.unpack_ncdhw(buf.view(c), out)
    return out.cpu()


# ----------------------------------------------------------------------------------------------- layout / resampling
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_pack_unpack_roundtrip(lib, dtype):
    x = torch.randn(2, 11, 5, 6, 7)
    back = from_blocked(lib, to_blocked(lib, x, dtype), 11)
    ref = x if dtype == torch.float32 else x.to(dtype).float()
    assert torch.equal(back, ref)


def test_pack_zero_fills_padding_channels(lib):
    x = torch.randn(1, 3, 4, 4, 4)
    buf = to_blocked(lib, x, torch.float32)
    assert torch.equal(buf.tensor[0, 0, ..., 3:].cpu(), torch.zeros(4, 4, 4, 5))


def test_avgpool_and_trilinear_match_torch(lib):
    x = torch.randn(2, 16, 8, 6, 10)
    src = to_blocked(lib, x, torch.float32)
    dst = lib.Blocked(2, 2, 4, 3, 5, torch.float32, "cuda")
    lib.avgpool2(src.view(16), dst.view(16))
    assert rel_err(from_blocked(lib, dst, 16), F.avg_pool3d(x, 2, 2, count_include_pad=False)) <= 1e-6
    up = lib.Blocked(2, 4, 16, 12, 20, torch.float32, "cuda")   # written into chunks [1, 3) of a wider buffer
    lib.upsample_trilinear2(src.view(16), up.view(16, 1))
    got = torch.empty(2, 16, 16, 12, 20, device="cuda")
    lib.unpack_ncdhw(up.view(16, 1), got)
    ref = F.interpolate(x, scale_factor=2, mode="trilinear", align_corners=True)
    assert rel_err(got.cpu(), ref) <= 2e-6


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 2e-6), (torch.bfloat16, 1e-2)])
def test_instance_norm_fused(lib, dtype, tol):
    g = torch.Generator().manual_seed(31)
    x = torch.randn(2, 11, 9, 10, 12, generator=g) * 3 + 5      # large mean: the two-pass variance matters
    res = torch.randn(2, 11, 9, 10, 12, generator=g)
    gamma, beta = torch.rand(11, generator=g) + 0.5, torch.randn(11, generator=g)
    xr = x.to(dtype).float()
    ref = F.leaky_relu(F.instance_norm(xr, weight=gamma, bias=beta, eps=1e-5), 0.1) + res.to(dtype).float()
    buf, rbuf = to_blocked(lib, x, dtype), to_blocked(lib, res, dtype)
    lib.instnorm(buf.view(11), dev(gamma), dev(beta), 1e-5, 0.1, rbuf.view(11))
    assert rel_err(from_blocked(lib, buf, 11), ref) <= tol
    assert buf.tensor[:, 1, ..., 3:].abs().max().item() == 0      # padding channels stay zero
    # affine=False, no residual, ReLU
    buf = to_blocked(lib, x, dtype)
    lib.instnorm(buf.view(11), None, None, 1e-5, 0.0)
    assert rel_err(from_blocked(lib, buf, 11), F.relu(F.instance_norm(xr, eps=1e-5))) <= tol


def test_softmax_and_stochastic_matrix(lib):
    from oracle import unet
    x = torch.randn(2, 5, 3, 4, 5)
    y = dev(x.clone())
    lib.softmax_ncdhw(y)
    assert rel_err(y.cpu(), torch.softmax(x, 1)) <= 1e-6
    x = torch.randn(1, 9, 3, 3, 3)
    y = dev(x.clone())
    lib.softmax_ncdhw(y, 3, 1.5)
    assert rel_err(y.cpu(), unet.stochastic_matrix(x, 3, 1.5)) <= 1e-6


# ----------------------------------------------------------------------------------------------- instance evaluation
def _blob_mask(shape, n_blobs, seed, big=False):
    rng = np.random.default_rng(seed)
    zz, yy, xx = np.meshgrid(*[np.arange(s) for s in shape], indexing="ij")
    m = np.zeros(shape, bool)
    for _ in range(n_blobs):
        c = [rng.uniform(0, s) for s in shape]
        r = rng.uniform(1.0, 6.0 if big else 3.5, 3)
        m |= ((zz - c[0]) / r[0]) ** 2 + ((yy - c[1]) / r[1]) ** 2 + ((xx - c[2]) / r[2]) ** 2 < 1
    m |= rng.random(shape) < 0.002                       # isolated voxels and diagonal contacts
    return m


@pytest.mark.parametrize("connectivity", [1, 2, 3])
@pytest.mark.parametrize("shape,dtype", [((40, 36, 28), torch.uint8), ((17, 50, 33), torch.int64), ((64, 64, 48), torch.int32)])
def test_connected_components_match_scipy(lib, connectivity, shape, dtype):
    """Device union-find labelling == scipy.ndimage.label (same partition AND the same raster-scan numbering as
    skimage.morphology.label, which the reference calls at instance_segmentation_evaluator.py:108-109)."""
    m = _blob_mask(shape, 25, 70 + connectivity + shape[0])
    ref, n_ref = evalstats.connected_components(m, connectivity)
    src = torch.from_numpy(m.astype(np.int64) * 3).to(dtype)          # foreground = any value > 0
    labels, n = lib.connected_components(dev(src), connectivity)
    assert n == n_ref and n > 10
    np.testing.assert_array_equal(labels.cpu().numpy(), ref)
    # degenerate inputs: empty and full masks
    z, nz = lib.connected_components(dev(torch.zeros(shape, dtype=dtype)), connectivity)
    assert nz == 0 and int(z.abs().sum()) == 0
    o, no = lib.connected_components(dev(torch.ones(shape, dtype=dtype)), connectivity)
    assert no == 1 and int((o != 1).sum()) == 0


def test_overlap_histogram_bit_exact(lib):
    pm, tm = _blob_mask((48, 40, 36), 30, 81, big=True), _blob_mask((48, 40, 36), 30, 82, big=True)
    hist_ref, n, m = evalstats.instance_overlap_histogram(pm, tm, 2)
    pc, m2 = lib.connected_components(dev(torch.from_numpy(pm.astype(np.uint8))), 2)
    tc, n2 = lib.connected_components(dev(torch.from_numpy(tm.astype(np.uint8))), 2)
    assert (n2, m2) == (n, m)
    hist = lib.overlap_histogram(tc, pc, n, m)
    np.testing.assert_array_equal(hist.cpu().numpy(), hist_ref)
    assert int(hist.sum()) == pm.size


# ----------------------------------------------------------------------------------------------- criterion
@pytest.mark.parametrize("square_dice,weights", [(True, None), (False, [0.3, 1.0, 2.0]), (True, [1.0, 0.5, 0.25])])
def test_hybrid_logistic_dice_loss_matches_oracle_and_autograd(lib, square_dice, weights):
    """Fused device loss vs the oracle restatement of criterions/hybrid_logistic_dice_loss.py:13-43 (itself pinned to the
    reference class by tests/golden/components.npz), values within 1e-5 and d loss / d prediction within 1e-4 of
    torch autograd through the oracle expression."""
    from oracle import unet
    from segmentation_pipeline.criterions import HybridLogisticDiceLoss
    g = torch.Generator().manual_seed(61)
    logits = torch.randn(2, 3, 9, 10, 11, generator=g)
    pred = torch.softmax(logits, 1)
    targ = F.one_hot(torch.randint(0, 3, (2, 9, 10, 11), generator=g), 3).permute(0, 4, 1, 2, 3).float()
    p_ref = pred.clone().requires_grad_(True)
    ref = unet.hybrid_logistic_dice_loss(p_ref, targ, dice_weight=0.3, logistic_class_weights=weights,
                                         square_dice=square_dice)
    ref["loss"].backward()
    p_dev = pred.cuda().requires_grad_(True)
    got = HybridLogisticDiceLoss(dice_weight=0.3, logistic_class_weights=weights, square_dice=square_dice)(p_dev, targ.cuda())
    (got["loss"] * 2.0).backward()                       # non-unit upstream gradient
    for k in ("loss", "dice_loss", "logistic_loss"):
        assert abs(float(got[k]) - float(ref[k])) <= 1e-5 * max(1.0, abs(float(ref[k]))), k
    assert rel_err(p_dev.grad.cpu(), 2.0 * p_ref.grad) <= 1e-4
    with pytest.raises(RuntimeError):
        HybridLogisticDiceLoss()(pred, targ)             # CPU tensors: loud error


# ----------------------------------------------------------------------------------------------- test-time augmentation
def test_tta_kernels_match_tensor_expressions(lib):
    """pack_ncdhw_tta / tta_accumulate / tta_finalize against the reference's flip / permute / stack / mean / argmax /
    mode / one_hot expressions (models/ensemble.py:16-35, 88-103) on a non-cubic volume, all 48 orientations."""
    import itertools
    g = torch.Generator().manual_seed(41)
    x = torch.randn(2, 3, 6, 5, 4, generator=g)
    flips = [()] + [c for k in (1, 2, 3) for c in itertools.combinations((2, 3, 4), k)]
    members, packs = [], []
    acc = torch.zeros(2, 5, 6, 5, 4, device="cuda")
    votes = torch.zeros(2, 5, 6, 5, 4, dtype=torch.uint8, device="cuda")
    e = 0
    for perm in itertools.permutations((2, 3, 4)):
        inv = tuple((torch.argsort(torch.tensor(perm)) + 2).tolist())
        xp = x.permute(0, 1, *perm)
        for dims in flips:
            xt = xp.flip(dims).contiguous()
            p0 = tuple(p - 2 for p in perm)
            fl = tuple((k + 2) in dims for k in range(3))
            # input side: packing with the transform == packing the transformed tensor
            buf = lib.Blocked(2, 1, *xt.shape[2:], torch.float32, "cuda")
            lib.pack_ncdhw_tta(dev(x), p0, fl, buf.view(3))
            assert torch.equal(from_blocked(lib, buf, 3), xt)
            # a fake member output in the member's space (integers -> many argmax ties, exact sums)
            y = torch.randint(0, 3, (2, 5, *xt.shape[2:]), generator=g).float()
            members.append(y.flip(dims).permute(0, 1, *inv))
            lib.tta_accumulate(dev(y), p0, fl, acc, None)
            lib.tta_accumulate(dev(y), p0, fl, None, votes)
            e += 1
    stacked = torch.stack(members)
    lib.tta_finalize(acc, None, None, e)
    assert rel_err(acc.cpu(), stacked.mean(0)) <= 1e-6
    onehot = torch.empty(votes.shape, dtype=torch.int64, device="cuda")
    lib.tta_finalize(None, votes, onehot, e)
    want = F.one_hot(torch.mode(stacked.argmax(2), dim=0).values, num_classes=5).movedim(-1, 1)
    assert torch.equal(onehot.cpu(), want)


# ----------------------------------------------------------------------------------------------- grid
@pytest.mark.parametrize("padding_mode,overlap,size,patch", [
    (None, (4, 2, 2), (20, 17, 13), (8, 8, 6)), ("edge", (4, 2, 2), (20, 17, 13), (8, 8, 6)),
    (0.5, (4, 4, 0), (20, 17, 13), (8, 8, 6)),
    # patch depth multiple of 4: the 4-wide kernel, with unaligned rows (D = 13) and with aligned ones (D = 32)
    ("edge", (4, 2, 2), (20, 17, 13), (8, 8, 8)), (0.5, (4, 2, 4), (20, 17, 13), (8, 8, 8)),
    ("edge", (4, 4, 8), (24, 16, 32), (8, 8, 16)), (None, (4, 4, 8), (24, 16, 32), (8, 8, 16)),
])
def test_grid_extract_bit_exact(lib, padding_mode, overlap, size, patch):
    rng = np.random.default_rng(3)
    vol = rng.standard_normal((3, *size)).astype(np.float32)
    padded = ogrid.pad_volume(vol, overlap, padding_mode)
    loc = ogrid.grid_locations(padded.shape[1:], patch, overlap)
    ref = ogrid.extract_patches(padded, loc)
    border = [o // 2 for o in overlap]
    mode = 0 if padding_mode is None else (1 if padding_mode == "edge" else 2)
    buf = lib.Blocked(len(loc), 1, *patch, torch.float32, "cuda")
    lib.grid_extract(dev(torch.from_numpy(vol)), loc.tolist(), border, mode,
                     0.0 if padding_mode in (None, "edge") else padding_mode, buf.view(3))
    got = from_blocked(lib, buf, 3).numpy()
    np.testing.assert_array_equal(got, ref)


@pytest.mark.parametrize("size,patch,overlap,pad", [
    ((20, 17, 13), (8, 8, 6), (4, 2, 2), None),        # scalar path (odd extents)
    ((32, 24, 40), (16, 16, 16), (8, 8, 8), "edge"),   # 128-bit path
    ((24, 24, 24), (16, 16, 16), (8, 8, 8), None),
])
def test_overlap_add_and_finalize_bit_exact(lib, size, patch, overlap, pad):
    rng = np.random.default_rng(4)
    border = [o // 2 if pad is not None else 0 for o in overlap]
    padded_shape = tuple(s + 2 * b for s, b in zip(size, border))
    loc = ogrid.grid_locations(padded_shape, patch, overlap)
    patches = rng.standard_normal((len(loc), 3, *patch)).astype(np.float32)
    out_ref, cnt_ref = ogrid.aggregate_average(patches, loc, padded_shape)
    probs_ref = ogrid.finalize(out_ref, cnt_ref, overlap, pad is not None)
    out = torch.zeros((3, *padded_shape), device="cuda")
    dp = dev(torch.from_numpy(patches))
    for b0 in range(0, len(loc), 5):     # ragged batches, as PatchPredict feeds them
        lib.overlap_add(out, dp[b0:b0 + 5].contiguous(), loc[b0:b0 + 5].tolist())
    np.testing.assert_array_equal(out.cpu().numpy(), out_ref)
    counts = []
    for s, p, o in zip(padded_shape, patch, overlap):
        c = np.zeros(s, np.int32)
        for st in ogrid.axis_starts(s, p, o):
            c[st:st + p] += 1
        counts.append(dev(torch.from_numpy(c)))
    probs = torch.empty((3, *size), device="cuda")
    lab64 = torch.empty(size, dtype=torch.int64, device="cuda")
    lab8 = torch.empty(size, dtype=torch.uint8, device="cuda")
    lib.finalize(out, counts, border, probs, lab64, lab8)
    np.testing.assert_array_equal(probs.cpu().numpy(), probs_ref)
    ref_lab = evalstats.argmax_labels(probs_ref)[0]
    np.testing.assert_array_equal(lab64.cpu().numpy(), ref_lab)
    np.testing.assert_array_equal(lab8.cpu().numpy(), ref_lab.astype(np.uint8))


def test_overlap_crop_matches_oracle(lib):
    rng = np.random.default_rng(5)
    for pad in (None, "edge"):
        size, patch, overlap = (20, 16, 12), (8, 8, 6), (4, 2, 2)
        border = [o // 2 if pad is not None else 0 for o in overlap]
        padded_shape = tuple(s + 2 * b for s, b in zip(size, border))
        loc = ogrid.grid_locations(padded_shape, patch, overlap)
        patches = rng.standard_normal((len(loc), 2, *patch)).astype(np.float32)
        ref = ogrid.aggregate_crop(patches, loc, padded_shape, overlap, pad is not None)
        out = torch.zeros((2, *padded_shape), device="cuda")
        lib.overlap_crop(out, dev(torch.from_numpy(patches)), loc.tolist(), [o // 2 for o in overlap], pad is not None)
        np.testing.assert_array_equal(out.cpu().numpy(), ref)


def _cuda_sliding_window(lib, vol, patch, overlap, padding_mode, overlap_mode, gain, batch=3):
    """The product's host grid + CUDA extraction / aggregation / finalize with a 'model' that scales patch n by
    gain(n, location) -- the device-side counterpart of oracle.grid.sliding_window."""
    from segmentation_pipeline.grid import PatchGrid
    grid = PatchGrid(vol.shape[1:], patch, overlap, padding_mode)
    c = vol.shape[0]
    dvol = dev(torch.from_numpy(vol))
    out = torch.zeros((c, *grid.padded_shape), device="cuda")
    n = 0
    for locs in grid.batches(batch):
        buf = lib.Blocked(len(locs), (c + 7) // 8, *grid.patch_size, torch.float32, "cuda")
        lib.grid_extract(dvol, locs, grid.border, grid.pad_mode_code, grid.pad_value, buf.view(c))
        y = torch.empty((len(locs), c, *grid.patch_size), device="cuda")
        lib.unpack_ncdhw(buf.view(c), y)
        for i, l in enumerate(locs):
            y[i] *= gain(n, l)
            n += 1
        if overlap_mode == "average":
            lib.overlap_add(out, y, locs)
        else:
            lib.overlap_crop(out, y, locs, [o // 2 for o in grid.patch_overlap], grid.volume_padded)
    counts = None
    if overlap_mode == "average":
        counts = [torch.tensor(cc, dtype=torch.int32, device="cuda") for cc in grid.axis_counts()]
    probs = torch.empty((c, *grid.spatial_shape), device="cuda")
    lib.finalize(out, counts, grid.border, probs, None, None)
    return grid, probs.cpu().numpy()


def test_cuda_grid_path_matches_pinned_torchio_vectors(lib):
    """tests/golden/grid_torchio.json: torchio's own unit-test fixtures + brute-force cases (oracle/make_grid_golden.py)."""
    import json
    import os
    from helpers import GOLDEN
    from segmentation_pipeline.grid import PatchGrid
    with open(os.path.join(GOLDEN, "grid_torchio.json")) as f:
        g = json.load(f)
    t = g["torchio_test_locations"]
    assert [list(l) for l in PatchGrid(t["image"], t["patch"], t["overlap"]).locations] == t["locations"]
    a = g["torchio_test_aggregator"]
    ones = np.ones((1, *a["image"]), np.float32)
    for mode in ("crop", "average"):
        _, res = _cuda_sliding_window(lib, ones, a["patch"], a["overlap"], None, mode,
                                      lambda n, l: float(a["patch_values"][f"{l[1]},{l[2]}"]))
        np.testing.assert_array_equal(res[0, 0], np.array(a[mode], np.float32))
    for case in g["brute_force"]:
        vol = np.array(case["volume"], np.float32)[None]
        grid, res = _cuda_sliding_window(lib, vol, case["patch"], case["overlap"], case["padding_mode"],
                                         case["overlap_mode"], lambda n, l: float(n + 1))
        assert [list(l) for l in grid.locations] == case["locations"]
        np.testing.assert_array_equal(res[0], np.array(case["output"], np.float32))


def test_cuda_grid_path_against_real_torchio_when_installed(lib):
    """Runs tio.GridSampler / tio.GridAggregator themselves when the box has torchio (the authoring image does not:
    the test then SKIPS and says so; the pinned vectors above are the torchio-derived check that always runs)."""
    tio = pytest.importorskip("torchio", reason="torchio is not installed on this box: real-torchio comparison skipped")
    from torch.utils.data import DataLoader
    rng = np.random.default_rng(11)
    vol = rng.standard_normal((2, 20, 18, 14)).astype(np.float32)
    for padding_mode, overlap_mode in ((None, "average"), ("edge", "average"), ("edge", "crop"), (None, "crop")):
        subject = tio.Subject(X=tio.ScalarImage(tensor=torch.from_numpy(vol)))
        sampler = tio.GridSampler(subject, (8, 8, 6), (4, 2, 2), padding_mode=padding_mode)
        aggregator = tio.GridAggregator(sampler, overlap_mode=overlap_mode)
        n = 0
        for batch in DataLoader(sampler, batch_size=3):
            data = batch["X"][tio.DATA].clone()
            for i in range(data.shape[0]):
                data[i] *= float(n + 1)
                n += 1
            aggregator.add_batch(data, batch[tio.LOCATION])
        ref = aggregator.get_output_tensor().numpy()
        _, res = _cuda_sliding_window(lib, vol, (8, 8, 6), (4, 2, 2), padding_mode, overlap_mode,
                                      lambda k, l: float(k + 1))
        np.testing.assert_array_equal(res, ref)
    print("real torchio", tio.__version__, "compared")


def test_argmax_ties_and_bit_exact(lib):
    rng = np.random.default_rng(6)
    for vox in (4 * 1000, 1003):
        p = rng.integers(0, 4, size=(5, vox)).astype(np.float32)   # many ties
        lab = torch.empty(vox, dtype=torch.int64, device="cuda")
        lab8 = torch.empty(vox, dtype=torch.uint8, device="cuda")
        lib.argmax(dev(torch.from_numpy(p)), lab, lab8)
        np.testing.assert_array_equal(lab.cpu().numpy(), np.argmax(p, axis=0))
        np.testing.assert_array_equal(lab8.cpu().numpy(), np.argmax(p, axis=0).astype(np.uint8))


@pytest.mark.parametrize("dtype,nc,vox", [(torch.uint8, 2, 1 << 20), (torch.uint8, 10, 999_983),
                                          (torch.int64, 17, 300_001), (torch.uint8, 40, 250_000)])
def test_confusion_bit_exact(lib, dtype, nc, vox):
    rng = np.random.default_rng(7)
    p = rng.integers(0, nc + 2, size=vox)      # includes values outside [0, nc): counted in the last class
    t = rng.integers(0, nc + 2, size=vox)
    if dtype == torch.int64:
        p[::97] = -3                           # negative labels too
    cm = torch.zeros((nc, nc), dtype=torch.int64, device="cuda")
    lib.confusion(dev(torch.from_numpy(p).to(dtype)), dev(torch.from_numpy(t).to(dtype)), nc, cm)
    np.testing.assert_array_equal(cm.cpu().numpy(), evalstats.confusion_matrix(p, t, nc))
    lib.confusion(dev(torch.from_numpy(p).to(dtype)), dev(torch.from_numpy(t).to(dtype)), nc, cm)   # accumulates
    np.testing.assert_array_equal(cm.cpu().numpy(), 2 * evalstats.confusion_matrix(p, t, nc))


# ----------------------------------------------------------------------------------------------- convolutions
def _epi(lib, cout, dst=None, out=None, softmax=False, bias=None, slope=1.0, residual=None):
    cpad = (cout + 7) // 8 * 8
    scale = torch.ones(cpad)
    shift = torch.zeros(cpad)
    if bias is not None:
        shift[:cout] = bias
    sl = torch.ones(cpad)
    sl[:cout] = slope
    keep = (dev(scale), dev(shift), dev(sl))
    epi = lib.make_epilogue(*keep, dst0=dst if dst is not None else lib.NULL_VIEW,
                            residual=residual if residual is not None else lib.NULL_VIEW, out_ncdhw=out,
                            softmax=softmax)
    return epi, keep


@pytest.mark.parametrize("k,stride,pad,transposed,ext", [(3, 1, 1, False, (5, 9, 7)), (4, 2, 1, False, (8, 6, 10)),
                                                         (4, 2, 1, True, (3, 5, 4))])
def test_conv_direct_fp32(lib, k, stride, pad, transposed, ext):
    from segmentation_pipeline.models import _plan
    g = torch.Generator().manual_seed(11)
    cin, cout = 11, 13
    x = torch.randn(2, cin, *ext, generator=g)
    if transposed:
        w = torch.randn(cin, cout, k, k, k, generator=g) * 0.1
        ref = F.conv_transpose3d(x, w, stride=stride, padding=pad)
    else:
        w = torch.randn(cout, cin, k, k, k, generator=g) * 0.1
        ref = F.conv3d(x, w, stride=stride, padding=pad)
    bias = torch.randn(cout, generator=g)
    ref = F.leaky_relu(ref + bias.view(1, -1, 1, 1, 1), 0.1)
    op = _plan.ConvOp({(3, False): 0, (4, False): 1, (4, True): 2}[(k, transposed)], _plan.Ref("in", 0, cin),
                      [(0, cin)], w, np.ones(cout, np.float32), np.zeros(cout, np.float32), np.ones(cout, np.float32))
    wd = dev(_plan.pack_direct_weight(op, 2))
    src = to_blocked(lib, x, torch.float32)
    dst = lib.Blocked(2, 2, *ref.shape[2:], torch.float32, "cuda")
    epi, keep = _epi(lib, cout, dst=dst.view(cout), bias=bias, slope=0.1)
    lib.conv3d_direct(src.view(cin), wd, cout, k, stride, pad, transposed, epi)
    assert rel_err(from_blocked(lib, dst, cout), ref) <= 1e-5


TC_CASES = [
    # mode, cin, cout, extent (n, z, y, x)
    (0, 16, 16, (1, 4, 16, 8)),       # one exact tile
    (0, 24, 40, (2, 13, 20, 19)),     # ragged tiles, lone chunk, N spill, batch
    (0, 2, 40, (1, 6, 18, 10)),       # first layer: 2 channels in one lone chunk
    (0, 80, 80, (1, 12, 16, 16)),     # widest N, two z tiles
    (0, 40, 8, (1, 5, 9, 9)),         # narrow output
    (1, 16, 16, (1, 8, 20, 12)),
    (1, 40, 40, (2, 12, 36, 20)),
    (2, 16, 16, (1, 3, 5, 9)),
    (2, 40, 40, (2, 7, 17, 11)),
    (2, 80, 80, (1, 4, 6, 6)),
    # >= 2 tiles per SM: the dual-issuer schedule (two A rings, shared weight ring, ghost rounds for CTAs with an odd
    # tile count), ragged extents
    (0, 40, 40, (3, 29, 50, 44)),     # resident weights, 5-chunk epilogue with residual
    (0, 80, 80, (2, 22, 50, 44)),     # streamed weights released by both issuers, 10-chunk epilogue
    (0, 24, 16, (5, 40, 36, 28)),     # generic epilogue under the dual schedule
    (1, 40, 40, (3, 44, 68, 60)),     # stride-2 conv, 24 streamed images per unit
    (2, 80, 80, (4, 15, 34, 30)),     # transposed conv (4 passes per tile pair)
    (2, 40, 40, (4, 15, 34, 30)),     # transposed conv on the single-issuer schedule (default rule)
    # 120 output channels in one launch (deep levels): TZ = 2, 15-chunk generic epilogue, single weight slot
    (0, 120, 120, (6, 3, 3, 3)),
    (0, 240, 120, (3, 6, 6, 6)),
    (1, 120, 120, (4, 6, 6, 6)),
    (2, 120, 120, (4, 3, 3, 3)),
]


@pytest.mark.parametrize("mode,cin,cout,ext", TC_CASES)
def test_conv_tc_matches_fp32_oracle(lib, mode, cin, cout, ext):
    from segmentation_pipeline.models import _plan
    import os
    if cout > 80 and os.environ.get("B200SEG_TC_VARIANT") == "1":
        pytest.skip("120-wide launches need the one-CTA-per-SM shared-memory budget")
    g = torch.Generator().manual_seed(100 + mode * 31 + cin + cout)
    n = ext[0]
    x = torch.randn(n, cin, *ext[1:], generator=g).to(torch.bfloat16).float()
    if mode == 2:
        w = (torch.randn(cin, cout, 4, 4, 4, generator=g) * 0.05).to(torch.bfloat16).float()
        ref = F.conv_transpose3d(x, w, stride=2, padding=1)
    elif mode == 1:
        w = (torch.randn(cout, cin, 4, 4, 4, generator=g) * 0.05).to(torch.bfloat16).float()
        ref = F.conv3d(x, w, stride=2, padding=1)
    else:
        w = (torch.randn(cout, cin, 3, 3, 3, generator=g) * 0.05).to(torch.bfloat16).float()
        ref = F.conv3d(x, w, padding=1)
    bias = torch.randn(cout, generator=g) * 0.1
    res = torch.randn(ref.shape, generator=g).to(torch.bfloat16).float()
    ref = F.relu(ref + bias.view(1, -1, 1, 1, 1)) + res
    chunks = (cin + 7) // 8
    phys = _plan.physical_weight(w, mode == 2, [(0, cin)], chunks, 0, cout)
    packed = dev(_plan.pack_tc_weight(mode, phys, chunks, cout))
    assert packed.numel() * 2 == lib.conv3d_tc_wbytes(mode, chunks, cout)
    src = to_blocked(lib, x, torch.bfloat16)
    rbuf = to_blocked(lib, res, torch.bfloat16)
    c8 = (cout + 7) // 8
    dst = lib.Blocked(n, c8 + 1, *ref.shape[2:], torch.bfloat16, "cuda")    # output lands in chunks [1, 1+c8)
    dst.tensor.zero_()
    epi, keep = _epi(lib, cout, dst=dst.view(cout, 1), bias=bias, slope=0.0, residual=rbuf.view(cout))
    lib.conv3d_tc(mode, src.view(cin), packed, cout, epi)
    torch.cuda.synchronize()
    got = torch.empty(ref.shape, device="cuda")
    lib.unpack_ncdhw(dst.view(cout, 1), got)
    assert rel_err(got.cpu(), ref) <= 1e-2
    assert dst.tensor[:, 0].abs().max().item() == 0          # neighbouring chunk untouched


@pytest.mark.parametrize("mode,cin,cout,ext", [(1, 40, 40, (2, 12, 36, 20)), (2, 40, 40, (2, 7, 17, 11)),
                                               (2, 80, 80, (4, 15, 34, 30)), (1, 80, 80, (2, 8, 20, 12)),
                                               (2, 24, 24, (1, 5, 9, 7))])
def test_conv_tc_identity_epilogue(lib, mode, cin, cout, ext):
    """slope01 = 2: the blur convolutions' epilogue (scale 1, shift 0, no activation) as a plain conversion -- must equal
    the general epilogue bit for bit."""
    from segmentation_pipeline.models import _plan
    g = torch.Generator().manual_seed(500 + mode + cin)
    x = torch.randn(ext[0], cin, *ext[1:], generator=g).to(torch.bfloat16).float()
    if mode == 2:
        w = (torch.randn(cin, cout, 4, 4, 4, generator=g) * 0.05).to(torch.bfloat16).float()
        ref = F.conv_transpose3d(x, w, stride=2, padding=1)
    else:
        w = (torch.randn(cout, cin, 4, 4, 4, generator=g) * 0.05).to(torch.bfloat16).float()
        ref = F.conv3d(x, w, stride=2, padding=1)
    chunks = (cin + 7) // 8
    packed = dev(_plan.pack_tc_weight(mode, _plan.physical_weight(w, mode == 2, [(0, cin)], chunks, 0, cout), chunks, cout))
    src = to_blocked(lib, x, torch.bfloat16)
    outs = []
    for flag in (1, 2):
        dst = lib.Blocked(ext[0], (cout + 7) // 8, *ref.shape[2:], torch.bfloat16, "cuda")
        dst.tensor.zero_()
        cpad = (cout + 7) // 8 * 8
        ones, zeros = dev(torch.ones(cpad)), dev(torch.zeros(cpad))
        epi = lib.make_epilogue(ones, zeros, ones, dst.view(cout), slope01=flag)
        lib.conv3d_tc(mode, src.view(cin), packed, cout, epi)
        torch.cuda.synchronize()
        outs.append(dst.tensor.clone())
    assert torch.equal(outs[0], outs[1])
    got = torch.empty(ref.shape, device="cuda")
    dst.tensor.copy_(outs[1])
    lib.unpack_ncdhw(dst.view(cout), got)
    assert rel_err(got.cpu(), ref) <= 1e-2


def test_plan_graph_replay_matches_eager():
    """Small inputs replay a captured CUDA graph of the whole op list (third call on a workspace): same bits as the
    eager launches, and the module-forward API hands out its own copy."""
    from helpers import load_case
    from test_gpu_models import build_model
    from segmentation_pipeline.models import set_precision
    meta, sd, x, y = load_case("models_modular_blur")
    model = build_model(meta)
    model.load_state_dict(sd)
    model.eval().cuda()
    set_precision("bf16")
    try:
        with torch.no_grad():
            outs = [model(x.cuda()) for _ in range(4)]           # eager, capture, replay, replay
            x2 = (x * 0.5).cuda()
            o2 = model(x2)
            compiled = next(iter(model.__dict__["_b200_cache"].values()))[1]
            ws = next(iter(compiled.workspaces.values()))
            assert "__graph__" in ws
    finally:
        set_precision("auto")
    for o in outs[1:]:
        assert torch.equal(o, outs[0])
    assert outs[0].data_ptr() != outs[1].data_ptr() and not torch.equal(o2, outs[0])
    assert rel_err(outs[0].cpu(), y) <= 2e-2


def test_conv_tc_softmax_head(lib):
    from segmentation_pipeline.models import _plan
    g = torch.Generator().manual_seed(77)
    for cout in (2, 10):
        x = torch.randn(2, 40, 7, 18, 9, generator=g).to(torch.bfloat16).float()
        w = (torch.randn(cout, 40, 3, 3, 3, generator=g) * 0.05).to(torch.bfloat16).float()
        bias = torch.randn(cout, generator=g)
        ref = torch.softmax(F.conv3d(x, w, bias, padding=1), 1)
        phys = _plan.physical_weight(w, False, [(0, 40)], 5, 0, cout)
        packed = dev(_plan.pack_tc_weight(0, phys, 5, cout))
        out = torch.empty(ref.shape, device="cuda")
        epi, keep = _epi(lib, cout, out=out, softmax=True, bias=bias)
        lib.conv3d_tc(0, to_blocked(lib, x, torch.bfloat16).view(40), packed, cout, epi)
        assert (out.cpu() - ref).abs().max() <= 2e-5
        assert torch.allclose(out.sum(1).cpu(), torch.ones(2, 7, 18, 9), atol=1e-5)


@pytest.mark.parametrize("cin,cout,ext,softmax", [(40, 2, (2, 7, 30, 14), True),     # out_conv: 3 x 3 tiles of 14 x 6
                                                  (40, 2, (1, 96, 96, 96), True),   # bench shape, dual issuers
                                                  (8, 1, (1, 3, 15, 7), False),     # lone chunk, one class, logits
                                                  (24, 4, (2, 13, 17, 9), True),    # widest tap block, ragged tiles
                                                  (16, 3, (1, 5, 14, 6), False)])
def test_conv_tc_taps_in_n_head(lib, cin, cout, ext, softmax):
    """B200SEG_TC_K3T (final layer with cout <= 4: nine in-plane taps as accumulator columns, summed in the
    epilogue) against the fp32 torch convolution of the same bf16-rounded operands."""
    from segmentation_pipeline.models import _plan
    g = torch.Generator().manual_seed(79 + cin + cout)
    n = ext[0]
    x = torch.randn(n, cin, *ext[1:], generator=g).to(torch.bfloat16).float()
    w = (torch.randn(cout, cin, 3, 3, 3, generator=g) * 0.05).to(torch.bfloat16).float()
    bias = torch.randn(cout, generator=g)
    ref = F.conv3d(x, w, bias, padding=1)
    if softmax:
        ref = torch.softmax(ref, 1)
    chunks = (cin + 7) // 8
    phys = _plan.physical_weight(w, False, [(0, cin)], chunks, 0, cout)
    packed = dev(_plan.pack_tc_weight(_plan.K3T, phys, chunks, cout))
    assert packed.numel() * 2 == lib.conv3d_tc_wbytes(_plan.K3T, chunks, cout)
    out = torch.full(ref.shape, float("nan"), device="cuda")
    epi, keep = _epi(lib, cout, out=out, softmax=softmax, bias=bias)
    lib.conv3d_tc(_plan.K3T, to_blocked(lib, x, torch.bfloat16).view(cin), packed, cout, epi)
    torch.cuda.synchronize()
    assert not torch.isnan(out).any()                       # every voxel is owned by exactly one tile
    assert (out.cpu() - ref).abs().max() <= (2e-5 if softmax else 2e-4)


def test_conv_tc_fused_two_destinations(lib):
    """conv0 || res_conv as one contraction: channels [0,40) -> ReLU'd tensor, [40,80) -> bias-only tensor."""
    from segmentation_pipeline.models import _plan
    g = torch.Generator().manual_seed(78)
    x = torch.randn(1, 16, 6, 16, 16, generator=g).to(torch.bfloat16).float()
    w = (torch.randn(80, 16, 3, 3, 3, generator=g) * 0.05).to(torch.bfloat16).float()
    bias = torch.randn(80, generator=g) * 0.1
    full = F.conv3d(x, w, bias, padding=1)
    phys = _plan.physical_weight(w, False, [(0, 16)], 2, 0, 80)
    packed = dev(_plan.pack_tc_weight(0, phys, 2, 80))
    d0 = lib.Blocked(1, 5, 6, 16, 16, torch.bfloat16, "cuda")
    d1 = lib.Blocked(1, 5, 6, 16, 16, torch.bfloat16, "cuda")
    slope = torch.ones(80)
    slope[:40] = 0.0
    keep = (dev(torch.ones(80)), dev(bias), dev(slope))
    epi = lib.make_epilogue(*keep, dst0=d0.view(40), dst1=d1.view(40), split=40)
    lib.conv3d_tc(0, to_blocked(lib, x, torch.bfloat16).view(16), packed, 80, epi)
    assert rel_err(from_blocked(lib, d0, 40), F.relu(full[:, :40])) <= 1e-2
    assert rel_err(from_blocked(lib, d1, 40), full[:, 40:]) <= 1e-2


def test_errors_are_loud(lib):
    x = lib.Blocked(1, 1, 4, 4, 4, torch.float32, "cuda")
    with pytest.raises(RuntimeError, match="bf16"):
        keep = (dev(torch.ones(8)), dev(torch.zeros(8)), dev(torch.ones(8)))
        epi = lib.make_epilogue(*keep, dst0=x.view(8))
        lib.conv3d_tc(0, x.view(8), dev(torch.zeros(16)), 8, epi)
    with pytest.raises(RuntimeError):
        lib.pack_ncdhw(torch.zeros(1, 1, 2, 2, 2), x.view(1))      # CPU tensor: no fallback
