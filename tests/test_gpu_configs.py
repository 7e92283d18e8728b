"""GPU parity at the BASELINE.json configurations.

config 1 (reference's CPU-runnable case) is run IN FULL against the oracle; configs 2/3/4 are checked with the
oracle on a sample the CPU finishes in seconds and, at full size, through size-independent properties
(probabilities sum to one, labels == argmax of the returned probabilities, identity-model round trip, bf16 vs
fp32 path agreement on the same kernels' inputs, confusion counts summing to the voxel count)."""
import numpy as np
import pytest
import torch

from oracle import evalstats, grid as ogrid, unet
from helpers import label_agreement_report, rel_err

pytestmark = pytest.mark.gpu


def _perturb_bn(model, seed):
    g = torch.Generator().manual_seed(seed)
    for m in model.modules():
        if isinstance(m, torch.nn.BatchNorm3d):
            m.running_mean.copy_(0.1 * torch.randn(m.running_mean.shape, generator=g))
            m.running_var.copy_(0.5 + torch.rand(m.running_var.shape, generator=g))
            m.weight.data.copy_(0.5 + torch.rand(m.weight.shape, generator=g))
            m.bias.data.copy_(0.1 * torch.randn(m.bias.shape, generator=g))


def _smooth_volume(shape, seed):
    g = torch.Generator().manual_seed(seed)
    c, w, h, d = shape
    coarse = torch.randn(1, c, max(w // 4, 1), max(h // 4, 1), max(d // 4, 1), generator=g)
    vol = torch.nn.functional.interpolate(coarse, size=(w, h, d), mode="trilinear", align_corners=False)[0]
    return (vol + 0.1 * torch.randn(shape, generator=g)).contiguous()


def test_config1_full_run_fp32_and_bf16():
    """ModularUNet(1, 2, [40, 80, 120], 3) defaults, 1 x 96^3 volume, patch 64^3, overlap 16 -> 8 patches."""
    from segmentation_pipeline import models as M
    from segmentation_pipeline.models import set_precision
    from segmentation_pipeline.prediction import PatchPredict
    torch.manual_seed(0)
    model = M.ModularUNet(1, 2, [40, 80, 120], 3)
    _perturb_bn(model, 1)
    model.eval()
    sd = {k: v.clone() for k, v in model.state_dict().items()}
    cfg = {"depth": 3, "filters": [40, 80, 120], "block": {"residual": False}, "down": "avgpool", "up": "trilinear"}
    vol = _smooth_volume((1, 96, 96, 96), 3)
    torch.set_num_threads(max(torch.get_num_threads(), 8))
    ref = ogrid.sliding_window(vol.numpy(), lambda p: unet.modular_unet_forward(sd, torch.from_numpy(p), cfg).numpy(),
                               64, 16, None, "average", patch_batch_size=8)
    model.cuda()
    predictor = PatchPredict(patch_batch_size=8, patch_size=64, patch_overlap=16, padding_mode=None)
    out = {}
    for precision in ("fp32", "bf16"):
        set_precision(precision)
        try:
            with torch.no_grad():
                probs, labels = predictor.predict_volume(model, vol.cuda())
        finally:
            set_precision("auto")
        out[precision] = (probs.cpu(), labels.cpu())
    assert rel_err(out["fp32"][0], torch.from_numpy(ref)) <= 1e-5
    assert rel_err(out["bf16"][0], torch.from_numpy(ref)) <= 2e-2
    # unfiltered label agreement (random-init head: see tests/test_gpu_labels.py for the confident-head criterion)
    ref_t = torch.from_numpy(ref)
    rep32 = label_agreement_report(ref_t, out["fp32"][0], "config 1, fp32 path, random head")
    rep16 = label_agreement_report(ref_t, out["bf16"][0], "config 1, bf16 path, random head")
    assert rep32["agreement"] >= 0.9999 and rep32["disagree_margin_max"] <= 1e-4
    assert rep16["agreement"] >= 0.97 and rep16["disagree_margin_max"] <= 2e-2
    # labels are exactly the argmax of the probabilities the same call returned
    for precision in out:
        np.testing.assert_array_equal(out[precision][1].numpy(), evalstats.argmax_labels(out[precision][0].numpy())[0])


def _msseg2_model():
    from segmentation_pipeline import models as M
    torch.manual_seed(0)
    model = M.ModularUNet(in_channels=2, out_channels=2, filters=[40, 40, 80, 80, 120, 120], depth=6,
                          block_params={'residual': True}, downsample_class=M.BlurConv3d,
                          downsample_params={'kernel_size': 3, 'stride': 2, 'padding': 1},
                          upsample_class=M.BlurConvTranspose3d,
                          upsample_params={'kernel_size': 3, 'stride': 2, 'padding': 1, 'output_padding': 0})
    _perturb_bn(model, 1)
    return model.eval()


def test_config2_network_one_patch_against_oracle():
    """The msseg2 network on one 96^3 patch: fp32 path <= 1e-5, bf16 path <= 2e-2 of the CPU oracle."""
    from segmentation_pipeline.models import set_precision
    model = _msseg2_model()
    sd = {k: v.clone() for k, v in model.state_dict().items()}
    cfg = {"depth": 6, "filters": [40, 40, 80, 80, 120, 120], "block": {"residual": True}, "down": "blur", "up": "blur"}
    x = _smooth_volume((2, 96, 96, 96), 5)[None]
    torch.set_num_threads(max(torch.get_num_threads(), 8))
    with torch.no_grad():
        ref = unet.modular_unet_forward(sd, x, cfg)
    model.cuda()
    for precision, tol in (("fp32", 1e-5), ("bf16", 2e-2)):
        set_precision(precision)
        try:
            with torch.no_grad():
                out = model(x.cuda()).cpu()
        finally:
            set_precision("auto")
        assert rel_err(out, ref) <= tol, precision
        rep = label_agreement_report(ref, out, f"config 2 network, one patch, {precision}, random head")
        assert rep["agreement"] >= (0.9999 if precision == "fp32" else 0.99)


def test_config2_full_volume_properties():
    """Full 2 x 256 x 256 x 192 volume, patch 96^3, overlap 48, edge padding, bf16: size-independent checks."""
    import b200seg
    from segmentation_pipeline.prediction import PatchPredict
    from segmentation_pipeline.models import set_precision
    model = _msseg2_model().cuda()
    vol = _smooth_volume((2, 256, 256, 192), 7).cuda()
    predictor = PatchPredict(patch_batch_size=24, patch_size=96, patch_overlap=48, padding_mode="edge")
    set_precision("bf16")
    try:
        with torch.no_grad():
            probs, labels = predictor.predict_volume(model, vol)
            probs2, labels2 = predictor.predict_volume(model, vol)
    finally:
        set_precision("auto")
    assert probs.shape == (2, 256, 256, 192) and labels.shape == (256, 256, 192)
    assert torch.equal(probs, probs2) and torch.equal(labels, labels2)              # deterministic (no atomics)
    assert torch.isfinite(probs).all()
    assert (probs.sum(0) - 1).abs().max().item() <= 1e-5                            # averaged softmax rows sum to 1
    assert torch.equal(labels.long(), probs.argmax(0))                              # fused argmax == argmax
    target = (torch.rand(labels.shape, device="cuda") > 0.5).to(torch.uint8)
    cm = torch.zeros((2, 2), dtype=torch.int64, device="cuda")
    b200seg.confusion(labels, target, 2, cm)
    assert int(cm.sum()) == labels.numel()
    assert int(cm[:, 1].sum()) == int((labels == 1).sum()) and int(cm[1].sum()) == int(target.sum())


def test_full_size_aggregation_identity_round_trip():
    """Extract -> overlap-add -> finalize of the patches themselves returns the volume bit-exactly wherever
    one patch covers a voxel and to 1 ulp elsewhere (config 2 geometry, 144 patches)."""
    import b200seg
    from segmentation_pipeline.grid import PatchGrid
    vol = torch.randn(2, 256, 256, 192, generator=torch.Generator().manual_seed(9)).cuda()
    grid = PatchGrid(vol.shape[1:], 96, 48, "edge")
    assert len(grid.locations) == 144
    out = torch.zeros((2, *grid.padded_shape), device="cuda")
    for locs in grid.batches(36):
        buf = b200seg.Blocked(len(locs), 1, 96, 96, 96, torch.float32, "cuda")
        b200seg.grid_extract(vol, locs, grid.border, 1, 0.0, buf.view(2))
        patches = torch.empty((len(locs), 2, 96, 96, 96), device="cuda")
        b200seg.unpack_ncdhw(buf.view(2), patches)
        b200seg.overlap_add(out, patches, locs)
    counts = [torch.tensor(c, dtype=torch.int32, device="cuda") for c in grid.axis_counts()]
    probs = torch.empty_like(vol)
    b200seg.finalize(out, counts, grid.border, probs, None, None)
    assert (probs - vol).abs().max().item() <= 1e-6 * vol.abs().max().item()
    once = (counts[0][24:-24, None, None] * counts[1][None, 24:-24, None] * counts[2][None, None, 24:-24]) == 1
    assert torch.equal(probs[:, once], vol[:, once])


def test_config3_nested_10class_sliding_window_against_oracle():
    """NestedResUNet(2, 10, 8) (qsm shape, shrunk filters), edge-padded sliding window vs the oracle."""
    from segmentation_pipeline import models as M
    from segmentation_pipeline.models import set_precision
    from segmentation_pipeline.prediction import PatchPredict
    torch.manual_seed(3)
    model = M.NestedResUNet(2, 10, 8)
    _perturb_bn(model, 4)
    model.eval()
    sd = {k: v.clone() for k, v in model.state_dict().items()}
    vol = _smooth_volume((2, 40, 32, 24), 11)
    ref = ogrid.sliding_window(vol.numpy(), lambda p: unet.nested_res_unet_forward(sd, torch.from_numpy(p)).numpy(),
                               16, 8, "edge", "average", patch_batch_size=7)
    model.cuda()
    predictor = PatchPredict(patch_batch_size=7, patch_size=16, patch_overlap=8, padding_mode="edge")
    for precision, tol in (("fp32", 1e-5), ("bf16", 2e-2)):
        set_precision(precision)
        try:
            with torch.no_grad():
                probs, labels = predictor.predict_volume(model, vol.cuda())
        finally:
            set_precision("auto")
        assert probs.shape[0] == 10
        assert rel_err(probs.cpu(), torch.from_numpy(ref)) <= tol, precision


def test_config4_full_width_standard_predict_sagittal_split():
    """BASELINE config 4 at its real width: NestedResUNet(3, 2, 40, dropout 0.2) in eval mode on 3 x 96 x 88 x 24 whole
    volumes through StandardPredict(sagittal_split=True) (reference prediction.py:73-102), vs the oracle."""
    from segmentation_pipeline import _tio, models as M
    from segmentation_pipeline.models import set_precision
    from segmentation_pipeline.prediction import StandardPredict
    torch.manual_seed(11)
    model = M.NestedResUNet(3, 2, 40, dropout_p=0.2)
    _perturb_bn(model, 12)
    model.eval()
    sd = {k: v.clone() for k, v in model.state_dict().items()}
    g = torch.Generator().manual_seed(13)
    vols = [torch.randn(3, 96, 88, 24, generator=g) for _ in range(2)]
    torch.set_num_threads(max(torch.get_num_threads(), 8))
    with torch.no_grad():
        ref = unet.reverse_split_and_flip(unet.nested_res_unet_forward(sd, unet.split_and_flip(torch.stack(vols))))
    model.cuda()
    for precision, tol in (("fp32", 1e-5), ("bf16", 2e-2)):
        subjects = [_tio.Subject(X=_tio.ScalarImage(tensor=v), name=f"s{i}") for i, v in enumerate(vols)]
        set_precision(precision)
        try:
            for _ in range(3):          # third call replays the captured CUDA graph of the plan
                out, batch = StandardPredict(sagittal_split=True).predict(model, torch.device("cuda"), subjects)
        finally:
            set_precision("auto")
        assert batch["y_pred"].shape == (2, 2, 96, 88, 24)
        assert rel_err(batch["y_pred"].cpu(), ref) <= tol, precision
        assert rel_err(out[1]["y_pred"]["data"], ref[1]) <= tol
