"""GPU parity of whole networks and of the sliding-window predictor against the committed golden fixtures
(produced by the reference's own modules) and the CPU oracle.  Tolerances are the ones BASELINE.json states:
fp32 path max rel err <= 1e-5, bf16 path <= 2e-2 on the network outputs; labels >= 99.9 % agreement in bf16;
counts / Dice bit-exact given identical labels."""
import numpy as np
import pytest
import torch

from oracle import evalstats, grid as ogrid, unet
from helpers import GOLDEN, label_agreement_report, load_case, rel_err

pytestmark = pytest.mark.gpu


def build_model(meta):
    from torch import nn
    from segmentation_pipeline import models as M
    if meta["class"] == "NestedResUNet":
        return M.NestedResUNet(meta["input_channels"], meta["output_channels"], meta["filters"],
                               dropout_p=meta.get("dropout_p", 0.0))
    block = meta["block"]
    bp = {"residual": block.get("residual", False)}
    if block.get("norm") == "none":
        bp["normalization_class"] = None
    if block.get("norm") == "instance":
        bp["normalization_class"] = nn.InstanceNorm3d
    if block.get("act") == "leaky_relu":
        bp["activation_class"] = nn.LeakyReLU
        bp["activation_params"] = {"negative_slope": block["slope"]}
    if block.get("conv") == "ws":
        bp["conv_class"] = M.WSConv3d
        bp["conv_params"] = {"kernel_size": 3, "padding": 1}
    kw = {}
    if meta["down"] == "blur":
        kw.update(downsample_class=M.BlurConv3d, downsample_params={"kernel_size": 3, "stride": 2, "padding": 1},
                  upsample_class=M.BlurConvTranspose3d,
                  upsample_params={"kernel_size": 3, "stride": 2, "padding": 1, "output_padding": 0})
    if meta["hypothesis"] == "identity":
        kw.update(hypothesis_class=nn.Identity, hypothesis_params={})
    return M.ModularUNet(meta["in_channels"], meta["out_channels"], meta["filters"], meta["depth"], block_params=bp,
                         **kw)


FP32_CASES = ["models_modular_blur", "models_modular_default", "models_modular_leaky_logits", "models_nested",
              "models_nested_10class", "models_modular_ws_instnorm"]


@pytest.mark.parametrize("name", FP32_CASES)
def test_fp32_network_matches_reference_fixture(name):
    from segmentation_pipeline.models import set_precision
    meta, sd, x, y = load_case(name)
    model = build_model(meta)
    model.load_state_dict(sd, strict=True)
    model.eval().cuda()
    set_precision("fp32")
    try:
        with torch.no_grad():
            out = model(x.cuda())
    finally:
        set_precision("auto")
    assert out.shape == y.shape and out.dtype == torch.float32
    assert rel_err(out.cpu(), y) <= 1e-5


@pytest.mark.parametrize("name", ["models_modular_blur", "models_modular_default", "models_modular_leaky_logits",
                                  "models_nested", "models_nested_10class", "models_modular_ws_instnorm"])
def test_bf16_network_within_tolerance(name):
    from segmentation_pipeline.models import set_precision
    meta, sd, x, y = load_case(name)
    model = build_model(meta)
    model.load_state_dict(sd, strict=True)
    model.eval().cuda()
    set_precision("bf16")
    try:
        with torch.no_grad():
            out = model(x.cuda())
    finally:
        set_precision("auto")
    assert rel_err(out.cpu(), y) <= 2e-2
    if meta["hypothesis"] == "softmax":
        # Unfiltered label agreement.  These fixtures are tiny RANDOM-INIT networks (p ~ 1/C everywhere), so the
        # argmax hangs on the last bits of the logits: the >= 99.9 % criterion is asserted on confident heads in
        # tests/test_gpu_labels.py; here every voxel is counted and a label may differ only where the reference's own
        # top-2 margin is below twice the probability error.
        rep = label_agreement_report(y, out.cpu(), f"{name} (random head, bf16)")
        assert rep["agreement"] >= 0.95
        assert rep["disagree_margin_max"] <= 2 * (out.cpu() - y).abs().max().item() + 1e-7


@pytest.mark.parametrize("head", ["softmax17", "stochastic5"])
def test_wide_heads_go_through_the_blocked_out_buffer(head):
    """More than 16 output channels (17 classes; a StochasticMatrix with C = 5 -> 25 channels): the last conv writes a
    blocked buffer, the engine unpacks it and runs the separate softmax pass (these branches used to fail)."""
    from segmentation_pipeline import models as M
    from segmentation_pipeline.models import set_precision
    torch.manual_seed(5)
    if head == "softmax17":
        model = M.ModularUNet(1, 17, [8, 8], 2)
        cfg = {"hypothesis": "softmax"}
    else:
        model = M.ModularUNet(1, 25, [8, 8], 2, hypothesis_class=M.StochasticMatrix,
                              hypothesis_params={"channels": 5, "diag_bias": 2.0})
        cfg = {"hypothesis": "stochastic_matrix", "sm_channels": 5, "sm_diag_bias": 2.0}
    model.eval()
    sd = {k: v.clone() for k, v in model.state_dict().items()}
    cfg.update({"depth": 2, "filters": [8, 8], "block": {"residual": False}, "down": "avgpool", "up": "trilinear"})
    x = torch.randn(2, 1, 16, 8, 8, generator=torch.Generator().manual_seed(6))
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
        assert out.shape == ref.shape
        assert rel_err(out, ref) <= tol, (head, precision)


def test_instance_norm_residual_block_keeps_conv0_res_fusion():
    """InstanceNorm3d blocks with a residual branch: conv0 and res_conv still run as ONE contraction (two destinations),
    followed by the in-place norm kernel; vs the oracle (fp32 <= 1e-5, bf16 <= 2e-2)."""
    from torch import nn
    from segmentation_pipeline import models as M
    from segmentation_pipeline.models import _engine, set_precision
    torch.manual_seed(15)
    model = M.ModularUNet(2, 2, [8, 16], 2, block_params={"residual": True, "normalization_class": nn.InstanceNorm3d,
                                                          "activation_class": nn.LeakyReLU,
                                                          "activation_params": {"negative_slope": 0.1}})
    model.eval()
    names = [getattr(op, "name", "") for op in _engine.lower(model).ops]
    assert "down0.conv0+res" in names and "down0.res_conv" not in names
    sd = {k: v.clone() for k, v in model.state_dict().items()}
    cfg = {"depth": 2, "filters": [8, 16], "block": {"residual": True, "norm": "instance", "act": "leaky_relu", "slope": 0.1},
           "down": "avgpool", "up": "trilinear"}
    x = torch.randn(2, 2, 16, 16, 8, generator=torch.Generator().manual_seed(16))
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


def test_other_device_than_current(monkeypatch):
    """predict(model, device, ...) with device != the current CUDA device.  Needs two GPUs; on one GPU
    the guard itself is checked: a tensor from another device index is refused loudly."""
    import b200seg
    if torch.cuda.device_count() < 2:
        t = torch.zeros(16, device="cuda:0")
        monkeypatch.setattr(torch.cuda, "current_device", lambda: 1)
        with pytest.raises(RuntimeError, match="current device"):
            b200seg._require_cuda(t)
        return
    meta, sd, x, y = load_case("models_modular_default")
    model = build_model(meta)
    model.load_state_dict(sd)
    model.eval().to("cuda:1")
    assert torch.cuda.current_device() == 0
    with torch.no_grad():
        out = model(x.to("cuda:1"))
    assert out.device == torch.device("cuda:1") and rel_err(out.cpu(), y) <= 1e-5


def test_invalidate_after_data_edit():
    meta, sd, x, y = load_case("models_modular_default")
    model = build_model(meta)
    model.load_state_dict(sd)
    model.eval().cuda()
    with torch.no_grad():
        a = model(x.cuda())
        model.out_conv.weight.data.mul_(2.0)          # .data edit: no version bump
        model.invalidate()
        b = model(x.cuda())
    assert not torch.equal(a, b)
    sd2 = dict(sd)
    sd2["out_conv.weight"] = sd["out_conv.weight"] * 2.0
    cfg = dict(meta)
    with torch.no_grad():
        ref = unet.modular_unet_forward(sd2, x, cfg)
    assert rel_err(b.cpu(), ref) <= 1e-5


def test_autocast_selects_bf16_path():
    meta, sd, x, y = load_case("models_modular_blur")
    model = build_model(meta)
    model.load_state_dict(sd)
    model.eval().cuda()
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
        out = model(x.cuda())
    assert out.dtype == torch.float32
    assert 1e-7 < rel_err(out.cpu(), y) <= 2e-2     # really the bf16 path, and within tolerance


def test_unsupported_and_training_mode_raise():
    from torch import nn
    from segmentation_pipeline import models as M
    model = M.ModularUNet(1, 2, [8, 8], 2, block_params={"activation_class": nn.Tanh, "activation_params": {}})
    model.eval().cuda()
    with pytest.raises(NotImplementedError):
        model(torch.zeros(1, 1, 8, 8, 8).cuda())     # Tanh is not lowered: loud error, no silent ATen path
    meta, sd, x, y = load_case("models_modular_default")
    model = build_model(meta).cuda()
    with pytest.raises(RuntimeError):
        model.eval()(x)           # CPU tensor
    # training mode: ModularUNet is lowered (tests/test_gpu_train.py); what is not raises instead of falling back
    inorm = M.ModularUNet(1, 2, [8, 8], 2, block_params={"normalization_class": nn.InstanceNorm3d}).cuda().train()
    with pytest.raises(NotImplementedError):
        inorm(torch.zeros(1, 1, 8, 8, 8).cuda())
    tanh = M.ModularUNet(1, 2, [8, 8], 2, block_params={"activation_class": nn.Tanh, "activation_params": {}}).cuda().train()
    with pytest.raises(NotImplementedError):
        tanh(torch.zeros(1, 1, 8, 8, 8).cuda())


def test_components_forward():
    from segmentation_pipeline import models as M
    from segmentation_pipeline.models import set_precision
    z = np.load(f"{GOLDEN}/components.npz")
    t = lambda k: torch.from_numpy(z[k])
    set_precision("fp32")
    try:
        bc = M.BlurConv3d(8, 8, kernel_size=3, stride=2, padding=1)
        bc.weight.data.copy_(t("blur_w"))
        assert rel_err(bc.eval().cuda()(t("blur_x").cuda()).cpu(), t("blur_y")) <= 1e-5
        bt = M.BlurConvTranspose3d(8, 8, kernel_size=3, stride=2, padding=1, output_padding=0)
        bt.weight.data.copy_(t("blurT_w"))
        assert rel_err(bt.eval().cuda()(t("blur_x").cuda()).cpu(), t("blurT_y")) <= 1e-5
        sm = M.StochasticMatrix(2, diag_bias=1.5)
        assert rel_err(sm(t("sm_x").cuda()).cpu(), t("sm_y")) <= 1e-6
        # ensembles around a native member
        base = M.ModularUNet(1, 2, [8, 8], 2)
        base.load_state_dict({k[7:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("ens_sd/")})
        base.eval().cuda()
        x = t("ens_x").cuda()
        with torch.no_grad():
            assert rel_err(M.EnsembleFlips(base, "mean")(x).cpu(), t("ens_flips_mean")) <= 1e-5
            assert rel_err(M.EnsembleOrientations(base, "mean")(x).cpu(), t("ens_orient_mean")) <= 1e-5
            maj = M.EnsembleFlips(base, "majority")(x).cpu()
            # two different members: EnsembleModels, fused, vs the oracle on both state_dicts
            other = M.ModularUNet(1, 2, [8, 8], 2)
            torch.manual_seed(9)
            for prm in other.parameters():
                torch.nn.init.normal_(prm, std=0.2)
            other.eval().cuda()
            got = M.EnsembleModels([base, other], "mean")(x).cpu()
        assert (maj.numpy() == z["ens_flips_majority"]).mean() >= 0.999
        assert maj.dtype == torch.int64
        cfg = {"depth": 2, "filters": [8, 8], "block": {"residual": False}, "down": "avgpool", "up": "trilinear"}
        sds = [{k: v.detach().cpu() for k, v in m.state_dict().items()} for m in (base, other)]
        want = unet.ensemble_models([lambda t, sd=sd: unet.modular_unet_forward(sd, t, cfg) for sd in sds], t("ens_x"))
        assert rel_err(got, want) <= 1e-5
    finally:
        set_precision("auto")


# ----------------------------------------------------------------------------------------------- predictors
def _small_model():
    meta, sd, _, _ = load_case("models_modular_blur")
    model = build_model(meta)
    model.load_state_dict(sd)
    return meta, sd, model.eval().cuda()


@pytest.mark.parametrize("device_batch", [0, 48])     # literal patch_batch_size (ragged batches) / regrouped on the device
@pytest.mark.parametrize("padding_mode,overlap_mode", [(None, "average"), ("edge", "average"), ("edge", "crop"),
                                                       (None, "hann"), ("edge", "hann")])
def test_patch_predict_matches_oracle_sliding_window(padding_mode, overlap_mode, device_batch, monkeypatch):
    from segmentation_pipeline import _tio, prediction
    from segmentation_pipeline.models import set_precision
    from segmentation_pipeline.prediction import PatchPredict, add_evaluation_labels
    monkeypatch.setattr(prediction, "DEVICE_BATCH", [device_batch])
    meta, sd, model = _small_model()
    g = torch.Generator().manual_seed(21)
    vol = torch.randn(2, 40, 36, 28, generator=g)
    ref = ogrid.sliding_window(vol.numpy(), lambda p: unet.modular_unet_forward(sd, torch.from_numpy(p), meta).numpy(),
                               (16, 16, 16), (8, 8, 4), padding_mode, overlap_mode, patch_batch_size=5)
    subject = _tio.Subject(X=_tio.ScalarImage(tensor=vol), name="s0")
    predictor = PatchPredict(patch_batch_size=5, patch_size=(16, 16, 16), patch_overlap=(8, 8, 4),
                             padding_mode=padding_mode, overlap_mode=overlap_mode)
    set_precision("fp32")
    try:
        subjects, batch = predictor.predict(model, torch.device("cuda"), [subject], {"label_values": {"lesion": 1}})
    finally:
        set_precision("auto")
    y_pred = subjects[0]["y_pred"]
    assert y_pred["data"].device.type == "cpu" and y_pred["label_values"] == {"lesion": 1}
    assert batch["X"].is_cuda and batch["y_pred"].shape == (1, 2, 40, 36, 28)
    assert rel_err(y_pred["data"], torch.from_numpy(ref)) <= 1e-5
    add_evaluation_labels(subjects)
    lab = subjects[0]["y_pred_eval"]["data"]
    assert lab.dtype == torch.int64 and lab.shape == (1, 40, 36, 28)
    # labels are the argmax of the probabilities we returned, bit-exactly (ties -> lowest index)
    np.testing.assert_array_equal(lab.numpy(), evalstats.argmax_labels(y_pred["data"].numpy()))


def test_device_regrouping_does_not_change_a_bit(monkeypatch):
    """patch_batch_size is a memory knob: 1, 5 (ragged) or regrouped to 48 on the device give identical outputs."""
    from segmentation_pipeline import prediction
    from segmentation_pipeline.models import set_precision
    from segmentation_pipeline.prediction import PatchPredict
    meta, sd, model = _small_model()
    vol = torch.randn(2, 40, 36, 28, generator=torch.Generator().manual_seed(24)).cuda()
    outs = []
    set_precision("bf16")
    try:
        with torch.no_grad():
            for batch, device_batch in ((1, 0), (5, 0), (1, 48), (16, 48)):
                monkeypatch.setattr(prediction, "DEVICE_BATCH", [device_batch])
                p = PatchPredict(patch_batch_size=batch, patch_size=(16, 16, 16), patch_overlap=(8, 8, 4),
                                 padding_mode="edge")
                for _ in range(3 if batch == 1 else 1):          # batch 1 literal: third call replays the CUDA graph
                    probs, labels = p.predict_volume(model, vol)
                outs.append((probs.clone(), labels.clone()))
    finally:
        set_precision("auto")
    for probs, labels in outs[1:]:
        assert torch.equal(probs, outs[0][0]) and torch.equal(labels, outs[0][1])


def test_patch_predict_config_attributes():
    from segmentation_pipeline.prediction import PatchPredict, StandardPredict
    p = PatchPredict(patch_batch_size=32, patch_size=96, patch_overlap=12, padding_mode=None,
                     overlap_mode='average', image_names=['X'])
    assert p.get_config() == dict(image_names=['X'], patch_batch_size=32, patch_size=96, patch_overlap=12,
                                  padding_mode=None, overlap_mode='average')
    assert StandardPredict(sagittal_split=True).get_config()["sagittal_split"] is True


def test_standard_predict_sagittal_split():
    from segmentation_pipeline import _tio, models as M
    from segmentation_pipeline.models import set_precision
    from segmentation_pipeline.prediction import StandardPredict
    meta, sd, x, _ = load_case("models_nested")
    model = M.NestedResUNet(3, 2, 8, dropout_p=0.2)
    model.load_state_dict(sd)
    model.eval().cuda()
    g = torch.Generator().manual_seed(22)
    vols = [torch.randn(3, 32, 16, 8, generator=g) for _ in range(2)]
    ref = unet.reverse_split_and_flip(unet.nested_res_unet_forward(sd, unet.split_and_flip(torch.stack(vols))))
    subjects = [_tio.Subject(X=_tio.ScalarImage(tensor=v), name=f"s{i}") for i, v in enumerate(vols)]
    set_precision("fp32")
    try:
        out, batch = StandardPredict(sagittal_split=True).predict(model, torch.device("cuda"), subjects)
    finally:
        set_precision("auto")
    assert rel_err(batch["y_pred"].cpu(), ref) <= 1e-5
    assert rel_err(out[1]["y_pred"]["data"], ref[1]) <= 1e-5


def test_segmentation_evaluator_matches_reference_fixture():
    from segmentation_pipeline import _tio
    from segmentation_pipeline.evaluators import LabelMapEvaluator, SegmentationEvaluator
    z = np.load(f"{GOLDEN}/evaluator.npz")
    label_values = dict(zip([str(s) for s in z["label_names"]], [int(v) for v in z["label_vals"]]))
    stats = [str(s) for s in z["stats"]]
    subjects = []
    for i in range(3):
        subjects.append(_tio.Subject(
            name=f"s{i}",
            pred=_tio.LabelMap(tensor=torch.from_numpy(z[f"pred{i}"]), label_values=label_values),
            targ=_tio.LabelMap(tensor=torch.from_numpy(z[f"targ{i}"]), label_values=label_values)))
    res = SegmentationEvaluator("pred", "targ", stats_to_output=stats)(subjects)
    got = res["subject_stats"][stats].to_numpy(dtype=np.float64)
    exp = z["subject_stats_values"]
    assert ((got == exp) | (np.isnan(got) & np.isnan(exp))).all()
    np.testing.assert_array_equal(res["summary_stats"].data.numpy(), z["summary_stats"])
    vol = LabelMapEvaluator("pred")(subjects)["subject_stats"][["volume"]].to_numpy(dtype=np.float64)
    np.testing.assert_array_equal(vol, z["volumes"])


def test_segmentation_evaluator_with_labels_outside_label_values():
    """label_values lists a SUBSET of the labels present: a voxel with target == v and an unlisted or
    larger prediction is a false negative, as the reference's boolean masks count it (segmentation_evaluator.py:69-77)."""
    from segmentation_pipeline import _tio
    from segmentation_pipeline.evaluators import LabelMapEvaluator, SegmentationEvaluator
    rng = np.random.default_rng(31)
    shape = (1, 24, 20, 18)
    pred = rng.integers(0, 7, size=shape)        # labels 0..6 present
    targ = rng.integers(0, 7, size=shape)
    label_values = {"a": 1, "c": 3}              # only two of them are evaluated; 4, 5, 6 exceed max(label_values)
    stats = list(evalstats.STATS)
    subject = _tio.Subject(name="s", pred=_tio.LabelMap(tensor=torch.from_numpy(pred), label_values=label_values),
                           targ=_tio.LabelMap(tensor=torch.from_numpy(targ), label_values=label_values))
    res = SegmentationEvaluator("pred", "targ", stats_to_output=stats)([subject])
    ref = evalstats.segmentation_stats(pred, targ, label_values, stats)
    for i, name in enumerate(label_values):
        row = res["subject_stats"].iloc[i]
        for st in stats:
            assert float(row[st]) == ref[name][st], (name, st)
    vol = LabelMapEvaluator("pred")([subject])["subject_stats"]["volume"].to_numpy()
    assert list(vol) == list(evalstats.label_volumes(pred, label_values).values())


def test_instance_segmentation_evaluator_matches_oracle():
    """InstanceSegmentationEvaluator (MSSEG lesion-detection metric): components + overlap table on the device, every
    statistic equal to the oracle restatement of instance_segmentation_evaluator.py:103-160."""
    from test_gpu_kernels import _blob_mask
    from segmentation_pipeline import _tio
    from segmentation_pipeline.evaluators import InstanceSegmentationEvaluator
    subjects, want = [], {}
    for i in range(3):
        targ = _blob_mask((48, 40, 36), 14, 90 + i, big=True)
        pred = np.roll(targ, (1, 2, 0), (0, 1, 2)) | _blob_mask((48, 40, 36), 4, 95 + i, big=True)   # shifted + extra lesions
        if i == 2:
            pred = targ.copy()                                                                # perfect prediction
        subjects.append(_tio.Subject(name=f"s{i}", pred=_tio.LabelMap(tensor=torch.from_numpy(pred[None].astype(np.int64))),
                                     targ=_tio.LabelMap(tensor=torch.from_numpy(targ[None].astype(np.int64)))))
        want[f"s{i}"] = evalstats.instance_stats(pred, targ, 2)
    ev = InstanceSegmentationEvaluator("pred", "targ")
    res = ev(subjects)["subject_stats"]
    for i in range(3):
        row = res.iloc[i]
        for stat in ev.stats_to_output:
            a, b = float(row[stat]), float(want[f"s{i}"][stat])
            assert a == b or (np.isnan(a) and np.isnan(b)), (i, stat, a, b)
    assert float(res.iloc[2]["detection_f1"]) == 1.0 and float(res.iloc[2]["dice"]) == 1.0
    assert float(res.iloc[0]["target_components"]) > 5


def test_slab_mode_single_rank_equals_patch_predict():
    """z-slab driver with the CUDA ops on one rank == PatchPredict.predict_volume, bit for bit."""
    from segmentation_pipeline.distributed import CudaSlabOps, slab_predict
    from segmentation_pipeline.grid import PatchGrid
    from segmentation_pipeline.models import set_precision
    from segmentation_pipeline.prediction import PatchPredict
    meta, sd, model = _small_model()
    vol = torch.randn(2, 40, 36, 28, generator=torch.Generator().manual_seed(23)).cuda()
    set_precision("bf16")
    try:
        with torch.no_grad():
            predictor = PatchPredict(patch_batch_size=5, patch_size=(16, 16, 16), patch_overlap=(8, 8, 4),
                                     padding_mode="edge")
            probs, labels = predictor.predict_volume(model, vol)
            grid = PatchGrid(vol.shape[1:], (16, 16, 16), (8, 8, 4), "edge")
            labels2, probs2 = slab_predict(vol, grid, CudaSlabOps(model, 5), gather_probs=True)
    finally:
        set_precision("auto")
    assert torch.equal(labels, labels2) and torch.equal(probs, probs2)


def test_label_map_evaluator_growth_curve_statistics():
    """LabelMapEvaluator with ``curve_params`` / ``curve_attribute`` (reference label_map_evaluator.py:83-99): volume
    error against a per-label polynomial of a subject attribute; missing arguments raise like the reference."""
    from segmentation_pipeline import _tio
    from segmentation_pipeline.evaluators import LabelMapEvaluator
    rng = np.random.default_rng(5)
    label_values = {"left": 1, "right": 2}
    curve = {"left": np.array([2.0, 100.0]), "right": np.array([0.5, -3.0, 400.0])}
    stats = ("volume", "error", "absolute_error", "squared_error", "percent_diff")
    subjects, maps = [], []
    for i, age in enumerate((20.0, 63.5)):
        data = rng.integers(0, 3, size=(1, 16, 12, 10))
        maps.append(data)
        subjects.append(_tio.Subject(name=f"s{i}", age=age,
                                     pred=_tio.LabelMap(tensor=torch.from_numpy(data), label_values=label_values)))
    res = LabelMapEvaluator("pred", curve_params=curve, curve_attribute="age", stats_to_output=stats)(subjects)
    frame = res["subject_stats"]
    row = 0
    for i, subject in enumerate(subjects):
        for name, value in label_values.items():
            volume = torch.tensor(int((maps[i] == value).sum()))
            expected = np.poly1d(curve[name])(subject["age"])
            err = volume - expected
            want = {"volume": volume, "error": err, "absolute_error": abs(err), "squared_error": err ** 2,
                    "percent_diff": (err / expected) * 100}
            for st in stats:
                assert float(frame.iloc[row][st]) == float(want[st].item()), (i, name, st)
            row += 1
    with pytest.raises(ValueError):
        LabelMapEvaluator("pred", stats_to_output=("volume", "error"))
    with pytest.raises(ValueError):
        LabelMapEvaluator("pred", curve_params=curve, stats_to_output=("percent_diff",))
