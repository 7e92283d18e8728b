"""CPU tests of the host-side mirror: constructor / state_dict compatibility with the reference fixtures, the
patch grid against the oracle, lowering of the networks to plans, and the C-ABI symbol table."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

from oracle import grid as ogrid
from helpers import ROOT, load_case


def test_state_dict_keys_and_shapes_match_reference():
    from test_gpu_models import build_model
    for name in ["models_modular_blur", "models_modular_default", "models_modular_leaky_logits",
                 "models_modular_ws_instnorm", "models_nested", "models_nested_10class"]:
        meta, sd, _, _ = load_case(name)
        model = build_model(meta)
        own = model.state_dict()
        assert list(own.keys()) == list(sd.keys()), name      # same keys in the same order
        for k in sd:
            assert tuple(own[k].shape) == tuple(sd[k].shape), (name, k)
        model.load_state_dict(sd, strict=True)


def test_random_init_matches_reference_when_available():
    from oracle import ref_import
    if not ref_import.reference_available():
        pytest.skip("reference tree not present")
    ref_models = ref_import.load_reference_models()
    from segmentation_pipeline import models as M
    torch.manual_seed(0)
    a = ref_models.NestedResUNet(3, 2, 8)
    torch.manual_seed(0)
    b = M.NestedResUNet(3, 2, 8)
    for (ka, va), (kb, vb) in zip(a.state_dict().items(), b.state_dict().items()):
        assert ka == kb and torch.equal(va, vb)
    kw = dict(block_params={'residual': True}, downsample_params={'kernel_size': 3, 'stride': 2, 'padding': 1},
              upsample_params={'kernel_size': 3, 'stride': 2, 'padding': 1, 'output_padding': 0})
    torch.manual_seed(1)
    a = ref_models.ModularUNet(2, 2, [8, 8, 16], 3, downsample_class=ref_models.BlurConv3d,
                               upsample_class=ref_models.BlurConvTranspose3d, **{k: dict(v) for k, v in kw.items()})
    torch.manual_seed(1)
    b = M.ModularUNet(2, 2, [8, 8, 16], 3, downsample_class=M.BlurConv3d, upsample_class=M.BlurConvTranspose3d,
                      **{k: dict(v) for k, v in kw.items()})
    for (ka, va), (kb, vb) in zip(a.state_dict().items(), b.state_dict().items()):
        assert ka == kb and torch.equal(va, vb)


def test_msseg2_model_parameter_count():
    from segmentation_pipeline import models as M
    m = M.ModularUNet(in_channels=2, out_channels=2, filters=[40, 40, 80, 80, 120, 120], depth=6,
                      block_params={'residual': True}, downsample_class=M.BlurConv3d,
                      downsample_params={'kernel_size': 3, 'stride': 2, 'padding': 1},
                      upsample_class=M.BlurConvTranspose3d,
                      upsample_params={'kernel_size': 3, 'stride': 2, 'padding': 1, 'output_padding': 0})
    assert sum(p.numel() for p in m.parameters()) == 9_472_282
    with pytest.raises(ValueError):
        M.ModularUNet(1, 2, [8, 8], 3)


def test_lowering_msseg2_plan_structure():
    from segmentation_pipeline import models as M
    from segmentation_pipeline.models import _engine, _plan
    m = M.ModularUNet(2, 2, [40, 40, 80, 80, 120, 120], 6, block_params={'residual': True},
                      downsample_class=M.BlurConv3d,
                      downsample_params={'kernel_size': 3, 'stride': 2, 'padding': 1},
                      upsample_class=M.BlurConvTranspose3d,
                      upsample_params={'kernel_size': 3, 'stride': 2, 'padding': 1, 'output_padding': 0}).eval()
    plan = _engine.lower(m)
    convs = [op for op in plan.ops if isinstance(op, _plan.ConvOp)]
    # 11 blocks x (conv0+res fused, conv1) + 5 down + 5 up + out_conv
    assert len(convs) == 11 * 2 + 5 + 5 + 1
    fused = [op for op in convs if op.dst1 is not None]
    assert len(fused) == 11 and all(op.cout == 2 * op.split for op in fused)
    assert plan.buffers["cat0"] == (10, 0) and plan.buffers["cat4"] == (30, 4)
    # decoder block 0 reads cat([up(40), skip(40)]) as one 80-channel contraction
    up0 = next(op for op in convs if op.name == "up0.conv0+res")
    assert up0.src == _plan.Ref("cat0", 0, 80) and up0.segments == [(0, 40), (5, 40)]
    assert convs[-1].final and convs[-1].softmax
    # BatchNorm folding: scale = gamma / sqrt(var + eps)
    bn = m.down_blocks[0].layers.norm1
    bn.running_var.fill_(3.0)
    bn.weight.data.fill_(2.0)
    bn.running_mean.fill_(0.5)
    bn.bias.data.fill_(-1.0)
    plan = _engine.lower(m)
    op = next(op for op in plan.ops if getattr(op, "name", "") == "down0.conv1")
    np.testing.assert_allclose(op.scale, 2.0 / np.sqrt(3.0 + 1e-5), rtol=1e-6)
    np.testing.assert_allclose(op.shift, -1.0 - 0.5 * op.scale, rtol=1e-6)
    assert op.residual == _plan.Ref("down0.res", 0, 40)


def test_lowering_nested_concat_order():
    from segmentation_pipeline import models as M
    from segmentation_pipeline.models import _engine, _plan
    plan = _engine.lower(M.NestedResUNet(3, 2, 8).eval())
    op = next(op for op in plan.ops if getattr(op, "name", "") == "conv1_1.conv0")
    # x1_1 = conv1_1(cat(x1_0, up(x2_0), down(x0_1))): skip first (nested_residual_unet.py:95)
    assert op.src == _plan.Ref("conv1_1.in", 0, 24) and op.segments == [(0, 8), (1, 8), (2, 8)]
    ups = [o for o in plan.ops if isinstance(o, _plan.UpsampleOp)]
    pools = [o for o in plan.ops if isinstance(o, _plan.PoolOp)]
    assert len(ups) == 6 and len(pools) == 6


@pytest.mark.parametrize("size,patch,overlap,padding", [
    ((96, 96, 96), 64, 16, None), ((256, 256, 192), 96, 48, "edge"), ((224, 224, 224), 96, 48, "edge"),
    ((20, 17, 13), (8, 8, 6), (4, 2, 2), None), ((20, 17, 13), (8, 8, 6), (4, 2, 2), 0.0)])
def test_patch_grid_matches_oracle(size, patch, overlap, padding):
    from segmentation_pipeline.grid import PatchGrid
    g = PatchGrid(size, patch, overlap, padding)
    border = ogrid.to_triple(overlap) // 2 if padding is not None else np.zeros(3, np.int64)
    padded = tuple(int(s + 2 * b) for s, b in zip(size, border))
    assert g.padded_shape == padded
    np.testing.assert_array_equal(np.array(g.locations), ogrid.grid_locations(padded, patch, overlap))
    ones = np.ones((len(g.locations), 1, *g.patch_size), np.float32)
    if np.prod(padded) < 1e6:
        _, cnt = ogrid.aggregate_average(ones, np.array(g.locations), padded)
        cw, ch, cd = (np.array(c) for c in g.axis_counts())
        np.testing.assert_array_equal(cnt[0], cw[:, None, None] * ch[None, :, None] * cd[None, None, :])


def test_patch_grid_errors():
    from segmentation_pipeline.grid import PatchGrid
    for args in [((10, 10, 10), 12, 0), ((10, 10, 10), 4, 4), ((10, 10, 10), 4, 1)]:
        with pytest.raises(ValueError):
            PatchGrid(*args)
    assert len(PatchGrid((304 - 48, 304 - 48, 240 - 48), 96, 48, "edge").locations) == 144


def test_forward_fails_loudly_without_gpu():
    from segmentation_pipeline import models as M
    m = M.ModularUNet(1, 2, [8, 8], 2).eval()
    with pytest.raises(RuntimeError, match="CUDA"):
        m(torch.zeros(1, 1, 8, 8, 8))


def test_c_abi_exports_every_declared_symbol():
    """The shared library loads without a GPU and exports exactly the entry points include/b200seg.h declares."""
    import b200seg
    header = open(os.path.join(ROOT, "include", "b200seg.h")).read()
    declared = set(re.findall(r"\b(b200seg_[a-z0-9_]+)\s*\(", header))
    assert declared == set(b200seg.EXPORTED_SYMBOLS)
    lib = b200seg.load_library()
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.b200seg_version() == 100
    assert lib.b200seg_conv3d_tc_wbytes(0, 5, 40) == 3 * 9 * 2 * (3 * 40 + 16) * 16
    # argument validation happens before any CUDA call
    assert lib.b200seg_conv3d_tc_wbytes(0, 5, 200) == -1
    assert b"cout" in lib.b200seg_last_error()
    # the engine-mode enum of the header, the Python packer and the C geometry agree (incl. the K3T final-layer mode)
    from segmentation_pipeline.models import _plan
    enum = re.search(r"enum \{ (B200SEG_TC_K3 = 0.*?) \};", header).group(1)
    values = dict((k.strip(), int(v)) for k, v in (item.split("=") for item in enum.split(",")))
    assert values == {"B200SEG_TC_K3": _plan.K3, "B200SEG_TC_DOWN": _plan.DOWN, "B200SEG_TC_UP": _plan.UP,
                      "B200SEG_TC_K3T": _plan.K3T}
    for mode, chunks, cout in [(_plan.K3, 5, 40), (_plan.DOWN, 5, 40), (_plan.UP, 10, 80), (_plan.K3T, 5, 2),
                               (_plan.K3T, 1, 4)]:
        g = _plan.tc_geometry(mode, chunks, cout)
        assert lib.b200seg_conv3d_tc_wbytes(mode, chunks, cout) == g["n_pass"] * g["n_bimg"] * g["bimg_elems"] * 2
    assert lib.b200seg_conv3d_tc_wbytes(_plan.K3T, 5, 5) == -1       # K3T serves at most 4 output channels


def test_training_spec_of_the_msseg2_network_and_refusals():
    """Host side of the training step (models/_train.py): the layer list and the flat parameter list autograd sees;
    the effective (blurred / standardised) weights are differentiable expressions of the module's parameters; module
    configurations outside the lowered space raise instead of falling back."""
    from segmentation_pipeline import models as M
    from segmentation_pipeline.models import _train
    model = M.ModularUNet(2, 2, [40, 40, 80, 80, 120, 120], 6, block_params={"residual": True},
                          downsample_class=M.BlurConv3d, downsample_params={"kernel_size": 3, "stride": 2, "padding": 1},
                          upsample_class=M.BlurConvTranspose3d,
                          upsample_params={"kernel_size": 3, "stride": 2, "padding": 1, "output_padding": 0})
    spec, params = _train.build_spec(model)
    assert spec["depth"] == 6 and len(spec["down"]) == 6 and len(spec["up"]) == 5
    assert [d["cout"] for d in spec["downs"]] == [40, 40, 80, 80, 120] and spec["ups"][0]["cin"] == 40
    assert spec["up"][4]["cin"] == 240 and spec["out"]["cout"] == 2 and spec["softmax"]
    # every parameter the reference's forward uses is reachable from the flat list (blur-conv biases are not: never applied)
    loss = sum((p.float() ** 2).sum() for p in params)
    loss.backward()
    unused = [n for n, p in model.named_parameters() if p.grad is None]
    assert sorted(unused) == sorted([f"downsampling.{i}.bias" for i in range(5)] + [f"upsampling.{i}.bias" for i in range(5)])
    blurred = params[spec["downs"][0]["w"]]
    assert blurred.shape == (40, 40, 4, 4, 4) and blurred.requires_grad and not blurred.is_leaf
    nested, nested_params = _train.build_spec(M.NestedResUNet(3, 2, 16, dropout_p=0.2))
    assert nested["kind"] == "nested" and nested["blocks"]["conv1_1"]["cin"] == 48 and len(nested_params) == 70
    assert nested["blocks"]["conv0_3"]["dropout_p"] == 0.2 and nested["blocks"]["conv0_1"]["res"] is not None
    assert _train.build_spec(M.ModularUNet(1, 2, [8, 8], 2, block_params={"dropout_p": 0.5}))[0]["down"][0]["dropout_p"] == 0.5
    with pytest.raises(NotImplementedError):
        _train.build_spec(M.NestedResUNet(1, 2, 12))                  # filters must be a multiple of 8
    with pytest.raises(NotImplementedError):
        _train.build_spec(M.ModularUNet(1, 2, [8, 8], 2, block_params={"normalization_class": torch.nn.InstanceNorm3d}))
    with pytest.raises(NotImplementedError):
        _train.build_spec(M.ModularUNet(1, 2, [12, 12], 2))          # concatenated widths must be multiples of 8
    default = _train.build_spec(M.ModularUNet(1, 2, [8, 16], 2))[0]
    assert default["downs"][0]["w"] is None and default["ups"][0]["w"] is None       # AvgPool3d / trilinear Upsample


def test_tc_gather_index_reproduces_the_host_packer():
    """The training step packs tensor-core operand images on the device through a cached gather index."""
    from segmentation_pipeline.models import _plan
    for mode, k, cin_chunks, cout in ((_plan.K3, 3, 5, 40), (_plan.K3, 3, 1, 80), (_plan.DOWN, 4, 10, 80),
                                      (_plan.UP, 4, 15, 120), (_plan.K3T, 3, 5, 2)):
        phys = torch.randn(k, k, k, cin_chunks * 8, _plan.c8(cout) * 8)
        want = _plan.pack_tc_weight(mode, phys, cin_chunks, cout)
        idx = _plan.tc_gather_index(mode, cin_chunks, cout)
        got = torch.where(idx >= 0, phys.reshape(-1)[idx.clamp(min=0)], torch.zeros(())).to(torch.bfloat16)
        assert got.shape == want.shape and bool((got == want).all())
