"""Training step on the device (reference segmentation_trainer.py:162-180): forward in
training mode (BatchNorm3d batch statistics), HybridLogisticDiceLoss, backward, SGD step -- against PyTorch autograd
over the CPU oracle (oracle/unet.py with ``bn_training``) on the same weights and inputs."""
import numpy as np
import pytest
import torch

from oracle import unet

pytestmark = pytest.mark.gpu


def assert_gradients_match(model, ref_sd, tight=2e-4, loose=3e-2, tight_fraction=0.9):
    """Per parameter tensor: e = max|g - g_ref| / max|g_ref| against autograd over the CPU oracle.  fp32 vs fp32: a ReLU
    whose pre-activation lies within rounding noise of zero may take the other branch on the two machines and move the
    gradients of the layers behind it by ~1 / sqrt(voxels) (measured: the CPU oracle in fp32 vs fp64 shows 4.6e-3 on one
    tensor of a 2 x 32^3 case).  Hence: at least 90 % of the tensors within ``tight`` (all of them
    on the boxes this was developed on), every tensor within ``loose``.  Returns the number of tensors compared."""
    errs = {}
    for name, p in model.named_parameters():
        want = ref_sd[name].grad
        if want is None:                       # parameters the reference's forward never applies (components.py:86,119)
            assert p.grad is None, name
            continue
        errs[name] = float((p.grad.cpu() - want).abs().max()) / (float(want.abs().max()) + 1e-12)
    worst = sorted(errs.items(), key=lambda kv: -kv[1])[:5]
    assert sum(e <= tight for e in errs.values()) >= tight_fraction * len(errs), worst
    assert worst[0][1] <= loose, worst
    return len(errs)


def _kernels_vs_torch_inputs(seed, n=2, c=12, ext=(6, 10, 9)):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(n, c, *ext, generator=g)


def _blocked(lib, t, dtype=torch.float32):
    n, c = t.shape[:2]
    buf = lib.Blocked(n, (c + 7) // 8, *t.shape[2:], dtype, t.device)
    lib.pack_ncdhw(t.contiguous(), buf.view(c))
    return buf


def _unblocked(lib, buf, c):
    out = torch.empty((buf.n, c, buf.z, buf.y, buf.x), dtype=torch.float32, device=buf.tensor.device)
    lib.unpack_ncdhw(buf.view(c), out)
    return out


def test_channel_moments_and_affine_act():
    import b200seg as lib
    x = (_kernels_vs_torch_inputs(0) * 3 + 1.5).cuda()
    buf = _blocked(lib, x)
    mean, var = lib.channel_moments(buf.view(12), 12, x.device)
    assert torch.allclose(mean[:12], x.mean(dim=(0, 2, 3, 4)), rtol=1e-5, atol=1e-6)
    assert torch.allclose(var[:12], x.var(dim=(0, 2, 3, 4), unbiased=False), rtol=1e-5, atol=1e-6)
    scale = torch.randn(16).cuda()
    shift = torch.randn(16).cuda()
    slope = torch.full((16,), 0.1).cuda()
    res = _kernels_vs_torch_inputs(1).cuda()
    dst = lib.Blocked(2, 2, 6, 10, 9, torch.float32, x.device)
    res_buf = _blocked(lib, res)           # keep the buffers alive: a View is a raw pointer
    lib.affine_act(buf.view(12), scale, shift, slope, dst.view(12), residual=res_buf.view(12))
    want = torch.nn.functional.leaky_relu(x * scale[:12].view(1, -1, 1, 1, 1) + shift[:12].view(1, -1, 1, 1, 1), 0.1) + res
    assert torch.allclose(_unblocked(lib, dst, 12), want, rtol=1e-6, atol=1e-6)


@pytest.mark.parametrize("slope,has_norm", [(0.0, True), (0.2, True), (1.0, True), (0.0, False)])
def test_bn_backward_matches_autograd(slope, has_norm):
    import b200seg as lib
    z = _kernels_vs_torch_inputs(2).cuda().requires_grad_(True)
    dy = _kernels_vs_torch_inputs(3).cuda()
    gamma = (torch.rand(12) + 0.5).cuda().requires_grad_(True)
    beta = torch.randn(12).cuda().requires_grad_(True)
    eps = 1e-5
    if has_norm:
        y = torch.nn.functional.batch_norm(z, None, None, gamma, beta, training=True, eps=eps)
    else:
        y = z
    a = torch.nn.functional.leaky_relu(y, slope) if slope != 1.0 else y
    a.backward(dy)
    mean = torch.zeros(16).cuda()
    rstd = torch.ones(16).cuda()
    scale = torch.ones(16).cuda()
    shift = torch.zeros(16).cuda()
    if has_norm:
        mean[:12] = z.detach().mean(dim=(0, 2, 3, 4))
        rstd[:12] = torch.rsqrt(z.detach().var(dim=(0, 2, 3, 4), unbiased=False) + eps)
        scale[:12] = gamma.detach() * rstd[:12]
        shift[:12] = beta.detach() - mean[:12] * scale[:12]
    dz = lib.Blocked(2, 2, 6, 10, 9, torch.float32, z.device)
    dy_buf, z_buf = _blocked(lib, dy), _blocked(lib, z.detach())
    sum_g, sum_gx = lib.bn_backward(dy_buf.view(12), z_buf.view(12), scale, shift,
                                    torch.full((16,), slope).cuda(), mean, rstd, has_norm, dz.view(12), 12, z.device)
    assert torch.allclose(_unblocked(lib, dz, 12), z.grad, rtol=1e-4, atol=1e-5)
    if has_norm:
        assert torch.allclose(sum_g[:12], beta.grad, rtol=1e-4, atol=1e-4)
        assert torch.allclose(sum_gx[:12], gamma.grad, rtol=1e-4, atol=1e-4)


@pytest.mark.parametrize("softmax", [True, False])
def test_softmax_backward(softmax):
    import b200seg as lib
    logits = _kernels_vs_torch_inputs(4, c=3).cuda().requires_grad_(True)
    dp = _kernels_vs_torch_inputs(5, c=3).cuda()
    p = torch.softmax(logits, dim=1) if softmax else logits * 1.0
    p.backward(dp)
    dst = lib.Blocked(2, 1, 6, 10, 9, torch.float32, dp.device)
    lib.softmax_backward(p.detach().contiguous(), dp, softmax, dst.view(3))
    assert torch.allclose(_unblocked(lib, dst, 3), logits.grad, rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("kind", ["k3", "down", "up"])
@pytest.mark.parametrize("cin,cout,ext", [(2, 16, (8, 8, 8)), (16, 24, (4, 6, 10)), (40, 40, (12, 12, 12)),
                                           (8, 16, (32, 32, 32)), (16, 8, (16, 16, 16)), (16, 16, (20, 36, 68))])
def test_wgrad_and_dgrad_match_autograd(kind, cin, cout, ext):
    """Weight and data gradients of the three convolution geometries vs torch.autograd (float64 on the CPU)."""
    import b200seg as lib
    from segmentation_pipeline.models import _train
    g = torch.Generator().manual_seed(cin * 100 + cout)
    n = 2
    x = torch.randn(n, cin, *ext, generator=g, dtype=torch.float64, requires_grad=True)
    runner = _train._Runner({}, torch.device("cuda"))
    if kind == "k3":
        w = torch.randn(cout, cin, 3, 3, 3, generator=g, dtype=torch.float64, requires_grad=True)
        y = torch.nn.functional.conv3d(x, w, padding=1)
    elif kind == "down":
        w = torch.randn(cout, cin, 4, 4, 4, generator=g, dtype=torch.float64, requires_grad=True)
        y = torch.nn.functional.conv3d(x, w, stride=2, padding=1)
    else:
        w = torch.randn(cin, cout, 4, 4, 4, generator=g, dtype=torch.float64, requires_grad=True)
        y = torch.nn.functional.conv_transpose3d(x, w, stride=2, padding=1)
    dy = torch.randn(y.shape, generator=g, dtype=torch.float64)
    y.backward(dy)
    xb = _blocked(lib, x.detach().float().cuda())
    dyb = _blocked(lib, dy.float().cuda())
    wf = w.detach().float().cuda()
    dx = lib.Blocked(n, (cin + 7) // 8, *ext, torch.float32, torch.device("cuda"))
    if kind == "k3":
        gw = runner._wgrad_conv(dyb.view(cout), xb.view(cin), cout, cin)
        runner.conv(dyb.view(cout), _train._pack(wf.flip(2, 3, 4).permute(2, 3, 4, 0, 1)), cin, dx.view(cin))
    elif kind == "down":
        gw = lib.wgrad(dyb.view(cout), xb.view(cin), 4, 2, 1, torch.device("cuda"))[:, :cout, :cin] \
            .permute(1, 2, 0).reshape(cout, cin, 4, 4, 4)
        runner.conv(dyb.view(cout), _train._pack(wf.permute(2, 3, 4, 0, 1)), cin, dx.view(cin), ksize=4, stride=2, pad=1,
                    transposed=True)
    else:
        gw = lib.wgrad(xb.view(cin), dyb.view(cout), 4, 2, 1, torch.device("cuda"))[:, :cin, :cout] \
            .permute(1, 2, 0).reshape(cin, cout, 4, 4, 4)
        runner.conv(dyb.view(cout), _train._pack(wf.permute(2, 3, 4, 1, 0)), cin, dx.view(cin), ksize=4, stride=2, pad=1)
    scale = float(w.grad.abs().max())
    assert float((gw.cpu().double() - w.grad).abs().max()) <= 2e-5 * scale
    scale = float(x.grad.abs().max())
    assert float((_unblocked(lib, dx, cin).cpu().double() - x.grad).abs().max()) <= 2e-5 * scale


@pytest.mark.parametrize("kind", ["k3", "down", "up"])
@pytest.mark.parametrize("cin,cout,ext", [(2, 16, (8, 8, 8)), (16, 24, (4, 6, 10)), (40, 40, (12, 12, 12)),
                                           (80, 40, (6, 20, 36)), (120, 56, (4, 4, 4))])
def test_tensor_core_wgrad_bf16(kind, cin, cout, ext):
    """b200seg_wgrad on bf16 views = the mma.sync kernel: bf16-exact inputs, fp32 accumulation -> equal to float64
    autograd up to fp32 summation error; several channel groups (> 48 / > 40 channels), rows that are not multiples
    of 16, all three geometries."""
    import b200seg as lib
    g = torch.Generator().manual_seed(cin * 100 + cout + 1)
    n = 2
    x = torch.randn(n, cin, *ext, generator=g).bfloat16().double().requires_grad_(True)
    if kind == "k3":
        w = torch.zeros(cout, cin, 3, 3, 3, dtype=torch.float64, requires_grad=True)
        y = torch.nn.functional.conv3d(x, w, padding=1)
    elif kind == "down":
        w = torch.zeros(cout, cin, 4, 4, 4, dtype=torch.float64, requires_grad=True)
        y = torch.nn.functional.conv3d(x, w, stride=2, padding=1)
    else:
        w = torch.zeros(cin, cout, 4, 4, 4, dtype=torch.float64, requires_grad=True)
        y = torch.nn.functional.conv_transpose3d(x, w, stride=2, padding=1)
    dy = torch.randn(y.shape, generator=g).bfloat16().double()
    y.backward(dy)
    dev = torch.device("cuda")
    xb = _blocked(lib, x.detach().float().cuda(), torch.bfloat16)
    dyb = _blocked(lib, dy.float().cuda(), torch.bfloat16)
    if kind == "k3":
        gw = lib.wgrad(dyb.view(cout), xb.view(cin), 3, 1, 1, dev)[:, :cout, :cin].permute(1, 2, 0).reshape(cout, cin, 3, 3, 3)
    elif kind == "down":
        gw = lib.wgrad(dyb.view(cout), xb.view(cin), 4, 2, 1, dev)[:, :cout, :cin].permute(1, 2, 0).reshape(cout, cin, 4, 4, 4)
    else:
        gw = lib.wgrad(xb.view(cin), dyb.view(cout), 4, 2, 1, dev)[:, :cin, :cout].permute(1, 2, 0).reshape(cin, cout, 4, 4, 4)
    scale = float(w.grad.abs().max())
    assert float((gw.cpu().double() - w.grad).abs().max()) <= 2e-5 * scale


def _msseg_like(filters, depth, residual=True, act=None):
    from segmentation_pipeline import models as M
    block_params = {"residual": residual}
    if act is not None:
        block_params.update(activation_class=torch.nn.LeakyReLU, activation_params={"negative_slope": act})
    return M.ModularUNet(2, 2, filters, depth, block_params=block_params,
                         downsample_class=M.BlurConv3d,
                         downsample_params={"kernel_size": 3, "stride": 2, "padding": 1},
                         upsample_class=M.BlurConvTranspose3d,
                         upsample_params={"kernel_size": 3, "stride": 2, "padding": 1, "output_padding": 0})


@pytest.mark.parametrize("filters,depth,residual,act", [([8, 16], 2, True, None), ([8, 8, 16], 3, False, 0.1)])
def test_training_step_matches_cpu_autograd(filters, depth, residual, act):
    """One trainer iteration (segmentation_trainer.py:162-180) on a small msseg2-style ModularUNet: probabilities,
    loss, EVERY parameter gradient, BatchNorm running statistics and the weights after the SGD step, vs autograd over
    the CPU oracle."""
    from segmentation_pipeline.criterions.hybrid_logistic_dice_loss import HybridLogisticDiceLoss
    torch.manual_seed(5)
    model = _msseg_like(filters, depth, residual, act)
    g = torch.Generator().manual_seed(6)
    for p in model.parameters():
        if p.dim() == 1:
            p.data.add_(0.1 * torch.randn(p.shape, generator=g))
    sd0 = {k: v.clone() for k, v in model.state_dict().items()}
    x = torch.randn(3, 2, 16, 16, 16, generator=g)
    labels = (torch.rand(3, 16, 16, 16, generator=g) < 0.3).long()
    target = torch.nn.functional.one_hot(labels, 2).movedim(-1, 1).float()

    # ---- CPU oracle: autograd through the functional restatement with batch-statistic BatchNorm
    ref_sd = {k: v.clone().requires_grad_(v.is_floating_point() and "running" not in k and "kernel" not in k)
              for k, v in sd0.items()}
    cfg = {"depth": depth, "filters": filters, "down": "blur", "up": "blur",
           "block": {"residual": residual, "bn_training": True, "act": "relu" if act is None else "leaky_relu",
                     "slope": act or 0.0}}
    ref_probs = unet.modular_unet_forward(ref_sd, x, cfg)
    ref_loss = unet.hybrid_logistic_dice_loss(ref_probs, target, logistic_class_weights=[1, 100])["loss"]
    ref_loss.backward()

    # ---- device
    model.cuda().train()
    opt = torch.optim.SGD(model.parameters(), lr=1e-3, momentum=0.95)
    criterion = HybridLogisticDiceLoss(logistic_class_weights=[1, 100])
    probs = model(x.cuda())
    assert probs.requires_grad
    loss = criterion(probs, target.cuda())["loss"]
    opt.zero_grad()
    loss.backward()
    assert float((probs.detach().cpu() - ref_probs.detach()).abs().max()) <= 1e-5
    assert abs(float(loss) - float(ref_loss)) <= 1e-5 * max(1.0, abs(float(ref_loss)))
    assert assert_gradients_match(model, ref_sd) >= 10
    for name, buf in model.named_buffers():
        if "running" in name:
            assert torch.allclose(buf.cpu(), ref_sd[name], rtol=1e-5, atol=1e-6), name
        if "num_batches_tracked" in name:
            assert int(buf) == 1
    opt.step()
    ref_params = [ref_sd[n] for n, _ in model.named_parameters()]
    ref_opt = torch.optim.SGD([p for p in ref_params if p.grad is not None], lr=1e-3, momentum=0.95)
    ref_opt.step()
    for (name, p), want in zip(model.named_parameters(), ref_params):
        assert torch.allclose(p.detach().cpu(), want.detach(), rtol=1e-5, atol=5e-5), name      # lr x the loose gradient bound


@pytest.mark.parametrize("filters,depth,residual,blur", [([8, 16], 2, True, True), ([16, 16, 24], 3, True, True),
                                                        ([8, 16, 24], 3, False, False)])
def test_mixed_precision_training_step_on_the_tensor_cores(filters, depth, residual, blur):
    """set_precision('bf16'): forward convolutions and data gradients on the tcgen05 engine, bf16 activations; fp32
    statistics / weight gradients / parameters.  Against fp32 autograd over the CPU oracle: probabilities within the
    bf16 tolerance of the inference path (2e-2), loss within 1 %.  Gradients: rounding the stored activations to bf16
    flips ReLU masks and moves a random-init network's gradients by 5-20 % of their norm -- the CPU oracle shows the
    same when only its FORWARD activations are rounded (``emulate_bf16``) -- so the bound per parameter is: no further
    from fp32 than 1.5 x that emulation + 3 % (the bf16 activation gradients of the backward pass)."""
    from segmentation_pipeline import models as M
    from segmentation_pipeline.criterions.hybrid_logistic_dice_loss import HybridLogisticDiceLoss
    from segmentation_pipeline.models import set_precision
    torch.manual_seed(5)
    model = _msseg_like(filters, depth, residual) if blur else M.ModularUNet(2, 2, filters, depth)
    g = torch.Generator().manual_seed(6)
    for p in model.parameters():
        if p.dim() == 1:
            p.data.add_(0.1 * torch.randn(p.shape, generator=g))
    sd0 = {k: v.clone() for k, v in model.state_dict().items()}
    x = torch.randn(3, 2, 32, 32, 32, generator=g)
    labels = (torch.rand(3, 32, 32, 32, generator=g) < 0.3).long()
    target = torch.nn.functional.one_hot(labels, 2).movedim(-1, 1).float()
    ref_sd = {k: v.clone().requires_grad_(v.is_floating_point() and "running" not in k and "kernel" not in k)
              for k, v in sd0.items()}
    cfg = {"depth": depth, "filters": filters, "down": "blur" if blur else "avgpool", "up": "blur" if blur else "trilinear",
           "block": {"residual": residual, "bn_training": True}}
    emu_sd = {k: (v if v.requires_grad else v.clone()) for k, v in ref_sd.items()}    # own running statistics
    with unet.emulate_bf16():
        emu_probs = unet.modular_unet_forward(emu_sd, x, cfg)
    unet.hybrid_logistic_dice_loss(emu_probs, target, logistic_class_weights=[1, 100])["loss"].backward()
    emulated = {k: v.grad.clone() for k, v in ref_sd.items() if v.grad is not None}
    for v in ref_sd.values():
        v.grad = None
    ref_probs = unet.modular_unet_forward(ref_sd, x, cfg)
    ref_loss = unet.hybrid_logistic_dice_loss(ref_probs, target, logistic_class_weights=[1, 100])["loss"]
    ref_loss.backward()
    model.cuda().train()
    set_precision("bf16")
    try:
        import b200seg
        launches = b200seg.launches()
        probs = model(x.cuda())
        loss = HybridLogisticDiceLoss(logistic_class_weights=[1, 100])(probs, target.cuda())["loss"]
        loss.backward()
    finally:
        set_precision("auto")
    assert b200seg.launches() > launches
    assert float((probs.detach().cpu() - ref_probs.detach()).abs().max()) <= 2e-2
    assert abs(float(loss.detach()) - float(ref_loss)) <= 1e-2 * abs(float(ref_loss))
    worst = ("", 0.0)
    for name, p in model.named_parameters():
        want = ref_sd[name].grad
        if want is None:
            assert p.grad is None, name
            continue
        err = float((p.grad.cpu() - want).norm()) / (float(want.norm()) + 1e-12)
        explained = float((emulated[name] - want).norm()) / (float(want.norm()) + 1e-12)
        assert err <= 1.5 * explained + 3e-2, (name, err, explained)
        if err > worst[1]:
            worst = (name, err, explained)
    print("worst relative gradient error (bf16 mixed precision; CPU forward emulation):", worst)
    for name, buf in model.named_buffers():
        if "running" in name:
            assert torch.allclose(buf.cpu(), ref_sd[name], rtol=2e-2, atol=2e-3), name


def test_weight_standardised_network_training_step():
    """WSConv3d blocks and weight-standardised blur convolutions: the standardisation (components.py:81-88, :113-116)
    is a tensor expression outside the device function, so autograd chains it -- checked against the oracle."""
    from segmentation_pipeline import models as M
    from segmentation_pipeline.criterions.hybrid_logistic_dice_loss import HybridLogisticDiceLoss
    torch.manual_seed(21)
    filters, depth = [8, 16], 2
    model = M.ModularUNet(2, 2, filters, depth,
                          block_params={"conv_class": M.WSConv3d, "conv_params": {"kernel_size": 3, "padding": 1},
                                        "residual": True, "residual_params": {"kernel_size": 3, "padding": 1}},
                          downsample_class=M.BlurConv3d,
                          downsample_params={"kernel_size": 3, "stride": 2, "padding": 1, "weight_standardization": True},
                          upsample_class=M.BlurConvTranspose3d,
                          upsample_params={"kernel_size": 3, "stride": 2, "padding": 1, "output_padding": 0,
                                           "weight_standardization": True},
                          hypothesis_class=torch.nn.Identity, hypothesis_params={})
    sd0 = {k: v.clone() for k, v in model.state_dict().items()}
    g = torch.Generator().manual_seed(22)
    x = torch.randn(2, 2, 16, 16, 16, generator=g)
    # unit-variance weights make the logits large (a softmax would saturate): compare the logits themselves, relatively,
    # under a fixed linear functional
    weights = torch.randn(2, 2, 16, 16, 16, generator=g)
    ref_sd = {k: v.clone().requires_grad_(v.is_floating_point() and "running" not in k and "kernel" not in k)
              for k, v in sd0.items()}
    cfg = {"depth": depth, "filters": filters, "down": "blur", "up": "blur", "down_ws": True, "up_ws": True,
           "hypothesis": "identity", "block": {"residual": True, "bn_training": True, "conv": "ws"}}
    ref_probs = unet.modular_unet_forward(ref_sd, x, cfg)
    (ref_probs * weights).sum().backward()
    model.cuda().train()
    probs = model(x.cuda())
    (probs * weights.cuda()).sum().backward()
    assert float((probs.detach().cpu() - ref_probs.detach()).abs().max()) <= 1e-5 * float(ref_probs.detach().abs().max())
    assert assert_gradients_match(model, ref_sd, tight=5e-4) >= 10


def test_default_modular_unet_training_step_matches_cpu_autograd():
    """The class defaults (AvgPool3d(2) down, trilinear Upsample(2, align_corners=True) up, no residual) -- the
    BASELINE config-1 network -- through one training step."""
    from segmentation_pipeline import models as M
    from segmentation_pipeline.criterions.hybrid_logistic_dice_loss import HybridLogisticDiceLoss
    torch.manual_seed(8)
    filters, depth = [8, 16, 24], 3
    model = M.ModularUNet(1, 3, filters, depth)
    sd0 = {k: v.clone() for k, v in model.state_dict().items()}
    g = torch.Generator().manual_seed(9)
    x = torch.randn(2, 1, 16, 24, 8, generator=g)
    target = torch.nn.functional.one_hot(torch.randint(0, 3, (2, 16, 24, 8), generator=g), 3).movedim(-1, 1).float()
    ref_sd = {k: v.clone().requires_grad_(v.is_floating_point() and "running" not in k) for k, v in sd0.items()}
    cfg = {"depth": depth, "filters": filters, "down": "avgpool", "up": "trilinear",
           "block": {"residual": False, "bn_training": True}}
    ref_probs = unet.modular_unet_forward(ref_sd, x, cfg)
    unet.hybrid_logistic_dice_loss(ref_probs, target)["loss"].backward()
    model.cuda().train()
    probs = model(x.cuda())
    HybridLogisticDiceLoss()(probs, target.cuda())["loss"].backward()
    assert float((probs.detach().cpu() - ref_probs.detach()).abs().max()) <= 1e-5
    assert_gradients_match(model, ref_sd)


def test_pool_and_upsample_adjoints_match_autograd():
    import b200seg as lib
    g = torch.Generator().manual_seed(3)
    x = torch.randn(2, 12, 4, 6, 10, generator=g, dtype=torch.float64, requires_grad=True)
    up = torch.nn.functional.interpolate(x, scale_factor=2, mode="trilinear", align_corners=True)
    dy = torch.randn(up.shape, generator=g, dtype=torch.float64)
    up.backward(dy)
    dy_buf = _blocked(lib, dy.float().cuda())
    dx = lib.Blocked(2, 2, 4, 6, 10, torch.float32, torch.device("cuda"))
    lib.upsample_trilinear2_backward(dy_buf.view(12), dx.view(12))
    assert float((_unblocked(lib, dx, 12).cpu().double() - x.grad).abs().max()) <= 1e-5 * float(x.grad.abs().max())
    big = torch.randn(2, 12, 8, 12, 20, generator=g, dtype=torch.float64, requires_grad=True)
    pooled = torch.nn.functional.avg_pool3d(big, 2, 2)
    dp = torch.randn(pooled.shape, generator=g, dtype=torch.float64)
    add = torch.randn(big.shape, generator=g, dtype=torch.float64)
    pooled.backward(dp)
    dp_buf, add_buf = _blocked(lib, dp.float().cuda()), _blocked(lib, add.float().cuda())
    dbig = lib.Blocked(2, 2, 8, 12, 20, torch.float32, torch.device("cuda"))
    lib.avgpool2_backward(dp_buf.view(12), dbig.view(12), add=add_buf.view(12))
    assert float((_unblocked(lib, dbig, 12).cpu().double() - (big.grad + add)).abs().max()) <= 1e-6


@pytest.mark.parametrize("dropout_p", [0.0, 0.3])
def test_nested_res_unet_training_step_matches_cpu_autograd(dropout_p, monkeypatch):
    """NestedResUNet (the dmri_hippo / qsm network; research/dmri_hippo/configs/*.py train it with dropout) through one
    training step: dense skip connections (tensors with several consumers), AvgPool / trilinear adjoints, Dropout3d.
    The Dropout3d draws are injected on both sides (the CPU oracle cannot share the device RNG stream)."""
    from segmentation_pipeline import models as M
    from segmentation_pipeline.criterions.hybrid_logistic_dice_loss import HybridLogisticDiceLoss
    from segmentation_pipeline.models import _train
    torch.manual_seed(12)
    model = M.NestedResUNet(3, 2, 8, dropout_p=dropout_p)
    sd0 = {k: v.clone() for k, v in model.state_dict().items()}
    g = torch.Generator().manual_seed(13)
    n = 2
    x = torch.randn(n, 3, 16, 24, 8, generator=g)
    target = torch.nn.functional.one_hot(torch.randint(0, 2, (n, 16, 24, 8), generator=g), 2).movedim(-1, 1).float()
    masks = {}
    if dropout_p:
        for name in _train.NESTED_BLOCKS:
            masks[name] = (torch.rand(n, 8, generator=g) >= dropout_p).float() / (1 - dropout_p)
        monkeypatch.setattr(_train, "MASK_OVERRIDE", lambda key, nn_, c, p: masks[key])
    ref_sd = {k: v.clone().requires_grad_(v.is_floating_point() and "running" not in k) for k, v in sd0.items()}
    ref_probs = unet.nested_res_unet_forward(ref_sd, x, {"bn_training": True, "dropout_masks": masks})
    ref_loss = unet.hybrid_logistic_dice_loss(ref_probs, target)["loss"]
    ref_loss.backward()
    model.cuda().train()
    probs = model(x.cuda())
    loss = HybridLogisticDiceLoss()(probs, target.cuda())["loss"]
    loss.backward()
    assert float((probs.detach().cpu() - ref_probs.detach()).abs().max()) <= 1e-5
    assert abs(float(loss.detach()) - float(ref_loss.detach())) <= 1e-5
    assert assert_gradients_match(model, ref_sd) == 70
    for name, buf in model.named_buffers():
        if "running" in name:
            assert torch.allclose(buf.cpu(), ref_sd[name], rtol=1e-5, atol=1e-6), name


def test_nested_res_unet_mixed_precision_and_real_dropout():
    """bf16 tensor-core path of the tape, and Dropout3d with the device RNG: channels are dropped per (sample, channel),
    survivors are scaled by 1 / (1 - p), gradients are finite."""
    from segmentation_pipeline import models as M
    from segmentation_pipeline.criterions.hybrid_logistic_dice_loss import HybridLogisticDiceLoss
    from segmentation_pipeline.models import set_precision
    torch.manual_seed(3)
    model = M.NestedResUNet(3, 2, 16, dropout_p=0.5).cuda().train()
    x = torch.randn(2, 3, 32, 32, 16).cuda()
    target = torch.nn.functional.one_hot(torch.randint(0, 2, (2, 32, 32, 16)), 2).movedim(-1, 1).float().cuda()
    for precision in ("fp32", "bf16"):
        set_precision(precision)
        try:
            model.zero_grad()
            loss = HybridLogisticDiceLoss()(model(x), target)["loss"]
            loss.backward()
        finally:
            set_precision("auto")
        assert torch.isfinite(loss.detach())
        assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in model.parameters())
        assert float(model.conv0_0.conv1.weight.grad.abs().max()) > 0


def test_modular_unet_dropout_training_step(monkeypatch):
    from segmentation_pipeline import models as M
    from segmentation_pipeline.models import _train
    torch.manual_seed(4)
    model = M.ModularUNet(1, 2, [8, 8], 2, block_params={"dropout_p": 0.25}).cuda().train()
    seen = []
    monkeypatch.setattr(_train, "MASK_OVERRIDE", lambda key, n, c, p: seen.append((key, n, c, p)))
    probs = model(torch.randn(2, 1, 16, 16, 16).cuda())
    probs.sum().backward()
    assert [k for k, *_ in seen] == [("down", 0), ("down", 1), ("up", 0)] and all(p == 0.25 for *_, p in seen)
    assert all(torch.isfinite(p.grad).all() for p in model.parameters() if p.grad is not None)


@pytest.mark.parametrize("momentum", [0.1, None])
def test_train_then_eval_uses_the_updated_weights_and_running_statistics(momentum):
    """segmentation_trainer.py:162-180 then :196-242: after ``optimizer.step()`` and ``model.eval()`` the inference plan
    must be rebuilt from the new weights AND the running statistics the training forward just updated (momentum, or the
    cumulative average of ``momentum=None``) -- compared with the oracle in eval mode on the oracle's own post-step state."""
    from segmentation_pipeline import models as M
    torch.manual_seed(31)
    filters, depth = [8, 16], 2
    model = M.ModularUNet(1, 2, filters, depth, block_params={"normalization_params": {"momentum": momentum}})
    sd0 = {k: v.clone() for k, v in model.state_dict().items()}
    g = torch.Generator().manual_seed(32)
    x = torch.randn(2, 1, 16, 16, 16, generator=g)
    weights = torch.randn(2, 2, 16, 16, 16, generator=g)
    model.cuda()
    with torch.no_grad():
        before = model.eval()(x.cuda()).cpu()                  # builds and caches the inference plan
    opt = torch.optim.SGD(model.parameters(), lr=0.05)
    model.train()
    (model(x.cuda()) * weights.cuda()).sum().backward()
    opt.step()
    model.eval()
    with torch.no_grad():
        after = model(x.cuda()).cpu()
    # oracle: the same step, then eval
    ref_sd = {k: v.clone().requires_grad_(v.is_floating_point() and "running" not in k) for k, v in sd0.items()}
    cfg = {"depth": depth, "filters": filters, "down": "avgpool", "up": "trilinear", "block": {"bn_training": True}}
    if momentum is None:        # F.batch_norm(momentum=None) is not the module's cumulative average: first batch -> factor 1
        for k in ref_sd:
            if "running_mean" in k:
                ref_sd[k].zero_()
            if "running_var" in k:
                ref_sd[k].zero_()
    import torch.nn.functional as F  # noqa: F401
    orig = unet.batch_norm_train
    try:
        if momentum is None:
            unet.batch_norm_train = lambda t, sd, prefix, eps=unet.BN_EPS, momentum=1.0: orig(t, sd, prefix, eps, 1.0)
        (unet.modular_unet_forward(ref_sd, x, cfg) * weights).sum().backward()
    finally:
        unet.batch_norm_train = orig
    with torch.no_grad():
        for v in ref_sd.values():
            if v.grad is not None:
                v -= 0.05 * v.grad
        want = unet.modular_unet_forward({k: v.detach() for k, v in ref_sd.items()}, x,
                                         {"depth": depth, "filters": filters, "down": "avgpool", "up": "trilinear", "block": {}})
    assert float((after - before).abs().max()) > 1e-3                       # the plan was rebuilt
    assert float((after - want).abs().max()) <= 2e-4            # (a ReLU-flip in the training step moves the new weights slightly)
    for name, buf in model.named_buffers():
        if "running" in name:
            assert torch.allclose(buf.cpu(), ref_sd[name], rtol=1e-5, atol=1e-6), name


def test_training_mode_unsupported_configuration_raises():
    from segmentation_pipeline import models as M
    model = M.ModularUNet(1, 2, [8, 8], 2, block_params={"normalization_class": torch.nn.InstanceNorm3d}).cuda().train()
    with pytest.raises(NotImplementedError):
        model(torch.randn(1, 1, 8, 8, 8).cuda())
