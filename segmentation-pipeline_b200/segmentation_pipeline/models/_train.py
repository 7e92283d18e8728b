"""Training-mode forward and backward of ``ModularUNet`` on the device -- the model part of the reference's trainer step
(``segmentation_trainer.py:162-180``: ``model.train()``, ``y_pred = model(X)``, ``loss.backward()``,
``optimizer.step()``), SURVEY.md section 8 rows a15 / f3.

Contract.  ``model(x)`` in training mode returns the probabilities as a tensor on the autograd tape; ``loss.backward()``
then fills ``.grad`` of every parameter the reference's forward uses, so the reference's optimizer step, gradient
clipping or a ``DistributedDataParallel`` wrapper work unchanged.  BatchNorm3d layers use BATCH statistics and update
``running_mean`` / ``running_var`` / ``num_batches_tracked`` exactly like ``nn.BatchNorm3d`` (momentum, unbiased running
variance).

How.  One ``torch.autograd.Function`` spans the whole network.  Its inputs are the *effective* weights: weight
standardisation (``components.py:81-88``) and the blur of ``BlurConv3d`` / ``BlurConvTranspose3d`` (``:111-121``,
``:144-154``) are tiny tensor expressions on the weights evaluated outside the function, so autograd chains their
derivatives by itself; everything that touches activations is a libb200seg kernel (fp32, blocked layout):

  forward   conv (``b200seg_conv3d_direct``: the register-tiled CUDA-core kernel, bias in the epilogue) -> per-channel
            batch moments (``b200seg_channel_moments``) -> BN apply + activation + residual add (``b200seg_affine_act``);
            the pre-BN output ``z`` of every convolution and the block inputs are kept for the backward
  backward  ``b200seg_softmax_backward`` -> per convolution: BN/activation backward (``b200seg_bn_backward``: d gamma,
            d beta, dz), weight gradient (``b200seg_wgrad``), data gradient = a convolution with re-arranged weights
            (flipped taps for 3x3x3; the strided conv's adjoint is the transposed geometry and vice versa), the
            two branches of a residual block / a skip connection summed through the convolution's residual epilogue.

Supported module space (anything else raises ``NotImplementedError`` -- never a silent ATen fallback):
``NestedResUNet`` (``nested_residual_unet.py:49-106``, run as a small tape because its skip tensors have several
consumers) and ``ModularUNet`` with ``Block3d`` blocks (``nn.Conv3d`` / ``WSConv3d`` 3x3x3 convolutions, ``BatchNorm3d`` or
no norm, ``ReLU`` / ``LeakyReLU`` / no activation, optional residual, ``Dropout3d``), ``AvgPool3d(2)`` or ``BlurConv3d`` down- and
trilinear ``Upsample(2)`` or ``BlurConvTranspose3d`` up-sampling (the class defaults, ``modular_unet.py:38-41``, and the
msseg2 configuration, ``research/msseg2/msseg2.py:84-93``), 3x3x3 ``out_conv``, ``Softmax(dim=1)`` or
``Identity`` hypothesis, channel counts that are multiples of 8 wherever tensors are concatenated.

Precision follows ``set_precision`` / autocast like the inference path: 'fp32' (default for fp32 modules) is the
CUDA-core path above; 'bf16' is mixed precision -- forward convolutions and data gradients run on the tcgen05 engine
(``b200seg_conv3d_tc``, all three geometries) with bf16 activations / activation gradients, while batch statistics,
weight gradients (fp32 accumulation over bf16 operands), parameters and the optimizer stay fp32."""
from __future__ import annotations

from typing import List, Optional

import torch
import torch.nn.functional as F
from torch import nn


def _lib():
    import b200seg
    b200seg.load_library()
    return b200seg


def _pad8(c: int) -> int:
    return (c + 7) // 8 * 8


class Unsupported(NotImplementedError):
    pass


def _triple(value):
    if isinstance(value, (tuple, list)):
        return tuple(int(v) if float(v) == int(v) else float(v) for v in value)
    return (int(value) if float(value) == int(value) else float(value),) * 3


# ------------------------------------------------------------------------------------------------- module -> spec
def _act_slope(module) -> float:
    if module is None:
        return 1.0
    if isinstance(module, nn.ReLU):
        return 0.0
    if isinstance(module, nn.LeakyReLU):
        return float(module.negative_slope)
    raise Unsupported(f"training: activation {type(module).__name__} is not lowered")


def _standardised(weight):
    w = weight - weight.mean(dim=(1, 2, 3, 4), keepdim=True)
    return w / (w.std(dim=(1, 2, 3, 4), keepdim=True) + 1e-5)


def _effective_conv_weight(conv, params: List[torch.Tensor]):
    """-> (index of the effective weight in params, index of the bias or None) for a 3x3x3 'same' convolution."""
    from .components import WSConv3d
    if not isinstance(conv, nn.Conv3d) or conv.kernel_size != (3, 3, 3) or conv.stride != (1, 1, 1) or \
            conv.padding != (1, 1, 1) or conv.dilation != (1, 1, 1) or conv.groups != 1 or conv.padding_mode != "zeros":
        raise Unsupported(f"training: {type(conv).__name__}{tuple(conv.kernel_size)} is not a plain 3x3x3 'same' convolution")
    if isinstance(conv, WSConv3d):
        params.append(_standardised(conv.weight))
        return len(params) - 1, None                         # the reference never applies a WSConv3d bias
    if type(conv) is not nn.Conv3d:
        raise Unsupported(f"training: convolution class {type(conv).__name__} is not lowered")
    params.append(conv.weight)
    wi = len(params) - 1
    bi = None
    if conv.bias is not None:
        params.append(conv.bias)
        bi = len(params) - 1
    return wi, bi


def _blurred(conv):
    weight = conv.weight
    if conv.weight_standardization:
        weight = _standardised(weight)
    return F.conv3d(weight, conv.kernel, padding=1, groups=conv.in_channels)


def _block_spec(block, params):
    from .components import Block3d
    if not isinstance(block, Block3d):
        raise Unsupported(f"training: block class {type(block).__name__} is not lowered")
    convs = []
    names = list(block.layers._modules.keys())
    index = 0
    while f"conv{index}" in block.layers._modules:
        conv = block.layers._modules[f"conv{index}"]
        norm = block.layers._modules.get(f"norm{index}")
        act = block.layers._modules.get(f"activation{index}")
        wi, bi = _effective_conv_weight(conv, params)
        entry = {"w": wi, "b": bi, "cin": conv.in_channels, "cout": conv.out_channels, "norm": None,
                 "slope": _act_slope(act)}
        if norm is not None:
            if not isinstance(norm, nn.BatchNorm3d) or not norm.affine:
                raise Unsupported(f"training: normalisation {type(norm).__name__} is not lowered (affine BatchNorm3d is)")
            params.append(norm.weight)
            params.append(norm.bias)
            entry["norm"] = {"g": len(params) - 2, "b": len(params) - 1, "module": norm}
        convs.append(entry)
        index += 1
    if index == 0 or len([n for n in names if n.startswith("conv")]) != index:
        raise Unsupported("training: unexpected Block3d layout")
    res = None
    if block.residual:
        wi, bi = _effective_conv_weight(block.res_conv, params)
        res = {"w": wi, "b": bi}
    return {"convs": convs, "res": res, "cin": convs[0]["cin"], "cout": convs[-1]["cout"],
            "dropout_p": 0.0 if block.dropout is None else float(block.dropout.p)}


def _nested_block_spec(block, params):
    """NestedResUNet.Block (nested_residual_unet.py:7-47): conv1-bn1-relu, conv2-bn2-relu, optional res_conv, dropout."""
    convs = []
    for conv, norm in ((block.conv1, block.bn1), (block.conv2, block.bn2)):
        wi, bi = _effective_conv_weight(conv, params)
        params.append(norm.weight)
        params.append(norm.bias)
        convs.append({"w": wi, "b": bi, "cin": conv.in_channels, "cout": conv.out_channels, "slope": 0.0,
                      "norm": {"g": len(params) - 2, "b": len(params) - 1, "module": norm}})
    res = None
    if block.residual:
        wi, bi = _effective_conv_weight(block.res_conv, params)
        res = {"w": wi, "b": bi}
    return {"convs": convs, "res": res, "cin": convs[0]["cin"], "cout": convs[-1]["cout"],
            "dropout_p": 0.0 if block.dropout is None else float(block.dropout.p)}


NESTED_BLOCKS = ("conv0_0", "conv1_0", "conv0_1", "conv2_0", "conv1_1", "conv0_2", "conv3_0", "conv2_1", "conv1_2", "conv0_3")


def build_nested_spec(model):
    """-> (spec, params) of a NestedResUNet (nested_residual_unet.py:49-106)."""
    params: List[torch.Tensor] = []
    spec = {"kind": "nested", "blocks": {}}
    for name in NESTED_BLOCKS:
        spec["blocks"][name] = _nested_block_spec(getattr(model, name), params)
    filters = spec["blocks"]["conv0_0"]["cout"]
    if filters % 8:
        raise Unsupported("training: NestedResUNet filters must be a multiple of 8 (channel concatenation)")
    wi, bi = _effective_conv_weight(model.out_conv, params)
    spec["out"] = {"w": wi, "b": bi, "cin": model.out_conv.in_channels, "cout": model.out_conv.out_channels}
    hyp = model.hypothesis
    if isinstance(hyp, nn.Softmax) and hyp.dim == 1:
        spec["softmax"] = True
    elif isinstance(hyp, nn.Identity):
        spec["softmax"] = False
    else:
        raise Unsupported(f"training: hypothesis {type(hyp).__name__} is not lowered (Softmax(dim=1) / Identity are)")
    if spec["out"]["cout"] > 16:
        raise Unsupported("training: more than 16 output channels are not lowered")
    return spec, params


def build_spec(model):
    """-> (spec, params): the layer list of a ModularUNet and the flat list of effective parameter tensors."""
    from .components import BlurConv3d, BlurConvTranspose3d
    from .modular_unet import ModularUNet
    from .nested_residual_unet import NestedResUNet
    if isinstance(model, NestedResUNet):
        return build_nested_spec(model)
    if not isinstance(model, ModularUNet):
        raise Unsupported(f"training: {type(model).__name__} is not lowered (ModularUNet and NestedResUNet are)")
    params: List[torch.Tensor] = []
    spec = {"kind": "modular", "depth": model.depth, "down": [], "downs": [], "up": [], "ups": []}
    for block in model.down_blocks:
        spec["down"].append(_block_spec(block, params))
    for i, down in enumerate(model.downsampling):
        width = spec["down"][i]["cout"]
        if isinstance(down, nn.AvgPool3d):
            if _triple(down.kernel_size) != (2, 2, 2) or _triple(down.stride) != (2, 2, 2) or _triple(down.padding) != (0, 0, 0) \
                    or down.ceil_mode:
                raise Unsupported("training: only AvgPool3d(kernel_size=2, stride=2) is lowered")
            spec["downs"].append({"w": None, "cin": width, "cout": width})
            continue
        if not isinstance(down, BlurConv3d) or down.kernel_size != (3, 3, 3) or down.stride != (2, 2, 2) or \
                down.padding != (1, 1, 1):
            raise Unsupported(f"training: downsampling {type(down).__name__} is not lowered "
                              f"(AvgPool3d(2) and BlurConv3d(kernel_size=3, stride=2, padding=1) are)")
        params.append(_blurred(down))
        spec["downs"].append({"w": len(params) - 1, "cin": down.in_channels, "cout": down.out_channels})
    for block in model.up_blocks:
        spec["up"].append(_block_spec(block, params))
    for i, up in enumerate(model.upsampling):
        if isinstance(up, nn.Upsample):
            if up.mode != "trilinear" or not up.align_corners or up.size is not None or \
                    _triple(up.scale_factor) != (2, 2, 2):
                raise Unsupported("training: only Upsample(scale_factor=2, mode='trilinear', align_corners=True) is lowered")
            width = spec["down"][i + 1]["cout"]
            spec["ups"].append({"w": None, "cin": width, "cout": width})
            continue
        if not isinstance(up, BlurConvTranspose3d) or up.kernel_size != (3, 3, 3) or up.stride != (2, 2, 2) or \
                up.padding != (1, 1, 1) or up.output_padding != (0, 0, 0):
            raise Unsupported(f"training: upsampling {type(up).__name__} is not lowered (trilinear Upsample(2) and "
                              f"BlurConvTranspose3d(kernel_size=3, stride=2, padding=1, output_padding=0) are)")
        params.append(_blurred(up))
        spec["ups"].append({"w": len(params) - 1, "cin": up.in_channels, "cout": up.out_channels})
    wi, bi = _effective_conv_weight(model.out_conv, params)
    spec["out"] = {"w": wi, "b": bi, "cin": model.out_conv.in_channels, "cout": model.out_conv.out_channels}
    hyp = model.hypothesis
    if isinstance(hyp, nn.Softmax) and hyp.dim == 1:
        spec["softmax"] = True
    elif isinstance(hyp, nn.Identity):
        spec["softmax"] = False
    else:
        raise Unsupported(f"training: hypothesis {type(hyp).__name__} is not lowered (Softmax(dim=1) / Identity are)")
    if spec["out"]["cout"] > 16:
        raise Unsupported("training: more than 16 output channels are not lowered")
    for i in range(model.depth - 1):
        if spec["ups"][i]["cout"] % 8 or spec["down"][i]["cout"] % 8:
            raise Unsupported("training: concatenated channel counts must be multiples of 8")
        if spec["up"][i]["cin"] != spec["ups"][i]["cout"] + spec["down"][i]["cout"]:
            raise Unsupported("training: up-block input width does not match upsample + skip")
    return spec, params


# ------------------------------------------------------------------------------------------------- weight layouts
def _pack(t: torch.Tensor) -> torch.Tensor:
    """(k, k, k, in, out) -> fp32 [k^3][round_up(in, 8)][round_up(out, 8)] for b200seg_conv3d_direct."""
    k, cin, cout = t.shape[0], t.shape[3], t.shape[4]
    if cin % 8 == 0 and cout % 8 == 0:
        return t.reshape(k ** 3, cin, cout).to(torch.float32).contiguous()
    out = torch.zeros((k ** 3, _pad8(cin), _pad8(cout)), dtype=torch.float32, device=t.device)
    out[:, :cin, :cout] = t.reshape(k ** 3, cin, cout)
    return out


_CONST_VEC = {}


def _vec(values: Optional[torch.Tensor], channels: int, fill: float, device) -> torch.Tensor:
    """fp32 vector of length round_up(channels, 8): ``values`` padded with ``fill`` (constant vectors are cached: a
    training step asks for a few hundred of them)."""
    if values is None:
        key = (_pad8(channels), float(fill), str(device))
        hit = _CONST_VEC.get(key)
        if hit is None:
            hit = _CONST_VEC[key] = torch.full((_pad8(channels),), fill, dtype=torch.float32, device=device)
        return hit
    if channels == _pad8(channels):
        return values.to(torch.float32).contiguous()
    out = torch.full((_pad8(channels),), fill, dtype=torch.float32, device=device)
    out[:channels] = values
    return out


MASK_OVERRIDE = None     # tests: callable (block key, n, channels, p) -> (n, channels) multiplier or None


class _Val:
    """A tensor of the training tape: a chunk range of a blocked buffer plus its gradient (another _Val) once known."""
    __slots__ = ("buf", "channels", "off", "grad")

    def __init__(self, buf, channels, off=0):
        self.buf, self.channels, self.off, self.grad = buf, channels, off, None

    def view(self):
        return self.buf.view(self.channels, self.off)

    @property
    def ext(self):
        return (self.buf.z, self.buf.y, self.buf.x)


_TC_INDEX = {}      # (mode, cin_chunks, cout, device) -> (clamped gather index, fp32 mask) on the device


def _tc_index(mode, cin_chunks, cout, device):
    from . import _plan
    key = (mode, cin_chunks, cout, str(device))
    hit = _TC_INDEX.get(key)
    if hit is None:
        idx = _plan.tc_gather_index(mode, cin_chunks, cout)
        hit = (idx.clamp(min=0).to(device), (idx >= 0).to(torch.float32).to(device))
        _TC_INDEX[key] = hit
    return hit


def _sub_view(lib, v, channels, chunk_off):
    return lib.View(v.data, v.dtype, v.n, channels, v.c8_total, v.c8_off + chunk_off, v.z, v.y, v.x)


class _Runner:
    """State of one training step: buffers kept by the forward for the backward.  ``precision``: 'fp32' (CUDA-core
    convolutions, fp32 activations) or 'bf16' (mixed precision: tensor-core convolutions for forward and dgrad, bf16
    activations and activation gradients; statistics, weight gradients, parameters and the optimizer stay fp32)."""

    def __init__(self, spec, device, precision="fp32"):
        self.spec = spec
        self.device = device
        self.precision = precision
        self.dtype = torch.bfloat16 if precision == "bf16" else torch.float32
        self.lib = _lib()
        self.saved = {}

    # -------------------------------------------------------------------------- primitives
    def buffer(self, n, channels, ext):
        return self.lib.Blocked(n, _pad8(channels) // 8, ext[0], ext[1], ext[2], self.dtype, self.device)

    def conv(self, src, weight_packed, cout, dst, bias=None, residual=None, ksize=3, stride=1, pad=1, transposed=False,
             out_ncdhw=None, softmax=False):
        """``weight_packed``: fp32 [k^3][round_up(cin, 8)][round_up(cout, 8)] (``_pack``)."""
        lib = self.lib
        if self.precision == "bf16":
            return self._conv_tc(src, weight_packed, cout, dst, bias, residual, ksize, stride, transposed, out_ncdhw,
                                 softmax)
        one = _vec(None, cout, 1.0, self.device)
        shift = _vec(bias, cout, 0.0, self.device)
        if out_ncdhw is not None:
            epi = lib.make_epilogue(one, shift, one, out_ncdhw=out_ncdhw, softmax=softmax, slope01=1)
        else:
            epi = lib.make_epilogue(one, shift, one, dst, residual=lib.NULL_VIEW if residual is None else residual,
                                    slope01=1)
        lib.conv3d_direct(src, weight_packed, cout, ksize, stride, pad, transposed, epi)

    def _conv_tc(self, src, phys, cout, dst, bias, residual, ksize, stride, transposed, out_ncdhw, softmax):
        """The same convolution on the tensor-core engine: the fp32 ``phys`` weight is gathered into the engine's bf16
        operand image on the device (the weights change every step, so the host-side packer of the inference plan is
        replaced by one ``take`` through a cached index map), wide outputs are cut into launches of <= 80 channels."""
        from . import _plan
        lib = self.lib
        mode = _plan.K3 if ksize == 3 else (_plan.UP if transposed else _plan.DOWN)
        cin_chunks = phys.shape[1] // 8
        out_ext = (src.z, src.y, src.x) if mode == _plan.K3 else \
            ((src.z * 2, src.y * 2, src.x * 2) if mode == _plan.UP else (src.z // 2, src.y // 2, src.x // 2))
        limit = 120 if max(out_ext) <= 12 else 80
        if out_ncdhw is not None:
            limit = 16
            if cout <= _plan.K3T_MAX_COUT and cin_chunks <= 12:
                mode = _plan.K3T
        for lo in range(0, cout, limit):
            hi = min(lo + limit, cout)
            width = _pad8(hi - lo)
            piece = phys[:, :, lo:lo + width].contiguous().reshape(-1)
            idx, mask = _tc_index(mode, cin_chunks, hi - lo, self.device)
            image = (piece[idx] * mask).to(torch.bfloat16).contiguous()
            one = _vec(None, hi - lo, 1.0, self.device)
            shift = _vec(None if bias is None else bias[lo:hi], hi - lo, 0.0, self.device)
            if out_ncdhw is not None:
                epi = lib.make_epilogue(one, shift, one, out_ncdhw=out_ncdhw, softmax=softmax, slope01=1)
            else:
                epi = lib.make_epilogue(one, shift, one, _sub_view(lib, dst, hi - lo, lo // 8),
                                        residual=lib.NULL_VIEW if residual is None
                                        else _sub_view(lib, residual, hi - lo, lo // 8),
                                        slope01=2 if bias is None else 1)
            lib.conv3d_tc(mode, src, image, hi - lo, epi)

    # -------------------------------------------------------------------------- forward
    def block_forward(self, key, bspec, params, src_view, n, ext, dst_buf, dst_off):
        """src_view -> block output written into chunk range [dst_off, ...) of dst_buf; returns its view."""
        lib = self.lib
        cout = bspec["cout"]
        state = {"src": src_view, "convs": []}
        res_buf = None
        if bspec["res"] is not None:
            res_buf = self.buffer(n, cout, ext)
            w = params[bspec["res"]["w"]]
            bias = None if bspec["res"]["b"] is None else params[bspec["res"]["b"]]
            self.conv(src_view, _pack(w.detach().permute(2, 3, 4, 1, 0)), cout, res_buf.view(cout), bias=bias)
        cur = src_view
        last = len(bspec["convs"]) - 1
        for j, cs in enumerate(bspec["convs"]):
            z = self.buffer(n, cs["cout"], ext)
            w = params[cs["w"]]
            bias = None if cs["b"] is None else params[cs["b"]]
            self.conv(cur, _pack(w.detach().permute(2, 3, 4, 1, 0)), cs["cout"], z.view(cs["cout"]), bias=bias)
            count = n * ext[0] * ext[1] * ext[2]
            if cs["norm"] is not None:
                bn = cs["norm"]["module"]
                mean, var = lib.channel_moments(z.view(cs["cout"]), cs["cout"], self.device)
                rstd = torch.rsqrt(var + bn.eps)
                gamma = _vec(params[cs["norm"]["g"]].detach(), cs["cout"], 0.0, self.device)
                beta = _vec(params[cs["norm"]["b"]].detach(), cs["cout"], 0.0, self.device)
                scale = gamma * rstd
                shift = beta - mean * scale
                self._update_running(bn, mean[:cs["cout"]], var[:cs["cout"]], count)
            else:
                mean = torch.zeros(_pad8(cs["cout"]), dtype=torch.float32, device=self.device)
                rstd = torch.ones_like(mean)
                scale = torch.ones_like(mean)
                shift = torch.zeros_like(mean)
            slope = _vec(None, cs["cout"], cs["slope"], self.device)
            if j == last:
                a_view = dst_buf.view(cout, dst_off)
                lib.affine_act(z.view(cs["cout"]), scale, shift, slope, a_view,
                               residual=lib.NULL_VIEW if res_buf is None else res_buf.view(cout))
            else:
                a_buf = self.buffer(n, cs["cout"], ext)
                a_view = a_buf.view(cs["cout"])
                lib.affine_act(z.view(cs["cout"]), scale, shift, slope, a_view)
                state.setdefault("keep", []).append(a_buf)
            state["convs"].append({"in": cur, "z": z, "scale": scale, "shift": shift, "slope": slope, "mean": mean,
                                   "rstd": rstd})
            cur = a_view
        if bspec.get("dropout_p", 0.0) > 0.0:
            # nn.Dropout3d after the residual add (components.py:70-71): one Bernoulli draw per (sample, channel)
            mask = self.dropout_mask(key, n, cout, bspec["dropout_p"])
            lib.channel_scale(cur, mask, cur)
            state["mask"] = mask
        state["ext"], state["n"] = ext, n
        self.saved[key] = state
        return cur

    def dropout_mask(self, key, n, channels, p):
        if MASK_OVERRIDE is not None:
            forced = MASK_OVERRIDE(key, n, channels, p)
            if forced is not None:
                out = torch.zeros((n, _pad8(channels)), dtype=torch.float32, device=self.device)
                out[:, :channels] = forced.to(self.device, torch.float32)
                return out
        keep = (torch.rand((n, _pad8(channels)), device=self.device) >= p).to(torch.float32)
        return (keep / (1.0 - p)).contiguous()

    @staticmethod
    def _update_running(bn, mean, var, count):
        if not bn.track_running_stats or bn.running_mean is None:
            return
        with torch.no_grad():
            bn.num_batches_tracked += 1
            momentum = bn.momentum if bn.momentum is not None else 1.0 / float(bn.num_batches_tracked)
            unbiased = var * (count / max(count - 1, 1))
            bn.running_mean.mul_(1 - momentum).add_(mean.to(bn.running_mean.dtype), alpha=momentum)
            bn.running_var.mul_(1 - momentum).add_(unbiased.to(bn.running_var.dtype), alpha=momentum)

    def forward(self, x, params):
        if self.spec.get("kind") == "nested":
            return self.nested_forward(x, params)
        lib, spec = self.lib, self.spec
        n, cin = x.shape[0], x.shape[1]
        ext = tuple(x.shape[2:])
        depth = spec["depth"]
        if any(e % (1 << (depth - 1)) for e in ext):
            raise RuntimeError(f"training: spatial extent {ext} must be divisible by 2^{depth - 1}")
        if cin != spec["down"][0]["cin"]:
            raise RuntimeError(f"expected {spec['down'][0]['cin']} input channels, got {cin}")
        xin = self.buffer(n, cin, ext)
        lib.pack_ncdhw(x.detach().to(torch.float32).contiguous(), xin.view(cin))
        exts = [tuple(e >> lvl for e in ext) for lvl in range(depth)]
        cur = xin.view(cin)
        self.saved["x"] = xin
        cats = []
        for i in range(depth):
            b = spec["down"][i]
            if i != depth - 1:
                up_c = spec["ups"][i]["cout"]
                cat = self.buffer(n, up_c + b["cout"], exts[i])        # [upsampled | skip], modular_unet.py:97
                cats.append(cat)
                out = self.block_forward(("down", i), b, params, cur, n, exts[i], cat, up_c // 8)
                d = spec["downs"][i]
                nxt = self.buffer(n, d["cout"], exts[i + 1])
                if d["w"] is None:
                    lib.avgpool2(out, nxt.view(d["cout"]))                          # nn.AvgPool3d(2), modular_unet.py:40-41
                else:
                    wb = params[d["w"]].detach()
                    self.conv(out, _pack(wb.permute(2, 3, 4, 1, 0)), d["cout"], nxt.view(d["cout"]), ksize=4, stride=2,
                              pad=1)
                self.saved[("downs", i)] = {"in": out, "out_buf": nxt}
                cur = nxt.view(d["cout"])
            else:
                bottom = self.buffer(n, b["cout"], exts[i])
                cur = self.block_forward(("down", i), b, params, cur, n, exts[i], bottom, 0)
                self.saved["bottom"] = bottom
        for i in reversed(range(depth - 1)):
            u = spec["ups"][i]
            cat = cats[i]
            if u["w"] is None:
                lib.upsample_trilinear2(cur, cat.view(u["cout"], 0))                # nn.Upsample, modular_unet.py:38-39
            else:
                wt = params[u["w"]].detach()
                self.conv(cur, _pack(wt.permute(2, 3, 4, 0, 1)), u["cout"], cat.view(u["cout"], 0), ksize=4, stride=2,
                          pad=1, transposed=True)
            self.saved[("ups", i)] = {"in": cur}
            b = spec["up"][i]
            outb = self.buffer(n, b["cout"], exts[i])
            cur = self.block_forward(("up", i), b, params, cat.view(b["cin"]), n, exts[i], outb, 0)
            self.saved[("upout", i)] = outb
        self.saved["cats"] = cats
        o = spec["out"]
        probs = torch.empty((n, o["cout"], *ext), dtype=torch.float32, device=self.device)
        w = params[o["w"]].detach()
        bias = None if o["b"] is None else params[o["b"]].detach()
        self.conv(cur, _pack(w.permute(2, 3, 4, 1, 0)), o["cout"], None, bias=bias, out_ncdhw=probs,
                  softmax=spec["softmax"])
        self.saved["out_in"] = cur
        self.saved["probs"] = probs
        self.saved["n"], self.saved["ext"], self.saved["exts"] = n, ext, exts
        return probs

    # -------------------------------------------------------------------------- backward
    def _wgrad_conv(self, dz_view, x_view, cout, cin, ksize=3, stride=1, pad=1):
        """Gradient of a (cout, cin, k, k, k) convolution weight."""
        g = self.lib.wgrad(dz_view, x_view, ksize, stride, pad, self.device)          # [tap][cout_pad][cin_pad]
        return g[:, :cout, :cin].permute(1, 2, 0).reshape(cout, cin, ksize, ksize, ksize)

    def _chan_sum(self, view, channels, count):
        mean, _ = self.lib.channel_moments(view, channels, self.device)
        return mean[:channels] * float(count)

    def block_backward(self, key, bspec, params, dout_view, grads, need_dsrc=True):
        """dout_view: gradient of the block output.  Returns the buffer holding the gradient of the block input (or
        None)."""
        lib = self.lib
        st = self.saved.pop(key)
        n, ext = st["n"], st["ext"]
        count = n * ext[0] * ext[1] * ext[2]
        cin, cout = bspec["cin"], bspec["cout"]
        if "mask" in st:
            masked = self.buffer(n, cout, ext)
            lib.channel_scale(dout_view, st["mask"], masked.view(cout))
            dout_view = masked.view(cout)
        dcur = dout_view
        dsrc = self.buffer(n, cin, ext) if need_dsrc else None
        for j in reversed(range(len(bspec["convs"]))):
            cs, sv = bspec["convs"][j], st["convs"][j]
            dz = self.buffer(n, cs["cout"], ext)
            sum_g, sum_gx = lib.bn_backward(dcur, sv["z"].view(cs["cout"]), sv["scale"], sv["shift"], sv["slope"],
                                            sv["mean"], sv["rstd"], cs["norm"] is not None, dz.view(cs["cout"]),
                                            cs["cout"], self.device)
            if cs["norm"] is not None:
                grads[cs["norm"]["g"]] = sum_gx[:cs["cout"]]
                grads[cs["norm"]["b"]] = sum_g[:cs["cout"]]
                if cs["b"] is not None:
                    grads[cs["b"]] = self._chan_sum(dz.view(cs["cout"]), cs["cout"], count)
            elif cs["b"] is not None:
                grads[cs["b"]] = sum_g[:cs["cout"]]
            grads[cs["w"]] = self._wgrad_conv(dz.view(cs["cout"]), sv["in"], cs["cout"], cs["cin"])
            w = params[cs["w"]].detach()
            if j > 0:
                dprev = self.buffer(n, cs["cin"], ext)
                self.conv(dz.view(cs["cout"]), _pack(w.flip(2, 3, 4).permute(2, 3, 4, 0, 1)), cs["cin"],
                          dprev.view(cs["cin"]))
                dcur = dprev.view(cs["cin"])
            elif need_dsrc:
                self.conv(dz.view(cs["cout"]), _pack(w.flip(2, 3, 4).permute(2, 3, 4, 0, 1)), cin, dsrc.view(cin))
            sv["z"] = None
        if bspec["res"] is not None:
            r = bspec["res"]
            grads[r["w"]] = self._wgrad_conv(dout_view, st["src"], cout, cin)
            if r["b"] is not None:
                grads[r["b"]] = self._chan_sum(dout_view, cout, count)
            if need_dsrc:
                w = params[r["w"]].detach()
                total = self.buffer(n, cin, ext)
                self.conv(dout_view, _pack(w.flip(2, 3, 4).permute(2, 3, 4, 0, 1)), cin, total.view(cin),
                          residual=dsrc.view(cin))
                dsrc = total
        return dsrc

    # -------------------------------------------------------------------------- NestedResUNet: a small tape
    def accumulate(self, val, g):
        """val.grad += g (g: _Val).  The first contribution is aliased, later ones are added into a fresh buffer."""
        if val.grad is None:
            val.grad = g
            return
        n = val.buf.n
        total = self.buffer(n, val.channels, val.ext)
        one, zero = _vec(None, val.channels, 1.0, self.device), _vec(None, val.channels, 0.0, self.device)
        self.lib.affine_act(g.view(), one, zero, one, total.view(val.channels), residual=val.grad.view())
        val.grad = _Val(total, val.channels)

    def nested_forward(self, x, params):
        """NestedResUNet.forward (nested_residual_unet.py:88-106) op by op; every op appends its backward to the tape."""
        lib, spec = self.lib, self.spec
        n, cin = x.shape[0], x.shape[1]
        ext = tuple(x.shape[2:])
        if any(e % 8 for e in ext):
            raise RuntimeError(f"training: spatial extent {ext} must be divisible by 8")
        if cin != spec["blocks"]["conv0_0"]["cin"]:
            raise RuntimeError(f"expected {spec['blocks']['conv0_0']['cin']} input channels, got {cin}")
        exts = [tuple(e >> lvl for e in ext) for lvl in range(4)]
        tape = []
        xin = self.buffer(n, cin, ext)
        lib.pack_ncdhw(x.detach().to(torch.float32).contiguous(), xin.view(cin))
        v_in = _Val(xin, cin)

        def block(name, src, level, is_input=False):
            b = spec["blocks"][name]
            if src.channels != b["cin"]:
                raise RuntimeError(f"{name}: expected {b['cin']} channels, got {src.channels}")
            out = self.buffer(n, b["cout"], exts[level])
            self.block_forward(name, b, params, src.view(), n, exts[level], out, 0)
            o = _Val(out, b["cout"])

            def bwd(grads):
                if o.grad is None:
                    return
                dsrc = self.block_backward(name, b, params, o.grad.view(), grads, need_dsrc=not is_input)
                if dsrc is not None:
                    self.accumulate(src, _Val(dsrc, b["cin"]))
            tape.append(bwd)
            return o

        def down(src, level):          # nn.AvgPool3d(2), level -> level + 1
            c = src.channels
            out = self.buffer(n, c, exts[level + 1])
            lib.avgpool2(src.view(), out.view(c))
            o = _Val(out, c)

            def bwd(grads):
                dx = self.buffer(n, c, exts[level])
                lib.avgpool2_backward(o.grad.view(), dx.view(c), add=lib.NULL_VIEW if src.grad is None else src.grad.view())
                src.grad = _Val(dx, c)
            tape.append(bwd)
            return o

        def up(src, level):            # trilinear Upsample(2), level -> level - 1
            c = src.channels
            out = self.buffer(n, c, exts[level - 1])
            lib.upsample_trilinear2(src.view(), out.view(c))
            o = _Val(out, c)

            def bwd(grads):
                dx = self.buffer(n, c, exts[level])
                lib.upsample_trilinear2_backward(o.grad.view(), dx.view(c))
                self.accumulate(src, _Val(dx, c))
            tape.append(bwd)
            return o

        def cat(parts, level):         # torch.cat(..., 1): copies into one buffer (skip tensors have several consumers)
            total = sum(p.channels for p in parts)
            buf = self.buffer(n, total, exts[level])
            off = 0
            for p in parts:
                lib.copy_view(p.view(), buf.view(p.channels, off // 8))
                off += p.channels
            o = _Val(buf, total)

            def bwd(grads):
                off = 0
                for p in parts:
                    self.accumulate(p, _Val(o.grad.buf, p.channels, o.grad.off + off // 8))
                    off += p.channels
            tape.append(bwd)
            return o

        x0_0 = block("conv0_0", v_in, 0, is_input=True)
        x1_0 = block("conv1_0", down(x0_0, 0), 1)
        x0_1 = block("conv0_1", cat([x0_0, up(x1_0, 1)], 0), 0)
        x2_0 = block("conv2_0", down(x1_0, 1), 2)
        x1_1 = block("conv1_1", cat([x1_0, up(x2_0, 2), down(x0_1, 0)], 1), 1)
        x0_2 = block("conv0_2", cat([x0_1, up(x1_1, 1)], 0), 0)
        x3_0 = block("conv3_0", down(x2_0, 2), 3)
        x2_1 = block("conv2_1", cat([x2_0, up(x3_0, 3), down(x1_1, 1)], 2), 2)
        x1_2 = block("conv1_2", cat([x1_1, up(x2_1, 2), down(x0_2, 0)], 1), 1)
        x0_3 = block("conv0_3", cat([x0_2, up(x1_2, 1)], 0), 0)
        o = spec["out"]
        probs = torch.empty((n, o["cout"], *ext), dtype=torch.float32, device=self.device)
        w = params[o["w"]].detach()
        bias = None if o["b"] is None else params[o["b"]].detach()
        self.conv(x0_3.view(), _pack(w.permute(2, 3, 4, 1, 0)), o["cout"], None, bias=bias, out_ncdhw=probs,
                  softmax=spec["softmax"])
        self.saved["tape"], self.saved["head_in"], self.saved["probs"] = tape, x0_3, probs
        self.saved["n"], self.saved["ext"] = n, ext
        return probs

    def nested_backward(self, dprobs, params):
        lib, spec = self.lib, self.spec
        n, ext = self.saved["n"], self.saved["ext"]
        grads = [None] * len(params)
        o = spec["out"]
        head_in = self.saved.pop("head_in")
        dlogits = self.buffer(n, o["cout"], ext)
        lib.softmax_backward(self.saved.pop("probs"), dprobs.detach().to(torch.float32).contiguous(), spec["softmax"],
                             dlogits.view(o["cout"]))
        grads[o["w"]] = self._wgrad_conv(dlogits.view(o["cout"]), head_in.view(), o["cout"], o["cin"])
        if o["b"] is not None:
            grads[o["b"]] = self._chan_sum(dlogits.view(o["cout"]), o["cout"], n * ext[0] * ext[1] * ext[2])
        w = params[o["w"]].detach()
        dhead = self.buffer(n, o["cin"], ext)
        self.conv(dlogits.view(o["cout"]), _pack(w.flip(2, 3, 4).permute(2, 3, 4, 0, 1)), o["cin"], dhead.view(o["cin"]))
        self.accumulate(head_in, _Val(dhead, o["cin"]))
        for bwd in reversed(self.saved.pop("tape")):
            bwd(grads)
        self.saved.clear()
        return grads

    def backward(self, dprobs, params):
        if not self.saved:
            raise RuntimeError("the activations of this training step were released by its first backward pass "
                               "(backward through the b200 network twice / retain_graph is not supported)")
        if self.spec.get("kind") == "nested":
            return self.nested_backward(dprobs, params)
        lib, spec = self.lib, self.spec
        n, ext, exts = self.saved["n"], self.saved["ext"], self.saved["exts"]
        depth = spec["depth"]
        grads = [None] * len(params)
        o = spec["out"]
        count0 = n * ext[0] * ext[1] * ext[2]
        dlogits = self.buffer(n, o["cout"], ext)
        lib.softmax_backward(self.saved.pop("probs"), dprobs.detach().to(torch.float32).contiguous(), spec["softmax"],
                             dlogits.view(o["cout"]))
        out_in = self.saved.pop("out_in")
        grads[o["w"]] = self._wgrad_conv(dlogits.view(o["cout"]), out_in, o["cout"], o["cin"])
        if o["b"] is not None:
            grads[o["b"]] = self._chan_sum(dlogits.view(o["cout"]), o["cout"], count0)
        w = params[o["w"]].detach()
        dcur_buf = self.buffer(n, o["cin"], ext)
        self.conv(dlogits.view(o["cout"]), _pack(w.flip(2, 3, 4).permute(2, 3, 4, 0, 1)), o["cin"], dcur_buf.view(o["cin"]))
        del dlogits
        dcur = dcur_buf.view(o["cin"])
        dskips = {}
        for i in range(depth - 1):                         # up blocks, level 0 first (reverse of the forward order)
            b, u = spec["up"][i], spec["ups"][i]
            dcat = self.block_backward(("up", i), b, params, dcur, grads)      # [d upsampled | d skip]
            self.saved.pop(("upout", i), None)
            dskips[i] = (dcat, u["cout"] // 8, spec["down"][i]["cout"])
            du = dcat.view(u["cout"], 0)
            up_in = self.saved.pop(("ups", i))["in"]
            dnext = self.buffer(n, u["cin"], exts[i + 1])
            if u["w"] is None:
                lib.upsample_trilinear2_backward(du, dnext.view(u["cin"]))
            else:
                # BlurConvTranspose3d: weight (cin, cout, 4, 4, 4) -- wgrad with A = x, B = dy
                g = lib.wgrad(up_in, du, 4, 2, 1, self.device)                 # [tap][cin_pad][cout_pad]
                grads[u["w"]] = g[:, :u["cin"], :u["cout"]].permute(1, 2, 0).reshape(u["cin"], u["cout"], 4, 4, 4)
                wt = params[u["w"]].detach()
                self.conv(du, _pack(wt.permute(2, 3, 4, 1, 0)), u["cin"], dnext.view(u["cin"]), ksize=4, stride=2, pad=1)
            dcur_buf, dcur = dnext, dnext.view(u["cin"])
        for i in reversed(range(depth)):                   # down blocks, deepest first
            b = spec["down"][i]
            dsrc = self.block_backward(("down", i), b, params, dcur, grads, need_dsrc=i > 0)
            if i == 0:
                break
            d = spec["downs"][i - 1]
            sv = self.saved.pop(("downs", i - 1))
            dcat, off, skip_c = dskips.pop(i - 1)
            dout = self.buffer(n, d["cin"], exts[i - 1])
            if d["w"] is None:
                lib.avgpool2_backward(dsrc.view(d["cout"]), dout.view(d["cin"]), add=dcat.view(skip_c, off))
            else:
                # BlurConv3d: weight (cout, cin, 4, 4, 4) -- wgrad with A = dz (coarse grid), B = x
                g = lib.wgrad(dsrc.view(d["cout"]), sv["in"], 4, 2, 1, self.device)      # [tap][cout_pad][cin_pad]
                grads[d["w"]] = g[:, :d["cout"], :d["cin"]].permute(1, 2, 0).reshape(d["cout"], d["cin"], 4, 4, 4)
                wb = params[d["w"]].detach()
                # adjoint of the strided conv = the transposed geometry; the skip connection's gradient rides in as
                # the residual of the epilogue
                self.conv(dsrc.view(d["cout"]), _pack(wb.permute(2, 3, 4, 0, 1)), d["cin"], dout.view(d["cin"]), ksize=4,
                          stride=2, pad=1, transposed=True, residual=dcat.view(skip_c, off))
            dcur_buf, dcur = dout, dout.view(d["cin"])
        self.saved.clear()
        return grads


class _TrainStep(torch.autograd.Function):
    @staticmethod
    def forward(ctx, runner, x, *params):
        ctx.runner = runner
        ctx.params = params
        with runner.lib.on_device(x):
            probs = runner.forward(x, params)
        return probs

    @staticmethod
    def backward(ctx, dprobs):
        runner = ctx.runner
        with runner.lib.on_device(dprobs):
            grads = runner.backward(dprobs, ctx.params)
        out = []
        for p, g in zip(ctx.params, grads):
            out.append(None if g is None else g.to(p.dtype).reshape(p.shape))
        return (None, None, *out)


def forward_train(model, x: torch.Tensor) -> torch.Tensor:
    """``model(x)`` in training mode (see the module docstring)."""
    if not x.is_cuda:
        raise RuntimeError("segmentation_pipeline.models (b200) runs on CUDA tensors only (no CPU fallback)")
    if x.dtype != torch.float32:
        raise Unsupported("training: fp32 inputs only")
    spec, params = build_spec(model)
    first = params[0]
    if first.device != x.device or first.dtype != torch.float32:
        raise RuntimeError(f"training: parameters must be fp32 on {x.device}")
    from . import _engine
    runner = _Runner(spec, x.device, _engine._resolve_precision(model, x))
    if not torch.is_grad_enabled() or not any(p.requires_grad for p in params):
        with runner.lib.on_device(x):
            probs = runner.forward(x, [p.detach() for p in params])
        runner.saved.clear()
        return probs
    return _TrainStep.apply(runner, x, *params)
