"""Lowering of the U-Net modules to a flat kernel plan, and the weight packers of libb200seg.

A *plan* is a list of ops over named blocked buffers ([N][C8][Z][Y][X][8], see include/b200seg.h):

    ConvOp      one convolution + fused epilogue (folded BatchNorm / bias, activation, residual add, optional
                second destination so that conv0 and res_conv -- which read the same tensor -- run as ONE
                contraction, optional softmax + fp32 NCDHW store for the last layer)
    PoolOp      AvgPool3d(2, 2)                      UpsampleOp   trilinear x2, align_corners=True
    SoftmaxOp   StochasticMatrix / un-fused softmax on the fp32 NCDHW output

Channel concatenation never materialises: every producer writes into a chunk range of the consumer's input
buffer (``Ref`` = buffer name + first chunk + logical channel count).

Reference behaviour restated here (file:line in /root/reference/segmentation_pipeline/models/):
  Block3d.forward components.py:62-73 - BlurConv3d.forward :111-121 - BlurConvTranspose3d.forward :144-154 -
  WSConv3d.forward :81-88 - ModularUNet.forward modular_unet.py:86-102 - NestedResUNet.forward
  nested_residual_unet.py:88-106 - StochasticMatrix.forward components.py:170-185.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F
from torch import nn

K3, DOWN, UP = 0, 1, 2          # tensor-core engine modes (B200SEG_TC_*)
K3T = 3                         # K3 for a final layer with cout <= 4: the nine in-plane taps are accumulator columns
K3T_MAX_COUT = 4
HX, HY = 10, 18                 # halo extent of the engine's 8 x 16 tile


def c8(channels: int) -> int:
    return (channels + 7) // 8


# ------------------------------------------------------------------------------------------------- IR
@dataclass(frozen=True)
class Ref:
    buf: str
    off: int      # first chunk inside the buffer
    c: int        # logical channels


@dataclass
class ConvOp:
    mode: int                       # K3 / DOWN / UP
    src: Ref
    segments: List[Tuple[int, int]]  # (chunk offset relative to src.off, channels) of each concatenated input
    weight: torch.Tensor            # fp32, regular: (Cout, Cin, k,k,k); UP: (Cin, Cout, k,k,k) -- effective taps
    scale: np.ndarray               # (Cout,)
    shift: np.ndarray
    slope: np.ndarray
    dst0: Optional[Ref] = None
    dst1: Optional[Ref] = None
    split: int = 0
    residual: Optional[Ref] = None
    final: bool = False             # write fp32 NCDHW output
    softmax: bool = False
    name: str = ""

    @property
    def cout(self) -> int:
        return int(self.scale.shape[0])


@dataclass
class PoolOp:
    src: Ref
    dst: Ref
    src_level: int = 0


@dataclass
class UpsampleOp:
    src: Ref
    dst: Ref
    src_level: int = 0


@dataclass
class SoftmaxOp:
    sm_channels: int = 0
    diag_bias: float = 0.0


@dataclass
class UnpackOp:
    """blocked 'out' buffer -> fp32 NCDHW result (heads wider than the fused final epilogue)."""


@dataclass
class InstNormOp:
    """nn.InstanceNorm3d in place on ``ref`` + activation (+ residual): cannot be folded into the conv epilogue
    because its statistics depend on the whole conv output."""
    ref: Ref
    gamma: Optional[np.ndarray]
    beta: Optional[np.ndarray]
    eps: float
    slope: float
    residual: Optional[Ref] = None


@dataclass
class Plan:
    in_channels: int
    out_channels: int
    buffers: Dict[str, Tuple[int, int]] = field(default_factory=dict)   # name -> (chunks, level)
    ops: list = field(default_factory=list)
    levels: int = 1                 # input extent must be divisible by 2**(levels-1)
    out_scale: int = 0              # log2 of output/input extent (0 for the U-Nets, -1 down, +1 up)
    op_levels: Dict[int, int] = field(default_factory=dict)

    def add_buffer(self, name: str, channels_or_chunks: int, level: int, chunks: bool = False) -> str:
        n = channels_or_chunks if chunks else c8(channels_or_chunks)
        self.buffers[name] = (n, level)
        return name


class UnsupportedModule(NotImplementedError):
    pass


# ------------------------------------------------------------------------------------------------- parameter folding
def _standardize(weight: torch.Tensor) -> torch.Tensor:
    weight = weight - weight.mean(dim=(1, 2, 3, 4), keepdim=True)
    return weight / (weight.std(dim=(1, 2, 3, 4), keepdim=True) + 1e-5)


def _conv_weight_bias(conv: nn.Module):
    """Effective fp32 (weight, bias) of a 3x3x3 'same' convolution module."""
    from .components import WSConv3d
    w = conv.weight.detach().float().cpu()
    if tuple(conv.kernel_size) != (3, 3, 3) or tuple(conv.stride) != (1, 1, 1) or \
            tuple(conv.padding) != (1, 1, 1) or tuple(conv.dilation) != (1, 1, 1) or conv.groups != 1:
        raise UnsupportedModule(f"only 3x3x3 stride-1 padding-1 convolutions are lowered, got {conv}")
    if isinstance(conv, WSConv3d):
        return _standardize(w), None   # reference never passes the bias (components.py:86)
    if type(conv) is not nn.Conv3d:
        raise UnsupportedModule(f"convolution class {type(conv).__name__} is not lowered")
    b = None if conv.bias is None else conv.bias.detach().float().cpu()
    return w, b


def _norm_affine(norm: Optional[nn.Module], cout: int):
    """(scale, shift) of an eval-mode normalisation layer."""
    if norm is None or isinstance(norm, nn.Identity):
        return np.ones(cout, np.float64), np.zeros(cout, np.float64)
    if isinstance(norm, (nn.BatchNorm3d, nn.InstanceNorm3d)):
        if norm.running_mean is None:
            raise UnsupportedModule("BatchNorm3d without running statistics")
        var = norm.running_var.detach().double().cpu().numpy()
        mean = norm.running_mean.detach().double().cpu().numpy()
        gamma = np.ones(cout) if norm.weight is None else norm.weight.detach().double().cpu().numpy()
        beta = np.zeros(cout) if norm.bias is None else norm.bias.detach().double().cpu().numpy()
        scale = gamma / np.sqrt(var + norm.eps)
        return scale, beta - mean * scale
    raise UnsupportedModule(f"normalisation {type(norm).__name__} is not lowered yet (BatchNorm3d eval / none are)")


def _instance_norm_params(norm: Optional[nn.Module], cout: int):
    """-> (gamma, beta, eps) when ``norm`` is an InstanceNorm3d that normalises with per-sample statistics
    (the default: track_running_stats=False, or training statistics are not available), else None."""
    if not isinstance(norm, nn.InstanceNorm3d):
        return None
    if norm.track_running_stats and norm.running_mean is not None:
        return None     # eval mode uses the running statistics: folds like BatchNorm
    gamma = None if norm.weight is None else norm.weight.detach().float().cpu().numpy()
    beta = None if norm.bias is None else norm.bias.detach().float().cpu().numpy()
    return gamma, beta, float(norm.eps)


def _act_slope(act: Optional[nn.Module], cout: int) -> np.ndarray:
    if act is None or isinstance(act, nn.Identity):
        return np.ones(cout, np.float32)
    if isinstance(act, nn.ReLU):
        return np.zeros(cout, np.float32)
    if isinstance(act, nn.LeakyReLU):
        return np.full(cout, act.negative_slope, np.float32)
    raise UnsupportedModule(f"activation {type(act).__name__} is not lowered (ReLU / LeakyReLU / none are)")


def _epilogue_arrays(cout, bias, norm, act):
    scale, shift = _norm_affine(norm, cout)
    if bias is not None:
        shift = shift + bias.double().numpy() * scale
    return scale.astype(np.float32), shift.astype(np.float32), _act_slope(act, cout)


def blur_effective_weight(module) -> torch.Tensor:
    """The 4^3 kernel BlurConv3d / BlurConvTranspose3d actually convolve with (components.py:118 / :151)."""
    w = module.weight.detach().float().cpu()
    if module.weight_standardization:
        w = _standardize(w)
    return F.conv3d(w, module.kernel.detach().float().cpu(), padding=1, groups=module.in_channels)


# ------------------------------------------------------------------------------------------------- lowering
def _lower_block(plan: Plan, level: int, name: str, src: Ref, segments, convs, norms, acts, res_conv,
                 out: Ref) -> None:
    """[conv_i -> norm_i -> act_i] * n, plus ``res_conv(x_in) + x`` after the last activation."""
    n = len(convs)
    cout = convs[0].out_channels
    res_ref = None
    ops_start = len(plan.ops)
    if res_conv is not None:
        rw, rb = _conv_weight_bias(res_conv)
        r_scale, r_shift, r_slope = _epilogue_arrays(cout, rb, None, None)
        res_ref = Ref(plan.add_buffer(f"{name}.res", cout, level), 0, cout)
    cur, cur_segments = src, segments
    for i, conv in enumerate(convs):
        w, b = _conv_weight_bias(conv)
        last = i == n - 1
        dst = out if last else Ref(plan.add_buffer(f"{name}.t{i}", cout, level), 0, cout)
        inst = _instance_norm_params(norms[i], cout)
        if inst is not None:
            # statistics need the whole conv output: raw conv (+bias) first, then the in-place norm kernel, which
            # also applies the activation and, on the last conv, the residual add
            scale, shift, slope = _epilogue_arrays(cout, b, None, None)
        else:
            scale, shift, slope = _epilogue_arrays(cout, b, norms[i], acts[i])
        if inst is not None:
            if i == 0 and res_conv is not None and not last and cout % 8 == 0:
                # conv0 (raw, + bias) and res_conv read the same tensor: ONE contraction with N = 2 * Cout and two
                # destinations, exactly as on the BatchNorm branch; the norm kernel then works in place on dst
                plan.ops.append(ConvOp(K3, cur, cur_segments, torch.cat([w, rw], 0), np.concatenate([scale, r_scale]),
                                       np.concatenate([shift, r_shift]), np.concatenate([slope, r_slope]), dst0=dst,
                                       dst1=res_ref, split=cout, name=f"{name}.conv0+res"))
            else:
                if i == 0 and res_conv is not None:
                    plan.ops.append(ConvOp(K3, cur, cur_segments, rw, r_scale, r_shift, r_slope, dst0=res_ref,
                                           name=f"{name}.res_conv"))
                plan.ops.append(ConvOp(K3, cur, cur_segments, w, scale, shift, slope, dst0=dst, name=f"{name}.conv{i}"))
            gamma, beta, eps = inst
            plan.ops.append(InstNormOp(dst, gamma, beta, eps, float(_act_slope(acts[i], 1)[0]),
                                       residual=res_ref if (last and res_conv is not None) else None))
        elif i == 0 and res_conv is not None and not last and cout % 8 == 0:
            # conv0 and res_conv read the same tensor: one contraction with N = 2*Cout, two destinations
            op = ConvOp(K3, cur, cur_segments, torch.cat([w, rw], 0), np.concatenate([scale, r_scale]),
                        np.concatenate([shift, r_shift]), np.concatenate([slope, r_slope]), dst0=dst, dst1=res_ref,
                        split=cout, name=f"{name}.conv0+res")
            plan.ops.append(op)
        else:
            if i == 0 and res_conv is not None:
                plan.ops.append(ConvOp(K3, cur, cur_segments, rw, r_scale, r_shift, r_slope, dst0=res_ref,
                                       name=f"{name}.res_conv"))
            plan.ops.append(ConvOp(K3, cur, cur_segments, w, scale, shift, slope, dst0=dst,
                                   residual=res_ref if last else None, name=f"{name}.conv{i}"))
        cur, cur_segments = dst, [(0, cout)]
    for k in range(ops_start, len(plan.ops)):
        plan.op_levels[k] = level


def _split_block3d(block):
    from .components import Block3d
    if not isinstance(block, Block3d):
        raise UnsupportedModule(f"block class {type(block).__name__} is not lowered (Block3d is)")
    convs, norms, acts = [], [], []
    i = 0
    while hasattr(block.layers, f"conv{i}"):
        convs.append(getattr(block.layers, f"conv{i}"))
        norms.append(getattr(block.layers, f"norm{i}", None))
        acts.append(getattr(block.layers, f"activation{i}", None))
        i += 1
    return convs, norms, acts, (block.res_conv if block.residual else None)


def _check_pool(m):
    def triple(v):
        return (v,) * 3 if isinstance(v, int) else tuple(v)
    if not isinstance(m, nn.AvgPool3d) or triple(m.kernel_size) != (2, 2, 2) or triple(m.stride) != (2, 2, 2) \
            or triple(m.padding) != (0, 0, 0) or m.ceil_mode:
        raise UnsupportedModule(f"downsample {m} is not lowered (AvgPool3d(2,2) and BlurConv3d(3, stride 2) are)")


def _check_upsample(m):
    sf = m.scale_factor
    sf = (sf,) * 3 if not isinstance(sf, (tuple, list)) else tuple(sf)
    if not isinstance(m, nn.Upsample) or tuple(float(s) for s in sf) != (2.0, 2.0, 2.0) or m.mode != 'trilinear' \
            or not m.align_corners:
        raise UnsupportedModule(f"upsample {m} is not lowered (trilinear x2 align_corners=True and "
                                f"BlurConvTranspose3d(3, stride 2) are)")


def _check_blur(m, transposed):
    ok = tuple(m.kernel_size) == (3, 3, 3) and tuple(m.stride) == (2, 2, 2) and tuple(m.padding) == (1, 1, 1) \
        and tuple(m.dilation) == (1, 1, 1) and m.groups == 1
    if transposed:
        ok = ok and tuple(m.output_padding) == (0, 0, 0)
    if not ok:
        raise UnsupportedModule(f"{type(m).__name__} is lowered only for kernel 3, stride 2, padding 1")


def _lower_hypothesis(plan: Plan, hypothesis, out_channels: int):
    """-> (fused_softmax flag for the last conv, extra ops)."""
    from .components import StochasticMatrix
    if isinstance(hypothesis, nn.Softmax):
        if hypothesis.dim != 1:
            raise UnsupportedModule("only Softmax(dim=1) is lowered")
        if out_channels <= FINAL_MAX_COUT:
            return True, []
        return False, [SoftmaxOp()]
    if isinstance(hypothesis, nn.Identity):
        return False, []
    if isinstance(hypothesis, StochasticMatrix):
        if hypothesis.channels ** 2 != out_channels:
            raise RuntimeError("Expected dim 1 of input tensor to be the square of the number of out channels")
        return False, [SoftmaxOp(hypothesis.channels, float(hypothesis.diag_bias or 0.0))]
    raise UnsupportedModule(f"hypothesis {type(hypothesis).__name__} is not lowered")


FINAL_MAX_COUT = 16      # widest layer the engines' fused NCDHW / softmax epilogue writes (make_depilogue, conv_direct)


def _lower_head(plan: Plan, hypothesis, src: Ref, segments, out_w, out_b) -> None:
    """out_conv + hypothesis.  Up to 16 output channels the last conv writes fp32 NCDHW itself (softmax fused); wider
    heads (a StochasticMatrix with C >= 5, more than 16 classes) go through a blocked 'out' buffer that the engine
    unpacks before the separate softmax pass."""
    out_ch = out_w.shape[0]
    fused, extra = _lower_hypothesis(plan, hypothesis, out_ch)
    scale, shift, slope = _epilogue_arrays(out_ch, out_b, None, None)
    plan.op_levels[len(plan.ops)] = 0
    if out_ch <= FINAL_MAX_COUT:
        plan.ops.append(ConvOp(K3, src, segments, out_w, scale, shift, slope, final=True, softmax=fused,
                               name="out_conv"))
    else:
        dst = Ref(plan.add_buffer("out", out_ch, 0), 0, out_ch)
        plan.ops.append(ConvOp(K3, src, segments, out_w, scale, shift, slope, dst0=dst, name="out_conv"))
        plan.ops.append(UnpackOp())
    plan.ops.extend(extra)


def lower_modular_unet(model) -> Plan:
    from .components import BlurConv3d, BlurConvTranspose3d
    depth = model.depth
    blocks = [_split_block3d(b) for b in model.down_blocks]
    filters = [b[0][0].out_channels for b in blocks]
    in_ch = blocks[0][0][0].in_channels
    out_w, out_b = _conv_weight_bias(model.out_conv)
    out_ch = out_w.shape[0]
    plan = Plan(in_ch, out_ch, levels=depth)
    plan.add_buffer("in", in_ch, 0)
    # concat buffer of decoder level i: [upsampled f[i+1] | skip f[i]]  (modular_unet.py:97: upsampled FIRST)
    for i in range(depth - 1):
        plan.add_buffer(f"cat{i}", c8(filters[i + 1]) + c8(filters[i]), i, chunks=True)
    x = Ref("in", 0, in_ch)
    for i in range(depth):
        convs, norms, acts, res = blocks[i]
        if i < depth - 1:
            out = Ref(f"cat{i}", c8(filters[i + 1]), filters[i])
        else:
            out = Ref(plan.add_buffer(f"bottom", filters[i], i), 0, filters[i])
        _lower_block(plan, i, f"down{i}", x, [(0, x.c)], convs, norms, acts, res, out)
        if i < depth - 1:
            nxt = Ref(plan.add_buffer(f"pool{i}", filters[i], i + 1), 0, filters[i])
            m = model.downsampling[i]
            if isinstance(m, BlurConv3d):
                _check_blur(m, False)
                ones, zeros = np.ones(filters[i], np.float32), np.zeros(filters[i], np.float32)
                plan.op_levels[len(plan.ops)] = i
                plan.ops.append(ConvOp(DOWN, out, [(0, filters[i])], blur_effective_weight(m), ones, zeros, ones.copy(),
                                       dst0=nxt, name=f"downsampling{i}"))
            else:
                _check_pool(m)
                plan.ops.append(PoolOp(out, nxt, i))
            x = nxt
        else:
            x = out
    for i in reversed(range(depth - 1)):
        up_dst = Ref(f"cat{i}", 0, filters[i + 1])
        m = model.upsampling[i]
        if isinstance(m, BlurConvTranspose3d):
            _check_blur(m, True)
            ones, zeros = np.ones(filters[i + 1], np.float32), np.zeros(filters[i + 1], np.float32)
            plan.op_levels[len(plan.ops)] = i + 1
            plan.ops.append(ConvOp(UP, x, [(0, x.c)], blur_effective_weight(m), ones, zeros, ones.copy(), dst0=up_dst,
                                   name=f"upsampling{i}"))
        else:
            _check_upsample(m)
            plan.ops.append(UpsampleOp(x, up_dst, i + 1))
        convs, norms, acts, res = _split_block3d(model.up_blocks[i])
        cat = Ref(f"cat{i}", 0, c8(filters[i + 1]) * 8 + filters[i])
        out = Ref(plan.add_buffer(f"up{i}", filters[i], i), 0, filters[i])
        _lower_block(plan, i, f"up{i}", cat, [(0, filters[i + 1]), (c8(filters[i + 1]), filters[i])], convs, norms,
                     acts, res, out)
        x = out
    _lower_head(plan, model.hypothesis, x, [(0, x.c)], out_w, out_b)
    return plan


def lower_nested_res_unet(model) -> Plan:
    """UNet++ wiring of nested_residual_unet.py:88-106; skip tensors come FIRST in every concat."""
    f = model.conv1_0.out_ch
    in_ch = model.conv0_0.conv1.in_channels
    out_w, out_b = _conv_weight_bias(model.out_conv)
    out_ch = out_w.shape[0]
    plan = Plan(in_ch, out_ch, levels=4)
    fc = c8(f)
    plan.add_buffer("in", in_ch, 0)
    # consumer input buffers; members listed in concat order
    cat = {
        "conv0_1": ("x0_0", "up(x1_0)"), "conv1_1": ("x1_0", "up(x2_0)", "down(x0_1)"),
        "conv0_2": ("x0_1", "up(x1_1)"), "conv2_1": ("x2_0", "up(x3_0)", "down(x1_1)"),
        "conv1_2": ("x1_1", "up(x2_1)", "down(x0_2)"), "conv0_3": ("x0_2", "up(x1_2)"),
    }
    level_of = {"conv0_0": 0, "conv1_0": 1, "conv0_1": 0, "conv2_0": 2, "conv1_1": 1, "conv0_2": 0, "conv3_0": 3,
                "conv2_1": 2, "conv1_2": 1, "conv0_3": 0}
    for name, members in cat.items():
        plan.add_buffer(f"{name}.in", fc * len(members), level_of[name], chunks=True)
    where = {}   # tensor name -> Ref of its home (first slot in the concat buffer of the block that leads with it)
    for name, members in cat.items():
        where[members[0]] = Ref(f"{name}.in", 0, f)

    def home(tensor: str, level: int) -> Ref:
        if tensor not in where:
            where[tensor] = Ref(plan.add_buffer(tensor, f, level), 0, f)
        return where[tensor]

    def slot(block: str, member: str) -> Ref:
        return Ref(f"{block}.in", cat[block].index(member) * fc, f)

    def run_block(block: str, src: Ref, segments, out: Ref):
        b = getattr(model, block)
        _lower_block(plan, level_of[block], block, src, segments, [b.conv1, b.conv2], [b.bn1, b.bn2],
                     [b.activation1, b.activation2], b.res_conv if b.residual else None, out)

    def cat_src(block: str):
        k = len(cat[block])
        return Ref(f"{block}.in", 0, (k - 1) * fc * 8 + f), [(j * fc, f) for j in range(k)]

    def pool_to(src: Ref, dst: Ref, src_level: int):
        plan.ops.append(PoolOp(src, dst, src_level))

    def up_to(src: Ref, dst: Ref, src_level: int):
        plan.ops.append(UpsampleOp(src, dst, src_level))

    x_in = Ref("in", 0, in_ch)
    run_block("conv0_0", x_in, [(0, in_ch)], home("x0_0", 0))
    p = Ref(plan.add_buffer("down(x0_0)", f, 1), 0, f)
    pool_to(where["x0_0"], p, 0)
    run_block("conv1_0", p, [(0, f)], home("x1_0", 1))
    up_to(where["x1_0"], slot("conv0_1", "up(x1_0)"), 1)
    run_block("conv0_1", *cat_src("conv0_1"), home("x0_1", 0))

    p = Ref(plan.add_buffer("down(x1_0)", f, 2), 0, f)
    pool_to(where["x1_0"], p, 1)
    run_block("conv2_0", p, [(0, f)], home("x2_0", 2))
    up_to(where["x2_0"], slot("conv1_1", "up(x2_0)"), 2)
    pool_to(where["x0_1"], slot("conv1_1", "down(x0_1)"), 0)
    run_block("conv1_1", *cat_src("conv1_1"), home("x1_1", 1))
    up_to(where["x1_1"], slot("conv0_2", "up(x1_1)"), 1)
    run_block("conv0_2", *cat_src("conv0_2"), home("x0_2", 0))

    p = Ref(plan.add_buffer("down(x2_0)", f, 3), 0, f)
    pool_to(where["x2_0"], p, 2)
    run_block("conv3_0", p, [(0, f)], home("x3_0", 3))
    up_to(where["x3_0"], slot("conv2_1", "up(x3_0)"), 3)
    pool_to(where["x1_1"], slot("conv2_1", "down(x1_1)"), 1)
    run_block("conv2_1", *cat_src("conv2_1"), home("x2_1", 2))
    up_to(where["x2_1"], slot("conv1_2", "up(x2_1)"), 2)
    pool_to(where["x0_2"], slot("conv1_2", "down(x0_2)"), 0)
    run_block("conv1_2", *cat_src("conv1_2"), home("x1_2", 1))
    up_to(where["x1_2"], slot("conv0_3", "up(x1_2)"), 1)
    run_block("conv0_3", *cat_src("conv0_3"), home("x0_3", 0))

    _lower_head(plan, model.hypothesis, where["x0_3"], [(0, f)], out_w, out_b)
    return plan


def lower_single(module) -> Plan:
    """Plans for the building blocks used on their own (Block3d, WSConv3d, BlurConv3d, BlurConvTranspose3d)."""
    from .components import Block3d, BlurConv3d, BlurConvTranspose3d, WSConv3d
    if isinstance(module, Block3d):
        convs, norms, acts, res = _split_block3d(module)
        cin, cout = convs[0].in_channels, convs[0].out_channels
        plan = Plan(cin, cout)
        plan.add_buffer("in", cin, 0)
        out = Ref(plan.add_buffer("out", cout, 0), 0, cout)
        _lower_block(plan, 0, "block", Ref("in", 0, cin), [(0, cin)], convs, norms, acts, res, out)
        return plan
    if isinstance(module, (BlurConv3d, BlurConvTranspose3d)):
        transposed = isinstance(module, BlurConvTranspose3d)
        _check_blur(module, transposed)
        cin, cout = module.in_channels, module.out_channels
        plan = Plan(cin, cout, levels=1 if transposed else 2, out_scale=1 if transposed else -1)
        plan.add_buffer("in", cin, 0)
        out = Ref(plan.add_buffer("out", cout, 1 if not transposed else -1), 0, cout)
        ones, zeros = np.ones(cout, np.float32), np.zeros(cout, np.float32)
        plan.op_levels[0] = 0
        plan.ops.append(ConvOp(UP if transposed else DOWN, Ref("in", 0, cin), [(0, cin)], blur_effective_weight(module),
                               ones, zeros, ones.copy(), dst0=out, name=type(module).__name__))
        return plan
    if isinstance(module, WSConv3d):
        w, _ = _conv_weight_bias(module)
        cout, cin = w.shape[:2]
        plan = Plan(cin, cout)
        plan.add_buffer("in", cin, 0)
        out = Ref(plan.add_buffer("out", cout, 0), 0, cout)
        ones, zeros = np.ones(cout, np.float32), np.zeros(cout, np.float32)
        plan.op_levels[0] = 0
        plan.ops.append(ConvOp(K3, Ref("in", 0, cin), [(0, cin)], w, ones, zeros, ones.copy(), dst0=out, name="WSConv3d"))
        return plan
    raise UnsupportedModule(f"{type(module).__name__} has no native lowering")


# ------------------------------------------------------------------------------------------------- weight packing
def physical_weight(op_weight: torch.Tensor, transposed: bool, segments, n_chunks: int, cout_lo: int, cout_hi: int
                    ) -> torch.Tensor:
    """-> fp32 (k, k, k, n_chunks*8, Cpad): taps x physical input channel x output channel of the slice
    [cout_lo, cout_hi), zero for padding channels.  ``segments`` maps the logical (concatenated) input
    channels to physical chunk positions."""
    w = op_weight
    if transposed:
        w = w.permute(1, 0, 2, 3, 4)              # (Cout, Cin, k,k,k): taps keep their transposed-conv meaning
    w = w[cout_lo:cout_hi]
    cout, cin, k = w.shape[0], w.shape[1], w.shape[2]
    cpad = c8(cout) * 8
    phys = torch.zeros((k, k, k, n_chunks * 8, cpad), dtype=torch.float32)
    logical = 0
    for chunk_off, channels in segments:
        phys[:, :, :, chunk_off * 8: chunk_off * 8 + channels, :cout] = \
            w[:, logical: logical + channels].permute(2, 3, 4, 1, 0)
        logical += channels
    if logical != cin:
        raise RuntimeError(f"segments cover {logical} channels, weight has {cin}")
    return phys


def pack_direct_weight(op: ConvOp, n_chunks: int) -> torch.Tensor:
    """fp32 [k^3][cin_phys][cout_pad] for b200seg_conv3d_direct."""
    phys = physical_weight(op.weight, op.mode == UP, op.segments, n_chunks, 0, op.cout)
    k = phys.shape[0]
    return phys.reshape(k * k * k, n_chunks * 8, phys.shape[-1]).contiguous()


def tc_geometry(mode: int, cin_chunks: int, cout: int) -> dict:
    """Mirror of tc_geometry() in csrc/conv_tc.cu."""
    cpad = c8(9 * cout if mode == K3T else cout) * 8
    blocks = {K3: 3, DOWN: 2, UP: 4, K3T: 3}[mode]
    nb = blocks * cpad + 16
    groups = (cin_chunks + 1) // 2
    # K3T: ONE image whose steps are the chunk groups (one MMA each; the whole plane is one shared-memory stage)
    steps_full = {K3: 9, K3T: groups}.get(mode, 4)
    steps_lone = {K3: 5, K3T: groups}.get(mode, 2)
    n_bimg = {DOWN: 8 * groups, K3T: 1}.get(mode, groups)
    return dict(cpad=cpad, blocks=blocks, nb=nb, steps_full=steps_full, steps_lone=steps_lone, groups=groups,
                lone_last=cin_chunks % 2, n_pass=4 if mode == UP else 1,
                n_bimg=n_bimg, bimg_elems=steps_full * 2 * nb * 8)


def step_units(mode: int, lone: bool, pp: int, st: int):
    """The two K-halves of MMA step ``st``: each is (sy, sx, chunk_in_group) or None (zero weights).
    (sy, sx) is the halo offset the A operand starts at -- must agree with step_desc() in csrc/conv_tc.cu."""
    if mode == K3:
        if not lone:
            return [(st // 3, st % 3, 0), (st // 3, st % 3, 1)]
        if st < 3:
            return [(st, 0, 0), (st, 1, 0)]
        if st == 3:
            return [(0, 2, 0), (1, 2, 0)]
        return [None, (2, 2, 0)]
    py, px = pp >> 1, pp & 1
    sy0 = (1 - py) if mode == DOWN else py
    sx0 = (1 - px) if mode == DOWN else px
    if not lone:
        sy, sx = sy0 + st // 2, sx0 + st % 2
        return [(sy, sx, 0), (sy, sx, 1)]
    return [(sy0 + st, sx0, 0), (sy0 + st, sx0 + 1, 0)]


def tap_of(mode: int, pp: int, sy: int, sx: int):
    """Kernel tap (ty, tx) read at halo offset (sy, sx)."""
    if mode == K3:
        return sy, sx
    py, px = pp >> 1, pp & 1
    if mode == DOWN:
        return 2 * sy - 1 + py, 2 * sx - 1 + px
    return py + 3 - 2 * sy, px + 3 - 2 * sx


def block_tz(mode: int, j: int, zpar: int) -> int:
    """Kernel tap along z served by row block ``j`` of a B image."""
    if mode in (K3, K3T):
        return 2 - j
    if mode == DOWN:
        return (3 if zpar else 2) - 2 * j
    return j


def pack_tc_weight(mode: int, phys: torch.Tensor, cin_chunks: int, cout: int) -> torch.Tensor:
    """bf16 operand image for b200seg_conv3d_tc: [pass][image][step][k-half][row][8] with
    row = (plane block j, cout) followed by 16 zero rows; ``phys`` from physical_weight().

    The layout is defined by pack_tc_image(); after the first weight of a geometry the image is produced through the
    cached gather index instead (one ``take``): re-packing a whole network -- every time the weights changed between two
    ``eval()`` forwards, e.g. the trainer's periodic validation -- costs milliseconds instead of ~0.5 s of Python loops."""
    try:
        idx = tc_gather_index(mode, cin_chunks, cout)
    except UnsupportedModule:
        return pack_tc_image(mode, phys, cin_chunks, cout).to(torch.bfloat16).contiguous()
    flat = phys.reshape(-1)
    img = torch.where(idx >= 0, flat[idx.clamp(min=0)], torch.zeros((), dtype=flat.dtype))
    return img.to(torch.bfloat16).contiguous()


_TC_GATHER = {}


def tc_gather_index(mode: int, cin_chunks: int, cout: int) -> torch.Tensor:
    """int64 tensor shaped like the operand image: for every element the flat index into ``phys`` (k, k, k,
    cin_chunks * 8, round_up(cout, 8)) it is copied from, -1 where the image holds a structural zero.  The packing is
    a pure gather, so running it once on an index-valued ``phys`` yields the map; the training step, which re-packs
    every weight after every optimizer step, then packs on the device with one ``take`` (models/_train.py)."""
    key = (mode, cin_chunks, cout)
    hit = _TC_GATHER.get(key)
    if hit is None:
        k = 3 if mode in (K3, K3T) else 4
        shape = (k, k, k, cin_chunks * 8, c8(cout) * 8)
        count = 1
        for d in shape:
            count *= d
        if count >= 1 << 24:
            raise UnsupportedModule("weight too large for the fp32-exact gather-index trick")
        marks = torch.arange(1, count + 1, dtype=torch.float32).reshape(shape)
        hit = pack_tc_image(mode, marks, cin_chunks, cout).to(torch.int64) - 1
        _TC_GATHER[key] = hit
    return hit


def pack_tc_image(mode: int, phys: torch.Tensor, cin_chunks: int, cout: int) -> torch.Tensor:
    """The operand image of pack_tc_weight in fp32 (before the bf16 rounding)."""
    g = tc_geometry(mode, cin_chunks, cout)
    cpad, nb = g["cpad"], g["nb"]
    img = torch.zeros((g["n_pass"], g["n_bimg"], g["steps_full"], 2, nb, 8), dtype=torch.float32)
    if mode == K3T:
        # step = chunk group; K halves = the two chunks of the group (a lone chunk pairs with zero weights);
        # rows of block j: column (dy*3+dx)*cout + co = weight of tap (2-j, dy, dx), output channel co
        for grp in range(g["groups"]):
            for half in range(2):
                chunk = 2 * grp + half
                if chunk >= cin_chunks:
                    continue
                for j in range(3):
                    for dy in range(3):
                        for dx in range(3):
                            r0 = j * cpad + (dy * 3 + dx) * cout
                            img[0, 0, grp, half, r0:r0 + cout, :] = phys[2 - j, dy, dx, chunk * 8:(chunk + 1) * 8, :cout].t()
        return img
    for ps in range(g["n_pass"]):
        for bi in range(g["n_bimg"]):
            grp = bi % g["groups"]
            lone = bool(g["lone_last"]) and grp == g["groups"] - 1
            if mode == DOWN:
                zpar, pp = bi // (4 * g["groups"]), (bi // g["groups"]) % 4
            elif mode == UP:
                zpar, pp = 0, ps
            else:
                zpar, pp = 0, 0
            for st in range(g["steps_lone"] if lone else g["steps_full"]):
                for half, unit in enumerate(step_units(mode, lone, pp, st)):
                    if unit is None:
                        continue
                    sy, sx, ck = unit
                    chunk = 2 * grp + ck
                    ty, tx = tap_of(mode, pp, sy, sx)
                    for j in range(g["blocks"]):
                        tz = block_tz(mode, j, zpar)
                        # (8 cin, Cpad) -> rows = cout, 8 contiguous input channels
                        img[ps, bi, st, half, j * cpad:(j + 1) * cpad, :] = \
                            phys[tz, ty, tx, chunk * 8:(chunk + 1) * 8, :].t()
    return img
