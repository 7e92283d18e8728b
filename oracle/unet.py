"""ORACLE (test infrastructure only) -- CPU fp32 restatement of the reference 3D U-Net forwards.

Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of
``bench.py`` may import this module.  The product (``segmentation-pipeline_b200/``) never does.

Every function works on a plain ``state_dict`` (name -> fp32 tensor) plus a small ``cfg`` dict, uses only
``torch.nn.functional`` on CPU tensors, and cites the reference lines it follows (paths relative to
``/root/reference/segmentation_pipeline/``).  This is a floating-point path, so the restatement is a torch
fp32 functional program rather than numpy (see the task's oracle rules for floating-point kernels).

PINNED: ``tests/golden/models_*.npz`` hold inputs, state_dicts and outputs produced by the reference's own
``segmentation_pipeline.models`` classes imported in-process (``oracle/make_golden.py``, which only runs in
the authoring container where ``/root/reference`` exists); ``tests/test_oracle_models.py`` checks this
restatement against them to ~1e-6.
"""
from __future__ import annotations

import itertools
from typing import Dict, Optional, Sequence

import torch
import torch.nn.functional as F

SD = Dict[str, torch.Tensor]
BN_EPS = 1e-5

# ---------------------------------------------------------------------------------------------- bf16 emulation
# The CUDA throughput path keeps activations and weights in bf16 (fp32 accumulation, fp32 epilogue).  To predict
# on the CPU what that does to labels, ``emulate_bf16()`` rounds exactly the tensors the kernels store in bf16:
# conv inputs / weights and every block-level activation that goes to HBM.  Outside the context ``_q`` is the
# identity and the oracle is the plain fp32 restatement.
_QUANT = [None]


def _q(t: torch.Tensor) -> torch.Tensor:
    return t if _QUANT[0] is None else _QUANT[0](t)


class emulate_bf16:
    """Context manager: round-to-nearest-even bf16 storage of activations and weights (fp32 arithmetic)."""

    def __enter__(self):
        self._prev = _QUANT[0]
        _QUANT[0] = lambda t: t.to(torch.bfloat16).to(torch.float32)
        return self

    def __exit__(self, *exc):
        _QUANT[0] = self._prev
        return False


# --------------------------------------------------------------------------------------------- components
def _standardize(weight: torch.Tensor) -> torch.Tensor:
    """Weight standardisation, models/components.py:83-84 (WSConv3d) and :114-116 / :147-149."""
    weight = weight - weight.mean(dim=(1, 2, 3, 4), keepdim=True)
    return weight / (weight.std(dim=(1, 2, 3, 4), keepdim=True) + 1e-5)


def ws_conv3d(x, weight, **kwargs):
    """models/components.py:81-88 -- note the reference never passes ``self.bias`` (bias unused)."""
    return F.conv3d(_q(x), _q(_standardize(weight)), **kwargs)


def blur_weight(weight: torch.Tensor, kernel: torch.Tensor, in_channels: int) -> torch.Tensor:
    """models/components.py:118 and :151: the 3^3 weight is box-filtered with the (out_channels,1,2,2,2)
    buffer as a grouped conv with ``groups=in_channels`` and padding 1 -> an effective 4^3 kernel.
    For BlurConv3d the weight is (Cout, Cin, 3,3,3): the conv treats Cout as batch and Cin as channels,
    so it only type-checks when kernel.shape[0] (= out_channels) == in_channels, as in every reference use."""
    return F.conv3d(weight, kernel, padding=1, groups=in_channels)


def blur_conv3d(x, weight, kernel, in_channels, weight_standardization=False, **kwargs):
    """BlurConv3d.forward, models/components.py:111-121.  ``kernel`` = ones/8/prod(stride) (:104-109).
    The bias parameter exists in the state_dict but is never applied (:119)."""
    if weight_standardization:
        weight = _standardize(weight)
    return _q(F.conv3d(_q(x), _q(blur_weight(weight, kernel, in_channels)), **kwargs))


def blur_conv_transpose3d(x, weight, kernel, in_channels, weight_standardization=False, **kwargs):
    """BlurConvTranspose3d.forward, models/components.py:144-154.  ``kernel`` = ones/sum(ones)*prod(stride)
    where sum runs over the WHOLE (out_channels,1,2,2,2) buffer, i.e. each tap is 1/out_channels for
    stride 2 (:136-141) -- a quirk that is part of the contract."""
    if weight_standardization:
        weight = _standardize(weight)
    return _q(F.conv_transpose3d(_q(x), _q(blur_weight(weight, kernel, in_channels)), **kwargs))


def batch_norm_eval(x, sd: SD, prefix: str, eps: float = BN_EPS):
    """nn.BatchNorm3d in eval mode (running statistics), models/components.py:53."""
    return F.batch_norm(x, sd[prefix + "running_mean"], sd[prefix + "running_var"], sd.get(prefix + "weight"),
                        sd.get(prefix + "bias"), training=False, eps=eps)


def _activation(x, act: Optional[str], slope: float = 0.01):
    if act is None or act == "none":
        return x
    if act == "relu":
        return F.relu(x)
    if act == "leaky_relu":
        return F.leaky_relu(x, slope)
    raise ValueError(act)


def batch_norm_train(x, sd: SD, prefix: str, eps: float = BN_EPS, momentum: float = 0.1):
    """nn.BatchNorm3d in TRAINING mode (models/components.py:53 under ``model.train()``,
    segmentation_trainer.py:162): batch statistics; ``sd``'s running_mean / running_var are updated in place."""
    return F.batch_norm(x, sd[prefix + "running_mean"], sd[prefix + "running_var"], sd.get(prefix + "weight"),
                        sd.get(prefix + "bias"), training=True, momentum=momentum, eps=eps)


def _norm(x, sd, prefix, norm: Optional[str], training: bool = False):
    if norm is None or norm == "none":
        return x
    if norm == "batch":
        return batch_norm_train(x, sd, prefix) if training else batch_norm_eval(x, sd, prefix)
    if norm == "instance":
        # nn.InstanceNorm3d defaults: affine=False, track_running_stats=False
        return F.instance_norm(x, weight=sd.get(prefix + "weight"), bias=sd.get(prefix + "bias"), eps=BN_EPS)
    raise ValueError(norm)


def block3d(x, sd: SD, prefix: str, cfg: dict):
    """Block3d.forward, models/components.py:62-73: [conv_i -> norm_i -> act_i]*num_convs, then
    ``res_conv(x_in) + x`` when residual; Dropout3d is the identity in eval mode."""
    x_in = x
    conv = cfg.get("conv", "conv")
    n = cfg.get("num_convs", 2)
    residual = cfg.get("residual", False)
    for i in range(n):
        w = sd[f"{prefix}layers.conv{i}.weight"]
        b = sd.get(f"{prefix}layers.conv{i}.bias")
        if conv == "ws":
            x = ws_conv3d(x, w, padding=1)
        else:
            x = F.conv3d(_q(x), _q(w), b, padding=1)
        x = _norm(x, sd, f"{prefix}layers.norm{i}.", cfg.get("norm", "batch"), cfg.get("bn_training", False))
        x = _activation(x, cfg.get("act", "relu"), cfg.get("slope", 0.01))
        if not (residual and i == n - 1):
            x = _q(x)          # stored activation (the last one is stored after the residual add)
    if residual:
        w = sd[f"{prefix}res_conv.weight"]
        b = sd.get(f"{prefix}res_conv.bias")
        r = ws_conv3d(x_in, w, padding=1) if conv == "ws" else F.conv3d(_q(x_in), _q(w), b, padding=1)
        x = _q(_q(r) + x)
    return x


def stochastic_matrix(x, channels: int, diag_bias=None):
    """StochasticMatrix.forward, models/components.py:170-185."""
    n, c2 = x.shape[:2]
    spatial = x.shape[2:]
    if c2 != channels * channels:
        raise RuntimeError("Expected dim 1 of input tensor to be the square of the number of out channels")
    x = x.reshape(n, channels, channels, *spatial)
    if diag_bias is not None:
        x = x + torch.eye(channels).reshape(1, channels, channels, *(1 for _ in spatial)) * diag_bias
    x = torch.softmax(x, dim=1)
    return x.reshape(n, c2, *spatial)


def _hypothesis(x, cfg: dict):
    hyp = cfg.get("hypothesis", "softmax")
    if hyp == "softmax":
        return torch.softmax(x, dim=1)
    if hyp == "identity":
        return x
    if hyp == "stochastic_matrix":
        return stochastic_matrix(x, cfg["sm_channels"], cfg.get("sm_diag_bias"))
    raise ValueError(hyp)


# --------------------------------------------------------------------------------------------- ModularUNet
def modular_unet_forward(sd: SD, x: torch.Tensor, cfg: dict) -> torch.Tensor:
    """ModularUNet.forward, models/modular_unet.py:86-102.

    cfg keys: depth, filters (list), block (dict for block3d), down in {'avgpool','blur'},
    up in {'trilinear','blur'}, hypothesis in {'softmax','identity','stochastic_matrix'}."""
    depth = cfg["depth"]
    filters = cfg["filters"]
    skips = []
    for i in range(depth):
        x = block3d(x, sd, f"down_blocks.{i}.", cfg["block"])
        if i != depth - 1:
            skips.append(x)
            if cfg.get("down", "avgpool") == "avgpool":
                # nn.AvgPool3d(kernel_size=2, stride=2, count_include_pad=False), modular_unet.py:40-41
                x = _q(F.avg_pool3d(x, 2, 2, count_include_pad=False))
            else:
                x = blur_conv3d(x, sd[f"downsampling.{i}.weight"], sd[f"downsampling.{i}.kernel"], filters[i],
                                cfg.get("down_ws", False), stride=2, padding=1)
    for i in reversed(range(depth - 1)):
        if cfg.get("up", "trilinear") == "trilinear":
            # nn.Upsample(scale_factor=2, mode='trilinear', align_corners=True), modular_unet.py:38-39
            x = _q(F.interpolate(x, scale_factor=2, mode="trilinear", align_corners=True))
        else:
            x = blur_conv_transpose3d(x, sd[f"upsampling.{i}.weight"], sd[f"upsampling.{i}.kernel"],
                                      filters[i + 1], cfg.get("up_ws", False), stride=2, padding=1,
                                      output_padding=0)
        # modular_unet.py:97 -- upsampled tensor FIRST, skip second
        x = block3d(torch.cat([x, skips[i]], dim=1), sd, f"up_blocks.{i}.", cfg["block"])
    if cfg.get("return_features", False):
        return x                    # input of out_conv (tests fit a linear read-out on it)
    x = F.conv3d(_q(x), _q(sd["out_conv.weight"]), sd.get("out_conv.bias"), padding=1)
    return _hypothesis(x, cfg)


# --------------------------------------------------------------------------------------------- NestedResUNet
def _nested_block(x, sd: SD, name: str, residual: bool, cfg: Optional[dict] = None):
    """NestedResUNet.Block.forward, models/nested_residual_unet.py:30-47.  ``cfg['bn_training']``: batch-statistic
    BatchNorm (``model.train()``); ``cfg['dropout_masks'][name]``: the (N, C) multiplier nn.Dropout3d drew for this
    block (0 or 1 / (1 - p) per sample and channel), applied after the residual add (:44-45)."""
    cfg = cfg or {}
    bn = batch_norm_train if cfg.get("bn_training", False) else batch_norm_eval
    x_in = x
    x = F.conv3d(_q(x), _q(sd[f"{name}.conv1.weight"]), None, padding=1)
    x = _q(F.relu(bn(x, sd, f"{name}.bn1.")))
    x = F.conv3d(x, _q(sd[f"{name}.conv2.weight"]), None, padding=1)
    x = F.relu(bn(x, sd, f"{name}.bn2."))
    if residual:
        r = F.conv3d(_q(x_in), _q(sd[f"{name}.res_conv.weight"]), sd[f"{name}.res_conv.bias"], padding=1)
        x = _q(r) + x
    mask = (cfg.get("dropout_masks") or {}).get(name)
    if mask is not None:
        x = x * mask.reshape(*mask.shape, 1, 1, 1)
    return _q(x)


def nested_res_unet_forward(sd: SD, x: torch.Tensor, cfg: Optional[dict] = None) -> torch.Tensor:
    """NestedResUNet.forward, models/nested_residual_unet.py:88-106 (UNet++ of depth 4; residual only on the
    level-0 blocks :72,74,78,83; skip tensors come FIRST in every concat)."""
    cfg = cfg or {}
    down = lambda t: _q(F.avg_pool3d(t, 2, 2, count_include_pad=False))
    up = lambda t: _q(F.interpolate(t, scale_factor=2, mode="trilinear", align_corners=True))
    x0_0 = _nested_block(x, sd, "conv0_0", True, cfg)
    x1_0 = _nested_block(down(x0_0), sd, "conv1_0", False, cfg)
    x0_1 = _nested_block(torch.cat((x0_0, up(x1_0)), 1), sd, "conv0_1", True, cfg)
    x2_0 = _nested_block(down(x1_0), sd, "conv2_0", False, cfg)
    x1_1 = _nested_block(torch.cat((x1_0, up(x2_0), down(x0_1)), 1), sd, "conv1_1", False, cfg)
    x0_2 = _nested_block(torch.cat((x0_1, up(x1_1)), 1), sd, "conv0_2", True, cfg)
    x3_0 = _nested_block(down(x2_0), sd, "conv3_0", False, cfg)
    x2_1 = _nested_block(torch.cat((x2_0, up(x3_0), down(x1_1)), 1), sd, "conv2_1", False, cfg)
    x1_2 = _nested_block(torch.cat((x1_1, up(x2_1), down(x0_2)), 1), sd, "conv1_2", False, cfg)
    x0_3 = _nested_block(torch.cat((x0_2, up(x1_2)), 1), sd, "conv0_3", True, cfg)
    if cfg.get("return_features", False):
        return x0_3                 # input of out_conv (tests fit a linear read-out on it)
    x_out = F.conv3d(_q(x0_3), _q(sd["out_conv.weight"]), sd["out_conv.bias"], padding=1)
    return _hypothesis(x_out, cfg)


# --------------------------------------------------------------------------------------------- ensembles / TTA
def apply_strategy(predictions: Sequence[torch.Tensor], strategy: str) -> torch.Tensor:
    """models/ensemble.py:16-35: 'mean' or 'majority' (argmax -> torch.mode over members -> one_hot)."""
    p = torch.stack(list(predictions))
    if strategy == "mean":
        return torch.mean(p, dim=0)
    if strategy == "majority":
        c = p.shape[2]
        y = torch.argmax(p, dim=2)
        y = torch.mode(y, dim=0).values
        return F.one_hot(y, num_classes=c).movedim(-1, 1)
    raise RuntimeError(f"Invalid prediction strategy {strategy}")


def flip_sets(spatial_dims=(2, 3, 4)):
    """models/ensemble.py:56-58: all subsets of the spatial dims, ordered by subset size."""
    out = []
    for order in range(len(spatial_dims) + 1):
        out += list(itertools.combinations(spatial_dims, order))
    return out


def ensemble_flips(model_fn, x, strategy="mean", spatial_dims=(2, 3, 4)):
    """EnsembleFlips.forward, models/ensemble.py:60-71."""
    preds = []
    for flip in flip_sets(spatial_dims):
        preds.append(model_fn(x.flip(flip)).flip(flip))
    return apply_strategy(preds, strategy)


def ensemble_orientations(model_fn, x, strategy="mean"):
    """EnsembleOrientations.forward, models/ensemble.py:88-103 (6 permutations x 8 flips)."""
    preds = []
    for perm in itertools.permutations((2, 3, 4)):
        inv = tuple((torch.argsort(torch.tensor(perm)) + 2).tolist())
        xp = x.permute(0, 1, *perm)
        for flip in flip_sets():
            preds.append(model_fn(xp.flip(flip)).flip(flip).permute(0, 1, *inv))
    return apply_strategy(preds, strategy)


def ensemble_models(model_fns, x, strategy="mean"):
    """EnsembleModels.forward, models/ensemble.py:44-47."""
    return apply_strategy([fn(x) for fn in model_fns], strategy)


# --------------------------------------------------------------------------------------------- predictor helpers
def split_and_flip(x):
    """prediction.py:16-20."""
    parts = list(x.split(x.shape[2] // 2, dim=2))
    parts[1] = parts[1].flip(2)
    return torch.cat(parts, dim=0)


def reverse_split_and_flip(x):
    """prediction.py:23-27."""
    parts = list(x.split(x.shape[0] // 2, dim=0))
    parts[1] = parts[1].flip(2)
    return torch.cat(parts, dim=2)


# --------------------------------------------------------------------------------------------- criterion
def hybrid_logistic_dice_loss(prediction, target, dice_weight=0.5, logistic_class_weights=None, square_dice=True):
    """HybridLogisticDiceLoss.forward, criterions/hybrid_logistic_dice_loss.py:13-43."""
    dims = (2, 3, 4)
    eps = 1e-8
    overlap = torch.sum(prediction * target, dim=dims)
    if square_dice:
        total = torch.sum(target * target, dim=dims) + torch.sum(prediction * prediction, dim=dims)
    else:
        total = torch.sum(target, dim=dims) + torch.sum(prediction, dim=dims)
    dice = 2 * overlap / (total + eps)
    safe = (prediction + eps) / (1 + eps)
    logistic = torch.mean(target * torch.log(safe), dim=dims)
    if logistic_class_weights is not None:
        logistic = logistic * torch.tensor(logistic_class_weights)[None]
    logistic_loss = torch.mean(-logistic)
    dice_loss = torch.mean(1 - dice)
    return {"loss": (1.0 - dice_weight) * logistic_loss + dice_weight * dice_loss, "dice_loss": dice_loss,
            "logistic_loss": logistic_loss}
