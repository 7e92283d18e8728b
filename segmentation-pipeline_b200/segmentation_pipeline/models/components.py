"""Building blocks of the 3D U-Nets -- host-side mirror of the reference's
``segmentation_pipeline/models/components.py`` (Block3d :17-73, WSConv3d :76-88, BlurConv3d :91-121,
BlurConvTranspose3d :124-154, StochasticMatrix :157-185).

The classes keep the reference's constructor signatures, attribute names and ``state_dict`` keys (checkpoints
load with ``strict=True``), but they hold parameters only: ``forward`` lowers the module to a kernel plan and
runs it through libb200seg on the GPU.  There is no ATen compute path and no CPU fallback.
"""
from __future__ import annotations

from collections import OrderedDict
from numbers import Number
from typing import Optional

import torch
from torch import nn

from . import _engine


def _product(values) -> int:
    result = 1
    for v in values:
        result *= v
    return result


class _NativeForward:
    """Mixin: ``forward`` = compile (cached) + run the native plan."""

    def forward(self, x, *args, **kwargs):  # noqa: D401
        return _engine.forward_native(self, x)

    def invalidate(self) -> None:
        """Drops the packed device weights / workspaces of this module; the next forward re-packs them.

        The cache key is (data_ptr, version counter) of every parameter and buffer, which catches ``load_state_dict``,
        ``.to()`` and ordinary in-place ops -- but NOT edits made through ``.data`` (``weight.data.mul_()``, EMA / SWA
        updates, manual BatchNorm-statistic patching), which do not bump the version counter: call ``invalidate()``
        after those.  One module instance serves ONE stream at a time (its workspaces are shared)."""
        self.__dict__.pop("_b200_cache", None)

    def load_state_dict(self, *args, **kwargs):
        self.invalidate()
        return super().load_state_dict(*args, **kwargs)

    def _apply(self, fn, *args, **kwargs):
        self.invalidate()
        return super()._apply(fn, *args, **kwargs)

    def __getstate__(self):
        state = dict(self.__dict__)
        state.pop("_b200_cache", None)
        return state


class Block3d(_NativeForward, nn.Module):
    """``num_convs`` x [conv -> norm -> activation] with an optional parallel residual conv whose result is
    added after the last activation (reference components.py:62-73)."""

    def __init__(
            self,
            in_channels,
            out_channels,
            conv_class=nn.Conv3d,
            conv_params=None,
            normalization_class=nn.BatchNorm3d,
            normalization_params=None,
            activation_class=nn.ReLU,
            activation_params=None,
            residual=False,
            residual_params=None,
            dropout_p=0.0,
            num_convs=2,
    ):
        super().__init__()
        conv_params = {'bias': False, 'kernel_size': 3, 'padding': 1} if conv_params is None else conv_params
        normalization_params = {} if normalization_params is None else normalization_params
        activation_params = {'inplace': True} if activation_params is None else activation_params
        residual_params = {'bias': True, 'kernel_size': 3, 'padding': 1} if residual_params is None \
            else residual_params

        self.residual = residual
        if residual:
            # registered first so that state_dict ordering matches the reference
            self.res_conv = conv_class(in_channels, out_channels, **residual_params)

        stages = OrderedDict()
        for index in range(num_convs):
            stages[f'conv{index}'] = conv_class(in_channels if index == 0 else out_channels, out_channels,
                                                **conv_params)
            if normalization_class is not None:
                stages[f'norm{index}'] = normalization_class(out_channels, **normalization_params)
            if activation_class is not None:
                stages[f'activation{index}'] = activation_class(**activation_params)
        self.layers = nn.Sequential(stages)

        self.dropout = nn.Dropout3d(p=dropout_p) if dropout_p != 0.0 else None


class WSConv3d(_NativeForward, nn.Conv3d):
    """Conv3d whose weight is standardised per output channel at call time (reference :81-88).  As in the
    reference, ``forward`` passes only ``**kwargs`` to the convolution, so the bias parameter is never used."""

    def __init__(self, in_channels, out_channels, kernel_size, **kwargs):
        super().__init__(in_channels, out_channels, kernel_size, **kwargs)
        self.kwargs = kwargs


class BlurConv3d(_NativeForward, nn.Conv3d):
    """Strided conv whose 3^3 weight is box-filtered to an effective 4^3 kernel (reference :111-121).  The
    ``kernel`` buffer is ones/8/prod(stride); the bias parameter exists but is never applied."""

    def __init__(self, in_channels, out_channels, kernel_size, weight_standardization=False, **kwargs):
        super().__init__(in_channels, out_channels, kernel_size, **kwargs)
        self.weight_standardization = weight_standardization
        self.register_buffer('kernel', torch.ones(out_channels, 1, 2, 2, 2) / 8)
        self.kernel = self.kernel / _product(self.stride)  # volume shrinks by stride^3
        self.kwargs = kwargs


class BlurConvTranspose3d(_NativeForward, nn.ConvTranspose3d):
    """Transposed counterpart (reference :144-154).  The blur buffer is normalised by the sum over the WHOLE
    (out_channels,1,2,2,2) tensor and multiplied by prod(stride) -- i.e. taps of 1/out_channels at stride 2;
    this quirk is part of the contract."""

    def __init__(self, in_channels, out_channels, kernel_size, weight_standardization=False, **kwargs):
        super().__init__(in_channels, out_channels, kernel_size, **kwargs)
        self.weight_standardization = weight_standardization
        ones = torch.ones(out_channels, 1, 2, 2, 2)
        self.register_buffer('kernel', ones / torch.sum(ones))
        self.kernel = self.kernel * _product(self.stride)  # volume grows by stride^3
        self.kwargs = kwargs

    def forward(self, x, output_size=None):
        return _engine.forward_native(self, x)


class StochasticMatrix(_NativeForward, nn.Module):
    """
    Reshapes a tensor with shape (N, C * C, ...) to (N, C, C, ...) and applies softmax on dim=1
    """

    def __init__(self, channels: int, diag_bias: Optional[Number] = None):
        super().__init__()
        self.channels = channels
        self.diag_bias = diag_bias
