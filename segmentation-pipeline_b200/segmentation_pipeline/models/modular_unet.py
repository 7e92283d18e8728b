"""ModularUNet -- host-side mirror of the reference's models/modular_unet.py:11-102.

Same constructor signature, sub-module names (``down_blocks``, ``downsampling``, ``up_blocks``,
``upsampling``, ``out_conv``, ``hypothesis``) and therefore the same ``state_dict`` keys.  ``forward`` runs the
whole network as one native kernel plan (see ``_plan.lower_modular_unet``): no ATen convolutions, no
``torch.cat`` (producers write straight into channel ranges of the consumer's input buffer).
"""
from __future__ import annotations

from typing import Dict, Optional, Sequence, Union

from torch import nn

from ..utils import is_sequence
from .components import Block3d, _NativeForward
from .utils import filter_kwargs


class ModularUNet(_NativeForward, nn.Module):
    def __init__(
            self,
            in_channels: int,
            out_channels: int,
            filters: Union[int, Sequence[int]],
            depth: int,
            block_class: nn.Module = Block3d,
            block_params: Optional[Dict] = None,
            upsample_class: nn.Module = nn.Upsample,
            upsample_params: Optional[Dict] = None,
            downsample_class: nn.Module = nn.AvgPool3d,
            downsample_params: Optional[Dict] = None,
            out_conv_class: nn.Module = nn.Conv3d,
            out_conv_params: Optional[Dict] = None,
            hypothesis_class: nn.Module = nn.Softmax,
            hypothesis_params: Optional[Dict] = None,
    ):
        super().__init__()

        if isinstance(filters, int):
            filters = [filters] * depth
        elif is_sequence(filters) and len(filters) != depth:
            raise ValueError(f"Sequence of filters {filters} does not match depth {depth}")

        block_params = {} if block_params is None else block_params
        if upsample_params is None:
            upsample_params = {'scale_factor': 2, 'mode': 'trilinear', 'align_corners': True}
        if downsample_params is None:
            downsample_params = {'kernel_size': 2, 'stride': 2, 'count_include_pad': False}
        if out_conv_params is None:
            out_conv_params = {'in_channels': filters[0], 'out_channels': out_channels, 'kernel_size': 3,
                               'padding': 1}
        if hypothesis_params is None:
            hypothesis_params = {"dim": 1}

        self.depth = depth

        # encoder: in -> f0, f0 -> f1, ...
        widths_in = [in_channels] + list(filters[:-1])
        self.down_blocks = nn.ModuleList(
            block_class(c_in, c_out, **block_params) for c_in, c_out in zip(widths_in, filters))

        self.downsampling = nn.ModuleList()
        for level in range(depth - 1):
            width = filters[level]
            downsample_params.update(filter_kwargs(downsample_class, in_channels=width, out_channels=width,
                                                   channels=width))
            self.downsampling.append(downsample_class(**downsample_params))

        # decoder block i consumes cat(upsampled f[i+1], skip f[i])
        self.up_blocks = nn.ModuleList(
            block_class(filters[level] + filters[level + 1], filters[level], **block_params)
            for level in range(depth - 1))

        self.upsampling = nn.ModuleList()
        for level in range(1, depth):
            width = filters[level]
            upsample_params.update(filter_kwargs(upsample_class, in_channels=width, out_channels=width,
                                                 channels=width))
            self.upsampling.append(upsample_class(**upsample_params))

        self.out_conv = out_conv_class(**out_conv_params)
        self.hypothesis = hypothesis_class(**hypothesis_params)
