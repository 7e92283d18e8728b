"""NestedResUNet (UNet++ of depth 4) -- host-side mirror of the reference's
models/nested_residual_unet.py:6-106: same constructor, same sub-module names (``conv{r}_{c}`` with
``res_conv / conv1 / bn1 / conv2 / bn2``, ``out_conv``) and ``state_dict`` keys.  ``forward`` is a native
kernel plan (``_plan.lower_nested_res_unet``)."""
from __future__ import annotations

from typing import Dict, Optional

from torch import nn

from .components import _NativeForward


class NestedResUNet(_NativeForward, nn.Module):
    class Block(_NativeForward, nn.Module):
        def __init__(self, in_ch, out_ch, residual=False, dropout_p=0.0):
            super().__init__()
            self.residual = residual
            self.out_ch = out_ch
            if residual:
                self.res_conv = nn.Conv3d(in_ch, out_ch, kernel_size=3, padding=1)
            self.conv1 = nn.Conv3d(in_ch, out_ch, bias=False, kernel_size=3, padding=1)
            self.bn1 = nn.BatchNorm3d(out_ch)
            self.activation1 = nn.ReLU(inplace=True)
            self.conv2 = nn.Conv3d(out_ch, out_ch, bias=False, kernel_size=3, padding=1)
            self.bn2 = nn.BatchNorm3d(out_ch)
            self.activation2 = nn.ReLU(inplace=True)
            self.dropout = nn.Dropout3d(p=dropout_p) if dropout_p != 0.0 else None

    def __init__(
            self,
            input_channels: int,
            output_channels: int,
            filters: int,
            dropout_p: float = 0.0,
            hypothesis_class: nn.Module = nn.Softmax,
            hypothesis_params: Optional[Dict] = None,
    ):
        super().__init__()
        if hypothesis_params is None:
            hypothesis_params = {"dim": 1}

        self.dropout = nn.Dropout3d(p=dropout_p) if dropout_p != 0.0 else None
        self.down = nn.AvgPool3d(kernel_size=2, stride=2, count_include_pad=False)
        self.up = nn.Upsample(scale_factor=2, mode='trilinear', align_corners=True)

        f = filters
        common = dict(dropout_p=dropout_p)
        # (name, input width, residual) in the reference's registration order
        layout = [
            ("conv0_0", input_channels, True), ("conv1_0", f, False), ("conv0_1", 2 * f, True),
            ("conv2_0", f, False), ("conv1_1", 3 * f, False), ("conv0_2", 2 * f, True),
            ("conv3_0", f, False), ("conv2_1", 3 * f, False), ("conv1_2", 3 * f, False), ("conv0_3", 2 * f, True),
        ]
        for name, width, residual in layout:
            setattr(self, name, self.Block(width, f, **common, residual=residual))

        self.out_conv = nn.Conv3d(f, output_channels, kernel_size=3, padding=1)
        self.hypothesis = hypothesis_class(**hypothesis_params)
