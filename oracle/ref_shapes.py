"""ORACLE (test infrastructure only) -- random state_dicts with the reference's parameter names and shapes, for tests
that need weights without instantiating a product module (``tests/`` only; the product never imports this).

Key names / shapes follow ``/root/reference/segmentation_pipeline/models/nested_residual_unet.py:9-28,58-86``
(130 entries for the whole network; SURVEY.md section 8b lists them)."""
from __future__ import annotations

import math

import torch


def _conv(g, cout, cin):
    bound = 1.0 / math.sqrt(cin * 27)
    return (torch.rand(cout, cin, 3, 3, 3, generator=g) * 2 - 1) * bound


def nested_state_dict(input_channels: int, output_channels: int, filters: int, seed: int = 0) -> dict:
    g = torch.Generator().manual_seed(seed)
    f = filters
    blocks = {"conv0_0": input_channels, "conv1_0": f, "conv0_1": 2 * f, "conv2_0": f, "conv1_1": 3 * f,
              "conv0_2": 2 * f, "conv3_0": f, "conv2_1": 3 * f, "conv1_2": 3 * f, "conv0_3": 2 * f}
    sd = {}
    for name, cin in blocks.items():
        sd[f"{name}.res_conv.weight"] = _conv(g, f, cin)
        sd[f"{name}.res_conv.bias"] = (torch.rand(f, generator=g) * 2 - 1) / math.sqrt(cin * 27)
        for i, c in ((1, cin), (2, f)):
            sd[f"{name}.conv{i}.weight"] = _conv(g, f, c)
            sd[f"{name}.bn{i}.weight"] = 0.5 + torch.rand(f, generator=g)
            sd[f"{name}.bn{i}.bias"] = 0.1 * torch.randn(f, generator=g)
            sd[f"{name}.bn{i}.running_mean"] = 0.1 * torch.randn(f, generator=g)
            sd[f"{name}.bn{i}.running_var"] = 0.5 + torch.rand(f, generator=g)
            sd[f"{name}.bn{i}.num_batches_tracked"] = torch.tensor(0)
    sd["out_conv.weight"] = _conv(g, output_channels, f)
    sd["out_conv.bias"] = torch.zeros(output_channels)
    return sd
