"""The two label transforms that sit on the hot path between aggregation and Dice -- device versions of the
reference's ``CustomArgMax`` / ``CustomOneHot`` (transforms/custom_label_transforms.py:211-272) as plain
functions on tensors.  The torchio Transform subclasses themselves (history, inverse, include/exclude) belong
to the torchio layer, which is not rebuilt."""
from __future__ import annotations

import torch


def custom_argmax(data: torch.Tensor) -> torch.Tensor:
    """``torch.argmax(data, dim=0, keepdim=True)`` (:267) for a (C, W, H, D) CUDA tensor -> int64 (1, W, H, D)."""
    import b200seg
    if not data.is_cuda:
        raise RuntimeError("custom_argmax runs on CUDA tensors (no CPU fallback)")
    probs = data.detach().to(torch.float32).contiguous()
    labels = torch.empty(probs.shape[1:], dtype=torch.int64, device=probs.device)
    b200seg.argmax(probs, labels, None)
    return labels[None]


def custom_one_hot(data: torch.Tensor, num_classes: int) -> torch.Tensor:
    """(1, W, H, D) integer labels -> (num_classes, W, H, D) one-hot of the same dtype (:236-238); pure index
    expansion, kept as a tensor expression."""
    labels = data[0].long()
    classes = torch.arange(num_classes, device=labels.device).view(-1, 1, 1, 1)
    return (labels[None] == classes).to(data.dtype)
