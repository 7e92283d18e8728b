from __future__ import annotations

from typing import Sequence

import torch


def no_op(x):
    return x


def is_sequence(x):
    return isinstance(x, Sequence) and not isinstance(x, str)


def as_list(x):
    if x is None:
        return []
    return list(x) if is_sequence(x) else [x]


def as_tuple(x):
    if x is None:
        return ()
    return tuple(x) if is_sequence(x) else (x,)


def auto_str(obj) -> str:
    fields = ", ".join(f"{k}={v}" for k, v in vars(obj).items())
    return f"{type(obj).__name__}({fields})"


def collate_subjects(subjects, image_names: Sequence[str], device: torch.device):
    """Stacks ``subject[name]['data']`` of every subject into (B, C, W, H, D) and moves it to ``device``
    (reference utils/utils.py:75-85)."""
    batch = {}
    for image_name in image_names:
        stacked = torch.stack([subject[image_name]["data"] for subject in subjects])
        batch[image_name] = stacked.to(device)
    return batch
