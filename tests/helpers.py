"""Shared helpers for the tests (fixture loading, synthetic inputs)."""
import json
import os

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")


def load_case(name):
    """-> (meta dict, state_dict of torch tensors, x, y) for a tests/golden/models_*.npz fixture."""
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    meta = json.loads(str(z["meta"]))
    sd = {k[3:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("sd/")}
    return meta, sd, torch.from_numpy(z["x"]), torch.from_numpy(z["y"])


def rel_err(a: torch.Tensor, b: torch.Tensor) -> float:
    """max|a-b| / max|b| -- the metric the parity tolerances are stated in (pure relative error is
    undefined near zero; SURVEY.md section 7 'hard parts')."""
    return float((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30))
