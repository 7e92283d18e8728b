"""Shared helpers for the tests (fixture loading, synthetic inputs)."""
import json
import math
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


# ------------------------------------------------------------------------------------------------- label-agreement harness
# The north-star criterion "argmax labels >= 99.9 % voxel agreement in bf16" is about a network whose outputs are
# CONFIDENT, as a trained segmentation network's are.  A random-initialised out_conv puts every voxel near p = 1/C, so
# its argmax is decided by the last bits of the logits and the agreement says nothing about the kernels.  Scaling
# out_conv cannot help (argmax is invariant to it).  SURVEY.md section 7 therefore prescribes a head "so that outputs
# are confident": here out_conv is a linear read-out FITTED on the CPU (multinomial logistic regression, fp64 L-BFGS)
# on the oracle's out_conv input features to the region map the synthetic volume was generated from -- the closed-form
# analogue of training the last layer.  The volume looks like the reference's data (MS lesions / deep grey-matter
# nuclei): a large background and a handful of compact foreground structures.  The same state_dict (with that head) is
# then run by the oracle and by the CUDA path and the labels are compared on EVERY voxel (no margin filter).
def blob_volume(channels: int, spatial, n_foreground: int, seed: int, noise: float = 0.05, distinct: bool = True):
    """-> (volume fp32 (C, W, H, D), region map int64 (W, H, D)).  Background = class 0 with code 0; foreground
    structure k is an ellipsoid (semi-axes 9..13 voxels at 96^3, scaled with the extent) whose voxels carry a fixed
    code on the unit circle of the first two channels (further channels: a fixed random code).  ``distinct``: every
    structure is its own class (multi-class configs) or all of them are class 1 (lesion / no-lesion)."""
    g = torch.Generator().manual_seed(seed)
    w, h, d = spatial
    zz, yy, xx = torch.meshgrid(torch.arange(w), torch.arange(h), torch.arange(d), indexing="ij")
    region = torch.zeros(tuple(spatial), dtype=torch.long)
    cells = [(a, b, c) for a in range(3) for b in range(3) for c in range(2)]
    order = torch.randperm(len(cells), generator=g).tolist()
    scale = min(w, h, d) / 96.0
    for k in range(n_foreground):
        a, b, c = cells[order[k % len(cells)]]
        jit = (torch.rand(3, generator=g) - 0.5) * torch.tensor([6.0, 6.0, 10.0]) * scale
        cz, cy, cx = (a + 0.5) * w / 3 + jit[0], (b + 0.5) * h / 3 + jit[1], (c + 0.5) * d / 2 + jit[2]
        r = (torch.rand(3, generator=g) * 4 + 9) * scale
        inside = ((zz - cz) / r[0]) ** 2 + ((yy - cy) / r[1]) ** 2 + ((xx - cx) / r[2]) ** 2 < 1
        region[inside] = k + 1
    ang = torch.arange(n_foreground) * (2 * math.pi / max(n_foreground, 1))
    codes = torch.zeros(n_foreground + 1, channels)
    codes[1:, 0] = torch.cos(ang)
    if channels > 1:
        codes[1:, 1] = torch.sin(ang)
    if channels > 2:
        codes[1:, 2:] = torch.randn(n_foreground, channels - 2, generator=g).clamp(-1, 1)
    vol = codes[region].permute(3, 0, 1, 2) + noise * torch.randn(channels, w, h, d, generator=g)
    if not distinct:
        region = (region > 0).long()
    return vol.contiguous(), region


def fit_readout(features: torch.Tensor, region: torch.Tensor, n_classes: int, l2: float = 1e-3, iters: int = 200,
                samples: int = 200000, seed: int = 0):
    """Multinomial logistic regression of the region map on the (F, W, H, D) features (standardised, L2-regularised,
    L-BFGS in fp64) -> (out_conv.weight (n_classes, F, 3, 3, 3) with only the centre tap set, out_conv.bias)."""
    f = features.reshape(features.shape[0], -1).T.double()
    idx = torch.randperm(f.shape[0], generator=torch.Generator().manual_seed(seed))[:samples]
    f, y = f[idx], region.reshape(-1)[idx]
    mu, sdv = f.mean(0), f.std(0).clamp_min(1e-6)
    fn = (f - mu) / sdv
    w = torch.zeros(f.shape[1], n_classes, dtype=torch.double, requires_grad=True)
    b = torch.zeros(n_classes, dtype=torch.double, requires_grad=True)
    opt = torch.optim.LBFGS([w, b], max_iter=iters, line_search_fn="strong_wolfe")

    def closure():
        opt.zero_grad()
        loss = torch.nn.functional.cross_entropy(fn @ w + b, y) + l2 * (w ** 2).sum()
        loss.backward()
        return loss

    opt.step(closure)
    wf = w.detach() / sdv[:, None]
    bf = b.detach() - (mu[:, None] * wf).sum(0)
    weight = torch.zeros((n_classes, features.shape[0], 3, 3, 3), dtype=torch.float32)
    weight[:, :, 1, 1, 1] = wf.T.float()
    return weight, bf.float().contiguous()


def label_agreement_report(ref_probs: torch.Tensor, got_probs: torch.Tensor, name: str = "") -> dict:
    """Unfiltered label agreement of two (N, C, ...) or (C, ...) probability tensors + the oracle's top-2 margin
    histogram of the DISAGREEING voxels (printed into the test log)."""
    cdim = 1 if ref_probs.dim() == 5 else 0
    lr, lg = ref_probs.argmax(cdim), got_probs.argmax(cdim)
    top2 = torch.topk(ref_probs, 2, dim=cdim).values
    margin = top2.select(cdim, 0) - top2.select(cdim, 1)
    dis = margin[lr != lg]
    edges = [0.0, 1e-4, 1e-3, 1e-2, 5e-2, 1.0 + 1e-6]
    hist = [int(((dis >= lo) & (dis < hi)).sum()) for lo, hi in zip(edges[:-1], edges[1:])]
    n_classes = ref_probs.shape[cdim]
    rep = {
        "name": name, "voxels": int(lr.numel()), "agreement": float((lr == lg).float().mean()),
        "class_fractions": [round(float((lr == c).float().mean()), 4) for c in range(n_classes)],
        "median_margin": float(margin.median()),
        "disagreeing": int(dis.numel()), "disagree_margin_max": float(dis.max()) if dis.numel() else 0.0,
        "disagree_margin_hist": dict(zip(["<1e-4", "<1e-3", "<1e-2", "<5e-2", ">=5e-2"], hist)),
    }
    print("LABEL-AGREEMENT " + json.dumps(rep))
    return rep
