"""ORACLE (test infrastructure only) -- fits the confident read-out head bench.py loads
(tests/golden/readout_msseg2.npz).

    python oracle/make_readout.py        (authoring container; ~1 min on 8 cores)

bench.py's model is the msseg2 ModularUNet with seeded random weights; its random out_conv puts every voxel at
p ~ 0.5, where argmax agreement only measures the last bits of the logits.  SURVEY.md section 7 prescribes a head whose
outputs are confident: here out_conv's centre tap is a multinomial logistic regression (fp64 L-BFGS, CPU) of the
lesion mask on the oracle's out_conv-input features of three 96^3 patches of bench.py's synthetic volume -- the
closed-form analogue of training the last layer.  Only the (2, 40) tap weights and the bias are stored."""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "segmentation-pipeline_b200"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def main() -> None:
    import bench
    from helpers import fit_readout
    from oracle import grid as ogrid, unet
    torch.set_num_threads(os.cpu_count() or 1)
    if os.path.exists(bench.READOUT):
        os.remove(bench.READOUT)
    model = bench.build_model()                       # random head (fixture absent)
    sd = {k: v.clone() for k, v in model.state_dict().items()}
    cfg = {"depth": 6, "filters": bench.FILTERS, "block": {"residual": True}, "down": "blur", "up": "blur",
           "return_features": True}
    vol, mask = bench.synthetic_volume(0, with_mask=True)
    padded = ogrid.pad_volume(vol.numpy(), bench.OVERLAP, bench.PADDING)
    pmask = ogrid.pad_volume(mask.numpy()[None], bench.OVERLAP, bench.PADDING)[0]
    loc = ogrid.grid_locations(padded.shape[1:], bench.PATCH, bench.OVERLAP)
    # patches with the most lesion voxels
    frac = [pmask[l[0]:l[3], l[1]:l[4], l[2]:l[5]].mean() for l in loc]
    picks = np.argsort(frac)[-3:]
    feats, regions = [], []
    for i in picks:
        l = loc[i]
        x = torch.from_numpy(ogrid.extract_patches(padded, l[None]))
        with torch.no_grad():
            feats.append(unet.modular_unet_forward(sd, x, cfg)[0])
        regions.append(torch.from_numpy(pmask[l[0]:l[3], l[1]:l[4], l[2]:l[5]].astype(np.int64)))
        print("patch", i, "lesion fraction", round(float(frac[i]), 4))
    feat = torch.cat([f.reshape(f.shape[0], -1) for f in feats], 1)[:, :, None, None]
    region = torch.cat([r.reshape(-1) for r in regions])[:, None, None]
    weight, bias = fit_readout(feat, region, 2, samples=400000)
    np.savez(bench.READOUT, weight=weight[:, :, 1, 1, 1].numpy(), bias=bias.numpy(),
             note="fitted by oracle/make_readout.py on patches %s of bench.synthetic_volume(0)" % picks.tolist())
    print("wrote", bench.READOUT, os.path.getsize(bench.READOUT), "bytes")


if __name__ == "__main__":
    main()
