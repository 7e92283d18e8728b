"""ORACLE fixture generator (test infrastructure only; runs ONLY where /root/reference exists).

    python oracle/make_golden.py          # rewrites tests/golden/*.npz

Imports the reference's own classes in-process (oracle/ref_import.py), runs them on seeded CPU fp32 inputs and
stores inputs, state_dicts and outputs as small .npz fixtures.  The fixtures travel to the GPU box; the
reference tree does not.  Every case records the reference constructor call it came from in ``meta``.
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np
import torch
from torch import nn

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import ref_import  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def perturb_bn(model: nn.Module, seed: int) -> None:
    """Random-init BatchNorm has mean 0 / var 1 / gamma 1 / beta 0, which would make BN folding a no-op;
    perturb running stats and affine terms so parity tests exercise it."""
    g = torch.Generator().manual_seed(seed)
    for m in model.modules():
        if isinstance(m, nn.BatchNorm3d):
            m.running_mean.copy_(0.1 * torch.randn(m.running_mean.shape, generator=g))
            m.running_var.copy_(0.5 + torch.rand(m.running_var.shape, generator=g))
            m.weight.data.copy_(0.5 + torch.rand(m.weight.shape, generator=g))
            m.bias.data.copy_(0.1 * torch.randn(m.bias.shape, generator=g))


def save_case(name: str, model: nn.Module, x: torch.Tensor, meta: dict, extra=None) -> None:
    model.eval()
    with torch.no_grad():
        y = model(x)
    blob = {"x": x.numpy(), "y": y.numpy(), "meta": np.array(json.dumps(meta))}
    for k, v in model.state_dict().items():
        blob["sd/" + k] = v.numpy()
    if extra:
        blob.update(extra)
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, **blob)
    print(f"{name}: x{tuple(x.shape)} -> y{tuple(y.shape)}  {os.path.getsize(path) / 1024:.0f} KiB")


def main() -> None:
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(4)
    M = ref_import.load_reference_models()

    # 1. msseg2-style ModularUNet (research/msseg2/msseg2.py:84-93), shrunk filters/depth
    torch.manual_seed(0)
    net = M.ModularUNet(in_channels=2, out_channels=2, filters=[8, 8, 16], depth=3, block_params={"residual": True},
                        downsample_class=M.BlurConv3d,
                        downsample_params={"kernel_size": 3, "stride": 2, "padding": 1},
                        upsample_class=M.BlurConvTranspose3d,
                        upsample_params={"kernel_size": 3, "stride": 2, "padding": 1, "output_padding": 0})
    perturb_bn(net, 1)
    x = torch.randn(2, 2, 16, 16, 16, generator=torch.Generator().manual_seed(10))
    save_case("models_modular_blur", net, x,
              {"class": "ModularUNet", "in_channels": 2, "out_channels": 2, "filters": [8, 8, 16], "depth": 3,
               "block": {"residual": True, "norm": "batch", "act": "relu"}, "down": "blur", "up": "blur",
               "hypothesis": "softmax"})

    # 2. default ModularUNet (AvgPool3d + trilinear Upsample), config-1 style, non-cubic input
    torch.manual_seed(1)
    net = M.ModularUNet(1, 2, [8, 16, 24], 3)
    perturb_bn(net, 2)
    x = torch.randn(1, 1, 16, 8, 24, generator=torch.Generator().manual_seed(11))
    save_case("models_modular_default", net, x,
              {"class": "ModularUNet", "in_channels": 1, "out_channels": 2, "filters": [8, 16, 24], "depth": 3,
               "block": {"residual": False, "norm": "batch", "act": "relu"}, "down": "avgpool", "up": "trilinear",
               "hypothesis": "softmax"})

    # 3. LeakyReLU + no norm + Identity hypothesis (logits), int filters broadcast, 3 classes
    torch.manual_seed(2)
    net = M.ModularUNet(2, 3, 8, 2,
                        block_params={"residual": True, "normalization_class": None,
                                      "activation_class": nn.LeakyReLU,
                                      "activation_params": {"negative_slope": 0.1}},
                        hypothesis_class=nn.Identity, hypothesis_params={})
    x = torch.randn(1, 2, 8, 8, 8, generator=torch.Generator().manual_seed(12))
    save_case("models_modular_leaky_logits", net, x,
              {"class": "ModularUNet", "in_channels": 2, "out_channels": 3, "filters": [8, 8], "depth": 2,
               "block": {"residual": True, "norm": "none", "act": "leaky_relu", "slope": 0.1},
               "down": "avgpool", "up": "trilinear", "hypothesis": "identity"})

    # 4. WSConv3d blocks + InstanceNorm3d
    torch.manual_seed(3)
    net = M.ModularUNet(1, 2, [8, 8], 2,
                        block_params={"conv_class": M.WSConv3d, "conv_params": {"kernel_size": 3, "padding": 1},
                                      "normalization_class": nn.InstanceNorm3d})
    x = torch.randn(2, 1, 8, 8, 8, generator=torch.Generator().manual_seed(13))
    save_case("models_modular_ws_instnorm", net, x,
              {"class": "ModularUNet", "in_channels": 1, "out_channels": 2, "filters": [8, 8], "depth": 2,
               "block": {"residual": False, "norm": "instance", "act": "relu", "conv": "ws"},
               "down": "avgpool", "up": "trilinear", "hypothesis": "softmax"})

    # 5. NestedResUNet (research/dmri_hippo/configs/main_config.py:123-127), shrunk
    torch.manual_seed(4)
    net = M.NestedResUNet(3, 2, 8, dropout_p=0.2)
    perturb_bn(net, 3)
    x = torch.randn(1, 3, 16, 16, 8, generator=torch.Generator().manual_seed(14))
    save_case("models_nested", net, x, {"class": "NestedResUNet", "input_channels": 3, "output_channels": 2,
                                        "filters": 8, "dropout_p": 0.2, "hypothesis": "softmax"})

    # 6. 10-class NestedResUNet (qsm shape)
    torch.manual_seed(5)
    net = M.NestedResUNet(2, 10, 8)
    perturb_bn(net, 4)
    x = torch.randn(1, 2, 8, 8, 8, generator=torch.Generator().manual_seed(15))
    save_case("models_nested_10class", net, x, {"class": "NestedResUNet", "input_channels": 2,
                                                "output_channels": 10, "filters": 8, "hypothesis": "softmax"})

    # 7. components in isolation
    torch.manual_seed(6)
    comp = {}
    bc = M.BlurConv3d(8, 8, kernel_size=3, stride=2, padding=1)
    xb = torch.randn(1, 8, 8, 8, 8, generator=torch.Generator().manual_seed(16))
    bt = M.BlurConvTranspose3d(8, 8, kernel_size=3, stride=2, padding=1, output_padding=0)
    sm = M.StochasticMatrix(2, diag_bias=1.5)
    xs = torch.randn(1, 4, 4, 4, 4, generator=torch.Generator().manual_seed(17))
    with torch.no_grad():
        comp.update({"blur_x": xb.numpy(), "blur_w": bc.weight.numpy(), "blur_kernel": bc.kernel.numpy(),
                     "blur_y": bc(xb).numpy(), "blurT_w": bt.weight.numpy(), "blurT_kernel": bt.kernel.numpy(),
                     "blurT_y": bt(xb).numpy(), "sm_x": xs.numpy(), "sm_y": sm(xs).numpy()})
    # 8. ensembles on the small blur model (EnsembleFlips mean / majority, EnsembleOrientations mean)
    torch.manual_seed(7)
    base = M.ModularUNet(1, 2, [8, 8], 2)
    perturb_bn(base, 5)
    base.eval()
    xe = torch.randn(1, 1, 8, 8, 8, generator=torch.Generator().manual_seed(18))
    with torch.no_grad():
        comp["ens_x"] = xe.numpy()
        comp["ens_flips_mean"] = M.EnsembleFlips(base, "mean")(xe).numpy()
        comp["ens_flips_majority"] = M.EnsembleFlips(base, "majority")(xe).numpy()
        comp["ens_orient_mean"] = M.EnsembleOrientations(base, "mean")(xe).numpy()
    for k, v in base.state_dict().items():
        comp["ens_sd/" + k] = v.numpy()
    # 9. criterion
    Loss = ref_import.load_reference_criterion()
    g = torch.Generator().manual_seed(19)
    pr = torch.softmax(torch.randn(2, 2, 6, 6, 6, generator=g), dim=1)
    tg = torch.nn.functional.one_hot(torch.randint(0, 2, (2, 6, 6, 6), generator=g), 2).movedim(-1, 1).float()
    out = Loss(logistic_class_weights=[1, 100])(pr, tg)
    comp.update({"loss_pred": pr.numpy(), "loss_target": tg.numpy(),
                 "loss_values": np.array([out["loss"].item(), out["dice_loss"].item(), out["logistic_loss"].item()],
                                         dtype=np.float64)})
    path = os.path.join(OUT, "components.npz")
    np.savez_compressed(path, **comp)
    print(f"components: {os.path.getsize(path) / 1024:.0f} KiB")

    # 10. evaluator arithmetic (segmentation_evaluator.py:56-102, label_map_evaluator.py:66-109)
    SegEval, LabEval, _ = ref_import.load_reference_evaluators()
    rng = np.random.default_rng(20)
    label_values = {"a": 1, "b": 2, "c": 4}

    class Img(dict):
        @property
        def data(self):
            return self["data"]

    subjects = []
    preds, targs = [], []
    for i, shape in enumerate([(1, 12, 10, 9), (1, 7, 7, 7), (1, 5, 6, 3)]):
        p = rng.integers(0, 5, size=shape).astype(np.int64)
        t = rng.integers(0, 5, size=shape).astype(np.int64)
        if i == 2:
            p[p == 4] = 0  # label 'c' predicted nowhere -> 0/0 and x/0 cases
            t[t == 2] = 0
        preds.append(p)
        targs.append(t)
        subjects.append({"name": f"s{i}",
                         "pred": Img(data=torch.from_numpy(p), label_values=label_values),
                         "targ": Img(data=torch.from_numpy(t), label_values=label_values)})
    stats = ("target_volume", "prediction_volume", "TP", "FP", "TN", "FN", "dice", "jaccard", "precision", "recall")
    res = SegEval("pred", "targ", stats_to_output=stats)(subjects)
    vol = LabEval("pred")(subjects)
    ev = {"label_names": np.array(list(label_values.keys())), "label_vals": np.array(list(label_values.values())),
          "stats": np.array(stats)}
    for i in range(3):
        ev[f"pred{i}"] = preds[i]
        ev[f"targ{i}"] = targs[i]
    df = res["subject_stats"]
    ev["subject_stats_columns"] = np.array(list(df.columns))
    ev["subject_stats_values"] = df[list(stats)].to_numpy(dtype=np.float64)  # rows = (subject, label) pairs
    ev["summary_stats"] = res["summary_stats"].data.numpy()
    ev["volumes"] = vol["subject_stats"][["volume"]].to_numpy(dtype=np.float64)
    path = os.path.join(OUT, "evaluator.npz")
    np.savez_compressed(path, **ev)
    print(f"evaluator: {os.path.getsize(path) / 1024:.0f} KiB; columns {list(df.columns)[:6]}...")


if __name__ == "__main__":
    if not ref_import.reference_available():
        sys.exit("reference tree not available: fixtures can only be regenerated in the authoring container")
    main()
