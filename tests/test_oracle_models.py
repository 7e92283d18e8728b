"""The oracle's model restatement (oracle/unet.py) against fixtures produced by the reference's own
segmentation_pipeline.models classes (oracle/make_golden.py).  CPU only."""
import numpy as np
import pytest
import torch

from oracle import unet
from helpers import GOLDEN, load_case, rel_err

MODULAR = ["models_modular_blur", "models_modular_default", "models_modular_leaky_logits",
           "models_modular_ws_instnorm"]


@pytest.mark.parametrize("name", MODULAR)
def test_modular_unet_matches_reference(name):
    meta, sd, x, y = load_case(name)
    out = unet.modular_unet_forward(sd, x, meta)
    assert out.shape == y.shape
    assert rel_err(out, y) <= 2e-6


@pytest.mark.parametrize("name", ["models_nested", "models_nested_10class"])
def test_nested_unet_matches_reference(name):
    meta, sd, x, y = load_case(name)
    out = unet.nested_res_unet_forward(sd, x, meta)
    assert rel_err(out, y) <= 2e-6
    assert torch.allclose(out.sum(dim=1), torch.ones_like(out.sum(dim=1)), atol=1e-5)


def test_components_match_reference():
    z = np.load(f"{GOLDEN}/components.npz")
    t = lambda k: torch.from_numpy(z[k])
    y = unet.blur_conv3d(t("blur_x"), t("blur_w"), t("blur_kernel"), 8, stride=2, padding=1)
    assert rel_err(y, t("blur_y")) <= 1e-6
    y = unet.blur_conv_transpose3d(t("blur_x"), t("blurT_w"), t("blurT_kernel"), 8, stride=2, padding=1,
                                   output_padding=0)
    assert rel_err(y, t("blurT_y")) <= 1e-6
    assert y.shape[-1] == 2 * t("blur_x").shape[-1]  # exact 2x upsample (needed by the concat)
    # quirk (components.py:136-141): transposed-blur taps are prod(stride)/(8*out_channels) = 1/out_channels
    assert abs(float(t("blurT_kernel").flatten()[0]) - 1.0 / 8) < 1e-7
    assert abs(float(t("blur_kernel").flatten()[0]) - 1.0 / 64) < 1e-7
    y = unet.stochastic_matrix(t("sm_x"), 2, 1.5)
    assert rel_err(y, t("sm_y")) <= 1e-6


def test_ensembles_match_reference():
    z = np.load(f"{GOLDEN}/components.npz")
    sd = {k[7:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("ens_sd/")}
    cfg = {"depth": 2, "filters": [8, 8], "block": {"residual": False}, "down": "avgpool", "up": "trilinear"}
    fn = lambda x: unet.modular_unet_forward(sd, x, cfg)
    x = torch.from_numpy(z["ens_x"])
    assert rel_err(unet.ensemble_flips(fn, x, "mean"), torch.from_numpy(z["ens_flips_mean"])) <= 2e-6
    maj = unet.ensemble_flips(fn, x, "majority")
    assert (maj.numpy() == z["ens_flips_majority"]).mean() >= 0.999
    assert rel_err(unet.ensemble_orientations(fn, x, "mean"), torch.from_numpy(z["ens_orient_mean"])) <= 2e-6


def test_loss_matches_reference():
    z = np.load(f"{GOLDEN}/components.npz")
    out = unet.hybrid_logistic_dice_loss(torch.from_numpy(z["loss_pred"]), torch.from_numpy(z["loss_target"]),
                                         logistic_class_weights=[1, 100])
    got = np.array([out["loss"].item(), out["dice_loss"].item(), out["logistic_loss"].item()])
    np.testing.assert_allclose(got, z["loss_values"], rtol=1e-6)


def test_split_and_flip_roundtrip():
    x = torch.randn(3, 2, 8, 6, 4)
    assert torch.equal(unet.reverse_split_and_flip(unet.split_and_flip(x)), x)
