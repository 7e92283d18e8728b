"""CPU check of the label-agreement harness (tests/helpers.py) with the oracle's bf16 EMULATION standing in for the
CUDA path: activations and weights rounded to bf16 exactly where the kernels store them (oracle/unet.py
``emulate_bf16``).  It predicts, without a GPU, what tests/test_gpu_labels.py measures on the B200: with a fitted,
confident head the unfiltered agreement is >= 99.9 %; the emulated and the exact forward differ by the bf16 tolerance."""
import torch

from oracle import unet
from helpers import blob_volume, fit_readout, label_agreement_report, rel_err


def _nested_sd(filters, classes, seed):
    from oracle.ref_shapes import nested_state_dict
    return nested_state_dict(2, classes, filters, seed)


def test_emulated_bf16_label_agreement_nested():
    torch.set_num_threads(max(torch.get_num_threads(), 8))
    sd = _nested_sd(24, 6, 3)
    vol, region = blob_volume(2, (64, 64, 64), 5, seed=5, distinct=True)
    x = vol[None]
    with torch.no_grad():
        feat = unet.nested_res_unet_forward(sd, x, {"return_features": True})[0]
    sd["out_conv.weight"], sd["out_conv.bias"] = fit_readout(feat, region, 6, samples=100000)
    with torch.no_grad():
        ref_logits = unet.nested_res_unet_forward(sd, x, {"hypothesis": "identity"})
        with unet.emulate_bf16():
            emu_logits = unet.nested_res_unet_forward(sd, x, {"hypothesis": "identity"})
    assert 1e-6 < rel_err(emu_logits, ref_logits) <= 2e-2       # the tolerance is stated on logits
    ref, emu = torch.softmax(ref_logits, 1), torch.softmax(emu_logits, 1)
    rep = label_agreement_report(ref, emu, "emulated bf16, NestedResUNet(2,6,24) 64^3")
    assert rep["median_margin"] >= 0.5
    assert rep["agreement"] >= 0.999
