"""TEST DRIVER (run in its own process by tests/test_dropin*.py) -- the UNMODIFIED reference ``SegmentationTrainer``
(/root/reference/segmentation_pipeline/segmentation_trainer.py:98-280) running one iteration against the b200 hot path:

  * ``b200_overlay.install(reference_root)`` -> ``import segmentation_pipeline`` executes the reference's own
    ``__init__.py``; ``prediction`` / ``models`` / the two count evaluators resolve to this repo, everything else
    (trainer, SubjectFolder, filters, transforms, data loaders, TorchContext, loggers) is the reference's file;
  * the dataset is the reference's ``SubjectFolder`` over a temporary folder tree, fed by a synthetic in-memory
    ``SubjectLoader``; the transform pipeline is ``CustomRemapLabels`` (a label swap) + ``CustomOneHot`` -- so that
    ``add_evaluation_labels`` has a real history to invert (reference prediction.py:155-170);
  * on the GPU the TRAINING step (segmentation_trainer.py:162-180) is the real thing too: ``StandardPredict(['X', 'y'])``
    over the b200 model in training mode, the b200 criterion, backward through the device autograd function, the
    reference's SGD step (on a CPU-only box a stub train predictor / criterion / optimizer stand in, the device path
    refuses CPU tensors); then the VALIDATION branch (:196-242): ``PatchPredict.predict`` -> ``add_evaluation_labels`` ->
    ``SegmentationEvaluator`` through the scheduled evaluation.

Prints one JSON line: which files served the key modules, and either the evaluation results (+ the oracle comparison)
or the error raised on the way (on a box without a GPU the predictor must refuse, loudly)."""
import argparse
import json
import os
import sys
import tempfile
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "segmentation-pipeline_b200"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reference", required=True)
    ap.add_argument("--device", default="cuda:0")
    ap.add_argument("--precision", default="fp32")
    args = ap.parse_args()

    shimmed = False
    try:
        import torchio  # noqa: F401
    except ImportError:
        sys.path.insert(0, os.path.join(ROOT, "tests", "shims"))
        shimmed = True
    sys.path.insert(0, os.path.join(ROOT, "tests", "shims")) if os.path.join(ROOT, "tests", "shims") not in sys.path else None
    import optional_stubs
    stubbed = optional_stubs.install()

    import numpy as np
    import torch
    import torchio as tio

    import b200_overlay
    finder = b200_overlay.install(args.reference)
    import segmentation_pipeline as sp
    from segmentation_pipeline.models import _engine

    out = {"torchio": tio.__version__, "torchio_shim": shimmed, "stubbed_optional_packages": stubbed,
           "served_from": {k: ("b200" if "segmentation-pipeline_b200" in v else "reference")
                           for k, v in sorted(finder.resolved.items())
                           if k.split(".")[-1] in ("segmentation_pipeline", "segmentation_trainer", "prediction",
                                                   "modular_unet", "segmentation_evaluator", "custom_label_transforms",
                                                   "subject_folder", "torch_context", "data_loader_factory", "post_processing")}}

    # ------------------------------------------------------------------ synthetic dataset behind the reference's SubjectFolder
    tmp = tempfile.mkdtemp(prefix="b200_dropin_")
    names = ["s0", "s1", "s2", "s3"]
    for n in names:
        os.makedirs(os.path.join(tmp, "subjects", n))
    shape = (40, 36, 28)

    def synth(name):
        g = torch.Generator().manual_seed(100 + int(name[1:]))
        coarse = torch.randn(1, 2, 10, 9, 7, generator=g)
        vol = torch.nn.functional.interpolate(coarse, size=shape, mode="trilinear", align_corners=False)[0]
        vol = (vol + 0.1 * torch.randn(2, *shape, generator=g)).contiguous()
        labels = (vol[0] > 0.3).long() + (vol[1] > 0.5).long()          # values 0, 1, 2
        return vol, labels[None]

    class SyntheticLoader(sp.data_processing.subject_loaders.SubjectLoader):
        def __call__(self, subject_data):
            vol, labels = synth(subject_data["name"])
            subject_data["X"] = tio.ScalarImage(tensor=vol)
            subject_data["y"] = tio.LabelMap(tensor=labels, label_values={"core": 1, "rim": 2})

    transforms = tio.Compose([
        sp.CustomRemapLabels(remapping={1: 2, 2: 1}, include=["y"]),          # swap the two labels (invertible)
        sp.CustomOneHot(num_classes=3, include=["y"]),
    ])
    cohorts = {"training": sp.RequireAttributes({"name": ["s0", "s1"]}),
               "validation": sp.RequireAttributes({"name": ["s2", "s3"]})}
    dataset = sp.SubjectFolder(tmp, "subjects", SyntheticLoader(), cohorts=cohorts, transforms=transforms)

    # ------------------------------------------------------------------ model (b200 classes through the reference namespace)
    torch.manual_seed(0)
    model = sp.ModularUNet(2, 3, [8, 16], 2, block_params={"residual": True})
    model.eval()
    sd = {k: v.clone() for k, v in model.state_dict().items()}
    sd0 = {k: v.clone() for k, v in sd.items()}
    device = torch.device(args.device)

    class StubTrainPredictor(sp.prediction.Predictor):
        """CPU-only wiring run: stands in for the training-mode forward (which needs the GPU): returns tensors the
        criterion can consume.  On the GPU the trainer gets the real thing -- StandardPredict(['X', 'y']) over the model
        in training mode, as research/msseg2/msseg2.py:138 configures it."""
        def predict(self, model, device, subjects, label_attributes=None):
            y = torch.stack([s["y"]["data"] for s in subjects]).float().to(device)
            return subjects, {"y": y, "y_pred": (y * 0 + 1.0 / y.shape[1]).requires_grad_(True)}

    grads = []

    def criterion(y_pred, y):
        """On the GPU: the b200 HybridLogisticDiceLoss (fused device kernels, autograd) through the reference namespace;
        on a CPU-only box the product criterion refuses CPU tensors, so the wiring test uses a trivial stand-in."""
        if device.type != "cuda":
            return {"loss": (y_pred * y).mean()}
        y_pred.register_hook(lambda g: grads.append(float(g.abs().sum())))
        return sp.HybridLogisticDiceLoss()(y_pred, y)

    logs = []

    class CaptureLogger(sp.Logger):
        def __init__(self):
            pass

        def setup(self, context):
            pass

        def save_context(self, context, save_path, iteration):
            pass

        def log(self, log_dict):
            logs.append(log_dict)

    dummy = torch.nn.Parameter(torch.zeros(1))
    on_gpu = device.type == "cuda"
    if on_gpu:
        model.to(device)
    context = types.SimpleNamespace(
        dataset=dataset, model=model, device=device,
        criterion=criterion,
        optimizer=torch.optim.SGD(model.parameters(), lr=1e-3, momentum=0.95) if on_gpu
        else torch.optim.SGD([dummy], lr=0.1))
    sampler = torch.utils.data.SequentialSampler
    evaluator = sp.SegmentationEvaluator("y_pred_eval", "y_eval", stats_to_output=("TP", "FP", "TN", "FN", "dice"))
    trainer = sp.SegmentationTrainer(
        training_batch_size=1, save_rate=1000, scoring_interval=1,
        scoring_function=lambda log: float(log["seg"]["validation"]["summary_stats"]["mean", "core", "dice"]),
        one_time_evaluators=[], training_evaluators=[],
        validation_evaluators=[sp.ScheduledEvaluation(evaluator, "seg", cohorts=["validation"])],
        max_iterations_with_no_improvement=10,
        train_predictor=sp.StandardPredict(image_names=["X", "y"]) if on_gpu else StubTrainPredictor(),
        validation_predictor=sp.PatchPredict(patch_batch_size=5, patch_size=(16, 16, 16), patch_overlap=(8, 8, 4),
                                             padding_mode="edge", overlap_mode="average"),
        train_dataloader_factory=sp.StandardDataLoader(sampler=sampler),
        validation_dataloader_factory=sp.StandardDataLoader(sampler=sampler))
    _engine.set_precision(args.precision)
    try:
        trainer.train(context, max_iterations=1, logger=CaptureLogger())
    except Exception as exc:  # noqa: BLE001
        import traceback
        frames = traceback.extract_tb(exc.__traceback__)
        out["error"] = {"type": type(exc).__name__, "message": str(exc)[:300],
                        "raised_in": [f"{os.path.basename(f.filename)}:{f.name}" for f in frames][-6:]}
        print("DROPIN " + json.dumps(out))
        return
    log = logs[-1]
    stats = log["seg"]["validation"]["subject_stats"]
    out["subject_stats"] = json.loads(stats.to_json(orient="split"))
    out["model_score"] = log["model_score"]
    out["timer_keys"] = sorted(log["timer"].keys())
    out["train_loss"] = {k: float(log[k]) for k in ("loss", "dice_loss", "logistic_loss") if k in log}
    out["train_grad_abs_sum"] = grads[-1] if grads else None
    out["criterion_module"] = sp.HybridLogisticDiceLoss.__module__ + " from " + \
        ("b200" if "segmentation-pipeline_b200" in finder.resolved.get("segmentation_pipeline.criterions.hybrid_logistic_dice_loss", "") else "reference")

    # ------------------------------------------------------------------ the same thing on the CPU oracle
    from oracle import evalstats, grid as ogrid, unet
    cfg = {"depth": 2, "filters": [8, 16], "block": {"residual": True}, "down": "avgpool", "up": "trilinear"}
    # the training step of the iteration (segmentation_trainer.py:162-180) on the first training subject: autograd
    # through the oracle with batch-statistic BatchNorm, then the first SGD step (momentum buffer = gradient)
    ref_sd = {k: v.clone().requires_grad_(v.is_floating_point() and "running" not in k) for k, v in sd.items()}
    x0, labels0 = synth("s0")
    remapped = torch.where(labels0 == 1, 2, torch.where(labels0 == 2, 1, labels0))
    y0 = torch.nn.functional.one_hot(remapped[0], 3).movedim(-1, 0)[None].float()
    train_cfg = dict(cfg, block=dict(cfg["block"], bn_training=True))
    ref_loss = unet.hybrid_logistic_dice_loss(unet.modular_unet_forward(ref_sd, x0[None], train_cfg), y0)["loss"]
    ref_loss.backward()
    with torch.no_grad():
        for v in ref_sd.values():
            if v.grad is not None:
                v -= 1e-3 * v.grad
    sd = {k: v.detach() for k, v in ref_sd.items()}
    out["oracle_train_loss"] = float(ref_loss)
    out["train_step_max_abs_weight_diff"] = max(
        float((v.detach().cpu().float() - sd[k].float()).abs().max()) for k, v in model.state_dict().items()
        if "num_batches_tracked" not in k)
    out["weights_moved"] = max(float((v.detach().cpu() - sd0[k]).abs().max()) for k, v in model.state_dict().items()
                               if v.is_floating_point())
    expected = {}
    for n in ("s2", "s3"):
        vol, labels = synth(n)
        probs = ogrid.sliding_window(vol.numpy(),
                                     lambda p: unet.modular_unet_forward(sd, torch.from_numpy(p), cfg).numpy(),
                                     (16, 16, 16), (8, 8, 4), "edge", "average", patch_batch_size=5)
        pred = evalstats.argmax_labels(probs)                      # in the one-hot (swapped) label space ...
        pred_orig = np.where(pred == 1, 2, np.where(pred == 2, 1, pred))    # ... the history inverse swaps back
        # y_eval must be the ORIGINAL label map; label_values were remapped by CustomRemapLabels only for
        # Sequence-style remappings, so the names keep their original ids
        expected[n] = evalstats.segmentation_stats(pred_orig, labels.numpy(), {"core": 1, "rim": 2},
                                                   ("TP", "FP", "TN", "FN", "dice"))
    out["oracle"] = expected
    print("DROPIN " + json.dumps(out))


if __name__ == "__main__":
    main()
