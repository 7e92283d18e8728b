"""Drop-in proof on the GPU: the unmodified reference ``SegmentationTrainer`` runs one iteration whose training step
(segmentation_trainer.py:162-180) goes through the b200 model in training mode (device forward / backward) and whose
validation branch (segmentation_trainer.py:196-242) goes through the b200 ``PatchPredict`` -> history-inverse
``add_evaluation_labels`` (label swap undone, argmax on the device) -> ``SegmentationEvaluator`` (device confusion
histogram); the per-subject TP / FP / TN / FN / Dice the trainer logged are compared with the CPU oracle."""
import pytest

from test_dropin import run_driver

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_reference_trainer_validation_branch_on_gpu(precision):
    res = run_driver("cuda:0", precision)
    assert "error" not in res, res.get("error")
    assert res["served_from"]["segmentation_pipeline.segmentation_trainer"] == "reference"
    assert res["served_from"]["segmentation_pipeline.prediction"] == "b200"
    stats = res["subject_stats"]
    cols = stats["columns"]
    rows = dict(zip(stats["index"], stats["data"]))           # plain RangeIndex; 'subject' and 'label' are columns
    # counts may differ from the oracle by voxels whose top-2 probabilities tie within the path's tolerance
    slack = 4 if precision == "fp32" else 80
    checked = 0
    for key, row in rows.items():
        subject, label = row[cols.index("subject")], row[cols.index("label")]
        want = res["oracle"][subject][label]
        for stat in ("TP", "FP", "TN", "FN"):
            assert abs(row[cols.index(stat)] - want[stat]) <= slack, (subject, label, stat)
        assert abs(row[cols.index("dice")] - want["dice"]) <= (1e-3 if precision == "fp32" else 2e-2)
        checked += 1
    assert checked == 4                                   # 2 validation subjects x 2 labels
    import json
    print("DROPIN-GPU " + json.dumps({"precision": precision, "torchio": res["torchio"],
                                      "served_from": res["served_from"], "subject_stats": stats["data"],
                                      "oracle": res["oracle"], "model_score": res["model_score"]}))
    assert "model_forward_evaluation" in res["timer_keys"] and "evaluation.seg.validation" in res["timer_keys"]
    # the training step of the same iteration (segmentation_trainer.py:162-180) ran on the device too: the model in
    # training mode behind StandardPredict(['X', 'y']), the b200 criterion, backward through the device autograd
    # function, the reference's optimizer step -- loss and updated weights (incl. BatchNorm running statistics) match an
    # autograd step over the CPU oracle
    assert res["criterion_module"].endswith("from b200")
    assert abs(res["train_loss"]["loss"] - res["oracle_train_loss"]) <= 1e-5 * max(1.0, abs(res["oracle_train_loss"]))
    assert res["train_grad_abs_sum"] is not None and res["train_grad_abs_sum"] > 0
    assert res["weights_moved"] > 1e-6
    assert res["train_step_max_abs_weight_diff"] <= 2e-5      # lr x gradient tolerance (see assert_gradients_match)
