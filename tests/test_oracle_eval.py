"""Evaluator arithmetic oracle (oracle/evalstats.py) against the reference's own SegmentationEvaluator /
LabelMapEvaluator outputs stored in tests/golden/evaluator.npz.  CPU only, bit-exact."""
import numpy as np

from oracle import evalstats
from helpers import GOLDEN


def _load():
    z = np.load(f"{GOLDEN}/evaluator.npz")
    names = [str(s) for s in z["label_names"]]
    vals = [int(v) for v in z["label_vals"]]
    return z, dict(zip(names, vals)), [str(s) for s in z["stats"]]


def test_segmentation_stats_match_reference_bit_exact():
    z, label_values, stats = _load()
    rows = []
    for i in range(3):
        res = evalstats.segmentation_stats(z[f"pred{i}"], z[f"targ{i}"], label_values, stats)
        for name in label_values:
            rows.append([res[name][s] for s in stats])
    got = np.array(rows, dtype=np.float64)
    exp = z["subject_stats_values"]
    assert got.shape == exp.shape
    same = (got == exp) | (np.isnan(got) & np.isnan(exp))
    assert same.all()
    assert np.isnan(exp).any() and np.isinf(exp).any() or np.isnan(exp).any()  # degenerate labels are covered


def test_confusion_matrix_reproduces_counts():
    z, label_values, _ = _load()
    for i in range(3):
        p, t = z[f"pred{i}"], z[f"targ{i}"]
        cm = evalstats.confusion_matrix(p, t, 5)
        assert cm.sum() == p.size
        for v in label_values.values():
            assert evalstats.counts_from_confusion(cm, v, p.size) == evalstats.label_counts(p, t, v)


def test_label_volumes_match_reference():
    z, label_values, _ = _load()
    got = []
    for i in range(3):
        vol = evalstats.label_volumes(z[f"pred{i}"], label_values)
        got += [vol[n] for n in label_values]
    np.testing.assert_array_equal(np.array(got, dtype=np.float64), z["volumes"].reshape(-1))


def test_argmax_ties_lowest_index():
    p = np.zeros((3, 2, 2, 2), dtype=np.float32)
    p[1, 0, 0, 0] = p[2, 0, 0, 0] = 0.5
    lab = evalstats.argmax_labels(p)
    assert lab.shape == (1, 2, 2, 2) and lab.dtype == np.int64
    assert lab[0, 0, 0, 0] == 1 and lab[0, 1, 1, 1] == 0
