"""CPU: the PixelMetricIgnore mirror (regda_b200/gast/metrics.py) against tests/golden/miou.npz, produced by the reference's own
PixelMetricIgnore.summary_all (regda/gast/metrics.py:26-65; make_golden.py miou_metric): 5-decimal rounding of the per-class
figures, ignored-class pop, rounded means, nan for a class that never occurs."""
import numpy as np
import torch

from conftest import load_golden
from regda_b200.gast.metrics import PixelMetricIgnore


def _cases():
    z = load_golden("miou.npz")
    for k in sorted({f.split("/")[0] for f in z.files}):
        yield {f.split("/")[1]: z[f] for f in z.files if f.startswith(k + "/")}


def test_summary_all_matches_reference():
    n = 0
    for c in _cases():
        C = int(c["num_classes"])
        op = PixelMetricIgnore(C, class_names=[f"c{i}" for i in range(C)], ignore_labels=[int(v) for v in c["ignore"]])
        gt, pred = torch.from_numpy(c["gt"]), torch.from_numpy(c["pred"])
        for i in range(gt.shape[0]):
            m = gt[i] >= 0
            op.forward(gt[i][m], pred[i][m])
        tb, miou = op.summary_all()
        rows = np.array([[float(v) for v in r[2:]] for r in tb.rows[:-1]])
        # the reference's table holds float32 values rounded to 5 decimals: equal as decimals
        np.testing.assert_allclose(rows, c["rows"], rtol=0, atol=1e-7, equal_nan=True)
        np.testing.assert_allclose(np.array([float(v) for v in tb.rows[-1][2:]]), c["means"], rtol=0, atol=1e-7, equal_nan=True)
        np.testing.assert_allclose(float(miou), float(c["miou"]), rtol=0, atol=1e-7, equal_nan=True)
        assert len(tb.rows) - 1 == C - len(c["ignore"])
        n += 1
    assert n == 3


def test_forward_rejects_out_of_range():
    op = PixelMetricIgnore(4)
    try:
        op.forward(torch.tensor([0, 1, 4]), torch.tensor([0, 1, 2]))
    except (ValueError, RuntimeError):
        return
    raise AssertionError("label == num_classes must raise (the reference's sparse matrix constructor does)")


def test_table_prints():
    op = PixelMetricIgnore(3, class_names=["a", "b", "c"], ignore_labels=[0])
    op.forward(torch.tensor([0, 1, 2, 2]), torch.tensor([0, 1, 2, 1]))
    tb, miou = op.summary_all()
    s = str(tb)
    assert "iou" in s and "mean" in s and abs(miou - 0.5) < 1e-9
