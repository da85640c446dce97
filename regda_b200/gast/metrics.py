"""PixelMetricIgnore -- host mirror of regda/gast/metrics.py:19-65, the metric regda/utils/eval.py:35-37,49,56 reports.

The reference subclasses `ever.api.metric.pixel.PixelMetric` (ever-beta==0.2.3, requirement.txt:33; absent from this image).
Its published algorithm, restated: a float32 [C,C] confusion matrix (rows = ground truth, columns = prediction) accumulated
from (y_true, y_pred) pairs; per class IoU = diag / (row + col - diag), precision = diag / col, recall = diag / row,
F_beta = (1+b^2) P R / (b^2 P + R) -- all in float32, 0/0 = nan.  `summary_all` (metrics.py:26-65) rounds every per-class
figure to 5 decimals, pops the ignored class ids (IsprsDA: class 0, eval.py:16-17), and returns (table, round(mean IoU, 5)).

The confusion matrix is accumulated on the device (one bincount per batch, int64) and read back once in summary_all."""
from __future__ import annotations

import numpy as np
import torch


class Table:
    """what `summary_all` returns first (the reference returns a prettytable.PrettyTable; only str() of it is ever used)"""

    def __init__(self, field_names):
        self.field_names = list(field_names)
        self.rows = []

    def add_row(self, row):
        self.rows.append(list(row))

    def __str__(self):
        cells = [self.field_names] + [[str(c) for c in r] for r in self.rows]
        widths = [max(len(str(r[i])) for r in cells) for i in range(len(self.field_names))]
        line = "+" + "+".join("-" * (w + 2) for w in widths) + "+"
        out = [line, "| " + " | ".join(str(c).center(w) for c, w in zip(cells[0], widths)) + " |", line]
        out += ["| " + " | ".join(str(c).center(w) for c, w in zip(r, widths)) + " |" for r in cells[1:]]
        return "\n".join(out + [line])


def _per_class(cm32):
    """float32 per-class figures of ever's PixelMetric.compute_{iou,F_measure,precision,recall}_per_class"""
    with np.errstate(divide="ignore", invalid="ignore"):
        over_pred = np.sum(cm32, axis=0)          # column sums: predicted as c
        over_true = np.sum(cm32, axis=1)          # row sums: labelled c
        diag = np.diag(cm32)
        iou = diag / (over_pred + over_true - diag)
        precision = diag / over_pred
        recall = diag / over_true
        f1 = (1 + 1.0 ** 2) * precision * recall / ((1.0 ** 2) * precision + recall)
    return iou, f1, precision, recall


class PixelMetricIgnore:
    def __init__(self, num_classes, logdir=None, logger=None, class_names=None, ignore_labels=None, device=None):
        self.num_classes = num_classes
        self.logger = logger
        self.logdir = logdir
        self._class_names = list(class_names) if class_names else None
        self.ignore_labels = sorted(ignore_labels or [], reverse=True)           # metrics.py:23-24
        self._total = torch.zeros((num_classes, num_classes), dtype=torch.int64, device=device)

    def reset(self):
        self._total.zero_()

    def forward(self, y_true, y_pred):
        """accumulate; both 1-D integer tensors of valid pixels (the caller masks `cls_gt >= 0`, eval.py:45-49).  A label or
        prediction outside [0, num_classes) raises, as the reference's sparse matrix constructor does."""
        y_true = torch.as_tensor(y_true).reshape(-1).to(self._total.device, torch.int64)
        y_pred = torch.as_tensor(y_pred).reshape(-1).to(self._total.device, torch.int64)
        idx = y_true * self.num_classes + y_pred
        cm = torch.bincount(idx, minlength=self.num_classes * self.num_classes)
        if cm.numel() != self.num_classes * self.num_classes or bool(((y_pred < 0) | (y_pred >= self.num_classes)).any()):
            raise ValueError("label / prediction outside [0, num_classes)")
        cm = cm.view(self.num_classes, self.num_classes)
        self._total += cm
        return cm

    __call__ = forward

    def summary_all(self, dec=5):
        dense_cm = self._total.cpu().numpy().astype(np.float32)                  # the reference accumulates a float32 matrix
        iou, f1, precision, recall = (np.round(v, dec).tolist() for v in _per_class(dense_cm))     # metrics.py:29-32
        names = list(self._class_names) if self._class_names else None
        for idx in self.ignore_labels:                                           # metrics.py:34-40
            for lst in (iou, f1, precision, recall):
                lst.pop(idx)
            if names:
                names.pop(idx)
        mrecall = np.round(np.array(recall).mean(), dec)                         # metrics.py:42-45
        miou = np.round(np.array(iou).mean(), dec)
        mf1 = np.round(np.array(f1).mean(), dec)
        mprec = np.round(np.array(precision).mean(), dec)
        if names:
            tb = Table(['name', 'class', 'iou', 'f1', 'precision', 'recall'])
            for i, row in enumerate(zip(iou, f1, precision, recall)):
                tb.add_row([names[i], i, *row])
            tb.add_row(['', 'mean', miou, mf1, mprec, mrecall])
        else:
            tb = Table(['class', 'iou', 'f1', 'precision', 'recall'])
            for i, row in enumerate(zip(iou, f1, precision, recall)):
                tb.add_row([i, *row])
            tb.add_row(['mean', miou, mf1, mprec, mrecall])
        if self.logger is not None:
            self.logger.info('\n' + str(tb))
        self.iou_per_class = iou
        return tb, miou
