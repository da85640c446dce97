"""evaluate() -- host mirror of regda/utils/eval.py:14-56; the metric is gast/metrics.py's PixelMetricIgnore (mirror of
regda/gast/metrics.py:19-65 over ever's PixelMetric: float32 IoU_c = cm[c,c] / (row_c + col_c - cm[c,c]), 5-decimal
rounding, ignored classes popped).  Forward-only reuse of the model; the confusion matrix is accumulated on the device
(one bincount per batch) and read back once.

`loader` yields (image float [b,3,H,W], label int64 [b,H,W]) -- the reference's DALoader(EVAL_DATA_CONFIG) does when its
`cls` entry is unpacked; images larger than `tile` are evaluated by 50 %-overlap sliding windows as pre_slide
(regda/utils/tools.py:61-97) does."""
from __future__ import annotations

import torch


def _origins(size, tile, stride):
    """window origins along one axis (tools.py:66-79): 0, stride, ... with the last window shifted back inside the image;
    a single origin 0 when the image is not larger than the tile"""
    if size <= tile:
        return [0]
    out = list(range(0, size - tile + 1, stride))
    if out[-1] != size - tile:
        out.append(size - tile)
    return out


def slide_predict(model, image, num_classes, tile=512, stride=None):
    """averaged eval-mode probabilities over 50 %-overlap tile x tile windows (tools.py:61-97: stride = ceil(tile / 2)); one
    window if it fits.
    (An axis shorter than the tile is taken whole -- the reference pads such a window to the tile size, pad_image
    tools.py:52-57; use utils.tools.pre_slide for that exact behaviour.)"""
    b, _, H, W = image.shape
    if H <= tile and W <= tile:
        return model(image)
    stride = stride or (tile + 1) // 2
    prob = torch.zeros((b, num_classes, H, W), device=image.device)
    cnt = torch.zeros((1, 1, H, W), device=image.device)
    for y in _origins(H, tile, stride):
        for x in _origins(W, tile, stride):
            prob[:, :, y:y + tile, x:x + tile] += model(image[:, :, y:y + tile, x:x + tile])
            cnt[:, :, y:y + tile, x:x + tile] += 1
    return prob / cnt


def confusion_matrix(pred, label, num_classes, ignore_label=-1):
    keep = (label != ignore_label) & (label >= 0) & (label < num_classes)
    idx = label[keep] * num_classes + pred[keep]
    return torch.bincount(idx, minlength=num_classes * num_classes).view(num_classes, num_classes)


def miou_from_confusion(cm, skip_class0=False):
    """float64 IoU per class / mean without the reference's rounding (diagnostics; evaluate() reports PixelMetricIgnore's)"""
    cm = cm.double()
    inter = cm.diag()
    union = cm.sum(0) + cm.sum(1) - inter
    iou = inter / union.clamp_min(1)
    if skip_class0:                                   # IsprsDA drops class 0 (regda/utils/eval.py:16-17)
        iou = iou[1:]
    return iou, float(iou.mean())


@torch.no_grad()
def evaluate(model, loader, num_classes, ignore_label=-1, skip_class0=False, tile=512, logger=None, tta=False, class_names=None):
    """regda/utils/eval.py:14-56: eval-mode (sliding-window) prediction, arg-max, PixelMetricIgnore over the pixels with
    label >= 0 (:45-49), summary_all() -> (table, mIoU) with the reference's 5-decimal rounding and class-0 pop
    (`skip_class0`: DATASETS == 'IsprsDA', :16-17).  `table.iou_per_class` holds the per-class list."""
    from ..gast.metrics import PixelMetricIgnore
    was_training = model.training
    model.eval()
    dev = next(model.parameters()).device
    metric_op = PixelMetricIgnore(num_classes, class_names=class_names, logger=logger, ignore_labels=[0] if skip_class0 else [],
                                  device=dev)
    for image, label in loader:
        image, label = image.to(dev), label.to(dev)
        if tta:                                       # eval.py:41 with tta=True: 8-view TTA inside every sliding window
            from .tools import pre_slide
            pred = pre_slide(model, image, num_classes=num_classes, tile_size=(tile, tile), tta=True).argmax(dim=1)
        else:
            pred = slide_predict(model, image, num_classes, tile).argmax(dim=1)
        mask = label >= 0                             # eval.py:46 (the ignore label is negative on this path)
        if ignore_label >= 0:
            mask &= label != ignore_label
        metric_op.forward(label[mask], pred[mask])
    tb, miou = metric_op.summary_all()
    tb.iou_per_class = metric_op.iou_per_class
    model.train(was_training)
    return tb, float(miou)
