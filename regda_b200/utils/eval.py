"""evaluate() -- host mirror of regda/utils/eval.py:14-56 + the confusion-matrix mIoU of regda/gast/metrics.py:19-65
(ever's PixelMetric: IoU_c = cm[c,c] / (row_c + col_c - cm[c,c])).  Forward-only reuse of the model; the confusion
matrix is accumulated on the device (one bincount per batch) and read back once.

`loader` yields (image float [b,3,H,W], label int64 [b,H,W]) -- the reference's DALoader(EVAL_DATA_CONFIG) does when its
`cls` entry is unpacked; images larger than `tile` are evaluated by 50 %-overlap sliding windows as pre_slide
(regda/utils/tools.py:61-97) does."""
from __future__ import annotations

import torch


def slide_predict(model, image, num_classes, tile=512, stride=256):
    """averaged eval-mode probabilities over overlapping tile x tile windows (tools.py:61-97); one window if it fits"""
    b, _, H, W = image.shape
    if H <= tile and W <= tile:
        return model(image)
    prob = torch.zeros((b, num_classes, H, W), device=image.device)
    cnt = torch.zeros((1, 1, H, W), device=image.device)
    ys = list(range(0, max(H - tile, 0) + 1, stride))
    xs = list(range(0, max(W - tile, 0) + 1, stride))
    if ys[-1] != H - tile:
        ys.append(H - tile)
    if xs[-1] != W - tile:
        xs.append(W - tile)
    for y in ys:
        for x in xs:
            prob[:, :, y:y + tile, x:x + tile] += model(image[:, :, y:y + tile, x:x + tile])
            cnt[:, :, y:y + tile, x:x + tile] += 1
    return prob / cnt


def confusion_matrix(pred, label, num_classes, ignore_label=-1):
    keep = (label != ignore_label) & (label >= 0) & (label < num_classes)
    idx = label[keep] * num_classes + pred[keep]
    return torch.bincount(idx, minlength=num_classes * num_classes).view(num_classes, num_classes)


def miou_from_confusion(cm, skip_class0=False):
    cm = cm.double()
    inter = cm.diag()
    union = cm.sum(0) + cm.sum(1) - inter
    iou = inter / union.clamp_min(1)
    if skip_class0:                                   # IsprsDA drops class 0 (regda/utils/eval.py:16-17)
        iou = iou[1:]
    return iou, float(iou.mean())


@torch.no_grad()
def evaluate(model, loader, num_classes, ignore_label=-1, skip_class0=False, tile=512, logger=None, tta=False):
    was_training = model.training
    model.eval()
    dev = next(model.parameters()).device
    cm = torch.zeros((num_classes, num_classes), dtype=torch.int64, device=dev)
    for image, label in loader:
        image, label = image.to(dev), label.to(dev)
        if tta:                                       # eval.py:41 with tta=True: 8-view TTA inside every sliding window
            from .tools import pre_slide
            pred = pre_slide(model, image, num_classes=num_classes, tile_size=(tile, tile), tta=True).argmax(dim=1)
        else:
            pred = slide_predict(model, image, num_classes, tile).argmax(dim=1)
        cm += confusion_matrix(pred, label, num_classes, ignore_label)
    iou, miou = miou_from_confusion(cm, skip_class0)
    if logger is not None:
        logger.info("IoU per class: " + ", ".join(f"{v:.4f}" for v in iou.tolist()) + f"; mIoU = {miou:.4f}")
    model.train(was_training)
    return iou.cpu(), miou
