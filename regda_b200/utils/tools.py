"""Host-side glue of the training loop -- drop-in for the on-path pieces of regda/utils/tools.py:
loss_calc (:240-260), adjust_learning_rate / lr_poly / lr_warmup (:191-207), seed_torch (:305-314),
import_config (:173-181)."""
from __future__ import annotations

import importlib
import os
import random

import numpy as np
import torch
import torch.nn.functional as tnf

from ..gast.balance import CrossEntropy


def loss_calc(pred, label, loss_fn, multi=False):
    """tools.py:240-260.  With the fused CrossEntropy the low-resolution logits go straight to the
    kernel (it upsamples internally); any other loss_fn gets the reference's explicit interpolate."""
    fused = isinstance(loss_fn, CrossEntropy)

    def one(p):
        if not fused and p.size()[-2:] != label.size()[-2:]:
            p = tnf.interpolate(p, size=label.size()[-2:], mode='bilinear', align_corners=True)
        return loss_fn(p, label.long())

    if multi is True:
        loss = 0
        num = 0
        for p in pred:
            loss = loss + one(p)
            num += 1
        return loss / num
    return one(pred)


def lr_poly(base_lr, i_iter, max_iter, power):
    return base_lr * ((1 - float(i_iter) / max_iter) ** power)


def lr_warmup(base_lr, i_iter, warmup_iter):
    return base_lr * (float(i_iter) / warmup_iter)


def adjust_learning_rate(optimizer, i_iter, cfg):
    if i_iter < cfg.PREHEAT_STEPS:
        lr = lr_warmup(cfg.LEARNING_RATE, i_iter, cfg.PREHEAT_STEPS)
    else:
        lr = lr_poly(cfg.LEARNING_RATE, i_iter, cfg.NUM_STEPS, cfg.POWER)
    optimizer.param_groups[0]['lr'] = lr
    if len(optimizer.param_groups) > 1:
        optimizer.param_groups[1]['lr'] = lr * 10
    return lr


def seed_torch(seed=2333):
    random.seed(seed)
    os.environ['PYTHONHASHSEED'] = str(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)
    torch.cuda.manual_seed(seed)
    torch.cuda.manual_seed_all(seed)


def import_config(config_name, prefix='configs', copy=False, create=False):
    cfg_path = '{}.{}'.format(prefix, config_name)
    m = importlib.import_module(name=cfg_path)
    if create:
        os.makedirs(m.SNAPSHOT_DIR, exist_ok=True)
    return m


# ---- offline teacher pass: tools.py:55-152 -------------------------------------------------------
def pad_image(img, target_size):
    """tools.py:52-57, verbatim semantics: F.pad's tuple is (left, right, top, bottom), so the reference pads the HEIGHT by
    rows_missing at the top and cols_missing at the bottom and never the width.  It only matters for images smaller than the
    tile (windows of larger images are shifted back inside and need no padding); kept as is for drop-in parity."""
    rows_missing = target_size[0] - img.shape[2]
    cols_missing = target_size[1] - img.shape[3]
    return tnf.pad(img, (0, 0, rows_missing, cols_missing), 'constant', 0)


# the 8 views of ttach.Compose([HorizontalFlip(), Rotate90([0, 90, 180, 270])]) (ttach==0.0.3, requirement.txt:165), in its
# itertools.product order: augment = flip (if set) then rot90(k); de-augment = rot90(-k) then flip
_TTA_VIEWS = [(flip, k) for flip in (False, True) for k in (0, 1, 2, 3)]


def _tta_augment(img, flip, k):
    if flip:
        img = img.flip(3)
    return torch.rot90(img, k, (2, 3)) if k else img


def _tta_deaugment(x, flip, k):
    if k:
        x = torch.rot90(x, -k, (2, 3))
    return x.flip(3) if flip else x


def tta_predict(model, img):
    """tools.py:132-152: mean of the de-augmented eval-mode predictions over the 8 flip / rot90 views.  The reference calls
    the model once per view; square inputs (pre_slide pads every tile to tile_size) go through ONE forward of the 8 views
    stacked as a batch instead -- eval-mode BatchNorm and InstanceNorm are per-sample, so the result is the same -- and the
    mean over ALL views and batch entries keeps the reference's `torch.mean(xs, dim=0, keepdim=True)` (b = 1 in the tools)."""
    b, _, h, w = img.shape
    if h == w:
        views = torch.cat([_tta_augment(img, f, k) for f, k in _TTA_VIEWS], 0)
        out = model(views)
        xs = torch.cat([_tta_deaugment(out[i * b:(i + 1) * b], f, k) for i, (f, k) in enumerate(_TTA_VIEWS)], 0)
    else:
        xs = torch.cat([_tta_deaugment(model(_tta_augment(img, f, k)), f, k) for f, k in _TTA_VIEWS], 0)
    return torch.mean(xs, dim=0, keepdim=True)


def pre_slide(model, image, num_classes=7, tile_size=(512, 512), tta=False):
    """tools.py:61-97: averaged predictions over 50 %-overlap sliding windows (the last window of a row / column is shifted
    back inside the image; windows smaller than tile_size are zero-padded)."""
    from math import ceil
    image_size = image.shape
    overlap = 1 / 2
    stride = ceil(tile_size[0] * (1 - overlap))
    tile_rows = int(ceil((image_size[2] - tile_size[0]) / stride) + 1)
    tile_cols = int(ceil((image_size[3] - tile_size[1]) / stride) + 1)
    full_probs = torch.zeros((image_size[0], num_classes, image_size[2], image_size[3]), device=image.device)
    count_predictions = torch.zeros((image_size[0], 1, image_size[2], image_size[3]), device=image.device)
    for row in range(tile_rows):
        for col in range(tile_cols):
            x1 = int(col * stride)
            y1 = int(row * stride)
            x2 = min(x1 + tile_size[1], image_size[3])
            y2 = min(y1 + tile_size[0], image_size[2])
            x1 = max(int(x2 - tile_size[1]), 0)
            y1 = max(int(y2 - tile_size[0]), 0)
            img = image[:, :, y1:y2, x1:x2]
            padded_img = pad_image(img, tile_size)
            padded = tta_predict(model, padded_img) if tta is True else model(padded_img)
            pre = padded[:, :, 0:img.shape[2], 0:img.shape[3]]
            count_predictions[:, :, y1:y2, x1:x2] += 1
            full_probs[:, :, y1:y2, x1:x2] += pre
    full_probs /= count_predictions
    return full_probs
