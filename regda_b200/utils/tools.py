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
