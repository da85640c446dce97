"""pseudo_selection -- drop-in for regda/gast/pseudo_generation.py:59-93 (two CUDA kernels:
per-(image, class) maximum, then the thresholded one-hot -> label) -- and gener_target_pseudo (:96-141), the offline
teacher pass that writes the soft pseudo labels the self-training step reads back."""
from __future__ import annotations

import torch

from .. import capi

_flags = {}


def _flag_word(device):
    f = _flags.get(device)
    if f is None:
        f = _flags[device] = torch.zeros(1, dtype=torch.int32, device=device)
    return f


def pseudo_selection(mask, cutoff_top=0.8, cutoff_low=0.6, return_type='ndarray', ignore_label=-1, check=True):
    """mask: float32 probabilities [b,c,h,w] -> pseudo label [b,h,w] int64 (np.ndarray or tensor).

    `check=True` reproduces the reference's `assert mask.max() <= 1 and mask.min() >= 0`
    (one host sync); pass False inside a captured / sync-free step."""
    assert return_type in ['ndarray', 'tensor']
    if not mask.is_cuda:
        raise RuntimeError("regda_b200.pseudo_selection needs a CUDA tensor (no CPU fallback)")
    bs, c, h, w = mask.size()
    soft = mask.detach().float().contiguous()
    out = torch.empty((bs, h, w), dtype=torch.int64, device=mask.device)
    flags = _flag_word(mask.device)
    ws = capi.workspace.get(capi.lib().regda_select_workspace_bytes(bs, c), mask.device)
    capi.call("regda_pseudo_select", capi.ptr(soft), capi.ptr(out), bs, c, h * w, float(cutoff_top), float(cutoff_low),
              int(ignore_label), capi.ptr(flags), capi.ptr(ws), ws.numel(), capi.stream())
    if check:
        f = int(flags.item())
        if f:
            flags.zero_()
            raise AssertionError("pseudo_selection: probabilities must lie in [0, 1]")
    if return_type == 'ndarray':
        return out.cpu().numpy()
    return out


def gener_target_pseudo(_cfg, model, pseudo_loader, save_pseudo_label_path, slide=True, save_prob=False, size=(1024, 1024),
                        ignore_label=-1, num_classes=None, viz_op=None, tile_size=(512, 512)):
    """pseudo_generation.py:96-141.  For every target image (loader batch size 1: `ret` float [1,3,H,W], `ret_gt['fname']`):
    eval-mode probabilities by 8-view TTA over 50 %-overlap sliding windows (pre_slide), then
      save_prob=True : torch.save(float32 [C, size_h, size_w]) to <path>/<fname>.pt -- the on-disk format
                       regda/datasets/basedata.py:86 reads back as `label_t_soft`;
      save_prob=False: pseudo_selection (if _cfg.PSEUDO_SELECT) or argmax, written as the uint8 image <path>/<fname> = label + 1
                       (0 = ignored), as the reference's cv2.imwrite does (:150-151).
    The forward runs on the hand-written inference kernels (bf16 tcgen05 convolutions, eval BatchNorm, fused upsample +
    softmax + head mean); the 8 TTA views of a tile are one batch.  `viz_op(pred, name)` is the optional colour dump the
    reference makes with VisualizeSegmm (plotting is out of scope); num_classes defaults to the model's."""
    import os

    import torch.nn.functional as tnf

    from ..utils.tools import pre_slide
    model.eval()
    os.makedirs(save_pseudo_label_path, exist_ok=True)
    if num_classes is None:
        num_classes = int(model.config.num_classes) if hasattr(model, "config") else 7
    dev = next(model.parameters()).device
    with torch.no_grad():
        for ret, ret_gt in pseudo_loader:
            ret = ret.to(dev)
            cls = pre_slide(model, ret, num_classes=num_classes, tile_size=tile_size, tta=True) if slide else model(ret)   # (b, c, h, w)
            if save_prob:
                torch.save(tnf.interpolate(cls, size, mode='bilinear', align_corners=True).squeeze(dim=0).cpu(),
                           os.path.join(save_pseudo_label_path, ret_gt['fname'][0] + '.pt'))                                  # (c, h, w)
                if viz_op is not None and getattr(_cfg, "SNAPSHOT_DIR", None) is not None:
                    sel = pseudo_selection(cls, ignore_label=ignore_label, cutoff_top=_cfg.CUTOFF_TOP, cutoff_low=_cfg.CUTOFF_LOW)
                    for fname, pred in zip(ret_gt['fname'], sel):
                        viz_op(pred, fname.replace('.tif', '.png'))
            else:
                if getattr(_cfg, "PSEUDO_SELECT", False):
                    sel = pseudo_selection(cls, ignore_label=ignore_label)                                                    # (b, h, w), -1..C-1
                else:
                    sel = cls.argmax(dim=1).cpu().numpy()
                import cv2
                import numpy as np
                cv2.imwrite(os.path.join(save_pseudo_label_path, ret_gt['fname'][0]), (sel + 1).reshape(*size).astype(np.uint8))   # :150-151
                if viz_op is not None and getattr(_cfg, "SNAPSHOT_DIR", None) is not None:
                    for fname, pred in zip(ret_gt['fname'], sel):
                        viz_op(pred, fname.replace('.tif', '.png'))
