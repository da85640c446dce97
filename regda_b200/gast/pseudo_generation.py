"""pseudo_selection -- drop-in for regda/gast/pseudo_generation.py:59-93 (two CUDA kernels:
per-(image, class) maximum, then the thresholded one-hot -> label)."""
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
