"""PrototypeContrastiveLoss -- drop-in for regda/loss.py:18-47, the alignment loss of the stage-2 step
(tools/train_align_reg.py:135,186-189).  One fused forward kernel + one backward kernel (regda_b200/csrc/pcl.cu): the masked
feature copy, its normalised copy and the [N,C] logits of the reference are never materialised."""
from __future__ import annotations

import torch
import torch.nn as nn

from . import capi

_flags = {}


def _flag_word(device):
    f = _flags.get(device)
    if f is None:
        f = _flags[device] = torch.zeros(1, dtype=torch.int32, device=device)
    return f


class _PclFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, feat, labels, proto, temperature, ignore_label):
        # feat: [N,K] float32 rows, or [b,K,h,w] (channels-last memory makes the [b*h*w, K] row view free)
        shape4 = None
        if feat.dim() == 4:
            shape4 = feat.shape
            rows = feat.permute(0, 2, 3, 1).reshape(-1, feat.shape[1])
        else:
            rows = feat
        rows = rows.float().contiguous()
        lab = labels.reshape(-1).long().contiguous()
        pr = proto.detach().float().contiguous()
        n, k = rows.shape
        c = pr.shape[0]
        assert lab.numel() == n and pr.shape[1] == k
        ws = torch.empty(capi.lib().regda_pcl_workspace_bytes(n), dtype=torch.uint8, device=rows.device)
        stats = torch.empty(2, dtype=torch.float32, device=rows.device)
        capi.call("regda_pcl_forward", capi.ptr(rows), capi.ptr(lab), capi.ptr(pr), capi.ptr(stats), n, k, c, int(ignore_label),
                  float(temperature), capi.ptr(_flag_word(rows.device)), capi.ptr(ws), ws.numel(), capi.stream())
        ctx.save_for_backward(rows, lab, pr, stats, ws)
        ctx.meta = (shape4, float(temperature), int(ignore_label), feat.dtype)
        return stats[0].clone()

    @staticmethod
    def backward(ctx, gout):
        rows, lab, pr, stats, ws = ctx.saved_tensors
        shape4, temperature, ignore_label, dtype = ctx.meta
        n, k = rows.shape
        up = gout.detach().float().reshape(1).contiguous()
        dfeat = torch.empty_like(rows)
        capi.call("regda_pcl_backward", capi.ptr(rows), capi.ptr(lab), capi.ptr(pr), capi.ptr(stats), capi.ptr(up), capi.ptr(dfeat), n, k,
                  pr.shape[0], ignore_label, temperature, capi.ptr(ws), ws.numel(), capi.stream())
        if shape4 is not None:
            b, _, h, w = shape4
            dfeat = dfeat.view(b, h, w, k).permute(0, 3, 1, 2)
        return dfeat.to(dtype), None, None, None, None


class PrototypeContrastiveLoss(nn.Module):
    def __init__(self, temperature=8.0, ignore_label=-1):
        super().__init__()
        self.temperature = temperature
        self.ignore_label = ignore_label

    def forward(self, Proto, feat, labels):
        """Proto (C, A) class means; feat (B, A, H, W) or (N, A); labels (B, 1, H, W) / (B, H, W) / (N,)"""
        assert not Proto.requires_grad and not labels.requires_grad and feat.requires_grad      # regda/loss.py:35
        if not feat.is_cuda:
            raise RuntimeError("regda_b200.PrototypeContrastiveLoss needs CUDA tensors (no CPU fallback)")
        return _PclFn.apply(feat, labels, Proto, self.temperature, self.ignore_label)

    def check(self, device=None):
        """host sync: raise if a label outside [0, C) that is not ignore_label was seen (nn.CrossEntropyLoss raises)"""
        for dev, f in _flags.items():
            if int(f.item()):
                f.zero_()
                raise IndexError("PrototypeContrastiveLoss: target out of bounds")
