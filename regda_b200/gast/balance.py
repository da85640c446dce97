"""CrossEntropy / ClassBalance -- drop-in for regda/gast/balance.py:15-101.

CrossEntropy is the default loss of both domains in tools/train_ssl_reg.py:133-158.  Its forward
here is one fused CUDA kernel (bilinear align_corners=True upsample of the low-resolution
logits + per-pixel CE + mean over ALL pixels) that also produces the gradient w.r.t. the
low-resolution logits, so loss_calc never materialises a [b,c,H,W] tensor.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from .. import capi


class _CEBilinear(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pred, label, ignore_label, class_weight, flags):
        b, c, h, w = pred.shape
        H, W = label.shape[-2:]
        p = pred.detach().float().contiguous()
        lab = label.long().contiguous()
        loss = torch.empty(1, dtype=torch.float32, device=pred.device)
        need_grad = pred.requires_grad
        dpred = torch.empty_like(p) if need_grad else None
        ws = capi.workspace.get(capi.lib().regda_ce_workspace_bytes(b, h, H), pred.device)
        capi.call("regda_ce_bilinear", capi.ptr(p), capi.ptr(lab), capi.ptr(loss), capi.ptr(dpred), b, c, h, w, H, W,
                  int(ignore_label), 1.0, capi.ptr(class_weight), capi.ptr(flags), capi.ptr(ws), ws.numel(), capi.stream())
        ctx.save_for_backward(dpred)
        ctx.in_dtype = pred.dtype
        return loss[0]

    @staticmethod
    def backward(ctx, g):
        (dpred,) = ctx.saved_tensors
        return (dpred * g).to(ctx.in_dtype), None, None, None, None


class ClassBalance(nn.Module):
    """balance.py:15-78 (flag-gated by --bcs/--bct; default off)."""

    def __init__(self, class_num=7, ignore_label=-1, decay=0.99, temperature=0.5):
        super().__init__()
        assert temperature > 0
        self.class_num = class_num
        self.ignore_label = ignore_label
        self.decay = decay
        self.temperature = temperature
        self.eps = 1e-7
        self.freq = torch.ones([class_num], device="cuda").float() / class_num
        self._flags = torch.zeros(1, dtype=torch.int32, device="cuda")

    def _local_freq(self, label):
        lab = label.long().contiguous()
        counts = torch.empty(self.class_num + 1, dtype=torch.int64, device=lab.device)
        capi.call("regda_class_count", capi.ptr(lab), lab.numel(), self.class_num, int(self.ignore_label), capi.ptr(counts),
                  capi.ptr(self._flags), capi.stream())
        return counts[:-1].float() / (counts[-1].float() + self.eps)

    def ema_update(self, label):
        # in place: the tensor's address is what a captured CUDA graph replays against
        self.freq.copy_((1.0 - self.decay) * self._local_freq(label) + self.decay * self.freq)

    def _get_class_wight(self):
        _prob = torch.softmax((1.0 - self.freq) / self.temperature, dim=0)
        return _prob / (_prob.max(dim=0, keepdim=True)[0] + self.eps)

    def get_class_weight(self, label):
        """EMA update + per-class weight [c] (the per-pixel gather of :30-33 happens inside the CE kernel)."""
        self.ema_update(label)
        return self._get_class_wight().detach().contiguous()

    def get_class_weight_4pixel(self, label):
        w = self.get_class_weight(label)
        lab = label.reshape(-1).long()
        return torch.where((lab >= 0) & (lab < self.class_num), w[lab.clamp(0, self.class_num - 1)], torch.zeros((), device=w.device))


class CrossEntropy(nn.Module):
    """balance.py:81-101."""

    def __init__(self, ignore_label=-1, class_balancer=None):
        super().__init__()
        self.ignore_label = ignore_label
        self.class_balancer = class_balancer
        self._flags = None

    def forward(self, preds, labels):
        """preds [B,C,h,w] logits (any resolution: they are bilinearly upsampled, align_corners=True,
        to the label size inside the kernel -- identity when the sizes match); labels [B,H,W]."""
        if not preds.is_cuda:
            raise RuntimeError("regda_b200.CrossEntropy needs CUDA tensors (no CPU fallback)")
        if self._flags is None or self._flags.device != preds.device:
            self._flags = torch.zeros(1, dtype=torch.int32, device=preds.device)
        cw = self.class_balancer.get_class_weight(labels) if self.class_balancer is not None else None
        return _CEBilinear.apply(preds, labels, self.ignore_label, cw, self._flags)
