"""CoralLoss -- drop-in for regda/gast/coral.py:14-47 (Deep CORAL, eq. (1) of arXiv:1607.01719), the domain-alignment loss of
stages 1-2 (`Aligner.align_domain`, regda/gast/alignment.py:79-84; tools/train_align_reg.py:187 with --align-domain 1, which
the reference's shipped recipe runs/regda/run_2potsdam.sh:15 turns on).

    loss = || C_s - C_t ||_F^2 / (4 d^2),   C = Xc^T Xc / (n - 1),   Xc = mean_rows(X) - X            (coral.py:36-46)

B200 mapping: the two d x d covariances (d = 2048, n = 8192 rows per domain at config P: 2 x 68.7 GFLOP) are exactly the weight-
gradient contraction of a 1x1 convolution -- sum over pixels of an outer product of two channel vectors -- so they run on the
tcgen05 weight-gradient kernel (regda_conv_wgrad_bf16 with dY = X = the centred rows, fp32 accumulation in tensor memory, fp32
TMA reduce-stores); the backward pass dX_s = Xc_s (C_s - C_t) / (d^2 (n_s - 1)), dX_t = -Xc_t (C_s - C_t) / (d^2 (n_t - 1)) is a
1x1 forward convolution with the (symmetric) covariance difference as the weight (regda_conv_fprop_bf16).  bf16 operands in
the bf16 training mode; in the float32 parity mode the same kernels through the three-way operand split of ops/tc.py.
The centring itself is one column mean + subtract over [n, d] (HBM-bound, 67 MB per domain)."""
from __future__ import annotations

import torch
import torch.nn as nn

from ..ops import tc


def _as_image(rows):
    """[n, d] row-major -> the [1, d, 1, n] channels-last view the convolution kernels take (same memory)"""
    n, d = rows.shape
    return rows.view(1, 1, n, d).permute(0, 3, 1, 2)


def _covariance_sum(xc, precise):
    """Xc^T Xc, float32 [d, d], on the tcgen05 weight-gradient kernel (xc float32 [n, d], centred)"""
    n, d = xc.shape
    gw = torch.zeros((d, d, 1, 1), dtype=torch.float32, device=xc.device).contiguous(memory_format=torch.channels_last)
    if precise:
        img = _as_image(xc.contiguous())
        tc.wgrad_accumulate_f32(img, img, gw, 1, 0, 1)
    else:
        img = _as_image(xc.to(torch.bfloat16).contiguous())
        tc.wgrad_accumulate(img, img, gw, 1, 0, 1)
    return gw.view(d, d)


def _times_symmetric(xc, m, precise):
    """Xc @ M for a symmetric float32 M [d, d], float32 [n, d], as a 1x1 forward convolution with weight M"""
    n, d = xc.shape
    w = m.view(d, d, 1, 1).contiguous(memory_format=torch.channels_last)
    if precise:
        y = tc.fprop_f32(_as_image(xc.contiguous()), w, 1, 0, 1)
    else:
        y = tc.fprop(_as_image(xc.to(torch.bfloat16).contiguous()), w.to(torch.bfloat16), 1, 0, 1, out_f32=True)
    return y.permute(0, 2, 3, 1).reshape(n, d)


class _CoralFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, source, target, is_sqrt, precise):
        d = source.shape[1]
        ns, nt = source.shape[0], target.shape[0]
        xs, xt = source.detach().float(), target.detach().float()
        xms = xs.mean(0, keepdim=True) - xs                                    # coral.py:36
        xmt = xt.mean(0, keepdim=True) - xt                                    # coral.py:40
        diff = _covariance_sum(xms, precise) / (ns - 1) - _covariance_sum(xmt, precise) / (nt - 1)      # coral.py:37, 41
        sq = (diff * diff).sum()                                               # coral.py:44
        loss = (sq.sqrt() if is_sqrt else sq) / (4 * d * d)                    # coral.py:45
        ctx.save_for_backward(xms, xmt, diff, sq)
        ctx.cfg = (is_sqrt, precise, d, ns, nt, source.dtype, target.dtype)
        return loss

    @staticmethod
    def backward(ctx, g):
        xms, xmt, diff, sq = ctx.saved_tensors
        is_sqrt, precise, d, ns, nt, sdt, tdt = ctx.cfg
        # dL/dC_s = 2 (C_s - C_t) / (4 d^2) [x 1 / (2 sqrt(S)) if is_sqrt]; C = Xm^T Xm / (n - 1) -> dL/dXm = 2 Xm dL/dC / (n - 1);
        # Xm = mean - X and the rows of Xm sum to zero -> dL/dX = -dL/dXm
        scale = g / (d * d)
        if is_sqrt:
            scale = scale / (2.0 * sq.sqrt().clamp_min(1e-30))
        gs = gt = None
        if ctx.needs_input_grad[0]:
            gs = -_times_symmetric(xms, diff * (scale / (ns - 1)), precise)
        if ctx.needs_input_grad[1]:
            gt = _times_symmetric(xmt, diff * (scale / (nt - 1)), precise)
        return (gs.to(sdt) if gs is not None else None), (gt.to(tdt) if gt is not None else None), None, None


class CoralLoss(nn.Module):
    def __init__(self, is_sqrt=False, precise=False):
        """precise: run the contractions at float32 accuracy (operand split) instead of bf16 operands"""
        super().__init__()
        self.is_sqrt = is_sqrt
        self.precise = precise

    def forward(self, source, target):
        """source, target: [instance num, feature dimension] (coral.py:26-33); feature dimension a multiple of 64"""
        if not (source.is_cuda and target.is_cuda):
            raise RuntimeError("regda_b200.CoralLoss needs CUDA tensors (no CPU fallback)")
        if source.dim() != 2 or target.dim() != 2 or source.shape[1] != target.shape[1] or source.shape[1] % 64:
            raise ValueError("CoralLoss: [n, d] inputs with the same d (a multiple of 64)")
        return _CoralFn.apply(source, target, self.is_sqrt, self.precise)
