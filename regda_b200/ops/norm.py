"""Host side of the fused train-mode BatchNorm kernels (regda_b200/csrc/norm.cu).

    out = bn_act(y, bn, residual=None, relu=True, groups=1)

is `relu(bn(y) + residual)` of the reference's Bottleneck / PPM blocks (regda/_resnets.py:92-112,
regda/models/Encoder.py:24-40) for a channels-last bf16 activation `y` and an nn.BatchNorm2d module
`bn` in training mode: batch statistics in fp32, running statistics and num_batches_tracked updated
in place (they are part of the checkpoint ABI), gradients of gamma / beta ACCUMULATED straight into
`bn.weight.grad` / `bn.bias.grad` (views of the trainer's flat gradient arena).

`groups` = G > 1 splits the batch into G equal consecutive parts with independent batch statistics:
the source and the target batch of one training step travel through the network as ONE tensor
(twice the rows per convolution launch, half the launches) while BatchNorm keeps the reference's
per-forward-call statistics (tools/train_ssl_reg.py:210-212 calls the model once per domain);
running statistics are updated once per group, in order, exactly like two consecutive calls.
"""
from __future__ import annotations

import os

import torch

from .. import capi


def supported(y, bn) -> bool:
    if not (y.is_cuda and y.dtype == torch.bfloat16 and y.dim() == 4 and bn.training and bn.affine and bn.track_running_stats):
        return False
    n, c, h, w = y.shape
    return bool(capi.lib().regda_bn_supported(n * h * w, c))


def _cl(t):
    return t if t.is_contiguous(memory_format=torch.channels_last) else t.contiguous(memory_format=torch.channels_last)


def _grad_buffer(p):
    if p.grad is None:
        p.grad = torch.zeros_like(p)
    return p.grad


FUSE_BN_BWD = os.environ.get("REGDA_FUSE_BN_BWD", "1") != "0"


class BnHandle:
    """Rides on the OUTPUT of a BatchNorm+ReLU node so that the convolution(s) consuming it can do the first half of
    this BatchNorm's backward in their data-gradient epilogue (regda_conv_dgrad_bnred_bf16): mask the gradient with
    `mask`, accumulate sum(dz), sum(dz*y) into `red`.  The model arms a handle (`arm(out, k)`) where it knows that the
    output has exactly k consumers and all of them are Conv2d calls; a consumer that cannot fuse sets `broken`."""
    __slots__ = ("y", "mask", "groups", "red", "expected", "seen", "done", "broken")

    def __init__(self):
        self.y = self.mask = self.red = None
        self.groups, self.expected, self.seen, self.done, self.broken = 1, 0, 0, 0, False

    @property
    def fused(self):
        return self.expected > 0 and not self.broken


def arm(t, consumers):
    h = getattr(t, "_bn_handle", None)
    if h is not None:
        h.expected = consumers


class _BnActFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, y, residual, gamma, beta, bn, relu, groups, ready_stats, handle=None):
        y = _cl(y)
        n, c, h, w = y.shape
        npix = n * h * w
        res = _cl(residual) if residual is not None else None
        out = torch.empty_like(y)
        assert n % groups == 0, "statistics groups must divide the batch"
        if ready_stats is not None:
            assert ready_stats.shape == (groups, 2, c) and ready_stats.dtype == torch.float32
            stats, have, zeroed = ready_stats, 1, True
        else:
            stats, zeroed = capi.zero_pool.take((groups, 2, c), y.device)
            have = 0
        mask = None
        if handle is not None and relu:
            mask = torch.empty(npix * c // 8, dtype=torch.uint8, device=y.device)
            handle.y, handle.mask, handle.groups = y, mask, groups
        ctx.handle = handle if mask is not None else None
        capi.call("regda_bn_forward_bf16", capi.ptr_any(y), capi.ptr_any(res) if res is not None else None, capi.ptr_any(out), npix, c,
                  groups, capi.ptr(gamma), capi.ptr(beta), capi.ptr(bn.running_mean), capi.ptr(bn.running_var),
                  capi.ptr(bn.num_batches_tracked), float(bn.eps), float(bn.momentum if bn.momentum is not None else 0.1), int(relu),
                  capi.ptr_any(stats), have, int(zeroed), capi.ptr(mask) if mask is not None else None, capi.stream())
        # the backward recomputes the ReLU mask of a residual-free layer from y (bit-identical to this forward), so the
        # output is only kept where a residual entered it
        ctx.save_for_backward(y, out if (relu and residual is not None) else None, stats)
        ctx.gamma, ctx.beta, ctx.relu, ctx.has_res, ctx.groups, ctx.eps = gamma, beta, relu, residual is not None, groups, float(bn.eps)
        return out

    @staticmethod
    def backward(ctx, dout):
        y, out, stats = ctx.saved_tensors
        n, c, h, w = y.shape
        dout = _cl(dout)
        dy = torch.empty_like(y)
        gamma, beta = ctx.gamma, ctx.beta
        arena = getattr(gamma, "_arena", None)
        if arena is not None:
            arena.backward_reached(gamma)          # gradient buckets of the data-parallel trainer (parallel.GradBuckets)
        dgamma = _grad_buffer(gamma) if gamma.requires_grad else None
        dbeta = _grad_buffer(beta) if beta.requires_grad else None
        hd = ctx.handle
        if hd is not None and hd.fused:
            # every consumer was a convolution whose data-gradient epilogue masked its share of dout and accumulated the
            # two reductions: dout IS dz (and the residual branch's gradient), only the apply pass is left
            assert hd.seen == hd.expected == hd.done and hd.red is not None, "BatchNorm handle: consumer bookkeeping mismatch"
            capi.call("regda_bn_backward_bf16", capi.ptr_any(dout), None, capi.ptr_any(y), capi.ptr_any(dy), None, n * h * w, c,
                      ctx.groups, capi.ptr(gamma), capi.ptr(beta), capi.ptr_any(stats), ctx.eps,
                      capi.ptr(dgamma) if dgamma is not None else None, capi.ptr(dbeta) if dbeta is not None else None,
                      int(ctx.relu), capi.ptr_any(hd.red), 1, 1, capi.stream())
            hd.red, hd.done = None, 0          # (a second backward over a retained graph starts clean)
            return dy, (dout if (ctx.has_res and ctx.needs_input_grad[1]) else None), None, None, None, None, None, None, None
        assert hd is None or hd.done == 0, "BatchNorm handle: a consumer fused although the handle is broken"
        dres = torch.empty_like(y) if (ctx.has_res and ctx.needs_input_grad[1]) else None
        red, zeroed = capi.zero_pool.take((ctx.groups, 2, c), y.device)
        capi.call("regda_bn_backward_bf16", capi.ptr_any(dout), capi.ptr_any(out) if out is not None else None, capi.ptr_any(y),
                  capi.ptr_any(dy), capi.ptr_any(dres) if dres is not None else None, n * h * w, c, ctx.groups, capi.ptr(gamma),
                  capi.ptr(beta), capi.ptr_any(stats), ctx.eps, capi.ptr(dgamma) if dgamma is not None else None,
                  capi.ptr(dbeta) if dbeta is not None else None, int(ctx.relu), capi.ptr_any(red), int(zeroed), 0, capi.stream())
        # gamma / beta gradients were accumulated in place: nothing flows back through autograd for them
        return dy, dres, None, None, None, None, None, None, None


def bn_act(y, bn, residual=None, relu=True, groups=1, stats=None):
    """stats: the [groups][2][c] sums produced by the convolution's epilogue (ops/tc.py fprop(..., stats_groups)), if any"""
    handle = BnHandle() if (relu and FUSE_BN_BWD and torch.is_grad_enabled() and y.requires_grad) else None
    out = _BnActFn.apply(y, residual, bn.weight, bn.bias, bn, relu, groups, stats, handle)
    if handle is not None:
        out._bn_handle = handle
    return out


def inference_supported(y, bn) -> bool:
    """eval-mode fast path: running statistics, no autograd"""
    if bn.training or torch.is_grad_enabled() or not (y.is_cuda and y.dtype == torch.bfloat16 and y.dim() == 4):
        return False
    if not (bn.affine and bn.track_running_stats) or bn.running_mean is None:
        return False
    return bool(capi.lib().regda_bn_supported(y.shape[0] * y.shape[2] * y.shape[3], y.shape[1]))


def bn_inference(y, bn, residual=None, relu=True):
    """relu?(bn(y) + residual) with the running statistics (model.eval(): the offline teacher pass and evaluate())"""
    y = _cl(y)
    n, c, h, w = y.shape
    res = _cl(residual) if residual is not None else None
    out = torch.empty_like(y)
    capi.call("regda_bn_inference_bf16", capi.ptr_any(y), capi.ptr_any(res) if res is not None else None, capi.ptr_any(out), n * h * w, c,
              capi.ptr(bn.weight), capi.ptr(bn.bias), capi.ptr(bn.running_mean), capi.ptr(bn.running_var), float(bn.eps), int(relu),
              capi.stream())
    return out


def bn_eager(x, bn, groups=1):
    """nn.BatchNorm2d with the same statistics-group semantics, through torch (tiny maps, float32 parity mode)."""
    if groups == 1 or not bn.training:
        return bn(x)
    return torch.cat([bn(part) for part in x.chunk(groups, dim=0)], dim=0)


class _MaxPool3s2Fn(torch.autograd.Function):
    """F.max_pool2d(x, 3, 2, 1) for channels-last bf16 (the stem pool, regda/_resnets.py:153)"""

    @staticmethod
    def forward(ctx, x):
        x = _cl(x)
        n, c, h, w = x.shape
        y = torch.empty((n, c, (h - 1) // 2 + 1, (w - 1) // 2 + 1), dtype=x.dtype, device=x.device, memory_format=torch.channels_last)
        arg = torch.empty(y.shape, dtype=torch.uint8, device=x.device, memory_format=torch.channels_last)
        capi.call("regda_maxpool3s2_fwd_bf16", capi.ptr_any(x), capi.ptr_any(y), capi.ptr_any(arg), n, h, w, c, capi.stream())
        ctx.save_for_backward(arg)
        ctx.xshape = (n, c, h, w)
        return y

    @staticmethod
    def backward(ctx, dy):
        (arg,) = ctx.saved_tensors
        n, c, h, w = ctx.xshape
        dy = _cl(dy)
        dx = torch.empty((n, c, h, w), dtype=dy.dtype, device=dy.device, memory_format=torch.channels_last)
        capi.call("regda_maxpool3s2_bwd_bf16", capi.ptr_any(arg), capi.ptr_any(dy), capi.ptr_any(dx), n, h, w, c, capi.stream())
        return dx


def max_pool3s2(x):
    if x.is_cuda and x.dtype == torch.bfloat16 and x.dim() == 4 and x.shape[1] % 8 == 0:
        return _MaxPool3s2Fn.apply(x)
    return torch.nn.functional.max_pool2d(x, 3, 2, 1)


class _InstanceNormFn(torch.autograd.Function):
    """nn.InstanceNorm2d(C) (no affine, no running statistics; regda/models/Encoder.py:118-123) on channels-last bf16:
    BatchNorm with one statistics group per image and gamma = 1, beta = 0 -- the same kernels."""

    @staticmethod
    def forward(ctx, x, eps):
        x = _cl(x)
        n, c, h, w = x.shape
        out = torch.empty_like(x)
        stats, zeroed = capi.zero_pool.take((n, 2, c), x.device)
        capi.call("regda_bn_forward_bf16", capi.ptr_any(x), None, capi.ptr_any(out), n * h * w, c, n, None, None, None, None, None,
                  float(eps), 0.0, 0, capi.ptr_any(stats), 0, int(zeroed), None, capi.stream())
        ctx.save_for_backward(x, stats)
        ctx.eps = float(eps)
        return out

    @staticmethod
    def backward(ctx, dout):
        x, stats = ctx.saved_tensors
        n, c, h, w = x.shape
        dout = _cl(dout)
        dx = torch.empty_like(x)
        red, zeroed = capi.zero_pool.take((n, 2, c), x.device)
        capi.call("regda_bn_backward_bf16", capi.ptr_any(dout), None, capi.ptr_any(x), capi.ptr_any(dx), None, n * h * w, c, n, None,
                  None, capi.ptr_any(stats), ctx.eps, None, None, 0, capi.ptr_any(red), int(zeroed), 0, capi.stream())
        return dx, None


def instance_norm(x, eps=1e-5):
    return _InstanceNormFn.apply(x, eps)


def instance_norm_supported(x):
    if not (x.is_cuda and x.dtype == torch.bfloat16 and x.dim() == 4):
        return False
    n, c, h, w = x.shape
    return bool(capi.lib().regda_bn_supported(h * w, c))
