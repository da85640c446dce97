"""Stem convolution Conv2d(3, 64, 7, stride 2, padding 3) (regda/_resnets.py:150) on the tcgen05 kernels.

Three input channels cannot form the 64-channel K-blocks of the implicit-GEMM kernel, so the stem is lowered explicitly:
`regda_stem_im2col_bf16` writes the patch matrix [n, oh, ow, 192] (7 filter rows of 21 taps, each padded to 24, + 24 zeros) and the forward / weight-gradient
kernels run it as a 1x1 convolution over 192 channels, with the BatchNorm statistics from the forward epilogue like every
other convolution of the network.  The image needs no data gradient.  (Replaces the library call this layer used to be.)
"""
from __future__ import annotations

import torch

from .. import capi
from . import tc

K_PAD = 192          # the patch row: 8 groups of 24 = filter row r as 21 real (s, c) taps + 3 zeros; group 7 all zero (csrc/stem.cu)


def _rows(t, cout):
    """the [cout, 7, 21] real-tap view of a [cout, 192(,1,1)] patch-ordered weight / weight-gradient tensor"""
    return t.view(cout, 8, 24)[:, :7, :21]


def supported(conv, x) -> bool:
    return (x.is_cuda and x.dtype in (torch.bfloat16, torch.float32) and x.dim() == 4 and x.shape[1] == 3 and conv.in_channels == 3
            and conv.out_channels % 64 == 0 and conv.kernel_size == 7 and conv.stride == 2 and conv.padding == 3 and conv.dilation == 1
            and conv.bias is None and not x.requires_grad)


class _StemConvFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, groups):
        """x: bf16 (channels-last image), or the loader's float32 NCHW image (rounded to bf16 inside the patch kernel)"""
        n, _, h, w = x.shape
        cout = weight.shape[0]
        oh, ow = (h - 1) // 2 + 1, (w - 1) // 2 + 1
        a = torch.empty((n, K_PAD, oh, ow), dtype=torch.bfloat16, device=x.device, memory_format=torch.channels_last)
        if x.dtype == torch.float32:
            x = x.contiguous()
            capi.call("regda_stem_im2col_f32nchw", capi.ptr(x), capi.ptr_any(a), n, h, w, capi.stream())
        else:
            x = x if x.is_contiguous(memory_format=torch.channels_last) else x.contiguous(memory_format=torch.channels_last)
            capi.call("regda_stem_im2col_bf16", capi.ptr_any(x), capi.ptr_any(a), n, h, w, capi.stream())
        w16 = tc.weight_shadow(weight)                       # bf16 [64,3,7,7] channels-last = OHWI rows of 147
        wp = torch.zeros((cout, K_PAD, 1, 1), dtype=torch.bfloat16, device=x.device).contiguous(memory_format=torch.channels_last)
        _rows(wp, cout).copy_(w16.permute(0, 2, 3, 1).reshape(cout, 7, 21))
        ctx.save_for_backward(a)
        ctx.weight = weight
        ctx.set_materialize_grads(False)
        if groups is None:                                   # no BatchNorm statistics wanted (inference, un-fused BatchNorm)
            return tc.fprop(a, wp, 1, 0, 1), None
        y, stats = tc.fprop(a, wp, 1, 0, 1, groups)
        ctx.mark_non_differentiable(stats)
        return y, stats

    @staticmethod
    def backward(ctx, gy, _gstats=None):
        if gy is None:
            return None, None, None
        a, = ctx.saved_tensors
        weight = ctx.weight
        cout = weight.shape[0]
        gy = gy.contiguous(memory_format=torch.channels_last)
        gw = torch.zeros((cout, K_PAD, 1, 1), dtype=torch.float32, device=gy.device).contiguous(memory_format=torch.channels_last)
        tc.wgrad_accumulate(gy, a, gw, 1, 0, 1)
        if weight.grad is None:
            weight.grad = torch.zeros_like(weight)
        # weight.grad is OHWI memory: 7 filter rows of 21 = (s, c) taps, the patch matrix's 24-padded row groups
        gview = weight.grad.permute(0, 2, 3, 1).reshape(cout, 7, 21)
        assert gview.data_ptr() == weight.grad.data_ptr(), "stem weight gradient must live in channels-last (OHWI) memory"
        gview.add_(_rows(gw, cout))
        return None, None, None


def _im2col(x16):
    n, _, h, w = x16.shape
    oh, ow = (h - 1) // 2 + 1, (w - 1) // 2 + 1
    x16 = x16 if x16.is_contiguous(memory_format=torch.channels_last) else x16.contiguous(memory_format=torch.channels_last)
    a = torch.empty((n, K_PAD, oh, ow), dtype=torch.bfloat16, device=x16.device, memory_format=torch.channels_last)
    capi.call("regda_stem_im2col_bf16", capi.ptr_any(x16), capi.ptr_any(a), n, h, w, capi.stream())
    return a


class _StemConvF32Fn(torch.autograd.Function):
    """float32 parity mode of the stem: the patch matrix is a pure gather, so im2col(x) = im2col(x_hi) + im2col(x_lo); the
    1x1 convolution over the 192-wide patches then runs as the split-operand product of ops/tc.py on the tcgen05 kernels."""

    @staticmethod
    def forward(ctx, x, weight):
        cout = weight.shape[0]
        ap = [_im2col(t) for t in tc.split_bf16(x)]
        wp = torch.zeros((cout, K_PAD), dtype=torch.float32, device=x.device)
        _rows(wp, cout).copy_(weight.detach().permute(0, 2, 3, 1).reshape(cout, 7, 21))
        wps = tc.split_bf16(wp.view(cout, K_PAD, 1, 1))
        y = (tc.fprop(ap[0], tc._cat_cl([wps[0]], 1), 1, 0, 1, out_f32=True) +
             tc.fprop(tc._cat_cl([ap[i] for i in tc._A_PARTS], 1), tc._cat_cl([wps[i] for i in tc._W_PARTS], 1), 1, 0, 1, out_f32=True))
        ctx.save_for_backward(*ap)
        ctx.weight = weight
        return y

    @staticmethod
    def backward(ctx, gy):
        ap = ctx.saved_tensors
        weight = ctx.weight
        cout = weight.shape[0]
        gp = [tc._nhwc(t) for t in tc.split_bf16(gy.contiguous(memory_format=torch.channels_last))]
        gw = torch.zeros((cout, K_PAD, 1, 1), dtype=torch.float32, device=gy.device).contiguous(memory_format=torch.channels_last)
        tc.wgrad_accumulate(gp[0], ap[0], gw, 1, 0, 1)
        for i, j in zip(tc._A_PARTS, tc._W_PARTS):
            tc.wgrad_accumulate(gp[i], ap[j], gw, 1, 0, 1)
        if weight.grad is None:
            weight.grad = torch.zeros_like(weight)
        gview = weight.grad.permute(0, 2, 3, 1).reshape(cout, 7, 21)
        assert gview.data_ptr() == weight.grad.data_ptr(), "stem weight gradient must live in channels-last (OHWI) memory"
        gview.add_(_rows(gw, cout))
        return None, None


def stem_conv(x, weight, groups, compute_dtype=None):
    """(y, bn_stats): y = conv7x7/2(x, weight) channels-last, bn_stats float32 [groups][2][cout] (bf16 compute), or
    (y float32, None) in the float32 parity mode (statistics are taken by the BatchNorm that follows).
    compute_dtype (default: x.dtype): bf16 compute accepts the float32 NCHW image of the data loader directly."""
    compute_dtype = compute_dtype or x.dtype
    if compute_dtype == torch.float32:
        return _StemConvF32Fn.apply(x.float(), weight), None
    return _StemConvFn.apply(x, weight, groups)
