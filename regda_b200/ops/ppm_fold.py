"""Folded PPM fuse convolution (reference regda/models/Encoder.py:43-52 + the 3x3 convolution of :33-34).

The reference concatenates the 2048-channel feature map with four bilinearly upsampled pyramid branches (4 x 512 channels that
carry only 1 + 4 + 9 + 36 = 50 distinct values per image) and runs a 3x3 convolution over all 4096 channels -- 42.7 % of the
model's multiply-adds.  Upsampling and convolution are linear, so the branch half is computed WITHOUT materialising the
upsampled maps:

    G_k[(img, cell)][(o, tap)] = sum_c p_k[(img, cell)][c] * W[o][tap][2048 + 512 k + c]            (4 tiny GEMMs: 50 rows per image)
    y_ppm[img][px][o]          = sum_(cell, tap) A[px][(cell, tap)] * G[img][(cell, tap)][o]        (one GEMM, K = 450 -> 512)
    y                          = conv3x3(fin, W[:, :, :2048]) + y_ppm                               (half the reference's K)

A[px][(cell, tap)] = bilinear weight (align_corners=False, zero outside the map = the convolution's zero padding) of pyramid cell
`cell` at pixel px + tap: a CONSTANT of the geometry, built once per (map size, pool scales).  Every GEMM runs on the tcgen05
convolution kernels as a 1x1 convolution (ops/tc.py); y_ppm enters the 3x3 convolution as its epilogue addend, so the BatchNorm
statistics of y still come out of that epilogue.  Backward: the same contractions transposed (data gradient of the 3x3 conv over
2048 instead of 4096 channels, dG = A^T dy per image, dp_k = dG_k W_k, dW_k = dG_k^T p_k), weight gradients on the side stream.
Per 16-image step and head this removes 2 x 155 GFLOP (fwd + 2 bwd GEMMs) and the 134 MB concatenated tensor for ~30 GFLOP of
small GEMMs.  Same function as the reference up to rounding (bf16 mode rounds G and y_ppm instead of the upsampled activations).
"""
from __future__ import annotations

import ctypes

import torch
import torch.nn.functional as F

from .. import capi
from . import conv as conv_ops
from . import tc

_basis_cache = {}


def _kp(scales, taps=9):
    return (sum(s * s for s in scales) * taps + 63) // 64 * 64


def basis_weight(h, w, scales, device):
    """A as the weight of a 1x1 convolution: bf16 [h*w (cout), KP (cin), 1, 1] channels-last, A[px][cell*9 + tap]"""
    key = (h, w, tuple(scales), str(device))
    a = _basis_cache.get(key)
    if a is None:
        cols = []
        for s in scales:
            eye = torch.eye(s * s, dtype=torch.float32).view(s * s, 1, s, s)
            up = F.interpolate(eye, (h, w), mode="bilinear", align_corners=False)          # Encoder.py:48-51
            padded = F.pad(up[:, 0], (1, 1, 1, 1))                                         # the 3x3 convolution's zero padding
            taps = [padded[:, r:r + h, c:c + w].reshape(s * s, h * w) for r in range(3) for c in range(3)]
            cols.append(torch.stack(taps, dim=1).reshape(s * s * 9, h * w))                # row = cell * 9 + tap
        a32 = torch.cat(cols, 0).t().contiguous()                                          # [hw][ncell * 9]
        kp = _kp(scales)
        full = torch.zeros(h * w, kp)
        full[:, :a32.shape[1]] = a32
        a = full.to(device=device, dtype=torch.bfloat16).view(h * w, kp, 1, 1).contiguous(memory_format=torch.channels_last)
        _basis_cache[key] = a
    return a


def supported(fin, branches, conv, scales) -> bool:
    if conv_ops.ENGINE == "cudnn" or not (fin.is_cuda and fin.dtype == torch.bfloat16 and fin.dim() == 4):
        return False
    b, cf, h, w = fin.shape
    O, ct, r, s = conv.weight.shape
    if (r, s) != (3, 3) or conv.stride != 1 or conv.padding != 1 or conv.dilation != 1 or conv.bias is not None:
        return False
    nb = len(branches)
    if nb < 1 or nb > 4 or len(scales) != nb:
        return False
    cb = branches[0].shape[1]
    if ct != cf + nb * cb or cf % 64 or cb % 64 or O % 64 or (h * w) % 64:
        return False
    return all(t.dtype == torch.bfloat16 and tuple(t.shape) == (b, cb, sc, sc) for t, sc in zip(branches, scales))


def _scales(scales):
    return (ctypes.c_int * 4)(*(list(scales) + [1] * (4 - len(scales)))), len(scales)


def _cl(t):
    return t if t.is_contiguous(memory_format=torch.channels_last) else t.contiguous(memory_format=torch.channels_last)


def _off(t, elems):
    """device pointer `elems` elements into t's memory (a channel block of an OHWI weight / weight gradient used in place)"""
    return ctypes.c_void_p(t.data_ptr() + elems * t.element_size())


def _ptrs(ts):
    return [capi.ptr_any(t) for t in ts] + [None] * (4 - len(ts))


class _FoldedFuseFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, fin, weight, scales, groups, tap, *branches):
        fin = _cl(fin)
        brs = [_cl(t) for t in branches]
        b, cf, h, w = fin.shape
        O, ct = weight.shape[0], weight.shape[1]
        nb, cb = len(brs), brs[0].shape[1]
        hw, kp = h * w, _kp(scales)
        dev = fin.device
        bf = dict(dtype=torch.bfloat16, device=dev)
        w16 = _cl(tc.weight_shadow(weight))         # OHWI [O][9][ct]: every part of it is used IN PLACE (channel stride ct)
        # G_k = p_k . W_k^T: 1x1 convolutions on the s x s maps (several images per tcgen05 M tile) against the 9*O rows (o, tap) of
        # the weight's channel block [cf + k*cb, +cb): output [b, 9*O, s, s] channels-last = [(img, cell)][o*9 + tap]
        gs = [torch.empty((b, s, s, 9 * O), **bf).permute(0, 3, 1, 2) for s in scales]
        for k, (p, gk, s) in enumerate(zip(brs, gs, scales)):
            tc.hint_static(weight)
            capi.call("regda_conv_fprop_addend_bf16", capi.ptr_any(p), _off(w16, cf + k * cb), ct, capi.ptr_any(gk), b, s, s, cb, 9 * O, 1, 1, 1, 0, 1,
                      None, None, 1, 1, capi.stream())
        conv_ops.stats["tcgen05_fprop"] += nb
        gt = torch.empty((b * O, kp), **bf)                                        # GT[img][o][(cell, tap)]
        arr, n = _scales(scales)
        capi.call("regda_ppm_g_pack", *_ptrs(gs), capi.ptr(gt), b, O, 9, kp, arr, n, 1, capi.stream())
        aw = basis_weight(h, w, scales, dev)
        # YT[(img, o)][px] = sum_kappa GT[(img, o)][kappa] A[px][kappa]: a 1x1 convolution over the "image" whose pixels are the (img, o) rows
        yt = tc.fprop(gt.view(1, 1, b * O, kp).permute(0, 3, 1, 2), aw, 1, 0, 1)    # [1, hw, 1, b*O] channels-last = [img][o][px]
        yppm = torch.empty((b, O, h, w), memory_format=torch.channels_last, **bf)   # [img][px][o]
        capi.call("regda_transpose_bf16", capi.ptr_any(yt), capi.ptr_any(yppm), b, O, hw, capi.stream())
        y = torch.empty((b, O, h, w), memory_format=torch.channels_last, **bf)
        stats = None
        zeroed = True
        if groups is not None:
            stats, zeroed = capi.zero_pool.take((groups, 2, O), dev)
        tc.hint_static(weight)
        capi.call("regda_conv_fprop_addend_bf16", capi.ptr_any(fin), capi.ptr_any(w16), ct, capi.ptr_any(y), b, h, w, cf, O, 3, 3, 1, 1, 1,
                  capi.ptr_any(yppm), capi.ptr_any(stats) if stats is not None else None, groups or 1, int(zeroed), capi.stream())
        conv_ops.stats["tcgen05_fprop"] += 1
        ctx.save_for_backward(fin, w16, aw, *brs)
        ctx.weight, ctx.scales, ctx.nb, ctx.tap = weight, tuple(scales), nb, tap
        ctx.set_materialize_grads(False)
        if stats is not None:
            ctx.mark_non_differentiable(stats)
        # fin_tap: fin itself as a second consumer handle -- the gradient of whoever reads the feature map through it is added in
        # the epilogue of this node's data-gradient kernel (ops/conv.py _ConvFn has the same mechanism for residual branches)
        return (y, stats, fin.view_as(fin)) if tap else (y, stats)

    @staticmethod
    def backward(ctx, gy, _gstats=None, g_tap=None):
        if gy is None:                           # only the tap carried a gradient
            return (g_tap, None, None, None, None) + (None,) * ctx.nb
        saved = ctx.saved_tensors
        fin, w16, aw = saved[0], saved[1], saved[2]
        nb, scales, weight = ctx.nb, ctx.scales, ctx.weight
        brs = saved[3:3 + nb]
        conv_ops._reached(weight)
        gy = _cl(gy)
        b, cf, h, w = fin.shape
        O, ct = weight.shape[0], weight.shape[1]
        cb = brs[0].shape[1]
        hw, kp = h * w, _kp(scales)
        dev = gy.device
        bf = dict(dtype=torch.bfloat16, device=dev)
        need_fin = ctx.needs_input_grad[0]
        need_br = any(ctx.needs_input_grad[5:])
        # dG = A^T dy per image: the data gradient of the YT convolution, on dy transposed to [img][o][px]
        gyt = torch.empty((b * O, hw), **bf)
        capi.call("regda_transpose_bf16", capi.ptr_any(gy), capi.ptr(gyt), b, hw, O, capi.stream())
        dgt = tc.dgrad(gyt.view(1, 1, b * O, hw).permute(0, 3, 1, 2), aw, (1, kp, 1, b * O), 1, 0, 1)     # [1, kp, 1, b*O] channels-last
        dgs = [torch.empty((b, s, s, 9 * O), **bf).permute(0, 3, 1, 2) for s in scales]      # [(img, cell)][o*9 + tap]
        arr, n = _scales(scales)
        capi.call("regda_ppm_g_pack", *_ptrs(dgs), capi.ptr_any(dgt), b, O, 9, kp, arr, n, 0, capi.stream())
        dps = [None] * nb
        if need_br:
            # dp_k = dG_k . W_k: data gradients of the branch GEMMs, reading the weight's channel block in place
            dps = [torch.empty_like(p) for p in brs]
            for k, (dg, dp, s) in enumerate(zip(dgs, dps, scales)):
                tc.hint_static(weight)
                capi.call("regda_conv_dgrad_wslice_bf16", capi.ptr_any(dg), _off(w16, cf + k * cb), ct, capi.ptr_any(dp), b, s, s, cb, 9 * O,
                          1, 1, 1, 0, 1, None, capi.stream())
            conv_ops.stats["tcgen05_dgrad"] += nb
        dfin = None
        if need_fin:
            dfin = torch.empty_like(fin)
            add = _cl(g_tap.to(torch.bfloat16)) if g_tap is not None else None
            tc.hint_static(weight)
            capi.call("regda_conv_dgrad_wslice_bf16", capi.ptr_any(gy), capi.ptr_any(w16), ct, capi.ptr_any(dfin), b, h, w, cf, O, 3, 3, 1, 1, 1,
                      capi.ptr_any(add) if add is not None else None, capi.stream())
        elif g_tap is not None:
            dfin = g_tap
        conv_ops.stats["tcgen05_dgrad"] += 1
        # weight gradients (feature part + the four branch parts) and their scatter into the OHWI gradient: off the critical path
        if weight.grad is None:
            weight.grad = torch.zeros_like(weight)

        assert weight.grad.dtype == torch.float32 and weight.grad.is_contiguous(memory_format=torch.channels_last)

        def wgrads():
            # every part is accumulated straight into the OHWI gradient: the feature part into columns [0, cf) of every tap, branch k
            # (dW_k = dG_k^T p_k, rows (o, tap)) into columns [cf + k*cb, +cb)
            capi.call("regda_conv_wgrad_wslice_bf16", capi.ptr_any(gy), capi.ptr_any(fin), capi.ptr_any(weight.grad), ct, b, h, w, cf, O, 3, 3,
                      1, 1, 1, capi.stream())
            for k, (dg, p, s) in enumerate(zip(dgs, brs, scales)):
                capi.call("regda_conv_wgrad_wslice_bf16", capi.ptr_any(dg), capi.ptr_any(p), _off(weight.grad, cf + k * cb), ct, b, s, s, cb, 9 * O,
                          1, 1, 1, 0, 1, capi.stream())

        if conv_ops._side_active:
            key, side = conv_ops._wgrad_stream(dev)
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                wgrads()
            for t in [gy, fin, *dgs, *brs]:
                t.record_stream(side)
            conv_ops._side_used.add(key)
        else:
            wgrads()
        conv_ops.stats["tcgen05_wgrad"] += 1
        return (dfin, None, None, None, None, *dps)


def fuse_conv(fin, branches, weight, scales, groups, tap=False):
    """(y, bn_stats [, fin_tap]): y = conv3x3(cat(fin, up(branch_k)...), weight) without the concatenation; bn_stats float32
    [groups][2][cout] from the convolution's epilogue (None when groups is None: inference); tap=True also hands fin back as the
    handle its next reader should use (that reader's gradient is then added inside this node's data-gradient kernel)"""
    for t in branches:
        hd = getattr(t, "_bn_handle", None)
        if hd is not None:
            hd.broken = True          # consumed by this op, not by a convolution whose epilogue carries BatchNorm reductions
    return _FoldedFuseFn.apply(fin, weight, tuple(scales), groups, tap, *branches)
