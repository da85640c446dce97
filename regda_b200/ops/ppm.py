"""Host side of the pyramid-pooling kernels (regda_b200/csrc/ppm.cu): the AdaptiveAvgPool2d of all
pool scales in one pass, and bilinear upsample + concat into the PPM head's 3x3-conv input
(reference regda/models/Encoder.py:43-52).  Channels-last bf16 feature maps."""
from __future__ import annotations

import ctypes

import torch

from .. import capi


def _scales(scales):
    arr = (ctypes.c_int * 4)(*(list(scales) + [1] * (4 - len(scales))))
    return arr, len(scales)


def _cl(t):
    return t if t.is_contiguous(memory_format=torch.channels_last) else t.contiguous(memory_format=torch.channels_last)


class _PoolFn(torch.autograd.Function):
    """feat [b,c,h,w] channels-last bf16 -> pooled float32 [b, sum s^2, c]"""

    @staticmethod
    def forward(ctx, feat, scales):
        feat = _cl(feat)
        b, c, h, w = feat.shape
        ncell = sum(s * s for s in scales)
        pooled = torch.empty((b, ncell, c), dtype=torch.float32, device=feat.device)
        arr, n = _scales(scales)
        capi.call("regda_ppm_pool_fwd", capi.ptr_any(feat), capi.ptr(pooled), b, h, w, c, arr, n, capi.stream())
        ctx.geom = (b, c, h, w, tuple(scales))
        return pooled

    @staticmethod
    def backward(ctx, dpooled):
        b, c, h, w, scales = ctx.geom
        dfeat = torch.empty((b, c, h, w), dtype=torch.bfloat16, device=dpooled.device, memory_format=torch.channels_last)
        arr, n = _scales(scales)
        capi.call("regda_ppm_pool_bwd", capi.ptr(dpooled.float().contiguous()), capi.ptr_any(dfeat), b, h, w, c, arr, n, capi.stream())
        return dfeat, None


class _UpcatFn(torch.autograd.Function):
    """cat([feat] + [bilinear_up(branch_k)], dim=1): feat [b,c,h,w], branch_k [b,cb,s_k,s_k], all channels-last bf16"""

    @staticmethod
    def forward(ctx, feat, scales, *branches):
        feat = _cl(feat)
        b, c, h, w = feat.shape
        cb = branches[0].shape[1]
        brs = [_cl(t.to(torch.bfloat16)) for t in branches]
        for t, s in zip(brs, scales):
            assert t.shape == (b, cb, s, s), (t.shape, s)
        cat = torch.empty((b, c + len(scales) * cb, h, w), dtype=torch.bfloat16, device=feat.device, memory_format=torch.channels_last)
        arr, n = _scales(scales)
        ptrs = [capi.ptr_any(t) for t in brs] + [None] * (4 - len(brs))
        capi.call("regda_ppm_upcat_fwd", capi.ptr_any(feat), *ptrs, capi.ptr_any(cat), b, h, w, c, cb, arr, n, capi.stream())
        ctx.geom = (b, c, h, w, cb, tuple(scales), [t.dtype for t in branches])
        return cat

    @staticmethod
    def backward(ctx, dcat):
        b, c, h, w, cb, scales, dtypes = ctx.geom
        dcat = _cl(dcat)
        dbr = [torch.empty((b, s, s, cb), dtype=torch.float32, device=dcat.device) for s in scales]
        arr, n = _scales(scales)
        ptrs = [capi.ptr(t) for t in dbr] + [None] * (4 - len(dbr))
        capi.call("regda_ppm_upcat_bwd", capi.ptr_any(dcat), *ptrs, b, h, w, c, cb, arr, n, capi.stream())
        grads = [t.permute(0, 3, 1, 2).to(dt) for t, dt in zip(dbr, dtypes)]
        return (dcat[:, :c], None, *grads)


def pool(feat, scales=(1, 2, 3, 6)):
    return _PoolFn.apply(feat, tuple(scales))


def upsample_concat(feat, branches, scales=(1, 2, 3, 6)):
    return _UpcatFn.apply(feat, tuple(scales), *branches)


def upsample_softmax_mean(x1, x2, size):
    """Eval tail of Deeplabv2.forward (regda/models/Encoder.py:152-155): mean over the heads of softmax(bilinear_align_corners(x))"""
    x1 = x1.float().contiguous()
    x2 = x2.float().contiguous() if x2 is not None else None
    b, c, h, w = x1.shape
    H, W = int(size[0]), int(size[1])
    out = torch.empty((b, c, H, W), dtype=torch.float32, device=x1.device)
    capi.call("regda_upsample_softmax_mean", capi.ptr(x1), capi.ptr(x2) if x2 is not None else None, capi.ptr(out), b, c, h, w, H, W,
              capi.stream())
    return out
