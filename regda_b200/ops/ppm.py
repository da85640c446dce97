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
    """feat [b,c,h,w] channels-last bf16 -> pooled float32 [b, sum s^2, c] [, feat_tap].

    feat_tap (tap=True) is feat itself handed back as a second consumer handle (the mechanism of ops/conv.py _ConvFn): the other
    readers of the feature map go through it, so their gradient arrives HERE and the pooling backward adds it while it writes
    the feature map's gradient -- the one place where that gradient is assembled, no autograd add kernels over the 67 MB map."""

    @staticmethod
    def forward(ctx, feat, scales, tap=False):
        feat = _cl(feat)
        b, c, h, w = feat.shape
        ncell = sum(s * s for s in scales)
        pooled = torch.empty((b, ncell, c), dtype=torch.float32, device=feat.device)
        arr, n = _scales(scales)
        capi.call("regda_ppm_pool_fwd", capi.ptr_any(feat), capi.ptr(pooled), b, h, w, c, arr, n, capi.stream())
        ctx.geom = (b, c, h, w, tuple(scales))
        ctx.set_materialize_grads(False)
        return (pooled, feat.view_as(feat)) if tap else pooled

    @staticmethod
    def backward(ctx, dpooled, g_tap=None):
        b, c, h, w, scales = ctx.geom
        if dpooled is None:
            return g_tap, None, None
        add = None
        if g_tap is not None:
            add = _cl(g_tap.to(torch.bfloat16))
        dfeat = torch.empty((b, c, h, w), dtype=torch.bfloat16, device=dpooled.device, memory_format=torch.channels_last)
        arr, n = _scales(scales)
        capi.call("regda_ppm_pool_bwd_add", capi.ptr(dpooled.float().contiguous()), capi.ptr_any(add) if add is not None else None,
                  capi.ptr_any(dfeat), b, h, w, c, arr, n, capi.stream())
        return dfeat, None, None


class _CellsFn(torch.autograd.Function):
    """pooled float32 [b, sum s^2, c] -> the branch inputs p_k bf16 [b, c, s_k, s_k] (channels-last), one launch each way
    (the slice / reshape / permute / cast chain of autograd ops this replaces was ~40 small kernels per step)"""

    @staticmethod
    def forward(ctx, pooled, scales):
        b, ncell, c = pooled.shape
        ps = [torch.empty((b, s, s, c), dtype=torch.bfloat16, device=pooled.device).permute(0, 3, 1, 2) for s in scales]
        arr, n = _scales(scales)
        capi.call("regda_ppm_cells", capi.ptr(pooled), *([capi.ptr_any(t) for t in ps] + [None] * (4 - n)), b, c, arr, n, 1, capi.stream())
        ctx.geom = (b, ncell, c, tuple(scales))
        return tuple(ps)

    @staticmethod
    def backward(ctx, *dps):
        b, ncell, c, scales = ctx.geom
        dev = next(t for t in dps if t is not None).device
        gs = []
        for t, s in zip(dps, scales):
            if t is None:
                t = torch.zeros((b, c, s, s), dtype=torch.bfloat16, device=dev)
            gs.append(t.to(torch.bfloat16).permute(0, 2, 3, 1).contiguous())           # dense [b][s][s][c]
        dpooled = torch.empty((b, ncell, c), dtype=torch.float32, device=dev)
        arr, n = _scales(scales)
        capi.call("regda_ppm_cells", capi.ptr(dpooled), *([capi.ptr_any(t) for t in gs] + [None] * (4 - n)), b, c, arr, n, 0, capi.stream())
        return dpooled, None


def cells(pooled, scales=(1, 2, 3, 6)):
    """the s_k x s_k pooled maps of every scale as bf16 channels-last tensors [b, c, s_k, s_k]"""
    return _CellsFn.apply(pooled.contiguous(), tuple(scales))


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


def pool(feat, scales=(1, 2, 3, 6), tap=False):
    """pooled [, feat_tap]: see _PoolFn"""
    return _PoolFn.apply(feat, tuple(scales), tap)


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
