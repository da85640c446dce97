"""Host side of the tcgen05 convolution kernels (regda_b200/csrc/conv_tc.cu, conv_wgrad.cu): shape
predicates, the bf16 weight shadow, and the calls through the C ABI.

Tensors are torch channels-last 4-D tensors; `x.permute(0,2,3,1)` of such a tensor is the contiguous
NHWC array the kernels take, and a channels-last [O,I,kh,kw] weight is the OHWI matrix, so no
copies are made.  The data gradient reads the forward weights in place (MN-major B operand) and
the weight gradient is accumulated straight into the fp32 gradient buffer.
"""
from __future__ import annotations

import torch

from .. import capi

_shadow = {}      # id(weight) -> (version, bf16 OHWI weight)     (only for parameters outside a ParamArena)


def _geom(xshape, wshape, stride, padding, dilation):
    n, cin, h, w = xshape
    cout, cin2, r, s = wshape
    assert cin == cin2
    return n, h, w, cin, cout, r, s, stride, padding, dilation


def supports_fprop(xshape, wshape, stride, padding, dilation, dtype):
    if dtype != torch.bfloat16:
        return False
    return bool(capi.lib().regda_conv_fprop_supported(*_geom(xshape, wshape, stride, padding, dilation)))


def supports_fprop_stats(xshape, wshape, stride, padding, dilation, dtype, groups):
    """forward kernel + the BatchNorm statistics of its output for `groups` statistics groups (every M tile in one group)"""
    if dtype != torch.bfloat16:
        return False
    return bool(capi.lib().regda_conv_fprop_stats_supported(*_geom(xshape, wshape, stride, padding, dilation), groups))


def _up2_shape(xshape, wshape, padding, dilation):
    """(oh1, ow1): the stride-1 output size of the convolution -- the size of the zero-inserted dY of its stride-2 form"""
    n, cin, h, w = xshape
    cout, _, r, s = wshape
    return h + 2 * padding - dilation * (r - 1), w + 2 * padding - dilation * (s - 1)


def supports_dgrad(xshape, wshape, stride, padding, dilation, dtype):
    """stride 2 runs as the stride-1 data gradient of the zero-inserted dY (regda_zero_insert2_bf16)"""
    if dtype != torch.bfloat16 or stride not in (1, 2):
        return False
    return bool(capi.lib().regda_conv_dgrad_supported(*_geom(xshape, wshape, 1, padding, dilation)))


def supports_dgrad_bnred(xshape, wshape, stride, padding, dilation, dtype, groups):
    if dtype != torch.bfloat16 or stride not in (1, 2):
        return False
    return bool(capi.lib().regda_conv_dgrad_bnred_supported(*_geom(xshape, wshape, 1, padding, dilation), groups))


def supports_wgrad(xshape, wshape, stride, padding, dilation, dtype):
    if dtype != torch.bfloat16:
        return False
    return bool(capi.lib().regda_conv_wgrad_supported(*_geom(xshape, wshape, stride, padding, dilation)))


def weight_shadow(weight):
    """bf16 OHWI copy of a float32 [O,I,kh,kw] channels-last parameter.  Parameters that live in a
    trainer.ParamArena carry `_bf16` (a view of the arena's bf16 shadow, rewritten by the SGD
    kernel every step); anything else gets a version-tracked cached copy."""
    w16 = getattr(weight, "_bf16", None)
    if w16 is not None:
        return w16
    key = id(weight)
    ent = _shadow.get(key)
    if ent is None or ent[0] != weight._version or ent[1].device != weight.device or ent[2] is not weight:
        w16 = weight.detach().to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
        ent = _shadow[key] = (weight._version, w16, weight)
    return ent[1]


def hint_static(weight):
    """tell the next forward / data-gradient convolution launched from this thread that its weight operand is the bf16 copy of an
    arena parameter -- written once per step by the optimizer kernel (or by ParamArena.sync_shadow, which fences it), never by the
    kernel that precedes the convolution -- so the kernel may request its first weight tiles before its PDL wait (conv_tc.cu)"""
    if getattr(weight, "_bf16", None) is not None:
        capi.lib().regda_conv_hint_static_weights()


def _nhwc(t):
    """the tensor in channels-last memory (a copy only when the producer left it in another stride order)"""
    if not t.is_contiguous(memory_format=torch.channels_last):
        t = t.contiguous(memory_format=torch.channels_last)
    return t


def out_hw(h, w, r, s, stride, padding, dilation):
    return (h + 2 * padding - dilation * (r - 1) - 1) // stride + 1, (w + 2 * padding - dilation * (s - 1) - 1) // stride + 1


def fprop(x, w16, stride, padding, dilation, stats_groups=None, out_f32=False):
    """y = conv(x, w16).  With stats_groups = G the kernel's epilogue also accumulates the train-mode BatchNorm
    statistics of y (float32 [G][2][cout]: per-group per-channel sum, sum of squares) and (y, stats) is returned.
    out_f32: y is the float32 accumulator (regda_conv_fprop_bf16_f32out)."""
    n, cin, h, w = x.shape
    cout, _, r, s = w16.shape
    oh, ow = out_hw(h, w, r, s, stride, padding, dilation)
    y = torch.empty((n, cout, oh, ow), dtype=torch.float32 if out_f32 else torch.bfloat16, device=x.device, memory_format=torch.channels_last)
    x, w16 = _nhwc(x), _nhwc(w16)
    if out_f32:
        assert stats_groups is None
        capi.call("regda_conv_fprop_bf16_f32out", capi.ptr_any(x), capi.ptr_any(w16), capi.ptr_any(y), n, h, w, cin, cout, r, s,
                  stride, padding, dilation, capi.stream())
        return y
    if stats_groups is None:
        capi.call("regda_conv_fprop_bf16", capi.ptr_any(x), capi.ptr_any(w16), capi.ptr_any(y), n, h, w, cin, cout, r, s,
                  stride, padding, dilation, capi.stream())
        return y
    stats, zeroed = capi.zero_pool.take((stats_groups, 2, cout), x.device)
    capi.call("regda_conv_fprop_stats_bf16", capi.ptr_any(x), capi.ptr_any(w16), capi.ptr_any(y), n, h, w, cin, cout, r, s,
              stride, padding, dilation, capi.ptr_any(stats), stats_groups, int(zeroed), capi.stream())
    return y, stats


def zero_insert2(gy, oh1, ow1):
    """dY [n,c,oh,ow] of a stride-2 convolution -> [n,c,oh1,ow1] with dY at the even positions and zeros elsewhere"""
    n, c, oh, ow = gy.shape
    gy = _nhwc(gy)
    up = torch.empty((n, c, oh1, ow1), dtype=torch.bfloat16, device=gy.device, memory_format=torch.channels_last)
    capi.call("regda_zero_insert2_bf16", capi.ptr_any(gy), capi.ptr_any(up), n, oh, ow, oh1, ow1, c, capi.stream())
    return up


def _dgrad_input(gy, xshape, wshape, stride, padding, dilation):
    if stride == 1:
        return _nhwc(gy)
    assert stride == 2
    return zero_insert2(gy, *_up2_shape(xshape, wshape, padding, dilation))


def dgrad(gy, w16, xshape, stride, padding, dilation, addend=None, out_f32=False):
    """dX from dY and the forward weights [O,I,kh,kw] (bf16, channels-last); `addend` (same shape as x) is added in the
    kernel's epilogue.  out_f32: dX (and the addend) are float32."""
    n, cin, h, w = xshape
    cout, _, r, s = w16.shape
    odt = torch.float32 if out_f32 else torch.bfloat16
    gx = torch.empty((n, cin, h, w), dtype=odt, device=gy.device, memory_format=torch.channels_last)
    gy, w16 = _dgrad_input(gy, xshape, w16.shape, stride, padding, dilation), _nhwc(w16)
    if addend is not None:
        assert tuple(addend.shape) == tuple(xshape) and addend.dtype == odt
        addend = _nhwc(addend)
    capi.call("regda_conv_dgrad_bf16_f32out" if out_f32 else "regda_conv_dgrad_bf16", capi.ptr_any(gy), capi.ptr_any(w16), capi.ptr_any(gx),
              n, h, w, cin, cout, r, s, 1, padding, dilation, capi.ptr_any(addend) if addend is not None else None, capi.stream())
    return gx


def dgrad_bnred(gy, w16, xshape, stride, padding, dilation, addend, bn_y, relu_mask, red, groups):
    """dz = (dgrad(gy) + addend) * relu_mask, red[groups][2][cin] += (sum dz, sum dz * bn_y): the data gradient fused with the
    reductions of the BatchNorm(+ReLU) backward whose output is this convolution's input (regda_conv_dgrad_bnred_bf16)."""
    n, cin, h, w = xshape
    cout, _, r, s = w16.shape
    gx = torch.empty((n, cin, h, w), dtype=torch.bfloat16, device=gy.device, memory_format=torch.channels_last)
    gy, w16, bn_y = _dgrad_input(gy, xshape, w16.shape, stride, padding, dilation), _nhwc(w16), _nhwc(bn_y)
    assert tuple(bn_y.shape) == tuple(xshape) and bn_y.dtype == torch.bfloat16 and red.shape == (groups, 2, cin)
    if addend is not None:
        assert tuple(addend.shape) == tuple(xshape) and addend.dtype == torch.bfloat16
        addend = _nhwc(addend)
    capi.call("regda_conv_dgrad_bnred_bf16", capi.ptr_any(gy), capi.ptr_any(w16), capi.ptr_any(gx), n, h, w, cin, cout, r, s,
              1, padding, dilation, capi.ptr_any(addend) if addend is not None else None, capi.ptr_any(bn_y),
              capi.ptr_any(relu_mask), capi.ptr_any(red), groups, capi.stream())
    return gx


def wgrad_accumulate(gy, x, gw, stride, padding, dilation):
    """gw (float32 [O,I,kh,kw], channels-last memory) += dW(gy, x)"""
    n, cin, h, w = x.shape
    cout, _, r, s = gw.shape
    assert gw.dtype == torch.float32 and gw.is_contiguous(memory_format=torch.channels_last)
    gy, x = _nhwc(gy), _nhwc(x)
    capi.call("regda_conv_wgrad_bf16", capi.ptr_any(gy), capi.ptr_any(x), capi.ptr_any(gw), n, h, w, cin, cout, r, s,
              stride, padding, dilation, capi.stream())


# ---- float32 parity path: every float32 operand as three bf16 parts, products of the parts accumulated in fp32 -----------------
def split_bf16(t):
    """float32 t -> (hi, mid, lo) bf16 with t = hi + mid + lo up to the last bit or two of the 24-bit mantissa"""
    hi = t.to(torch.bfloat16)
    r1 = t - hi.float()
    mid = r1.to(torch.bfloat16)
    lo = (r1 - mid.float()).to(torch.bfloat16)
    return hi, mid, lo


def _cat_cl(parts, dim):
    return torch.cat(parts, dim=dim).contiguous(memory_format=torch.channels_last)


# Products kept: hi*hi (the main term) and the five cross terms down to 2^-16 of it (hi*mid, mid*hi, mid*mid, hi*lo, lo*hi); the
# dropped ones are <= 2^-24.  The tensor cores accumulate K in steps of 16 and -- unlike an IEEE fp32 sum -- TRUNCATE when they
# align an update with a large accumulator: ~2^-25 of the running sum per update, same sign (measured: 1.4e-4 of the output after
# the 6912 updates of the PPM fuse convolution's K = 36864 x 3, tests/test_conv_gpu.py).  The float32 path therefore (a) keeps
# the big hi*hi term and the 2^-8-times smaller cross terms in SEPARATE accumulators, and (b) cuts the reduction into chunks of
# at most kF32MaxK products of the main term per accumulator, summed in IEEE float32 outside the kernel (<= 72 updates: ~2e-6).
kF32MaxK = 1152
_A_PARTS = (0, 1, 1, 0, 2)          # activation part of cross term j  (hi, mid, mid, hi, lo)
_W_PARTS = (1, 0, 1, 2, 0)          # weight part of cross term j      (mid, hi, mid, lo, hi)


def _chunks(channels, taps):
    """channel ranges [lo, hi) (multiples of 64) with (hi - lo) * taps <= kF32MaxK (at least 64 channels)"""
    step = max(64, (kF32MaxK // taps) // 64 * 64)
    return [(lo, min(lo + step, channels)) for lo in range(0, channels, step)]


def supports_f32(xshape, wshape, stride, padding, dilation):
    n, cin, h, w = xshape
    cout = wshape[0]
    if cin % 64 or cout % 64 or stride not in (1, 2):
        return False
    return (supports_fprop(xshape, wshape, stride, padding, dilation, torch.bfloat16) and supports_dgrad(xshape, wshape, stride, padding, dilation, torch.bfloat16)
            and supports_wgrad(xshape, wshape, stride, padding, dilation, torch.bfloat16))


def fprop_f32(x, weight, stride, padding, dilation):
    """float32 convolution on the bf16 tensor-core kernels (see the note above): fp32 accumulation in tensor memory, float32
    partial outputs added in float32 -- float32-class accuracy with no library convolution."""
    xp = split_bf16(x)
    wp = split_bf16(weight.detach())
    r, s = weight.shape[2], weight.shape[3]
    y = None
    for lo, hi in _chunks(x.shape[1], r * s):
        main = fprop(_cat_cl([xp[0][:, lo:hi]], 1), _cat_cl([wp[0][:, lo:hi]], 1), stride, padding, dilation, out_f32=True)
        cross = fprop(_cat_cl([xp[i][:, lo:hi] for i in _A_PARTS], 1), _cat_cl([wp[i][:, lo:hi] for i in _W_PARTS], 1), stride, padding,
                      dilation, out_f32=True)
        part = main + cross
        y = part if y is None else y + part
    return y


def dgrad_f32(gy, weight, xshape, stride, padding, dilation, addend=None):
    """reduction over cout: dY parts against the weight parts stacked along O, read MN-major in place; chunked over cout"""
    gp = split_bf16(gy)
    wp = split_bf16(weight.detach())
    r, s = weight.shape[2], weight.shape[3]
    gx = addend
    for lo, hi in _chunks(gy.shape[1], r * s):
        main = dgrad(_cat_cl([gp[0][:, lo:hi]], 1), _cat_cl([wp[0][lo:hi]], 0), xshape, stride, padding, dilation, out_f32=True)
        cross = dgrad(_cat_cl([gp[i][:, lo:hi] for i in _A_PARTS], 1), _cat_cl([wp[i][lo:hi] for i in _W_PARTS], 0), xshape, stride, padding,
                      dilation, out_f32=True)
        part = main + cross
        gx = part if gx is None else gx + part
    return gx


def wgrad_accumulate_f32(gy, x, gw, stride, padding, dilation):
    """pixels are the reduction: the six part products are separate launches that add into gw through the kernel's fp32
    TMA reduce-stores (IEEE adds in L2); split-K keeps every accumulator short"""
    gp = [_nhwc(t) for t in split_bf16(gy)]
    xp = [_nhwc(t) for t in split_bf16(x)]
    wgrad_accumulate(gp[0], xp[0], gw, stride, padding, dilation)
    for i, j in zip(_A_PARTS, _W_PARTS):
        wgrad_accumulate(gp[i], xp[j], gw, stride, padding, dilation)
