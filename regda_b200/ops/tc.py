"""Host side of the tcgen05 convolution kernels (regda_b200/csrc/conv_tc.cu): shape predicates,
bf16 weight shadows in the kernels' layouts, and the calls through the C ABI.

Tensors are torch channels-last 4-D tensors; `x.permute(0,2,3,1)` of such a tensor is the
contiguous NHWC array the kernels take, so no copies are made.
"""
from __future__ import annotations

import torch

from .. import capi

_shadow = {}      # id(weight) -> (version, bf16 OHWI weight)
_shadow_t = {}    # id(weight) -> (version, bf16 [cin][r][s][cout] flipped weight for dgrad)


def _geom(xshape, wshape, stride, padding, dilation):
    n, cin, h, w = xshape
    cout, cin2, r, s = wshape
    assert cin == cin2
    return n, h, w, cin, cout, r, s, stride, padding, dilation


def supports_fprop(xshape, wshape, stride, padding, dilation, dtype):
    if dtype != torch.bfloat16:
        return False
    return bool(capi.lib().regda_conv_fprop_supported(*_geom(xshape, wshape, stride, padding, dilation)))


def supports_dgrad(xshape, wshape, stride, padding, dilation, dtype):
    """dX of a stride-1 conv = conv of dY with the flipped, transposed weights and pad' = dil*(r-1) - pad."""
    if dtype != torch.bfloat16 or stride != 1:
        return False
    n, cin, h, w = xshape
    cout, _, r, s = wshape
    oh = h + 2 * padding - dilation * (r - 1)
    ow = w + 2 * padding - dilation * (s - 1)
    pad2 = dilation * (r - 1) - padding
    if pad2 < 0 or r != s:
        return False
    return bool(capi.lib().regda_conv_fprop_supported(n, oh, ow, cout, cin, r, s, 1, pad2, dilation))


def supports_wgrad(xshape, wshape, stride, padding, dilation, dtype):
    return False


def weight_shadow(weight):
    """bf16 copy of a float32 [O,I,kh,kw] channels-last parameter = the OHWI matrix [O][kh*kw*I]."""
    key = id(weight)
    ent = _shadow.get(key)
    if ent is None or ent[0] != weight._version or ent[1].device != weight.device:
        w16 = weight.detach().to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
        ent = _shadow[key] = (weight._version, w16)
    return ent[1]


def weight_shadow_t(weight):
    """bf16 [I][kh][kw][O] with both filter axes flipped (the dgrad operand), as a channels-last [I,O,kh,kw] tensor."""
    key = id(weight)
    ent = _shadow_t.get(key)
    if ent is None or ent[0] != weight._version or ent[1].device != weight.device:
        wt = weight.detach().to(torch.bfloat16).flip(2, 3).transpose(0, 1).contiguous(memory_format=torch.channels_last)
        ent = _shadow_t[key] = (weight._version, wt)
    return ent[1]


def _nhwc(t):
    """the tensor in channels-last memory (a copy only when the producer left it in another stride order,
    e.g. the bilinear-upsampled PPM branches)"""
    if not t.is_contiguous(memory_format=torch.channels_last):
        t = t.contiguous(memory_format=torch.channels_last)
    return t


def fprop(x, w16, stride, padding, dilation):
    n, cin, h, w = x.shape
    cout, _, r, s = w16.shape
    oh = h + 2 * padding - dilation * (r - 1)
    ow = w + 2 * padding - dilation * (s - 1)
    y = torch.empty((n, cout, oh, ow), dtype=torch.bfloat16, device=x.device, memory_format=torch.channels_last)
    x, w16 = _nhwc(x), _nhwc(w16)
    capi.call("regda_conv_fprop_bf16", capi.ptr_any(x), capi.ptr_any(w16), capi.ptr_any(y), n, h, w, cin, cout, r, s,
              stride, padding, dilation, capi.stream())
    return y


def dgrad(gy, wt16, xshape, stride, padding, dilation):
    n, cin, h, w = xshape
    _, cout, r, s = wt16.shape          # wt16 is [I,O,kh,kw]
    pad2 = dilation * (r - 1) - padding
    gx = fprop(gy, wt16, 1, pad2, dilation)
    assert gx.shape == (n, cin, h, w), (gx.shape, xshape)
    return gx


def wgrad(gy, x, wshape, stride, padding, dilation):
    raise NotImplementedError
