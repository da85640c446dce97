"""Host side of the tcgen05 convolution kernels (filled in as the kernels land)."""
from __future__ import annotations


def supports_fprop(xshape, wshape, stride, padding, dilation, dtype):
    return False


def supports_dgrad(xshape, wshape, stride, padding, dilation, dtype):
    return False


def supports_wgrad(xshape, wshape, stride, padding, dilation, dtype):
    return False
