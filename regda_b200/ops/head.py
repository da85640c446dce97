"""Host side of the PPM head's tail (regda_b200/csrc/misc.cu): Dropout2d(0.1) -> Conv2d(512, C, 1) + bias
(reference regda/models/Encoder.py:39-40) as one streaming kernel forward and one backward.

    logits = dropout_classifier(y, conv.weight, conv.bias, p, training)      # float32 [b, C, h, w] (the reference's NCHW logits)

y is the channels-last activation after the fuse conv's BatchNorm + ReLU (bf16, or float32 in parity mode).  The Dropout2d
mask (one keep/scale factor per image and channel) comes from a counter-based generator whose state lives in device memory,
so a replayed CUDA graph draws a fresh mask every step; torch's own Dropout2d random stream is not reproduced (parity tests
run with p = 0, as they must for any comparison of a dropout network).  Gradients of the weight / bias are ACCUMULATED into
`.grad` (views of the trainer's gradient arena), like every other parameter gradient of the model."""
from __future__ import annotations

import torch

from .. import capi

_rng_state = {}


def _state(device):
    key = torch.device(device).index or 0
    st = _rng_state.get(key)
    if st is None:
        seed = int(torch.initial_seed()) & 0x7FFFFFFFFFFFFFFF
        st = _rng_state[key] = torch.tensor([seed, 0], dtype=torch.int64, device=device)
    return st


def seed(value: int, device=None):
    """(re)seed the Dropout2d generator of `device` (default: current CUDA device)"""
    dev = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
    st = _state(dev)
    st.copy_(torch.tensor([int(value) & 0x7FFFFFFFFFFFFFFF, 0], dtype=torch.int64))


def supported(y, weight) -> bool:
    if not (y.is_cuda and y.dim() == 4 and y.dtype in (torch.bfloat16, torch.float32)):
        return False
    ncls, cin = weight.shape[0], weight.shape[1]
    octs = cin // 8
    return (tuple(weight.shape[2:]) == (1, 1) and cin == y.shape[1] and cin % 8 == 0 and ncls <= 8 and octs <= 256 and 256 % octs == 0
            and ncls * cin * 4 <= 48 * 1024)


def _grad_buffer(p):
    if p.grad is None:
        p.grad = torch.zeros_like(p)
    return p.grad


class _DropoutClassifierFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, y, weight, bias, p):
        if not y.is_contiguous(memory_format=torch.channels_last):
            y = y.contiguous(memory_format=torch.channels_last)
        b, cin, h, w = y.shape
        ncls = weight.shape[0]
        keep = None
        if p > 0.0:
            keep = torch.empty((b, cin), dtype=torch.float32, device=y.device)
            capi.call("regda_dropout2d_mask", capi.ptr(_state(y.device)), float(p), capi.ptr(keep), b * cin, capi.stream())
        out = torch.empty((b, ncls, h, w), dtype=torch.float32, device=y.device)
        wf = weight.detach().reshape(ncls, cin)
        assert wf.is_contiguous() and wf.dtype == torch.float32
        capi.call("regda_classifier_fwd", capi.ptr_any(y), int(y.dtype == torch.float32), capi.ptr(wf),
                  capi.ptr(bias.detach()) if bias is not None else None, capi.ptr(keep) if keep is not None else None, capi.ptr(out),
                  b, h * w, cin, ncls, capi.stream())
        ctx.save_for_backward(y, keep)
        ctx.weight, ctx.bias = weight, bias
        return out

    @staticmethod
    def backward(ctx, dout):
        y, keep = ctx.saved_tensors
        weight, bias = ctx.weight, ctx.bias
        arena = getattr(weight, "_arena", None)
        if arena is not None:
            arena.backward_reached(weight)
        b, cin, h, w = y.shape
        ncls = weight.shape[0]
        dout = dout.float().contiguous()
        dy = torch.empty_like(y)                       # channels-last, same dtype as y
        gw = _grad_buffer(weight)
        gb = _grad_buffer(bias) if bias is not None and bias.requires_grad else None
        gwf = gw.reshape(ncls, cin)
        assert gwf.data_ptr() == gw.data_ptr() and gw.dtype == torch.float32
        capi.call("regda_classifier_bwd", capi.ptr_any(y), int(y.dtype == torch.float32), capi.ptr(weight.detach().reshape(ncls, cin)),
                  capi.ptr(keep) if keep is not None else None, capi.ptr(dout), capi.ptr_any(dy), capi.ptr_any(gwf),
                  capi.ptr(gb) if gb is not None else None, b, h * w, cin, ncls, capi.stream())
        # weight / bias gradients were accumulated in place
        return dy, None, None, None


def dropout_classifier(y, weight, bias, p, training):
    hd = getattr(y, "_bn_handle", None)
    if hd is not None:
        hd.broken = True              # this consumer is not a convolution whose epilogue could carry the BatchNorm reductions
    return _DropoutClassifierFn.apply(y, weight, bias, float(p) if training else 0.0)
