"""Convolution front end of the model: `Conv2d` keeps the reference's parameter ABI
(weight [O,I,kh,kw] float32, optional bias) and dispatches the arithmetic.

Engines
  tcgen05 : the hand-written sm_100a implicit-GEMM kernels (regda_b200/csrc/conv_tc.cu fprop / dgrad,
            conv_wgrad.cu) -- bf16 operands, fp32 accumulation in tensor memory, TMA-fed.  bf16 activations run
            directly; float32 activations (the parity mode) run on the SAME kernels with every operand split into
            three bf16 parts (ops/tc.py fprop_f32).  A shape the kernels do not cover RAISES: there is no
            library fallback on the product path.
  cudnn   : torch.nn.functional.conv2d (library call) -- only as the measured baseline
            (bench.py `gpu_library_baseline`, scripts/bench_conv.py) and for A/B tests; never chosen implicitly.
Select with set_engine("tcgen05" | "cudnn") (REGDA_CONV presets it; "auto" is an alias of "tcgen05").
"""
from __future__ import annotations

import math
import os

import torch
import torch.nn as nn
import torch.nn.functional as F

ENGINE = {"auto": "tcgen05"}.get(os.environ.get("REGDA_CONV", "tcgen05"), os.environ.get("REGDA_CONV", "tcgen05"))
FUSE_BN_STATS = True          # BatchNorm statistics / backward reductions in the convolution epilogues
stats = {"tcgen05_fprop": 0, "tcgen05_dgrad": 0, "tcgen05_wgrad": 0, "cudnn": 0}

# Weight gradients are off the backward pass's critical path (only the optimiser reads them), so they are launched on a side
# stream: the data-gradient chain continues on the main stream and the two kernel families fill each other's launch gaps and
# tails.  The side stream forks from the main stream when dY exists (event) and is joined once, after loss.backward()
# (join_wgrad_stream, called by the trainer); inside a CUDA-graph capture this becomes a parallel branch of the graph.
WGRAD_STREAM = os.environ.get("REGDA_WGRAD_STREAM", "1") != "0"
_side_streams = {}
_side_used = set()
_side_active = False          # only inside `with wgrad_side_stream():` -- a caller that does not join must not get async gradients


def _wgrad_stream(device):
    key = torch.device(device).index or 0
    st = _side_streams.get(key)
    if st is None:
        st = _side_streams[key] = torch.cuda.Stream(device=device)
    return key, st


class wgrad_side_stream:
    """`with wgrad_side_stream(): loss.backward()` -- weight-gradient kernels of the enclosed backward pass go to the side
    stream; leaving the block makes the current stream wait for them."""

    def __enter__(self):
        global _side_active
        _side_active = WGRAD_STREAM
        return self

    def __exit__(self, *exc):
        global _side_active
        _side_active = False
        join_wgrad_stream()
        return False


def join_wgrad_stream():
    """main stream waits for every weight-gradient kernel launched on the side stream since the last join"""
    for key in list(_side_used):
        torch.cuda.current_stream(key).wait_stream(_side_streams[key])
    _side_used.clear()


def set_engine(name: str):
    global ENGINE
    assert name in ("auto", "tcgen05", "cudnn")
    ENGINE = "tcgen05" if name == "auto" else name


def _tc():
    from . import tc
    return tc


def _reached(param):
    """tell the trainer's gradient buckets that the backward pass has arrived at `param` (regda_b200/parallel.py GradBuckets)"""
    arena = getattr(param, "_arena", None)
    if arena is not None:
        arena.backward_reached(param)


class _ConvFn(torch.autograd.Function):
    """x: channels-last bf16 [N,C,H,W]; weight: float32 master [O,I,kh,kw] (channels-last memory).
    The weight gradient is ACCUMULATED into weight.grad by the wgrad kernel (no autograd add)."""

    @staticmethod
    def forward(ctx, x, weight, stride, padding, dilation, stats_groups=None, tap=False, bn_handle=None):
        """returns y [, bn_stats] [, x_tap].  x_tap (tap=True) is x itself handed back as a second consumer handle: a
        residual branch that reads the same x goes through it, so this node's backward receives that branch's gradient and
        the dgrad kernel adds it in its epilogue (dx = dgrad(dy) + d_tap) instead of autograd launching an add."""
        tc = _tc()
        w16 = tc.weight_shadow(weight)
        ctx.save_for_backward(x, w16)
        ctx.weight = weight
        ctx.geom = (stride, padding, dilation)
        ctx.has_stats, ctx.tap = stats_groups is not None, tap
        # x is the output of a BatchNorm+ReLU whose backward reductions this node's dgrad epilogue can do (ops/norm.py BnHandle)
        ctx.bn_handle = None
        if bn_handle is not None and bn_handle.expected > 0:
            if FUSE_BN_STATS and tc.supports_dgrad_bnred(x.shape, weight.shape, stride, padding, dilation, x.dtype, bn_handle.groups):
                bn_handle.seen += 1
                ctx.bn_handle = bn_handle
            else:
                bn_handle.broken = True
        ctx.set_materialize_grads(False)        # no zero-filled gradient tensors for the statistics / unused tap outputs
        stats["tcgen05_fprop"] += 1
        outs = []
        tc.hint_static(weight)
        if stats_groups is None:
            outs.append(tc.fprop(x, w16, stride, padding, dilation))
        else:
            y, bn_stats = tc.fprop(x, w16, stride, padding, dilation, stats_groups)
            ctx.mark_non_differentiable(bn_stats)
            outs += [y, bn_stats]
        if tap:
            outs.append(x.view_as(x))
        return outs[0] if len(outs) == 1 else tuple(outs)

    @staticmethod
    def backward(ctx, gy, *rest):
        tc = _tc()
        x, w16 = ctx.saved_tensors
        weight = ctx.weight
        _reached(weight)
        stride, padding, dilation = ctx.geom
        g_tap = rest[-1] if ctx.tap else None
        if gy is None:                           # only the tap branch carried a gradient
            assert ctx.bn_handle is None or not ctx.bn_handle.fused, "fused BatchNorm backward needs this node's gradient"
            return g_tap, None, None, None, None, None, None, None
        gy = gy.contiguous(memory_format=torch.channels_last)
        gx = gw = None
        if ctx.needs_input_grad[0]:
            if not tc.supports_dgrad(x.shape, weight.shape, stride, padding, dilation, x.dtype):
                raise RuntimeError(f"no tcgen05 data-gradient kernel for conv {tuple(x.shape)} x {tuple(weight.shape)} stride {stride}")
            stats["tcgen05_dgrad"] += 1
            fuse = g_tap is not None and g_tap.dtype == torch.bfloat16 and FUSE_BN_STATS
            hd = ctx.bn_handle
            tc.hint_static(weight)
            if hd is not None and hd.fused:
                assert g_tap is None or fuse
                if hd.red is None:
                    hd.red, zeroed = capi_zero_take((hd.groups, 2, x.shape[1]), x.device)
                    if not zeroed:
                        hd.red.zero_()
                gx = tc.dgrad_bnred(gy, w16, x.shape, stride, padding, dilation, g_tap, hd.y, hd.mask, hd.red, hd.groups)
                hd.done += 1
            else:
                gx = tc.dgrad(gy, w16, x.shape, stride, padding, dilation, addend=g_tap if fuse else None)
                if g_tap is not None and not fuse:
                    gx = gx + g_tap
        if ctx.needs_input_grad[1]:
            if not tc.supports_wgrad(x.shape, weight.shape, stride, padding, dilation, x.dtype):
                raise RuntimeError(f"no tcgen05 weight-gradient kernel for conv {tuple(x.shape)} x {tuple(weight.shape)}")
            stats["tcgen05_wgrad"] += 1
            if weight.grad is None:
                weight.grad = torch.zeros_like(weight)
            if _side_active:
                key, side = _wgrad_stream(gy.device)
                side.wait_stream(torch.cuda.current_stream())           # dY (and the zeroed gradient arena) are ready
                with torch.cuda.stream(side):
                    tc.wgrad_accumulate(gy, x, weight.grad, stride, padding, dilation)
                gy.record_stream(side)                                  # the allocator must not recycle dY / x under the kernel
                x.record_stream(side)
                _side_used.add(key)
            else:
                tc.wgrad_accumulate(gy, x, weight.grad, stride, padding, dilation)
        return gx, gw, None, None, None, None, None, None


class _ConvF32Fn(torch.autograd.Function):
    """float32 parity mode: x float32 channels-last, weight float32 master; every contraction runs on the tcgen05 kernels
    through the three-way bf16 operand split of ops/tc.py (fprop_f32 / dgrad_f32 / wgrad_accumulate_f32).  The weight gradient is
    accumulated into weight.grad like the bf16 path."""

    @staticmethod
    def forward(ctx, x, weight, stride, padding, dilation):
        tc = _tc()
        x = x.contiguous(memory_format=torch.channels_last)
        ctx.save_for_backward(x)
        ctx.weight = weight
        ctx.geom = (stride, padding, dilation)
        stats["tcgen05_fprop"] += 1
        return tc.fprop_f32(x, weight, stride, padding, dilation)

    @staticmethod
    def backward(ctx, gy):
        tc = _tc()
        (x,) = ctx.saved_tensors
        weight = ctx.weight
        _reached(weight)
        stride, padding, dilation = ctx.geom
        gy = gy.contiguous(memory_format=torch.channels_last)
        gx = None
        if ctx.needs_input_grad[0]:
            stats["tcgen05_dgrad"] += 1
            gx = tc.dgrad_f32(gy, weight, x.shape, stride, padding, dilation)
        if ctx.needs_input_grad[1]:
            stats["tcgen05_wgrad"] += 1
            if weight.grad is None:
                weight.grad = torch.zeros_like(weight)
            tc.wgrad_accumulate_f32(gy, x, weight.grad, stride, padding, dilation)
        return gx, None, None, None, None


def capi_zero_take(shape, device):
    from .. import capi
    return capi.zero_pool.take(shape, device)


class Conv2d(nn.Module):
    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, bias=True):
        super().__init__()
        self.in_channels, self.out_channels = in_channels, out_channels
        self.kernel_size, self.stride, self.padding, self.dilation = kernel_size, stride, padding, dilation
        w = torch.empty(out_channels, in_channels, kernel_size, kernel_size)
        nn.init.kaiming_normal_(w, mode="fan_out", nonlinearity="relu")          # _resnets.py:163
        self.weight = nn.Parameter(w.contiguous(memory_format=torch.channels_last))
        if bias:
            bound = 1.0 / math.sqrt(in_channels * kernel_size * kernel_size)
            self.bias = nn.Parameter(torch.empty(out_channels).uniform_(-bound, bound))
        else:
            self.register_parameter("bias", None)

    def extra_repr(self):
        return (f"{self.in_channels}, {self.out_channels}, kernel_size={self.kernel_size}, stride={self.stride}, "
                f"padding={self.padding}, dilation={self.dilation}, bias={self.bias is not None}")

    def forward_with_bn_stats(self, x, groups, tap=False):
        """(y, stats[, x_tap]): the convolution plus the BatchNorm statistics of its output from the kernel's epilogue
        (stats is None when the statistics groups do not align with this shape's tiles -- tiny maps -- or on the library
        engine); with tap=True also the handle a residual branch should read x through (see _ConvFn.forward)."""
        if (ENGINE != "cudnn" and self.bias is None and x.is_cuda and x.dtype == torch.bfloat16 and FUSE_BN_STATS
                and _tc().supports_fprop_stats(x.shape, self.weight.shape, self.stride, self.padding, self.dilation, x.dtype, groups)):
            return _ConvFn.apply(x, self.weight, self.stride, self.padding, self.dilation, groups, tap, getattr(x, "_bn_handle", None))
        hd = getattr(x, "_bn_handle", None)
        if hd is not None:
            hd.broken = True
        return (self.forward(x), None, x) if tap else (self.forward(x), None)

    def forward(self, x):
        hd = getattr(x, "_bn_handle", None)
        if hd is not None:
            hd.broken = True                     # plain forward: this consumer does not take part in the fused BatchNorm backward
        geom = (self.stride, self.padding, self.dilation)
        if ENGINE != "cudnn":
            # the product path: hand-written kernels or an error -- never a silent library / CPU fallback
            if not x.is_cuda:
                raise RuntimeError("regda_b200.Conv2d needs a CUDA tensor (no CPU fallback)")
            tc = _tc()
            if x.dtype == torch.bfloat16 and tc.supports_fprop(x.shape, self.weight.shape, *geom, x.dtype):
                y = _ConvFn.apply(x, self.weight, *geom)
            elif x.dtype == torch.float32 and tc.supports_f32(x.shape, self.weight.shape, *geom):
                y = _ConvF32Fn.apply(x, self.weight, *geom)
            else:
                raise RuntimeError(f"no tcgen05 kernel for conv {tuple(x.shape)} ({x.dtype}) x {tuple(self.weight.shape)} "
                                   f"stride {self.stride} (channel counts must be multiples of 64); set_engine('cudnn') is the "
                                   "library baseline, not a fallback")
            if self.bias is not None:
                y = y + self.bias.to(y.dtype).view(1, -1, 1, 1)
            return y
        stats["cudnn"] += 1
        b = self.bias.to(x.dtype) if self.bias is not None else None
        return F.conv2d(x, self.weight.to(x.dtype), b, self.stride, self.padding, self.dilation)
