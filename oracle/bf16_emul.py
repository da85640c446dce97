"""TEST INFRASTRUCTURE ONLY -- bf16-emulating forward of the Deeplabv2 oracle (oracle/step_oracle.py DeeplabOracle).

The benchmarked configuration of regda_b200 computes in bf16: every activation tensor the kernels WRITE is rounded to bf16,
every convolution reads bf16 weights, and everything in between (tensor-memory accumulation, BatchNorm statistics and
normalisation constants, pooling sums, bilinear interpolation, the classifier) is fp32.  This module restates the reference's
forward (regda/models/Encoder.py:8-65,129-155, regda/resnet.py:140-166, regda/_resnets.py:92-112) with EXACT arithmetic
(float64) between those rounding points and a round-to-nearest-even bf16 cast AT each of them, i.e. the function the bf16
kernels compute up to fp32 accumulation order.  Comparing the CUDA bf16 path with it on the reference-generated golden inputs
(tests/test_bf16_parity_gpu.py) ties the tcgen05 network to the reference's architecture and weights at 1e-3-class tolerance --
a bound that the float32 golden comparison cannot give for bf16 arithmetic.

Rounding points (regda_b200 file that rounds):
  image -> bf16 (models/Encoder.py Deeplabv2.forward); conv weights -> bf16 (ops/tc.py weight_shadow / optim.cu shadow);
  conv output -> bf16 (csrc/conv_tc.cu epilogue); BatchNorm statistics from the ROUNDED conv output (same epilogue);
  BatchNorm(+residual)(+ReLU) output -> bf16 (csrc/norm.cu bn_apply_kernel); InstanceNorm output -> bf16 (same kernel);
  pooled pyramid maps -> bf16 (models/Encoder.py PPMBilinear.forward); the folded fuse convolution (ops/ppm_fold.py): branch
  GEMM G -> bf16, interpolation weights A -> bf16, their product y_ppm -> bf16 before it joins the 3x3 convolution's accumulator;
  the classifier reads the fp32 master weights and writes fp32 logits (csrc/misc.cu) -- no rounding.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F


ACC32 = False     # True: convolutions accumulate in float32 (torch CPU order) instead of exactly -- a second, equally valid
                  # evaluation of the same bf16 function, used to measure the arithmetic's own run-to-run noise floor
ROUND = True      # False: no rounding at all -- the restatement must then reproduce the reference's float32 goldens (the pin
                  # of this file: tests/test_oracle_golden.py::test_bf16_emulation_without_rounding_is_the_reference)


def r16(t):
    """round to nearest-even bf16, keep carrying float64"""
    return t.float().bfloat16().double() if ROUND else t.double()


def _conv(x, conv):
    w = r16(conv.weight.detach())
    if ACC32:
        return r16(F.conv2d(x.float(), w.float(), None, conv.stride, conv.padding, conv.dilation))
    return r16(F.conv2d(x, w, None, conv.stride, conv.padding, conv.dilation))


def _bn_train(y, bn, relu, residual=None, groups=1):
    """train-mode BatchNorm on the bf16 tensor y with `groups` statistics groups over equal batch parts, fp32-style constants"""
    outs = []
    for part, res in zip(y.chunk(groups, 0), (residual.chunk(groups, 0) if residual is not None else [None] * groups)):
        mean = part.mean(dim=(0, 2, 3), keepdim=True)
        var = (part * part).mean(dim=(0, 2, 3), keepdim=True) - mean * mean
        rstd = 1.0 / torch.sqrt(var.clamp_min(0.0) + bn.eps)
        g = bn.weight.detach().double().view(1, -1, 1, 1) if bn.weight is not None else 1.0
        b = bn.bias.detach().double().view(1, -1, 1, 1) if bn.bias is not None else 0.0
        o = (part - mean) * rstd * g + b
        if res is not None:
            o = o + res
        if relu:
            o = o.clamp_min(0.0)
        outs.append(r16(o))
    return torch.cat(outs, 0)


def _bottleneck(blk, x, groups):
    o = _bn_train(_conv(x, blk.conv1), blk.bn1, True, groups=groups)
    o = _bn_train(_conv(o, blk.conv2), blk.bn2, True, groups=groups)
    identity = x if blk.downsample is None else _bn_train(_conv(x, blk.downsample[0]), blk.downsample[1], False, groups=groups)
    return _bn_train(_conv(o, blk.conv3), blk.bn3, True, residual=identity, groups=groups)


def _instance_norm(x, eps):
    mean = x.mean(dim=(2, 3), keepdim=True)
    var = (x * x).mean(dim=(2, 3), keepdim=True) - mean * mean
    return r16((x - mean) / torch.sqrt(var.clamp_min(0.0) + eps))


def _head(head, fin, groups):
    """PPM head as regda_b200/ops/ppm_fold.py evaluates it: the 3x3 fuse convolution over cat(fin, up(p_1..4)) split by linearity
    into conv3x3(fin, W[:, :2048]) + sum_(cell, tap) A[px][(cell, tap)] G[(cell, tap)][o], G_k = p_k . W_k (bf16), A = the bilinear
    interpolation weights of the pyramid cells at the tap-shifted pixels (bf16), the second term rounded to bf16 before it joins
    the accumulator.  With ROUND off this is exactly cat + conv (the reference, Encoder.py:43-52)."""
    b, cf, h, w = fin.shape
    conv = head.conv_last[0]
    W = r16(conv.weight.detach())                                   # [O, 4096, 3, 3]
    O = W.shape[0]
    yppm = torch.zeros(b, O, h, w, dtype=torch.float64)
    off = cf
    for br in head.ppm:
        pooled = r16(br[0](fin))                                   # AdaptiveAvgPool2d in exact arithmetic, then the bf16 cast
        p = _bn_train(_conv(pooled, br[1]), br[2], True, groups=groups)          # [b, 512, s, s]
        s2, cb = p.shape[2] * p.shape[3], p.shape[1]
        Wk = W[:, off:off + cb]                                                  # [O, cb, 3, 3]
        off += cb
        G = r16(torch.einsum("bcj,ocrs->bjrso", p.reshape(b, cb, s2), Wk))       # [b, cell, r, s, O]
        eye = torch.eye(s2, dtype=torch.float64).view(s2, 1, p.shape[2], p.shape[3])
        up = F.interpolate(eye, (h, w), mode="bilinear", align_corners=False)[:, 0]          # B[cell][y][x]
        pad = F.pad(up, (1, 1, 1, 1))
        for r in range(3):
            for c in range(3):
                A = r16(pad[:, r:r + h, c:c + w])                                # basis weight at the tap-shifted pixel (zero padded)
                yppm = yppm + torch.einsum("jyx,bjo->boyx", A, G[:, :, r, c])
    yppm = r16(yppm)
    main = F.conv2d(fin, W[:, :cf], None, 1, 1, 1)
    y = _bn_train(r16(main + yppm), head.conv_last[1], True, groups=groups)
    cls = head.conv_last[4]                                        # Dropout2d is the identity in the parity runs (p = 0)
    return F.conv2d(y, cls.weight.detach().double(), cls.bias.detach().double())


@torch.no_grad()
def forward_train(model, x, groups=1, taps=None):
    """train-mode forward of step_oracle.DeeplabOracle as the bf16 kernels compute it: returns (x1, x2, feat) float32.
    groups = 2 restates Deeplabv2.forward_pair (source and target batch in one tensor, BatchNorm statistics per domain)."""
    rn = model.encoder.resnet
    t = r16(x.double())
    t = _bn_train(_conv(t, rn.conv1), rn.bn1, True, groups=groups)
    if taps is not None:
        taps["stem"] = t.float()
    t = F.max_pool2d(t, 3, 2, 1)
    for li, layer in enumerate((rn.layer1, rn.layer2, rn.layer3, rn.layer4), start=1):
        for bi, blk in enumerate(layer):
            t = _bottleneck(blk, t, groups)
            if taps is not None:
                taps[f"layer{li}.{bi}"] = t.float()
    fin = _instance_norm(t, model.instance_norm.eps)
    x1 = _head(model.layer5, fin, groups)
    x2 = _head(model.layer6, fin, groups)
    if taps is not None:
        taps.update(fin=fin.float(), x1=x1.float(), x2=x2.float())
    return x1.float(), x2.float(), fin.float()
