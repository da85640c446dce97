"""Deeplabv2 (ResNet OS16 + InstanceNorm + two PPM heads) -- drop-in for
regda/models/Encoder.py:87-165 with regda/resnet.py:43-66,140-207 and regda/_resnets.py:72-212,
for the configuration the self-training tools build (multi_layer=True, cascade=False,
use_ppm=True, is_ins_norm=True; tools/train_ssl_reg.py:94-111).

* state_dict keys and shapes are identical to the reference's (688 keys for ResNet-101), so
  stage-2 checkpoints load with strict=True and checkpoints written here load in tools/eval.py;
* train mode returns (x1, x2, feat) with x1/x2 float32 [b,C,h/16,w/16] and feat float32
  [b,2048,h/16,w/16] (channels-last memory); eval mode returns the averaged softmax at input
  resolution, as the reference does;
* activations flow channels-last in `compute_dtype` (bf16 by default, float32 for parity
  runs); every convolution goes through regda_b200.ops.conv.conv2d, which dispatches to the
  hand-written tcgen05 implicit-GEMM kernels for the shapes they cover.
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

import os

from ..ops.conv import Conv2d
from ..ops import conv as conv_ops
from ..ops import head as fhead
from ..ops import norm as fnorm
from ..ops import ppm as fppm
from ..ops import ppm_fold as ffold
from ..ops import stem as fstem

FUSED = os.environ.get("REGDA_FUSED", "1") != "0"     # hand-written BN / PPM kernels for bf16 training forward+backward


def set_fused(flag: bool):
    global FUSED
    FUSED = bool(flag)


def _fused(x, bn):
    return FUSED and fnorm.supported(x, bn)


# number of BatchNorm statistics groups of the forward pass in flight: 1 for a plain model(x) call, 2 inside
# Deeplabv2.forward_pair (source and target batch concatenated; see ops/norm.py)
_GROUPS = 1


def _inference(x):
    """eval-mode, no-grad, bf16 on the GPU: the hand-written inference kernels apply (offline teacher pass / evaluate())"""
    return FUSED and not torch.is_grad_enabled() and x.is_cuda and x.dtype == torch.bfloat16


def _conv_bn(conv, bn, x, residual=None, relu=False, tap=False):
    """relu?(bn(conv(x)) + residual); on the hand-written path the BatchNorm statistics come out of the convolution
    kernel's epilogue, so BatchNorm makes one pass (apply) instead of two over the activation.  tap=True returns
    (out, x_tap): x_tap is x for a residual branch, routed so that its gradient is added inside the dgrad kernel."""
    if not bn.training and _inference(x):
        y = conv(x)
        if fnorm.inference_supported(y, bn):
            out = fnorm.bn_inference(y, bn, residual, relu)
        else:
            out = _bn(y, bn, residual, relu)
        return (out, x) if tap else out
    if FUSED and bn.training and x.is_cuda and x.dtype == torch.bfloat16:
        res = conv.forward_with_bn_stats(x, _GROUPS, tap)
        y, st = res[0], res[1]
        if st is not None and fnorm.supported(y, bn):
            out = fnorm.bn_act(y, bn, residual=residual, relu=relu, groups=_GROUPS, stats=st)
        else:
            out = _bn(y, bn, residual, relu)
        return (out, res[2]) if tap else out
    out = _bn(conv(x), bn, residual, relu)
    return (out, x) if tap else out


def _bn(x, bn, residual=None, relu=False):
    """relu?(bn(x) + residual) through the fused kernels when they apply, else through torch; honours _GROUPS"""
    if _fused(x, bn):
        return fnorm.bn_act(x, bn, residual=residual, relu=relu, groups=_GROUPS)
    if FUSED and fnorm.inference_supported(x, bn):
        return fnorm.bn_inference(x, bn, residual, relu)
    out = fnorm.bn_eager(x, bn, _GROUPS)
    if residual is not None:
        out = out + residual
    return F.relu(out, inplace=True) if relu else out

_DEPTHS = {"resnet50": (3, 4, 6, 3), "resnet101": (3, 4, 23, 3)}


class AttrDict(dict):
    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = v


def _merge(dst, src):
    for k, v in src.items():
        if isinstance(v, dict):
            node = dst.get(k)
            if not isinstance(node, dict):
                node = AttrDict()
            dst[k] = _merge(AttrDict(node), v)
        else:
            dst[k] = v
    return dst


class Bottleneck(nn.Module):
    """_resnets.py:72-112 (stride on the 3x3, torchvision v1.5 style)."""
    expansion = 4

    def __init__(self, inplanes, planes, stride=1, dilation=1, downsample=False):
        super().__init__()
        self.conv1 = Conv2d(inplanes, planes, 1, bias=False)
        self.bn1 = nn.BatchNorm2d(planes)
        self.conv2 = Conv2d(planes, planes, 3, stride=stride, padding=dilation, dilation=dilation, bias=False)
        self.bn2 = nn.BatchNorm2d(planes)
        self.conv3 = Conv2d(planes, planes * 4, 1, bias=False)
        self.bn3 = nn.BatchNorm2d(planes * 4)
        self.downsample = None
        if downsample:
            self.downsample = nn.Sequential(Conv2d(inplanes, planes * 4, 1, stride=stride, bias=False), nn.BatchNorm2d(planes * 4))

    def forward(self, x):
        if (FUSED and self.training and x.is_cuda and x.dtype == torch.bfloat16) or _GROUPS > 1 or (not self.training and _inference(x)):
            # fnorm.arm(t, k): t (a BatchNorm+ReLU output) has exactly k consumers, all of them the Conv2d calls below, so
            # their data-gradient epilogues may carry the first half of that BatchNorm's backward (ops/norm.py BnHandle)
            # x is read twice (conv1 and the residual branch): the second reader goes through the tap conv1 hands back, so that
            # its gradient arrives in conv1's backward and is added inside the data-gradient epilogue instead of an autograd add
            fnorm.arm(x, 1)
            out, x_tap = _conv_bn(self.conv1, self.bn1, x, relu=True, tap=True)
            identity = x_tap if self.downsample is None else _conv_bn(self.downsample[0], self.downsample[1], x_tap)
            fnorm.arm(out, 1)
            out = _conv_bn(self.conv2, self.bn2, out, relu=True)
            fnorm.arm(out, 1)
            return _conv_bn(self.conv3, self.bn3, out, residual=identity, relu=True)
        y = self.conv1(x)
        out = F.relu(self.bn1(y), inplace=True)
        out = F.relu(self.bn2(self.conv2(out)), inplace=True)
        out = self.bn3(self.conv3(out))
        identity = x if self.downsample is None else self.downsample(x)
        out += identity
        return F.relu(out, inplace=True)


class ResNet(nn.Module):
    """_resnets.py:115-212 without avgpool/fc, after the output-stride-16 surgery of
    resnet.py:62-63,192-207: layer4's stride-2 convs run at stride 1 (3x3: dilation 1) and the
    other 3x3 convs of layer4 get dilation 2."""

    def __init__(self, resnet_type="resnet101"):
        super().__init__()
        self.conv1 = Conv2d(3, 64, 7, stride=2, padding=3, bias=False)
        self.bn1 = nn.BatchNorm2d(64)
        inplanes = 64
        for li, (planes, nblk) in enumerate(zip((64, 128, 256, 512), _DEPTHS[resnet_type]), start=1):
            blocks = []
            for bi in range(nblk):
                stride = 2 if (bi == 0 and li in (2, 3)) else 1
                dil = 2 if (li == 4 and bi > 0) else 1
                blocks.append(Bottleneck(inplanes, planes, stride, dil, downsample=(bi == 0)))
                inplanes = planes * 4
            setattr(self, f"layer{li}", nn.Sequential(*blocks))

    def forward(self, x, compute_dtype=None):
        """x: the image in the compute dtype, or -- compute_dtype = bf16 -- still the loader's float32 NCHW tensor (the stem's
        patch kernel rounds it on the way, saving the layout / dtype conversion pass)"""
        cdt = compute_dtype or x.dtype
        if conv_ops.ENGINE != "cudnn" and fstem.supported(self.conv1, x):
            # the 3-channel stem as an explicit patch matrix + 1x1 tcgen05 convolution (float32 parity mode: split-operand patch
            # matrices on the same kernels); in bf16 training the BatchNorm statistics come from its epilogue
            stats = FUSED and self.training and cdt == torch.bfloat16 and fnorm.supported(x.new_empty((1, 64, 1, 1), dtype=cdt), self.bn1)
            y, st = fstem.stem_conv(x, self.conv1.weight, _GROUPS if stats else None, cdt)
            x = fnorm.bn_act(y, self.bn1, relu=True, groups=_GROUPS, stats=st) if stats else _bn(y, self.bn1, relu=True)
        else:
            x = self.conv1(x)             # the library baseline (conv_ops.ENGINE == "cudnn"); anything else raises in Conv2d
            x = _bn(x, self.bn1, relu=True)
        x = fnorm.max_pool3s2(x) if FUSED else F.max_pool2d(x, 3, 2, 1)
        return self.layer4(self.layer3(self.layer2(self.layer1(x))))


class ResNetEncoder(nn.Module):
    """resnet.py:43-207: only the pieces that hold parameters / shape the forward."""

    def __init__(self, config):
        super().__init__()
        rt = str(config.get("resnet_type", "resnet50")).lower()
        if rt == "resnet":
            rt = "resnet50"
        if rt not in _DEPTHS:
            raise ValueError(f"unsupported resnet_type {rt!r} (resnet50 / resnet101)")
        if int(config.get("output_stride", 16)) != 16:
            raise ValueError("only output_stride=16 is on the self-training path")
        self.resnet = ResNet(rt)

    def forward(self, x, compute_dtype=None):
        return self.resnet(x, compute_dtype)


class PPMBilinear(nn.Module):
    """Encoder.py:8-65 (use_aux=False)."""

    def __init__(self, num_classes=7, fc_dim=2048, use_aux=False, pool_scales=(1, 2, 3, 6), dropout=0.1):
        super().__init__()
        if use_aux:
            raise NotImplementedError("use_aux=True is not used by the self-training tools")
        self.pool_scales = tuple(pool_scales)
        self.ppm = nn.ModuleList([
            nn.Sequential(nn.AdaptiveAvgPool2d(s), Conv2d(fc_dim, 512, 1, bias=False), nn.BatchNorm2d(512), nn.ReLU(inplace=True))
            for s in pool_scales])
        self.conv_last = nn.Sequential(
            Conv2d(fc_dim + len(pool_scales) * 512, 512, 3, padding=1, bias=False), nn.BatchNorm2d(512), nn.ReLU(inplace=True),
            nn.Dropout2d(dropout), Conv2d(512, num_classes, 1, bias=True))

    def fused_ok(self, conv_out):
        return (FUSED and (self.training or not torch.is_grad_enabled()) and conv_out.is_cuda and conv_out.dtype == torch.bfloat16
                and len(self.pool_scales) <= 4
                and conv_out.shape[1] % 8 == 0 and self.ppm[0][1].out_channels % 8 == 0)

    def forward(self, conv_out, pooled=None, tap=False):
        """tap=True (folded fuse convolution only) returns (logits, conv_out_tap): the handle the NEXT reader of the feature map
        should use, so that its gradient is added inside this head's data-gradient kernel (ops/ppm_fold.py)"""
        if tap:
            assert self.fused_ok(conv_out) and pooled is not None
        if self.fused_ok(conv_out):
            # one-pass pyramid pooling (shared by both heads when the caller passes `pooled`), tiny per-branch
            # 1x1 conv + BN + ReLU on the pooled maps, fused bilinear-upsample + concat, tcgen05 3x3 conv.
            if pooled is None:
                pooled = fppm.pool(conv_out, self.pool_scales)
            # the s x s pooled maps as bf16 NHWC tensors (one launch; shared by both heads when the caller passes them as a list)
            cells = pooled if isinstance(pooled, (list, tuple)) else fppm.cells(pooled, self.pool_scales)
            # 1x1 conv on the s x s map (several images per tcgen05 M tile) + BatchNorm + ReLU, all hand-written kernels
            branches = [_conv_bn(branch[1], branch[2], p, relu=True) for p, branch in zip(cells, self.ppm)]
            conv, bn = self.conv_last[0], self.conv_last[1]
            if ffold.supported(conv_out, branches, conv, self.pool_scales) and (bn.training or fnorm.inference_supported(conv_out, bn)):
                # the upsampled branches enter the 3x3 convolution through two small GEMMs instead of 2048 materialised channels
                # (ops/ppm_fold.py): half the reference's K, no concatenated tensor
                train = bn.training and fnorm.supported(conv_out.new_empty((1, conv.out_channels, 1, 1)), bn)
                res = ffold.fuse_conv(conv_out, branches, conv.weight, self.pool_scales, _GROUPS if train else None, tap)
                y, st = res[0], res[1]
                y = fnorm.bn_act(y, bn, relu=True, groups=_GROUPS, stats=st) if train else _bn(y, bn, relu=True)
                if tap:
                    return self._classify(y), res[2]
            else:
                cat = fppm.upsample_concat(conv_out, branches, self.pool_scales)
                y = _conv_bn(conv, bn, cat, relu=True)
            return (self._classify(y), conv_out) if tap else self._classify(y)
        size = conv_out.shape[-2:]
        outs = [conv_out]
        for branch in self.ppm:
            t = branch[3](fnorm.bn_eager(branch[1](branch[0](conv_out)), branch[2], _GROUPS))
            outs.append(F.interpolate(t, size, mode="bilinear", align_corners=False))
        y = self.conv_last[0](torch.cat(outs, 1))
        y = self.conv_last[2](fnorm.bn_eager(y, self.conv_last[1], _GROUPS))
        return self._classify(y)

    def _classify(self, y):
        """Dropout2d -> 1x1 conv + bias (Encoder.py:39-40): one streaming kernel (ops/head.py), float32 logits out"""
        drop, cls = self.conv_last[3], self.conv_last[4]
        if conv_ops.ENGINE != "cudnn" and fhead.supported(y, cls.weight):
            return fhead.dropout_classifier(y, cls.weight, cls.bias, drop.p, self.training)
        return cls(drop(y))


class Deeplabv2(nn.Module):
    def __init__(self, config, compute_dtype=torch.bfloat16):
        super().__init__()
        cfg = AttrDict(backbone=AttrDict(resnet_type="resnet50", output_stride=16, pretrained=False), multi_layer=False,
                       cascade=False, use_ppm=False, ppm=AttrDict(num_classes=7, use_aux=False, fc_dim=2048),
                       inchannels=2048, num_classes=7, is_ins_norm=False)
        self._cfg = _merge(cfg, dict(config))
        c = self._cfg
        if not (c.multi_layer and not c.cascade and c.use_ppm):
            raise NotImplementedError("regda_b200.Deeplabv2 implements the multi_layer / no-cascade / PPM form "
                                      "that tools/train_ssl_reg.py:94-111 builds")
        self.compute_dtype = compute_dtype
        self.encoder = ResNetEncoder(c.backbone)
        self.layer5 = PPMBilinear(**c.ppm)
        self.layer6 = PPMBilinear(**c.ppm)
        if c.is_ins_norm:
            self.instance_norm = nn.InstanceNorm2d(c.inchannels)
        self.to(memory_format=torch.channels_last)

    @property
    def config(self):
        return self._cfg

    def forward_pair(self, x_s, x_t, feat_dtype=None):
        """Train-mode forward of the source and the target batch as ONE tensor (tools/train_ssl_reg.py:210-212 makes two
        calls): every convolution sees twice the rows, BatchNorm keeps one statistics group per domain, so the results
        are those of two separate calls.  Returns ((x1_s, x2_s, feat_s), (x1_t, x2_t, feat_t))."""
        global _GROUPS
        assert self.training and x_s.shape == x_t.shape
        b = x_s.shape[0]
        _GROUPS = 2
        try:
            x1, x2, feat = self.forward(torch.cat([x_s, x_t], 0), feat_dtype=feat_dtype)
        finally:
            _GROUPS = 1
        return (x1[:b], x2[:b], feat[:b]), (x1[b:], x2[b:], feat[b:])

    def forward(self, x, feat_dtype=None):
        """feat_dtype: dtype of the returned feature map in train mode (default float32, as the reference returns it; the trainer
        asks for the bf16 tensor the InstanceNorm kernel wrote -- its Aligner kernels read bf16 rows -- which saves a 200 MB
        float32 copy per step)"""
        if (self.compute_dtype == torch.bfloat16 and x.dtype == torch.float32 and x.is_cuda and conv_ops.ENGINE != "cudnn"
                and fstem.supported(self.encoder.resnet.conv1, x)):
            feat = self.encoder(x, torch.bfloat16)           # the stem reads the float32 NCHW image directly
        else:
            xin = x.to(self.compute_dtype).contiguous(memory_format=torch.channels_last)
            feat = self.encoder(xin)
        if self._cfg.is_ins_norm and FUSED and (self.training or not torch.is_grad_enabled()) and fnorm.instance_norm_supported(feat):
            # hand-written path: per-image statistics groups of the BatchNorm kernels, bf16 in / bf16 out; the Aligner's
            # float32 feature view is one conversion of the result
            fin = fnorm.instance_norm(feat, self.instance_norm.eps)
            feat = None                                       # made from fin (or from its last tap) after the heads
        else:
            if self._cfg.is_ins_norm:
                feat = self.instance_norm(feat.float())      # float32 statistics and output (feeds the Aligner)
            else:
                feat = feat.float()
            fin = feat.to(self.compute_dtype)
        if self.layer5.fused_ok(fin) and self.layer5.pool_scales == self.layer6.pool_scales:
            if self.training and feat is None and torch.is_grad_enabled():
                # the feature map has four readers (pooling, the two heads' fuse convolutions, the caller's Aligner losses): chain
                # them through taps -- pooling <- head 5 <- head 6 <- caller -- so that every reader's gradient is added inside the
                # next one's backward kernel instead of by three autograd adds over the 67 MB map
                pooled, fin5 = fppm.pool(fin, self.layer5.pool_scales, tap=True)
                pooled = fppm.cells(pooled, self.layer5.pool_scales)
                x1, fin6 = self.layer5(fin5, pooled, tap=True)
                x2, fin = self.layer6(fin6, pooled, tap=True)
                x1, x2 = x1.float(), x2.float()
            else:
                pooled = fppm.cells(fppm.pool(fin, self.layer5.pool_scales), self.layer5.pool_scales)    # both heads pool the same feature map
                x1 = self.layer5(fin, pooled).float()
                x2 = self.layer6(fin, pooled).float()
        else:
            x1 = self.layer5(fin).float()
            x2 = self.layer6(fin).float()
        if feat is None:
            # (the float32 copy is made on the dense NHWC view: a contiguous, vectorised cast instead of a strided one)
            feat = fin if (feat_dtype == torch.bfloat16 and self.training) else fin.permute(0, 2, 3, 1).float().permute(0, 3, 1, 2)
        if self.training:
            return x1, x2, feat
        if FUSED and x1.is_cuda and not torch.is_grad_enabled() and x1.shape[1] <= 16:
            return fppm.upsample_softmax_mean(x1, x2, x.shape[-2:])       # one pass, no [b,c,H,W] intermediates
        x1 = F.interpolate(x1, x.shape[-2:], mode="bilinear", align_corners=True)
        x2 = F.interpolate(x2, x.shape[-2:], mode="bilinear", align_corners=True)
        return (x1.softmax(dim=1) + x2.softmax(dim=1)) / 2
