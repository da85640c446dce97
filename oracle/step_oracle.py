"""TEST INFRASTRUCTURE ONLY -- plain PyTorch fp32 (CPU) restatement of the floating-point
stages of RegDA's self-training inner step and of the whole step.

Used as (a) the checker for the CUDA kernels in tests/ and smoke(), (b) the `port` CPU
baseline / `--impl reference` arm of bench.py (the reference is Python and does not travel
to the GPU box).  Never imported by regda_b200/.

Parity status: PINNED against the unmodified reference run in the build container
(tests/golden/*.npz via tests/golden/make_golden.py; tests/test_oracle_golden.py) and, when
/root/reference is present, against the live reference (tests/test_oracle_vs_reference.py).

Each function cites the reference lines it restates.
"""
from __future__ import annotations

import math
from collections import OrderedDict

import torch
import torch.nn as nn
import torch.nn.functional as F

EPS = 1e-7  # Aligner.eps, alignment.py:43


# --------------------------------------------------------------------------------------
# integer stages in torch (slow-but-simple twins of oracle/regda_oracle.c; used where a
# tensor-in/tensor-out helper is handier than numpy)
# --------------------------------------------------------------------------------------
def lrh(labels, regions, class_num, ignore_label, percent):
    """local_region_homog.py:125-152 through the C oracle."""
    from . import cbind
    return torch.from_numpy(cbind.lrh(labels.cpu().numpy(), regions.cpu().numpy(), class_num, ignore_label, percent))


def lrh_torch(labels, regions, class_num, ignore_label, percent):
    """local_region_homog.py:107-152 restated op-for-op in torch (int64 one-hot counts,
    scatter_add_ per image, float32 ratio test, gather, where).  This is the multi-threaded
    CPU arm bench.py times as `--impl reference` (the C function above is the scalar checker);
    tests/test_oracle_golden.py holds it to the same golden vectors."""
    b = labels.shape[0]
    lab = labels.reshape(b, -1)
    reg = regions.reshape(b, -1)
    code = torch.where(lab == ignore_label, torch.full_like(lab, class_num), lab)          # :118
    onehot = F.one_hot(code, class_num + 1)[..., :class_num]                               # :121
    n_reg = int(reg.max()) + 1                                                             # scatter's index.max()+1
    counts = torch.zeros(b, n_reg, class_num, dtype=torch.int64)
    counts.scatter_add_(1, reg.unsqueeze(-1).expand(-1, -1, class_num), onehot)            # :140
    valid = counts.sum(-1)                                                                 # :141
    best, arg = counts.max(-1)                                                             # :142 (first max)
    ratio = best / (valid + 1e-5)                                                          # :143 float32
    arg = torch.where(ratio < percent, torch.full_like(arg, ignore_label), arg)            # :144
    out = torch.gather(arg, 1, reg)                                                        # :147
    out = torch.where(reg == 0, torch.full_like(out, ignore_label), out)                   # :149
    return torch.where(out == ignore_label, lab, out).view_as(labels)                      # :151


def pseudo_select(soft, cutoff_top=0.8, cutoff_low=0.6, ignore_label=-1):
    """pseudo_generation.py:59-93 through the C oracle."""
    from . import cbind
    return torch.from_numpy(cbind.pseudo_select(soft.detach().cpu().numpy(), cutoff_top, cutoff_low, ignore_label))


def downscale_label(label, scale=16, n_classes=6, ignore_label=-1, min_ratio=0.75):
    """alignment.py:466-481 through the C oracle; returns [b,1,h,w] int64."""
    from . import cbind
    return torch.from_numpy(cbind.downscale_label(label.cpu().numpy(), scale, n_classes, ignore_label, min_ratio))


# --------------------------------------------------------------------------------------
# Aligner arithmetic
# --------------------------------------------------------------------------------------
def pearson_dist(rows, protos):
    """alignment.py:396-423.  rows [n,k], protos [m,k] -> [n,m] in [0,1].

    dist = (1 - cov/(k-1+eps) / (std_row*std_proto + eps)) / 2, unbiased std.
    (Evaluated as a matmul of centred rows instead of the reference's [n,m,k] broadcast:
    same sums, different association order -- float tolerance applies.)"""
    k = rows.shape[-1]
    rc = rows - rows.mean(dim=-1, keepdim=True)
    pc = protos - protos.mean(dim=-1, keepdim=True)
    cov = rc @ pc.t() / (k - 1 + EPS)
    denom = rows.std(dim=-1).unsqueeze(1) * protos.std(dim=-1).unsqueeze(0) + EPS
    return ((1.0 - cov / denom) * 0.5).detach()


def _peak_normalised_softmax(x, temp):
    p = torch.softmax(x / temp, dim=1)
    return p / (p.max(dim=1, keepdim=True)[0] + 1e-7)


def label_refine(feat_t, preds_t, soft, prototypes, temp=2.0):
    """alignment.py:194-265 with label_t_sup=None, refine=True, mode='all'.

    feat_t [b,k,h,w]; preds_t = [p1,p2] each [b,c,h,w]; soft [b,c,H,W]; prototypes [c,k]."""
    b, k, h, w = feat_t.shape
    H, W = soft.shape[-2:]
    rows = feat_t.detach().permute(0, 2, 3, 1).reshape(-1, k)
    simi = 1.0 / pearson_dist(rows, prototypes)                                   # :216
    simi = simi.view(b, h, w, -1).permute(0, 3, 1, 2)
    simi = F.interpolate(simi, (H, W), mode="bilinear", align_corners=True)        # :218-219
    weight = _peak_normalised_softmax(simi, 1.0)                                   # :220-222
    ups = [F.interpolate(p.detach(), (H, W), mode="bilinear", align_corners=True) for p in preds_t]
    pred_w = sum(torch.softmax(u / temp, dim=1) for u in ups) * (1.0 / len(ups))   # :228-231
    weight = weight + pred_w / (pred_w.max(dim=1, keepdim=True)[0] + 1e-7)         # :235-236
    out = weight * soft                                                            # :263
    return out / (out.sum(dim=1, keepdim=True) + EPS)                              # :264, 288-298


def class_sums(feat, label_ds, class_num):
    """Per-class feature sums and counts (the segmented reduction inside
    alignment.py:300-327 / :107-119).  feat [b,k,h,w]; label_ds [b,1,h,w] int64."""
    b, k, h, w = feat.shape
    rows = feat.detach().permute(0, 2, 3, 1).reshape(-1, k)
    lab = label_ds.reshape(-1)
    sums = torch.zeros(class_num, k, dtype=rows.dtype)
    cnt = torch.zeros(class_num, dtype=rows.dtype)
    for c in range(class_num):
        m = lab == c
        cnt[c] = m.sum()
        if m.any():
            sums[c] = rows[m].sum(0)
    return sums, cnt


def prototype_ema(prototypes, sums, cnt, decay):
    """alignment.py:319-325, 435-438: local = sum/(n+eps); classes with n<1 keep the old
    prototype; new = (1-decay)*local + decay*old."""
    local = sums / (cnt.unsqueeze(1) + EPS)
    local = torch.where(cnt.unsqueeze(1) < 1, prototypes, local)
    return (1.0 - decay) * local + decay * prototypes


def update_prototype(prototypes, feat_s, label_s, class_num, ignore_label, decay):
    """alignment.py:86-90."""
    lab = downscale_label(label_s, 16, class_num, ignore_label, 0.75)
    sums, cnt = class_sums(feat_s, lab, class_num)
    return prototype_ema(prototypes, sums, cnt, decay), lab


# --------------------------------------------------------------------------------------
# losses / optimiser
# --------------------------------------------------------------------------------------
def ce_loss_multi(preds, label, ignore_label=-1):
    """tools.py:240-252 + balance.py:88-101: for each head bilinear-upsample
    (align_corners=True) to the label size, per-pixel CE with ignore_index, **mean over all
    pixels including ignored ones**, then the mean over heads."""
    total = 0.0
    for p in preds:
        if p.shape[-2:] != label.shape[-2:]:
            p = F.interpolate(p, size=label.shape[-2:], mode="bilinear", align_corners=True)
        total = total + F.cross_entropy(p, label.long(), ignore_index=ignore_label, reduction="none").mean()
    return total / len(preds)


def learning_rate(i_iter, base_lr, warmup_iters, num_steps, power):
    """tools.py:191-207: linear warm-up then poly decay."""
    if i_iter < warmup_iters:
        return base_lr * (float(i_iter) / warmup_iters)
    return base_lr * ((1 - float(i_iter) / num_steps) ** power)


def class_balance_freq(freq, label, class_num, ignore_label, decay=0.99):
    """balance.py:35-53: EMA of the per-batch class frequency."""
    from . import cbind
    counts, n_valid = cbind.class_count(label.cpu().numpy(), class_num, ignore_label)
    local = torch.from_numpy(counts).float() / (float(n_valid) + 1e-7)
    return (1.0 - decay) * local + decay * freq


def class_balance_weights(freq, temperature):
    """balance.py:38-43."""
    p = torch.softmax((1.0 - freq) / temperature, dim=0)
    return p / (p.max() + 1e-7)


# --------------------------------------------------------------------------------------
# model: ResNet (bottleneck) OS16 + InstanceNorm + two PPM heads.  Same state_dict keys
# as the reference (Encoder.py:87-165, resnet.py:43-66,192-207, _resnets.py:72-212).
# --------------------------------------------------------------------------------------
_DEPTHS = {"resnet50": (3, 4, 6, 3), "resnet101": (3, 4, 23, 3)}


def conv_plan(resnet_type="resnet101"):
    """List of (name, cin, cout, k, stride, dilation, has_downsample) for the encoder
    bottlenecks after the output-stride-16 surgery of resnet.py:62-63,192-207: the stride-2
    convs of layer4 become stride 1 (3x3: dilation 1), every other 3x3 in layer4 gets
    dilation 2."""
    plan = []
    inplanes = 64
    for li, (planes, nblk) in enumerate(zip((64, 128, 256, 512), _DEPTHS[resnet_type]), start=1):
        for bi in range(nblk):
            stride = 2 if (bi == 0 and li in (2, 3)) else 1
            dil = 2 if (li == 4 and bi > 0) else 1
            plan.append(dict(name=f"layer{li}.{bi}", cin=inplanes, mid=planes, cout=planes * 4,
                             stride=stride, dilation=dil, downsample=(bi == 0)))
            inplanes = planes * 4
    return plan


class _Bottleneck(nn.Module):
    def __init__(self, cin, mid, cout, stride, dilation, downsample):
        super().__init__()
        self.conv1 = nn.Conv2d(cin, mid, 1, bias=False)
        self.bn1 = nn.BatchNorm2d(mid)
        self.conv2 = nn.Conv2d(mid, mid, 3, stride=stride, padding=dilation, dilation=dilation, bias=False)
        self.bn2 = nn.BatchNorm2d(mid)
        self.conv3 = nn.Conv2d(mid, cout, 1, bias=False)
        self.bn3 = nn.BatchNorm2d(cout)
        self.downsample = None
        if downsample:
            self.downsample = nn.Sequential(nn.Conv2d(cin, cout, 1, stride=stride, bias=False), nn.BatchNorm2d(cout))

    def forward(self, x):
        y = F.relu(self.bn1(self.conv1(x)))
        y = F.relu(self.bn2(self.conv2(y)))
        y = self.bn3(self.conv3(y))
        sc = x if self.downsample is None else self.downsample(x)
        return F.relu(y + sc)


class _ResNet(nn.Module):
    def __init__(self, resnet_type):
        super().__init__()
        self.conv1 = nn.Conv2d(3, 64, 7, stride=2, padding=3, bias=False)
        self.bn1 = nn.BatchNorm2d(64)
        layers = {1: [], 2: [], 3: [], 4: []}
        for blk in conv_plan(resnet_type):
            li = int(blk["name"][5])
            layers[li].append(_Bottleneck(blk["cin"], blk["mid"], blk["cout"], blk["stride"], blk["dilation"], blk["downsample"]))
        for li in range(1, 5):
            setattr(self, f"layer{li}", nn.Sequential(*layers[li]))

    def forward(self, x):
        x = F.relu(self.bn1(self.conv1(x)))
        x = F.max_pool2d(x, 3, 2, 1)
        return self.layer4(self.layer3(self.layer2(self.layer1(x))))


class _Encoder(nn.Module):
    def __init__(self, resnet_type):
        super().__init__()
        self.resnet = _ResNet(resnet_type)

    def forward(self, x):
        return self.resnet(x)


class _PPMHead(nn.Module):
    """Encoder.py:8-65: pyramid pooling (1,2,3,6) + 3x3 fuse conv + classifier."""

    def __init__(self, num_classes, fc_dim=2048, scales=(1, 2, 3, 6), dropout=0.1):
        super().__init__()
        self.scales = scales
        self.ppm = nn.ModuleList([
            nn.Sequential(nn.AdaptiveAvgPool2d(s), nn.Conv2d(fc_dim, 512, 1, bias=False), nn.BatchNorm2d(512), nn.ReLU())
            for s in scales])
        self.conv_last = nn.Sequential(
            nn.Conv2d(fc_dim + 512 * len(scales), 512, 3, padding=1, bias=False), nn.BatchNorm2d(512), nn.ReLU(),
            nn.Dropout2d(dropout), nn.Conv2d(512, num_classes, 1))

    def forward(self, f):
        size = f.shape[-2:]
        cat = [f] + [F.interpolate(br(f), size, mode="bilinear", align_corners=False) for br in self.ppm]
        return self.conv_last(torch.cat(cat, 1))


class DeeplabOracle(nn.Module):
    """Encoder.py:129-155 for multi_layer=True, cascade=False, use_ppm=True, is_ins_norm=True."""

    def __init__(self, resnet_type="resnet101", num_classes=6, dropout=0.1):
        super().__init__()
        self.encoder = _Encoder(resnet_type)
        self.layer5 = _PPMHead(num_classes, dropout=dropout)
        self.layer6 = _PPMHead(num_classes, dropout=dropout)
        self.instance_norm = nn.InstanceNorm2d(2048)

    def forward(self, x):
        feat = self.instance_norm(self.encoder(x))
        x1, x2 = self.layer5(feat), self.layer6(feat)
        if self.training:
            return x1, x2, feat
        x1 = F.interpolate(x1, x.shape[-2:], mode="bilinear", align_corners=True)
        x2 = F.interpolate(x2, x.shape[-2:], mode="bilinear", align_corners=True)
        return (x1.softmax(dim=1) + x2.softmax(dim=1)) / 2


def seeded_state_dict(model_or_keys, seed=2333):
    """Deterministic weights keyed by parameter *name* (so the reference, the oracle and
    the CUDA model can be filled identically on any host with this torch build).
    conv weights ~ N(0, sqrt(2/fan_out)) (the kaiming fan_out rule of _resnets.py:163),
    BN gamma ~ U(0.5,1.5), beta ~ N(0,0.1), running_mean ~ N(0,0.1), running_var ~ U(0.5,1.5); the last BatchNorm of
    every bottleneck (bn3) gets gamma ~ U(0.05,0.25) -- the usual small-residual initialisation -- so that the
    random network is well conditioned: with O(1) residual gains and the tiny batch-statistic sample counts of the
    64x64 / 96x96 fixtures, float32 re-association noise is amplified to percents in the stem gradient, which says
    nothing about kernel parity."""
    ref = model_or_keys.state_dict() if isinstance(model_or_keys, nn.Module) else model_or_keys
    out = OrderedDict()
    for name, t in ref.items():
        h = 0
        for ch in name:
            h = (h * 131 + ord(ch)) % 2147483647
        g = torch.Generator().manual_seed(seed * 7919 + h)
        if name.endswith("num_batches_tracked"):
            out[name] = torch.zeros_like(t)
        elif t.dim() == 4:
            fan_out = t.shape[0] * t.shape[2] * t.shape[3]
            out[name] = torch.randn(t.shape, generator=g) * math.sqrt(2.0 / fan_out)
        elif name.endswith("bn3.weight"):
            out[name] = torch.rand(t.shape, generator=g) * 0.2 + 0.05
        elif name.endswith("running_var") or (name.endswith("weight") and t.dim() == 1):
            out[name] = torch.rand(t.shape, generator=g) + 0.5
        else:
            out[name] = torch.randn(t.shape, generator=g) * 0.1
    return out


# --------------------------------------------------------------------------------------
# the whole inner step, tools/train_ssl_reg.py:198-241
# --------------------------------------------------------------------------------------
class StepState:
    def __init__(self, model, prototypes, lr=1e-2, momentum=0.9, weight_decay=5e-4):
        self.model = model
        self.prototypes = prototypes
        self.opt = torch.optim.SGD(model.parameters(), lr=lr, momentum=momentum, weight_decay=weight_decay)


def inner_step(state, images_s, label_s, images_t, soft_t, regs_t, *, class_num=6, ignore_label=-1,
               percent=0.5, cutoff_top=0.8, cutoff_low=0.6, temp=2.0, decay=0.996, lr=None, max_norm=32.0,
               sam_refine=True):
    m = state.model
    m.train()
    if lr is not None:
        state.opt.param_groups[0]["lr"] = lr
    ps1, ps2, feat_s = m(images_s)                                                  # :210
    pt1, pt2, feat_t = m(images_t)                                                  # :212
    soft = label_refine(feat_t, [pt1, pt2], soft_t, state.prototypes, temp)         # :214
    hard = pseudo_select(soft, cutoff_top, cutoff_low, ignore_label)                # :218
    if sam_refine:
        hard = lrh(hard, regs_t.squeeze(1), class_num, ignore_label, percent)       # :223
    state.prototypes, _ = update_prototype(state.prototypes, feat_s, label_s, class_num, ignore_label, decay)  # :225
    loss_s = ce_loss_multi([ps1, ps2], label_s, ignore_label)                       # :228
    loss_t = ce_loss_multi([pt1, pt2], hard, ignore_label)                          # :233
    loss = loss_s + loss_t
    state.opt.zero_grad()
    loss.backward()                                                                 # :238
    gnorm = torch.nn.utils.clip_grad_norm_(m.parameters(), max_norm=max_norm, norm_type=2)  # :239
    state.opt.step()                                                                # :241
    return dict(loss=float(loss), loss_source=float(loss_s), loss_target=float(loss_t), grad_norm=float(gnorm),
                hard=hard, soft=soft)


# ---- offline teacher pass ---------------------------------------------------------------------------
# Test infrastructure like the rest of this file.  Restates regda/utils/tools.py:52-57 (pad_image), :61-97 (pre_slide) and
# :132-152 (tta_predict, through ttach==0.0.3's Compose([HorizontalFlip(), Rotate90([0, 90, 180, 270])])); pinned by
# tests/golden/teacher_pass.npz, which the reference's own functions produced (tests/golden/make_golden.py teacher_pass).
def tta_predict(model, img):
    """mean over the 8 views (flip in {no, yes}) x (k quarter turns, k = 0..3): view = rot90^k(flip(img)); the prediction is
    brought back by rot90^-k then flip.  One model call per view, as the reference does."""
    outs = []
    for flip in (False, True):
        for k in range(4):
            v = img.flip(3) if flip else img
            v = torch.rot90(v, k, (2, 3))
            o = torch.rot90(model(v), -k, (2, 3))
            outs.append(o.flip(3) if flip else o)
    return torch.cat(outs, 0).mean(dim=0, keepdim=True)


def window_origins(size, tile):
    """origins of the 50 %-overlap windows along one axis: 0, tile/2, ... with the last one shifted back inside"""
    stride = math.ceil(tile * 0.5)
    n = int(math.ceil((size - tile) / stride) + 1)          # tools.py:66-67 (1 when the image is smaller than the tile)
    return [max(min(i * stride + tile, size) - tile, 0) for i in range(n)]


def pre_slide(model, image, num_classes, tile=(512, 512), tta=False):
    b, _, H, W = image.shape
    prob = torch.zeros(b, num_classes, H, W)
    cnt = torch.zeros(b, 1, H, W)
    for y1 in window_origins(H, tile[0]):
        for x1 in window_origins(W, tile[1]):
            y2, x2 = min(y1 + tile[0], H), min(x1 + tile[1], W)
            win = image[:, :, y1:y2, x1:x2]
            # pad_image: F.pad(img, (0, 0, rows_missing, cols_missing)) = height padded by rows_missing on TOP and cols_missing
            # at the bottom, width untouched (the reference's argument order, tools.py:56)
            win_p = F.pad(win, (0, 0, tile[0] - win.shape[2], tile[1] - win.shape[3]))
            out = tta_predict(model, win_p) if tta else model(win_p)
            prob[:, :, y1:y2, x1:x2] += out[:, :, :win.shape[2], :win.shape[3]]
            cnt[:, :, y1:y2, x1:x2] += 1
    return prob / cnt


# --------------------------------------------------------------------------------------
# stage-2 step, tools/train_align_reg.py:144-196 with regda/loss.py:18-47
# --------------------------------------------------------------------------------------
def pcl_loss(prototypes, feat, labels, temperature=8.0, ignore_label=-1):
    """PrototypeContrastiveLoss.forward (regda/loss.py:24-47): cross entropy of the cosine similarities (divided by the
    temperature) between every non-ignored feature row and the class prototypes, mean over those rows."""
    k = feat.shape[1]
    rows = feat.permute(0, 2, 3, 1).reshape(-1, k) if feat.dim() == 4 else feat
    lab = labels.reshape(-1)
    keep = lab != ignore_label
    rows, lab = rows[keep], lab[keep]
    rows = rows / rows.norm(dim=1, keepdim=True).clamp_min(1e-12)
    pr = prototypes / prototypes.norm(dim=1, keepdim=True).clamp_min(1e-12)
    return F.cross_entropy(rows @ pr.t() / temperature, lab)


def coral_loss(source, target, is_sqrt=False):
    """CoralLoss.forward (regda/gast/coral.py:26-47): squared Frobenius distance of the two feature covariances / (4 d^2)"""
    d = source.shape[1]
    ns, nt = source.shape[0], target.shape[0]
    xm = torch.mean(source, 0, keepdim=True) - source
    xc = xm.t() @ xm / (ns - 1)
    xmt = torch.mean(target, 0, keepdim=True) - target
    xct = xmt.t() @ xmt / (nt - 1)
    loss = torch.sum((xc - xct) * (xc - xct))
    return (loss.sqrt() if is_sqrt else loss) / (4 * d * d)


def align_domain(feat_s, feat_t):
    """Aligner.align_domain (regda/gast/alignment.py:79-84)"""
    k = feat_s.shape[1]
    return coral_loss(feat_s.permute(0, 2, 3, 1).reshape(-1, k), feat_t.permute(0, 2, 3, 1).reshape(-1, k))


def align_step(state, images_s, label_s, images_t, regs_t, *, class_num=6, ignore_label=-1, percent=0.5, cutoff_top=0.8,
               cutoff_low=0.6, temp=2.0, decay=0.996, lr=None, max_norm=32.0, sam_refine=True, pcl_temp=8.0, use_coral=False):
    m = state.model
    m.train()
    if lr is not None:
        state.opt.param_groups[0]["lr"] = lr
    ps1, ps2, feat_s = m(images_s)                                                  # :155
    state.prototypes, label_s_down = update_prototype(state.prototypes, feat_s.detach(), label_s, class_num, ignore_label, decay)  # :158
    state.prototypes = state.prototypes.detach()
    pt1, pt2, feat_t = m(images_t)                                                  # :163
    size = images_t.shape[-2:]
    x1 = F.interpolate(pt1, size, mode="bilinear", align_corners=True)
    x2 = F.interpolate(pt2, size, mode="bilinear", align_corners=True)
    soft_t = ((x1.softmax(dim=1) + x2.softmax(dim=1)) * 0.5).detach()               # :165-167
    soft = label_refine(feat_t, [pt1, pt2], soft_t, state.prototypes, temp)         # :168
    hard = pseudo_select(soft, cutoff_top, cutoff_low, ignore_label)                # :170
    if sam_refine:
        hard = lrh(hard, regs_t.squeeze(1), class_num, ignore_label, percent)       # :176-178
    label_t = downscale_label(hard, 16, class_num, ignore_label)                    # :182
    loss_seg = ce_loss_multi([ps1, ps2], label_s, ignore_label)                     # :186
    loss_align = (pcl_loss(state.prototypes, feat_s, label_s_down, pcl_temp, ignore_label) +
                  pcl_loss(state.prototypes, feat_t, label_t, pcl_temp, ignore_label)) * 0.5   # :188-189
    loss_domain = align_domain(feat_s, feat_t) if use_coral else 0.0                # :187
    loss = loss_seg + loss_domain + loss_align
    state.opt.zero_grad()
    loss.backward()                                                                 # :193
    gnorm = torch.nn.utils.clip_grad_norm_(m.parameters(), max_norm=max_norm, norm_type=2)
    state.opt.step()
    return dict(loss=float(loss), loss_seg=float(loss_seg), loss_align=float(loss_align), loss_domain=float(loss_domain),
                grad_norm=float(gnorm), hard=hard, label_t=label_t)
