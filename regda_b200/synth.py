"""Seeded synthetic inputs of the shapes the self-training step consumes.

Host-side helper for bench.py / tests / smoke(); works on any torch device.  Seed 2333 is
the reference's seed_torch value (tools/train_ssl_reg.py:274).
"""
from __future__ import annotations

import math

import torch

SEED = 2333


def _gen(device, seed):
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    return g


def region_maps(b, h, w, n_regions, device="cpu", seed=SEED, bg_frac=0.15, id_span=None):
    """int64 [b,h,w] blocky partition with ~n_regions regions per tile, wavy boundaries, ids
    drawn sparsely from 1..id_span (default 2*n_regions) and ~bg_frac of the regions set to
    0 (SAM leaves small / unsegmented areas 0, local_region_homog.py:51-55)."""
    g = _gen(device, seed)
    id_span = id_span or 2 * n_regions
    gw = max(1, int(round(math.sqrt(n_regions * w / h))))
    gh = max(1, int(math.ceil(n_regions / gw)))
    ys = torch.arange(h, device=device, dtype=torch.float32).view(1, h, 1)
    xs = torch.arange(w, device=device, dtype=torch.float32).view(1, 1, w)
    ph = torch.rand(b, 4, device=device, generator=g).mul(6.283).view(b, 4, 1, 1)
    ch, cw = h / gh, w / gw
    yy = ys + 0.35 * ch * torch.sin(xs * (6.283 / (2.3 * cw)) + ph[:, 0]) + 0.2 * ch * torch.sin(xs * (6.283 / (0.9 * cw)) + ph[:, 1])
    xx = xs + 0.35 * cw * torch.sin(ys * (6.283 / (2.1 * ch)) + ph[:, 2]) + 0.2 * cw * torch.sin(ys * (6.283 / (0.8 * ch)) + ph[:, 3])
    cy = (yy / ch).floor().clamp_(0, gh - 1).long()
    cx = (xx / cw).floor().clamp_(0, gw - 1).long()
    cell = cy * gw + cx                                                   # [b,h,w] in [0, gh*gw)
    ncell = gh * gw
    ids = torch.stack([torch.randperm(max(id_span, ncell), device=device, generator=g)[:ncell] + 1 for _ in range(b)])
    ids = torch.where(torch.rand(b, ncell, device=device, generator=g) < bg_frac, torch.zeros_like(ids), ids)
    return torch.gather(ids, 1, cell.view(b, -1)).view(b, h, w)


def lrh_labels(regions, class_num=6, ignore_label=-1, seed=SEED, noise=0.2, ignore_frac=0.1):
    """int64 labels: per-region majority class + `noise` random pixels + `ignore_frac` ignored."""
    device = regions.device
    g = _gen(device, seed + 1)
    b = regions.shape[0]
    rmax = int(regions.max()) + 1
    major = torch.randint(0, class_num, (b, rmax), device=device, generator=g)
    lab = torch.gather(major, 1, regions.view(b, -1)).view_as(regions)
    u = torch.rand(regions.shape, device=device, generator=g)
    rnd = torch.randint(0, class_num, regions.shape, device=device, generator=g)
    lab = torch.where(u < noise, rnd, lab)
    lab = torch.where(u > 1.0 - ignore_frac, torch.full_like(lab, ignore_label), lab)
    return lab


def smooth_noise(shape, device, g, cells=8):
    """low-frequency noise in roughly [-1,1]: bilinear upsample of a coarse normal grid."""
    *lead, h, w = shape
    n = 1
    for d in lead:
        n *= d
    coarse = torch.randn(n, 1, cells, cells, device=device, generator=g)
    out = torch.nn.functional.interpolate(coarse, size=(h, w), mode="bilinear", align_corners=True)
    return out.view(*shape)


def step_inputs(b, h, w, class_num=6, n_regions=200, device="cpu", seed=SEED, ignore_label=-1, k=2048):
    """One step's worth of inputs: (images_s, label_s, images_t, soft_t, regs_t, prototypes)."""
    g = _gen(device, seed + 2)
    images_s = torch.randn(b, 3, h, w, device=device, generator=g).clamp_(max=1.0)   # augmentation.py:118-122
    images_t = torch.randn(b, 3, h, w, device=device, generator=g).clamp_(max=1.0)
    seg = region_maps(b, h, w, 64, device=device, seed=seed + 3, bg_frac=0.0)
    label_s = seg % class_num
    u = torch.rand(b, h, w, device=device, generator=g)
    label_s = torch.where(u < 0.05, torch.full_like(label_s, ignore_label), label_s)
    soft_t = torch.softmax(3.0 * smooth_noise((b, class_num, h, w), device, g), dim=1)
    regs_t = region_maps(b, h, w, n_regions, device=device, seed=seed + 4).unsqueeze(1)
    prototypes = torch.randn(class_num, k, device=device, generator=g).abs()
    return images_s, label_s, images_t, soft_t, regs_t, prototypes
