"""Generate the committed golden vectors by running the UNMODIFIED reference.

Run in the build container (where /root/reference is mounted):

    python tests/golden/make_golden.py

Every output array in tests/golden/*.npz is produced by the reference's own code
(imported through oracle/ref_loader.py); inputs are seeded and stored next to them.
The reference ships no tests or known-answer vectors of its own,
so these files are the pin for both the CPU oracle and the CUDA path.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.abspath(os.path.join(HERE, "..", "..")))

from oracle import ref_loader, step_oracle as so  # noqa: E402

SEED = 2333  # the reference's seed_torch value, tools/train_ssl_reg.py:274


def blocky_regions(g, b, h, w, n_regions, id_span, bg_frac=0.15):
    """Voronoi partition with sparse ids in 1..id_span and ~bg_frac background (id 0)."""
    out = torch.zeros(b, h, w, dtype=torch.int64)
    ys, xs = torch.meshgrid(torch.arange(h), torch.arange(w), indexing="ij")
    for i in range(b):
        n = max(1, n_regions)
        cy = torch.randint(0, h, (n,), generator=g)
        cx = torch.randint(0, w, (n,), generator=g)
        d = (ys[None] - cy[:, None, None]) ** 2 + (xs[None] - cx[:, None, None]) ** 2
        cell = d.argmin(0)
        ids = torch.randperm(id_span, generator=g)[:n] + 1
        ids[torch.rand(n, generator=g) < bg_frac] = 0
        out[i] = ids[cell]
    return out


def lrh_cases(ns):
    g = torch.Generator().manual_seed(SEED)
    cases = []

    def run(name, labels, regions, percent, class_num=6, ignore=-1):
        ref = ns.Homogenizer(percent=percent, class_num=class_num, ignore_label=ignore)
        l0, r0 = labels.clone(), regions.clone()
        out = ref(labels, regions)
        assert torch.equal(l0, labels) and torch.equal(r0, regions)  # inputs are never mutated
        cases.append(dict(name=name, labels=labels.numpy(), regions=regions.numpy(), out=out.numpy(),
                          percent=percent, class_num=class_num, ignore=ignore))

    # seeded random tiles: majority class per region + 20 % noise + 10 % ignore
    for k, (n_reg, pct) in enumerate([(5, 0.5), (40, 0.5), (400, 0.5), (40, 0.9), (40, 0.3)]):
        regs = blocky_regions(g, 2, 48, 48, n_reg, 2 * n_reg)
        major = torch.randint(0, 6, (2, 2 * n_reg + 1), generator=g)
        lab = torch.gather(major, 1, regs.view(2, -1)).view(2, 48, 48)
        noise = torch.rand(2, 48, 48, generator=g)
        lab = torch.where(noise < 0.2, torch.randint(0, 6, (2, 48, 48), generator=g), lab)
        lab = torch.where(noise > 0.9, torch.full_like(lab, -1), lab)
        run(f"random_{k}", lab, regs, pct)

    # exact-percent regions straddling the float32 (n + 1e-5 == n) boundary at n = 256
    for pct, ns_ in [(0.5, [4, 100, 254, 256, 512, 1024, 4096]), (0.9, [10, 1000, 10000]), (0.25, [8, 256])]:
        for n in ns_:
            k = int(round(n * pct))
            lab = torch.full((1, 1, n + 3), -1, dtype=torch.int64)
            lab[0, 0, :k] = 2
            lab[0, 0, k:n] = 4 if pct >= 0.5 else 5
            if pct < 0.5:   # split the remainder so class 2 stays the unique max
                rest = n - k
                lab[0, 0, k:k + rest // 3] = 0
                lab[0, 0, k + rest // 3:k + 2 * (rest // 3)] = 1
            regs = torch.full((1, 1, n + 3), 7, dtype=torch.int64)
            regs[0, 0, n:] = 0
            run(f"exact_p{pct}_n{n}", lab, regs, pct)

    # ties -> lowest class id; all-ignored region; region 0; ignored px inside a passing region
    lab = torch.tensor([[[3, 3, 1, 1, -1, -1, -1, 5, 5, 5, 0, -1, 2, 2]]])
    reg = torch.tensor([[[9, 9, 9, 9, 4, 4, 4, 0, 0, 0, 6, 6, 6, 6]]])
    for pct in (0.5, 0.0, -1.0, 0.75, 1.0):
        run(f"edge_ties_p{pct}", lab, reg, pct)
    # per-image histograms: same region id in two images must not merge
    lab = torch.tensor([[[1, 1, 1, 2]], [[2, 2, 2, 1]]])
    reg = torch.tensor([[[5, 5, 5, 5]], [[5, 5, 5, 5]]])
    run("per_image", lab, reg, 0.5)
    # label == class_num passes through and counts nowhere; ignore_label = 255 variant
    lab = torch.tensor([[[6, 6, 6, 1, 1, 255, 255, 0]]])
    reg = torch.tensor([[[3, 3, 3, 3, 3, 3, 8, 8]]])
    run("label_eq_classnum_ign255", lab, reg, 0.5, 6, 255)
    # very sparse / large ids
    lab = torch.randint(-1, 6, (1, 16, 16), generator=g)
    reg = torch.randint(0, 4, (1, 16, 16), generator=g) * 20011
    run("sparse_big_ids", lab, reg, 0.4)
    # 7-class (LoveDA) and 16-class variants
    for cn in (7, 16):
        regs = blocky_regions(g, 1, 40, 56, 12, 30)
        lab = torch.randint(-1, cn, (1, 40, 56), generator=g)
        run(f"classes_{cn}", lab, regs, 0.2, cn, -1)

    flat = {}
    for i, c in enumerate(cases):
        for k, v in c.items():
            flat[f"{i:03d}/{k}"] = np.asarray(v)
    np.savez_compressed(os.path.join(HERE, "lrh.npz"), **flat)
    print("lrh cases:", len(cases))


def select_and_downscale(ns):
    g = torch.Generator().manual_seed(SEED + 1)
    d = {}
    soft = torch.softmax(3 * torch.randn(2, 6, 24, 24, generator=g), 1)
    soft[0, 1, 0, 0] = 0.0   # make channel maxima interesting: force an exact-threshold pixel
    m = soft[1, 2].max()
    soft[1, 2, 3, 3] = (m * 0.8)  # == threshold -> strict '>' must reject
    d["soft"] = soft.numpy()
    for i, (top, low) in enumerate([(0.8, 0.6), (0.8, 0.3), (0.5, 0.2), (1.0, 0.0)]):
        out = ns.pseudo_selection(soft.clone(), top, low, "tensor", -1)
        d[f"sel_{i}"] = out.numpy()
        d[f"sel_{i}_args"] = np.array([top, low], dtype=np.float64)
    out = ns.pseudo_selection(soft.clone(), 0.8, 0.6, "tensor", 255)
    d["sel_ign255"] = out.numpy()

    lab = torch.randint(-1, 6, (2, 64, 96), generator=g)
    lab[:, :32, :32] = 2
    lab[:, 32:48, :16] = -1
    lab[0, :16, 32:48] = torch.where(torch.rand(16, 16, generator=g) < 0.75, torch.tensor(4), torch.tensor(1))
    lab[1, 48:64, 48:64] = 3
    lab[1, 48:60, 48:64] = 3          # exactly 192/256 = 0.75 of class 3 below
    lab[1, 60:64, 48:64] = 5
    d["ds_label"] = lab.numpy()
    for i, (scale, mr) in enumerate([(16, 0.75), (8, 0.5), (16, 0.3), (4, 0.75)]):
        out = ns.DownscaleLabel(scale_factor=scale, n_classes=6, ignore_label=-1, min_ratio=mr)(lab)
        d[f"ds_{i}"] = out.numpy()
        d[f"ds_{i}_args"] = np.array([scale, mr], dtype=np.float64)
    np.savez_compressed(os.path.join(HERE, "select_downscale.npz"), **d)


def aligner_and_loss(ns):
    g = torch.Generator().manual_seed(SEED + 2)
    C, K, b, h, w, H, W = 6, 64, 2, 4, 4, 64, 64
    d = {}
    al = ns.Aligner(ns.logger, K, C, -1, 0.996)
    proto = torch.randn(C, K, generator=g).abs()
    al.prototypes = proto.clone()
    feat = torch.randn(b, K, h, w, generator=g)
    p1 = 3 * torch.randn(b, C, h, w, generator=g)
    p2 = 3 * torch.randn(b, C, h, w, generator=g)
    soft = torch.softmax(3 * torch.randn(b, C, H, W, generator=g), 1)
    d.update(proto=proto.numpy(), feat=feat.numpy(), p1=p1.numpy(), p2=p2.numpy(), soft=soft.numpy())
    d["pearson"] = al._pearson_dist(feat.permute(0, 2, 3, 1).reshape(-1, K), proto).numpy()
    d["refined_T2"] = al.label_refine(None, feat, [p1, p2], soft, True, "all", 2.0).numpy()
    d["refined_T1"] = al.label_refine(None, feat, [p1, p2], soft, True, "all", 1.0).numpy()
    d["refined_single_pred"] = al.label_refine(None, feat, p1, soft, True, "all", 2.0).numpy()

    lab = torch.randint(-1, C, (b, H, W), generator=g)
    lab[0, :32, :32] = 1
    lab[0, 32:, :16] = 3
    lab[1, :48, 16:] = 5   # class 0, 2, 4 get no 16x16 block -> keep old prototype
    d["label_s"] = lab.numpy()
    ds = al.update_prototype(feat, lab)
    d["label_ds"] = ds.numpy()
    d["proto_after"] = al.prototypes.numpy()
    # update_avg / init_avg (tools/init_prototypes.py:101-111)
    al2 = ns.Aligner(ns.logger, K, C, -1, 0.996)
    al2.update_avg(feat, lab)
    al2.update_avg(feat * 0.5 + 1.0, lab)
    al2.init_avg()
    d["proto_init_avg"] = al2.prototypes.numpy()

    # loss_calc + CrossEntropy, with gradients w.r.t. the low-res logits
    ce = ns.CrossEntropy(ignore_label=-1, class_balancer=None)
    q1 = p1.clone().requires_grad_(True)
    q2 = p2.clone().requires_grad_(True)
    loss = ns.loss_calc([q1, q2], lab, ce, multi=True)
    loss.backward()
    d.update(loss=loss.detach().numpy(), dp1=q1.grad.numpy(), dp2=q2.grad.numpy())
    lab_all_ign = torch.full_like(lab, -1)
    d["loss_all_ignored"] = ns.loss_calc([p1, p2], lab_all_ign, ce, multi=True).numpy()

    # ClassBalance (flag-gated, balance.py:15-78)
    cb = ns.ClassBalance(class_num=C, ignore_label=-1, decay=0.99, temperature=2.0)
    wpx = cb.get_class_weight_4pixel(lab)
    d["cb_freq"] = cb.freq.numpy()
    d["cb_class_weight"] = cb._get_class_wight().numpy()
    d["cb_weight_sum"] = np.array(float(wpx.sum()))
    np.savez_compressed(os.path.join(HERE, "aligner_loss.npz"), **d)


def _no_dropout(m):
    for mod in m.modules():
        if isinstance(mod, torch.nn.Dropout2d):
            mod.p = 0.0


def model_and_step(ns):
    g = torch.Generator().manual_seed(SEED + 3)
    C = 6
    for rt, hw in (("resnet50", 64), ("resnet101", 96)):
        d = {}
        m = ref_loader.build_reference_model(ns, rt, C)
        m.load_state_dict(so.seeded_state_dict(m, SEED))
        _no_dropout(m)
        x = torch.randn(2, 3, hw, hw, generator=g).clamp(max=1.0)
        lab = torch.randint(-1, C, (2, hw, hw), generator=g)
        d.update(x=x.numpy(), label=lab.numpy())
        m.eval()
        with torch.no_grad():
            d["eval_prob"] = m(x).numpy()
        m.train()
        x1, x2, feat = m(x)
        d.update(x1=x1.detach().numpy(), x2=x2.detach().numpy(), feat=feat.detach().numpy())
        ce = ns.CrossEntropy(ignore_label=-1, class_balancer=None)
        loss = ns.loss_calc([x1, x2], lab, ce, multi=True)
        loss.backward()
        d["loss"] = loss.detach().numpy()
        names, norms = [], []
        for n_, p in m.named_parameters():
            names.append(n_)
            norms.append(float(p.grad.norm()))
        d["grad_names"] = np.array(names)
        d["grad_norms"] = np.array(norms, dtype=np.float64)
        d["grad_conv1"] = m.encoder.resnet.conv1.weight.grad.numpy()
        d["grad_cls5"] = m.layer5.conv_last[4].weight.grad.numpy()
        d["grad_cls5_bias"] = m.layer5.conv_last[4].bias.grad.numpy()
        d["bn1_running_mean"] = m.encoder.resnet.bn1.running_mean.numpy()
        d["bn1_running_var"] = m.encoder.resnet.bn1.running_var.numpy()
        np.savez_compressed(os.path.join(HERE, f"model_{rt}.npz"), **d)

    # one full inner step through the reference's own objects (tools/train_ssl_reg.py:198-241)
    d = {}
    hw = 64
    m = ref_loader.build_reference_model(ns, "resnet50", C)
    m.load_state_dict(so.seeded_state_dict(m, SEED))
    _no_dropout(m)
    m.train()
    al = ns.Aligner(ns.logger, 2048, C, -1, 0.996)
    proto = torch.randn(C, 2048, generator=g).abs()
    al.prototypes = proto.clone()
    hom = ns.Homogenizer(percent=0.5, class_num=C, ignore_label=-1)
    ce = ns.CrossEntropy(ignore_label=-1, class_balancer=None)
    opt = torch.optim.SGD(m.parameters(), lr=1e-2, momentum=0.9, weight_decay=5e-4)
    xs = torch.randn(2, 3, hw, hw, generator=g).clamp(max=1.0)
    xt = torch.randn(2, 3, hw, hw, generator=g).clamp(max=1.0)
    ls = torch.randint(-1, C, (2, hw, hw), generator=g)
    ls[:, :32, :32] = 2
    soft = torch.softmax(6 * torch.randn(2, C, hw, hw, generator=g), 1)
    regs = blocky_regions(g, 2, hw, hw, 9, 20).unsqueeze(1)
    d.update(xs=xs.numpy(), xt=xt.numpy(), ls=ls.numpy(), soft=soft.numpy(), regs=regs.numpy(), proto=proto.numpy())
    losses = []
    for it in range(2):
        ps1, ps2, fs = m(xs)
        pt1, pt2, ft = m(xt)
        soft_r = al.label_refine(None, ft, [pt1, pt2], soft, refine=True, mode="all", temp=2.0)
        hard = ns.pseudo_selection(soft_r, cutoff_top=0.8, cutoff_low=0.6, return_type="tensor", ignore_label=-1)
        hard = hom(hard, regs.squeeze(1))
        al.update_prototype(fs, ls)
        l_s = ns.loss_calc([ps1, ps2], ls, loss_fn=ce, multi=True)
        l_t = ns.loss_calc([pt1, pt2], hard, loss_fn=ce, multi=True)
        loss = l_s + l_t
        opt.zero_grad()
        loss.backward()
        gn = torch.nn.utils.clip_grad_norm_(filter(lambda p: p.requires_grad, m.parameters()), max_norm=32, norm_type=2)
        opt.step()
        losses.append([float(loss), float(l_s), float(l_t), float(gn)])
        if it == 0:
            d["hard_0"] = hard.numpy()
            d["soft_refined_0"] = soft_r.detach().numpy()
    d["losses"] = np.array(losses, dtype=np.float64)
    d["proto_after"] = al.prototypes.numpy()
    d["conv1_after"] = m.encoder.resnet.conv1.weight.detach().numpy()
    d["cls6_after"] = m.layer6.conv_last[4].weight.detach().numpy()
    np.savez_compressed(os.path.join(HERE, "step_resnet50.npz"), **d)


def teacher_pass(ns):
    """Offline teacher pass: the reference's own pre_slide / tta_predict (regda/utils/tools.py:61-152)
    around its eval-mode Deeplabv2, with the ttach restatement of oracle/ref_loader.py.  Tiles of 64 so the CPU run is quick:
    a 96x112 image (2 x 3 windows, the last column shifted back inside) and a 48x64 image (smaller than the tile: the
    reference's pad_image pads the TOP of the height, tools.py:56)."""
    from regda.utils.tools import pre_slide, tta_predict
    g = torch.Generator().manual_seed(SEED + 11)
    C = 6
    m = ref_loader.build_reference_model(ns, "resnet50", C)
    m.load_state_dict(so.seeded_state_dict(m, SEED))
    m.eval()
    d = {}
    with torch.no_grad():
        x = torch.randn(1, 3, 96, 112, generator=g).clamp(max=1.0)
        d["image"] = x.numpy()
        d["tile_tta"] = tta_predict(m, x[:, :, :64, :64]).numpy()
        d["slide_tta"] = pre_slide(m, x, num_classes=C, tile_size=(64, 64), tta=True).numpy()
        d["slide_plain"] = pre_slide(m, x, num_classes=C, tile_size=(64, 64), tta=False).numpy()
        xs = torch.randn(1, 3, 48, 64, generator=g).clamp(max=1.0)
        d["image_small"] = xs.numpy()
        d["slide_small_tta"] = pre_slide(m, xs, num_classes=C, tile_size=(64, 64), tta=True).numpy()
    np.savez_compressed(os.path.join(HERE, "teacher_pass.npz"), **d)


def align_step(ns):
    """Stage-2 step through the reference's own objects, tools/train_align_reg.py:144-196 (default flags:
    refine-label on, align-domain off, pcl-temp 8), ResNet-50, 2 + 2 tiles of 64x64, two iterations; plus
    PrototypeContrastiveLoss alone with its gradient."""
    import torch.nn.functional as tnf
    from regda.loss import PrototypeContrastiveLoss
    g = torch.Generator().manual_seed(SEED + 21)
    C, hw = 6, 64
    d = {}
    # the loss alone
    pcl = PrototypeContrastiveLoss(temperature=8.0, ignore_label=-1)
    proto = torch.randn(C, 256, generator=g).abs()
    feat = torch.randn(2, 256, 5, 7, generator=g).requires_grad_(True)
    lab = torch.randint(-1, C, (2, 1, 5, 7), generator=g)
    loss = pcl(proto, feat, lab)
    (loss * 0.5).backward()
    d.update(pcl_proto=proto.numpy(), pcl_feat=feat.detach().numpy(), pcl_label=lab.numpy(), pcl_loss=loss.detach().numpy(),
             pcl_dfeat_half=feat.grad.numpy())
    # the step
    m = ref_loader.build_reference_model(ns, "resnet50", C)
    m.load_state_dict(so.seeded_state_dict(m, SEED))
    _no_dropout(m)
    m.train()
    al = ns.Aligner(ns.logger, 2048, C, -1, 0.996)
    proto = torch.randn(C, 2048, generator=g).abs()
    al.prototypes = proto.clone()
    hom = ns.Homogenizer(percent=0.5, class_num=C, ignore_label=-1)
    ce = ns.CrossEntropy(ignore_label=-1, class_balancer=None)
    opt = torch.optim.SGD(m.parameters(), lr=1e-2, momentum=0.9, weight_decay=5e-4)
    xs = torch.randn(2, 3, hw, hw, generator=g).clamp(max=1.0)
    xt = torch.randn(2, 3, hw, hw, generator=g).clamp(max=1.0)
    ls = torch.randint(-1, C, (2, hw, hw), generator=g)
    ls[:, :32, :32] = 2
    ls[:, 32:, 16:48] = 4
    regs = blocky_regions(g, 2, hw, hw, 9, 20).unsqueeze(1)
    d.update(xs=xs.numpy(), xt=xt.numpy(), ls=ls.numpy(), regs=regs.numpy(), proto=proto.numpy())
    losses = []
    for it in range(2):
        ps1, ps2, fs = m(xs)
        label_s_down = al.update_prototype(fs, ls)
        pt1, pt2, ft = m(xt)
        x1 = tnf.interpolate(pt1, xt.shape[-2:], mode='bilinear', align_corners=True)
        x2 = tnf.interpolate(pt2, xt.shape[-2:], mode='bilinear', align_corners=True)
        soft = ((x1.softmax(dim=1) + x2.softmax(dim=1)) * 0.5).detach()
        soft = al.label_refine(None, ft, [pt1, pt2], soft, refine=True, mode="all", temp=2.0)
        hard = ns.pseudo_selection(soft, cutoff_top=0.8, cutoff_low=0.6, return_type="tensor", ignore_label=-1)
        hard = hom(hard, regs.squeeze(1))
        label_t = al.downscale_gt(hard)
        l_seg = ns.loss_calc([ps1, ps2], ls, loss_fn=ce, multi=True)
        l_al = (pcl(al.prototypes, fs, label_s_down) + pcl(al.prototypes, ft, label_t)) * 0.5
        loss = l_seg + l_al
        opt.zero_grad()
        loss.backward()
        gn = torch.nn.utils.clip_grad_norm_(filter(lambda p: p.requires_grad, m.parameters()), max_norm=32, norm_type=2)
        opt.step()
        losses.append([float(loss), float(l_seg), float(l_al), float(gn)])
        if it == 0:
            d["hard_0"] = hard.numpy()
            d["label_t_0"] = label_t.numpy()
            d["label_s_down_0"] = label_s_down.numpy()
    d["losses"] = np.array(losses, dtype=np.float64)
    d["proto_after"] = al.prototypes.numpy()
    d["conv1_after"] = m.encoder.resnet.conv1.weight.detach().numpy()
    np.savez_compressed(os.path.join(HERE, "align_step_resnet50.npz"), **d)


def coral(ns):
    """CORAL: the reference's CoralLoss alone with its gradients (regda/gast/coral.py:26-47), and two
    iterations of the stage-2 step with --align-domain 1 (tools/train_align_reg.py:187; the shipped recipe
    runs/regda/run_2potsdam.sh:15 turns it on), ResNet-50, 2 + 2 tiles of 64x64."""
    import torch.nn.functional as tnf
    from regda.gast.coral import CoralLoss
    from regda.loss import PrototypeContrastiveLoss
    g = torch.Generator().manual_seed(SEED + 41)
    d = {}
    for k, (nsrc, ntgt, dim, is_sqrt) in enumerate([(96, 96, 128, False), (200, 136, 64, False), (64, 64, 256, True)]):
        mix = torch.randn(dim, dim, generator=g) / dim ** 0.5
        src = (torch.randn(nsrc, dim, generator=g) @ mix + 0.3).requires_grad_(True)
        tgt = (torch.randn(ntgt, dim, generator=g) * 1.3 - 0.2).requires_grad_(True)
        loss = CoralLoss(is_sqrt=is_sqrt)(src, tgt)
        loss.backward()
        d.update({f"loss{k}/src": src.detach().numpy(), f"loss{k}/tgt": tgt.detach().numpy(), f"loss{k}/is_sqrt": np.array(is_sqrt),
                  f"loss{k}/loss": loss.detach().numpy(), f"loss{k}/dsrc": src.grad.numpy(), f"loss{k}/dtgt": tgt.grad.numpy()})
    C, hw = 6, 64
    pcl = PrototypeContrastiveLoss(temperature=8.0, ignore_label=-1)
    m = ref_loader.build_reference_model(ns, "resnet50", C)
    m.load_state_dict(so.seeded_state_dict(m, SEED))
    _no_dropout(m)
    m.train()
    al = ns.Aligner(ns.logger, 2048, C, -1, 0.996)
    proto = torch.randn(C, 2048, generator=g).abs()
    al.prototypes = proto.clone()
    hom = ns.Homogenizer(percent=0.5, class_num=C, ignore_label=-1)
    ce = ns.CrossEntropy(ignore_label=-1, class_balancer=None)
    opt = torch.optim.SGD(m.parameters(), lr=1e-2, momentum=0.9, weight_decay=5e-4)
    xs = torch.randn(2, 3, hw, hw, generator=g).clamp(max=1.0)
    xt = (torch.randn(2, 3, hw, hw, generator=g) * 0.8 + 0.2).clamp(max=1.0)
    ls = torch.randint(-1, C, (2, hw, hw), generator=g)
    ls[:, :32, :32] = 2
    ls[:, 32:, 16:48] = 4
    regs = blocky_regions(g, 2, hw, hw, 9, 20).unsqueeze(1)
    d.update(xs=xs.numpy(), xt=xt.numpy(), ls=ls.numpy(), regs=regs.numpy(), proto=proto.numpy())
    losses = []
    for it in range(2):
        ps1, ps2, fs = m(xs)
        label_s_down = al.update_prototype(fs, ls)
        pt1, pt2, ft = m(xt)
        x1 = tnf.interpolate(pt1, xt.shape[-2:], mode='bilinear', align_corners=True)
        x2 = tnf.interpolate(pt2, xt.shape[-2:], mode='bilinear', align_corners=True)
        soft = ((x1.softmax(dim=1) + x2.softmax(dim=1)) * 0.5).detach()
        soft = al.label_refine(None, ft, [pt1, pt2], soft, refine=True, mode="all", temp=2.0)
        hard = ns.pseudo_selection(soft, cutoff_top=0.8, cutoff_low=0.6, return_type="tensor", ignore_label=-1)
        hard = hom(hard, regs.squeeze(1))
        label_t = al.downscale_gt(hard)
        l_seg = ns.loss_calc([ps1, ps2], ls, loss_fn=ce, multi=True)
        l_dom = al.align_domain(fs, ft)
        l_al = (pcl(al.prototypes, fs, label_s_down) + pcl(al.prototypes, ft, label_t)) * 0.5
        loss = l_seg + l_dom + l_al
        opt.zero_grad()
        loss.backward()
        gn = torch.nn.utils.clip_grad_norm_(filter(lambda p: p.requires_grad, m.parameters()), max_norm=32, norm_type=2)
        opt.step()
        losses.append([float(loss), float(l_seg), float(l_al), float(l_dom), float(gn)])
    d["losses"] = np.array(losses, dtype=np.float64)
    d["proto_after"] = al.prototypes.numpy()
    d["conv1_after"] = m.encoder.resnet.conv1.weight.detach().numpy()
    np.savez_compressed(os.path.join(HERE, "coral.npz"), **d)


def miou_metric(ns):
    """mIoU through the reference's own PixelMetricIgnore.summary_all (regda/gast/metrics.py:19-65: 5-decimal rounding of the
    per-class figures, class-0 pop for IsprsDA, rounded means) fed as regda/utils/eval.py:43-49 feeds it, on top of the
    restated ever PixelMetric of oracle/ref_loader.py."""
    from regda.gast.metrics import PixelMetricIgnore
    g = torch.Generator().manual_seed(SEED + 31)
    d = {}
    for k, (C, ignore, absent) in enumerate([(6, [0], None), (7, [], None), (6, [0], 4)]):
        gt = torch.randint(-1, C, (3, 40, 56), generator=g)
        # a prediction that agrees with the label on ~70 % of the pixels
        pred = torch.where(torch.rand(gt.shape, generator=g) < 0.7, gt.clamp(min=0), torch.randint(0, C, gt.shape, generator=g))
        if absent is not None:      # a class that never occurs in label or prediction: 0/0 = nan in the reference
            gt = torch.where(gt == absent, torch.zeros_like(gt), gt)
            pred = torch.where(pred == absent, torch.zeros_like(pred), pred)
        op = PixelMetricIgnore(C, class_names=[f"c{i}" for i in range(C)], logdir=None, logger=None, ignore_labels=list(ignore))
        for i in range(gt.shape[0]):                       # one forward per image, as the eval loop does
            cls_gt = gt[i].numpy().astype(np.int32)
            mask = cls_gt >= 0
            op.forward(cls_gt[mask].ravel(), pred[i].numpy()[mask].ravel())
        tb, miou = op.summary_all()
        d[f"{k}/gt"] = gt.numpy()
        d[f"{k}/pred"] = pred.numpy()
        d[f"{k}/num_classes"] = np.array(C)
        d[f"{k}/ignore"] = np.array(ignore, dtype=np.int64)
        d[f"{k}/miou"] = np.array(miou, dtype=np.float64)
        d[f"{k}/rows"] = np.array([[float(v) for v in r[2:]] for r in tb.rows[:-1]], dtype=np.float64)     # iou, f1, precision, recall
        d[f"{k}/means"] = np.array([float(v) for v in tb.rows[-1][2:]], dtype=np.float64)
    np.savez_compressed(os.path.join(HERE, "miou.npz"), **d)


if __name__ == "__main__":
    assert ref_loader.reference_available(), "run this where /root/reference is mounted"
    torch.set_num_threads(os.cpu_count() or 1)
    ns = ref_loader.load()
    if "--only-teacher" in sys.argv:
        teacher_pass(ns)
        sys.exit(0)
    if "--only-align" in sys.argv:
        align_step(ns)
        sys.exit(0)
    if "--only-coral" in sys.argv:
        coral(ns)
        sys.exit(0)
    if "--only-miou" in sys.argv:
        miou_metric(ns)
        sys.exit(0)
    lrh_cases(ns)
    select_and_downscale(ns)
    aligner_and_loss(ns)
    model_and_step(ns)
    teacher_pass(ns)
    align_step(ns)
    miou_metric(ns)
    coral(ns)
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(HERE, f)) // 1024, "KiB")
