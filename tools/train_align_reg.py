#!/usr/bin/env python
"""Stage-2 prototype alignment -- the reference's tools/train_align_reg.py (CLI :33-59, loop :144-216) on the B200 kernels.

    python tools/train_align_reg.py --config-path st.regda.2potsdam --ckpt-model <stage-1 .pth> --ckpt-proto <prototypes .pth> \
        --sam-refine --percent 0.5 [--steps N] [--cuda-graph 1]
    torchrun --nproc-per-node 8 tools/train_align_reg.py ...        # image-sharded data parallel (regda_b200/parallel.py)

Per iteration (regda_b200/trainer.py AlignStep): paired source/target forward, prototype EMA from the source features, target
soft labels from the model's own two heads, label refinement + selection + LRH, source segmentation loss and the
prototype-contrastive loss on both domains (+ the CORAL domain loss with `--align-domain 1`, as the shipped recipe
runs/regda/run_2potsdam.sh:15 sets it), backward, clip, SGD.  Same flags and meaning as the reference; `--ls OhemCrossEntropy`
belongs to code that is out of scope and is refused.
Data: synthetic tensors of the reference's shapes (`--data synthetic`), as tools/train_ssl_reg.py."""
from __future__ import annotations

import argparse
import os
import os.path as osp
import sys
import time

ROOT = osp.dirname(osp.dirname(osp.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from regda_b200 import synth  # noqa: E402
from regda_b200.gast.alignment import Aligner  # noqa: E402
from regda_b200.gast.balance import ClassBalance, CrossEntropy  # noqa: E402
from regda_b200.models.Encoder import Deeplabv2  # noqa: E402
from regda_b200.trainer import AlignStep, GraphedStep  # noqa: E402
from regda_b200.utils.local_region_homog import Homogenizer  # noqa: E402
from regda_b200.utils.tools import adjust_learning_rate, import_config, seed_torch  # noqa: E402


def str2bool(v):
    return str(v).lower() in ("1", "true", "yes", "y", "t")


def parse():
    p = argparse.ArgumentParser(description='RegDA stage-2 alignment by the prototype-contrastive loss (B200-native hot path)')
    p.add_argument('--config-path', type=str, default='st.regda.2potsdam')
    p.add_argument('--ckpt-model', type=str, default='', help='model ckpt from stage 1 (reference state_dict keys)')
    p.add_argument('--ckpt-proto', type=str, default='', help='prototypes [C,2048] from tools/init_prototypes.py')
    p.add_argument('--gen', type=str2bool, default=1)
    p.add_argument('--align-domain', type=str2bool, default=0, help='CORAL domain alignment (regda/gast/coral.py) between the source and target features')
    p.add_argument('--refine-label', type=str2bool, default=1)
    p.add_argument('--refine-mode', type=str, default='all', choices=['all'])
    p.add_argument('--refine-temp', type=float, default=2.0)
    p.add_argument('--sam-refine', action='store_true', help='Local Region Homogenizing')
    p.add_argument('--percent', type=float, default=0.5)
    p.add_argument('--ls', type=str, default='CrossEntropy', choices=['CrossEntropy'])
    p.add_argument('--bcs', type=str2bool, default=0)
    p.add_argument('--class-temp', type=float, default=2.0)
    p.add_argument('--pcl-temp', type=float, default=8.0)
    p.add_argument('--data', type=str, default='synthetic', choices=['synthetic'])
    p.add_argument('--steps', type=int, default=0, help='override STAGE2_STEPS (0 = config)')
    p.add_argument('--cuda-graph', type=str2bool, default=1)
    p.add_argument('--region-bound', type=int, default=0)
    return p.parse_args()


def main():
    args = parse()
    cfg = import_config(args.config_path, create=True)
    rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1)); local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1"); os.environ.setdefault("MASTER_PORT", "29534")
        dist.init_process_group("nccl", device_id=dev)
    seed_torch(2333 + rank)
    log = (lambda s: print(s, flush=True)) if rank == 0 else (lambda s: None)

    class_num, ignore_label = cfg.CLASS_NUM, cfg.IGNORE_LABEL
    stop_steps = args.steps or cfg.STAGE2_STEPS
    cfg.NUM_STEPS = stop_steps * 1.5                          # train_align_reg.py:84
    cfg.PREHEAT_STEPS = max(int(stop_steps / 20), 1)          # :85
    model_name = str(cfg.MODEL).lower()
    model_name = 'resnet50' if model_name == 'resnet' else model_name
    model = Deeplabv2(dict(backbone=dict(resnet_type=model_name, output_stride=16, pretrained=False), multi_layer=True, cascade=False,
                           use_ppm=True, ppm=dict(num_classes=class_num, use_aux=False, fc_dim=2048), inchannels=2048,
                           num_classes=class_num, is_ins_norm=True))
    if args.ckpt_model:
        model.load_state_dict(torch.load(args.ckpt_model, map_location='cpu'), strict=True)
    else:
        log('WARNING: no --ckpt-model: training from random weights')
    model = model.to(dev).train()
    aligner = Aligner(logger=None, feat_channels=2048, class_num=class_num, ignore_label=ignore_label, decay=0.996,
                      resume=args.ckpt_proto or None, device=dev)
    cb_s = ClassBalance(class_num=class_num, ignore_label=ignore_label, decay=0.99, temperature=args.class_temp)
    loss_s = CrossEntropy(ignore_label=ignore_label, class_balancer=cb_s if args.bcs else None)

    h, w = cfg.SYNTHETIC["size"]
    xs, ls, xt, _, regs, proto = synth.step_inputs(cfg.BATCH_SIZE, h, w, class_num, cfg.SYNTHETIC["regions_per_tile"], device=dev, seed=2333 + rank)
    if not args.ckpt_proto:
        aligner.prototypes = proto.clone()
    batch = [xs, ls, xt, regs]                                # NOTE: one fixed seeded batch per rank, replayed every step
    bound = args.region_bound or int(regs.max()) + 1
    hom = Homogenizer(percent=args.percent, class_num=class_num, ignore_label=ignore_label, region_bound=bound, strict=False)
    step = AlignStep(model, aligner, hom, class_num=class_num, ignore_label=ignore_label, cutoff_top=cfg.CUTOFF_TOP,
                     cutoff_low=cfg.CUTOFF_LOW, refine_temp=args.refine_temp, sam_refine=args.sam_refine,
                     refine_label=bool(args.refine_label), momentum=cfg.MOMENTUM, weight_decay=cfg.WEIGHT_DECAY,
                     loss_fn_s=loss_s, world_size=world, pcl_temp=args.pcl_temp, align_domain=bool(args.align_domain))
    use_graph = bool(args.cuda_graph) and not args.bcs
    runner = GraphedStep(step, batch, lr=0.0) if use_graph else None

    class _Opt:
        param_groups = [dict(lr=0.0)]

    t0 = time.time()
    os.makedirs(cfg.SNAPSHOT_DIR, exist_ok=True)
    for i_iter in range(stop_steps):
        lr = adjust_learning_rate(_Opt, i_iter, cfg)            # :150
        if runner is not None:
            o = runner(*batch, lr=lr)
            out = dict(loss=o["loss"], loss_seg=o["loss_source"], loss_align=o["loss_target"])
        else:
            out = step(*batch, lr)
        if i_iter == 0 or (i_iter + 1) % 50 == 0:               # :197-201
            log(f"iter={i_iter + 1}, total={float(out['loss']):.3f}, loss_seg={float(out['loss_seg']):.3f}, "
                f"loss_align={float(out['loss_align']):.3e}, "
                f"loss_domain={(float(step.loss_domain) if step.loss_domain is not None else 0.0):.3e} lr={lr:.3e}")
            hom.check()
            step.loss_fn_pcl.check()
        if (i_iter + 1) % cfg.EVAL_EVERY == 0 or (i_iter + 1) >= stop_steps:        # :203-211
            if rank == 0:
                torch.save({k: v.detach().cpu() for k, v in model.state_dict().items()}, osp.join(cfg.SNAPSHOT_DIR, cfg.TARGET_SET + '_curr.pth'))
                torch.save(aligner.prototypes.cpu(), osp.join(cfg.SNAPSHOT_DIR, 'prototypes_curr.pth'))
    torch.cuda.synchronize()
    dt = time.time() - t0
    imgs = 2 * cfg.BATCH_SIZE * world * stop_steps
    log(f">>>> Using {dt / 3600:.3f} hours, {imgs / dt:.1f} images/s over {world} GPU(s).")
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
