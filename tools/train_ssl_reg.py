#!/usr/bin/env python
"""Stage-3 self-training -- the reference's tools/train_ssl_reg.py (CLI :34-65, loop :176-266) on the B200 kernels.

    python tools/train_ssl_reg.py --config-path st.regda.2potsdam --ckpt-model <stage-2 .pth> --ckpt-proto <prototypes .pth> \
        --sam-refine --percent 0.5 [--data synthetic] [--steps N] [--cuda-graph 1]
    torchrun --nproc-per-node 8 tools/train_ssl_reg.py ...        # image-sharded data parallel (regda_b200/parallel.py)

Same flags and meaning as the reference for the switches on the hot path (--refine-label/--refine-temp, --sam-refine,
--percent, --bcs/--bct/--class-temp); the alternative target losses (--lt != none) belong to code that is out of scope and
are refused.  Data: the reference's DALoader is CPU file I/O (out of scope); `--data synthetic` (default) trains on the
seeded synthetic tensors of the same shapes, `--data reference` uses the reference's own DALoader objects when the
reference package and its data are importable.  Checkpoints keep the reference's state_dict keys."""
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

from regda_b200 import parallel, synth  # noqa: E402
from regda_b200.gast.alignment import Aligner  # noqa: E402
from regda_b200.gast.balance import ClassBalance, CrossEntropy  # noqa: E402
from regda_b200.gast.pseudo_generation import gener_target_pseudo  # noqa: E402
from regda_b200.models.Encoder import Deeplabv2  # noqa: E402
from regda_b200.trainer import GraphedStep, SelfTrainingStep  # noqa: E402
from regda_b200.utils.local_region_homog import Homogenizer  # noqa: E402
from regda_b200.utils.tools import adjust_learning_rate, import_config, seed_torch  # noqa: E402


def str2bool(v):
    return str(v).lower() in ("1", "true", "yes", "y", "t")


def parse():
    p = argparse.ArgumentParser(description='RegDA stage-3 self-training (B200-native hot path)')
    p.add_argument('--config-path', type=str, default='st.regda.2potsdam')
    p.add_argument('--ckpt-model', type=str, default='', help='model ckpt from stage 2 (reference state_dict keys)')
    p.add_argument('--ckpt-proto', type=str, default='', help='prototypes [C,2048] from tools/init_prototypes.py')
    p.add_argument('--refine-label', type=str2bool, default=1)
    p.add_argument('--refine-mode', type=str, default='all', choices=['all'])
    p.add_argument('--refine-temp', type=float, default=2.0)
    p.add_argument('--sam-refine', action='store_true', help='Local Region Homogenizing')
    p.add_argument('--percent', type=float, default=0.5, help='class-frequency threshold of LRH')
    p.add_argument('--ls', type=str, default='CrossEntropy', choices=['CrossEntropy'])
    p.add_argument('--bcs', type=str2bool, default=0)
    p.add_argument('--lt', type=str, default='none', choices=['none'])
    p.add_argument('--bct', type=str2bool, default=0)
    p.add_argument('--class-temp', type=float, default=2.0)
    p.add_argument('--data', type=str, default='synthetic', choices=['synthetic', 'reference'])
    p.add_argument('--steps', type=int, default=0, help='override STAGE3_STEPS (0 = config)')
    p.add_argument('--cuda-graph', type=str2bool, default=1, help='replay the whole step as one CUDA graph')
    p.add_argument('--gene-every', type=int, default=0, help='regenerate the soft pseudo labels with the current model every N '
                   'iterations (the reference\'s GENE_EVERY block, tools/train_ssl_reg.py:180-194); synthetic data: over this rank\'s target tiles, '
                   'written to SNAPSHOT_DIR/pseudo_label/*.pt and read back')
    p.add_argument('--region-bound', type=int, default=0, help='upper bound of region ids + 1 (0 = measured from the data once)')
    return p.parse_args()


class SyntheticLoader:
    """endless stream of seeded synthetic (source, target) batches resident on the device"""

    def __init__(self, cfg, device, seed):
        h, w = cfg.SYNTHETIC["size"]
        self.t = synth.step_inputs(cfg.BATCH_SIZE, h, w, cfg.CLASS_NUM, cfg.SYNTHETIC["regions_per_tile"], device=device, seed=seed)

    def next(self):
        return self.t[:5]


def main():
    args = parse()
    cfg = import_config(args.config_path, create=True)
    rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1)); local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1"); os.environ.setdefault("MASTER_PORT", "29533")
        dist.init_process_group("nccl", device_id=dev)
    seed_torch(2333 + rank)                                   # the reference's seed (tools/train_ssl_reg.py:274), shard-offset
    log = (lambda s: print(s, flush=True)) if rank == 0 else (lambda s: None)

    class_num, ignore_label = cfg.CLASS_NUM, cfg.IGNORE_LABEL
    stop_steps = args.steps or cfg.STAGE3_STEPS
    cfg.NUM_STEPS = stop_steps * 1.5                          # :85
    cfg.PREHEAT_STEPS = max(int(stop_steps / 20), 1)          # :86
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
    cb_t = ClassBalance(class_num=class_num, ignore_label=ignore_label, decay=0.99, temperature=args.class_temp)
    loss_s = CrossEntropy(ignore_label=ignore_label, class_balancer=cb_s if args.bcs else None)
    loss_t = CrossEntropy(ignore_label=ignore_label, class_balancer=cb_t if args.bct else None)

    if args.data == 'reference':
        from ever.core.iterator import Iterator                     # the reference's own loaders (not part of this repo)
        from regda.datasets.daLoader import DALoader
        rcfg = __import__('configs.' + args.config_path, fromlist=['x'])
        src_it, tgt_it = Iterator(DALoader(rcfg.SOURCE_DATA_CONFIG, rcfg.DATASETS)), Iterator(DALoader(rcfg.TARGET_DATA_CONFIG, rcfg.DATASETS))

        def next_batch():
            (xs, ls), (xt, lt) = src_it.next()[0], tgt_it.next()[0]
            return xs.to(dev), ls['cls'].to(dev), xt.to(dev), lt['cls'].to(dev), lt['sup'].to(dev)
    else:
        loader = SyntheticLoader(cfg, dev, 2333 + rank)      # NOTE: one fixed seeded batch per rank, replayed every step
        next_batch = loader.next
        if not args.ckpt_proto:
            aligner.prototypes = loader.t[5].clone()

    first = next_batch()
    bound = args.region_bound or int(first[4].max()) + 1       # one sync at start-up instead of one per step (scatter's index.max())
    hom = Homogenizer(percent=args.percent, class_num=class_num, ignore_label=ignore_label, region_bound=bound, strict=False)
    step = SelfTrainingStep(model, aligner, hom, class_num=class_num, ignore_label=ignore_label, cutoff_top=cfg.CUTOFF_TOP,
                            cutoff_low=cfg.CUTOFF_LOW, refine_temp=args.refine_temp, sam_refine=args.sam_refine,
                            refine_label=bool(args.refine_label), momentum=cfg.MOMENTUM, weight_decay=cfg.WEIGHT_DECAY,
                            loss_fn_s=loss_s, loss_fn_t=loss_t, world_size=world)
    use_graph = bool(args.cuda_graph) and not (args.bcs or args.bct)        # ClassBalance keeps host-visible EMA state
    runner = GraphedStep(step, list(first), lr=0.0) if use_graph else None

    class _Opt:                                                 # adjust_learning_rate's optimizer surface (tools.py:199-207)
        param_groups = [dict(lr=0.0)]

    def regenerate_pseudo_labels():
        """:180-194 -- offline teacher pass with the current weights (8-view TTA, sliding windows), soft labels through the
        reference's on-disk format (<pseudo_label>/<fname>.pt, float32 [C,H,W]) and back into the target batch"""
        path = osp.join(cfg.SNAPSHOT_DIR, 'pseudo_label')
        xs, ls, xt, _, regs = loader.t[:5]
        names = [f'synthetic_r{rank}_{i:04d}.tif' for i in range(xt.shape[0])]
        hw = tuple(xt.shape[-2:])
        tile = (min(512, hw[0]), min(512, hw[1]))
        gener_target_pseudo(cfg, model, [(xt[i:i + 1].float(), {'fname': [n]}) for i, n in enumerate(names)], path, size=hw,
                            save_prob=True, slide=True, ignore_label=ignore_label, num_classes=class_num, tile_size=tile)
        model.train()
        soft = torch.stack([torch.load(osp.join(path, n + '.pt')) for n in names]).to(dev)
        loader.t = (xs, ls, xt, soft, regs) + tuple(loader.t[5:])
        log(f'###### generated {len(names)} soft pseudo labels in {path} ######')

    best = dict(miou=-1.0)

    def checkpoint_and_evaluate(i_iter):
        """:253-265 -- <TARGET_SET>_curr.pth every EVAL_EVERY iterations (and at iteration 0 / the end), evaluate(), and
        <TARGET_SET>_best.pth + prototypes_best.pth (what the downstream stages and tools/eval.py load) when the mIoU improves.
        Data parallel: BatchNorm running statistics are per rank (each rank normalises its own images, like the reference's
        single process); the checkpoint carries their mean over the ranks."""
        import shutil
        from regda_b200.utils.eval import evaluate
        if world > 1:
            for buf in model.buffers():
                if buf.dtype.is_floating_point:
                    dist.all_reduce(buf, op=dist.ReduceOp.SUM)
                    buf.div_(world)
        if rank != 0:
            return
        ckpt = osp.join(cfg.SNAPSHOT_DIR, cfg.TARGET_SET + '_curr.pth')
        torch.save({k: v.detach().cpu() for k, v in model.state_dict().items()}, ckpt)
        torch.save(aligner.prototypes.cpu(), osp.join(cfg.SNAPSHOT_DIR, 'prototypes_curr.pth'))
        if args.data == 'reference':
            from regda.datasets.daLoader import DALoader
            eval_loader = ((ret, ret_gt['cls']) for ret, ret_gt in DALoader(rcfg.EVAL_DATA_CONFIG, rcfg.DATASETS))
            tile = 512
        else:       # synthetic: the labelled source-like tiles stand in for the evaluation split
            xs, ls = loader.t[0], loader.t[1]
            eval_loader = [(xs[i:i + 1], ls[i:i + 1]) for i in range(xs.shape[0])]
            tile = min(512, xs.shape[-2], xs.shape[-1])
        _, miou = evaluate(model, eval_loader, class_num, ignore_label=ignore_label, skip_class0=(cfg.DATASETS == 'IsprsDA'), tile=tile)
        model.train()
        log(f'iter={i_iter + 1}, mIoU = {miou:.5f} (best so far {max(best["miou"], -1.0):.5f})')
        if miou == miou and miou > best['miou']:
            best['miou'] = miou
            shutil.copyfile(ckpt, osp.join(cfg.SNAPSHOT_DIR, cfg.TARGET_SET + '_best.pth'))
            shutil.copyfile(osp.join(cfg.SNAPSHOT_DIR, 'prototypes_curr.pth'), osp.join(cfg.SNAPSHOT_DIR, 'prototypes_best.pth'))

    t0 = time.time()
    os.makedirs(cfg.SNAPSHOT_DIR, exist_ok=True)
    batch = first
    for i_iter in range(stop_steps):
        if args.gene_every > 0 and args.data == 'synthetic' and i_iter % args.gene_every == 0:
            regenerate_pseudo_labels()
            batch = next_batch()
        lr = adjust_learning_rate(_Opt, i_iter, cfg)            # :178
        out = runner(*batch, lr=lr) if runner is not None else step(*batch, lr)
        if i_iter == 0 or (i_iter + 1) % 50 == 0:               # :246-251 (the only host sync: reading the loss to log it)
            log(f"iter={i_iter + 1}, total={float(out['loss']):.3f}, loss_source={float(out['loss_source']):.3f}, "
                f"loss_target={float(out['loss_target']):.3f},, lr = {lr:.3e}")
            hom.check()
        if i_iter == 0 or (i_iter + 1) % cfg.EVAL_EVERY == 0 or (i_iter + 1) >= stop_steps:        # :253-265
            checkpoint_and_evaluate(i_iter)
        batch = next_batch()
    torch.cuda.synchronize()
    dt = time.time() - t0
    imgs = 2 * cfg.BATCH_SIZE * world * stop_steps
    log(f">>>> Using {dt / 3600:.3f} hours, {imgs / dt:.1f} images/s over {world} GPU(s).")
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
