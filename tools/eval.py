#!/usr/bin/env python
"""Evaluation -- the reference's tools/eval.py (CLI :17-25) + regda/utils/eval.py:14-56 on the B200 inference kernels.

    python tools/eval.py --config-path st.regda.2potsdam --ckpt-path <model .pth> [--tta 1] [--test 1]

Loads a checkpoint with the reference's state_dict keys, runs eval-mode sliding-window prediction (optionally with the 8-view
TTA of tta_predict) and reports per-class IoU / mIoU from a device-side confusion matrix (IsprsDA drops class 0, eval.py:16-17).
Data: `--data synthetic` evaluates seeded synthetic tiles (the reference's DALoader is CPU file I/O, out of scope);
`--data reference` uses the reference's own DALoader(EVAL_DATA_CONFIG / TEST_DATA_CONFIG) when that package is importable."""
from __future__ import annotations

import argparse
import os
import os.path as osp
import sys

ROOT = osp.dirname(osp.dirname(osp.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def str2bool(v):
    return str(v).lower() in ("1", "true", "yes", "y", "t")


def parse(argv=None):
    p = argparse.ArgumentParser(description='Run predict methods.')
    p.add_argument('--config-path', type=str, default='st.regda.2potsdam', help='config path')
    p.add_argument('--ckpt-path', type=str, default='', help='ckpt path (reference state_dict keys)')
    p.add_argument('--multi-layer', type=str2bool, default=True, choices=[True])
    p.add_argument('--ins-norm', type=str2bool, default=True)
    p.add_argument('--test', type=str2bool, default=False, help='evaluate the test set?')
    p.add_argument('--tta', type=str2bool, default=False, help='8-view test-time augmentation')
    p.add_argument('--data', type=str, default='synthetic', choices=['synthetic', 'reference'])
    p.add_argument('--tiles', type=int, default=8, help='synthetic data: number of tiles')
    return p.parse_args(argv)


def main(argv=None):
    args = parse(argv)
    import torch

    from regda_b200 import synth
    from regda_b200.models.Encoder import Deeplabv2
    from regda_b200.utils.eval import evaluate
    from regda_b200.utils.tools import import_config, seed_torch
    seed_torch(2333)
    if not torch.cuda.is_available():
        raise SystemExit("tools/eval.py needs a CUDA device: regda_b200 has no CPU fallback")
    cfg = import_config(args.config_path, create=False)
    class_num = cfg.CLASS_NUM
    model_name = str(cfg.MODEL).lower()
    model_name = 'resnet50' if model_name == 'resnet' else model_name
    model = Deeplabv2(dict(backbone=dict(resnet_type=model_name, output_stride=16, pretrained=False), multi_layer=True, cascade=False,
                           use_ppm=True, ppm=dict(num_classes=class_num, use_aux=False, fc_dim=2048), inchannels=2048,
                           num_classes=class_num, is_ins_norm=args.ins_norm))
    if args.ckpt_path:
        model.load_state_dict(torch.load(args.ckpt_path, map_location='cpu'), strict=True)
        print(f'[Load params] from {args.ckpt_path}')
    else:
        print('WARNING: no --ckpt-path: evaluating random weights')
    model = model.cuda()
    tile = 512
    if args.data == 'reference':
        from regda.datasets.daLoader import DALoader                 # the reference's own loaders (not part of this repo)
        rcfg = __import__('configs.' + args.config_path, fromlist=['x'])
        ref_loader = DALoader(rcfg.TEST_DATA_CONFIG if args.test else rcfg.EVAL_DATA_CONFIG, rcfg.DATASETS)
        loader = ((ret, ret_gt['cls']) for ret, ret_gt in ref_loader)
    else:
        h, w = cfg.SYNTHETIC["size"]
        xs, ls, *_ = synth.step_inputs(args.tiles, h, w, class_num, cfg.SYNTHETIC["regions_per_tile"], device="cuda", seed=2333)
        loader = [(xs[i:i + 1], ls[i:i + 1]) for i in range(args.tiles)]
        # tiles smaller than the reference's 512 window would hit its pad_image quirk (tools.py:56 pads the TOP of the height and
        # pre_slide then crops the padding, see regda_b200/utils/tools.py): evaluate synthetic tiles with a window of their own size
        tile = min(tile, h, w)
    tb, miou = evaluate(model, loader, class_num, ignore_label=cfg.IGNORE_LABEL, skip_class0=True, tile=tile, tta=args.tta)
    print(tb)
    print("IoU per class: " + ", ".join(f"{v:.5f}" for v in tb.iou_per_class) + f"; mIoU = {miou:.5f}")
    return miou


if __name__ == '__main__':
    main()
