#!/usr/bin/env python
"""Class prototypes from one pass over source batches -- the reference's tools/init_prototypes.py:50-113:
model(images_s) -> aligner.update_avg(feat_s, label_s) (running per-class feature sums, alignment.py:107-119) ->
aligner.init_avg() -> torch.save(prototypes.cpu()).  Same kernels as the step's prototype update
(regda_class_sums(accumulate=1) + regda_prototype_init_avg).  `--data synthetic` iterates seeded synthetic batches."""
from __future__ import annotations

import argparse
import os.path as osp
import sys

ROOT = osp.dirname(osp.dirname(osp.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

from regda_b200 import synth  # noqa: E402
from regda_b200.gast.alignment import Aligner  # noqa: E402
from regda_b200.models.Encoder import Deeplabv2  # noqa: E402
from regda_b200.utils.tools import import_config, seed_torch  # noqa: E402


def main():
    p = argparse.ArgumentParser()
    p.add_argument('--config-path', type=str, default='st.regda.2potsdam')
    p.add_argument('--ckpt-model', type=str, default='')
    p.add_argument('--out', type=str, default='')
    p.add_argument('--batches', type=int, default=8)
    args = p.parse_args()
    cfg = import_config(args.config_path, create=True)
    seed_torch(2333)
    dev = torch.device("cuda", 0)
    name = str(cfg.MODEL).lower()
    model = Deeplabv2(dict(backbone=dict(resnet_type='resnet50' if name == 'resnet' else name, output_stride=16, pretrained=False),
                           multi_layer=True, cascade=False, use_ppm=True, ppm=dict(num_classes=cfg.CLASS_NUM, use_aux=False, fc_dim=2048),
                           inchannels=2048, num_classes=cfg.CLASS_NUM, is_ins_norm=True))
    if args.ckpt_model:
        model.load_state_dict(torch.load(args.ckpt_model, map_location='cpu'), strict=True)
    model = model.to(dev).train()                    # the reference leaves the model in train mode here (init_prototypes.py:96-101)
    aligner = Aligner(None, 2048, cfg.CLASS_NUM, cfg.IGNORE_LABEL, 0.996, device=dev)
    h, w = cfg.SYNTHETIC["size"]
    with torch.no_grad():
        for i in range(args.batches):
            xs, ls = synth.step_inputs(cfg.BATCH_SIZE, h, w, cfg.CLASS_NUM, cfg.SYNTHETIC["regions_per_tile"], device=dev, seed=2333 + i)[:2]
            _, _, feat = model(xs)
            aligner.update_avg(feat, ls)
    aligner.init_avg()
    out = args.out or osp.join(cfg.SNAPSHOT_DIR, 'prototypes_best.pth')
    torch.save(aligner.prototypes.cpu(), out)
    print('prototypes', tuple(aligner.prototypes.shape), '->', out)


if __name__ == '__main__':
    main()
