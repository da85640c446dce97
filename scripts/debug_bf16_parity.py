"""Layer-by-layer comparison of the bf16 CUDA model with the bf16-emulating oracle on a golden input (debug aid)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from oracle import bf16_emul as be, step_oracle as so
from regda_b200.models.Encoder import Deeplabv2

rt = sys.argv[1] if len(sys.argv) > 1 else "resnet50"
z = np.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", f"model_{rt}.npz"))
cfg = dict(backbone=dict(resnet_type=rt, output_stride=16, pretrained=False), multi_layer=True, cascade=False, use_ppm=True,
           ppm=dict(num_classes=6, use_aux=False, fc_dim=2048), inchannels=2048, num_classes=6, is_ins_norm=True)
m = Deeplabv2(cfg, compute_dtype=torch.bfloat16)
sd = so.seeded_state_dict(m, 2333)
m.load_state_dict(sd)
o = so.DeeplabOracle(rt, 6, dropout=0.0)
o.load_state_dict(sd)
for mod in m.modules():
    if isinstance(mod, torch.nn.Dropout2d):
        mod.p = 0.0
m = m.cuda().train(); o.train()
rec = {}
rn = m.encoder.resnet
rn.bn1.register_forward_hook(lambda mod, i, out: None)
for li in range(1, 5):
    for bi, blk in enumerate(getattr(rn, f"layer{li}")):
        blk.register_forward_hook(lambda mod, i, out, k=f"layer{li}.{bi}": rec.__setitem__(k, out.detach().float().cpu()))
x = torch.from_numpy(z["x"])
x1, x2, feat = m(x.cuda())
taps = {}
w1, w2, wf = be.forward_train(o, x, taps=taps)
def rel(a, b):
    e = (a - b).abs()
    return f"max {float(e.max() / b.abs().max()):.2e} mean {float(e.mean() / b.abs().mean()):.2e} frac>1ulp {float((e > 0.008 * b.abs().clamp_min(1e-3)).float().mean()):.3f}"
for k in taps:
    if k in rec:
        print(f"{k:12s} {tuple(rec[k].shape)} {rel(rec[k], taps[k])}")
print("feat", rel(feat.detach().float().cpu(), wf))
print("x1  ", rel(x1.detach().float().cpu(), w1))
print("x2  ", rel(x2.detach().float().cpu(), w2))
