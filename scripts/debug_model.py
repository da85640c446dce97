"""Layer-by-layer comparison of the fused bf16 model against the eager bf16 model (same weights, same input)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import step_oracle as so
from regda_b200.models import Encoder as E
from regda_b200.ops import conv as C

cfg = dict(backbone=dict(resnet_type="resnet50", output_stride=16, pretrained=False), multi_layer=True, cascade=False, use_ppm=True,
           ppm=dict(num_classes=6, use_aux=False, fc_dim=2048), inchannels=2048, num_classes=6, is_ins_norm=True)
torch.manual_seed(0)
x = torch.randn(4, 3, 256, 256, device="cuda").clamp(max=1.0)
acts = {}


def run(tag, fused, engine, dtype=torch.bfloat16):
    E.set_fused(fused)
    C.set_engine(engine)
    m = E.Deeplabv2(cfg, compute_dtype=dtype)
    m.load_state_dict(so.seeded_state_dict(m, 2333), strict=True)
    for mod in m.modules():
        if isinstance(mod, torch.nn.Dropout2d):
            mod.p = 0.0
    m = m.cuda().train()
    rec = {}
    hooks = []
    for name, mod in m.named_modules():
        if name.count(".") <= 3 and name:
            hooks.append(mod.register_forward_hook(lambda mod, i, o, name=name: rec.__setitem__(name, o.detach().float()) if torch.is_tensor(o) else None))
    x1, x2, feat = m(x)
    rec["x1"], rec["x2"], rec["feat"] = x1.detach().float(), x2.detach().float(), feat.detach().float()
    acts[tag] = rec


run("f32", False, "cudnn", torch.float32)
run("eager", False, "cudnn")
run("tc", False, "auto")
run("fused_cudnn", True, "cudnn")
run("fused", True, "auto")
names = list(acts["f32"].keys())
print(f"{'module':40s} {'eager':>10s} {'tc':>10s} {'fusedcudnn':>10s} {'fused':>10s}   (max abs err vs f32 / max abs f32)")
for n in names:
    a = acts["f32"][n]
    row = []
    for t in ("eager", "tc", "fused_cudnn", "fused"):
        b = acts[t].get(n)
        row.append(float((a - b).abs().max() / a.abs().max()) if b is not None and b.shape == a.shape else float("nan"))
    if n.count(".") <= 2 or max(r for r in row if r == r) > 0.05:
        print(f"{n:40s} " + " ".join(f"{r:10.4f}" for r in row))
