#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_conv_gpu.py -x -q -k "bnred" 2>&1 | tail -25 > gpurun_out/c10_tests.txt
cat gpurun_out/c10_tests.txt
python - <<'PY' 2>&1 | tail -30
import torch, sys
sys.path.insert(0, ".")
from regda_b200.models import Encoder as E
from regda_b200.ops import norm as fnorm
def run(fuse):
    torch.manual_seed(7)
    fnorm.FUSE_BN_BWD = fuse
    blocks = [E.Bottleneck(256, 64), E.Bottleneck(256, 64), E.Bottleneck(256, 64)]
    net = torch.nn.Sequential(*blocks).cuda().train().to(memory_format=torch.channels_last)
    x = torch.randn(4, 256, 24, 40, device="cuda").bfloat16().contiguous(memory_format=torch.channels_last).requires_grad_(True)
    E._GROUPS = 2
    out = net(x)
    g = torch.randn(out.shape, device="cuda").bfloat16().contiguous(memory_format=torch.channels_last)
    out.backward(g)
    E._GROUPS = 1
    return out.detach().float(), x.grad.float(), {n: p.grad.float().clone() for n, p in net.named_parameters()}
a = run(False); b = run(False); c = run(True); d = run(True)
def rep(tag, p, q):
    print(tag, "out", float((p[0]-q[0]).abs().max()), "dx max", float((p[1]-q[1]).abs().max()), "mean", float((p[1]-q[1]).abs().mean()), "scale", float(q[1].abs().max()), float(q[1].abs().mean()))
    for n in list(p[2])[:40]:
        e = float((p[2][n]-q[2][n]).abs().max()); s = float(q[2][n].abs().max())
        if e > 0.03 * s: print("   ", n, e, s)
rep("unfused vs unfused", a, b)
rep("fused vs fused", c, d)
rep("fused vs unfused", c, a)
PY
