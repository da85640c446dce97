"""GPU: the BENCHMARKED configuration (bf16 compute, tcgen05 convolutions, fused BatchNorm, PPM / classifier kernels) against
the bf16-emulating oracle (oracle/bf16_emul.py) on the reference-generated golden inputs.

The emulation is the reference's forward with exact arithmetic between, and a bf16 cast at, the points where the kernels round
(pinned to the reference's float32 goldens with the casts off: tests/test_oracle_golden.py).  What is left between it and the
CUDA path is fp32 accumulation order -- which moves a few values across a bf16 rounding boundary per layer (one bf16 ulp =
4e-3 relative on that element).  Tolerances: logits / features within 1e-2 of the tensor's scale at the worst element and 1e-3
on average; the loss within 2e-3 relative.  Every convolution must have run on the hand-written kernels."""
import numpy as np
import pytest
import torch

from conftest import load_golden
from oracle import bf16_emul as be
from oracle import step_oracle as so

pytestmark = pytest.mark.gpu


def _cfg(rt, C=6):
    return dict(backbone=dict(resnet_type=rt, output_stride=16, pretrained=False), multi_layer=True, cascade=False, use_ppm=True,
                ppm=dict(num_classes=C, use_aux=False, fc_dim=2048), inchannels=2048, num_classes=C, is_ins_norm=True)


def _models(rt):
    from regda_b200.models.Encoder import Deeplabv2
    m = Deeplabv2(_cfg(rt), compute_dtype=torch.bfloat16)
    sd = so.seeded_state_dict(m, 2333)
    m.load_state_dict(sd, strict=True)
    o = so.DeeplabOracle(rt, 6, dropout=0.0)
    o.load_state_dict(sd, strict=True)
    for mod in m.modules():
        if isinstance(mod, torch.nn.Dropout2d):
            mod.p = 0.0
    return m.cuda().train(), o.train()


def _check(got, want, name, worst=1e-2, mean=1e-3):
    got, want = got.detach().float().cpu(), want.float()
    err = (got - want).abs()
    scale = float(want.abs().max())
    assert float(err.max()) <= worst * scale, (name, float(err.max()) / scale)
    assert float(err.mean()) <= mean * float(want.abs().mean()) + 1e-7, (name, float(err.mean()) / float(want.abs().mean()))


@pytest.mark.parametrize("rt", ["resnet50", "resnet101"])
def test_bf16_tcgen05_model_matches_bf16_emulating_oracle(rt):
    from regda_b200.gast.balance import CrossEntropy
    from regda_b200.ops import conv as C
    from regda_b200.utils.tools import loss_calc
    z = load_golden(f"model_{rt}.npz")
    m, o = _models(rt)
    x = torch.from_numpy(z["x"])
    before = dict(C.stats)
    x1, x2, feat = m(x.cuda())
    n_convs = sum(1 for mod in m.modules() if isinstance(mod, C.Conv2d)) - 3          # stem + two classifiers have their own ops
    assert C.stats["cudnn"] == before["cudnn"] and C.stats["tcgen05_fprop"] - before["tcgen05_fprop"] == n_convs + 1     # (+ the stem's 1x1)
    w1, w2, wfeat = be.forward_train(o, x)
    _check(feat, wfeat, "feat")
    _check(x1, w1, "x1")
    _check(x2, w2, "x2")
    label = torch.from_numpy(z["label"])
    loss = loss_calc([x1, x2], label.cuda(), CrossEntropy(-1), multi=True)
    want = so.ce_loss_multi([w1, w2], label, -1)
    assert abs(float(loss) - float(want)) <= 2e-3 * abs(float(want)), (float(loss), float(want))
    # for scale: the same bf16 path against the FLOAT32 reference golden is percents away (bf16 arithmetic, not kernel error)
    ref = torch.from_numpy(z["x1"])
    assert float((x1.detach().cpu() - ref).abs().max()) > 3 * float((x1.detach().cpu() - w1).abs().max())


def test_bf16_paired_forward_matches_bf16_emulating_oracle():
    """Deeplabv2.forward_pair (source + target batch as one tensor, BatchNorm statistics per domain -- the benchmarked form) on
    the golden step inputs: equals the emulation with two statistics groups, i.e. the reference's two separate model calls"""
    z = load_golden("step_resnet50.npz")
    m, o = _models("resnet50")
    xs, xt = torch.from_numpy(z["xs"]), torch.from_numpy(z["xt"])
    (s1, s2, fs), (t1, t2, ft) = m.forward_pair(xs.cuda(), xt.cuda())
    w1, w2, wf = be.forward_train(o, torch.cat([xs, xt], 0), groups=2)
    b = xs.shape[0]
    _check(torch.cat([fs, ft]), wf, "feat")
    _check(torch.cat([s1, t1]), w1, "x1")
    _check(torch.cat([s2, t2]), w2, "x2")
    # two groups == two separate calls in the emulation as well
    a1, _, _ = be.forward_train(o, xs)
    assert torch.equal(a1, w1[:b])
