"""GPU: the stage-2 step -- PrototypeContrastiveLoss kernels and AlignStep -- against the golden vectors
the reference's own objects produced (tests/golden/align_step_resnet50.npz, tools/train_align_reg.py:144-196) and against the
oracle on random inputs.  Tolerances: the loss kernels 1e-5 relative (float32); the whole step in float32 compute mode 2e-3 on
the losses (north_star: 1e-3 on logits / loss; the step adds two float32 re-associations); bf16 compute mode 5e-2."""
import numpy as np
import pytest
import torch

from conftest import load_golden
from oracle import step_oracle as so

pytestmark = pytest.mark.gpu


def test_pcl_loss_and_gradient_match_reference_fixture():
    from regda_b200.loss import PrototypeContrastiveLoss
    z = load_golden("align_step_resnet50.npz")
    pcl = PrototypeContrastiveLoss(temperature=8.0, ignore_label=-1)
    feat = torch.from_numpy(z["pcl_feat"]).cuda().contiguous(memory_format=torch.channels_last).requires_grad_(True)
    loss = pcl(torch.from_numpy(z["pcl_proto"]).cuda(), feat, torch.from_numpy(z["pcl_label"]).cuda())
    (loss * 0.5).backward()
    assert abs(float(loss) - float(z["pcl_loss"])) <= 1e-5 * abs(float(z["pcl_loss"]))
    torch.testing.assert_close(feat.grad.cpu(), torch.from_numpy(z["pcl_dfeat_half"]), rtol=1e-4, atol=1e-8)
    pcl.check()


@pytest.mark.parametrize("shape", [(8192, 2048, 6), (1000, 512, 7), (33, 64, 3)], ids=str)
def test_pcl_random_vs_oracle(shape):
    from regda_b200.loss import PrototypeContrastiveLoss
    n, k, c = shape
    g = torch.Generator().manual_seed(5)
    proto = torch.randn(c, k, generator=g).abs()
    feat = torch.randn(n, k, generator=g)
    lab = torch.randint(-1, c, (n,), generator=g)
    fr = feat.clone().requires_grad_(True)
    want = so.pcl_loss(proto, fr, lab, 8.0, -1)
    want.backward()
    fg = feat.cuda().requires_grad_(True)
    got = PrototypeContrastiveLoss(8.0, -1)(proto.cuda(), fg, lab.cuda())
    got.backward()
    assert abs(float(got) - float(want)) <= 1e-5 * abs(float(want))
    torch.testing.assert_close(fg.grad.cpu(), fr.grad, rtol=1e-4, atol=1e-9)
    assert float(fg.grad[lab.cuda() == -1].abs().max()) == 0.0


def _objects(dtype, sync_free=False):
    from regda_b200.gast.alignment import Aligner
    from regda_b200.models.Encoder import Deeplabv2
    from regda_b200.trainer import AlignStep
    from regda_b200.utils.local_region_homog import Homogenizer
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    z = load_golden("align_step_resnet50.npz")
    cfg = dict(backbone=dict(resnet_type="resnet50", output_stride=16, pretrained=False), multi_layer=True, cascade=False, use_ppm=True,
               ppm=dict(num_classes=6, use_aux=False, fc_dim=2048), inchannels=2048, num_classes=6, is_ins_norm=True)
    m = Deeplabv2(cfg, compute_dtype=dtype)
    m.load_state_dict(so.seeded_state_dict(m, 2333), strict=True)
    for mod in m.modules():
        if isinstance(mod, torch.nn.Dropout2d):
            mod.p = 0.0
    m = m.cuda().train()
    al = Aligner(None, 2048, 6, -1, 0.996)
    al.prototypes = torch.from_numpy(z["proto"]).cuda()
    hom = (Homogenizer(percent=0.5, class_num=6, ignore_label=-1, region_bound=int(z["regs"].max()) + 1, strict=False) if sync_free
           else Homogenizer(percent=0.5, class_num=6, ignore_label=-1))
    step = AlignStep(m, al, hom, class_num=6, ignore_label=-1, pcl_temp=8.0)
    t = [torch.from_numpy(z[k]).cuda() for k in ("xs", "ls", "xt", "regs")]
    return z, m, al, step, t


def test_align_step_float32_matches_reference():
    z, m, al, step, t = _objects(torch.float32)
    outs = [step(*t, 1e-2) for _ in range(2)]
    want = z["losses"]
    for it in range(2):
        for k, name in enumerate(("loss", "loss_seg", "loss_align", "grad_norm")):
            got = float(outs[it][name])
            tol = 1e-2 if name == "grad_norm" else 2e-3     # the second step's global gradient norm compounds the stem layers' float32 re-association (measured 5e-3)
            assert abs(got - want[it, k]) <= tol * abs(want[it, k]), (it, name, got, want[it, k])
    assert (outs[0]["hard"].cpu().numpy() != z["hard_0"]).mean() < 2e-3
    # prototypes are O(1) values moved by 0.4 % of a class mean per step: 1e-4 of their scale absolute (near-zero entries), 2e-3 relative
    torch.testing.assert_close(al.prototypes.cpu(), torch.from_numpy(z["proto_after"]), rtol=2e-3, atol=1e-4)
    rel = float((m.encoder.resnet.conv1.weight.detach().cpu() - torch.from_numpy(z["conv1_after"])).abs().max() / torch.from_numpy(z["conv1_after"]).abs().max())
    # two clipped SGD steps move the stem weights by ~30 % of their scale each (random init, gradient norm ~200 clipped to 32); the
    # stem gradient itself carries up to 8e-2 of float32 re-association noise between the GPU library kernels and the CPU
    # reference (tests/test_step_gpu.py::test_model_float32_matches_reference), which is what shows up here (measured 3e-2)
    assert rel < 6e-2


def test_align_step_bf16_and_cuda_graph():
    from regda_b200.trainer import GraphedStep
    z, m, al, step, t = _objects(torch.bfloat16)
    o = step(*t, 1e-2)
    want = z["losses"][0]
    assert abs(float(o["loss"]) - want[0]) <= 5e-2 * abs(want[0])
    assert abs(float(o["loss_align"]) - want[2]) <= 5e-2 * abs(want[2])
    assert torch.isfinite(step.arena.param).all()
    z2, m2, al2, step2, t2 = _objects(torch.bfloat16, sync_free=True)
    g = GraphedStep(step2, t2, lr=0.0, warmup=2)
    losses = [float(g(*t2, lr=1e-2)["loss"]) for _ in range(3)]
    assert all(np.isfinite(losses))
    assert abs(losses[0] - float(o["loss"])) <= 3e-2 * abs(float(o["loss"]))


@pytest.mark.parametrize("precise", [True, False])
def test_coral_loss_and_gradients_match_reference_fixture(precise):
    """CoralLoss on the tcgen05 kernels (covariances = the weight-gradient contraction, backward = a 1x1 forward convolution) against
    the reference's own CoralLoss values and gradients (tests/golden/coral.npz): float32-accuracy mode at 1e-4, bf16-operand mode
    (what bf16 training uses) at 2e-2 of the gradient scale."""
    from regda_b200.gast.coral import CoralLoss
    z = load_golden("coral.npz")
    for k in range(3):
        src = torch.from_numpy(z[f"loss{k}/src"]).cuda().requires_grad_(True)
        tgt = torch.from_numpy(z[f"loss{k}/tgt"]).cuda().requires_grad_(True)
        loss = CoralLoss(is_sqrt=bool(z[f"loss{k}/is_sqrt"]), precise=precise)(src, tgt)
        loss.backward()
        want = float(z[f"loss{k}/loss"])
        ltol, gtol = (1e-4, 2e-4) if precise else (1e-2, 2e-2)
        assert abs(float(loss) - want) <= ltol * abs(want), (k, float(loss), want)
        for got, key in ((src.grad, "dsrc"), (tgt.grad, "dtgt")):
            ref = torch.from_numpy(z[f"loss{k}/{key}"]).cuda()
            assert float((got - ref).abs().max()) <= gtol * float(ref.abs().max()), (k, key)


def test_align_step_with_coral_matches_reference_fixture():
    """two iterations of the stage-2 step with --align-domain 1 (tools/train_align_reg.py:187), float32 compute, against the
    reference's own objects (tests/golden/coral.npz)"""
    from regda_b200.gast.alignment import Aligner
    from regda_b200.models.Encoder import Deeplabv2
    from regda_b200.ops import conv as C
    from regda_b200.trainer import AlignStep
    from regda_b200.utils.local_region_homog import Homogenizer
    from oracle import step_oracle as so
    z = load_golden("coral.npz")
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    cfg = dict(backbone=dict(resnet_type="resnet50", output_stride=16, pretrained=False), multi_layer=True, cascade=False, use_ppm=True,
               ppm=dict(num_classes=6, use_aux=False, fc_dim=2048), inchannels=2048, num_classes=6, is_ins_norm=True)
    m = Deeplabv2(cfg, compute_dtype=torch.float32)
    m.load_state_dict(so.seeded_state_dict(m, 2333), strict=True)
    for mod in m.modules():
        if isinstance(mod, torch.nn.Dropout2d):
            mod.p = 0.0
    m = m.cuda().train()
    al = Aligner(None, 2048, 6, -1, 0.996)
    al.prototypes = torch.from_numpy(z["proto"]).cuda()
    hom = Homogenizer(percent=0.5, class_num=6, ignore_label=-1)
    step = AlignStep(m, al, hom, class_num=6, ignore_label=-1, align_domain=True)
    t = [torch.from_numpy(z[k]).cuda() for k in ("xs", "ls", "xt", "regs")]
    before = dict(C.stats)
    want = z["losses"]                     # [it][total, seg, align, domain, grad_norm]
    for it in range(2):
        o = step(*t, 1e-2)
        got = [float(o["loss"]), float(o["loss_seg"]), float(o["loss_align"]), float(o["loss_domain"]), float(o["grad_norm"])]
        for k, name in enumerate(("loss", "loss_seg", "loss_align", "loss_domain", "grad_norm")):
            tol = 5e-3 if name in ("grad_norm", "loss_domain") else 2e-3
            assert abs(got[k] - want[it, k]) <= tol * abs(want[it, k]), (it, name, got[k], want[it, k])
    assert C.stats["cudnn"] == before["cudnn"]
    torch.testing.assert_close(al.prototypes.cpu(), torch.from_numpy(z["proto_after"]), rtol=1e-3, atol=5e-5)
