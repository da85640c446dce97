"""GPU: the model and the whole self-training step against golden vectors produced by the
unmodified reference (tests/golden/model_*.npz, step_resnet50.npz).

float32 compute (parity mode): logits / loss within 1e-3 relative (north_star tolerance) -- every convolution of this mode
runs on the hand-written tcgen05 kernels (operands split into bf16 hi + lo parts, ops/tc.py), asserted through the dispatch
counters: no library convolution takes part;
bf16 compute (benchmark mode): checked against the bf16-emulating oracle in tests/test_bf16_parity_gpu.py."""
import numpy as np
import pytest
import torch

from conftest import load_golden
from oracle import step_oracle as so

pytestmark = pytest.mark.gpu


def _cfg(rt, C=6):
    return dict(backbone=dict(resnet_type=rt, output_stride=16, pretrained=True), multi_layer=True, cascade=False, use_ppm=True,
                ppm=dict(num_classes=C, use_aux=False, fc_dim=2048), inchannels=2048, num_classes=C, is_ins_norm=True)


def _model(rt, dtype):
    from regda_b200.models.Encoder import Deeplabv2
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    m = Deeplabv2(_cfg(rt), compute_dtype=dtype)
    m.load_state_dict(so.seeded_state_dict(m, 2333), strict=True)
    for mod in m.modules():
        if isinstance(mod, torch.nn.Dropout2d):
            mod.p = 0.0
    return m.cuda()


def _rel(a, b):
    return float((a - b).abs().max() / b.abs().max())


@pytest.mark.parametrize("rt", ["resnet50", "resnet101"])
def test_model_float32_matches_reference(rt):
    from regda_b200.gast.balance import CrossEntropy
    from regda_b200.utils.tools import loss_calc
    from regda_b200.ops import conv as C
    z = load_golden(f"model_{rt}.npz")
    m = _model(rt, torch.float32)
    x = torch.from_numpy(z["x"]).cuda()
    before = dict(C.stats)
    m.eval()
    with torch.no_grad():
        p = m(x)
    assert _rel(p.cpu(), torch.from_numpy(z["eval_prob"])) < 1e-3
    m.train()
    x1, x2, feat = m(x)
    assert x1.dtype == torch.float32 and feat.shape == (2, 2048, x.shape[2] // 16, x.shape[3] // 16)
    assert _rel(x1.detach().cpu(), torch.from_numpy(z["x1"])) < 1e-3
    assert _rel(x2.detach().cpu(), torch.from_numpy(z["x2"])) < 1e-3
    assert _rel(feat.detach().cpu(), torch.from_numpy(z["feat"])) < 1e-3
    loss = loss_calc([x1, x2], torch.from_numpy(z["label"]).cuda(), CrossEntropy(-1), multi=True)
    loss.backward()
    # the 1e-3 gate ran on the repo's own kernels: every convolution (stem, bottlenecks, PPM branches, fuse conv) forward,
    # data gradient and weight gradient went through tcgen05, none through the library
    n_convs = sum(1 for mod in m.modules() if isinstance(mod, C.Conv2d)) - 2          # the two classifiers are ops/head.py
    assert C.stats["cudnn"] == before["cudnn"]
    assert C.stats["tcgen05_fprop"] - before["tcgen05_fprop"] >= 2 * (n_convs - 1)    # eval + train forward (stem counted apart)
    assert C.stats["tcgen05_wgrad"] - before["tcgen05_wgrad"] >= n_convs - 1
    assert abs(float(loss) - float(z["loss"])) < 1e-3 * abs(float(z["loss"]))
    names = [str(n) for n in z["grad_names"]]
    got = dict(m.named_parameters())
    assert names == list(got.keys())
    norms = np.array([float(got[n].grad.norm()) for n in names])
    bad = np.abs(norms - z["grad_norms"]) > 5e-2 * np.abs(z["grad_norms"]) + 1e-5 * z["grad_norms"].max()
    assert not bad.any(), [(names[i], norms[i], z["grad_norms"][i]) for i in np.nonzero(bad)[0][:8]]
    # the stem gradient has crossed every layer of the backward pass: float32 re-association between the GPU library
    # kernels and the CPU reference shows up here first (5e-2 for ResNet-101 with batch 2 and random weights)
    assert _rel(m.encoder.resnet.conv1.weight.grad.cpu(), torch.from_numpy(z["grad_conv1"])) < 8e-2
    assert _rel(m.layer5.conv_last[4].weight.grad.cpu(), torch.from_numpy(z["grad_cls5"])) < 1e-3
    torch.testing.assert_close(m.encoder.resnet.bn1.running_mean.cpu(), torch.from_numpy(z["bn1_running_mean"]), rtol=1e-4, atol=1e-6)
    torch.testing.assert_close(m.encoder.resnet.bn1.running_var.cpu(), torch.from_numpy(z["bn1_running_var"]), rtol=1e-4, atol=1e-6)


def _step_objects(dtype, sync_free=False):
    from regda_b200.gast.alignment import Aligner
    from regda_b200.trainer import SelfTrainingStep
    from regda_b200.utils.local_region_homog import Homogenizer
    z = load_golden("step_resnet50.npz")
    m = _model("resnet50", dtype).train()
    al = Aligner(None, 2048, 6, -1, 0.996)
    al.prototypes = torch.from_numpy(z["proto"]).cuda()
    # sync_free: the region-id bound is given up front and domain errors are polled with check(), so nothing in the
    # step synchronises the host (required for CUDA-graph capture)
    hom = (Homogenizer(percent=0.5, class_num=6, ignore_label=-1, region_bound=int(z["regs"].max()) + 1, strict=False) if sync_free
           else Homogenizer(percent=0.5, class_num=6, ignore_label=-1))
    step = SelfTrainingStep(m, al, hom, class_num=6, ignore_label=-1)
    t = [torch.from_numpy(z[k]).cuda() for k in ("xs", "ls", "xt", "soft", "regs")]
    return z, m, al, step, t


def test_full_step_float32_matches_reference():
    from regda_b200.ops import conv as C
    z, m, al, step, t = _step_objects(torch.float32)
    before = dict(C.stats)
    outs = [step(*t, 1e-2) for _ in range(2)]
    assert C.stats["cudnn"] == before["cudnn"] and C.stats["tcgen05_fprop"] > before["tcgen05_fprop"] + 100
    want = z["losses"]
    for it in range(2):
        for k, name in enumerate(("loss", "loss_source", "loss_target", "grad_norm")):
            got = float(outs[it][name])
            tol = 5e-3 if name == "grad_norm" else 2e-3      # the global gradient norm carries the stem layers' noise
            assert abs(got - want[it, k]) <= tol * abs(want[it, k]), (it, name, got, want[it, k])
    z0 = load_golden("step_resnet50.npz")
    hard = outs[0]["hard"].cpu().numpy()
    assert (hard != z0["hard_0"]).mean() < 2e-3        # threshold-adjacent pixels may flip with float re-association
    # two EMA steps of 0.004 x (class mean of the normalised features, |feat| <= ~4 matched at 1e-3 of that scale)
    torch.testing.assert_close(al.prototypes.cpu(), torch.from_numpy(z["proto_after"]), rtol=1e-3, atol=5e-5)
    # the stem weight after two clipped SGD steps inherits the percent-level float32 re-association noise of the stem
    # gradient (see test_model_float32_matches_reference) scaled by lr; the library's float32 convolution algorithms are
    # not run-to-run deterministic (observed 2e-3 .. 6e-3 of the weight scale over runs of the same code)
    # (measured with the float32 path on the tcgen05 kernels: 0.6e-2 .. 2.8e-2)
    assert _rel(m.encoder.resnet.conv1.weight.detach().cpu(), torch.from_numpy(z["conv1_after"])) < 4e-2
    assert _rel(m.layer6.conv_last[4].weight.detach().cpu(), torch.from_numpy(z["cls6_after"])) < 1e-3


def test_full_step_bf16_runs_and_stays_close():
    z, m, al, step, t = _step_objects(torch.bfloat16)
    o = step(*t, 1e-2)
    want = z["losses"][0]
    assert abs(float(o["loss"]) - want[0]) <= 5e-2 * abs(want[0])
    assert torch.isfinite(step.arena.param).all()


def test_cuda_graph_replay_matches_eager():
    """the graph warm-up must leave no trace.  A GraphedStep built with a NON-ZERO warm-up learning rate starts
    from exactly the state it was given (weights, momentum, prototypes, BatchNorm running statistics and counters), and its
    first replayed step moves the weights like the first eager step (momentum buffer = gradient, not 2.7x the gradient as
    after three warm-up accumulations).  bf16 on 64x64 tiles is chaotic (BatchNorm over 32 samples amplifies the fp32-atomic
    last bits), so losses / weight updates are compared at percent level; the exact statements are about state."""
    from regda_b200.trainer import GraphedStep
    z, m, al, step, t = _step_objects(torch.bfloat16, sync_free=True)
    z2, m2, al2, step2, t2 = _step_objects(torch.bfloat16, sync_free=True)
    w0 = step2.arena.param.clone()
    bn0 = {k: v.clone() for k, v in m2.state_dict().items() if "running_" in k}
    g = GraphedStep(step2, t2, lr=1e-2, warmup=3)        # real steps at lr 1e-2 during warm-up: everything must be put back
    assert torch.equal(step2.arena.param, w0) and torch.equal(step2.arena.param_bf16, w0.bfloat16())
    assert float(step2.arena.momentum.abs().max()) == 0.0
    assert torch.equal(al2.prototypes, al.prototypes)
    assert int(m2.encoder.resnet.bn1.num_batches_tracked) == 0
    assert all(torch.equal(v, bn0[k]) for k, v in m2.state_dict().items() if "running_" in k)
    og = g(*t2, lr=1e-2)
    oe = step(*t, 1e-2)
    lg, le = float(og["loss"]), float(oe["loss"])
    assert abs(lg - le) <= 2e-2 * abs(le), (lg, le)
    dw_g, dw_e = step2.arena.param - w0, step.arena.param - w0
    ratio = float(dw_g.norm() / dw_e.norm())
    assert 0.85 <= ratio <= 1.18, ratio                  # the stale-momentum bug gave 3.4
    cos = float((dw_g * dw_e).sum() / (dw_g.norm() * dw_e.norm()))
    assert cos > 0.8, cos                                # (measured 0.90; a wrong update direction is ~0)
    torch.testing.assert_close(al2.prototypes, al.prototypes, rtol=5e-2, atol=1e-4)
    for _ in range(2):
        og = g(*t2, lr=1e-2)
        step(*t, 1e-2)
    assert np.isfinite(float(og["loss"]))
    bn_g, bn_e = m2.encoder.resnet.layer3[0].bn2, m.encoder.resnet.layer3[0].bn2
    assert int(bn_g.num_batches_tracked) == int(bn_e.num_batches_tracked) == 6        # 3 steps x 2 domain batches


def test_state_dict_abi():
    from regda_b200.models.Encoder import Deeplabv2
    m = Deeplabv2(_cfg("resnet101"))
    sd = m.state_dict()
    assert len(sd) == 688
    o = so.DeeplabOracle("resnet101", 6)
    assert list(sd.keys()) == list(o.state_dict().keys())
    for (k, a), b in zip(sd.items(), o.state_dict().values()):
        assert a.shape == b.shape, k


def test_step_at_config_L_geometry():
    """BASELINE.json configs[3] geometry (ResNet-50, 7 classes, 1024x1024 tiles): 256x256 layer1 maps (two 128-pixel TMA patches
    per row), 64x64 feature maps, megapixel LRH.  One bf16 step on 2 + 2 tiles runs on the hand-written kernels only, and the
    pseudo-label chain on its tensors matches the oracle (LRH bit-exact on the selected labels)."""
    import bench_step
    from oracle import cbind
    from regda_b200.ops import conv as C
    dev = torch.device("cuda", 0)
    model, step, runner, tensors = bench_step.build(dev, 1, resnet="resnet50", use_graph=False, n_regions=300, batch=2, hw=(1024, 1024), classes=7)
    for mod in model.modules():
        if isinstance(mod, torch.nn.Dropout2d):
            mod.p = 0.0
    before = dict(C.stats)
    proto0 = step.aligner.prototypes.clone()
    o = step(*tensors, 0.0)
    assert np.isfinite(float(o["loss"])) and C.stats["cudnn"] == before["cudnn"]
    assert o["hard"].shape == (2, 1024, 1024)
    images_s, label_s, images_t, soft_t, regs_t = tensors
    with torch.no_grad():
        (_, _, _), (pt1, pt2, feat_t) = model.forward_pair(images_s, images_t)
        assert feat_t.shape == (2, 2048, 64, 64)
        step.aligner.prototypes.copy_(proto0)
        sel = step.aligner.refine_select(feat_t, [pt1, pt2], soft_t, 2.0, 0.8, 0.6)
        hard = step.homogenizer(sel, regs_t.squeeze(1))
    want_sel = so.pseudo_select(so.label_refine(feat_t.float().cpu(), [pt1.float().cpu(), pt2.float().cpu()], soft_t.cpu(), proto0.cpu(), 2.0),
                                0.8, 0.6, -1)
    assert float((sel.cpu() != want_sel).float().mean()) < 2e-3
    want = cbind.lrh(sel.cpu().numpy(), regs_t.squeeze(1).cpu().numpy(), 7, -1, 0.5)
    assert np.array_equal(hard.cpu().numpy(), want)
