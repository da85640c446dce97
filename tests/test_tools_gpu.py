"""GPU: the host-surface pieces around the step -- EMA mirror (regda/utils/ema.py), evaluate()/mIoU (regda/utils/eval.py +
gast/metrics.py), and the trainer / prototype tools end to end on the tiny synthetic config."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


def test_ema_update_apply_restore():
    from regda_b200.ops.conv import Conv2d
    from regda_b200.utils.ema import ExponentialMovingAverage
    torch.manual_seed(0)
    m = torch.nn.Sequential(Conv2d(8, 16, 3, padding=1, bias=True), torch.nn.BatchNorm2d(16)).cuda()
    ema = ExponentialMovingAverage(m, 0.9)
    ema.register()
    before = {n: p.detach().clone() for n, p in m.named_parameters()}
    with torch.no_grad():
        for p in m.parameters():
            p.add_(torch.randn_like(p))
    ema.update()
    for n, p in m.named_parameters():
        want = 0.1 * p.detach() + 0.9 * before[n]                      # ema.py:49-50
        torch.testing.assert_close(ema.shadow[n], want, rtol=1e-6, atol=1e-7)
    cur = {n: p.detach().clone() for n, p in m.named_parameters()}
    ema.apply_shadow()
    for n, p in m.named_parameters():
        torch.testing.assert_close(p.detach(), ema.shadow[n])
    ema.restore()
    for n, p in m.named_parameters():
        assert torch.equal(p.detach(), cur[n])


def test_miou_matches_reference_formula():
    from regda_b200.utils.eval import confusion_matrix, miou_from_confusion
    g = torch.Generator().manual_seed(3)
    pred = torch.randint(0, 6, (4, 64, 64), generator=g).cuda()
    label = torch.randint(-1, 6, (4, 64, 64), generator=g).cuda()
    cm = confusion_matrix(pred, label, 6, -1)
    assert int(cm.sum()) == int((label != -1).sum())
    iou, miou = miou_from_confusion(cm, skip_class0=True)
    for c in range(1, 6):
        inter = int(((pred == c) & (label == c)).sum())
        union = int((((pred == c) & (label != -1)) | (label == c)).sum())
        assert abs(float(iou[c - 1]) - inter / union) < 1e-12
    assert abs(miou - float(iou.mean())) < 1e-12


def test_evaluate_matches_reference_metric_on_reference_predictions():
    """evaluate() end to end against the reference: the image and the reference model's sliding-window probabilities come from
    tests/golden/teacher_pass.npz (the reference's own pre_slide on its own Deeplabv2); the expected table is the reference
    metric's arithmetic (tests/test_metrics.py pins the mirror to it) over the arg-max of those probabilities."""
    from conftest import load_golden
    from oracle import step_oracle as so
    from regda_b200.gast.metrics import PixelMetricIgnore
    from regda_b200.models.Encoder import Deeplabv2
    from regda_b200.utils.eval import evaluate
    z = load_golden("teacher_pass.npz")
    cfg = dict(backbone=dict(resnet_type="resnet50", output_stride=16, pretrained=False), multi_layer=True, cascade=False, use_ppm=True,
               ppm=dict(num_classes=6, use_aux=False, fc_dim=2048), inchannels=2048, num_classes=6, is_ins_norm=True)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    m = Deeplabv2(cfg, compute_dtype=torch.float32)
    m.load_state_dict(so.seeded_state_dict(m, 2333), strict=True)
    m = m.cuda().train()                           # evaluate() must switch to eval mode and restore train mode
    image = torch.from_numpy(z["image"])
    g = torch.Generator().manual_seed(5)
    ref_pred = torch.from_numpy(z["slide_plain"]).argmax(dim=1)
    label = torch.where(torch.rand(ref_pred.shape, generator=g) < 0.6, ref_pred, torch.randint(-1, 6, ref_pred.shape, generator=g))
    want = PixelMetricIgnore(6, ignore_labels=[0])
    want.forward(label[label >= 0], ref_pred[label >= 0])
    _, want_miou = want.summary_all()
    tb, miou = evaluate(m, [(image, label)], 6, ignore_label=-1, skip_class0=True, tile=64)
    assert m.training
    assert len(tb.iou_per_class) == 5
    # float32 compute: at most a handful of arg-max ties may resolve differently
    assert abs(miou - float(want_miou)) <= 2e-3, (miou, want_miou)
    np_rows = [r[2:] for r in tb.rows[:-1]]
    assert all(0.0 <= float(v) <= 1.0 for r in np_rows for v in r)


def test_evaluate_sliding_window_geometry():
    """one axis shorter than the tile, the other longer (negative window origin): every pixel is predicted
    exactly as the windows cover it, no duplicated window"""
    from regda_b200.utils.eval import _origins, slide_predict
    assert _origins(192, 512, 256) == [0] and _origins(600, 512, 256) == [0, 88] and _origins(1024, 512, 256) == [0, 256, 512]
    assert _origins(112, 64, 32) == [0, 32, 48] and _origins(96, 64, 32) == [0, 32]         # the reference's windows (tools.py:66-79)
    calls = []

    def fake_model(x):
        calls.append(tuple(x.shape[-2:]))
        return torch.ones((x.shape[0], 2) + tuple(x.shape[-2:]), device=x.device)

    out = slide_predict(fake_model, torch.zeros(1, 3, 192, 600, device="cuda"), 2, tile=512)
    assert calls == [(192, 512), (192, 512)]
    assert torch.equal(out, torch.ones_like(out))


def test_weight_shadow_follows_parameter_writes():
    """the bf16 copy the tcgen05 convolutions read must follow load_state_dict() and EMA apply_shadow()/restore().
    Observed at the output of layer1 in eval mode: convolutions + running-statistics BatchNorm only, so two forwards with the
    same weights agree bit for bit (further down, InstanceNorm's fp32 atomics make repeated forwards differ in the last bits)."""
    from regda_b200.models.Encoder import Deeplabv2
    from regda_b200.trainer import ParamArena
    from regda_b200.utils.ema import ExponentialMovingAverage
    cfg = dict(backbone=dict(resnet_type="resnet50", output_stride=16, pretrained=False), multi_layer=True, cascade=False, use_ppm=True,
               ppm=dict(num_classes=6, use_aux=False, fc_dim=2048), inchannels=2048, num_classes=6, is_ins_norm=True)
    torch.manual_seed(0)
    m = Deeplabv2(cfg).cuda().eval()
    arena = ParamArena(m)
    x = torch.randn(2, 3, 256, 256, device="cuda")
    seen = {}
    m.encoder.resnet.layer1.register_forward_hook(lambda mod, i, o: seen.__setitem__("l1", o.detach().float().clone()))

    def l1():
        m(x)
        return seen["l1"]

    with torch.no_grad():
        p0 = l1()
        assert torch.equal(l1(), p0)
        sd = {k: v.clone() for k, v in m.state_dict().items()}
        sd2 = {k: (v * 1.5 if k.endswith("conv2.weight") else v) for k, v in sd.items()}
        m.load_state_dict(sd2)
        p1 = l1()
        assert float((p1 - p0).abs().max()) > 1e-2 * float(p0.abs().max()), "forward ignored load_state_dict (stale bf16 weights)"
        m.load_state_dict(sd)
        assert torch.equal(l1(), p0)
        ema = ExponentialMovingAverage(m, 0.5)
        ema.register()
        for n, p in m.named_parameters():
            if n.endswith("conv2.weight"):
                p.mul_(1.5)
        arena.sync_shadow()
        p2 = l1()
        assert not torch.equal(p2, p0)
        ema.apply_shadow()                      # back to the registered (original) weights
        assert torch.equal(l1(), p0)
        ema.restore()
        assert torch.equal(l1(), p2)


def test_trainer_and_prototype_tools_end_to_end(tmp_path):
    env = dict(os.environ, PYTHONPATH=ROOT)
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "init_prototypes.py"), "--config-path", "st.regda.tiny", "--batches", "2",
                        "--out", str(tmp_path / "proto.pth")], capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    proto = torch.load(tmp_path / "proto.pth")
    assert proto.shape == (6, 2048) and torch.isfinite(proto).all()
    for graph in ("0", "1"):
        r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "train_ssl_reg.py"), "--config-path", "st.regda.tiny", "--ckpt-proto",
                            str(tmp_path / "proto.pth"), "--sam-refine", "--percent", "0.5", "--cuda-graph", graph] +
                           (["--gene-every", "2"] if graph == "0" else []),
                           capture_output=True, text=True, timeout=900, env=env, cwd=ROOT)
        assert r.returncode == 0, r.stderr[-2000:]
        assert "iter=1, total=" in r.stdout and "images/s" in r.stdout, r.stdout[-1000:]
        if graph == "0":                       # the GENE_EVERY teacher pass ran and its .pt soft labels are in the reference's format
            assert "soft pseudo labels" in r.stdout, r.stdout[-1000:]
            soft = torch.load("/tmp/regda_tiny/pseudo_label/synthetic_r0_0000.tif.pt")
            assert soft.dtype == torch.float32 and soft.dim() == 3 and soft.shape[0] == 6
    sd = torch.load("/tmp/regda_tiny/Potsdam_curr.pth")
    assert "encoder.resnet.conv1.weight" in sd and "layer6.conv_last.4.bias" in sd


def test_stage2_align_tool_end_to_end():
    env = dict(os.environ, PYTHONPATH=ROOT)
    for graph in ("0", "1"):
        r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "train_align_reg.py"), "--config-path", "st.regda.tiny", "--steps", "4",
                            "--sam-refine", "--percent", "0.5", "--cuda-graph", graph], capture_output=True, text=True, timeout=900, env=env, cwd=ROOT)
        assert r.returncode == 0, r.stderr[-2000:]
        assert "iter=1, total=" in r.stdout and "loss_align=" in r.stdout and "images/s" in r.stdout, r.stdout[-1000:]
