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


def test_evaluate_runs_sliding_window():
    from regda_b200.models.Encoder import Deeplabv2
    from regda_b200.utils.eval import evaluate
    cfg = dict(backbone=dict(resnet_type="resnet50", output_stride=16, pretrained=False), multi_layer=True, cascade=False, use_ppm=True,
               ppm=dict(num_classes=6, use_aux=False, fc_dim=2048), inchannels=2048, num_classes=6, is_ins_norm=True)
    torch.manual_seed(0)
    m = Deeplabv2(cfg).cuda()
    data = [(torch.randn(1, 3, 192, 160), torch.randint(-1, 6, (1, 192, 160)))]
    iou, miou = evaluate(m, data, 6, ignore_label=-1, skip_class0=True, tile=128)
    assert iou.shape == (5,) and 0.0 <= miou <= 1.0 and not m.training is None


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
