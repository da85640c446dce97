"""CPU: the oracle (C restatement + torch fp32 restatement) against the committed golden
vectors, which were produced by the unmodified reference (tests/golden/make_golden.py)."""
import numpy as np
import pytest
import torch

from conftest import load_golden, lrh_golden_cases
from oracle import cbind, step_oracle as so

CASES = lrh_golden_cases()


@pytest.mark.parametrize("case", CASES, ids=[str(c["name"]) for c in CASES])
def test_lrh_oracle_bit_exact(case):
    out = cbind.lrh(case["labels"], case["regions"], int(case["class_num"]), int(case["ignore"]), float(case["percent"]))
    assert np.array_equal(out, case["out"])


def test_lrh_float32_boundary_is_what_the_survey_says():
    """exact-percent regions: rejected below 256 valid px, accepted from 256 (float32 n+1e-5)."""
    seen = {}
    for c in CASES:
        name = str(c["name"])
        if name.startswith("exact_p0.5_n"):
            n = int(name.split("_n")[1])
            homogenised = bool((c["out"][0, 0, :n] == 2).all())
            seen[n] = homogenised
    assert seen == {4: False, 100: False, 254: False, 256: True, 512: True, 1024: True, 4096: True}


def test_lrh_oracle_errors():
    lab = np.zeros((1, 2, 2), dtype=np.int64)
    reg = np.zeros((1, 2, 2), dtype=np.int64)
    lab[0, 0, 0] = 9
    with pytest.raises(cbind.OracleError):
        cbind.lrh(lab, reg, 6, -1, 0.5)
    lab[0, 0, 0] = 0
    reg[0, 1, 1] = -3
    with pytest.raises(cbind.OracleError):
        cbind.lrh(lab, reg, 6, -1, 0.5)


def test_pseudo_select_oracle_bit_exact():
    z = load_golden("select_downscale.npz")
    for i in range(4):
        top, low = z[f"sel_{i}_args"]
        assert np.array_equal(cbind.pseudo_select(z["soft"], top, low, -1), z[f"sel_{i}"])
    assert np.array_equal(cbind.pseudo_select(z["soft"], 0.8, 0.6, 255), z["sel_ign255"])


def test_downscale_oracle_bit_exact():
    z = load_golden("select_downscale.npz")
    for i in range(4):
        scale, mr = z[f"ds_{i}_args"]
        assert np.array_equal(cbind.downscale_label(z["ds_label"], int(scale), 6, -1, mr), z[f"ds_{i}"])


def test_aligner_oracle():
    z = load_golden("aligner_loss.npz")
    t = {k: torch.from_numpy(z[k]) for k in z.files if z[k].dtype.kind in "fi"}
    K = t["feat"].shape[1]
    pd = so.pearson_dist(t["feat"].permute(0, 2, 3, 1).reshape(-1, K), t["proto"])
    torch.testing.assert_close(pd, t["pearson"], rtol=1e-5, atol=1e-6)
    for key, temp in (("refined_T2", 2.0), ("refined_T1", 1.0)):
        r = so.label_refine(t["feat"], [t["p1"], t["p2"]], t["soft"], t["proto"], temp)
        torch.testing.assert_close(r, t[key], rtol=1e-4, atol=1e-7)
    r = so.label_refine(t["feat"], [t["p1"]], t["soft"], t["proto"], 2.0)
    torch.testing.assert_close(r, t["refined_single_pred"], rtol=1e-4, atol=1e-7)
    new, ds = so.update_prototype(t["proto"], t["feat"], t["label_s"], 6, -1, 0.996)
    assert torch.equal(ds, t["label_ds"])
    torch.testing.assert_close(new, t["proto_after"], rtol=1e-5, atol=1e-6)


def test_loss_oracle():
    z = load_golden("aligner_loss.npz")
    p1 = torch.from_numpy(z["p1"]).requires_grad_(True)
    p2 = torch.from_numpy(z["p2"]).requires_grad_(True)
    lab = torch.from_numpy(z["label_s"])
    loss = so.ce_loss_multi([p1, p2], lab, -1)
    loss.backward()
    torch.testing.assert_close(loss.detach(), torch.from_numpy(z["loss"]), rtol=1e-6, atol=0)
    torch.testing.assert_close(p1.grad, torch.from_numpy(z["dp1"]), rtol=1e-5, atol=1e-8)
    torch.testing.assert_close(p2.grad, torch.from_numpy(z["dp2"]), rtol=1e-5, atol=1e-8)
    l0 = so.ce_loss_multi([p1.detach(), p2.detach()], torch.full_like(lab, -1), -1)
    assert float(l0) == float(z["loss_all_ignored"]) == 0.0


def test_class_balance_oracle():
    z = load_golden("aligner_loss.npz")
    lab = torch.from_numpy(z["label_s"])
    freq = so.class_balance_freq(torch.ones(6) / 6, lab, 6, -1, 0.99)
    torch.testing.assert_close(freq, torch.from_numpy(z["cb_freq"]), rtol=1e-6, atol=1e-8)
    torch.testing.assert_close(so.class_balance_weights(freq, 2.0), torch.from_numpy(z["cb_class_weight"]), rtol=1e-6, atol=1e-8)


@pytest.mark.parametrize("rt", ["resnet50", "resnet101"])
def test_model_oracle_matches_reference_fixture(rt):
    z = load_golden(f"model_{rt}.npz")
    torch.manual_seed(0)
    m = so.DeeplabOracle(rt, 6, dropout=0.0)
    m.load_state_dict(so.seeded_state_dict(m, 2333))
    x = torch.from_numpy(z["x"])
    m.eval()
    with torch.no_grad():
        torch.testing.assert_close(m(x), torch.from_numpy(z["eval_prob"]), rtol=1e-4, atol=1e-5)
    m.train()
    x1, x2, feat = m(x)
    torch.testing.assert_close(x1, torch.from_numpy(z["x1"]), rtol=1e-4, atol=1e-4)
    torch.testing.assert_close(feat, torch.from_numpy(z["feat"]), rtol=1e-4, atol=1e-4)
    loss = so.ce_loss_multi([x1, x2], torch.from_numpy(z["label"]), -1)
    loss.backward()
    assert abs(float(loss) - float(z["loss"])) <= 1e-5 * abs(float(z["loss"]))
    names = [str(n) for n in z["grad_names"]]
    assert names == [n for n, _ in m.named_parameters()]
    got = np.array([float(p.grad.norm()) for _, p in m.named_parameters()])
    np.testing.assert_allclose(got, z["grad_norms"], rtol=2e-3, atol=1e-6)


def test_full_step_oracle_matches_reference_fixture():
    z = load_golden("step_resnet50.npz")
    m = so.DeeplabOracle("resnet50", 6, dropout=0.0)
    m.load_state_dict(so.seeded_state_dict(m, 2333))
    st = so.StepState(m, torch.from_numpy(z["proto"]).clone())
    t = lambda k: torch.from_numpy(z[k])  # noqa: E731
    for it in range(2):
        r = so.inner_step(st, t("xs"), t("ls"), t("xt"), t("soft"), t("regs"))
        np.testing.assert_allclose([r["loss"], r["loss_source"], r["loss_target"], r["grad_norm"]], z["losses"][it], rtol=2e-4)
        if it == 0:
            assert (r["hard"].numpy() != z["hard_0"]).mean() < 1e-3   # fp-threshold flips only
    torch.testing.assert_close(st.prototypes, t("proto_after"), rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(m.encoder.resnet.conv1.weight.detach(), t("conv1_after"), rtol=1e-3, atol=1e-5)


@pytest.mark.parametrize("case", CASES, ids=[str(c["name"]) for c in CASES])
def test_lrh_torch_port_bit_exact(case):
    """the multi-threaded torch port timed by `bench.py --impl reference` is held to the same pins"""
    out = so.lrh_torch(torch.from_numpy(case["labels"]), torch.from_numpy(case["regions"]),
                       int(case["class_num"]), int(case["ignore"]), float(case["percent"]))
    assert np.array_equal(out.numpy(), case["out"])


def _teacher_model():
    m = so.DeeplabOracle("resnet50", 6, dropout=0.0)
    m.load_state_dict(so.seeded_state_dict(m, 2333))
    return m.eval()


def test_teacher_pass_oracle_matches_reference_fixture():
    """offline teacher pass: the oracle's pre_slide / tta_predict against what the reference's own
    functions produced (tests/golden/teacher_pass.npz)"""
    z = load_golden("teacher_pass.npz")
    m = _teacher_model()
    x, xs = torch.from_numpy(z["image"]), torch.from_numpy(z["image_small"])
    with torch.no_grad():
        torch.testing.assert_close(so.tta_predict(m, x[:, :, :64, :64]), torch.from_numpy(z["tile_tta"]), rtol=1e-4, atol=1e-5)
        torch.testing.assert_close(so.pre_slide(m, x, 6, (64, 64), tta=False), torch.from_numpy(z["slide_plain"]), rtol=1e-4, atol=1e-5)
        torch.testing.assert_close(so.pre_slide(m, x, 6, (64, 64), tta=True), torch.from_numpy(z["slide_tta"]), rtol=1e-4, atol=1e-5)
        torch.testing.assert_close(so.pre_slide(m, xs, 6, (64, 64), tta=True), torch.from_numpy(z["slide_small_tta"]), rtol=1e-4, atol=1e-5)


def test_teacher_pass_host_logic_matches_reference_fixture():
    """the product's host-side pre_slide / tta_predict (device-agnostic window arithmetic, TTA view order, the stacked-views
    batch, the pad quirk) around the oracle model on the CPU"""
    from regda_b200.utils.tools import pre_slide, tta_predict
    z = load_golden("teacher_pass.npz")
    m = _teacher_model()
    x, xs = torch.from_numpy(z["image"]), torch.from_numpy(z["image_small"])
    with torch.no_grad():
        torch.testing.assert_close(tta_predict(m, x[:, :, :64, :64]), torch.from_numpy(z["tile_tta"]), rtol=1e-4, atol=1e-5)
        torch.testing.assert_close(pre_slide(m, x, num_classes=6, tile_size=(64, 64), tta=True), torch.from_numpy(z["slide_tta"]), rtol=1e-4, atol=1e-5)
        torch.testing.assert_close(pre_slide(m, x, num_classes=6, tile_size=(64, 64), tta=False), torch.from_numpy(z["slide_plain"]), rtol=1e-4, atol=1e-5)
        torch.testing.assert_close(pre_slide(m, xs, num_classes=6, tile_size=(64, 64), tta=True), torch.from_numpy(z["slide_small_tta"]), rtol=1e-4, atol=1e-5)


def test_pcl_oracle_matches_reference_fixture():
    z = load_golden("align_step_resnet50.npz")
    feat = torch.from_numpy(z["pcl_feat"]).requires_grad_(True)
    loss = so.pcl_loss(torch.from_numpy(z["pcl_proto"]), feat, torch.from_numpy(z["pcl_label"]), 8.0, -1)
    (loss * 0.5).backward()
    assert abs(float(loss) - float(z["pcl_loss"])) <= 1e-6 * abs(float(z["pcl_loss"]))
    torch.testing.assert_close(feat.grad, torch.from_numpy(z["pcl_dfeat_half"]), rtol=1e-5, atol=1e-8)


def test_align_step_oracle_matches_reference_fixture():
    """stage-2 step: the oracle's restatement of tools/train_align_reg.py:144-196 against two iterations
    of the reference's own objects"""
    z = load_golden("align_step_resnet50.npz")
    m = so.DeeplabOracle("resnet50", 6, dropout=0.0)
    m.load_state_dict(so.seeded_state_dict(m, 2333))
    st = so.StepState(m, torch.from_numpy(z["proto"]).clone())
    t = lambda k: torch.from_numpy(z[k])  # noqa: E731
    for it in range(2):
        r = so.align_step(st, t("xs"), t("ls"), t("xt"), t("regs"))
        np.testing.assert_allclose([r["loss"], r["loss_seg"], r["loss_align"], r["grad_norm"]], z["losses"][it], rtol=2e-4)
        if it == 0:
            assert (r["hard"].numpy() != z["hard_0"]).mean() < 1e-3   # fp-threshold flips only
            assert (r["label_t"].numpy() != z["label_t_0"]).mean() < 1e-2
    torch.testing.assert_close(st.prototypes, t("proto_after"), rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(m.encoder.resnet.conv1.weight.detach(), t("conv1_after"), rtol=1e-3, atol=1e-5)


def test_bf16_emulation_without_rounding_is_the_reference():
    """oracle/bf16_emul.py restates the model forward with a bf16 cast at every point where the CUDA kernels round; with the
    casts switched off it must reproduce the reference's own float32 outputs (tests/golden/model_resnet50.npz) -- that pins
    the restatement, so that the GPU comparison against the ROUNDING version (tests/test_bf16_parity_gpu.py) is a comparison
    with the reference's architecture and weights."""
    from oracle import bf16_emul as be
    z = load_golden("model_resnet50.npz")
    m = so.DeeplabOracle("resnet50", 6, dropout=0.0)
    m.load_state_dict(so.seeded_state_dict(m, 2333))
    m.train()
    be.ROUND = False
    try:
        x1, x2, feat = be.forward_train(m, torch.from_numpy(z["x"]))
    finally:
        be.ROUND = True
    for got, key in ((x1, "x1"), (x2, "x2"), (feat, "feat")):
        ref = torch.from_numpy(z[key])
        assert float((got - ref).abs().max()) <= 2e-4 * float(ref.abs().max()), key
    # and with rounding it is a different function (bf16 arithmetic drifts by percents on these random weights)
    y1, _, _ = be.forward_train(m, torch.from_numpy(z["x"]))
    assert float((y1 - torch.from_numpy(z["x1"])).abs().max()) > 1e-3 * float(np.abs(z["x1"]).max())


def test_coral_oracle_matches_reference():
    """so.coral_loss (restating regda/gast/coral.py:26-47) against the reference's own CoralLoss outputs and gradients"""
    z = load_golden("coral.npz")
    for k in range(3):
        src = torch.from_numpy(z[f"loss{k}/src"]).requires_grad_(True)
        tgt = torch.from_numpy(z[f"loss{k}/tgt"]).requires_grad_(True)
        loss = so.coral_loss(src, tgt, bool(z[f"loss{k}/is_sqrt"]))
        loss.backward()
        assert abs(float(loss) - float(z[f"loss{k}/loss"])) <= 1e-6 * abs(float(z[f"loss{k}/loss"]))
        np.testing.assert_allclose(src.grad.numpy(), z[f"loss{k}/dsrc"], rtol=1e-4, atol=1e-9)
        np.testing.assert_allclose(tgt.grad.numpy(), z[f"loss{k}/dtgt"], rtol=1e-4, atol=1e-9)
