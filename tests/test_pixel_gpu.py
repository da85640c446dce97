"""GPU: the full-resolution pixel chain and the prototype / loss / optimiser kernels against the
golden vectors of the unmodified reference (tests/golden/*.npz) and the CPU oracle.
Integer outputs bit-exact; float outputs within the tolerance written next to each check
(north_star: 1e-3 relative on fp logits/loss; most checks are far tighter)."""
import numpy as np
import pytest
import torch

from conftest import load_golden
from oracle import cbind, step_oracle as so

pytestmark = pytest.mark.gpu


def _t(z, k):
    return torch.from_numpy(z[k]).cuda()


# ---- pseudo_selection -------------------------------------------------------------------------
def test_pseudo_selection_golden_bit_exact():
    from regda_b200.gast.pseudo_generation import pseudo_selection
    z = load_golden("select_downscale.npz")
    soft = _t(z, "soft")
    for i in range(4):
        top, low = z[f"sel_{i}_args"]
        out = pseudo_selection(soft, top, low, "tensor", -1)
        assert out.dtype == torch.int64
        assert np.array_equal(out.cpu().numpy(), z[f"sel_{i}"])
    assert np.array_equal(pseudo_selection(soft, 0.8, 0.6, "ndarray", 255), z["sel_ign255"])


@pytest.mark.parametrize("shape", [(8, 6, 512, 512), (2, 7, 100, 37), (1, 16, 64, 64), (3, 1, 9, 9)])
def test_pseudo_selection_random_vs_oracle(shape):
    from regda_b200.gast.pseudo_generation import pseudo_selection
    g = torch.Generator(device="cuda").manual_seed(5)
    soft = torch.softmax(4 * torch.randn(shape, device="cuda", generator=g), 1)
    out = pseudo_selection(soft, 0.8, 0.6, "tensor", -1)
    want = cbind.pseudo_select(soft.cpu().numpy(), 0.8, 0.6, -1)
    assert np.array_equal(out.cpu().numpy(), want)


def test_pseudo_selection_asserts_on_out_of_range_probabilities():
    from regda_b200.gast.pseudo_generation import pseudo_selection
    soft = torch.full((1, 3, 4, 4), 0.3, device="cuda")
    soft[0, 1, 2, 2] = 1.5
    with pytest.raises(AssertionError):
        pseudo_selection(soft, 0.8, 0.6, "tensor", -1)
    soft[0, 1, 2, 2] = -0.1
    with pytest.raises(AssertionError):
        pseudo_selection(soft, 0.8, 0.6, "tensor", -1)


# ---- DownscaleLabel ---------------------------------------------------------------------------
def test_downscale_golden_bit_exact():
    from regda_b200.gast.alignment import DownscaleLabel
    z = load_golden("select_downscale.npz")
    lab = _t(z, "ds_label")
    for i in range(4):
        scale, mr = z[f"ds_{i}_args"]
        out = DownscaleLabel(scale_factor=int(scale), n_classes=6, ignore_label=-1, min_ratio=float(mr))(lab)
        assert np.array_equal(out.cpu().numpy(), z[f"ds_{i}"])


def test_downscale_random_vs_oracle():
    from regda_b200 import synth
    from regda_b200.gast.alignment import DownscaleLabel
    _, label_s, *_ = synth.step_inputs(4, 512, 512, device="cuda")
    out = DownscaleLabel(16, 6, -1, 0.75)(label_s)
    want = cbind.downscale_label(label_s.cpu().numpy(), 16, 6, -1, 0.75)
    assert np.array_equal(out.cpu().numpy(), want)
    out = DownscaleLabel(16, 6, -1, 0.75)(label_s.unsqueeze(1))
    assert np.array_equal(out.cpu().numpy(), want)


# ---- Aligner ----------------------------------------------------------------------------------
def _aligner(K=64, C=6, proto=None):
    from regda_b200.gast.alignment import Aligner
    al = Aligner(None, K, C, -1, 0.996)
    if proto is not None:
        al.prototypes = proto.clone()
    return al


def test_pearson_and_label_refine_golden():
    z = load_golden("aligner_loss.npz")
    al = _aligner(proto=_t(z, "proto"))
    feat = _t(z, "feat")
    K = feat.shape[1]
    pd = al._pearson_dist(feat.permute(0, 2, 3, 1).reshape(-1, K), al.prototypes)
    torch.testing.assert_close(pd.cpu(), torch.from_numpy(z["pearson"]), rtol=1e-5, atol=1e-6)
    soft, p1, p2 = _t(z, "soft"), _t(z, "p1"), _t(z, "p2")
    for key, temp in (("refined_T2", 2.0), ("refined_T1", 1.0)):
        r = al.label_refine(None, feat, [p1, p2], soft, True, "all", temp)
        torch.testing.assert_close(r.cpu(), torch.from_numpy(z[key]), rtol=1e-4, atol=1e-7)
    r = al.label_refine(None, feat, p1, soft, True, "all", 2.0)
    torch.testing.assert_close(r.cpu(), torch.from_numpy(z["refined_single_pred"]), rtol=1e-4, atol=1e-7)
    assert al.label_refine(None, feat, [p1, p2], soft, False) is soft
    with pytest.raises(NotImplementedError):
        al.label_refine(torch.zeros(1), feat, [p1, p2], soft)


def test_refine_select_fused_matches_refine_then_select():
    """the fused kernel pair must agree with the oracle's label_refine -> pseudo_select except at
    pixels whose refined probability sits within 1e-5 of its threshold (float re-association)."""
    from regda_b200 import synth
    b, C, K, h, w, H, W = 2, 6, 256, 16, 16, 256, 256
    g = torch.Generator(device="cuda").manual_seed(3)
    feat = torch.randn(b, K, h, w, device="cuda", generator=g)
    p1 = 3 * torch.randn(b, C, h, w, device="cuda", generator=g)
    p2 = 3 * torch.randn(b, C, h, w, device="cuda", generator=g)
    soft = torch.softmax(3.0 * synth.smooth_noise((b, C, H, W), "cuda", g), dim=1)
    proto = torch.randn(C, K, device="cuda", generator=g).abs()
    al = _aligner(K, C, proto)
    hard = al.refine_select(feat, [p1, p2], soft, 2.0, 0.8, 0.6)
    refined = al.label_refine(None, feat, [p1, p2], soft, True, "all", 2.0)
    ref_soft = so.label_refine(feat.cpu(), [p1.cpu(), p2.cpu()], soft.cpu(), proto.cpu(), 2.0)
    torch.testing.assert_close(refined.cpu(), ref_soft, rtol=1e-4, atol=1e-6)
    # fused == (our refine -> exact select): same kernel arithmetic, must be identical
    from regda_b200.gast.pseudo_generation import pseudo_selection
    assert torch.equal(hard, pseudo_selection(refined, 0.8, 0.6, "tensor", -1))
    want = torch.from_numpy(cbind.pseudo_select(ref_soft.numpy(), 0.8, 0.6, -1))
    diff = hard.cpu() != want
    if diff.any():
        thr = torch.clamp(ref_soft.flatten(2).max(-1)[0] * 0.8, min=0.6).view(b, C, 1, 1)
        margin = (ref_soft - thr).abs().min(1)[0]
        assert float(margin[diff].max()) < 1e-5
    assert diff.float().mean() < 1e-3
    assert (hard >= 0).any() and (hard == -1).any()


def test_update_prototype_and_init_avg_golden():
    z = load_golden("aligner_loss.npz")
    al = _aligner(proto=_t(z, "proto"))
    feat, lab = _t(z, "feat"), _t(z, "label_s")
    ds = al.update_prototype(feat, lab)
    assert np.array_equal(ds.cpu().numpy(), z["label_ds"])
    torch.testing.assert_close(al.prototypes.cpu(), torch.from_numpy(z["proto_after"]), rtol=1e-5, atol=1e-6)
    al2 = _aligner()
    al2.update_avg(feat, lab)
    al2.update_avg(feat * 0.5 + 1.0, lab)
    al2.init_avg()
    torch.testing.assert_close(al2.prototypes.cpu(), torch.from_numpy(z["proto_init_avg"]), rtol=1e-5, atol=1e-6)


def test_update_prototype_full_size_vs_oracle():
    from regda_b200 import synth
    b, C, K = 8, 6, 2048
    _, label_s, _, _, _, proto = synth.step_inputs(b, 512, 512, device="cuda")
    g = torch.Generator(device="cuda").manual_seed(9)
    feat = torch.randn(b, K, 32, 32, device="cuda", generator=g).contiguous(memory_format=torch.channels_last)
    al = _aligner(K, C, proto)
    al.update_prototype(feat, label_s)
    want, _ = so.update_prototype(proto.cpu(), feat.cpu(), label_s.cpu(), C, -1, 0.996)
    torch.testing.assert_close(al.prototypes.cpu(), want, rtol=1e-4, atol=1e-5)


# ---- loss -------------------------------------------------------------------------------------
def test_ce_loss_and_gradients_golden():
    from regda_b200.gast.balance import CrossEntropy
    from regda_b200.utils.tools import loss_calc
    z = load_golden("aligner_loss.npz")
    p1 = _t(z, "p1").requires_grad_(True)
    p2 = _t(z, "p2").requires_grad_(True)
    lab = _t(z, "label_s")
    ce = CrossEntropy(ignore_label=-1, class_balancer=None)
    loss = loss_calc([p1, p2], lab, ce, multi=True)
    loss.backward()
    assert abs(float(loss) - float(z["loss"])) <= 1e-5 * abs(float(z["loss"]))
    torch.testing.assert_close(p1.grad.cpu(), torch.from_numpy(z["dp1"]), rtol=1e-4, atol=1e-8)
    torch.testing.assert_close(p2.grad.cpu(), torch.from_numpy(z["dp2"]), rtol=1e-4, atol=1e-8)
    la = loss_calc([p1.detach(), p2.detach()], torch.full_like(lab, -1), ce, multi=True)
    assert float(la) == float(z["loss_all_ignored"]) == 0.0


@pytest.mark.parametrize("shape", [(8, 6, 32, 32, 512, 512), (2, 7, 64, 64, 1024, 1024), (2, 6, 40, 40, 40, 40), (1, 3, 5, 7, 33, 50)])
def test_ce_random_vs_torch(shape):
    from regda_b200.gast.balance import CrossEntropy
    b, c, h, w, H, W = shape
    g = torch.Generator(device="cuda").manual_seed(1)
    p = (2 * torch.randn(b, c, h, w, device="cuda", generator=g)).requires_grad_(True)
    lab = torch.randint(-1, c, (b, H, W), device="cuda", generator=g)
    loss = CrossEntropy(-1)(p, lab)
    loss.backward()
    q = p.detach().cpu().double().requires_grad_(True)
    want = so.ce_loss_multi([q], lab.cpu(), -1)
    want.backward()
    assert abs(float(loss) - float(want)) <= 2e-6 * abs(float(want))
    torch.testing.assert_close(p.grad.cpu().double(), q.grad, rtol=2e-4, atol=1e-9)


def test_class_balance_golden():
    from regda_b200.gast.balance import ClassBalance
    z = load_golden("aligner_loss.npz")
    lab = _t(z, "label_s")
    cb = ClassBalance(class_num=6, ignore_label=-1, decay=0.99, temperature=2.0)
    wpx = cb.get_class_weight_4pixel(lab)
    torch.testing.assert_close(cb.freq.cpu(), torch.from_numpy(z["cb_freq"]), rtol=1e-6, atol=1e-8)
    torch.testing.assert_close(cb._get_class_wight().cpu(), torch.from_numpy(z["cb_class_weight"]), rtol=1e-5, atol=1e-7)
    assert abs(float(wpx.sum()) - float(z["cb_weight_sum"])) <= 1e-4 * abs(float(z["cb_weight_sum"]))
    counts, nv = cbind.class_count(lab.cpu().numpy(), 6, -1)
    assert int((lab != -1).sum()) == nv


# ---- optimiser --------------------------------------------------------------------------------
@pytest.mark.parametrize("n", [1 << 20, 1000003, 7])
def test_clip_and_sgd_match_torch(n):
    from regda_b200 import capi
    g = torch.Generator(device="cuda").manual_seed(2)
    p0 = torch.randn(n, device="cuda", generator=g)
    grads = [torch.randn(n, device="cuda", generator=g) * s for s in (0.5, 0.01, 3.0)]
    ref_p = p0.clone().requires_grad_(True)
    opt = torch.optim.SGD([ref_p], lr=1e-2, momentum=0.9, weight_decay=5e-4)
    p, buf = p0.clone(), torch.zeros(n, device="cuda")
    shadow = torch.zeros(n, dtype=torch.bfloat16, device="cuda")
    sumsq = torch.zeros(1, device="cuda")
    ws = capi.workspace.get(capi.lib().regda_sumsq_workspace_bytes(n), "cuda")
    for it, gr in enumerate(grads):
        ref_p.grad = gr.clone()
        norm = torch.nn.utils.clip_grad_norm_([ref_p], max_norm=32.0, norm_type=2)
        opt.step()
        capi.call("regda_sumsq", capi.ptr(gr), n, capi.ptr(sumsq), 0, capi.ptr(ws), ws.numel(), capi.stream())
        assert abs(float(sumsq.sqrt()) - float(norm)) <= 1e-5 * float(norm)
        capi.call("regda_sgd_step", capi.ptr(p), capi.ptr(gr), capi.ptr(buf), capi.ptr(shadow), n, capi.ptr(sumsq), 32.0, 1.0,
                  1e-2, None, 0.9, 5e-4, int(it == 0), capi.stream())
        torch.testing.assert_close(p, ref_p.detach(), rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(shadow.float(), p.bfloat16().float(), rtol=0, atol=0)


def test_ema_update_matches_reference_formula():
    from regda_b200 import capi
    g = torch.Generator(device="cuda").manual_seed(4)
    sh, p = torch.randn(10007, device="cuda", generator=g), torch.randn(10007, device="cuda", generator=g)
    want = (1.0 - 0.999) * p + 0.999 * sh
    capi.call("regda_ema_update", capi.ptr(sh), capi.ptr(p), 10007, 0.999, capi.stream())
    torch.testing.assert_close(sh, want, rtol=1e-6, atol=1e-7)
